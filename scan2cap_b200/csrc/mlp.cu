// Grouped-MLP kernels for sm_100a: the per-group shared MLP of PointnetSAModuleVotes / PointnetFPModule
// (reference lib/pointnet2/pytorch_utils.py:11-36,67-120 SharedMLP = [1x1 Conv2d(no bias) -> BatchNorm2d -> ReLU]*,
// applied at lib/pointnet2/pointnet2_modules.py:251-257 and followed by max_pool2d over nsample).
//
// The reference runs three library kernels per layer (cuDNN/cuBLAS conv, BatchNorm, ReLU), each streaming the
// whole (B*npoint*nsample) x C activation through HBM.  Here one layer is ONE kernel built on the 5th-generation
// tensor cores:
//   C[R x N] = f(A)[R x K] * W[N x K]^T        R = every (scene, group, sample) row, K = Cin, N = Cout
//     prologue  f(a) = relu(a * scale[k] + shift[k])   -- the PREVIOUS layer's BatchNorm + ReLU, applied while the
//                                                         tile is staged (the normalised activation never exists in HBM)
//     MMA       tcgen05.mma.cta_group::1.kind::tf32, M = 128, accumulators in TMEM; fp32 operands are split
//               a = a_hi + a_lo, w = w_hi + w_lo (tf32 parts) and three MMAs (hi*hi, hi*lo, lo*hi) are accumulated
//               -- "3xTF32": fp32-faithful results (needed for the 1e-3 parity bar) at tensor-core rate
//     epilogue  tcgen05.ld -> registers -> shared -> coalesced store of the PRE-BatchNorm output, fused with the
//               per-channel sum / sum of squares this layer's BatchNorm needs (batch statistics)
// Operands are staged by the CTA's threads (not TMA) because of the element-wise prologue; they are written in the
// canonical K-major SWIZZLE_128B layout the UMMA shared-memory descriptors expect.
// The kernel is persistent (one CTA per SM loops over 128-row tiles): TMEM is allocated once, and the BatchNorm
// statistics are accumulated per CTA and flushed with one double-precision atomic per channel at the end.
#include "s2c_common.cuh"

namespace s2c {
namespace {

constexpr int BM = 128;        // rows per tile = UMMA M
constexpr int BK = 32;         // fp32 elements per 128-byte swizzle row
constexpr int NTHREADS = 128;  // 4 warps: warp w owns TMEM lanes [32w, 32w+32)
constexpr int STAGES = 2;

struct GemmArgs {
  const float *A;            // (R, lda)
  long long lda;
  int K;                     // valid columns of A
  const float *pro_scale;    // [K] or null
  const float *pro_shift;    // [K] or null
  const float *W;            // (N, K) row-major
  int N;                     // multiple of 16, <= 256
  float *C;                  // (R, ldc)
  long long ldc;
  double *stat_sum;          // [N] or null
  double *stat_sumsq;        // [N] or null
  long long R;
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]   (kind::tf32, single CTA)
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every tcgen05 op issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t *v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (1, unused for swizzled K-major) |
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B) | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=TF32 [7,10)=2, b=TF32 [10,13)=2,
// K-major A and B (bits 15,16 = 0), n_dim = N>>3 at [17,23), m_dim = M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of the 16-byte chunk `seg` (4 fp32) of row r inside a [rows x 32 fp32] K-major SWIZZLE_128B block
__device__ __forceinline__ uint32_t swz(int r, int seg) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((seg ^ (r & 7)) << 4));
}

__device__ __forceinline__ void split_tf32(float v, float &hi, float &lo) {
  hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);  // the 19 bits the tensor core reads
  lo = v - hi;                                             // exact in fp32
}

__device__ __forceinline__ void store_split(unsigned char *hi_base, unsigned char *lo_base, uint32_t off, float4 v) {
  float4 h, l;
  split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
  *reinterpret_cast<float4 *>(hi_base + off) = h;
  *reinterpret_cast<float4 *>(lo_base + off) = l;
}

// ---------------------------------------------------------------------------------------------- the GEMM kernel
__global__ void __launch_bounds__(NTHREADS, 1)
mlp_gemm_kernel(GemmArgs g) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // layout: per stage [A_hi 16K | A_lo 16K | B_hi N*128 | B_lo N*128], then the epilogue staging, then barriers
  const int N = g.N;
  const uint32_t a_bytes = BM * BK * 4;         // 16 KB
  const uint32_t b_bytes = (uint32_t)N * BK * 4;  // N rows x 128 B (N multiple of 16 -> multiple of 2 KB)
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  unsigned char *stage_base = smem;
  float *sC = reinterpret_cast<float *>(smem + STAGES * stage_bytes);           // [4 warps][32 rows][33]
  float *s_sum = sC + 4 * 32 * 33;                                              // [4 warps][N] per-warp partial sums
  float *s_sq = s_sum + 4 * N;                                                  // [4 warps][N]
  float *s_scale = s_sq + 4 * N;                                                // [Kpad] prologue scale
  float *s_shift = s_scale + ((g.K + BK - 1) / BK) * BK;                        // [Kpad] prologue shift
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_shift + ((g.K + BK - 1) / BK) * BK);  // [STAGES] empty + [1] done
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + STAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KC = (g.K + BK - 1) / BK;
  const bool has_pro = g.pro_scale != nullptr;
  const bool has_stats = g.stat_sum != nullptr;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < N) tmem_cols <<= 1;

  if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
  if (tid == 0) {
    for (int s = 0; s < STAGES + 1; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int k = tid; k < KC * BK; k += NTHREADS) {
    s_scale[k] = (has_pro && k < g.K) ? g.pro_scale[k] : 1.f;
    s_shift[k] = (has_pro && k < g.K) ? g.pro_shift[k] : 0.f;
  }
  for (int i = tid; i < 8 * N; i += NTHREADS) s_sum[i] = 0.f;  // s_sum and s_sq are contiguous
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = make_idesc(BM, N);
  const bool a_vec = ((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
  const bool w_vec = ((g.K & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.W) & 15) == 0);

  const long long num_tiles = (g.R + BM - 1) / BM;
  uint32_t it = 0;         // global k-chunk counter (selects the stage and its barrier phase)
  uint32_t tile_count = 0;
  for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_count) {
    const long long row0 = tile * BM;
    for (int kc = 0; kc < KC; ++kc, ++it) {
      const int s = it % STAGES;
      unsigned char *a_hi = stage_base + (size_t)s * stage_bytes;
      unsigned char *a_lo = a_hi + a_bytes;
      unsigned char *b_hi = a_lo + a_bytes;
      unsigned char *b_lo = b_hi + b_bytes;
      if (it >= STAGES) mbar_wait(&bars[s], ((it / STAGES) - 1) & 1);  // the MMAs that read this stage are done
      const int k0 = kc * BK;
      // ---- A chunk: 128 rows x 32 k.  thread -> (row = pass*16 + tid/8, 16-byte segment = tid%8)
      {
        const int seg = tid & 7;
        const int kk = k0 + seg * 4;
        const float4 sc = *reinterpret_cast<const float4 *>(s_scale + kk);
        const float4 sh = *reinterpret_cast<const float4 *>(s_shift + kk);
#pragma unroll
        for (int pass = 0; pass < BM / 16; ++pass) {
          const int r = pass * 16 + (tid >> 3);
          const long long row = row0 + r;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row < g.R) {
            const float *p = g.A + row * g.lda + kk;
            if (a_vec && kk + 3 < g.K) {
              v = __ldg(reinterpret_cast<const float4 *>(p));
            } else {
              if (kk + 0 < g.K) v.x = __ldg(p + 0);
              if (kk + 1 < g.K) v.y = __ldg(p + 1);
              if (kk + 2 < g.K) v.z = __ldg(p + 2);
              if (kk + 3 < g.K) v.w = __ldg(p + 3);
            }
            if (has_pro) {
              v.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f); v.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
              v.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f); v.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
              if (kk + 0 >= g.K) v.x = 0.f;
              if (kk + 1 >= g.K) v.y = 0.f;
              if (kk + 2 >= g.K) v.z = 0.f;
              if (kk + 3 >= g.K) v.w = 0.f;
            }
          }
          store_split(a_hi, a_lo, swz(r, seg), v);
        }
      }
      // ---- W chunk: N rows x 32 k
      {
        const int seg = tid & 7;
        const int kk = k0 + seg * 4;
        for (int r = tid >> 3; r < N; r += NTHREADS / 8) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          const float *p = g.W + (size_t)r * g.K + kk;
          if (w_vec && kk + 3 < g.K) {
            v = __ldg(reinterpret_cast<const float4 *>(p));
          } else {
            if (kk + 0 < g.K) v.x = __ldg(p + 0);
            if (kk + 1 < g.K) v.y = __ldg(p + 1);
            if (kk + 2 < g.K) v.z = __ldg(p + 2);
            if (kk + 3 < g.K) v.w = __ldg(p + 3);
          }
          store_split(b_hi, b_lo, swz(r, seg), v);
        }
      }
      fence_async_proxy();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
        for (int j = 0; j < BK / 8; ++j) {  // UMMA K = 8 tf32 = 32 bytes inside the 128-byte swizzle row
          const uint32_t o = j * 32;
          umma_tf32(tmem_base, make_desc(ah + o), make_desc(bh + o), idesc, (kc | j) ? 1u : 0u);
          umma_tf32(tmem_base, make_desc(ah + o), make_desc(bl + o), idesc, 1u);
          umma_tf32(tmem_base, make_desc(al + o), make_desc(bh + o), idesc, 1u);
        }
        umma_commit(&bars[s]);                       // frees this stage
        if (kc == KC - 1) umma_commit(&bars[STAGES]);  // accumulator complete
      }
    }
    // ---- epilogue: TMEM -> registers -> shared (per-warp 32x33) -> coalesced global store + column statistics
    mbar_wait(&bars[STAGES], tile_count & 1);
    tc_fence_after();
    float *wC = sC + warp * 32 * 33;
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c) wC[lane * 33 + c] = __uint_as_float(v[c]);
      __syncwarp();
      float sum = 0.f, sq = 0.f;
      const int col = c0 + lane;
#pragma unroll 8
      for (int i = 0; i < 32; ++i) {
        const long long row = row0 + warp * 32 + i;
        const float y = wC[i * 33 + lane];
        if (row < g.R && col < N) {
          g.C[row * g.ldc + col] = y;
          sum += y;
          sq = fmaf(y, y, sq);
        }
      }
      if (has_stats && col < N) {
        s_sum[warp * N + col] += sum;
        s_sq[warp * N + col] += sq;
      }
      __syncwarp();
    }
    tc_fence_before();
    __syncthreads();  // every warp has drained its TMEM lanes before the next tile overwrites the accumulator
  }
  if (has_stats) {
    for (int c = tid; c < N; c += NTHREADS) {
      const double s = (double)s_sum[c] + (double)s_sum[N + c] + (double)s_sum[2 * N + c] + (double)s_sum[3 * N + c];
      const double q = (double)s_sq[c] + (double)s_sq[N + c] + (double)s_sq[2 * N + c] + (double)s_sq[3 * N + c];
      atomicAdd(g.stat_sum + c, s);
      atomicAdd(g.stat_sumsq + c, q);
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

size_t gemm_smem_bytes(int N, int K) {
  const int KC = (K + BK - 1) / BK;
  size_t stage = 2 * (size_t)BM * BK * 4 + 2 * (size_t)N * BK * 4;
  return STAGES * stage + (size_t)(4 * 32 * 33 + 8 * N + 2 * KC * BK) * 4 + (STAGES + 1) * 8 + 16;
}

}  // namespace
}  // namespace s2c

extern "C" int s2c_mlp_layer_fwd(const float *A, long long lda, long long R, int K, const float *pro_scale,
                                 const float *pro_shift, const float *W, int N, float *C, long long ldc,
                                 double *stat_sum, double *stat_sumsq, void *stream) {
  using namespace s2c;
  S2C_REQUIRE(R >= 0 && K >= 1 && N >= 16 && N <= 256 && (N % 16) == 0, "mlp_layer_fwd: need K>=1, N in [16,256] multiple of 16 (K=%d N=%d)", K, N);
  S2C_REQUIRE(lda >= K && ldc >= N, "mlp_layer_fwd: lda=%lld < K=%d or ldc=%lld < N=%d", lda, K, ldc, N);
  S2C_REQUIRE((pro_scale == nullptr) == (pro_shift == nullptr), "mlp_layer_fwd: scale/shift must both be given or both null");
  S2C_REQUIRE((stat_sum == nullptr) == (stat_sumsq == nullptr), "mlp_layer_fwd: stat_sum/stat_sumsq must both be given or both null");
  if (R == 0) return S2C_OK;
  S2C_REQUIRE(A && W && C, "mlp_layer_fwd: null pointer");
  GemmArgs g;
  g.A = A; g.lda = lda; g.K = K; g.pro_scale = pro_scale; g.pro_shift = pro_shift; g.W = W; g.N = N; g.C = C; g.ldc = ldc;
  g.stat_sum = stat_sum; g.stat_sumsq = stat_sumsq; g.R = R;
  const size_t smem = gemm_smem_bytes(N, K);
  S2C_REQUIRE(smem <= 227 * 1024, "mlp_layer_fwd: shared memory %zu B exceeds 227 KB (N=%d K=%d)", smem, N, K);
  S2C_CUDA(cudaFuncSetAttribute(mlp_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "mlp_layer_fwd smem attr");
  const long long tiles = (R + BM - 1) / BM;
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  mlp_gemm_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(g);
  S2C_CHECK_LAUNCH("mlp_layer_fwd launch");
  return S2C_OK;
}

// ================================================================================================
// Pooling end of the stack:  out[g, c] = max_s relu(Y[g*ns + s, c] * scale[c] + shift[c])
// (BatchNorm + ReLU of the LAST layer folded into F.max_pool2d, pointnet2_modules.py:255-257), plus the
// arg-max sample the backward pass routes the gradient to (first maximum, like max_pool2d).
// ================================================================================================
namespace s2c {
namespace {

__global__ void __launch_bounds__(256)
pool_fwd_kernel(const float *__restrict__ Y, long long ldy, long long G, int ns, int N, const float *__restrict__ scale,
                const float *__restrict__ shift, float *__restrict__ out, int *__restrict__ argmax) {
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  if (c >= N) return;
  const float sc = scale[c], sh = shift[c];
  for (long long grp = (long long)blockIdx.y * 4 + (threadIdx.x >> 6); grp < G; grp += (long long)gridDim.y * 4) {
    const float *y = Y + grp * ns * ldy + c;
    float best = -1.f;
    int bi = 0;
    for (int s = 0; s < ns; ++s) {
      const float v = fmaxf(fmaf(__ldg(y + (long long)s * ldy), sc, sh), 0.f);
      if (v > best) { best = v; bi = s; }
    }
    out[grp * N + c] = best;
    if (argmax) argmax[grp * N + c] = bi;
  }
}

// per-channel  sum_g = sum of the pooled gradient where the ReLU was active,  sum_gy = sum of g * y(argmax)
__global__ void __launch_bounds__(256)
pool_bwd_stats_kernel(const float *__restrict__ dpool, const int *__restrict__ argmax, const float *__restrict__ Y,
                      long long ldy, long long G, int ns, int N, const float *__restrict__ scale,
                      const float *__restrict__ shift, double *__restrict__ sum_g, double *__restrict__ sum_gy) {
  __shared__ float s1[4][64], s2[4][64];
  const int cl = threadIdx.x & 63, q = threadIdx.x >> 6;
  const int c = blockIdx.x * 64 + cl;
  float a = 0.f, b = 0.f;
  if (c < N) {
    const float sc = scale[c], sh = shift[c];
    for (long long grp = (long long)blockIdx.y * 4 + q; grp < G; grp += (long long)gridDim.y * 4) {
      const int s = argmax[grp * N + c];
      const float y = __ldg(Y + (grp * ns + s) * ldy + c);
      const float gval = (fmaf(y, sc, sh) > 0.f) ? dpool[grp * N + c] : 0.f;
      a += gval;
      b = fmaf(gval, y, b);
    }
  }
  s1[q][cl] = a; s2[q][cl] = b;
  __syncthreads();
  if (q == 0 && c < N) {
    atomicAdd(sum_g + c, (double)s1[0][cl] + (double)s1[1][cl] + (double)s1[2][cl] + (double)s1[3][cl]);
    atomicAdd(sum_gy + c, (double)s2[0][cl] + (double)s2[1][cl] + (double)s2[2][cl] + (double)s2[3][cl]);
  }
}

}  // namespace
}  // namespace s2c

extern "C" int s2c_pool_fwd(const float *Y, long long ldy, long long G, int ns, int N, const float *scale,
                            const float *shift, float *out, int *argmax, void *stream) {
  using namespace s2c;
  S2C_REQUIRE(G >= 0 && ns >= 1 && N >= 1 && ldy >= N, "pool_fwd: bad sizes");
  if (G == 0) return S2C_OK;
  S2C_REQUIRE(Y && scale && shift && out, "pool_fwd: null pointer");
  long long gy = (G + 3) / 4;
  if (gy > 65535) gy = 65535;  // the kernel strides over the groups
  dim3 grid((unsigned)ceil_div(N, 64), (unsigned)gy);
  pool_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Y, ldy, G, ns, N, scale, shift, out, argmax);
  S2C_CHECK_LAUNCH("pool_fwd");
  return S2C_OK;
}

extern "C" int s2c_pool_bwd_stats(const float *dpool, const int *argmax, const float *Y, long long ldy, long long G,
                                  int ns, int N, const float *scale, const float *shift, double *sum_g,
                                  double *sum_gy, void *stream) {
  using namespace s2c;
  S2C_REQUIRE(G >= 0 && ns >= 1 && N >= 1 && ldy >= N, "pool_bwd_stats: bad sizes");
  if (G == 0) return S2C_OK;
  S2C_REQUIRE(dpool && argmax && Y && scale && shift && sum_g && sum_gy, "pool_bwd_stats: null pointer");
  long long gy = (G + 3) / 4;
  if (gy > 256) gy = 256;
  dim3 grid((unsigned)ceil_div(N, 64), (unsigned)gy);
  pool_bwd_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dpool, argmax, Y, ldy, G, ns, N, scale, shift, sum_g, sum_gy);
  S2C_CHECK_LAUNCH("pool_bwd_stats");
  return S2C_OK;
}
