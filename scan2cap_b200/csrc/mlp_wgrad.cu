// Weight gradient of one shared-MLP layer on the tcgen05 tensor cores:
//
//     dW_l [C_l x C_prev] = sum over all R rows of  dY_l[r, :]^T  x  X'[r, :]          X' = relu(bn_{l-1}(Y_prev)) or the raw input
//
// This is a GEMM whose reduction dimension is the ROW index (R ~ 10^5..10^6) and whose output is tiny, the shape the
// library handles worst (the reference's cuDNN wgrad / cuBLAS "nt" split-K kernels take ~6 ms per step here).
// Both operands are consumed MN-MAJOR: in this GEMM the reduction index K is the data row and the operand "rows" are
// channels, which is exactly how dY and X lie in memory (channels contiguous).  Per 32-row chunk a thread loads 16
// bytes (4 channels of one data row) and stores them with ONE 16-byte shared-memory store per hi / lo half into the
// MN-major SWIZZLE_128B_BASE32B layout (32 channels = one 128-byte line, 4 data rows = one 512-byte swizzle atom); the
// tensor core transposes through the descriptor (a_major = b_major = MN).  (Round 1 staged the operands transposed into the
// K-major layout with eight rotated 4-byte stores per 16 loaded bytes: ncu showed the kernel instruction-bound on
// those stores and their address arithmetic, 22-30 % of its HBM roofline.)  dY is optionally formed on the fly (BatchNorm-backward affine
// a*g + b*y + c), X' by the BatchNorm + ReLU prologue; fp32 operands are split hi/lo (3xTF32).  Each persistent CTA
// accumulates its share of the rows in TMEM and adds its [C_l x C_prev] partial to the result with fp32 atomics.
#include <math.h>

#include "s2c_common.cuh"

namespace s2c {
namespace {

constexpr int CK = 32;  // rows (reduction elements) per pipeline chunk
constexpr int OS = 2;   // operand stages
constexpr int kLoadWarps = 8, kThreads = (kLoadWarps + 1) * 32;  // 8 load/transform warps + 1 MMA warp

struct WgradArgs {
  const float *dY; long long lddy;   // (R, C) dense dY, or g when `a` is given
  const float *Y; long long ldy;     // (R, C) pre-BN output of this layer (affine mode only)
  const float *a, *b, *c;            // [C] or null: dY = a*g + b*y + c
  const float *X; long long ldx;     // (R, P) previous layer's pre-BN output, or the raw input
  const float *xs, *xh;              // [P] or null: X' = relu(X*xs + xh)
  float *dW; long long lddw;         // (C, P) += (the caller zero-fills)
  long long R; int C, P;
  int MH, NB;                        // M halves (C <= 128*MH), 32-column blocks of P
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
  uint32_t *u = reinterpret_cast<uint32_t *>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
        "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
        "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// MN-major shared-memory descriptor.  For 32-bit (tf32) operands the only MN-major layout the tensor core takes is
// SWIZZLE_128B_BASE32B (layout type 1): 32 channels = one 128-byte line, FOUR data rows 128 bytes apart form a 512-byte
// swizzle atom in which the 32-byte groups of a line are XOR-permuted by (row & 3); the next 4 data rows lie SBO = 512
// bytes further, the next 32 channels LBO bytes further (one 4 KB block).
constexpr uint32_t kLBO = 32 * 32 * 4;  // = BLK
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(kLBO >> 4) << 16) | (32ull << 32) | (1ull << 46) | (1ull << 61);
}
// tf32 x tf32 -> f32 instruction descriptor, A and B both MN-major (bits 15 / 16)
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// one elected lane of a converged warp (keeps the MMA operands in uniform registers: see elect_one in mlp2.cu)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void split_tf32(float v, float &hi, float &lo) {
  hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  lo = v - hi;
}
// 16-byte store of a thread's 4 consecutive channels (seg = which 4-channel group of the 32-channel block) of data row
// k, hi / lo split: byte offset  k * 128 + (((seg / 2) ^ (k % 4)) * 32) + (seg % 2) * 16   (SWIZZLE_128B_BASE32B)
__device__ __forceinline__ uint32_t store_offset(int k, int seg) {
  return (uint32_t)(k * 128 + ((((seg >> 1) ^ k) & 3) << 5) + ((seg & 1) << 4));
}
__device__ __forceinline__ void store_split(unsigned char *hi_base, unsigned char *lo_base, uint32_t off, float4 v) {
  float4 h, l;
  split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
  *reinterpret_cast<float4 *>(hi_base + off) = h;
  *reinterpret_cast<float4 *>(lo_base + off) = l;
}
__device__ __forceinline__ float4 ld4_guard(const float *base, long long ld, long long row, long long R, int col, int ncols) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < R) {
    const float *p = base + row * ld + col;
    if (col + 3 < ncols) {
      v = __ldg(reinterpret_cast<const float4 *>(p));
    } else {
      if (col + 0 < ncols) v.x = __ldg(p + 0);
      if (col + 1 < ncols) v.y = __ldg(p + 1);
      if (col + 2 < ncols) v.z = __ldg(p + 2);
    }
  }
  return v;
}

constexpr uint32_t BLK = 32 * CK * 4;  // 32 operand rows (channels) x 32 k (data rows) fp32 = 4 KB

// CBT / NBT: compile-time bounds on the 32-channel blocks of dY / X a thread handles (register arrays);
// PF: register prefetch of the next chunk (small shapes only) and two CTAs per SM
template <int CBT, int NBT, bool PF>
__global__ void __launch_bounds__(kThreads, PF ? 2 : 1)
mlp_wgrad_kernel(WgradArgs g) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int MB = 4 * g.MH;   // A blocks per chunk (M padded to 128 per half)
  const int NB = g.NB;
  const uint32_t a_bytes = (uint32_t)MB * BLK, b_bytes = (uint32_t)NB * BLK;
  const uint32_t stage_bytes = 2 * (a_bytes + b_bytes);  // [A_hi | A_lo | B_hi | B_lo]
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + OS * stage_bytes);  // op_full[2], op_empty[2], done
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ncols = g.MH * NB * 32;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < ncols) tmem_cols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < OS; ++s) { mbar_init(&bars[s], kLoadWarps * 32); mbar_init(&bars[2 + s], 1); }
    mbar_init(&bars[4], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kLoadWarps) tmem_alloc(tmem_slot, tmem_cols);
  // zero the operand stages once: channel blocks beyond C (M padding to 128) are never written again
  for (uint32_t i = tid; i < OS * stage_bytes / 16; i += kThreads) reinterpret_cast<float4 *>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_async_proxy();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long num_chunks = (g.R + CK - 1) / CK;
  const long long my_chunks = (num_chunks > (long long)blockIdx.x) ? (num_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp < kLoadWarps) {
    // ===================== load + transform: global -> prologue -> hi/lo split -> MN-major operand blocks
    const int r = tid >> 3, seg = tid & 7;  // 256 threads = 32 rows x 8 sixteen-byte segments
    const uint32_t st_off = store_offset(r, seg);
    const int CB = (g.C + 31) / 32;         // real channel blocks of A
    const bool affine = g.a != nullptr;
    const bool xpro = g.xs != nullptr;
    float4 va[CBT], vy[CBT], vb[NBT];       // the chunk being staged
    float4 na[CBT], ny[CBT], nbv[NBT];      // PF: the next chunk, in flight while this one is staged
    auto issue_loads = [&](long long i, float4 *pa, float4 *py, float4 *pb) {
      const long long row = ((long long)blockIdx.x + i * gridDim.x) * CK + r;
#pragma unroll
      for (int mb = 0; mb < CBT; ++mb) {
        pa[mb] = make_float4(0.f, 0.f, 0.f, 0.f); py[mb] = pa[mb];
        if (mb < CB) {
          pa[mb] = ld4_guard(g.dY, g.lddy, row, g.R, mb * 32 + seg * 4, g.C);
          if (affine) py[mb] = ld4_guard(g.Y, g.ldy, row, g.R, mb * 32 + seg * 4, g.C);
        }
      }
#pragma unroll
      for (int nb = 0; nb < NBT; ++nb) {
        pb[nb] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nb < NB) pb[nb] = ld4_guard(g.X, g.ldx, row, g.R, nb * 32 + seg * 4, g.P);
      }
    };
    if (PF && my_chunks > 0) issue_loads(0, na, ny, nbv);
    for (long long i = 0; i < my_chunks; ++i) {
      const int s = (int)(i % OS);
      const long long row = ((long long)blockIdx.x + i * gridDim.x) * CK + r;
      unsigned char *a_hi = smem + (size_t)s * stage_bytes, *a_lo = a_hi + a_bytes, *b_hi = a_lo + a_bytes, *b_lo = b_hi + b_bytes;
      if (PF) {
#pragma unroll
        for (int mb = 0; mb < CBT; ++mb) { va[mb] = na[mb]; vy[mb] = ny[mb]; }
#pragma unroll
        for (int nb = 0; nb < NBT; ++nb) vb[nb] = nbv[nb];
        if (i + 1 < my_chunks) issue_loads(i + 1, na, ny, nbv);  // stays in flight during the stores below
      } else {
        issue_loads(i, va, vy, vb);  // all loads of the chunk first (memory-level parallelism), then the stage wait
      }
      if (i >= OS) mbar_wait(&bars[2 + s], (uint32_t)(((i / OS) - 1) & 1));
#pragma unroll
      for (int mb = 0; mb < CBT; ++mb) {
        if (mb < CB) {
          float4 v = va[mb];
          const int ch = mb * 32 + seg * 4;  // operand row of v.x (rows >= C stay zero from the one-time clear)
          if (affine) {
            const float4 y = vy[mb];
            float ca[4], cb[4], cc[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const bool ok = ch + e < g.C;
              ca[e] = ok ? __ldg(g.a + ch + e) : 0.f; cb[e] = ok ? __ldg(g.b + ch + e) : 0.f; cc[e] = ok ? __ldg(g.c + ch + e) : 0.f;
            }
            v.x = fmaf(ca[0], v.x, fmaf(cb[0], y.x, cc[0])); v.y = fmaf(ca[1], v.y, fmaf(cb[1], y.y, cc[1]));
            v.z = fmaf(ca[2], v.z, fmaf(cb[2], y.z, cc[2])); v.w = fmaf(ca[3], v.w, fmaf(cb[3], y.w, cc[3]));
            if (row >= g.R) v = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          store_split(a_hi, a_lo, st_off + (uint32_t)mb * BLK, v);
        }
      }
#pragma unroll
      for (int nb = 0; nb < NBT; ++nb) {
        if (nb < NB) {
          float4 v = vb[nb];
          const int ch = nb * 32 + seg * 4;
          if (xpro) {
            float xs[4], xh[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const bool ok = ch + e < g.P;
              xs[e] = ok ? __ldg(g.xs + ch + e) : 0.f; xh[e] = ok ? __ldg(g.xh + ch + e) : 0.f;
            }
            v.x = fmaxf(fmaf(v.x, xs[0], xh[0]), 0.f); v.y = fmaxf(fmaf(v.y, xs[1], xh[1]), 0.f);
            v.z = fmaxf(fmaf(v.z, xs[2], xh[2]), 0.f); v.w = fmaxf(fmaf(v.w, xs[3], xh[3]), 0.f);
            if (row >= g.R) v = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          store_split(b_hi, b_lo, st_off + (uint32_t)nb * BLK, v);
        }
      }
      fence_async_proxy();
      mbar_arrive(&bars[s]);
    }
  } else {
    // ===================== MMA issuer: D[128 x N] (per M half) += A[128 ch x 8 rows] * B[N ch x 8 rows]^T
    const int n0 = (NB > 8 ? 8 : NB) * 32, n1 = (NB > 8 ? NB - 8 : 0) * 32;
    const uint32_t idesc0 = make_idesc(128, n0), idesc1 = n1 ? make_idesc(128, n1) : 0u;
    const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);  // warp-uniform by construction
    for (long long i = 0; i < my_chunks; ++i) {
      const int s = (int)(i % OS);
      mbar_wait(&bars[s], (uint32_t)((i / OS) & 1));
      tc_fence_after();
      const uint32_t ah = smem_u32(smem + (size_t)s * stage_bytes), al = ah + a_bytes, bh = al + a_bytes, bl = bh + b_bytes;
      if (elect_one()) {
        for (int mh = 0; mh < g.MH; ++mh) {
          const uint32_t d0 = tbase + (uint32_t)(mh * NB * 32);
#pragma unroll
          for (int ks = 0; ks < CK / 8; ++ks) {  // UMMA K = 8 data rows = two 512-byte swizzle atoms of every 32-channel block
            const uint32_t acc = (i | ks) ? 1u : 0u;
            const uint32_t ao = (uint32_t)mh * 4 * BLK + ks * 1024, bo = ks * 1024;
            umma_tf32(d0, make_desc(ah + ao), make_desc(bh + bo), idesc0, acc);
            umma_tf32(d0, make_desc(ah + ao), make_desc(bl + bo), idesc0, 1u);
            umma_tf32(d0, make_desc(al + ao), make_desc(bh + bo), idesc0, 1u);
            if (n1) {
              const uint32_t bo1 = 8 * BLK + ks * 1024;  // channels 256.. of B
              umma_tf32(d0 + n0, make_desc(ah + ao), make_desc(bh + bo1), idesc1, acc);
              umma_tf32(d0 + n0, make_desc(ah + ao), make_desc(bl + bo1), idesc1, 1u);
              umma_tf32(d0 + n0, make_desc(al + ao), make_desc(bh + bo1), idesc1, 1u);
            }
          }
        }
        umma_commit(&bars[2 + s]);
        if (i == my_chunks - 1) umma_commit(&bars[4]);
      }
      __syncwarp();
    }
  }
  // ===================== epilogue: this CTA's partial [C x P] -> global with fp32 atomics (warps 0-3, TMEM quarter = warp)
  if (warp < 4 && my_chunks > 0) {
    const bool vec_red = ((g.lddw & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.dW) & 15) == 0);
    mbar_wait(&bars[4], 0);
    tc_fence_after();
    for (int mh = 0; mh < g.MH; ++mh) {
      const int m = mh * 128 + warp * 32 + lane;
      for (int nb = 0; nb < NB; ++nb) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)((mh * NB + nb) * 32), v);
        if (m < g.C) {
          float *o = g.dW + (long long)m * g.lddw + nb * 32;
          if (vec_red) {  // one 16-byte reduction per 4 columns (4x fewer atomic operations than scalar adds)
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (nb * 32 + j + 3 < g.P) {
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]),
                             "f"(v[j + 3])
                             : "memory");
              } else {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  if (nb * 32 + j + u < g.P) atomicAdd(o + j + u, v[j + u]);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb * 32 + j < g.P) atomicAdd(o + j, v[j]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kLoadWarps) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace
}  // namespace s2c

extern "C" int s2c_mlp_layer_bwd_weight(const float *dY, long long lddy, const float *Y, long long ldy, const float *a,
                                        const float *b, const float *c, const float *X, long long ldx, const float *xs,
                                        const float *xh, long long R, int C, int P, float *dW, long long lddw,
                                        void *stream) {
  using namespace s2c;
  S2C_REQUIRE(R >= 0 && C >= 1 && P >= 1 && lddy >= C && ldx >= P && lddw >= P, "mlp_layer_bwd_weight: bad sizes");
  S2C_REQUIRE((lddy & 3) == 0 && (ldx & 3) == 0, "mlp_layer_bwd_weight: leading dimensions must be multiples of 4");
  S2C_REQUIRE((a == nullptr) == (b == nullptr) && (a == nullptr) == (c == nullptr), "mlp_layer_bwd_weight: a, b, c must be given together");
  S2C_REQUIRE(a == nullptr || (Y != nullptr && ldy >= C && (ldy & 3) == 0), "mlp_layer_bwd_weight: affine mode needs Y");
  S2C_REQUIRE((xs == nullptr) == (xh == nullptr), "mlp_layer_bwd_weight: xs/xh must be given together");
  const int MH = (C + 127) / 128, NB = (P + 31) / 32;
  if (C > 256 || NB > 9 || MH * NB > 16) {
    set_error("mlp_layer_bwd_weight: unsupported shape C=%d P=%d (need C<=256, P<=288, ceil(C/128)*ceil(P/32)<=16)", C, P);
    return S2C_ERR_UNSUPPORTED;
  }
  if (R == 0) return S2C_OK;
  S2C_REQUIRE(dY && X && dW, "mlp_layer_bwd_weight: null pointer");
  S2C_REQUIRE(((uintptr_t)dY & 15) == 0 && ((uintptr_t)X & 15) == 0 && (!Y || ((uintptr_t)Y & 15) == 0), "mlp_layer_bwd_weight: operands must be 16-byte aligned");
  WgradArgs g;
  g.dY = dY; g.lddy = lddy; g.Y = Y; g.ldy = ldy; g.a = a; g.b = b; g.c = c; g.X = X; g.ldx = ldx; g.xs = xs; g.xh = xh;
  g.dW = dW; g.lddw = lddw; g.R = R; g.C = C; g.P = P; g.MH = MH; g.NB = NB;
  const size_t smem = (size_t)OS * 2 * (4 * MH + NB) * BLK + 8 * 8 + 16;
  S2C_REQUIRE(smem <= 227 * 1024, "mlp_layer_bwd_weight: shared memory %zu B exceeds 227 KB", smem);
  const long long chunks = (R + CK - 1) / CK;
  const int CB = (C + 31) / 32;
  // chunks per CTA: every CTA ends with C*P/4 vector reductions into dW (~70 per ns device-wide) and spends ~2 us per
  // chunk, so short reductions (R of a few thousand rows) want fewer, longer-running CTAs: minimise
  //   cpc * 2000 ns + (chunks / cpc) * C*P / 280 ns   ->   cpc = sqrt(chunks * C*P / 560000)
  long long cpc = (long long)(sqrt((double)chunks * (double)C * (double)P / 560000.0) + 0.5);
  if (cpc < 1) cpc = 1;
#define S2C_WGRAD(CBT, NBT, PF)                                                                                         \
  do {                                                                                                                  \
    auto kern = mlp_wgrad_kernel<CBT, NBT, PF>;                                                                         \
    S2C_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "wgrad smem attr");    \
    const long long cap = (long long)kNumSMs * ((PF) ? 2 : 1);                                                          \
    const long long want = (chunks + cpc - 1) / cpc;                                                                    \
    const int grid = (int)(want < cap ? want : cap);                                                                    \
    kern<<<grid, kThreads, smem, (cudaStream_t)stream>>>(g);                                                            \
    S2C_CHECK_LAUNCH("mlp_wgrad launch");                                                                               \
    return S2C_OK;                                                                                                      \
  } while (0)
  // small shapes: two CTAs per SM (<= 113 KB of shared memory each) with register prefetch of the next chunk
  if (CB <= 2 && NB <= 2 && smem <= 113 * 1024) S2C_WGRAD(2, 2, true);
  if (CB <= 4 && NB <= 2 && smem <= 113 * 1024) S2C_WGRAD(4, 2, true);
  if (CB <= 4 && NB <= 4) S2C_WGRAD(4, 4, false);
  if (CB <= 8 && NB <= 4) S2C_WGRAD(8, 4, false);
  S2C_WGRAD(8, 9, false);
#undef S2C_WGRAD
}
