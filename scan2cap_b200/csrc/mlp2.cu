// Pipelined, warp-specialised version of the grouped-MLP layer kernel (see mlp.cu for the arithmetic):
//
//   loader warp      cp.async.bulk (TMA engine, no register staging) : raw fp32 rows of the activation tile -> smem ring,
//                    and the pre-split / pre-swizzled weight chunk straight into the operand stage; mbarrier complete_tx
//   transform warps  raw tile -> [previous layer's BatchNorm + ReLU] -> tf32 hi/lo split -> K-major SWIZZLE_128B operand
//   MMA warp         one thread issues tcgen05.mma.cta_group::1.kind::tf32 (3xTF32), tcgen05.commit -> mbarriers
//   epilogue warps   tcgen05.ld from the DOUBLE-BUFFERED TMEM accumulator -> 128-byte row stores + per-channel sum / sumsq
//                    (warp transpose-reduce), overlapping the next tile's loads and MMAs
// Every hand-off is an mbarrier; nothing in the steady state is a CTA-wide barrier.
#include <cuda.h>

#include "s2c_common.cuh"

namespace s2c {
namespace {

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int kMaxOS = 4, kMaxWS = 8, kMaxRS = 8;  // ring depths are chosen per shape by the host (choose_pipe)
constexpr int kLoaderWarp = 0, kMmaWarp = 1, kFirstTransformWarp = 2;
// warps: loader, MMA issuer, TW transform warps, 4 epilogue warps.  TW = 4 with the operand staged through shared memory;
// 8 on the tensor-memory operand path, where staging a chunk (one tile row per thread) was the slowest pipeline stage
// (s2c_mlp_probe: 0.65-0.85 us per chunk, half of it barrier / tcgen05.st round trips): two groups of four warps take
// the even and the odd chunks, so two chunks are being staged at any time.
constexpr int transform_warps(bool atm) { return atm ? 8 : 4; }
constexpr int block_threads(bool atm) { return (6 + transform_warps(atm)) * 32; }

// operand prologues (applied by the transform warps while staging the A tile)
constexpr int PRO_BNRELU = 1;   // a' = relu(a*p0[k] + p1[k])             (p0 == null: identity)         -- forward
constexpr int PRO_AFFINE2 = 2;  // a' = p0[k]*g + p1[k]*y + p2[k]          two raw tiles (g = A, y = A2)  -- BatchNorm backward
constexpr int PRO_POOL = 3;     // as AFFINE2 with g = (argmax == sample && relu active) ? dpool : 0     -- last layer
// epilogues
constexpr int EPI_STORE_STATS = 0;  // C = acc ; optional column sum / sum of squares
constexpr int EPI_MASK_STATS = 1;   // C = acc where relu(bn(Yprev)) was active else 0 ; column sum(C), sum(C*Yprev)

struct Gemm2Args {
  const float *A; long long lda; int K;
  const float *p0, *p1, *p2, *p3, *p4;  // per-K prologue coefficients (see PRO_*); p3/p4 = last layer's scale/shift (PRO_POOL)
  const float *dpool; const int *argmax; int ns;  // PRO_POOL: (G, K) pooled gradient and arg-max sample
  int GT;  // PRO_POOL: groups one 128-row tile can touch = rows of the (argmax, dpool) slabs staged with every raw chunk
  const unsigned char *wprep;  // KC chunks of [hi: N x 128 B swizzled][lo: N x 128 B swizzled]
  float *C; long long ldc;
  float *dY_out;  // optional (R, K): the staged operand a' written back (the weight-gradient GEMM consumes it)
  const float *Yprev; long long ldyp; const float *e_scale, *e_shift;  // EPI_MASK_STATS
  double *stat_sum; double *stat_sumsq;
  long long R;
  int RS;    // raw-tile ring depth (TMA -> transform warps)
  int OS;    // A-operand ring depth (transform warps -> MMA)
  int WS;    // weight-chunk ring depth (TMA -> MMA); decoupled from the A ring so weight chunks are prefetched WS chunks
             // ahead instead of only after the MMAs that last used their operand stage have retired
  int wres;  // 1: WS == number of K chunks, every chunk is loaded ONCE and stays resident for all tiles of the CTA
  unsigned long long *probe; int probe_cap;  // optional profiling aid (s2c_mlp_probe): clock64 stamps of CTA 0's roles
};

// probe slots: [0] kernel start; chunk i (i < 64): 16 + 16*i + {0: transform begins waiting, 1: raw tile landed, 2: operand
// stage free, 3: operand staged, 4: MMA warp saw operand + weights, 5: MMAs issued, 6: raw TMA issued, 7: weight TMA issued};
// tile t: 16 + 16*t + {8: accumulator complete (epilogue), 9: epilogue done}
#define S2C_PROBE(slot)                                                                                   \
  do {                                                                                                    \
    if (g.probe != nullptr && blockIdx.x == 0 && (slot) < g.probe_cap) g.probe[(slot)] = (unsigned long long)clock64(); \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
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
// global -> shared bulk copy executed by the TMA engine; completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
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
// the same with the A operand in tensor memory (M = 128: lane = row, one 32-bit column per K element)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// 32 columns of this thread's TMEM lane (lane = 32 * (warp % 4) + lane id)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float *v) {
  const uint32_t *u = reinterpret_cast<const uint32_t *>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]), "r"(u[9]), "r"(u[10]),
      "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]), "r"(u[16]), "r"(u[17]), "r"(u[18]), "r"(u[19]), "r"(u[20]),
      "r"(u[21]), "r"(u[22]), "r"(u[23]), "r"(u[24]), "r"(u[25]), "r"(u[26]), "r"(u[27]), "r"(u[28]), "r"(u[29]), "r"(u[30]),
      "r"(u[31])
      : "memory");
}
// one elected lane of a converged warp: unlike `lane == 0`, the compiler keeps the region's operands in UNIFORM registers,
// so every tcgen05.mma below is one UTCHMMA.  Behind `if (lane == 0)` each MMA was wrapped in an ELECT / R2UR.BROADCAST x5 /
// BRA.U.ANY waterfall (descriptors moved from vector to uniform registers per instruction): ~100 cycles per MMA, which
// s2c_mlp_probe showed as 0.7 us of MMA issue per 32-column chunk whatever the tile width or the operand source.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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
// the same load without the wait: the registers are valid only after tmem_ld_wait(v) -- lets the epilogue fetch the next
// column block from TMEM while it transposes / stores the current one
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float *v) {
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
}
// wait for the outstanding tcgen05.ld; the "+r" operands tie the loaded registers to the wait so that no use of them
// can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait(float *v) {
  uint32_t *u = reinterpret_cast<uint32_t *>(v);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]), "+r"(u[8]),
                 "+r"(u[9]), "+r"(u[10]), "+r"(u[11]), "+r"(u[12]), "+r"(u[13]), "+r"(u[14]), "+r"(u[15])
               :
               : "memory");
  asm volatile(""
               : "+r"(u[16]), "+r"(u[17]), "+r"(u[18]), "+r"(u[19]), "+r"(u[20]), "+r"(u[21]), "+r"(u[22]), "+r"(u[23]),
                 "+r"(u[24]), "+r"(u[25]), "+r"(u[26]), "+r"(u[27]), "+r"(u[28]), "+r"(u[29]), "+r"(u[30]), "+r"(u[31])
               :
               : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {  // K-major SWIZZLE_128B, SBO = 1024 B, version 1
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ __forceinline__ uint32_t swz(int r, int seg) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((seg ^ (r & 7)) << 4));
}
__device__ __forceinline__ void split_tf32(float v, float &hi, float &lo) {
  hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  lo = v - hi;
}
__device__ __forceinline__ void store_split(unsigned char *hi_base, unsigned char *lo_base, uint32_t off, float4 v) {
  float4 h, l;
  split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
  *reinterpret_cast<float4 *>(hi_base + off) = h;
  *reinterpret_cast<float4 *>(lo_base + off) = l;
}

// weight preparation: W (N, K) fp32 -> per K-chunk [hi | lo] blocks in the operand layout (done once per call)
// kmajor = 1: operand row n, column k = W[n*ld + k]   (forward: W is (N, K))
// kmajor = 0: operand row n, column k = W[k*ld + n]   (backward data: the operand is W^T of a (K, N) weight)
__global__ void w_prep_kernel(const float *__restrict__ W, int N, int K, int ld, int kmajor, unsigned char *__restrict__ out) {
  const int KC = (K + BK - 1) / BK;
  const int total = KC * N * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int seg = i & 7, r = (i >> 3) % N, kc = i / (8 * N);
    const int kk = kc * BK + seg * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t sk = kmajor ? 1 : (size_t)ld, sn = kmajor ? (size_t)ld : 1;
    const float *p = W + (size_t)r * sn + (size_t)kk * sk;
    if (kk + 0 < K) v.x = p[0];
    if (kk + 1 < K) v.y = p[sk];
    if (kk + 2 < K) v.z = p[2 * sk];
    if (kk + 3 < K) v.w = p[3 * sk];
    unsigned char *hi = out + (size_t)kc * N * 256;
    store_split(hi, hi + (size_t)N * 128, swz(r, seg), v);
  }
}

// 2-D TMA load of one [128 rows x 32 fp32] box of the activation matrix (rows / columns outside the tensor are
// zero-filled by the hardware); completes 16 KB on the mbarrier
__device__ __forceinline__ void tma_load_box(void *dst_smem, const CUtensorMap *tmap, int x, int y, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst_smem)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}

// ATM: the A operand travels through TENSOR memory instead of shared memory (forward kernels).  Measured with
// s2c_mlp_probe, the layer kernel is bound by shared-memory bandwidth: per 128 x 32 chunk the hi / lo operand costs 32 KB
// of stores and 12 x 4 KB of MMA operand reads, 40 % of all shared-memory traffic.  With ATM the transform warps own one
// tile row per thread (row = TMEM lane), read it from the TMA-swizzled raw tile, and hand hi / lo to the tensor core with
// two tcgen05.st; the MMAs read A from TMEM (tcgen05.mma [d], [a], b_desc) and only B from shared memory.
template <int N, int PRO, int EPI, bool ATM>
__global__ void __launch_bounds__(block_threads(ATM), 1)
mlp_gemm2_kernel(Gemm2Args g, const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2,
                 const __grid_constant__ CUtensorMap tmap_am, const __grid_constant__ CUtensorMap tmap_dp) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr uint32_t A_BYTES = BM * BK * 4;            // 16 KB per hi / lo / raw block
  constexpr uint32_t W_BYTES = N * BK * 4;             // N x 128 B per hi / lo block
  constexpr uint32_t RAW_TILES = (PRO == PRO_AFFINE2) ? 2 : 1;   // raw tiles per chunk (g and y for the affine prologue)
  // PRO_POOL: every raw stage also carries the [GT groups x 32 channels] slabs of argmax (int32) and dpool (fp32) that
  // cover the tile's rows, fetched by the TMA engine with the y tile -- the transform warps never touch global memory
  const uint32_t SLAB_BYTES = (PRO == PRO_POOL) ? (uint32_t)g.GT * 128u : 0u;
  const uint32_t RAW_BYTES = RAW_TILES * A_BYTES + 2 * SLAB_BYTES;
  constexpr int NPRO = (PRO == PRO_BNRELU) ? 2 : (PRO == PRO_AFFINE2) ? 3 : 5;
  constexpr uint32_t ACC_COLS = 2 * N, A_COLS = ATM ? 2 * 64 : 0;  // accumulators (double-buffered) | two [hi | lo] A stages
  constexpr uint32_t TMEM_COLS = (ACC_COLS + A_COLS <= 32) ? 32 : (ACC_COLS + A_COLS <= 64) ? 64 : (ACC_COLS + A_COLS <= 128) ? 128
                                 : (ACC_COLS + A_COLS <= 256) ? 256 : 512;
  static_assert(!ATM || PRO == PRO_BNRELU, "the tensor-memory operand path is built for the forward prologue");
  static_assert(ACC_COLS + A_COLS <= 512, "tensor memory");
  constexpr int TW = transform_warps(ATM), kFirstEpilogueWarp = kFirstTransformWarp + TW, kThreads = block_threads(ATM);
  const int RS = g.RS, OS = g.OS, WS = g.WS;
  const bool wres = g.wres != 0;
  const int KC = (g.K + BK - 1) / BK;
  unsigned char *a_base = smem;                                  // OS x [A hi | A lo]
  unsigned char *w_base = smem + (ATM ? 0 : (size_t)OS * 2 * A_BYTES);  // WS x [W hi | W lo]
  unsigned char *raw_base = w_base + (size_t)WS * 2 * W_BYTES;  // RS x raw stage
  float *s_pro = reinterpret_cast<float *>(raw_base + (size_t)RS * RAW_BYTES);  // [NPRO][KC*BK] coefficients
  float *s_epi = s_pro + NPRO * KC * BK;                                         // [4 warps][32 rows][36] epilogue staging
  uint64_t *bars = reinterpret_cast<uint64_t *>(s_epi + 4 * 32 * 36);
  uint64_t *raw_full = bars, *raw_empty = bars + 8, *op_full = bars + 16, *op_empty = bars + 20, *acc_full = bars + 24,
           *acc_empty = bars + 26, *w_full = bars + 28, *w_empty = bars + 36;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 44);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool has_pro = g.p0 != nullptr;
  if (tid == 0) {
    for (int s = 0; s < RS; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 128); }  // four warps stage a chunk
    for (int s = 0; s < OS; ++s) { mbar_init(&op_full[s], 128); mbar_init(&op_empty[s], 1); }
    for (int s = 0; s < WS; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, TMEM_COLS);
  {
    const float *src[5] = {g.p0, g.p1, g.p2, g.p3, g.p4};
    for (int i = tid; i < NPRO * KC * BK; i += kThreads) {
      const int a = i / (KC * BK), k = i - a * (KC * BK);
      // identity defaults: scale-like arrays (index 0, and 3 for PRO_POOL) -> 1, the others -> 0
      const float dflt = (a == 0 || a == 3) ? 1.f : 0.f;
      s_pro[i] = (src[a] != nullptr && k < g.K) ? src[a][k] : dflt;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) S2C_PROBE(0);
  const uint32_t tmem_base = *tmem_slot;
  const long long num_tiles = (g.R + BM - 1) / BM;
  const long long my_tiles = (num_tiles > (long long)blockIdx.x) ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long total_chunks = my_tiles * KC;

  if (warp == kLoaderWarp) {
    // ===================== loader: raw activation rows (ring of RS) and weight chunks (ring of WS, or resident)
    if (lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
    long long ia = 0, iw = 0;          // next raw chunk / next weight chunk to issue
    long long ta = blockIdx.x; int ka = 0;  // (tile, kc) of chunk ia
    int kw = 0;
    const long long w_total = wres ? (total_chunks < KC ? total_chunks : (long long)KC) : total_chunks;
    // The warp shares its scheduler with a transform and an epilogue warp: it must not spin.  When only raw chunks are
    // left to issue (always, once resident weights are in) it blocks on the stage's mbarrier (try_wait suspends the warp
    // in hardware); while both rings are being fed it polls them alternately with a sleep between unsuccessful rounds.
    while (ia < total_chunks || iw < w_total) {
      bool progressed = false;
      if (ia < total_chunks) {
        const int rs = (int)(ia % RS);
        // lane 0 polls and broadcasts: the decision (and the counters below) must be warp-uniform
        if (ia < RS || __shfl_sync(0xffffffffu, lane == 0 ? (int)mbar_test(&raw_empty[rs], (uint32_t)(((ia / RS) - 1) & 1)) : 0, 0)) {
          if (lane == 0) {
            S2C_PROBE(16 + 16 * (int)ia + 6);
            mbar_arrive_expect_tx(&raw_full[rs], RAW_BYTES);
            unsigned char *dst = raw_base + (size_t)rs * RAW_BYTES;
            if (PRO == PRO_POOL) {  // y tile + the pooled-gradient slabs of the tile's groups; g is rebuilt from them
              const int g0 = (int)((ta * BM) / g.ns);
              tma_load_box(dst, &tmap_a2, ka * BK, (int)(ta * BM), &raw_full[rs]);
              tma_load_box(dst + A_BYTES, &tmap_am, ka * BK, g0, &raw_full[rs]);
              tma_load_box(dst + A_BYTES + SLAB_BYTES, &tmap_dp, ka * BK, g0, &raw_full[rs]);
            } else {
              tma_load_box(dst, &tmap_a, ka * BK, (int)(ta * BM), &raw_full[rs]);
              if (PRO == PRO_AFFINE2) tma_load_box(dst + A_BYTES, &tmap_a2, ka * BK, (int)(ta * BM), &raw_full[rs]);
            }
          }
          __syncwarp();
          ++ia;
          if (++ka == KC) { ka = 0; ta += gridDim.x; }
          progressed = true;
        }
      }
      if (iw < w_total) {
        const int ws = (int)(iw % WS);
        if (iw < WS || __shfl_sync(0xffffffffu, lane == 0 ? (int)mbar_test(&w_empty[ws], (uint32_t)(((iw / WS) - 1) & 1)) : 0, 0)) {
          if (lane == 0) {
            S2C_PROBE(16 + 16 * (int)iw + 7);
            unsigned char *dst = w_base + (size_t)ws * 2 * W_BYTES;
            mbar_arrive_expect_tx(&w_full[ws], 2 * W_BYTES);
            bulk_g2s(dst, g.wprep + (size_t)kw * 2 * W_BYTES, 2 * W_BYTES, &w_full[ws]);
          }
          __syncwarp();
          ++iw;
          if (++kw == KC) kw = 0;
          progressed = true;
        }
      }
      if (!progressed) {
        if (iw >= w_total) mbar_wait(&raw_empty[(int)(ia % RS)], (uint32_t)(((ia / RS) - 1) & 1));
        else if (ia >= total_chunks) mbar_wait(&w_empty[(int)(iw % WS)], (uint32_t)(((iw / WS) - 1) & 1));
        else __nanosleep(64);
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer
    const uint32_t idesc = make_idesc(BM, N);
    const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);  // warp-uniform by construction (see elect_one)
    int os = 0, ws = 0;
    uint32_t oph = 0, wph = 0;  // phase parities of the A-operand / weight rings
    int pit = 0;
    for (long long t = 0; t < my_tiles; ++t) {
      const int ab = (int)(t & 1);
      if (t >= 2) mbar_wait(&acc_empty[ab], (uint32_t)(((t >> 1) - 1) & 1));
      tc_fence_after();
      const uint32_t d_tmem = tbase + (uint32_t)(ab * N);
      for (int kc = 0; kc < KC; ++kc) {
        mbar_wait(&op_full[os], oph);
        mbar_wait(&w_full[ws], wres ? 0u : wph);  // resident chunks complete phase 0 once and stay valid
        tc_fence_after();
        const uint32_t ah = smem_u32(a_base + (size_t)os * 2 * A_BYTES), al = ah + A_BYTES;
        const uint32_t bh = smem_u32(w_base + (size_t)ws * 2 * W_BYTES), bl = bh + W_BYTES;
        const uint32_t ath = tbase + ACC_COLS + (uint32_t)os * 64u, atl = ath + 32u;  // ATM: A stage in tensor memory
        // only the 8-column steps that hold columns of the matrix: a first layer over [xyz, 0 | features] rows has
        // K = 8, 132 or 260 valid columns, i.e. one step in its last (or only) chunk instead of four
        const int steps = min(BK / 8, (g.K - kc * BK + 7) / 8);
        if (elect_one()) {
          S2C_PROBE(16 + 16 * pit + 4);
#pragma unroll
          for (int j = 0; j < BK / 8; ++j) {
            if (j >= steps) break;
            const uint32_t o = j * 32;
            if (ATM) {
              umma_tf32_ts(d_tmem, ath + j * 8, make_desc(bh + o), idesc, (kc | j) ? 1u : 0u);
              umma_tf32_ts(d_tmem, ath + j * 8, make_desc(bl + o), idesc, 1u);
              umma_tf32_ts(d_tmem, atl + j * 8, make_desc(bh + o), idesc, 1u);
            } else {
              umma_tf32(d_tmem, make_desc(ah + o), make_desc(bh + o), idesc, (kc | j) ? 1u : 0u);
              umma_tf32(d_tmem, make_desc(ah + o), make_desc(bl + o), idesc, 1u);
              umma_tf32(d_tmem, make_desc(al + o), make_desc(bh + o), idesc, 1u);
            }
          }
          umma_commit(&op_empty[os]);
          if (!wres) umma_commit(&w_empty[ws]);
          if (kc == KC - 1) umma_commit(&acc_full[ab]);
          S2C_PROBE(16 + 16 * pit + 5);
        }
        __syncwarp();
        ++pit;
        if (++os == OS) { os = 0; oph ^= 1u; }
        if (++ws == WS) { ws = 0; wph ^= 1u; }
      }
    }
  } else if (warp < kFirstEpilogueWarp) {
    // ===================== transform: raw rows -> prologue -> hi/lo split -> swizzled operand
    const int tt = tid - kFirstTransformWarp * 32;  // 0..127
    const int seg = tt & 7;
    long long it = 0;
    int rs = 0, os = 0;
    uint32_t rph = 0, oph = 0;  // phase parities of the raw / A-operand rings
    for (long long t = 0; t < my_tiles; ++t) {
      const long long row0 = ((long long)blockIdx.x + t * gridDim.x) * BM;
      for (int kc = 0; kc < KC; ++kc, ++it) {
        const int kk = kc * BK + seg * 4;
        const int KP = KC * BK;
        const float4 c0 = *reinterpret_cast<const float4 *>(s_pro + kk);
        const float4 c1 = *reinterpret_cast<const float4 *>(s_pro + KP + kk);
        float4 c2 = make_float4(0.f, 0.f, 0.f, 0.f), c3 = c2, c4 = c2;
        if (PRO != PRO_BNRELU) c2 = *reinterpret_cast<const float4 *>(s_pro + 2 * KP + kk);
        if (PRO == PRO_POOL) {
          c3 = *reinterpret_cast<const float4 *>(s_pro + 3 * KP + kk);
          c4 = *reinterpret_cast<const float4 *>(s_pro + 4 * KP + kk);
        }
        // PRO_POOL: (argmax, dpool) depend on the row's GROUP only; the slab row (group - first group of the tile) and
        // the sample index advance incrementally from pass to pass (16 rows apart)
        int pgrp = 0, psmp = 0;
        if (PRO == PRO_POOL) {
          const long long base = row0 + (tt >> 3);
          const long long gq = base / g.ns;
          pgrp = (int)(gq - row0 / g.ns);
          psmp = (int)(base - gq * g.ns);
        }
        if (ATM && ((int)(it & 1) != ((warp - kFirstTransformWarp) >> 2))) {  // the other group's chunk
          if (++rs == RS) { rs = 0; rph ^= 1u; }
          if (++os == OS) { os = 0; oph ^= 1u; }
          continue;
        }
        if ((tt & 127) == 0) S2C_PROBE(16 + 16 * (int)it + 0);
        mbar_wait(&raw_full[rs], rph);
        if ((tt & 127) == 0) S2C_PROBE(16 + 16 * (int)it + 1);
        if (it >= OS) mbar_wait(&op_empty[os], oph ^ 1u);
        if ((tt & 127) == 0) S2C_PROBE(16 + 16 * (int)it + 2);
        const unsigned char *raw = raw_base + (size_t)rs * RAW_BYTES;
        unsigned char *a_hi = a_base + (size_t)os * 2 * A_BYTES, *a_lo = a_hi + A_BYTES;
        if constexpr (ATM) {
          // one tile row per thread: row = TMEM lane = 32 * (warp % 4) + lane.  The raw tile was written by the TMA engine
          // with the 128-byte swizzle (16-byte chunk c of row r sits at chunk c ^ (r & 7)), so the eight lanes of a
          // shared-memory wavefront -- eight consecutive rows, same chunk -- hit eight different bank groups.
          tc_fence_after();
          const int q = warp & 3, r = q * 32 + lane;
          const bool row_ok = row0 + r < g.R;
          float4 v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const float4 *>(raw + r * 128 + ((j ^ (r & 7)) << 4));
          float hi[32], lo[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k4 = kc * BK + j * 4;
            if (has_pro) {
              const float4 s0 = *reinterpret_cast<const float4 *>(s_pro + k4);            // warp-wide broadcast
              const float4 s1 = *reinterpret_cast<const float4 *>(s_pro + KC * BK + k4);
              v[j].x = fmaxf(fmaf(v[j].x, s0.x, s1.x), 0.f); v[j].y = fmaxf(fmaf(v[j].y, s0.y, s1.y), 0.f);
              v[j].z = fmaxf(fmaf(v[j].z, s0.z, s1.z), 0.f); v[j].w = fmaxf(fmaf(v[j].w, s0.w, s1.w), 0.f);
            }
            if (!row_ok || k4 + 0 >= g.K) v[j].x = 0.f;
            if (!row_ok || k4 + 1 >= g.K) v[j].y = 0.f;
            if (!row_ok || k4 + 2 >= g.K) v[j].z = 0.f;
            if (!row_ok || k4 + 3 >= g.K) v[j].w = 0.f;
            split_tf32(v[j].x, hi[4 * j + 0], lo[4 * j + 0]); split_tf32(v[j].y, hi[4 * j + 1], lo[4 * j + 1]);
            split_tf32(v[j].z, hi[4 * j + 2], lo[4 * j + 2]); split_tf32(v[j].w, hi[4 * j + 3], lo[4 * j + 3]);
          }
          const uint32_t at = tmem_base + ((uint32_t)(q * 32) << 16) + ACC_COLS + (uint32_t)os * 64u;
          tmem_st32(at, hi);
          tmem_st32(at + 32u, lo);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&op_full[os]);
          mbar_arrive(&raw_empty[rs]);
          if ((tt & 127) == 0) S2C_PROBE(16 + 16 * (int)it + 3);
          if (++rs == RS) { rs = 0; rph ^= 1u; }
          if (++os == OS) { os = 0; oph ^= 1u; }
          continue;
        }
        // The shared-memory loads of PB passes are issued together, then the arithmetic and the operand stores follow:
        // the raw ring and the operand stages are carved out of one array, so the compiler must keep every load behind
        // the stores that precede it in program order -- with one pass per iteration each pass exposed a full
        // shared-memory round trip (s2c_mlp_probe: 650-750 ns of staging per 16 KB chunk, the slowest pipeline stage).
        constexpr int PB = (PRO == PRO_BNRELU) ? 8 : 4;
        const bool staged = kk < ((g.K + 7) & ~7);  // columns beyond the last 8-column MMA step are never read
#pragma unroll
        for (int p0 = 0; p0 < BM / 16; p0 += PB) {
          if (!staged) break;
          float4 vv[PB], yy[PB], dpv[PB];
          int4 amv[PB];
          int smpv[PB];
#pragma unroll
          for (int j = 0; j < PB; ++j) {
            const int r = (p0 + j) * 16 + (tt >> 3);
            vv[j] = *reinterpret_cast<const float4 *>(raw + r * 128 + seg * 16);
            if (PRO == PRO_AFFINE2) yy[j] = *reinterpret_cast<const float4 *>(raw + A_BYTES + r * 128 + seg * 16);
            if (PRO == PRO_POOL) {
              smpv[j] = psmp;
              amv[j] = make_int4(-1, -1, -1, -1);  // never equals a sample index: rows / columns outside the matrix
              dpv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (row0 + r < g.R && kk < g.K) {
                amv[j] = *reinterpret_cast<const int4 *>(raw + A_BYTES + pgrp * 128 + seg * 16);
                dpv[j] = *reinterpret_cast<const float4 *>(raw + A_BYTES + SLAB_BYTES + pgrp * 128 + seg * 16);
              }
              const int nx = psmp + 16;  // the next pass is 16 rows further
              if (g.ns >= 16) {          // at most one group boundary per pass: no division
                const bool wrap = nx >= g.ns;
                pgrp += wrap ? 1 : 0;
                psmp = wrap ? nx - g.ns : nx;
              } else {
                const int q = nx / g.ns;
                pgrp += q;
                psmp = nx - q * g.ns;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < PB; ++j) {
            const int r = (p0 + j) * 16 + (tt >> 3);
            const long long row = row0 + r;
            float4 v = vv[j];
            if (PRO == PRO_BNRELU) {
              if (has_pro) {
                v.x = fmaxf(fmaf(v.x, c0.x, c1.x), 0.f); v.y = fmaxf(fmaf(v.y, c0.y, c1.y), 0.f);
                v.z = fmaxf(fmaf(v.z, c0.z, c1.z), 0.f); v.w = fmaxf(fmaf(v.w, c0.w, c1.w), 0.f);
              }
            } else if (PRO == PRO_AFFINE2) {
              const float4 y = yy[j];
              v.x = fmaf(c0.x, v.x, fmaf(c1.x, y.x, c2.x)); v.y = fmaf(c0.y, v.y, fmaf(c1.y, y.y, c2.y));
              v.z = fmaf(c0.z, v.z, fmaf(c1.z, y.z, c2.z)); v.w = fmaf(c0.w, v.w, fmaf(c1.w, y.w, c2.w));
            } else {  // PRO_POOL: v holds y; g is the pooled gradient at the arg-max sample where the ReLU was active
              const float4 y = v;
              const int4 am = amv[j];
              const float4 dp = dpv[j];
              const int smp = smpv[j];
              float4 gg;
              gg.x = (am.x == smp && fmaf(y.x, c3.x, c4.x) > 0.f) ? dp.x : 0.f;
              gg.y = (am.y == smp && fmaf(y.y, c3.y, c4.y) > 0.f) ? dp.y : 0.f;
              gg.z = (am.z == smp && fmaf(y.z, c3.z, c4.z) > 0.f) ? dp.z : 0.f;
              gg.w = (am.w == smp && fmaf(y.w, c3.w, c4.w) > 0.f) ? dp.w : 0.f;
              v.x = fmaf(c0.x, gg.x, fmaf(c1.x, y.x, c2.x)); v.y = fmaf(c0.y, gg.y, fmaf(c1.y, y.y, c2.y));
              v.z = fmaf(c0.z, gg.z, fmaf(c1.z, y.z, c2.z)); v.w = fmaf(c0.w, gg.w, fmaf(c1.w, y.w, c2.w));
            }
            const bool row_ok = row < g.R;  // TMA zero-fills out-of-range elements, but the prologue may map 0 to non-zero
            if (!row_ok || kk + 0 >= g.K) v.x = 0.f;
            if (!row_ok || kk + 1 >= g.K) v.y = 0.f;
            if (!row_ok || kk + 2 >= g.K) v.z = 0.f;
            if (!row_ok || kk + 3 >= g.K) v.w = 0.f;
            store_split(a_hi, a_lo, swz(r, seg), v);
            if (PRO != PRO_BNRELU && g.dY_out != nullptr && row_ok && kk < g.K)
              *reinterpret_cast<float4 *>(g.dY_out + row * g.K + kk) = v;
          }
        }
        fence_async_proxy();
        mbar_arrive(&op_full[os]);
        mbar_arrive(&raw_empty[rs]);
        if ((tt & 127) == 0) S2C_PROBE(16 + 16 * (int)it + 3);
        if (++rs == RS) { rs = 0; rph ^= 1u; }
        if (++os == OS) { os = 0; oph ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> per-warp smem transpose -> coalesced global rows (+ statistics)
    // tcgen05.ld hands every thread one accumulator ROW; storing that directly would scatter each warp store over
    // 32 cache lines.  The 32x32 block goes through shared memory instead and leaves as 4 rows x 128 B per instruction;
    // the same (row, 16-byte segment) mapping loads the ReLU-mask operand and needs only two shuffles for the column sums.
    const int q = warp & 3;  // the TMEM lane quarter this warp may access
    const bool has_stats = g.stat_sum != nullptr;
    float *sE = s_epi + q * 32 * 36;
    const int er = lane >> 3, es = lane & 7;  // read-back mapping: rows er, er+4, ..., segment es
    float4 acc_s[N / 32], acc_q[N / 32];
#pragma unroll
    for (int i = 0; i < N / 32; ++i) { acc_s[i] = make_float4(0.f, 0.f, 0.f, 0.f); acc_q[i] = acc_s[i]; }
    // EPI_MASK_STATS: the ReLU-mask operand Yprev comes from global memory, 8 rows x 16 bytes per thread and column
    // block.  Loading each value right before its use serialises ~16 L2/HBM round trips per tile on the four epilogue
    // warps (ncu: 35 % of the kernel's stall samples sat on the first use of those loads, and 55 tiles x 16 x ~0.5 us
    // is the whole kernel time).  So the loads of column block cb+1 are issued before the math of block cb, and
    // block 0's before the wait for the accumulator.
    float4 ypre[8];
    auto load_y = [&](int cb, long long row_base) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long row = row_base + i * 4 + er;
        ypre[i] = row < g.R ? __ldg(reinterpret_cast<const float4 *>(g.Yprev + row * g.ldyp + cb * 32 + es * 4))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    for (long long t = 0; t < my_tiles; ++t) {
      const int ab = (int)(t & 1);
      const long long row_base = ((long long)blockIdx.x + t * gridDim.x) * BM + q * 32;
      if (EPI == EPI_MASK_STATS) load_y(0, row_base);
      mbar_wait(&acc_full[ab], (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      if (q == 0 && lane == 0) S2C_PROBE(16 + 16 * (int)t + 8);
      // column block cb + 1 is fetched from TMEM while block cb is transposed and stored (not in the 128-wide ReLU-mask
      // epilogue: with the mask operand's own prefetch registers it would spill)
      constexpr bool kPipeTmem = (EPI == EPI_STORE_STATS) || N == 64;
      float vbuf[kPipeTmem ? 2 : 1][32];
      if (kPipeTmem) tmem_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * N), vbuf[0]);
#pragma unroll
      for (int cb = 0; cb < N / 32; ++cb) {
        float *v = vbuf[kPipeTmem ? (cb & 1) : 0];
        if (kPipeTmem) {
          tmem_ld_wait(v);
          if (cb + 1 < N / 32)
            tmem_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * N + (cb + 1) * 32), vbuf[(cb + 1) & 1]);
        } else {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * N + cb * 32), v);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4 *>(sE + lane * 36 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        float4 e_sc = make_float4(1.f, 1.f, 1.f, 1.f), e_sh = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 ycur[8];
        if (EPI == EPI_MASK_STATS) {
          e_sc = __ldg(reinterpret_cast<const float4 *>(g.e_scale + cb * 32 + es * 4));
          e_sh = __ldg(reinterpret_cast<const float4 *>(g.e_shift + cb * 32 + es * 4));
#pragma unroll
          for (int i = 0; i < 8; ++i) ycur[i] = ypre[i];
          if (cb + 1 < N / 32) load_y(cb + 1, row_base);
        }
        float4 ps = make_float4(0.f, 0.f, 0.f, 0.f), pq = ps;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + er;
          const long long row = row_base + rr;
          float4 a = *reinterpret_cast<const float4 *>(sE + rr * 36 + es * 4);
          if (row < g.R) {
            if (EPI == EPI_MASK_STATS) {
              // gradient w.r.t. the previous layer's rectified output: mask by its ReLU, reduce sum(g) and sum(g*y)
              const float4 y = ycur[i];
              a.x = fmaf(y.x, e_sc.x, e_sh.x) > 0.f ? a.x : 0.f; a.y = fmaf(y.y, e_sc.y, e_sh.y) > 0.f ? a.y : 0.f;
              a.z = fmaf(y.z, e_sc.z, e_sh.z) > 0.f ? a.z : 0.f; a.w = fmaf(y.w, e_sc.w, e_sh.w) > 0.f ? a.w : 0.f;
              pq.x = fmaf(a.x, y.x, pq.x); pq.y = fmaf(a.y, y.y, pq.y); pq.z = fmaf(a.z, y.z, pq.z); pq.w = fmaf(a.w, y.w, pq.w);
            } else {
              pq.x = fmaf(a.x, a.x, pq.x); pq.y = fmaf(a.y, a.y, pq.y); pq.z = fmaf(a.z, a.z, pq.z); pq.w = fmaf(a.w, a.w, pq.w);
            }
            ps.x += a.x; ps.y += a.y; ps.z += a.z; ps.w += a.w;
            *reinterpret_cast<float4 *>(g.C + row * g.ldc + cb * 32 + es * 4) = a;
          }
        }
        if (has_stats) {
          acc_s[cb].x += ps.x; acc_s[cb].y += ps.y; acc_s[cb].z += ps.z; acc_s[cb].w += ps.w;
          acc_q[cb].x += pq.x; acc_q[cb].y += pq.y; acc_q[cb].z += pq.z; acc_q[cb].w += pq.w;
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[ab]);
      if (q == 0 && lane == 0) S2C_PROBE(16 + 16 * (int)t + 9);
    }
    if (has_stats) {
#pragma unroll
      for (int cb = 0; cb < N / 32; ++cb) {
        float4 a = acc_s[cb], b = acc_q[cb];
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {  // the four lanes that share a segment
          a.x += __shfl_xor_sync(0xffffffffu, a.x, o); a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
          a.z += __shfl_xor_sync(0xffffffffu, a.z, o); a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
          b.x += __shfl_xor_sync(0xffffffffu, b.x, o); b.y += __shfl_xor_sync(0xffffffffu, b.y, o);
          b.z += __shfl_xor_sync(0xffffffffu, b.z, o); b.w += __shfl_xor_sync(0xffffffffu, b.w, o);
        }
        if (lane < 8) {
          double *s1 = g.stat_sum + cb * 32 + lane * 4, *s2 = g.stat_sumsq + cb * 32 + lane * 4;
          atomicAdd(s1 + 0, (double)a.x); atomicAdd(s1 + 1, (double)a.y); atomicAdd(s1 + 2, (double)a.z); atomicAdd(s1 + 3, (double)a.w);
          atomicAdd(s2 + 0, (double)b.x); atomicAdd(s2 + 1, (double)b.y); atomicAdd(s2 + 2, (double)b.z); atomicAdd(s2 + 3, (double)b.w);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, TMEM_COLS);
}

struct PipeCfg { int OS, WS, RS, wres; size_t smem; };

unsigned long long *g_probe_buf = nullptr;  // s2c_mlp_probe(): device buffer the next launches stamp (null = off)
int g_probe_cap = 0;

size_t gemm2_smem(int N, int K, int OS, int WS, int RS, int raw_tiles, int npro, int GT) {
  const int KC = (K + BK - 1) / BK;
  return (size_t)OS * (2 * BM * BK * 4) + (size_t)WS * (2 * (size_t)N * BK * 4) +
         (size_t)RS * (raw_tiles * BM * BK * 4 + 2 * (size_t)GT * 128) + (size_t)npro * KC * BK * 4 + 4 * 32 * 36 * 4 + 44 * 8 + 16;
}

// Ring depths for one launch.  Measured with s2c_mlp_probe / tools/mlp_pipe_sweep.py (profiles/r02_mlp_pipe_probe.txt):
// the layer kernel is bound by SHARED-MEMORY BANDWIDTH, not by a latency the rings could hide -- per 128 x 32 chunk the
// transform warps move 48 KB (16 KB raw in, 32 KB hi/lo out), the twelve 3xTF32 MMAs read 12 x (4 KB of A + N x 32 B of
// B) and the epilogue stages every output tile through shared memory once more, ~1500 cycles of the SM's 128 B/clk in
// total (measured: 0.95 us per chunk at any depth of the weight ring, resident weights included).  What the depths do
// change is the number of raw bytes the TMA engine keeps in flight, which the HBM-bound backward kernels need: so the
// weight ring stays at two stages and every remaining byte goes to the raw ring.
PipeCfg choose_pipe(int N, int K, int raw_tiles, int npro, int GT, bool atm = false) {
  const size_t budget = 227 * 1024;
  const int KC = (K + BK - 1) / BK;
  if (atm) {
    // A operand in tensor memory (two stages there, one per group of transform warps): shared memory holds only weights and raw tiles.  Weights stay
    // resident when four raw stages still fit beside them, else three streamed stages; raw ring up to six deep.
    auto fits = [&](int WS, int RS) { return gemm2_smem(N, K, 0, WS, RS, raw_tiles, npro, GT) <= budget; };
    auto fill = [&](int WS) { int RS = 0; while (RS < 6 && fits(WS, RS + 1)) ++RS; return RS; };
    PipeCfg c = {2, 0, 0, 0, 0};
    if (KC <= kMaxWS && fill(KC) >= 4) { c.WS = KC; c.wres = 1; }
    else { c.WS = KC < 3 ? KC : 3; c.wres = KC <= 3 ? 1 : 0; }
    c.RS = fill(c.WS);
    if (c.RS < 1) c.RS = 1;
    c.smem = gemm2_smem(N, K, 0, c.WS, c.RS, raw_tiles, npro, GT);
    return c;
  }
  const int rs_cap = 4;
  auto fits = [&](int OS, int WS, int RS) { return gemm2_smem(N, K, OS, WS, RS, raw_tiles, npro, GT) <= budget; };
  auto fill_rs = [&](int OS, int WS) { int RS = 0; while (RS < rs_cap && fits(OS, WS, RS + 1)) ++RS; return RS; };
  const int ws0 = KC < 2 ? KC : 2;
  PipeCfg c = {2, ws0, fill_rs(2, ws0), KC <= 2 ? 1 : 0, 0};
  // tuning / debugging knobs (tools/mlp_pipe_sweep.py): S2C_MLP_OS, S2C_MLP_WS (0 = resident), S2C_MLP_RS
  const char *eo = getenv("S2C_MLP_OS"), *ew = getenv("S2C_MLP_WS"), *er = getenv("S2C_MLP_RS");
  if (eo || ew || er) {
    const PipeCfg chosen = c;
    if (eo && atoi(eo) >= 1 && atoi(eo) <= kMaxOS) c.OS = atoi(eo);
    if (ew) {
      const int w = atoi(ew);
      if (w == 0 && KC <= kMaxWS) { c.WS = KC; c.wres = 1; }
      else if (w >= 1 && w <= kMaxWS) { c.WS = w; c.wres = (w == KC) ? 1 : 0; if (w > KC) { c.WS = KC; c.wres = 1; } }
    }
    c.RS = fill_rs(c.OS, c.WS);
    if (er && atoi(er) >= 1 && atoi(er) <= kMaxRS && atoi(er) < c.RS) c.RS = atoi(er);
    if (c.RS < 1) c = chosen;  // the requested rings do not fit: keep the automatic choice
  }
  if (c.RS < 1) c.RS = 1;
  c.smem = gemm2_smem(N, K, c.OS, c.WS, c.RS, raw_tiles, npro, GT);
  return c;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {  // driver entry point through the runtime: no link-time dependency on libcuda
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// tensor map of a row-major (R, ld) matrix of 4-byte elements with K valid columns, box = [box_rows x 32 columns]
int make_tmap(CUtensorMap *tmap, const void *base, long long R, int K, long long ld, int box_rows = BM, bool i32 = false,
              bool swizzle128 = false) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) {
    set_error("mlp: cuTensorMapEncodeTiled is not available from this driver");
    return S2C_ERR_UNSUPPORTED;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)R};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tmap, i32 ? CU_TENSOR_MAP_DATA_TYPE_INT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("mlp: cuTensorMapEncodeTiled failed (%d) for R=%lld K=%d ld=%lld", (int)r, R, K, ld);
    return S2C_ERR_CUDA;
  }
  return S2C_OK;
}

template <int N, int PRO, int EPI, bool ATM = false>
int launch_gemm2(const Gemm2Args &g0, const float *A2, long long lda2, cudaStream_t st) {
  Gemm2Args g = g0;
  constexpr int raw_tiles = (PRO == PRO_AFFINE2) ? 2 : 1;
  constexpr int npro = (PRO == PRO_BNRELU) ? 2 : (PRO == PRO_AFFINE2) ? 3 : 5;
  CUtensorMap tmap, tmap2, tmap_am, tmap_dp;
  int rc = make_tmap(&tmap, g.A ? g.A : A2, g.R, g.K, g.A ? g.lda : lda2, BM, false, ATM);
  if (rc) return rc;
  rc = make_tmap(&tmap2, A2 ? A2 : g.A, g.R, g.K, A2 ? lda2 : g.lda);
  if (rc) return rc;
  g.GT = 0;
  if (PRO == PRO_POOL) {  // groups a 128-row tile can touch: 128/ns when ns divides 128 (tiles start on a group boundary)
    g.GT = (BM % g.ns == 0) ? BM / g.ns : (BM - 1) / g.ns + 2;
    const long long G = g.R / g.ns;
    rc = make_tmap(&tmap_am, g.argmax, G, g.K, g.K, g.GT, true);
    if (rc) return rc;
    rc = make_tmap(&tmap_dp, g.dpool, G, g.K, g.K, g.GT, false);
    if (rc) return rc;
  } else {
    tmap_am = tmap; tmap_dp = tmap;
  }
  const PipeCfg pc = choose_pipe(N, g.K, raw_tiles, npro, g.GT, ATM);
  const size_t smem = pc.smem;
  if (smem > 227 * 1024) {
    set_error("mlp: shared memory %zu B exceeds 227 KB (N=%d K=%d)", smem, N, g.K);
    return S2C_ERR_UNSUPPORTED;
  }
  g.RS = pc.RS; g.OS = pc.OS; g.WS = pc.WS; g.wres = pc.wres;
  g.probe = g_probe_buf; g.probe_cap = g_probe_cap;
  auto kern = mlp_gemm2_kernel<N, PRO, EPI, ATM>;
  S2C_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "mlp_gemm2 smem attr");
  const long long tiles = (g.R + BM - 1) / BM;
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  kern<<<grid, block_threads(ATM), smem, st>>>(g, tmap, tmap2, tmap_am, tmap_dp);
  S2C_CHECK_LAUNCH("mlp_gemm2 launch");
  return S2C_OK;
}

}  // namespace
}  // namespace s2c

// Profiling aid: CTA 0 of every following layer-kernel launch stamps clock64() at its pipeline hand-offs into `buf`
// (`capacity` 8-byte slots, device memory; layout at S2C_PROBE in this file).  buf = null switches it off.
extern "C" int s2c_mlp_probe(unsigned long long *buf, int capacity) {
  s2c::g_probe_buf = buf;
  s2c::g_probe_cap = buf ? capacity : 0;
  return S2C_OK;
}

// Pipelined layer kernel.  Extra requirements over s2c_mlp_layer_fwd: N in {64,128,256}; K, lda, ldc multiples of 4;
// A and C 16-byte aligned; `wprep` = caller-provided workspace of ceil(K/32) * N * 256 bytes (16-byte aligned).
extern "C" int s2c_mlp_layer_fwd_v2(const float *A, long long lda, long long R, int K, const float *pro_scale,
                                    const float *pro_shift, const float *W, int N, float *C, long long ldc,
                                    double *stat_sum, double *stat_sumsq, void *wprep, void *stream) {
  using namespace s2c;
  S2C_REQUIRE(R >= 0 && K >= 4, "mlp_layer_fwd_v2: bad sizes");
  S2C_REQUIRE(N == 64 || N == 128 || N == 256, "mlp_layer_fwd_v2: N=%d must be 64, 128 or 256", N);
  S2C_REQUIRE((K & 3) == 0 && (lda & 3) == 0 && (ldc & 3) == 0 && lda >= K && ldc >= N, "mlp_layer_fwd_v2: K, lda, ldc must be multiples of 4 (K=%d lda=%lld ldc=%lld)", K, lda, ldc);
  S2C_REQUIRE((pro_scale == nullptr) == (pro_shift == nullptr), "mlp_layer_fwd_v2: scale/shift must both be given or both null");
  S2C_REQUIRE((stat_sum == nullptr) == (stat_sumsq == nullptr), "mlp_layer_fwd_v2: stat pointers must both be given or both null");
  if (R == 0) return S2C_OK;
  S2C_REQUIRE(A && W && C && wprep, "mlp_layer_fwd_v2: null pointer");
  S2C_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)C & 15) == 0 && ((uintptr_t)wprep & 15) == 0, "mlp_layer_fwd_v2: A, C and wprep must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int KC = (K + BK - 1) / BK;
  if (N != 256) {
    w_prep_kernel<<<ceil_div(KC * N * 8, 256), 256, 0, st>>>(W, N, K, K, 1, (unsigned char *)wprep);
    S2C_CHECK_LAUNCH("w_prep");
  }
  Gemm2Args g = {};
  g.A = A; g.lda = lda; g.K = K; g.p0 = pro_scale; g.p1 = pro_shift; g.wprep = (const unsigned char *)wprep;
  g.C = C; g.ldc = ldc; g.stat_sum = stat_sum; g.stat_sumsq = stat_sumsq; g.R = R;
  // forward: A operand through tensor memory (S2C_MLP_ATM=0: the shared-memory operand path, for A/B measurements)
  const char *atm_env = getenv("S2C_MLP_ATM");
  const bool atm = !(atm_env && atoi(atm_env) == 0);
  if (N == 64) return atm ? launch_gemm2<64, PRO_BNRELU, EPI_STORE_STATS, true>(g, nullptr, 0, st)
                          : launch_gemm2<64, PRO_BNRELU, EPI_STORE_STATS>(g, nullptr, 0, st);
  if (N == 128) return atm ? launch_gemm2<128, PRO_BNRELU, EPI_STORE_STATS, true>(g, nullptr, 0, st)
                           : launch_gemm2<128, PRO_BNRELU, EPI_STORE_STATS>(g, nullptr, 0, st);
  // N = 256: two passes over 128 output channels each (the operand + staging footprint of a 256-wide tile does not
  // leave room for a double-buffered pipeline in 227 KB); w_prep wrote the chunks of all 256 rows, so prepare per half
  for (int h = 0; h < 2; ++h) {
    unsigned char *wp = (unsigned char *)wprep + (size_t)h * KC * 128 * 256;
    w_prep_kernel<<<ceil_div(KC * 128 * 8, 256), 256, 0, st>>>(W + (size_t)h * 128 * K, 128, K, K, 1, wp);
    S2C_CHECK_LAUNCH("w_prep");
    Gemm2Args gh = g;
    gh.wprep = wp; gh.C = C + h * 128;
    if (stat_sum) { gh.stat_sum = stat_sum + h * 128; gh.stat_sumsq = stat_sumsq + h * 128; }
    const int rc = atm ? launch_gemm2<128, PRO_BNRELU, EPI_STORE_STATS, true>(gh, nullptr, 0, st)
                       : launch_gemm2<128, PRO_BNRELU, EPI_STORE_STATS>(gh, nullptr, 0, st);
    if (rc) return rc;
  }
  return S2C_OK;
}

// Backward "data" kernel of one layer l of the shared MLP (tensor cores, same pipeline as the forward):
//     dY_l   = a[k]*g_l + b[k]*Y_l + c[k]          BatchNorm backward of layer l, affine per channel (a,b,c from the caller)
//     dX     = dY_l * W_l                           (R, N = C_{l-1})
//     g_prev = dX where relu(bn_{l-1}(Y_prev)) was active, else 0     -> C
//     stat_sum += sum_rows g_prev,  stat_sumsq += sum_rows g_prev * Y_prev     (float64, zeroed by the caller)
//   g_l is either dense (G, ldg) or -- for the LAST layer -- rebuilt on the fly from the pooled gradient:
//   dpool/argmax (R/ns, K) with last_scale/last_shift = the last layer's folded BatchNorm (ReLU mask of layer l).
//   W is layer l's Conv2d weight (K = C_l rows, N = C_{l-1} columns, row-major).  N in {64,128,256}; K multiple of 4.
//   dY_out (R, K), optional: dY_l written back for the weight-gradient GEMM  dW_l = dY_l^T * relu(bn(Y_prev)).
extern "C" int s2c_mlp_layer_bwd_data(const float *G, long long ldg, const float *Y, long long ldy, long long R, int K,
                                      const float *a, const float *b, const float *c, const float *dpool,
                                      const int *argmax, int ns, const float *last_scale, const float *last_shift,
                                      const float *W, int N, const float *Yprev, long long ldyp, const float *prev_scale,
                                      const float *prev_shift, float *C, long long ldc, float *dY_out, double *stat_sum,
                                      double *stat_sumsq, void *wprep, void *stream) {
  using namespace s2c;
  S2C_REQUIRE(R >= 0 && K >= 4 && (K & 3) == 0, "mlp_layer_bwd_data: K=%d must be a positive multiple of 4", K);
  S2C_REQUIRE(N == 64 || N == 128 || N == 256, "mlp_layer_bwd_data: N=%d must be 64, 128 or 256", N);
  S2C_REQUIRE((ldy & 3) == 0 && (ldyp & 3) == 0 && (ldc & 3) == 0 && ldy >= K && ldyp >= N && ldc >= N, "mlp_layer_bwd_data: bad leading dimensions");
  if (R == 0) return S2C_OK;
  const bool pool = dpool != nullptr;
  S2C_REQUIRE(pool || (G && (ldg & 3) == 0 && ldg >= K), "mlp_layer_bwd_data: need either a dense gradient or dpool/argmax");
  S2C_REQUIRE(!pool || (argmax && ns >= 1 && last_scale && last_shift && R % ns == 0), "mlp_layer_bwd_data: incomplete pooled-gradient arguments");
  S2C_REQUIRE(Y && a && b && c && W && Yprev && prev_scale && prev_shift && C && wprep && stat_sum && stat_sumsq, "mlp_layer_bwd_data: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int KC = (K + BK - 1) / BK;
  // B operand = W_l^T: B[n][k] = W[k][n]
  if (N != 256) {
    w_prep_kernel<<<ceil_div(KC * N * 8, 256), 256, 0, st>>>(W, N, K, N, 0, (unsigned char *)wprep);
    S2C_CHECK_LAUNCH("w_prep");
  }
  Gemm2Args g = {};
  g.K = K; g.p0 = a; g.p1 = b; g.p2 = c; g.wprep = (const unsigned char *)wprep; g.C = C; g.ldc = ldc; g.dY_out = dY_out;
  g.Yprev = Yprev; g.ldyp = ldyp; g.e_scale = prev_scale; g.e_shift = prev_shift;
  g.stat_sum = stat_sum; g.stat_sumsq = stat_sumsq; g.R = R;
  if (pool) {
    g.A = nullptr; g.lda = 0; g.p3 = last_scale; g.p4 = last_shift; g.dpool = dpool; g.argmax = argmax; g.ns = ns;
  } else {
    g.A = G; g.lda = ldg;
  }
  const int halves = N == 256 ? 2 : 1, NH = N / halves;   // 256 output channels: two 128-wide passes (see the forward)
  for (int h = 0; h < halves; ++h) {
    Gemm2Args gh = g;
    if (halves == 2) {
      unsigned char *wp = (unsigned char *)wprep + (size_t)h * KC * 128 * 256;
      w_prep_kernel<<<ceil_div(KC * 128 * 8, 256), 256, 0, st>>>(W + h * 128, 128, K, N, 0, wp);
      S2C_CHECK_LAUNCH("w_prep");
      gh.wprep = wp; gh.C = C + h * 128; gh.Yprev = Yprev + h * 128; gh.e_scale = prev_scale + h * 128; gh.e_shift = prev_shift + h * 128;
      gh.stat_sum = stat_sum + h * 128; gh.stat_sumsq = stat_sumsq + h * 128;
      if (h == 1) gh.dY_out = nullptr;  // dY does not depend on the output half: written once
    }
    int rc;
    if (pool) rc = NH == 64 ? launch_gemm2<64, PRO_POOL, EPI_MASK_STATS>(gh, Y, ldy, st) : launch_gemm2<128, PRO_POOL, EPI_MASK_STATS>(gh, Y, ldy, st);
    else rc = NH == 64 ? launch_gemm2<64, PRO_AFFINE2, EPI_MASK_STATS>(gh, Y, ldy, st) : launch_gemm2<128, PRO_AFFINE2, EPI_MASK_STATS>(gh, Y, ldy, st);
    if (rc) return rc;
  }
  return S2C_OK;
}

// Input gradient of the FIRST layer of a shared MLP (no previous BatchNorm / ReLU to mask with, no statistics):
//     dY = a[k]*G + b[k]*Y + c[k]            BatchNorm backward of layer 0 (affine per channel)
//     C  = dY * W[:, w0 : w0+N]              (R, N) written with leading dimension ldc
//   W is the layer's Conv2d weight (K rows) with row stride ldw; the caller offsets W (and C) to select a block of N
//   input columns -- e.g. the feature block of the grouped rows [xyz, 0, features], or one half of a concatenation.
//   N in {64,128,256}.  dY_out (R, K), optional: dY written back for the weight-gradient GEMM.
//   Replaces cuDNN's dgrad of the first 1x1 convolution (pytorch_utils.py:88-95 under autograd).
extern "C" int s2c_mlp_layer_bwd_input(const float *G, long long ldg, const float *Y, long long ldy, long long R, int K,
                                       const float *a, const float *b, const float *c, const float *W, long long ldw,
                                       int N, float *C, long long ldc, float *dY_out, void *wprep, void *stream) {
  using namespace s2c;
  S2C_REQUIRE(R >= 0 && K >= 4 && (K & 3) == 0, "mlp_layer_bwd_input: K=%d must be a positive multiple of 4", K);
  S2C_REQUIRE(N == 64 || N == 128 || N == 256, "mlp_layer_bwd_input: N=%d must be 64, 128 or 256", N);
  S2C_REQUIRE((ldg & 3) == 0 && (ldy & 3) == 0 && (ldc & 3) == 0 && ldg >= K && ldy >= K && ldc >= N && ldw >= N,
              "mlp_layer_bwd_input: bad leading dimensions");
  if (R == 0) return S2C_OK;
  S2C_REQUIRE(G && Y && a && b && c && W && C && wprep, "mlp_layer_bwd_input: null pointer");
  S2C_REQUIRE(((uintptr_t)G & 15) == 0 && ((uintptr_t)Y & 15) == 0 && ((uintptr_t)C & 15) == 0 && ((uintptr_t)wprep & 15) == 0,
              "mlp_layer_bwd_input: G, Y, C and wprep must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int KC = (K + BK - 1) / BK;
  Gemm2Args g = {};
  g.A = G; g.lda = ldg; g.K = K; g.p0 = a; g.p1 = b; g.p2 = c; g.C = C; g.ldc = ldc; g.dY_out = dY_out; g.R = R;
  const int halves = N == 256 ? 2 : 1, NH = N / halves;
  for (int h = 0; h < halves; ++h) {
    Gemm2Args gh = g;
    unsigned char *wp = (unsigned char *)wprep + (size_t)h * KC * NH * 256;
    // B operand = (W block)^T: B[n][k] = W[k*ldw + n]
    w_prep_kernel<<<ceil_div(KC * NH * 8, 256), 256, 0, st>>>(W + h * NH, NH, K, (int)ldw, 0, wp);
    S2C_CHECK_LAUNCH("w_prep");
    gh.wprep = wp; gh.C = C + h * NH;
    if (h == 1) gh.dY_out = nullptr;
    const int rc = NH == 64 ? launch_gemm2<64, PRO_AFFINE2, EPI_STORE_STATS>(gh, Y, ldy, st)
                            : launch_gemm2<128, PRO_AFFINE2, EPI_STORE_STATS>(gh, Y, ldy, st);
    if (rc) return rc;
  }
  return S2C_OK;
}
