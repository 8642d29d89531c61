// Furthest point sampling for sm_100a.
//
// Replaces furthest_point_sampling_kernel (reference lib/pointnet2/_ext_src/src/sampling_gpu.cu:69-173,
// host wrapper :175-229, binding sampling.cpp:66-87) and, optionally, the gather_points call that
// follows it in PointnetSAModuleVotes.forward (pointnet2_modules.py:238-240).
//
// (Measured on B200, 8 scenes x 40 000 -> 2048 points: reference kernel 27.7 ms, this kernel 1.98 ms.)
// The reference runs ONE 512-thread block per scene and, for each of the m-1 sequential picks,
// re-reads all N points and the (B,N) scratch `temp` from L1/L2 and walks a 9-level
// __syncthreads tree.  FPS is a latency-bound serial chain, so the design here minimises the
// latency of one pick:
//   * one thread-block CLUSTER per scene (1..16 CTAs x 1024 threads); every point and its running
//     min-distance live in REGISTERS for the whole kernel (no scratch buffer, no per-pick memory traffic);
//   * per pick: PPT fused distance updates per thread, a two-instruction warp arg-max
//     (redux.sync max on the distance bits, redux.sync min on the tie-break key), a CTA-level
//     reduction through shared memory (one __syncthreads), then ONE 20-byte record per CTA pushed
//     into the shared memory of every CTA of the cluster with st.async (DSMEM write that signals the
//     destination's mbarrier with its byte count) -- no cluster-wide barrier: each CTA just waits on
//     its own mbarrier for CL x 20 bytes, and every warp reduces the CL records redundantly;
//     (DSMEM moves only ~20 B/cycle/SM, so pushing per-warp records would make the exchange the bottleneck)
//   * the winner's coordinates travel with the record, so the next pick needs no global load.
//
// Bit-exactness.  The reference's result is  argmax_k temp[k]  where ties are resolved by its
// reduction tree: within a thread the smallest k wins (strict '>'), and between threads the
// shared-memory tree keeps the lower slot at each level, which makes the winner the tied thread
// with the smallest BIT-REVERSED thread id (tid = k mod bs, bs = opt_n_threads(N), cuda_utils.h:15-19).
// We reproduce this for any thread->point mapping by reducing the pair
//     (distance bits,  key = bitrev_log2bs(k mod bs) << 23 | k / bs)      max distance, then min key.
// Threads with no admissible point contribute (-1, index 0) exactly like the reference.
#include "s2c_common.cuh"

namespace s2c {
namespace {

constexpr uint32_t kNoKey = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(a), "r"(parity)
      : "memory");
}
// DSMEM store that completes `bytes` on the destination CTA's mbarrier
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, uint32_t remote_bar, uint32_t a, uint32_t b, uint32_t c,
                                            uint32_t d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(remote_addr),
               "r"(a), "r"(b), "r"(c), "r"(d), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t remote_addr, uint32_t remote_bar, uint32_t a) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr), "r"(a),
               "r"(remote_bar)
               : "memory");
}

// PPT: points per thread (registers), T: threads per CTA, CL: CTAs per cluster (1 = plain launch)
template <int PPT, int T, int CL>
__global__ void __launch_bounds__(T, 1)
fps_kernel(const float *__restrict__ xyz, int N, int m, int log2bs, int *__restrict__ idx,
           float *__restrict__ new_xyz) {
  constexpr int W = T / 32;
  __shared__ uint4 wrec[2][32];   // per-warp records of this CTA  {dist bits ^ 0x80000000, key, x bits, y bits}
  __shared__ float wrecz[2][32];
  __shared__ uint4 crec[2][16];   // per-CTA records of the whole cluster (written remotely with st.async)
  __shared__ float crecz[2][16];
  __shared__ __align__(8) uint64_t full[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = (CL > 1) ? cluster_ctarank() : 0u;
  const int scene = blockIdx.x / CL;
  xyz += (size_t)scene * N * 3;
  idx += (size_t)scene * m;
  if (new_xyz) new_xyz += (size_t)scene * m * 3;

  // slot i of this thread holds point k = i*(CL*T) + rank*T + tid ; (CL*T) % bs == 0, so every slot of a
  // thread has the same k mod bs and "smallest slot wins" == the reference's per-thread "smallest k wins".
  const int k0 = (int)rank * T + tid;
  float px[PPT], py[PPT], pz[PPT], tmp[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int k = i * (CL * T) + k0;
    px[i] = py[i] = pz[i] = 0.f;
    tmp[i] = -1.f;  // inadmissible: fminf(d, -1) = -1 is never '>' the running best (-1)
    if (k < N) {
      px[i] = xyz[(size_t)k * 3 + 0];
      py[i] = xyz[(size_t)k * 3 + 1];
      pz[i] = xyz[(size_t)k * 3 + 2];
      // sampling_gpu.cu:100-101: "if (mag <= 1e-3) continue;" compares in double
      if (!((double)sqnorm3(px[i], py[i], pz[i]) <= 1e-3)) tmp[i] = 1e10f;  // sampling.cpp:74-76
    }
  }
  const float x0 = xyz[0], y0 = xyz[1], z0 = xyz[2];
  const uint32_t bsm1 = (1u << log2bs) - 1u;
  const uint32_t rev = log2bs ? (__brev((uint32_t)k0 & bsm1) >> (32 - log2bs)) : 0u;

  if (rank == 0 && tid == 0) {
    idx[0] = 0;
    if (new_xyz) { new_xyz[0] = x0; new_xyz[1] = y0; new_xyz[2] = z0; }
  }
  if (CL > 1) {
    if (tid == 0) {
      mbar_init(&full[0], 1);
      mbar_init(&full[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_sync_all();  // barriers initialised and every CTA resident before the first DSMEM store
  }

  float x1 = x0, y1 = y0, z1 = z0;
  for (int j = 1; j < m; ++j) {
    const int buf = j & 1;
    float best = -1.f;
    int bi = 0;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const float d = sqdist3(px[i], py[i], pz[i], x1, y1, z1);
      const float d2 = fminf(d, tmp[i]);
      tmp[i] = d2;
      if (d2 > best) { best = d2; bi = i; }
    }
    // ---- stage 1: warp arg-max ------------------------------------------------------------
    const int dbits = __float_as_int(best);  // best >= 0 (monotone as int) or -1.0f (negative int)
    const int wmax = __reduce_max_sync(0xffffffffu, dbits);
    const uint32_t k = (uint32_t)(bi * (CL * T) + k0);
    const uint32_t key = (dbits == wmax && dbits >= 0) ? ((rev << 23) | (k >> log2bs)) : kNoKey;
    const uint32_t wkey = __reduce_min_sync(0xffffffffu, key);
    if (key == wkey && (lane == 0 || wkey != kNoKey)) {  // the unique winner lane (lane 0 if the warp has none)
      float wx = 0.f, wy = 0.f, wz = 0.f;
#pragma unroll
      for (int i = 0; i < PPT; ++i)
        if (i == bi) { wx = px[i]; wy = py[i]; wz = pz[i]; }
      // unsigned-ordered distance; "no candidate" (-1.0f) sorts below every candidate
      wrec[buf][warp] = make_uint4((uint32_t)wmax ^ 0x80000000u, wkey, __float_as_uint(wx), __float_as_uint(wy));
      wrecz[buf][warp] = wz;
    }
    __syncthreads();
    uint32_t rh = 0u, rk = kNoKey, rx = 0u, ry = 0u;
    float rz = 0.f;
    if (CL > 1) {
      // ---- stage 2: warp 0 reduces the CTA's W records and pushes ONE record to every CTA of the cluster
      if (warp == 0) {
        uint4 r = make_uint4(0u, kNoKey, 0u, 0u);
        float z = 0.f;
        if (lane < W) { r = wrec[buf][lane]; z = wrecz[buf][lane]; }
        const uint32_t gh = __reduce_max_sync(0xffffffffu, r.x);
        const uint32_t gk = __reduce_min_sync(0xffffffffu, r.x == gh ? r.y : kNoKey);
        const unsigned w2 = __ballot_sync(0xffffffffu, r.x == gh && r.y == gk);
        const int src = __ffs(w2) - 1;
        const uint32_t sx = __shfl_sync(0xffffffffu, r.z, src), sy = __shfl_sync(0xffffffffu, r.w, src);
        const float sz = __shfl_sync(0xffffffffu, z, src);
        if (lane == 0) mbar_arrive_expect_tx(&full[buf], CL * 20);
        if (lane < CL) {
          const uint32_t bar = map_to_cta((uint32_t)__cvta_generic_to_shared(&full[buf]), (uint32_t)lane);
          st_async_v4(map_to_cta((uint32_t)__cvta_generic_to_shared(&crec[buf][rank]), (uint32_t)lane), bar, gh, gk, sx, sy);
          st_async_b32(map_to_cta((uint32_t)__cvta_generic_to_shared(&crecz[buf][rank]), (uint32_t)lane), bar,
                       __float_as_uint(sz));
        }
      }
      // ---- stage 3: wait for the CL records of this pick, every warp reduces them redundantly
      mbar_wait(&full[buf], (uint32_t)(((j - 1) >> 1) & 1));  // buffer `buf` is on its ((j-1)/2)-th use
      if (lane < CL) {
        const uint4 r = crec[buf][lane];
        rh = r.x; rk = r.y; rx = r.z; ry = r.w;
        rz = crecz[buf][lane];
      }
    } else {
      if (lane < W) {
        const uint4 r = wrec[buf][lane];
        rh = r.x; rk = r.y; rx = r.z; ry = r.w;
        rz = wrecz[buf][lane];
      }
    }
    const uint32_t gh = __reduce_max_sync(0xffffffffu, rh);
    const uint32_t gk = __reduce_min_sync(0xffffffffu, rh == gh ? rk : kNoKey);
    int old = 0;
    if (gh >= 0x80000000u) {  // at least one admissible point in the scene
      const unsigned w2 = __ballot_sync(0xffffffffu, rh == gh && rk == gk);
      const int src = __ffs(w2) - 1;
      x1 = __uint_as_float(__shfl_sync(0xffffffffu, rx, src));
      y1 = __uint_as_float(__shfl_sync(0xffffffffu, ry, src));
      z1 = __shfl_sync(0xffffffffu, rz, src);
      old = (int)(((gk & 0x7FFFFFu) << log2bs) | (log2bs ? (__brev(gk >> 23) >> (32 - log2bs)) : 0u));
    } else {  // reference: every thread reports (-1, 0) -> index 0
      x1 = x0; y1 = y0; z1 = z0;
    }
    if (rank == 0 && tid == 0) {
      idx[j] = old;
      if (new_xyz) { new_xyz[j * 3 + 0] = x1; new_xyz[j * 3 + 1] = y1; new_xyz[j * 3 + 2] = z1; }
    }
  }
}

template <int PPT, int T, int CL>
int launch_fps(const float *xyz, int B, int N, int m, int log2bs, int *idx, float *new_xyz, cudaStream_t st) {
  auto kern = fps_kernel<PPT, T, CL>;
  if (CL > 8) {
    S2C_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1),
             "fps: allow 16-CTA clusters");
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * CL));
  cfg.blockDim = dim3(T);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (CL > 1) ? 1 : 0;
  S2C_CUDA(cudaLaunchKernelEx(&cfg, kern, xyz, N, m, log2bs, idx, new_xyz), "fps launch");
  return S2C_OK;
}

int host_opt_n_threads(int work) {  // cuda_utils.h:15-19, integer form (exact for every int)
  int p = 0;
  while ((2 << p) <= work) ++p;
  int v = 1 << p;
  return v > 512 ? 512 : (v < 1 ? 1 : v);
}

}  // namespace
}  // namespace s2c

// Force a cluster size (0 = automatic).  Exposed for tuning/tests through the environment
// variable S2C_FPS_CLUSTER read once per process.
static int fps_cluster_override() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("S2C_FPS_CLUSTER");
    v = e ? atoi(e) : 0;
  }
  return v;
}

extern "C" int s2c_furthest_point_sampling(const float *xyz, int B, int N, int m, int *idx, float *new_xyz,
                                           void *stream) {
  using namespace s2c;
  S2C_REQUIRE(B >= 0 && N >= 1 && m >= 0, "furthest_point_sampling: bad sizes B=%d N=%d m=%d", B, N, m);
  if (B == 0 || m == 0) return S2C_OK;
  S2C_REQUIRE(xyz && idx, "furthest_point_sampling: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int bs = host_opt_n_threads(N);
  int log2bs = 0;
  while ((1 << log2bs) < bs) ++log2bs;

#define S2C_FPS(PPT, T, CL) return launch_fps<PPT, T, CL>(xyz, B, N, m, log2bs, idx, new_xyz, st)
  int cl = fps_cluster_override();
  if (cl == 0) {
    // automatic: single CTA while the scene fits 8 points/thread, else the smallest cluster that
    // keeps <= 6 points per thread (<= 8 CTAs, portable), else 16 CTAs.
    if (N <= 8192) cl = 1;
    else if (N <= 2 * 1024 * 6) cl = 2;
    else if (N <= 4 * 1024 * 6) cl = 4;
    else if (N <= 8 * 1024 * 6) cl = 8;
    else cl = 16;
  }
  if (cl == 1) {
    if (N <= 512) S2C_FPS(1, 512, 1);
    if (N <= 1024) S2C_FPS(2, 512, 1);
    if (N <= 2048) S2C_FPS(4, 512, 1);
    if (N <= 4096) S2C_FPS(8, 512, 1);
    if (N <= 8192) S2C_FPS(8, 1024, 1);
    if (N <= 12288) S2C_FPS(12, 1024, 1);
    cl = 2;
  }
  static int threads_override = -1;
  if (threads_override < 0) {
    const char *e = getenv("S2C_FPS_THREADS");
    threads_override = e ? atoi(e) : 512;
  }
  // 512-thread CTAs (16 warps, up to 20 points per thread) are the default for clusters: measured 1.98 ms vs
  // 2.28 ms with 1024 threads for 8 scenes x (40 000 -> 2048) on B200 (profiles/r01_fps_tuning.txt);
  // S2C_FPS_THREADS=1024 selects the other variant.
  if (threads_override == 512) {
    const int p5 = ceil_div(N, cl * 512);
#define S2C_FPS_CL512(CL)                  \
  if (cl == CL) {                          \
    if (p5 <= 2) S2C_FPS(2, 512, CL);      \
    if (p5 <= 4) S2C_FPS(4, 512, CL);      \
    if (p5 <= 6) S2C_FPS(6, 512, CL);      \
    if (p5 <= 8) S2C_FPS(8, 512, CL);      \
    if (p5 <= 10) S2C_FPS(10, 512, CL);    \
    if (p5 <= 12) S2C_FPS(12, 512, CL);    \
    if (p5 <= 16) S2C_FPS(16, 512, CL);    \
    if (p5 <= 20) S2C_FPS(20, 512, CL);    \
  }
    S2C_FPS_CL512(2)
    S2C_FPS_CL512(4)
    S2C_FPS_CL512(8)
    S2C_FPS_CL512(16)
#undef S2C_FPS_CL512
  }
  const int ppt = ceil_div(N, cl * 1024);
#define S2C_FPS_CL(CL)                      \
  if (cl == CL) {                           \
    if (ppt <= 1) S2C_FPS(1, 1024, CL);     \
    if (ppt <= 2) S2C_FPS(2, 1024, CL);     \
    if (ppt <= 3) S2C_FPS(3, 1024, CL);     \
    if (ppt <= 4) S2C_FPS(4, 1024, CL);     \
    if (ppt <= 5) S2C_FPS(5, 1024, CL);     \
    if (ppt <= 6) S2C_FPS(6, 1024, CL);     \
    if (ppt <= 8) S2C_FPS(8, 1024, CL);     \
    if (ppt <= 10) S2C_FPS(10, 1024, CL);   \
    if (ppt <= 12) S2C_FPS(12, 1024, CL);   \
    if (CL == 16 && ppt <= 13) S2C_FPS(13, 1024, CL); \
    if (CL == 16 && ppt <= 16) S2C_FPS(16, 1024, CL); \
  }
  S2C_FPS_CL(2)
  S2C_FPS_CL(4)
  S2C_FPS_CL(8)
  S2C_FPS_CL(16)
  if (cl < 16) {  // forced small cluster that cannot hold the scene: fall through to the largest
    const int p16 = ceil_div(N, 16 * 1024);
    if (p16 <= 12) {
      if (p16 <= 2) S2C_FPS(2, 1024, 16);
      if (p16 <= 4) S2C_FPS(4, 1024, 16);
      if (p16 <= 6) S2C_FPS(6, 1024, 16);
      if (p16 <= 8) S2C_FPS(8, 1024, 16);
      S2C_FPS(12, 1024, 16);
    }
  }
  set_error("furthest_point_sampling: N=%d exceeds the register-resident capacity (262144 points/scene)", N);
  return S2C_ERR_UNSUPPORTED;
#undef S2C_FPS
#undef S2C_FPS_CL
}
