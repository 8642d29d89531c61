// Ball query and fused query+group for sm_100a.
//
// Replaces query_ball_point_kernel (reference lib/pointnet2/_ext_src/src/ball_query_gpu.cu:9-44,
// binding ball_query.cpp:8-32) and, in the fused entry point, the whole of QueryAndGroup.forward
// (lib/pointnet2/pointnet2_utils.py:317-376): ball_query -> group_points(xyz^T) - new_xyz ->
// (/radius) -> group_points(features) -> cat.
//
// Reference: one block per scene, each THREAD scans all n points serially for its centres.
// Here: one WARP scans for CPW centres at once over point tiles staged in shared memory as SoA
// (each coordinate is read once per warp per CPW centres, conflict-free), hits are compacted in
// index order with ballot + popc so the "first nsample in index order" rule is kept bit-exactly,
// and a warp stops as soon as all its centres are full.  The neighbour list stays in shared memory
// and the gather (+ centre subtraction, 1/radius scaling, channel concat) is done by the same warp
// with coalesced stores -- the (B,M,ns) index tensor and the two grouped tensors of the reference
// never make a round trip through HBM.
#include "s2c_common.cuh"
#include "group_epilogue.cuh"

namespace s2c {
namespace {

constexpr int kWarps = 8;        // warps per CTA
                                 // centres per warp: template parameter CPW (4 when the grid is large enough, else 2 / 1)
constexpr int kTile = 2048;      // points per shared-memory tile

template <bool GROUP, int kCPW>
__global__ void __launch_bounds__(kWarps * 32, 4)
ball_query_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz, int n, int M, float radius,
                  int nsample, int *__restrict__ idx, int *__restrict__ cnt_out, GroupArgs ga) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *sx = reinterpret_cast<float *>(smem_raw);
  float *sy = sx + kTile;
  float *sz = sy + kTile;
  constexpr int kCentresPerCta = kWarps * kCPW;
  int *sidx = reinterpret_cast<int *>(sz + kTile);  // [kCentresPerCta][nsample]

  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * kCentresPerCta + warp * kCPW;
  xyz += (size_t)b * n * 3;
  new_xyz += (size_t)b * M * 3;

  const float radius2 = __fmul_rn(radius, radius);  // ball_query_gpu.cu:22 (fp32)
  float cx[kCPW], cy[kCPW], cz[kCPW];
  int cnt[kCPW];
#pragma unroll
  for (int q = 0; q < kCPW; ++q) {
    const int j = c0 + q;
    const bool ok = j < M;
    cx[q] = ok ? new_xyz[j * 3 + 0] : 0.f;
    cy[q] = ok ? new_xyz[j * 3 + 1] : 0.f;
    cz[q] = ok ? new_xyz[j * 3 + 2] : 0.f;
    cnt[q] = ok ? 0 : nsample;  // out-of-range centres are "full" from the start
  }
  int *my_idx = sidx + (size_t)(warp * kCPW) * nsample;

  bool warp_done = true;
#pragma unroll
  for (int q = 0; q < kCPW; ++q) warp_done = warp_done && (cnt[q] >= nsample);

  for (int t0 = 0; t0 < n; t0 += kTile) {
    if (__syncthreads_and(warp_done)) break;  // also protects the tile against early overwrite
    const int tn = min(kTile, n - t0);
    for (int i = threadIdx.x; i < tn * 3; i += kWarps * 32) {
      const float v = xyz[(size_t)t0 * 3 + i];
      const int p = i / 3, c = i - p * 3;
      (c == 0 ? sx : (c == 1 ? sy : sz))[p] = v;
    }
    __syncthreads();
    if (!warp_done) {
      for (int base = 0; base < tn; base += 32) {
        const int p = base + lane;
        const bool in = p < tn;
        const float x = in ? sx[p] : 0.f, y = in ? sy[p] : 0.f, z = in ? sz[p] : 0.f;
        bool all_full = true;
#pragma unroll
        for (int q = 0; q < kCPW; ++q) {
          const float d2 = sqdist3(cx[q], cy[q], cz[q], x, y, z);  // (new - p), ball_query_gpu.cu:30-31
          const bool hit = in && (d2 < radius2) && (cnt[q] < nsample);
          const unsigned mask = __ballot_sync(0xffffffffu, hit);
          if (mask) {
            const int pos = cnt[q] + __popc(mask & ((1u << lane) - 1u));
            if (hit && pos < nsample) my_idx[q * nsample + pos] = t0 + p;
            cnt[q] = min(nsample, cnt[q] + __popc(mask));
          }
          all_full = all_full && (cnt[q] >= nsample);
        }
        if (all_full) { warp_done = true; break; }
      }
    }
  }
  __syncwarp();

  // ---- epilogue: pad (first hit, or 0 for an empty ball), write idx, gather ---------------
#pragma unroll 1
  for (int q = 0; q < kCPW; ++q) {
    const int j = c0 + q;
    if (j >= M) break;
    const int c = cnt[q];
    int *li = my_idx + q * nsample;
    const int first = c > 0 ? li[0] : 0;
    __syncwarp();
    for (int s = c + lane; s < nsample; s += 32) li[s] = first;
    __syncwarp();
    if (idx) {
      int *o = idx + ((size_t)b * M + j) * nsample;
      for (int s = lane; s < nsample; s += 32) o[s] = li[s];
    }
    if (cnt_out && lane == 0) cnt_out[(size_t)b * M + j] = c;
    if (GROUP) {
      const float *f = ga.features ? ga.features + (size_t)b * ga.feat_scene_stride : nullptr;
      group_epilogue(ga, xyz, f, li, nsample, lane, cx[q], cy[q], cz[q], b, M, j);
    }
  }
}

template <bool GROUP, int CPW>
int launch_cpw(const float *new_xyz, const float *xyz, int B, int n, int M, float radius, int nsample, int *idx,
               int *cnt, const GroupArgs &ga, cudaStream_t st) {
  const size_t smem = (size_t)3 * kTile * sizeof(float) + (size_t)kWarps * CPW * nsample * sizeof(int);
  auto kern = ball_query_kernel<GROUP, CPW>;
  S2C_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ball_query smem attr");
  dim3 grid((unsigned)ceil_div(M, kWarps * CPW), (unsigned)B);
  kern<<<grid, kWarps * 32, smem, st>>>(new_xyz, xyz, n, M, radius, nsample, idx, cnt, ga);
  S2C_CHECK_LAUNCH("ball_query launch");
  return S2C_OK;
}

template <bool GROUP>
int launch(const float *new_xyz, const float *xyz, int B, int n, int M, float radius, int nsample, int *idx,
           int *cnt, const GroupArgs &ga, cudaStream_t st) {
  // 4 centres per warp amortise the shared-memory reads of the scan; with few centres (SA3/SA4, vote aggregation)
  // fewer centres per warp keep at least ~2 CTAs per SM busy in the store-bound gather epilogue
  const long long want = 2LL * kNumSMs;
  if ((long long)ceil_div(M, kWarps * 4) * B >= want) return launch_cpw<GROUP, 4>(new_xyz, xyz, B, n, M, radius, nsample, idx, cnt, ga, st);
  if ((long long)ceil_div(M, kWarps * 2) * B >= want) return launch_cpw<GROUP, 2>(new_xyz, xyz, B, n, M, radius, nsample, idx, cnt, ga, st);
  return launch_cpw<GROUP, 1>(new_xyz, xyz, B, n, M, radius, nsample, idx, cnt, ga, st);
}

}  // namespace
}  // namespace s2c

extern "C" int s2c_ball_query(const float *new_xyz, const float *xyz, int B, int n, int M, float radius,
                              int nsample, int *idx, int *cnt, void *stream) {
  using namespace s2c;
  S2C_REQUIRE(B >= 0 && n >= 0 && M >= 0, "ball_query: bad sizes B=%d n=%d M=%d", B, n, M);
  S2C_REQUIRE(nsample >= 1 && nsample <= 1024, "ball_query: nsample=%d outside [1,1024]", nsample);
  if (B == 0 || M == 0) return S2C_OK;
  S2C_REQUIRE(new_xyz && idx && (xyz || n == 0), "ball_query: null pointer");
  S2C_REQUIRE(B <= 65535, "ball_query: B=%d > 65535", B);
  GroupArgs ga = {};
  return launch<false>(new_xyz, xyz, B, n, M, radius, nsample, idx, cnt, ga, (cudaStream_t)stream);
}

extern "C" int s2c_query_and_group(const float *xyz, const float *new_xyz, const float *features, int B, int n,
                                   int M, int C, int feat_layout, long long feat_stride, float radius,
                                   int nsample, int normalize_xyz, int out_layout, int *idx, float *grouped,
                                   void *stream) {
  using namespace s2c;
  S2C_REQUIRE(B >= 0 && n >= 1 && M >= 0 && C >= 0, "query_and_group: bad sizes B=%d n=%d M=%d C=%d", B, n, M, C);
  S2C_REQUIRE(nsample >= 1 && nsample <= 1024, "query_and_group: nsample=%d outside [1,1024]", nsample);
  S2C_REQUIRE(feat_layout == 0 || feat_layout == 1, "query_and_group: feat_layout must be 0 or 1");
  S2C_REQUIRE(out_layout >= 0 && out_layout <= 2, "query_and_group: out_layout must be 0, 1 or 2");
  if (B == 0 || M == 0) return S2C_OK;
  S2C_REQUIRE(xyz && new_xyz && grouped, "query_and_group: null pointer");
  S2C_REQUIRE(C == 0 || features, "query_and_group: features is null but C=%d", C);
  S2C_REQUIRE(B <= 65535, "query_and_group: B=%d > 65535", B);
  GroupArgs ga = {};
  ga.features = features;
  ga.grouped = grouped;
  ga.C = C;
  if (feat_layout == 0) {  // (B,C,n)
    ga.feat_point_stride = 1;
    ga.feat_chan_stride = n;
    ga.feat_scene_stride = (long long)C * n;
  } else {  // (B,n,stride)
    S2C_REQUIRE(feat_stride >= C, "query_and_group: feat_stride %lld < C=%d", feat_stride, C);
    ga.feat_point_stride = feat_stride;
    ga.feat_chan_stride = 1;
    ga.feat_scene_stride = (long long)n * feat_stride;
  }
  ga.out_layout = out_layout;
  ga.normalize = normalize_xyz ? 1 : 0;
  ga.inv_radius = normalize_xyz ? (1.0f / radius) : 1.0f;
  return launch<true>(new_xyz, xyz, B, n, M, radius, nsample, idx, nullptr, ga, (cudaStream_t)stream);
}
