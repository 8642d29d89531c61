// Ball query / fused query+group with a spatial pre-filter, bit-exact with the reference's semantics.
//
// The reference kernel (lib/pointnet2/_ext_src/src/ball_query_gpu.cu:9-44) and the brute-force kernel of
// ball_query.cu test every (centre, point) pair: M*n = 82 M distance tests per scene at SA1.  The RESULT, however,
// only depends on the points inside the ball: "the nsample smallest indices k with d2(k) < r^2, in increasing order".
// So:
//   1. grid_build_kernel (one 8-CTA cluster per scene, phases separated by cluster barriers): bounding box, uniform grid with cell edge >= r*(1+1e-3) (capped at
//      65 536 cells), counting sort of the point indices by cell (order inside a cell is irrelevant).
//   2. query kernel (one warp per centre): visit the 3x3 runs of x-adjacent cells around the centre (each run is one
//      contiguous range of the sorted array), test ONLY those candidates with exactly the reference's fp32
//      expression, collecting the hit indices in a small per-warp list; the list sorted by index (each hit's rank =
//      number of smaller hits) and cut at nsample is the reference's neighbour list.  Padding / empty-ball rules as
//      in the reference.  The 9 runs are scanned as ONE flattened candidate range; more than 1024 hits (denser than
//      any indoor scan) fall back to a repeated minimum search over the candidates.
//   The gather epilogue (centre subtraction, 1/r, concat, channels-last output) is the one of ball_query.cu.
// Candidate count per centre drops from n (40 000) to ~200, so the kernel becomes a gather bound by the bytes it
// writes instead of by distance arithmetic.
//
// Exactness of the pre-filter: a hit has |dx| <= sqrt(d2) < r*(1+2^-22) per axis.  Cell coordinates are
// floor(fl(fl(x-lo)*inv)) -- monotone in x -- with at most 2^16 cells per axis range, so the fp32 rounding of the
// argument is < 1e-2 of a cell, far below the 1e-3*r/cs... margin built into the cell edge: a hit can never be more
// than one cell away from its centre's cell in any axis.
#include "s2c_common.cuh"
#include "group_epilogue.cuh"

namespace s2c {
namespace {

constexpr int kMaxCells = 65536;
constexpr int kBuildThreads = 1024;

struct GridParams {   // per scene, written by the build kernel
  float lo[3];
  float inv;          // 1 / cell edge
  int n[3];           // cells per axis
  int ncell;
};

__device__ __forceinline__ int cell_coord(float x, float lo, float inv, int ncells) {
  float f = __fmul_rn(__fsub_rn(x, lo), inv);
  f = fminf(fmaxf(f, -2.0f), (float)ncells + 1.0f);  // keeps far-away centres representable; inside points are unaffected
  return (int)floorf(f);
}

// ---- 1. build: bbox -> grid params -> counts -> exclusive scan -> scatter ---------------------------------------
// One thread-block CLUSTER of kBuildCluster CTAs per scene; the phases are separated by cluster barriers
// (release/acquire at cluster scope orders the global-memory traffic between the CTAs of the scene).
constexpr int kBuildCluster = 8;

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kBuildThreads)
grid_build_kernel(const float *__restrict__ xyz, int n, float radius, GridParams *__restrict__ params,
                  int *__restrict__ cell_start /* (B, kMaxCells+1) */, float4 *__restrict__ sorted /* (B, n): x,y,z,index */,
                  int *__restrict__ cursor /* (B, kMaxCells) scratch */, float *__restrict__ bbox /* (B, CL, 6) scratch */) {
  __shared__ float s_red[6][32];
  __shared__ GridParams gp;
  __shared__ int s_scan[kBuildThreads / 32];
  const int b = blockIdx.x / kBuildCluster, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster_rank();
  const int gtid = rank * kBuildThreads + tid, gthreads = kBuildCluster * kBuildThreads;
  xyz += (size_t)b * n * 3;
  cell_start += (size_t)b * (kMaxCells + 1);
  sorted += (size_t)b * n;
  cursor += (size_t)b * kMaxCells;
  bbox += (size_t)b * kBuildCluster * 6;
  // ---- phase 1: bounding box (per-CTA partials through global scratch)
  float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (int k = gtid; k < n; k += gthreads) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = xyz[(size_t)k * 3 + a];
      mn[a] = fminf(mn[a], v); mx[a] = fmaxf(mx[a], v);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    for (int o = 16; o; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
    if (lane == 0) { s_red[a][warp] = mn[a]; s_red[3 + a][warp] = mx[a]; }
  }
  __syncthreads();
  if (tid < 6) {
    float v = s_red[tid][0];
    for (int w = 1; w < kBuildThreads / 32; ++w) v = tid < 3 ? fminf(v, s_red[tid][w]) : fmaxf(v, s_red[tid][w]);
    bbox[rank * 6 + tid] = v;
  }
  cluster_sync_all();
  if (tid == 0) {  // every CTA derives the same grid parameters
    float lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
      lo[a] = bbox[a]; hi[a] = bbox[3 + a];
      for (int r = 1; r < kBuildCluster; ++r) { lo[a] = fminf(lo[a], bbox[r * 6 + a]); hi[a] = fmaxf(hi[a], bbox[r * 6 + 3 + a]); }
    }
    float cs = radius * 1.001f;  // cell edge: strictly larger than any per-axis offset of a hit
    if (!(cs > 0.f)) cs = 1.f;
    for (;;) {  // coarsen until the grid fits (cells larger than needed are still correct)
      long long tot = 1;
      int dims[3];
      for (int a = 0; a < 3; ++a) {
        const float ext = fmaxf(hi[a] - lo[a], 0.f);
        const float q = ext / cs;
        dims[a] = q < 65000.f ? (int)q + 1 : 65001;
        tot *= dims[a];
      }
      if (tot <= kMaxCells) {
        gp.n[0] = dims[0]; gp.n[1] = dims[1]; gp.n[2] = dims[2]; gp.ncell = (int)tot;
        break;
      }
      cs *= 1.26f;
    }
    gp.lo[0] = lo[0]; gp.lo[1] = lo[1]; gp.lo[2] = lo[2];
    gp.inv = 1.0f / cs;
    if (rank == 0) params[b] = gp;
  }
  __syncthreads();
  const int ncell = gp.ncell;
  // ---- phase 2: clear the histogram
  for (int c = gtid; c < ncell; c += gthreads) cursor[c] = 0;
  cluster_sync_all();
  // ---- phase 3: counts (cursor doubles as the histogram)
  for (int k = gtid; k < n; k += gthreads) {
    const int ix = min(max(cell_coord(xyz[(size_t)k * 3 + 0], gp.lo[0], gp.inv, gp.n[0]), 0), gp.n[0] - 1);
    const int iy = min(max(cell_coord(xyz[(size_t)k * 3 + 1], gp.lo[1], gp.inv, gp.n[1]), 0), gp.n[1] - 1);
    const int iz = min(max(cell_coord(xyz[(size_t)k * 3 + 2], gp.lo[2], gp.inv, gp.n[2]), 0), gp.n[2] - 1);
    atomicAdd(&cursor[(iz * gp.n[1] + iy) * gp.n[0] + ix], 1);
  }
  cluster_sync_all();
  // ---- phase 4: exclusive scan of the counts (CTA 0).  Each warp owns one contiguous chunk of cells and walks it
  //      32 cells at a time (coalesced) with a running carry; the 32 chunk totals are scanned by warp 0 and the chunk
  //      base is added in a second, cache-hot pass that also initialises the scatter cursors.
  if (rank == 0) {
    constexpr int kW = kBuildThreads / 32;
    const int chunk = ((ncell + kW - 1) / kW + 31) & ~31;
    const int c0 = warp * chunk, c1 = min(c0 + chunk, ncell);
    int carry = 0;
    for (int cb = c0; cb < c1; cb += 32) {
      const int c = cb + lane;
      const int cnt = c < c1 ? cursor[c] : 0;
      int inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      if (c < c1) cell_start[c] = carry + inc - cnt;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) s_scan[warp] = carry;
    __syncthreads();
    if (warp == 0) {
      const int tot = s_scan[lane];
      int inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      s_scan[lane] = inc - tot;
    }
    __syncthreads();
    const int base = s_scan[warp];
    for (int c = c0 + lane; c < c1; c += 32) {
      const int v = cell_start[c] + base;
      cell_start[c] = v;
      cursor[c] = v;  // becomes the write cursor of the scatter
    }
    if (tid == 0) cell_start[ncell] = n;
  }
  cluster_sync_all();
  // ---- phase 5: scatter
  for (int k = gtid; k < n; k += gthreads) {
    const float x = xyz[(size_t)k * 3 + 0], y = xyz[(size_t)k * 3 + 1], z = xyz[(size_t)k * 3 + 2];
    const int ix = min(max(cell_coord(x, gp.lo[0], gp.inv, gp.n[0]), 0), gp.n[0] - 1);
    const int iy = min(max(cell_coord(y, gp.lo[1], gp.inv, gp.n[1]), 0), gp.n[1] - 1);
    const int iz = min(max(cell_coord(z, gp.lo[2], gp.inv, gp.n[2]), 0), gp.n[2] - 1);
    const int pos = atomicAdd(&cursor[(iz * gp.n[1] + iy) * gp.n[0] + ix], 1);
    // the candidate scan reads coordinates and index with ONE coalesced 16-byte load per point
    sorted[pos] = make_float4(x, y, z, __int_as_float(k));
  }
}

// ---- 2. query (+ gather) ---------------------------------------------------------------------------------
// Per-warp shared memory: hk[kHitCap] the hit indices of the current centre (unordered) | li[nsample] its neighbour
// list.  The footprint does not depend on n, so 48 warps per SM stay resident for the gather at any scene size.
constexpr int kHitCap = 1024;

template <bool GROUP>
__global__ void __launch_bounds__(256, 5)
grid_query_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz, int n, int M, float radius, int nsample,
                  const GridParams *__restrict__ params, const int *__restrict__ cell_start, const float4 *__restrict__ sorted,
                  int *__restrict__ idx, GroupArgs ga) {
  extern __shared__ __align__(16) int smem_i[];
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int *hk = smem_i + (size_t)warp * (kHitCap + nsample);
  int *li = hk + kHitCap;
  const int b = blockIdx.y;
  xyz += (size_t)b * n * 3;
  new_xyz += (size_t)b * M * 3;
  cell_start += (size_t)b * (kMaxCells + 1);
  sorted += (size_t)b * n;
  const GridParams gp = params[b];
  const float radius2 = __fmul_rn(radius, radius);
  const float *f = (GROUP && ga.features) ? ga.features + (size_t)b * ga.feat_scene_stride : nullptr;

  for (int j = blockIdx.x * warps + warp; j < M; j += gridDim.x * warps) {
    const float cx = new_xyz[j * 3 + 0], cy = new_xyz[j * 3 + 1], cz = new_xyz[j * 3 + 2];
    const int ix = cell_coord(cx, gp.lo[0], gp.inv, gp.n[0]);
    const int iy = cell_coord(cy, gp.lo[1], gp.inv, gp.n[1]);
    const int iz = cell_coord(cz, gp.lo[2], gp.inv, gp.n[2]);
    const int x0 = max(ix - 1, 0), x1 = min(ix + 1, gp.n[0] - 1);
    // the 3x3 (z,y) neighbourhood = 9 runs of x-adjacent cells, each one contiguous range of `sorted`:
    // lane r < 9 fetches run r's bounds, then the 9 ranges are scanned as one flattened candidate range
    int rs = 0, rl = 0;
    if (lane < 9 && x0 <= x1) {
      const int z = iz + lane / 3 - 1, y = iy + lane % 3 - 1;
      if (z >= 0 && z < gp.n[2] && y >= 0 && y < gp.n[1]) {
        const int row = (z * gp.n[1] + y) * gp.n[0];
        rs = __ldg(cell_start + row + x0);
        rl = __ldg(cell_start + row + x1 + 1) - rs;
      }
    }
    int inc = rl;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    const int total = __shfl_sync(0xffffffffu, inc, 8);
    const int my_excl = inc - rl, my_adj = rs - my_excl;  // candidate t of run r lives at sorted[t + adj_r]
    int excl[9], adj[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      excl[r] = __shfl_sync(0xffffffffu, my_excl, r);
      adj[r] = __shfl_sync(0xffffffffu, my_adj, r);
    }
    // ---- collect the hits (any order); H counts all of them, only the first kHitCap are stored
    int H = 0;
    for (int t0 = 0; t0 < total; t0 += 32) {
      const int t = t0 + lane;
      bool hit = false;
      int k = 0;
      if (t < total) {
        int a = adj[0];
#pragma unroll
        for (int r = 1; r < 9; ++r) a = t >= excl[r] ? adj[r] : a;  // runs are in increasing t order: the last match wins
        const float4 p = __ldg(sorted + t + a);
        k = __float_as_int(p.w);
        hit = sqdist3(cx, cy, cz, p.x, p.y, p.z) < radius2;
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const int pos = H + __popc(m & ((1u << lane) - 1u));
        if (pos < kHitCap) hk[pos] = k;
      }
      H += __popc(m);
    }
    __syncwarp();
    int cnt;
    if (H <= kHitCap) {
      // ---- ordered selection by rank: the neighbour list is the hits sorted by index, cut at nsample
      for (int i = lane; i < H; i += 32) {
        const int k = hk[i];
        int rank = 0;
        for (int q = 0; q < H; ++q) rank += hk[q] < k ? 1 : 0;  // broadcast reads
        if (rank < nsample) li[rank] = k;
      }
      cnt = min(H, nsample);
    } else {
      // ---- more hits than the list holds (very dense neighbourhoods): repeated minimum search over the candidates
      cnt = 0;
      int prev = -1;
      while (cnt < nsample) {
        int best = 0x7fffffff;
        for (int t = lane; t < total; t += 32) {
          int a = adj[0];
#pragma unroll
          for (int r = 1; r < 9; ++r) a = t >= excl[r] ? adj[r] : a;
          const float4 p = __ldg(sorted + t + a);
          const int k = __float_as_int(p.w);
          if (k > prev && k < best && sqdist3(cx, cy, cz, p.x, p.y, p.z) < radius2) best = k;
        }
        best = __reduce_min_sync(0xffffffffu, best);
        if (best == 0x7fffffff) break;
        if (lane == 0) li[cnt] = best;
        prev = best;
        ++cnt;
      }
    }
    __syncwarp();
    const int first = cnt > 0 ? li[0] : 0;
    __syncwarp();
    for (int s = cnt + lane; s < nsample; s += 32) li[s] = first;
    __syncwarp();
    if (idx) {
      int *o = idx + ((size_t)b * M + j) * nsample;
      for (int s = lane; s < nsample; s += 32) o[s] = li[s];
    }
    if (GROUP) group_epilogue(ga, xyz, f, li, nsample, lane, cx, cy, cz, b, M, j);
    __syncwarp();
  }
}

}  // namespace
}  // namespace s2c

extern "C" long long s2c_ball_query_grid_workspace_bytes(int B, int n) {
  using namespace s2c;
  // params (64 B per scene) | sorted (B, n) float4 | cell_start (B, kMaxCells+1) | cursor (B, kMaxCells)
  return (long long)B * 64 + (long long)B * n * 16 + (long long)B * (kMaxCells + 1) * 4 + (long long)B * kMaxCells * 4 +
         (long long)B * kBuildCluster * 6 * 4 + 512;
}

extern "C" int s2c_query_and_group_grid(const float *xyz, const float *new_xyz, const float *features, int B, int n, int M,
                                        int C, int feat_layout, long long feat_stride, float radius, int nsample,
                                        int normalize_xyz, int out_layout, int *idx, float *grouped, void *workspace,
                                        long long workspace_bytes, void *stream) {
  using namespace s2c;
  S2C_REQUIRE(B >= 0 && n >= 1 && M >= 0 && C >= 0, "query_and_group_grid: bad sizes");
  S2C_REQUIRE(nsample >= 1 && nsample <= 1024, "query_and_group_grid: nsample=%d outside [1,1024]", nsample);
  S2C_REQUIRE(radius > 0.f, "query_and_group_grid: radius must be positive");
  S2C_REQUIRE(feat_layout == 0 || feat_layout == 1, "query_and_group_grid: feat_layout must be 0 or 1");
  S2C_REQUIRE(out_layout >= 0 && out_layout <= 2, "query_and_group_grid: out_layout must be 0, 1 or 2");
  if (B == 0 || M == 0) return S2C_OK;
  S2C_REQUIRE(xyz && new_xyz && (idx || grouped) && workspace, "query_and_group_grid: null pointer");
  S2C_REQUIRE(C == 0 || features || !grouped, "query_and_group_grid: features is null but C=%d", C);
  S2C_REQUIRE(workspace_bytes >= s2c_ball_query_grid_workspace_bytes(B, n), "query_and_group_grid: workspace too small");
  S2C_REQUIRE(B <= 65535, "query_and_group_grid: B too large");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char *ws = (unsigned char *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  GridParams *params = (GridParams *)ws;
  float4 *sorted = (float4 *)(ws + (size_t)B * 64);
  int *cell_start = (int *)(sorted + (size_t)B * n);
  int *cursor = cell_start + (size_t)B * (kMaxCells + 1);
  float *bbox = (float *)(cursor + (size_t)B * kMaxCells);
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * kBuildCluster));
    cfg.blockDim = dim3(kBuildThreads);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kBuildCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    S2C_CUDA(cudaLaunchKernelEx(&cfg, grid_build_kernel, xyz, n, radius, params, cell_start, sorted, cursor, bbox), "grid_build launch");
  }
  GroupArgs ga = {};
  ga.features = features; ga.grouped = grouped; ga.C = C;
  if (feat_layout == 0) { ga.feat_point_stride = 1; ga.feat_chan_stride = n; ga.feat_scene_stride = (long long)C * n; }
  else {
    S2C_REQUIRE(feat_stride >= C, "query_and_group_grid: feat_stride %lld < C=%d", feat_stride, C);
    ga.feat_point_stride = feat_stride; ga.feat_chan_stride = 1; ga.feat_scene_stride = (long long)n * feat_stride;
  }
  ga.out_layout = out_layout; ga.normalize = normalize_xyz ? 1 : 0; ga.inv_radius = normalize_xyz ? (1.0f / radius) : 1.0f;
  const int warps = 8;
  const size_t smem = (size_t)warps * (kHitCap + nsample) * 4;
  S2C_REQUIRE(smem <= 227 * 1024, "query_and_group_grid: nsample=%d too large", nsample);
  const int ctas_x = min(ceil_div(M, warps), max(1, 16 * kNumSMs / max(B, 1)));
  dim3 grid((unsigned)ctas_x, (unsigned)B);
  if (grouped) {
    S2C_CUDA(cudaFuncSetAttribute(grid_query_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "grid_query smem");
    grid_query_kernel<true><<<grid, warps * 32, smem, st>>>(new_xyz, xyz, n, M, radius, nsample, params, cell_start, sorted, idx, ga);
  } else {
    S2C_CUDA(cudaFuncSetAttribute(grid_query_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "grid_query smem");
    grid_query_kernel<false><<<grid, warps * 32, smem, st>>>(new_xyz, xyz, n, M, radius, nsample, params, cell_start, sorted, idx, ga);
  }
  S2C_CHECK_LAUNCH("grid_query");
  return S2C_OK;
}
