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
// floor(f(x)), f(x) = fl(fl(x-lo)*inv), monotone in x.  With at most kMaxCellsPerAxis = 1024 cells per axis, f <= 1024
// and its two roundings are off by at most 2 * 1024 * 2^-24 = 1.2e-4 of a cell each, so for a hit
//   |f(x1) - f(x2)| <= (r / cs) * (1 + 2^-22) + 2.4e-4 <= 1 / 1.001 + 2.5e-4 < 1
// (cs >= 1.001 r; coarsening only enlarges cs): a hit can never be more than one cell away from its centre's cell in
// any axis.  (Round 1 allowed 65 001 cells on one axis, where the rounding, 8e-3 of a cell, exceeded the 1e-3 margin:
// elongated, line-like clouds could lose a hit -- tests/test_native_ops_gpu.py::test_ball_query_line_cloud.)
#include <cuda.h>

#include "s2c_common.cuh"
#include "group_epilogue.cuh"

namespace s2c {
namespace {

constexpr int kMaxCells = 65536;
constexpr int kMaxCellsPerAxis = 1024;
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
                  int *__restrict__ cursor /* (B, kMaxCells) scratch */, float *__restrict__ bbox /* (B, CL, 6) scratch */,
                  unsigned int *__restrict__ next_centre /* work counter of the gather kernel */) {
  __shared__ float s_red[6][32];
  __shared__ GridParams gp;
  __shared__ int s_scan[kBuildThreads / 32];
  const int b = blockIdx.x / kBuildCluster, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster_rank();
  const int gtid = rank * kBuildThreads + tid, gthreads = kBuildCluster * kBuildThreads;
  if (blockIdx.x == 0 && tid == 0) *next_centre = 0u;
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
      if (tot <= kMaxCells && max(dims[0], max(dims[1], dims[2])) <= kMaxCellsPerAxis) {
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
#pragma unroll 4
  for (int k = gtid; k < n; k += gthreads) {
    const int ix = min(max(cell_coord(xyz[(size_t)k * 3 + 0], gp.lo[0], gp.inv, gp.n[0]), 0), gp.n[0] - 1);
    const int iy = min(max(cell_coord(xyz[(size_t)k * 3 + 1], gp.lo[1], gp.inv, gp.n[1]), 0), gp.n[1] - 1);
    const int iz = min(max(cell_coord(xyz[(size_t)k * 3 + 2], gp.lo[2], gp.inv, gp.n[2]), 0), gp.n[2] - 1);
    atomicAdd(&cursor[(iz * gp.n[1] + iy) * gp.n[0] + ix], 1);
  }
  cluster_sync_all();
  // ---- phase 4: exclusive scan of the counts (CTA 0).  Each warp owns one contiguous chunk of cells and reads it 32
  //      cells at a time (coalesced).  Pass 1 sums the chunk with independent loads; the 32 chunk totals are scanned by
  //      warp 0; pass 2 walks the chunk again (cache-hot) FOUR 32-cell groups at a time -- their loads are issued
  //      together, only the shuffle scans are serial -- writes the prefix of every cell and initialises the scatter
  //      cursors.  (The first version carried the prefix through one load -> scan -> store round trip per 32 cells:
  //      18 serial L2 latencies per warp, 35 % of the kernel's stall samples sat on the cluster barrier behind it.)
  if (rank == 0) {
    constexpr int kW = kBuildThreads / 32;
    const int chunk = ((ncell + kW - 1) / kW + 31) & ~31;
    const int c0 = warp * chunk, c1 = min(c0 + chunk, ncell);
    int sum = 0;
#pragma unroll 8
    for (int c = c0 + lane; c < c1; c += 32) sum += cursor[c];
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_scan[warp] = sum;
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
    int carry = s_scan[warp];
    for (int cb = c0; cb < c1; cb += 128) {
      int cnt[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = cb + u * 32 + lane;
        cnt[u] = c < c1 ? cursor[c] : 0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = cb + u * 32 + lane;
        int inc = cnt[u];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += v;
        }
        if (c < c1) {
          const int v = carry + inc - cnt[u];
          cell_start[c] = v;
          cursor[c] = v;  // becomes the write cursor of the scatter
        }
        carry += __shfl_sync(0xffffffffu, inc, 31);
      }
    }
    if (tid == 0) cell_start[ncell] = n;
  }
  cluster_sync_all();
  // ---- phase 5: scatter (unrolled: the returning atomics of several points are in flight together)
#pragma unroll 4
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

// One centre: fills li[0..nsample) (shared memory of the calling warp) with the reference's neighbour list.
// kCap: capacity of the calling warp's hit list hk (more hits than that take the repeated-minimum fallback).
constexpr int kHitCapQ = 512;   // query warps inside the TMA gather kernel (shared memory is the ring's there)
template <int kCap = 1024>
__device__ __forceinline__ void grid_query_centre(const GridParams &gp, const int *__restrict__ cell_start,
                                                  const float4 *__restrict__ sorted, float cx, float cy, float cz,
                                                  float radius2, int nsample, int lane, int *hk, int *li) {
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
      if (pos < kCap) hk[pos] = k;
    }
    H += __popc(m);
  }
  __syncwarp();
  int cnt;
  if (H <= kCap) {
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
}

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
    grid_query_centre(gp, cell_start, sorted, cx, cy, cz, radius2, nsample, lane, hk, li);
    if (idx) {
      int *o = idx + ((size_t)b * M + j) * nsample;
      for (int s = lane; s < nsample; s += 32) o[s] = li[s];
    }
    if (GROUP) group_epilogue(ga, xyz, f, li, nsample, lane, cx, cy, cz, b, M, j);
    __syncwarp();
  }
}

// ---- 2b. TMA gather of the grouped rows -----------------------------------------------------------------------
// Padded channels-last output (out_layout 2) from 16-byte aligned point-major feature rows: one centre's block of the
// grouped tensor is ONE contiguous run of nsample rows [x, y, z, 0 | C features] (Cp = C + 4 floats).  The query
// kernel above (40 resident warps per SM hide its latency-bound candidate scans) leaves the neighbour lists in `idx`;
// this kernel moves the bytes, and the indexed feature rows never go through registers.  Each warp owns a ring of
// kStages shared-memory tiles of kTileRows rows and keeps it full across centre boundaries:
//   cp.async.bulk.tensor.2d ... tile::gather4   4 indexed rows of the (B*n, C) feature matrix per instruction, landing
//                                               at row pitch Cp*4 bytes: the box starts at column -4, whose 16 bytes
//                                               are outside the tensor and therefore zero-filled by the TMA unit,
//   (the warp overwrites those 16 bytes of every row with the centred, scaled coordinates)
//   cp.async.bulk.global.shared::cta            one contiguous store of the finished tile.
// A warp has kStages-2 tile loads and up to two tile stores in flight while it is not even scheduled, instead of one
// 16-byte load per lane whose latency the warp has to sit out (51 % of the stall samples of the LDG/STG epilogue at
// C = 132 were on the store that consumes the gathered load: profiles/r01_ncu_mlp_backward_s3.txt).
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_gather4(void *dst, const CUtensorMap *tmap, int col, int r0, int r1, int r2, int r3,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::
          "r"(smem_addr(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_addr(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_store(float *dst, const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(src)), "r"(bytes)
               : "memory");
}


// The neighbour lists are computed by the second half of the CTA's warps, a few centres ahead of the movers (the
// candidate scans are latency bound and leave the memory system idle; run as a separate kernel they were a 23 us serial
// prefix).  Query warp q is paired with mover q: it takes the centres the mover takes, in order, and fills the mover's
// next free ENTRY in shared memory -- the neighbour list and the centred / scaled coordinates of its rows -- so the
// mover itself never touches global memory except through the TMA unit.  Hand-over: done[q] / taken[q] counters in
// shared memory, fence.cta on both sides.
template <int kTileRows, int kStages, int kWarps, int kEntries>
__global__ void __launch_bounds__(2 * kWarps * 32, 1)
group_rows_tma_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz, int n, int M, long long centres,
                      int nsample, int *__restrict__ idx, GroupArgs ga, const __grid_constant__ CUtensorMap tmap,
                      float radius, const GridParams *__restrict__ params, const int *__restrict__ cell_start,
                      const float4 *__restrict__ sorted, unsigned int *__restrict__ next_centre) {
  static_assert(kTileRows % 4 == 0 && kTileRows <= 32 && kStages >= 3 && kEntries >= 2,
                "tile = whole gather4 instructions, one lane per row; ring of at least 3 tiles; 2+ entries per pair");
  constexpr int kAhead = kStages - 2;  // tile loads in flight per mover
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int CP = ga.C + 4;
  const uint32_t row_bytes = (uint32_t)CP * 4u;
  const uint32_t tile_bytes = row_bytes * kTileRows;
  const int ns4 = (nsample + 3) & ~3;
  // layout: [rings: warps x stages x tile_bytes] [entries: warps x kEntries x (xs: ns4 float4 | li: ns4 int | centre id)]
  //         [slot meta: warps x stages x 16 B] [mbarriers: warps x stages] [done | taken: 2 x warps ints, 128 B]
  //         [query warps' hit lists: warps x kHitCapQ ints]
  const size_t entry_bytes = (size_t)ns4 * 20 + 16;
  unsigned char *p0 = smem_raw + (size_t)kWarps * kStages * tile_bytes;
  unsigned char *p1 = p0 + (size_t)kWarps * kEntries * entry_bytes;
  unsigned char *p2 = p1 + (size_t)kWarps * kStages * 24;
  volatile int *done = reinterpret_cast<volatile int *>(p2);
  volatile int *taken = done + kWarps;
  if (threadIdx.x < 2 * kWarps) done[threadIdx.x] = 0;
  __syncthreads();
  const int pair = warp < kWarps ? warp : warp - kWarps;
  unsigned char *entries = p0 + (size_t)pair * kEntries * entry_bytes;
  // Centres are handed out dynamically (one atomic per centre on a counter the build kernel zeroed): a CTA that
  // starts late -- the sampling chain of the deeper levels runs on a side stream and occupies SMs -- simply takes fewer.

  if (warp >= kWarps) {
    // ---------------------------------------------------------------- query role
    int *hk = reinterpret_cast<int *>(p2 + 128) + (size_t)pair * kHitCapQ;
    const float radius2 = __fmul_rn(radius, radius);
    for (int k = 0;; ++k) {
      if (lane == 0) while (k - taken[pair] >= kEntries) __nanosleep(64);   // the mover still uses entry k % kEntries
      __syncwarp();
      __threadfence_block();
      float4 *xs = reinterpret_cast<float4 *>(entries + (size_t)(k % kEntries) * entry_bytes);
      int *li = reinterpret_cast<int *>(xs + ns4);
      long long c = 0;
      if (lane == 0) c = (long long)atomicAdd(next_centre, 1u);
      c = __shfl_sync(0xffffffffu, c, 0);
      if (c >= centres) {   // no work left: tell the mover
        if (lane == 0) {
          *reinterpret_cast<long long *>(li + ns4) = -1;
          __threadfence_block();
          done[pair] = k + 1;
        }
        return;
      }
      const int b = (int)(c / M);
      const GridParams gp = params[b];
      const float *ctr = new_xyz + (size_t)c * 3;
      const float cx = __ldg(ctr), cy = __ldg(ctr + 1), cz = __ldg(ctr + 2);
      grid_query_centre<kHitCapQ>(gp, cell_start + (size_t)b * (kMaxCells + 1), sorted + (size_t)b * n, cx, cy, cz,
                                  radius2, nsample, lane, hk, li);
      const float *px = xyz + (size_t)b * n * 3;
      int *o = idx + (size_t)c * nsample;
      for (int s = lane; s < ns4; s += 32) {
        const int kk = li[min(s, nsample - 1)];
        if (s < nsample) o[s] = kk; else li[s] = kk;
        float4 v;
        v.x = __fsub_rn(__ldg(px + (size_t)kk * 3 + 0), cx);
        v.y = __fsub_rn(__ldg(px + (size_t)kk * 3 + 1), cy);
        v.z = __fsub_rn(__ldg(px + (size_t)kk * 3 + 2), cz);
        if (ga.normalize) {
          v.x = __fmul_rn(v.x, ga.inv_radius); v.y = __fmul_rn(v.y, ga.inv_radius); v.z = __fmul_rn(v.z, ga.inv_radius);
        }
        v.w = 0.f;
        xs[s] = v;
      }
      if (lane == 0) *reinterpret_cast<long long *>(li + ns4) = c;
      __threadfence_block();
      __syncwarp();
      if (lane == 0) done[pair] = k + 1;
    }
  }

  // ------------------------------------------------------------------ mover role
  unsigned char *ring = smem_raw + (size_t)warp * kStages * tile_bytes;
  uint4 *meta = reinterpret_cast<uint4 *>(p1) + warp * kStages;   // {out lo, out hi, xs shared address, rows | last << 16}
  uint64_t *bars = reinterpret_cast<uint64_t *>(p1 + (size_t)kWarps * kStages * 16) + warp * kStages;
  if (lane == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
  }
  __syncwarp();
  const int T = (nsample + kTileRows - 1) / kTileRows;
  uint32_t issued = 0, finished = 0;   // tiles of this warp (ring slot = tile number % kStages)
  int released = 0;                    // centres whose entry has been handed back to the query warp

  auto finish = [&]() {   // wait for the oldest tile's rows, patch the coordinate slot of every row, store the tile
    const uint32_t slot = finished % kStages;
    mbar_wait(bars + slot, (finished / kStages) & 1u);
    const uint4 m = meta[slot];
    const uint32_t rows = m.w & 0xffffu;
    unsigned char *tile = ring + (size_t)slot * tile_bytes;
    if (lane < (int)rows) {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(m.z + 16u * lane));
      *reinterpret_cast<float4 *>(tile + (size_t)lane * row_bytes) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      float *out = reinterpret_cast<float *>(((uint64_t)m.y << 32) | (uint64_t)m.x);
      bulk_store(out, tile, rows * row_bytes);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (m.w >> 16) {   // last tile of its centre: the entry may be refilled
        ++released;
        taken[warp] = released;
      }
    } else if (m.w >> 16) {
      ++released;
    }
    ++finished;
  };

  for (int k = 0;; ++k) {
    if (lane == 0) while (done[warp] <= k) __nanosleep(64);   // the paired query warp has filled entry k % kEntries
    __syncwarp();
    __threadfence_block();
    const float4 *xs = reinterpret_cast<const float4 *>(entries + (size_t)(k % kEntries) * entry_bytes);
    const int *li = reinterpret_cast<const int *>(xs + ns4);
    const long long c = *reinterpret_cast<const long long *>(li + ns4);
    if (c < 0) break;
    const int b = (int)(c / M);
    const int row_base = b * n;   // row of this scene's point 0 in the (B*n, C) feature matrix
    float *out = ga.grouped + (size_t)c * (size_t)nsample * CP;
    for (int t = 0; t < T; ++t) {
      if (issued - finished == (uint32_t)kAhead) finish();
      // the store that used this slot last (tile issued - kStages) is at least two store groups old
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncwarp();
      const uint32_t slot = issued % kStages;
      const int rows = min(kTileRows, nsample - t * kTileRows);
      const int quads = (rows + 3) >> 2;
      if (lane == 0) {
        const uint64_t o = reinterpret_cast<uint64_t>(out + (size_t)t * kTileRows * CP);
        meta[slot] = make_uint4((uint32_t)o, (uint32_t)(o >> 32), smem_addr(xs + t * kTileRows),
                                (uint32_t)rows | (t == T - 1 ? 0x10000u : 0u));
        mbar_arrive_expect_tx(bars + slot, (uint32_t)quads * 4u * row_bytes);
      }
      __syncwarp();
      if (lane < quads) {
        const int4 r = *reinterpret_cast<const int4 *>(li + t * kTileRows + lane * 4);
        tma_gather4(ring + (size_t)slot * tile_bytes + (size_t)lane * 4 * row_bytes, &tmap, -4, row_base + r.x,
                    row_base + r.y, row_base + r.z, row_base + r.w, bars + slot);
      }
      ++issued;
    }
    // an entry is refilled only after its tiles are patched; with fewer tiles per centre than the look-ahead the
    // pipeline would otherwise hold more centres than there are entries
    while ((int)(k + 1 - released) >= kEntries) finish();
  }
  while (finished < issued) finish();
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// (rows, C) fp32 feature matrix with row stride `ld` floats; box = [C + 4 columns x 1 row]: tile::gather4 loads four
// such rows per instruction, and a box that starts at column -4 gets its first 16 bytes zero-filled
int make_feature_tmap(CUtensorMap *tmap, const float *base, long long rows, int C, long long ld) {
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<EncodeTiledFn>(p);
  }
  if (!enc) {
    set_error("query_and_group_grid: cuTensorMapEncodeTiled is not available from this driver");
    return S2C_ERR_UNSUPPORTED;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)(C + 4), 1};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("query_and_group_grid: cuTensorMapEncodeTiled failed (%d) for rows=%lld C=%d ld=%lld", (int)r, rows, C, ld);
    return S2C_ERR_CUDA;
  }
  return S2C_OK;
}

int g_tma_variant = 0;  // ring geometry of the TMA gather (tuning knob, s2c_query_and_group_grid_tune)

}  // namespace
}  // namespace s2c

extern "C" int s2c_query_and_group_grid_tune(int variant) {
  s2c::g_tma_variant = variant;
  return S2C_OK;
}

extern "C" long long s2c_ball_query_grid_workspace_bytes(int B, int n) {
  using namespace s2c;
  // params (64 B per scene) | sorted (B, n) float4 | cell_start (B, kMaxCells+1) | cursor (B, kMaxCells)
  return (long long)B * 64 + (long long)B * n * 16 + (long long)B * (kMaxCells + 1) * 4 + (long long)B * kMaxCells * 4 +
         (long long)B * kBuildCluster * 6 * 4 + 64 + 512;
}

namespace s2c {
namespace {
struct GridWorkspace {
  GridParams *params; float4 *sorted; int *cell_start; int *cursor; float *bbox; unsigned int *counter;
};
GridWorkspace carve_grid(void *workspace, int B, int n) {
  GridWorkspace w;
  unsigned char *ws = (unsigned char *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  w.params = (GridParams *)ws;
  w.sorted = (float4 *)(ws + (size_t)B * 64);
  w.cell_start = (int *)(w.sorted + (size_t)B * n);
  w.cursor = w.cell_start + (size_t)B * (kMaxCells + 1);
  w.bbox = (float *)(w.cursor + (size_t)B * kMaxCells);
  w.counter = (unsigned int *)(w.bbox + (size_t)B * kBuildCluster * 6);
  return w;
}
int launch_grid_build(const float *xyz, int B, int n, float radius, const GridWorkspace &w, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * kBuildCluster));
  cfg.blockDim = dim3(kBuildThreads);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kBuildCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  S2C_CUDA(cudaLaunchKernelEx(&cfg, grid_build_kernel, xyz, n, radius, w.params, w.cell_start, w.sorted, w.cursor, w.bbox, w.counter), "grid_build launch");
  return S2C_OK;
}
int query_and_group_grid_impl(const float *xyz, const float *new_xyz, const float *features, int B, int n, int M,
                              int C, int feat_layout, long long feat_stride, float radius, int nsample,
                              int normalize_xyz, int out_layout, int *idx, float *grouped, void *workspace,
                              long long workspace_bytes, void *stream, bool prebuilt) {
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
  const GridWorkspace gw = carve_grid(workspace, B, n);
  GridParams *params = gw.params;
  float4 *sorted = gw.sorted;
  int *cell_start = gw.cell_start;
  unsigned int *counter = gw.counter;
  if (prebuilt) {  // the grid of these points / this radius is already in the workspace: only re-arm the work counter
    S2C_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), st), "query_and_group_grid counter reset");
  } else {
    if (int rc = launch_grid_build(xyz, B, n, radius, gw, st)) return rc;
  }
  GroupArgs ga = {};
  ga.features = features; ga.grouped = grouped; ga.C = C;
  if (feat_layout == 0) { ga.feat_point_stride = 1; ga.feat_chan_stride = n; ga.feat_scene_stride = (long long)C * n; }
  else {
    S2C_REQUIRE(feat_stride >= C, "query_and_group_grid: feat_stride %lld < C=%d", feat_stride, C);
    ga.feat_point_stride = feat_stride; ga.feat_chan_stride = 1; ga.feat_scene_stride = (long long)n * feat_stride;
  }
  ga.out_layout = out_layout; ga.normalize = normalize_xyz ? 1 : 0; ga.inv_radius = normalize_xyz ? (1.0f / radius) : 1.0f;
  // ---- TMA gather path: padded channels-last output from 16-byte aligned point-major feature rows whose padded row
  //      (C + 4 floats) is a multiple of 32 bytes (gather4 lands 4 rows per instruction at 128-byte aligned offsets)
  const bool tma_ok = grouped && features && out_layout == 2 && feat_layout == 1 && C >= 32 && (C & 3) == 0 &&
                      ((C + 4) & 7) == 0 && C + 4 <= 256 && (feat_stride & 3) == 0 &&
                      (reinterpret_cast<uintptr_t>(features) & 15) == 0 && (long long)B * n < 2147483647LL &&
                      nsample >= 4 && idx != nullptr && g_tma_variant >= 0;
  if (tma_ok) {
    CUtensorMap tmap;
    if (int rc = make_feature_tmap(&tmap, features, (long long)B * n, C, feat_stride)) return rc;
    // 2. the bytes
    const int ns4 = (nsample + 3) & ~3;
    auto launch = [&](auto kern, int tile_rows, int stages, int warps, int nentries) -> int {
      const size_t smem = (size_t)warps * stages * tile_rows * (C + 4) * 4 + (size_t)warps * nentries * (ns4 * 20 + 16) +
                          (size_t)warps * stages * 24 + 128 + (size_t)warps * kHitCapQ * 4;
      if (smem > 227 * 1024) return -1;
      S2C_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "group_rows_tma smem");
      const long long centres = (long long)B * M;
      const int ctas = (int)min((long long)kNumSMs, ceil_div_ll(centres, warps));
      kern<<<ctas, 2 * warps * 32, smem, st>>>(new_xyz, xyz, n, M, centres, nsample, idx, ga, tmap, radius, params,
                                               cell_start, sorted, counter);
      S2C_CHECK_LAUNCH("group_rows_tma");
      return S2C_OK;
    };
    int rc = -1;
    switch (g_tma_variant) {
      case 1: rc = launch(group_rows_tma_kernel<8, 4, 8, 3>, 8, 4, 8, 3); break;
      case 2: rc = launch(group_rows_tma_kernel<8, 4, 9, 3>, 8, 4, 9, 3); break;
      case 3: rc = launch(group_rows_tma_kernel<8, 5, 7, 3>, 8, 5, 7, 3); break;
      case 4: rc = launch(group_rows_tma_kernel<8, 3, 12, 2>, 8, 3, 12, 2); break;
      case 5: rc = launch(group_rows_tma_kernel<16, 3, 6, 3>, 16, 3, 6, 3); break;
      default: rc = launch(group_rows_tma_kernel<8, 3, 10, 3>, 8, 3, 10, 3); break;   // 65.5 % of HBM at C=132, ns=64
    }
    if (rc == -1) rc = launch(group_rows_tma_kernel<8, 3, 4, 2>, 8, 3, 4, 2);   // smaller footprint for large nsample
    if (rc != -1) return rc;   // (-1: no ring fits for this nsample -> the LDG/STG epilogue below)
  }
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
}  // namespace
}  // namespace s2c

extern "C" int s2c_query_and_group_grid(const float *xyz, const float *new_xyz, const float *features, int B, int n, int M,
                                        int C, int feat_layout, long long feat_stride, float radius, int nsample,
                                        int normalize_xyz, int out_layout, int *idx, float *grouped, void *workspace,
                                        long long workspace_bytes, void *stream) {
  return s2c::query_and_group_grid_impl(xyz, new_xyz, features, B, n, M, C, feat_layout, feat_stride, radius, nsample,
                                        normalize_xyz, out_layout, idx, grouped, workspace, workspace_bytes, stream, false);
}

// The uniform grid depends only on (xyz, radius): a caller that knows the next batch's coordinates can build it ahead of
// time (s2c_ball_query_grid_build, e.g. on a copy stream during the previous training step) and run only the
// query / gather kernel on the critical path (s2c_query_and_group_grid_prebuilt, same arguments and results as
// s2c_query_and_group_grid; `workspace` must hold the grid built from the same xyz / B / n / radius).
extern "C" int s2c_ball_query_grid_build(const float *xyz, int B, int n, float radius, void *workspace,
                                         long long workspace_bytes, void *stream) {
  using namespace s2c;
  S2C_REQUIRE(B >= 0 && n >= 1, "ball_query_grid_build: bad sizes");
  S2C_REQUIRE(radius > 0.f, "ball_query_grid_build: radius must be positive");
  if (B == 0) return S2C_OK;
  S2C_REQUIRE(xyz && workspace, "ball_query_grid_build: null pointer");
  S2C_REQUIRE(workspace_bytes >= s2c_ball_query_grid_workspace_bytes(B, n), "ball_query_grid_build: workspace too small");
  S2C_REQUIRE(B <= 65535, "ball_query_grid_build: B too large");
  return launch_grid_build(xyz, B, n, radius, carve_grid(workspace, B, n), (cudaStream_t)stream);
}

extern "C" int s2c_query_and_group_grid_prebuilt(const float *xyz, const float *new_xyz, const float *features, int B, int n,
                                                 int M, int C, int feat_layout, long long feat_stride, float radius,
                                                 int nsample, int normalize_xyz, int out_layout, int *idx, float *grouped,
                                                 void *workspace, long long workspace_bytes, void *stream) {
  return s2c::query_and_group_grid_impl(xyz, new_xyz, features, B, n, M, C, feat_layout, feat_stride, radius, nsample,
                                        normalize_xyz, out_layout, idx, grouped, workspace, workspace_bytes, stream, true);
}
