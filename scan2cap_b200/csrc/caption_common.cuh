// Helpers shared by the two implementations of the caption decoder recurrence: caption.cu (one thread-block cluster,
// weights streamed from L2) and caption_grid.cu (persistent cooperative grid, weights resident in shared memory).
#pragma once
#include "s2c_common.cuh"

// threads per CTA of the including kernel file (a compile-time constant gives tighter loops than S2C_CAP_NT)
#ifndef S2C_CAP_NT
#define S2C_CAP_NT blockDim.x
#endif

namespace s2c {
namespace {

constexpr int kRows = 8;  // batch rows per cluster

__device__ __forceinline__ uint32_t cl_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cl_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ldw4(const float *p) {  // weights: read-only for the whole kernel
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stamp(long long *ts, int c, int t, int slot) {
  if (ts != nullptr && c == 0 && threadIdx.x == 0) {
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    ts[t * 8 + slot] = (long long)v;
  }
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// out[i][r] = sum_k w_i[k] * x[r][k], i < 4, r < 8.  x: shared memory, row stride xld (multiple of 4), K % 4 == 0.
// Result: lane l returns out[l >> 3][l & 7].
template <bool SM = false>  // SM: the weight rows are in shared memory (plain 16-byte loads) instead of global memory
__device__ __forceinline__ float gemv_quad(const float *__restrict__ w0, const float *__restrict__ w1,
                                           const float *__restrict__ w2, const float *__restrict__ w3, int K,
                                           const float *x, int xld, int lane) {
  float a[32];
#pragma unroll
  for (int m = 0; m < 32; ++m) a[m] = 0.f;
#pragma unroll 4
  for (int k = lane * 4; k < K; k += 128) {
    float4 q0, q1, q2, q3;
    if (SM) {
      q0 = *reinterpret_cast<const float4 *>(w0 + k); q1 = *reinterpret_cast<const float4 *>(w1 + k);
      q2 = *reinterpret_cast<const float4 *>(w2 + k); q3 = *reinterpret_cast<const float4 *>(w3 + k);
    } else {
      q0 = ldw4(w0 + k); q1 = ldw4(w1 + k); q2 = ldw4(w2 + k); q3 = ldw4(w3 + k);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float4 xv = *reinterpret_cast<const float4 *>(x + r * xld + k);
      a[r] = fmaf(q0.w, xv.w, fmaf(q0.z, xv.z, fmaf(q0.y, xv.y, fmaf(q0.x, xv.x, a[r]))));
      a[8 + r] = fmaf(q1.w, xv.w, fmaf(q1.z, xv.z, fmaf(q1.y, xv.y, fmaf(q1.x, xv.x, a[8 + r]))));
      a[16 + r] = fmaf(q2.w, xv.w, fmaf(q2.z, xv.z, fmaf(q2.y, xv.y, fmaf(q2.x, xv.x, a[16 + r]))));
      a[24 + r] = fmaf(q3.w, xv.w, fmaf(q3.z, xv.z, fmaf(q3.y, xv.y, fmaf(q3.x, xv.x, a[24 + r]))));
    }
  }
  // halving butterfly: after the step with offset s, bit s of the lane id selects bit s of the element index
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int m = 0; m < s; ++m) {
      const float send = up ? a[m] : a[m + s];
      const float keep = up ? a[m + s] : a[m];
      a[m] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return a[0];
}

// rows [j0, j0+4) of a row-major matrix (row stride ld), clamped to jmax-1 (results of clamped rows are discarded)
#define S2C_QUAD_PTRS(W, ld, j0, jmax)                                           \
  (W) + (size_t)min((j0) + 0, (jmax)-1) * (ld), (W) + (size_t)min((j0) + 1, (jmax)-1) * (ld), \
      (W) + (size_t)min((j0) + 2, (jmax)-1) * (ld), (W) + (size_t)min((j0) + 3, (jmax)-1) * (ld)

// global (nb rows of ncols floats, row stride ldg) -> shared (8 rows, stride xld), rows >= nb zero; L2 loads (.cg):
// the data was written by other CTAs of the cluster earlier in this kernel
// BATCH = 4 (persistent-grid kernels, 256 threads): four independent loads per thread before the first shared-memory
// store -- an in-order warp otherwise pays one L2 round trip per element it moves.  BATCH = 1 (cluster kernels, 512
// threads at the 128-register limit): the plain loop; the batched form measured 10 % slower there (spills).
template <int BATCH = 1>
__device__ __forceinline__ void load_rows(float *xs, int xld, const float *g, size_t ldg, int ncols, int nb) {
  const int c4 = ncols >> 2, total = kRows * c4;
  if (BATCH == 1) {
    for (int i = threadIdx.x; i < total; i += S2C_CAP_NT) {
      const int r = i / c4, c = (i - r * c4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g != nullptr && r < nb) v = __ldcg(reinterpret_cast<const float4 *>(g + (size_t)r * ldg + c));
      *reinterpret_cast<float4 *>(xs + r * xld + c) = v;
    }
    return;
  }
  for (int i0 = threadIdx.x; i0 < total; i0 += 4 * S2C_CAP_NT) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * S2C_CAP_NT;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < total) {
        const int r = i / c4, c = (i - r * c4) * 4;
        if (g != nullptr && r < nb) v[u] = __ldcg(reinterpret_cast<const float4 *>(g + (size_t)r * ldg + c));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * S2C_CAP_NT;
      if (i < total) {
        const int r = i / c4, c = (i - r * c4) * 4;
        *reinterpret_cast<float4 *>(xs + r * xld + c) = v[u];
      }
    }
  }
}
// n floats global (L2) -> shared, four independent loads per thread at a time
__device__ __forceinline__ void load_flat(float *dst, const float *src, int n) {
  for (int i0 = threadIdx.x; i0 < n; i0 += 4 * S2C_CAP_NT) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * S2C_CAP_NT;
      v[u] = i < n ? __ldcg(src + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * S2C_CAP_NT;
      if (i < n) dst[i] = v[u];
    }
  }
}

struct Slices {
  int hs, j0, j1;  // hidden units of this CTA
  int e0, e1;      // embedding units of this CTA
  int f0, f1;      // feature units of this CTA
};
__device__ __forceinline__ Slices make_slices(int c, int CL, int H, int E, int F) {
  Slices s;
  s.hs = H / CL;
  s.j0 = c * s.hs; s.j1 = s.j0 + s.hs;
  const int es = ((E + CL - 1) / CL + 3) & ~3;
  s.e0 = min(c * es, E); s.e1 = min(s.e0 + es, E);
  const int fs = ((F + CL - 1) / CL + 3) & ~3;
  s.f0 = min(c * fs, F); s.f1 = min(s.f0 + fs, F);
  return s;
}

}  // namespace
}  // namespace s2c

using namespace s2c;

namespace s2c {
namespace {

// Shared-memory plan (floats): XA[8][XLD] | XB[8][XLD] | G[6*hs][8] | probs[8][K] | sc[8][K] | att[8][F] | vk[8][K] (int) |
//   nv[8] | uniform[8] | pb[12] | pair_r[kMaxPairs] | pair_k[kMaxPairs] | objs | mcache | dmacc
//   forward: XLD = F+H ; backward: XLD = 3H.  The valid set of every scene is constant over the words, so (cache
//   level >= 1) the valid proposals' feature rows and (level 2) their map_feat rows -- forward: this CTA's share of the
//   (row, proposal) pairs, whole rows; backward: all pairs, this CTA's hidden units, plus the d_mapped accumulators --
//   stay in shared memory for the whole kernel.
constexpr int kMaxPairs = 96;  // (row, valid proposal) pairs the caches hold (8 scenes x (10 locals + self) = 88)
struct SmemPlan {
  float *XA, *XB, *G, *probs, *sc, *att, *objs, *mcache, *dmacc;
  int *vk, *nv, *uniform, *pb;  // pb[r]: first pair of row r; pb[8]: total pairs, or -1 when the caches are off
  int *pair_r, *pair_k;
  int mask;  // bit 0: objs cache, bit 1: mcache / dmacc
};
__host__ __device__ __forceinline__ size_t mcache_floats(bool bwd, int CL, int hs, int H) {
  return bwd ? (size_t)kMaxPairs * hs : (size_t)((kMaxPairs + CL - 1) / CL) * H;
}
__device__ __forceinline__ SmemPlan plan(float *base, int xld, int hs, int K, int F, int H, int CL, bool bwd, int mask) {
  SmemPlan p;
  p.XA = base; base += kRows * xld;
  p.XB = base; base += kRows * xld;
  p.G = base; base += 6 * hs * kRows;
  p.probs = base; base += kRows * K;
  p.sc = base; base += kRows * K;
  p.att = base; base += kRows * F;
  p.vk = reinterpret_cast<int *>(base); base += kRows * K;
  p.nv = reinterpret_cast<int *>(base); base += kRows;
  p.uniform = reinterpret_cast<int *>(base); base += kRows;
  p.pb = reinterpret_cast<int *>(base); base += 12;
  p.pair_r = reinterpret_cast<int *>(base); base += kMaxPairs;
  p.pair_k = reinterpret_cast<int *>(base); base += kMaxPairs;
  p.objs = base; base += (mask & 1) ? (size_t)kMaxPairs * F : 0;
  p.mcache = base; base += (mask & 2) ? mcache_floats(bwd, CL, hs, H) : 0;
  p.dmacc = base;
  p.mask = mask;
  return p;
}
size_t plan_bytes(int xld, int hs, int K, int F, int H, int CL, bool bwd, int mask) {
  size_t fl = (size_t)2 * kRows * xld + (size_t)6 * hs * kRows + (size_t)3 * kRows * K + (size_t)kRows * F + 2 * kRows + 12 +
              2 * kMaxPairs;
  if (mask & 1) fl += (size_t)kMaxPairs * F;
  if (mask & 2) fl += mcache_floats(bwd, CL, hs, H) * (bwd ? 2 : 1);
  return sizeof(float) * fl + 16;
}

// valid-object lists of the cluster's rows (constant over the steps).  A row without any valid object gets the
// uniform distribution over all K objects (softmax of K equal -1e30 scores), flagged in `uniform`.
__device__ __forceinline__ void build_valid_lists(const SmemPlan &sp, const float *valid, const float *obj, int rb, int nb,
                                                  int K, int F) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < kRows) {
    const int r = warp;
    int cnt = 0;
    if (r < nb) {
      for (int k0 = 0; k0 < K; k0 += 32) {
        const int k = k0 + lane;
        const bool v = k < K && valid[(size_t)(rb + r) * K + k] != 0.f;
        const unsigned m = __ballot_sync(0xffffffffu, v);
        if (v) sp.vk[r * K + cnt + __popc(m & ((1u << lane) - 1u))] = k;
        cnt += __popc(m);
      }
      int uni = 0;
      if (cnt == 0) {
        uni = 1;
        for (int k = lane; k < K; k += 32) sp.vk[r * K + k] = k;
        cnt = K;
      }
      if (lane == 0) { sp.nv[r] = cnt; sp.uniform[r] = uni; }
    } else if (lane == 0) {
      sp.nv[r] = 0; sp.uniform[r] = 0;
    }
  }
  for (int i = threadIdx.x; i < kRows * K; i += S2C_CAP_NT) sp.probs[i] = 0.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int r = 0; r < kRows; ++r) { sp.pb[r] = tot; tot += sp.nv[r]; }
    sp.pb[kRows] = (sp.mask != 0 && tot <= kMaxPairs) ? tot : -1;
  }
  __syncthreads();
  if (sp.pb[kRows] >= 0) {  // the valid set is the same for every word: keep those proposals' features on chip
    const int f4 = F >> 2;
    for (int r = 0; r < kRows; ++r)
      for (int ii = threadIdx.x; ii < sp.nv[r]; ii += S2C_CAP_NT) {
        sp.pair_r[sp.pb[r] + ii] = r;
        sp.pair_k[sp.pb[r] + ii] = sp.vk[r * K + ii];
      }
    for (int r = 0; r < ((sp.mask & 1) ? nb : 0); ++r) {
      const int n = sp.nv[r];
      for (int i = threadIdx.x; i < n * f4; i += S2C_CAP_NT) {
        const int ii = i / f4, c = (i - ii * f4) * 4;
        const int k = sp.vk[r * K + ii];
        *reinterpret_cast<float4 *>(sp.objs + (size_t)(sp.pb[r] + ii) * F + c) =
            __ldg(reinterpret_cast<const float4 *>(obj + ((size_t)(rb + r) * K + k) * F + c));
      }
    }
  }
  __syncthreads();
}
// feature row of the ii-th valid proposal of row r: shared-memory cache, or global memory when the cache is off
__device__ __forceinline__ const float *obj_row(const SmemPlan &sp, const float *obj, int rb, int r, int ii, int K, int F) {
  return ((sp.mask & 1) && sp.pb[kRows] >= 0) ? sp.objs + (size_t)(sp.pb[r] + ii) * F
                           : obj + ((size_t)(rb + r) * K + sp.vk[r * K + ii]) * F;
}


}  // namespace
}  // namespace s2c
