// Gather epilogue shared by the two fused query+group kernels (ball_query.cu: brute-force scan for small point sets,
// ball_query_grid.cu: uniform-grid pre-filter).  Given one centre's neighbour list (in shared memory) a warp writes
// that centre's block of the grouped tensor:
//     cat([ (xyz[idx] - centre) * (1/r if normalize), features[idx] ])          (QueryAndGroup.forward,
//                                                         lib/pointnet2/pointnet2_utils.py:317-376)
// in one of three layouts.  Arithmetic is bit-identical to the reference's: fp32 subtract, then a multiply by the
// fp32 reciprocal of the radius (torch's tensor / python-scalar on CUDA multiplies by the reciprocal).
#pragma once
#include "s2c_common.cuh"

namespace s2c {

struct GroupArgs {
  const float *features;  // may be null (C == 0)
  float *grouped;         // may be null (query only)
  int C;
  long long feat_point_stride, feat_chan_stride, feat_scene_stride;
  int out_layout;  // 0: (B,3+C,M,ns)  1: (B,M,ns,3+C)  2: (B,M,ns,Cp) rows [x,y,z,0 | C features | zero pad], Cp = 4 + ceil4(C)
  float inv_radius;  // 1 if !normalize_xyz
  int normalize;
};

// li: this centre's nsample neighbour indices (shared memory, already padded); (cx,cy,cz): the centre;
// xyz / f: this scene's coordinates / features; bMj = b*M + j.
__device__ __forceinline__ void group_epilogue(const GroupArgs &ga, const float *__restrict__ xyz,
                                               const float *__restrict__ f, const int *li, int nsample, int lane,
                                               float cx, float cy, float cz, int b, int M, int j) {
  const int CC = 3 + ga.C;
  if (ga.out_layout == 0) {
    // (B,3+C,M,ns): lanes run over s, one 4*ns-byte contiguous run per channel
    float *o = ga.grouped + (((size_t)b * CC) * M + j) * nsample;
    const size_t cstride = (size_t)M * nsample;
    for (int s = lane; s < nsample; s += 32) {
      const int k = li[s];
      float rx = __fsub_rn(xyz[(size_t)k * 3 + 0], cx);
      float ry = __fsub_rn(xyz[(size_t)k * 3 + 1], cy);
      float rz = __fsub_rn(xyz[(size_t)k * 3 + 2], cz);
      if (ga.normalize) {
        rx = __fmul_rn(rx, ga.inv_radius); ry = __fmul_rn(ry, ga.inv_radius); rz = __fmul_rn(rz, ga.inv_radius);
      }
      st_stream(o + s, rx);
      st_stream(o + cstride + s, ry);
      st_stream(o + 2 * cstride + s, rz);
      const float *fk = f + (size_t)k * ga.feat_point_stride;
      for (int ch = 0; ch < ga.C; ++ch)
        st_stream(o + (size_t)(3 + ch) * cstride + s, __ldg(fk + (size_t)ch * ga.feat_chan_stride));
    }
    return;
  }
  // (B,M,ns,Cp): the centre's whole block is one contiguous run of ns*Cp floats
  if (ga.out_layout == 2) {
    // rows [x, y, z, 0 | features | zero pad]: the feature block starts 16-byte aligned, so it can be written (and,
    // when the source rows are aligned too, read) 4 channels at a time, and its gradient comes back as an aligned
    // block the tensor-core dgrad / the vectorised scatter-add can address directly.
    const int CP = 4 + ((ga.C + 3) & ~3);
    float *o = ga.grouped + ((size_t)b * M + j) * (size_t)nsample * CP;
    const bool unit = ga.feat_chan_stride == 1;
    const bool vec = unit && (ga.feat_point_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(f) & 15) == 0 && (ga.C & 3) == 0;
    // slot 0 of every sample: the three offsets and the zero that aligns the feature block
    for (int s = lane; s < nsample; s += 32) {
      const int k = li[s];
      float4 v;
      v.x = __fsub_rn(xyz[(size_t)k * 3 + 0], cx);
      v.y = __fsub_rn(xyz[(size_t)k * 3 + 1], cy);
      v.z = __fsub_rn(xyz[(size_t)k * 3 + 2], cz);
      if (ga.normalize) {
        v.x = __fmul_rn(v.x, ga.inv_radius); v.y = __fmul_rn(v.y, ga.inv_radius); v.z = __fmul_rn(v.z, ga.inv_radius);
      }
      v.w = 0.f;
      st_stream4(o + (size_t)s * CP, v);
    }
    if (ga.C == 0) return;
    if (vec) {
      // aligned source rows (features of a previous set-abstraction level): 16-byte loads and stores
      const int W = ga.C >> 2;
      const int total = nsample * W;
      const float invW = 1.0f / (float)W;
#pragma unroll 4
      for (int q = lane; q < total; q += 32) {
        const int s = __float2int_rz(((float)q + 0.5f) * invW);  // q / W (exact: |(q+.5)/W - integer| >= .5/W)
        const int w = q - s * W;
        const float4 v = __ldg(reinterpret_cast<const float4 *>(f + (size_t)li[s] * ga.feat_point_stride) + w);
        st_stream4(o + (size_t)s * CP + 4 + 4 * w, v);
      }
    } else {
      // unaligned / strided source rows (the feature columns of point_clouds, row stride 3+C floats): lanes run over
      // the channels of one neighbour, so both the loads and the stores of a warp instruction cover one contiguous
      // 128-byte run (4-byte accesses, fully coalesced) instead of 32 partly used sectors
      const int Cw = CP - 4;  // feature columns incl. the zero pad
      const int total = nsample * Cw;
      const float invC = 1.0f / (float)Cw;
      const size_t cs = (size_t)ga.feat_chan_stride;
#pragma unroll 4
      for (int t = lane; t < total; t += 32) {
        const int s = __float2int_rz(((float)t + 0.5f) * invC);
        const int w = t - s * Cw;
        const float v = w < ga.C ? __ldg(f + (size_t)li[s] * ga.feat_point_stride + (size_t)w * cs) : 0.f;
        st_stream(o + (size_t)s * CP + 4 + w, v);
      }
    }
    return;
  }
  const int CP = CC;
  float *o = ga.grouped + ((size_t)b * M + j) * (size_t)nsample * CP;
  const int total = nsample * CP;
  for (int t = lane; t < total; t += 32) {
    const int s = t / CP, ch = t - s * CP;
    const int k = li[s];
    float v;
    if (ch >= CC) {
      v = 0.f;
    } else if (ch < 3) {
      v = __fsub_rn(xyz[(size_t)k * 3 + ch], ch == 0 ? cx : (ch == 1 ? cy : cz));
      if (ga.normalize) v = __fmul_rn(v, ga.inv_radius);
    } else {
      v = __ldg(f + (size_t)k * ga.feat_point_stride + (size_t)(ch - 3) * ga.feat_chan_stride);
    }
    st_stream(o + t, v);
  }
}

}  // namespace s2c
