// Fused VoteNet detection loss, forward AND backward in one launch.
//
// Replaces compute_vote_loss / compute_objectness_loss / compute_box_and_sem_cls_loss and the detection part of
// get_scene_cap_loss of the reference (lib/loss_helper.py:24-187, 381-491; utils/nn_distance.py:32-59), which in
// framework form are ~200 small kernels forward and ~300 backward per training step (dense (B,K,G,3) difference
// tensors, gathers, one-hots, masked means).  The inputs are tiny -- B*1024 votes, B*256 proposals with 97 head
// outputs, B*128 ground-truth slots -- so ONE CTA walks them twice: pass 1 evaluates every term and its batch-wide
// sums, pass 2 writes d(det_loss)/d(input) for
//     det_loss = vote + 0.5*objectness + box + 0.1*sem_cls ,
//     box      = center + 0.1*heading_cls + heading_reg + 0.1*size_cls + size_reg            (loss_helper.py:409, 472-476)
// Label decisions (arg-min over the ground-truth centres, the 0.3 / 0.6 m objectness thresholds) are evaluated with
// the reference's fp32 operation order so the integer outputs (objectness_label, object_assignment) stay bit-exact.
#include <math.h>

#include "s2c_common.cuh"

namespace s2c {
namespace {

constexpr int kThreads = 1024;
constexpr float kNear = 0.3f, kFar = 0.6f;

enum Sum { S_VOTE, S_VMASK, S_OBJ, S_OMASK, S_LABEL, S_C1, S_C2, S_BMASK, S_HCLS, S_HREG, S_SCLS, S_SREG, S_SEM,
           S_ACC, S_COUNT };

struct LossArgs {
  int B, S, N, K, G, NH, NS, NC, W;        // W = 2 + 3 + 2*NH + 4*NS + NC head outputs per proposal
  const float *vote_xyz, *seed_xyz;       // (B,S,3)
  const int *seed_inds; long long seed_ld; // (B,S) int32, row stride
  const float *vote_label;                 // (B,N,9)
  const long long *vote_label_mask;        // (B,N)
  const float *agg_xyz;                    // (B,K,3) aggregated_vote_xyz
  const float *net;                        // (B,K,W) head outputs
  const float *center;                     // (B,K,3) = agg_xyz + net[...,2:5]
  const float *center_label;               // (B,G,3)
  const long long *heading_class_label;    // (B,G)
  const float *heading_residual_label;     // (B,G)
  const long long *size_class_label;       // (B,G)
  const float *size_residual_label;        // (B,G,3)
  const long long *sem_cls_label;          // (B,G)
  const float *box_label_mask;             // (B,G)
  const float *mean_size;                  // (NS,3)
  // outputs
  float *stats;                            // [16]: see s2c.h
  long long *objectness_label;             // (B,K)
  float *objectness_mask;                  // (B,K)
  long long *object_assignment;            // (B,K)
  float *d_vote_xyz;                       // (B,S,3)
  float *d_net;                            // (B,K,W) (columns 2..4 stay zero: the centre gradient is returned in d_center)
  float *d_center;                         // (B,K,3)
  int *scratch;                            // (B*K) int: arg-min GT of every predicted centre | (B*G) int: arg-min proposal of every GT
};

__device__ __forceinline__ float sq3(float ax, float ay, float az, float bx, float by, float bz) {
  // torch: sum((a - b) ** 2, -1) -- sequential fp32 adds, no contraction
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ float huber(float e) {  // delta = 1 (utils/nn_distance.py:11-28)
  const float a = fabsf(e), q = fminf(a, 1.0f);
  return 0.5f * q * q + (a - q);
}
__device__ __forceinline__ float huber_grad(float e) { return fminf(fmaxf(e, -1.0f), 1.0f); }

// log-sum-exp of n logits at stride 1
__device__ __forceinline__ float lse(const float *x, int n) {
  float m = x[0];
  for (int i = 1; i < n; ++i) m = fmaxf(m, x[i]);
  float s = 0.f;
  for (int i = 0; i < n; ++i) s += expf(x[i] - m);
  return m + logf(s);
}

__device__ double block_sum(double v, double *red) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.0;
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  const double r = red[0];
  return r;
}

__global__ void __launch_bounds__(kThreads, 1) detection_loss_kernel(LossArgs a) {
  __shared__ double red[32];
  __shared__ double tot[S_COUNT];
  const int tid = threadIdx.x;
  const int BK = a.B * a.K, BS = a.B * a.S, BG = a.B * a.G;
  const int oh = 5, ohr = 5 + a.NH, os = 5 + 2 * a.NH, osr = os + a.NS, osem = os + 4 * a.NS;
  double acc[S_COUNT];
#pragma unroll
  for (int i = 0; i < S_COUNT; ++i) acc[i] = 0.0;

  // ---- votes (loss_helper.py:24-69): L1 distance of every seed's vote to the nearest of its 3 GT votes, masked mean
  for (int i = tid; i < BS; i += kThreads) {
    const int b = i / a.S;
    const int ind = a.seed_inds[(size_t)b * a.seed_ld + (i - b * a.S)];
    const float m = (float)a.vote_label_mask[(size_t)b * a.N + ind];
    const float *gt = a.vote_label + ((size_t)b * a.N + ind) * 9;
    const float *sx = a.seed_xyz + (size_t)i * 3, *vx = a.vote_xyz + (size_t)i * 3;
    float best = 3.4e38f;
    for (int j = 0; j < 3; ++j) {
      float d = 0.f;
      for (int c = 0; c < 3; ++c) d = __fadd_rn(d, fabsf(__fsub_rn(vx[c], __fadd_rn(gt[3 * j + c], sx[c]))));
      best = fminf(best, d);
    }
    acc[S_VOTE] += (double)(best * m);
    acc[S_VMASK] += (double)m;
  }
  // ---- proposals: objectness labels from the aggregated vote positions, then every masked box / class term
  for (int i = tid; i < BK; i += kThreads) {
    const int b = i / a.K;
    const float *p = a.agg_xyz + (size_t)i * 3;
    const float *gl = a.center_label + (size_t)b * a.G * 3;
    float d1 = 3.4e38f;
    int j1 = 0;
    for (int j = 0; j < a.G; ++j) {
      const float d = sq3(p[0], p[1], p[2], gl[3 * j], gl[3 * j + 1], gl[3 * j + 2]);
      if (d < d1) { d1 = d; j1 = j; }   // first minimum
    }
    const float e = __fsqrt_rn(__fadd_rn(d1, 1e-6f));
    const int label = e < kNear ? 1 : 0;
    const float mask = (label || e > kFar) ? 1.f : 0.f;
    a.objectness_label[i] = label;
    a.objectness_mask[i] = mask;
    a.object_assignment[i] = j1;
    const float *x = a.net + (size_t)i * a.W;
    // objectness: weighted CE, weights (0.2, 0.8)
    const float l2 = lse(x, 2);
    acc[S_OBJ] += (double)((label ? 0.8f : 0.2f) * (l2 - x[label]) * mask);
    acc[S_OMASK] += (double)mask;
    acc[S_LABEL] += (double)label;
    acc[S_ACC] += (double)((((x[1] > x[0]) ? 1 : 0) == label ? 1.f : 0.f) * mask);
    // centre, direction 1: predicted centre -> nearest GT centre
    const float *c = a.center + (size_t)i * 3;
    float dc = 3.4e38f;
    int jc = 0;
    for (int j = 0; j < a.G; ++j) {
      const float d = sq3(c[0], c[1], c[2], gl[3 * j], gl[3 * j + 1], gl[3 * j + 2]);
      if (d < dc) { dc = d; jc = j; }
    }
    a.scratch[i] = jc;
    if (label) {
      acc[S_C1] += (double)dc;
      const size_t g = (size_t)b * a.G + j1;
      // heading: class CE + Huber on the normalised residual of the labelled bin
      const int hcl = (int)a.heading_class_label[g];
      acc[S_HCLS] += (double)(lse(x + oh, a.NH) - x[oh + hcl]);
      const float hlab = a.heading_residual_label[g] / (3.14159265358979323846f / (float)a.NH);
      acc[S_HREG] += (double)huber(x[ohr + hcl] - hlab);
      // size: class CE + Huber on the normalised residual of the labelled cluster (mean over x, y, z)
      const int scl = (int)a.size_class_label[g];
      acc[S_SCLS] += (double)(lse(x + os, a.NS) - x[os + scl]);
      float hs = 0.f;
      for (int d = 0; d < 3; ++d)
        hs += huber(x[osr + scl * 3 + d] - a.size_residual_label[g * 3 + d] / a.mean_size[scl * 3 + d]);
      acc[S_SREG] += (double)(hs / 3.0f);
      acc[S_SEM] += (double)(lse(x + osem, a.NC) - x[osem + (int)a.sem_cls_label[g]]);
    }
  }
  // ---- centre, direction 2: every GT centre -> nearest predicted centre
  for (int i = tid; i < BG; i += kThreads) {
    const int b = i / a.G;
    const float *g = a.center_label + (size_t)i * 3;
    const float *cs = a.center + (size_t)b * a.K * 3;
    float d2 = 3.4e38f;
    int k2 = 0;
    for (int k = 0; k < a.K; ++k) {
      const float d = sq3(cs[3 * k], cs[3 * k + 1], cs[3 * k + 2], g[0], g[1], g[2]);
      if (d < d2) { d2 = d; k2 = k; }
    }
    a.scratch[BK + i] = k2;
    const float bm = a.box_label_mask[i];
    acc[S_C2] += (double)(d2 * bm);
    acc[S_BMASK] += (double)bm;
  }
  for (int s = 0; s < S_COUNT; ++s) {
    const double v = block_sum(acc[s], red);
    if (tid == 0) tot[s] = v;
  }
  __syncthreads();
  const float dv = (float)(tot[S_VMASK] + 1e-6), dom = (float)(tot[S_OMASK] + 1e-6), dl = (float)(tot[S_LABEL] + 1e-6),
              dbm = (float)(tot[S_BMASK] + 1e-6);
  if (tid == 0) {
    const float vote = (float)tot[S_VOTE] / dv, obj = (float)tot[S_OBJ] / dom;
    const float c1 = (float)tot[S_C1] / dl, c2 = (float)tot[S_C2] / dbm;
    const float hcls = (float)tot[S_HCLS] / dl, hreg = (float)tot[S_HREG] / dl, scls = (float)tot[S_SCLS] / dl,
                sreg = (float)tot[S_SREG] / dl, sem = (float)tot[S_SEM] / dl;
    const float center = c1 + c2;
    const float box = center + 0.1f * hcls + hreg + 0.1f * scls + sreg;
    float *s = a.stats;
    s[0] = vote + 0.5f * obj + box + 0.1f * sem;   // det_loss
    s[1] = vote; s[2] = obj; s[3] = center; s[4] = hcls; s[5] = hreg; s[6] = scls; s[7] = sreg; s[8] = sem; s[9] = box;
    s[10] = (float)tot[S_ACC] / dom;                                   // obj_acc
    s[11] = (float)tot[S_LABEL] / (float)BK;                           // pos_ratio
    s[12] = (float)tot[S_OMASK] / (float)BK - s[11];                   // neg_ratio
    s[13] = s[14] = s[15] = 0.f;
  }

  // ================================= gradients of det_loss =================================
  for (int i = tid; i < BS; i += kThreads) {   // votes
    const int b = i / a.S;
    const int ind = a.seed_inds[(size_t)b * a.seed_ld + (i - b * a.S)];
    const float m = (float)a.vote_label_mask[(size_t)b * a.N + ind];
    const float *gt = a.vote_label + ((size_t)b * a.N + ind) * 9;
    const float *sx = a.seed_xyz + (size_t)i * 3, *vx = a.vote_xyz + (size_t)i * 3;
    float best = 3.4e38f;
    int jb = 0;
    for (int j = 0; j < 3; ++j) {
      float d = 0.f;
      for (int c = 0; c < 3; ++c) d = __fadd_rn(d, fabsf(__fsub_rn(vx[c], __fadd_rn(gt[3 * j + c], sx[c]))));
      if (d < best) { best = d; jb = j; }
    }
    const float w = m / dv;
    for (int c = 0; c < 3; ++c) {
      const float diff = __fsub_rn(vx[c], __fadd_rn(gt[3 * jb + c], sx[c]));
      a.d_vote_xyz[(size_t)i * 3 + c] = w * (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f));
    }
  }
  for (int i = tid; i < BK; i += kThreads) {   // head outputs and predicted centres (direction 1)
    const int b = i / a.K;
    const float *x = a.net + (size_t)i * a.W;
    float *dx = a.d_net + (size_t)i * a.W;
    for (int q = 0; q < a.W; ++q) dx[q] = 0.f;
    const int label = (int)a.objectness_label[i];
    const float mask = a.objectness_mask[i];
    {  // objectness
      const float l2 = lse(x, 2);
      const float w = 0.5f * (label ? 0.8f : 0.2f) * mask / dom;
      dx[0] = w * (expf(x[0] - l2) - (label == 0 ? 1.f : 0.f));
      dx[1] = w * (expf(x[1] - l2) - (label == 1 ? 1.f : 0.f));
    }
    float *dc = a.d_center + (size_t)i * 3;
    dc[0] = dc[1] = dc[2] = 0.f;
    if (label) {
      const float w = 1.0f / dl;
      const float *c = a.center + (size_t)i * 3;
      const float *g1 = a.center_label + ((size_t)b * a.G + a.scratch[i]) * 3;
      for (int d = 0; d < 3; ++d) dc[d] = w * 2.0f * __fsub_rn(c[d], g1[d]);
      const size_t g = (size_t)b * a.G + (int)a.object_assignment[i];
      const int hcl = (int)a.heading_class_label[g];
      const float lh = lse(x + oh, a.NH);
      for (int h = 0; h < a.NH; ++h) dx[oh + h] = 0.1f * w * (expf(x[oh + h] - lh) - (h == hcl ? 1.f : 0.f));
      const float hlab = a.heading_residual_label[g] / (3.14159265358979323846f / (float)a.NH);
      dx[ohr + hcl] = w * huber_grad(x[ohr + hcl] - hlab);
      const int scl = (int)a.size_class_label[g];
      const float ls = lse(x + os, a.NS);
      for (int q = 0; q < a.NS; ++q) dx[os + q] = 0.1f * w * (expf(x[os + q] - ls) - (q == scl ? 1.f : 0.f));
      for (int d = 0; d < 3; ++d)
        dx[osr + scl * 3 + d] = w * huber_grad(x[osr + scl * 3 + d] - a.size_residual_label[g * 3 + d] / a.mean_size[scl * 3 + d]) / 3.0f;
      const int sem = (int)a.sem_cls_label[g];
      const float lm = lse(x + osem, a.NC);
      for (int q = 0; q < a.NC; ++q) dx[osem + q] = 0.1f * w * (expf(x[osem + q] - lm) - (q == sem ? 1.f : 0.f));
    }
  }
  __syncthreads();
  for (int i = tid; i < BG; i += kThreads) {   // predicted centres (direction 2): scatter onto the arg-min proposal
    const float bm = a.box_label_mask[i];
    if (bm == 0.f) continue;
    const int b = i / a.G;
    const int k2 = a.scratch[BK + i];
    const float *g = a.center_label + (size_t)i * 3;
    const float *c = a.center + ((size_t)b * a.K + k2) * 3;
    const float w = bm / dbm;
    for (int d = 0; d < 3; ++d) atomicAdd(a.d_center + ((size_t)b * a.K + k2) * 3 + d, w * 2.0f * __fsub_rn(c[d], g[d]));
  }
}

}  // namespace
}  // namespace s2c

extern "C" int s2c_detection_loss(int B, int S, int N, int K, int G, int NH, int NS, int NC, const float *vote_xyz,
                                  const float *seed_xyz, const int *seed_inds, long long seed_ld, const float *vote_label,
                                  const long long *vote_label_mask, const float *agg_xyz, const float *net,
                                  const float *center, const float *center_label, const long long *heading_class_label,
                                  const float *heading_residual_label, const long long *size_class_label,
                                  const float *size_residual_label, const long long *sem_cls_label,
                                  const float *box_label_mask, const float *mean_size, float *stats,
                                  long long *objectness_label, float *objectness_mask, long long *object_assignment,
                                  float *d_vote_xyz, float *d_net, float *d_center, int *scratch, void *stream) {
  using namespace s2c;
  S2C_REQUIRE(B >= 1 && S >= 1 && N >= 1 && K >= 1 && G >= 1, "detection_loss: bad sizes");
  S2C_REQUIRE(NH >= 1 && NH <= 64 && NS >= 1 && NS <= 256 && NC >= 1 && NC <= 256, "detection_loss: bad class counts");
  S2C_REQUIRE(vote_xyz && seed_xyz && seed_inds && vote_label && vote_label_mask && agg_xyz && net && center &&
                  center_label && heading_class_label && heading_residual_label && size_class_label &&
                  size_residual_label && sem_cls_label && box_label_mask && mean_size,
              "detection_loss: null input");
  S2C_REQUIRE(stats && objectness_label && objectness_mask && object_assignment && d_vote_xyz && d_net && d_center && scratch,
              "detection_loss: null output");
  S2C_REQUIRE(seed_ld >= S, "detection_loss: seed_ld=%lld < S=%d", seed_ld, S);
  LossArgs a;
  a.B = B; a.S = S; a.N = N; a.K = K; a.G = G; a.NH = NH; a.NS = NS; a.NC = NC; a.W = 2 + 3 + 2 * NH + 4 * NS + NC;
  a.vote_xyz = vote_xyz; a.seed_xyz = seed_xyz; a.seed_inds = seed_inds; a.seed_ld = seed_ld; a.vote_label = vote_label;
  a.vote_label_mask = vote_label_mask; a.agg_xyz = agg_xyz; a.net = net; a.center = center; a.center_label = center_label;
  a.heading_class_label = heading_class_label; a.heading_residual_label = heading_residual_label;
  a.size_class_label = size_class_label; a.size_residual_label = size_residual_label; a.sem_cls_label = sem_cls_label;
  a.box_label_mask = box_label_mask; a.mean_size = mean_size; a.stats = stats; a.objectness_label = objectness_label;
  a.objectness_mask = objectness_mask; a.object_assignment = object_assignment; a.d_vote_xyz = d_vote_xyz; a.d_net = d_net;
  a.d_center = d_center; a.scratch = scratch;
  detection_loss_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(a);
  S2C_CHECK_LAUNCH("detection_loss");
  return S2C_OK;
}
