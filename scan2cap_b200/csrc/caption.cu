// Teacher-forced top-down caption decoder (training mode) as ONE persistent kernel per direction.
//
// Replaces the per-step Python loop of TopDownSceneCaptionModule.forward_sample_batch
// (reference models/caption_module.py:428-500, step function :250-292): per word
//     u   = relu(W_td [w_t ; h2 ; target] + b)           (word / target terms are hoisted by the caller)
//     h1  = GRUCell_1(u, h1)
//     p   = softmax_k( attend(tanh(map_feat(obj_k) + map_hidd(h1))) masked by the local-context mask )
//     att = sum_k p_k obj_k ;  l = relu(W_lang [att ; h1] + b) ;  h2 = GRUCell_2(l, h2)
// which the framework path issues as ~28 kernels per word forward and ~75 backward (B = 8 rows each).
//
// Here one thread-block CLUSTER (16 CTAs, 8 if 16 cannot be scheduled) walks all T words: every mat-vec is split
// over the CTAs by output unit, the ~12 MB of weights stream from L2 each step, the B x {300,512} activations are
// exchanged through global memory between cluster barriers (release/acquire), and everything the backward pass
// needs is written once per step.  The backward kernel runs the same recurrence in reverse (transposed weights
// prepared by the caller) and emits the per-step gate gradients; the weight gradients are then plain GEMMs over
// the (T*B)-row stacks, done by the caller.
//
// Mat-vec micro-kernel: a warp owns 4 output rows x 8 batch rows; lanes split K (float4 loads of the weight rows,
// activations broadcast from shared memory); the 32 partial sums are reduced with a halving butterfly (31 shuffles)
// that leaves out[i = lane/8][r = lane%8] in each lane.
#include <stdlib.h>

#define S2C_CAP_NT 512
#include "caption_common.cuh"

namespace s2c {
namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;

// ================================================================== forward
template <int CL>
__global__ void __launch_bounds__(kThreads, 1)
caption_fwd_kernel(const s2c_caption_params P, const int level) {
  extern __shared__ __align__(16) float smem[];
  const int B = P.B, T = P.T, K = P.K, E = P.E, H = P.H, F = P.F;
  const int c = (int)cl_rank();
  const int rb = (blockIdx.x / CL) * kRows;
  const int nb = min(kRows, B - rb);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Slices S = make_slices(c, CL, H, E, F);
  const int XLD = F + H;
  const SmemPlan sp = plan(smem, XLD, S.hs, K, F, H, CL, false, level == 2 ? 3 : level);
  build_valid_lists(sp, P.valid, P.obj, rb, nb, K, F);
  const int li = lane >> 3, lr = lane & 7;  // this lane's (row-in-quad, batch row) after gemv_quad
  // level 2: the (row, proposal) pairs are dealt round-robin to the CTAs; each keeps the map_feat rows of its pairs
  const bool split = level >= 2 && sp.pb[kRows] >= 0;
  const int n_own = split ? (sp.pb[kRows] - c + CL - 1) / CL : 0;
  if (split) {
    const int h4 = H >> 2;
    for (int i = threadIdx.x; i < n_own * h4; i += kThreads) {
      const int o = i / h4, hh = (i - o * h4) * 4;
      const int p = c + o * CL;
      *reinterpret_cast<float4 *>(sp.mcache + (size_t)o * H + hh) = __ldg(reinterpret_cast<const float4 *>(
          P.mapped + ((size_t)(rb + sp.pair_r[p]) * K + sp.pair_k[p]) * H + hh));
    }
    __syncthreads();
  }

  for (int t = 0; t < T; ++t) {
    const size_t tb = (size_t)t * B + rb;  // first (t, b) row of this cluster in the (T,B,.) buffers
    const size_t tb_prev = (size_t)(t - 1) * B + rb;
    // ---- S1: u = relu(pre_word_t + pre_tgt + W_tdh h2)
    stamp(P.dbg_ts, c, t, 0);
    load_rows(sp.XB, XLD, t > 0 ? P.h2 + tb_prev * H : nullptr, H, H, nb);
    __syncthreads();
    for (int q = warp; q * 4 < S.e1 - S.e0; q += kWarps) {
      const int e = S.e0 + q * 4;
      const float v = gemv_quad(S2C_QUAD_PTRS(P.w_tdh, P.ld_tdh, e, S.e1), H, sp.XB, XLD, lane);
      const int ee = e + li;
      if (ee < S.e1 && lr < nb) {
        const float pre = P.pre_word[((size_t)(rb + lr) * T + t) * E + ee] + P.pre_tgt[(size_t)(rb + lr) * E + ee] + v;
        P.u[(tb + lr) * E + ee] = fmaxf(pre, 0.f);
      }
    }
    stamp(P.dbg_ts, c, t, 1);
    cl_sync();
    // ---- S2: GRU cell 1 on (u, h1_prev)
    stamp(P.dbg_ts, c, t, 2);
    load_rows(sp.XA, XLD, P.u + tb * E, E, E, nb);
    load_rows(sp.XB, XLD, t > 0 ? P.h1 + tb_prev * H : nullptr, H, H, nb);
    __syncthreads();
    for (int q = warp; q < 6 * S.hs / 4; q += kWarps) {
      const int vr = q * 4;                 // virtual row: [gate g of W_ih | gate g of W_hh] x hs units
      const int m = vr / (3 * S.hs);        // 0: W_ih (K = E, x = XA), 1: W_hh (K = H, x = XB)
      const int g = (vr - m * 3 * S.hs) / S.hs, jl = vr - m * 3 * S.hs - g * S.hs;
      const int row = g * H + S.j0 + jl;
      const float *W = m ? P.w_hh1 : P.w_ih1;
      const int KK = m ? H : E;
      const float v = gemv_quad(S2C_QUAD_PTRS(W, KK, row, 3 * H), KK, m ? sp.XB : sp.XA, XLD, lane);
      sp.G[(vr + li) * kRows + lr] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S.hs * kRows; i += kThreads) {
      const int jl = i >> 3, r = i & 7, j = S.j0 + jl;
      if (r < nb) {
        const float gir = sp.G[(0 * S.hs + jl) * kRows + r] + P.b_ih1[j], giz = sp.G[(1 * S.hs + jl) * kRows + r] + P.b_ih1[H + j],
                    gin = sp.G[(2 * S.hs + jl) * kRows + r] + P.b_ih1[2 * H + j];
        const float ghr = sp.G[(3 * S.hs + jl) * kRows + r] + P.b_hh1[j], ghz = sp.G[(4 * S.hs + jl) * kRows + r] + P.b_hh1[H + j],
                    ghn = sp.G[(5 * S.hs + jl) * kRows + r] + P.b_hh1[2 * H + j];
        const float rg = sigmoidf_(gir + ghr), zg = sigmoidf_(giz + ghz), ng = tanhf(gin + rg * ghn);
        const float hp = sp.XB[r * XLD + j];
        const float hn = (1.f - zg) * ng + zg * hp;
        const size_t o = (tb + r) * H + j;
        P.r1[o] = rg; P.z1[o] = zg; P.n1[o] = ng; P.hn1[o] = ghn; P.h1[o] = hn;
      }
    }
    stamp(P.dbg_ts, c, t, 3);
    cl_sync();
    // ---- S3: q = W_hidd h1
    load_rows(sp.XB, XLD, P.h1 + tb * H, H, H, nb);
    __syncthreads();
    for (int q = warp; q * 4 < S.hs; q += kWarps) {
      const int j = S.j0 + q * 4;
      const float v = gemv_quad(S2C_QUAD_PTRS(P.w_hidd, H, j, S.j1), H, sp.XB, XLD, lane);
      if (lr < nb) P.q[(tb + lr) * H + j + li] = v;
    }
    cl_sync();
    stamp(P.dbg_ts, c, t, 4);
    // ---- S4: attention over the valid objects (every CTA, redundantly), then l = relu(W_lang [att ; h1] + b)
    load_rows(sp.XA, XLD, P.q + tb * H, H, H, nb);
    __syncthreads();
    {
      if (split) {
        // scores of this CTA's pairs (map_feat rows on chip), published through `scores`; one more cluster barrier
        for (int o = warp; o < n_own; o += kWarps) {
          const int p = c + o * CL, r = sp.pair_r[p];
          const float *mp = sp.mcache + (size_t)o * H;
          float s = 0.f;
          for (int h = lane * 4; h < H; h += 128) {
            const float4 m4 = *reinterpret_cast<const float4 *>(mp + h);
            const float4 qv = *reinterpret_cast<const float4 *>(sp.XA + r * XLD + h);
            const float4 wv = __ldg(reinterpret_cast<const float4 *>(P.w_att + h));
            s = fmaf(tanhf(m4.x + qv.x), wv.x, s); s = fmaf(tanhf(m4.y + qv.y), wv.y, s);
            s = fmaf(tanhf(m4.z + qv.z), wv.z, s); s = fmaf(tanhf(m4.w + qv.w), wv.w, s);
          }
#pragma unroll
          for (int o2 = 16; o2; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
          if (lane == 0 && r < nb) P.scores[(tb + r) * K + sp.pair_k[p]] = s;
        }
        cl_sync();
        for (int p = threadIdx.x; p < sp.pb[kRows]; p += kThreads) {
          const int r = sp.pair_r[p];
          if (r < nb) sp.sc[r * K + p - sp.pb[r]] = __ldcg(P.scores + (tb + r) * K + sp.pair_k[p]);
        }
      } else {
      // scores: one warp per (row, valid object)
      for (int r = 0; r < nb; ++r) {
        const int n = sp.nv[r];
        if (!sp.uniform[r]) {
          for (int i = warp; i < n; i += kWarps) {
            const int k = sp.vk[r * K + i];
            const float *mp = P.mapped + ((size_t)(rb + r) * K + k) * H;
            float s = 0.f;
            for (int h0 = 0; h0 < H; h0 += 512) {  // 4 independent 16-byte loads per lane before any tanh
              float4 m4[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int h = h0 + j * 128 + lane * 4;
                m4[j] = h < H ? __ldg(reinterpret_cast<const float4 *>(mp + h)) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int h = h0 + j * 128 + lane * 4;
                if (h < H) {
                  const float4 qv = *reinterpret_cast<const float4 *>(sp.XA + r * XLD + h);
                  const float4 wv = __ldg(reinterpret_cast<const float4 *>(P.w_att + h));
                  s = fmaf(tanhf(m4[j].x + qv.x), wv.x, s); s = fmaf(tanhf(m4[j].y + qv.y), wv.y, s);
                  s = fmaf(tanhf(m4[j].z + qv.z), wv.z, s); s = fmaf(tanhf(m4[j].w + qv.w), wv.w, s);
                }
              }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) sp.sc[r * K + i] = s;
          }
        }
      }
      __syncthreads();
      }
      __syncthreads();
      if (warp < nb) {  // softmax of row r = warp over its valid list
        const int r = warp, n = sp.nv[r];
        if (sp.uniform[r]) {
          for (int i = lane; i < n; i += 32) sp.probs[r * K + i] = 1.0f / (float)K;
        } else {
          float mx = -3.4e38f;
          for (int i = lane; i < n; i += 32) mx = fmaxf(mx, sp.sc[r * K + i]);
#pragma unroll
          for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          float sum = 0.f;
          for (int i = lane; i < n; i += 32) {
            const float e = expf(sp.sc[r * K + i] - mx);
            sp.sc[r * K + i] = e;
            sum += e;
          }
#pragma unroll
          for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          for (int i = lane; i < n; i += 32) sp.probs[r * K + sp.vk[r * K + i]] = sp.sc[r * K + i] / sum;
        }
      }
      __syncthreads();
      // attended features
      for (int i = threadIdx.x; i < kRows * F; i += kThreads) {
        const int r = i / F, f = i - r * F;
        float a = 0.f;
        if (r < nb) {
          const int n = sp.nv[r];
          for (int ii = 0; ii < n; ++ii)
            a = fmaf(sp.probs[r * K + sp.vk[r * K + ii]], obj_row(sp, P.obj, rb, r, ii, K, F)[f], a);
        }
        sp.att[i] = a;
      }
      __syncthreads();
      // x = [att ; h1] (h1 is still in XB), and the step's attention outputs (written once, by CTA 0)
      for (int i = threadIdx.x; i < kRows * (F + H); i += kThreads) {
        const int r = i / (F + H), cc = i - r * (F + H);
        sp.XA[r * XLD + cc] = cc < F ? sp.att[r * F + cc] : sp.XB[r * XLD + cc - F];
      }
      if (c == 0) {
        for (int i = threadIdx.x; i < nb * K; i += kThreads) P.probs[tb * K + i] = sp.probs[i];
        for (int i = threadIdx.x; i < nb * F; i += kThreads) P.att[tb * F + i] = sp.att[i];
      }
      __syncthreads();
    }
    stamp(P.dbg_ts, c, t, 5);
    for (int q = warp; q * 4 < S.e1 - S.e0; q += kWarps) {
      const int e = S.e0 + q * 4;
      const float v = gemv_quad(S2C_QUAD_PTRS(P.w_lang, F + H, e, S.e1), F + H, sp.XA, XLD, lane);
      const int ee = e + li;
      if (ee < S.e1 && lr < nb) P.lang[(tb + lr) * E + ee] = fmaxf(v + P.b_lang[ee], 0.f);
    }
    cl_sync();
    stamp(P.dbg_ts, c, t, 6);
    // ---- S5: GRU cell 2 on (l, h2_prev)
    load_rows(sp.XA, XLD, P.lang + tb * E, E, E, nb);
    load_rows(sp.XB, XLD, t > 0 ? P.h2 + tb_prev * H : nullptr, H, H, nb);
    __syncthreads();
    for (int q = warp; q < 6 * S.hs / 4; q += kWarps) {
      const int vr = q * 4;
      const int m = vr / (3 * S.hs);
      const int g = (vr - m * 3 * S.hs) / S.hs, jl = vr - m * 3 * S.hs - g * S.hs;
      const int row = g * H + S.j0 + jl;
      const float *W = m ? P.w_hh2 : P.w_ih2;
      const int KK = m ? H : E;
      const float v = gemv_quad(S2C_QUAD_PTRS(W, KK, row, 3 * H), KK, m ? sp.XB : sp.XA, XLD, lane);
      sp.G[(vr + li) * kRows + lr] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S.hs * kRows; i += kThreads) {
      const int jl = i >> 3, r = i & 7, j = S.j0 + jl;
      if (r < nb) {
        const float gir = sp.G[(0 * S.hs + jl) * kRows + r] + P.b_ih2[j], giz = sp.G[(1 * S.hs + jl) * kRows + r] + P.b_ih2[H + j],
                    gin = sp.G[(2 * S.hs + jl) * kRows + r] + P.b_ih2[2 * H + j];
        const float ghr = sp.G[(3 * S.hs + jl) * kRows + r] + P.b_hh2[j], ghz = sp.G[(4 * S.hs + jl) * kRows + r] + P.b_hh2[H + j],
                    ghn = sp.G[(5 * S.hs + jl) * kRows + r] + P.b_hh2[2 * H + j];
        const float rg = sigmoidf_(gir + ghr), zg = sigmoidf_(giz + ghz), ng = tanhf(gin + rg * ghn);
        const float hp = sp.XB[r * XLD + j];
        const float hn = (1.f - zg) * ng + zg * hp;
        const size_t o = (tb + r) * H + j;
        P.r2[o] = rg; P.z2[o] = zg; P.n2[o] = ng; P.hn2[o] = ghn; P.h2[o] = hn;
      }
    }
    cl_sync();
  }
}

// ================================================================== backward
// Transposed weights (row-major): wt_tdh (H,E), wt_ih* (E,3H), wt_hh* (H,3H), wt_hidd (H,H), wt_lang (F+H,E).
template <int CL>
__global__ void __launch_bounds__(kThreads, 1)
caption_bwd_kernel(const s2c_caption_params P, const int level) {
  extern __shared__ __align__(16) float smem[];
  const int B = P.B, T = P.T, K = P.K, E = P.E, H = P.H, F = P.F;
  const int c = (int)cl_rank();
  const int cid = blockIdx.x / CL;
  const int rb = cid * kRows;
  const int nb = min(kRows, B - rb);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Slices S = make_slices(c, CL, H, E, F);
  const int XLD = 3 * H;
  const SmemPlan sp = plan(smem, XLD, S.hs, K, F, H, CL, true, level == 2 ? 3 : level);
  // carried gradients of this CTA's hidden units: acc1/acc2 [hs][8] live in G's tail? -> dedicated arrays in G:
  // G layout here: [0, hs*8): dh1 carried / total, [hs*8, 2*hs*8): dh2 carried / total, [2*hs*8, 3*hs*8): scratch
  float *d1 = sp.G, *d2 = sp.G + S.hs * kRows, *dsv = sp.sc;  // dsv[r][i]: d score of (row, i-th valid object)
  build_valid_lists(sp, P.valid, P.obj, rb, nb, K, F);
  for (int i = threadIdx.x; i < 2 * S.hs * kRows; i += kThreads) sp.G[i] = 0.f;
  // level 2: map_feat values of all valid pairs for this CTA's hidden units, and their gradient accumulators, on chip
  const bool mc_on = level >= 2 && sp.pb[kRows] >= 0;
  if (mc_on) {
    for (int i = threadIdx.x; i < sp.pb[kRows] * S.hs; i += kThreads) {
      const int p = i / S.hs, jl = i - p * S.hs;
      sp.mcache[i] = __ldg(P.mapped + ((size_t)(rb + sp.pair_r[p]) * K + sp.pair_k[p]) * H + S.j0 + jl);
      sp.dmacc[i] = 0.f;
    }
  }
  __syncthreads();
  const int li = lane >> 3, lr = lane & 7;
  float dwatt = 0.f;  // thread (jl = tid/8, r = tid%8) of the first hs*8 threads: partial d w_att[j0+jl]

  for (int t = T - 1; t >= 0; --t) {
    const size_t tb = (size_t)t * B + rb;
    const size_t tb_prev = (size_t)(t - 1) * B + rb;
    // ---- B1: GRU cell 2 backward (element-wise, own hidden units)
    for (int i = threadIdx.x; i < S.hs * kRows; i += kThreads) {
      const int jl = i >> 3, r = i & 7, j = S.j0 + jl;
      if (r < nb) {
        const size_t o = (tb + r) * H + j;
        const float dh = P.d_h2[o] + d2[i];
        const float rg = P.r2[o], zg = P.z2[o], ng = P.n2[o], hn = P.hn2[o];
        const float hp = t > 0 ? P.h2[(tb_prev + r) * H + j] : 0.f;
        const float dn = dh * (1.f - zg), dz = dh * (hp - ng);
        d2[i] = dh * zg;
        const float dnp = dn * (1.f - ng * ng), dzp = dz * zg * (1.f - zg), drp = dnp * hn * rg * (1.f - rg);
        const size_t g = (tb + r) * 3 * H + j;
        P.dgi2[g] = drp; P.dgi2[g + H] = dzp; P.dgi2[g + 2 * H] = dnp;
        P.dgh2[g] = drp; P.dgh2[g + H] = dzp; P.dgh2[g + 2 * H] = dnp * rg;
      }
    }
    cl_sync();
    // ---- B2: d l_pre = (W_ih2^T dgi2) * [l > 0]   and   dh2 += W_hh2^T dgh2 (own units)
    load_rows(sp.XA, XLD, P.dgi2 + tb * 3 * H, 3 * H, 3 * H, nb);
    load_rows(sp.XB, XLD, P.dgh2 + tb * 3 * H, 3 * H, 3 * H, nb);
    __syncthreads();
    {
      const int nq_e = (S.e1 - S.e0 + 3) / 4, nq_h = S.hs / 4;
      for (int q = warp; q < nq_e + nq_h; q += kWarps) {
        if (q < nq_e) {
          const int e = S.e0 + q * 4;
          const float v = gemv_quad(S2C_QUAD_PTRS(P.wt_ih2, 3 * H, e, S.e1), 3 * H, sp.XA, XLD, lane);
          const int ee = e + li;
          if (ee < S.e1 && lr < nb) {
            const size_t o = (tb + lr) * E + ee;
            P.dlang[o] = P.lang[o] > 0.f ? v : 0.f;
          }
        } else {
          const int jl = (q - nq_e) * 4;
          const float v = gemv_quad(S2C_QUAD_PTRS(P.wt_hh2, 3 * H, S.j0 + jl, S.j1), 3 * H, sp.XB, XLD, lane);
          d2[(jl + li) * kRows + lr] += v;
        }
      }
    }
    cl_sync();
    // ---- B3: [d att ; d h1] = W_lang^T d l_pre
    load_rows(sp.XA, XLD, P.dlang + tb * E, E, E, nb);
    __syncthreads();
    {
      const int nq_f = (S.f1 - S.f0 + 3) / 4, nq_h = S.hs / 4;
      for (int q = warp; q < nq_f + nq_h; q += kWarps) {
        if (q < nq_f) {
          const int f = S.f0 + q * 4;
          const float v = gemv_quad(S2C_QUAD_PTRS(P.wt_lang, E, f, S.f1), E, sp.XA, XLD, lane);
          if (f + li < S.f1 && lr < nb) P.datt[(tb + lr) * F + f + li] = v;
        } else {
          const int jl = (q - nq_f) * 4;
          const float v = gemv_quad(S2C_QUAD_PTRS(P.wt_lang, E, F + S.j0 + jl, F + S.j1), E, sp.XA, XLD, lane);
          d1[(jl + li) * kRows + lr] += v;
        }
      }
    }
    cl_sync();
    // ---- B4: attention backward
    load_rows(sp.att, F, P.datt + tb * F, F, F, nb);             // d att (8,F)
    for (int i = threadIdx.x; i < nb * K; i += kThreads) sp.probs[i] = __ldcg(P.probs + tb * K + i);
    __syncthreads();
    // dp[r][i] = d att . obj_k (+ d probs) -> sp.sc, per (row, valid object): one warp each
    for (int r = 0; r < nb; ++r) {
      const int n = sp.nv[r];
      for (int i = warp; i < n; i += kWarps) {
        const int k = sp.vk[r * K + i];
        const float *ob = obj_row(sp, P.obj, rb, r, i, K, F);
        float s = 0.f;
        for (int f = lane; f < F; f += 32) s = fmaf(sp.att[r * F + f], ob[f], s);
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sp.sc[r * K + i] = s + (P.d_probs ? P.d_probs[(tb + r) * K + k] : 0.f);
      }
    }
    __syncthreads();
    if (warp < nb) {  // ds = p * (dp - sum p dp)   (zero for a uniform row: its scores were all masked)
      const int r = warp, n = sp.nv[r];
      float dot = 0.f;
      for (int i = lane; i < n; i += 32) dot = fmaf(sp.probs[r * K + sp.vk[r * K + i]], sp.sc[r * K + i], dot);
#pragma unroll
      for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      const bool uni = sp.uniform[r] != 0;
      for (int i = lane; i < n; i += 32) {
        const float p = sp.probs[r * K + sp.vk[r * K + i]];
        dsv[r * K + i] = uni ? 0.f : p * (sp.sc[r * K + i] - dot);
      }
    }
    __syncthreads();
    // own hidden units: d pre-tanh -> d mapped (accumulated over the steps), dq, d w_att
    for (int i = threadIdx.x; i < S.hs * kRows; i += kThreads) {
      const int jl = i >> 3, r = i & 7, j = S.j0 + jl;
      if (r < nb) {
        const float qv = P.q[(tb + r) * H + j], wa = P.w_att[j];
        float dq = 0.f;
        if (!sp.uniform[r] && mc_on) {
          const int n = sp.nv[r], p0 = sp.pb[r];
          for (int ii = 0; ii < n; ++ii) {
            const float cb = tanhf(sp.mcache[(p0 + ii) * S.hs + jl] + qv);
            const float ds = dsv[r * K + ii];
            const float dpre = ds * wa * (1.f - cb * cb);
            sp.dmacc[(p0 + ii) * S.hs + jl] += dpre;
            dq += dpre;
            dwatt = fmaf(ds, cb, dwatt);
          }
        } else if (!sp.uniform[r]) {
          const int n = sp.nv[r];
          for (int i0 = 0; i0 < n; i0 += 4) {  // 4 proposals at a time: all loads issued before the first use
            float mv[4], dm[4];
            size_t mo[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int ii = min(i0 + u, n - 1);
              mo[u] = ((size_t)(rb + r) * K + sp.vk[r * K + ii]) * H + j;
              mv[u] = __ldg(P.mapped + mo[u]);
              dm[u] = P.d_mapped[mo[u]];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (i0 + u < n) {
                const float cb = tanhf(mv[u] + qv);
                const float ds = dsv[r * K + i0 + u];
                const float dpre = ds * wa * (1.f - cb * cb);
                P.d_mapped[mo[u]] = dm[u] + dpre;
                dq += dpre;
                dwatt = fmaf(ds, cb, dwatt);
              }
            }
          }
        }
        P.dq[(tb + r) * H + j] = dq;
      }
    }
    // own feature units: d obj += p * d att
    for (int r = 0; r < nb; ++r) {
      const int n = sp.nv[r], fw = S.f1 - S.f0;
      for (int i = threadIdx.x; i < n * fw; i += kThreads) {
        const int ii = i / fw, f = S.f0 + i - ii * fw;
        const int k = sp.vk[r * K + ii];
        P.d_obj[((size_t)(rb + r) * K + k) * F + f] += sp.probs[r * K + k] * sp.att[r * F + f];
      }
    }
    cl_sync();
    // ---- B5: dh1 += W_hidd^T dq (own units), then B6: GRU cell 1 backward (element-wise, own units)
    load_rows(sp.XA, XLD, P.dq + tb * H, H, H, nb);
    __syncthreads();
    for (int q = warp; q * 4 < S.hs; q += kWarps) {
      const int jl = q * 4;
      const float v = gemv_quad(S2C_QUAD_PTRS(P.wt_hidd, H, S.j0 + jl, S.j1), H, sp.XA, XLD, lane);
      d1[(jl + li) * kRows + lr] += v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S.hs * kRows; i += kThreads) {
      const int jl = i >> 3, r = i & 7, j = S.j0 + jl;
      if (r < nb) {
        const size_t o = (tb + r) * H + j;
        const float dh = d1[i];
        const float rg = P.r1[o], zg = P.z1[o], ng = P.n1[o], hn = P.hn1[o];
        const float hp = t > 0 ? P.h1[(tb_prev + r) * H + j] : 0.f;
        const float dn = dh * (1.f - zg), dz = dh * (hp - ng);
        d1[i] = dh * zg;
        const float dnp = dn * (1.f - ng * ng), dzp = dz * zg * (1.f - zg), drp = dnp * hn * rg * (1.f - rg);
        const size_t g = (tb + r) * 3 * H + j;
        P.dgi1[g] = drp; P.dgi1[g + H] = dzp; P.dgi1[g + 2 * H] = dnp;
        P.dgh1[g] = drp; P.dgh1[g + H] = dzp; P.dgh1[g + 2 * H] = dnp * rg;
      }
    }
    cl_sync();
    // ---- B7: d u_pre = (W_ih1^T dgi1) * [u > 0]   and   dh1 += W_hh1^T dgh1 (own units)
    load_rows(sp.XA, XLD, P.dgi1 + tb * 3 * H, 3 * H, 3 * H, nb);
    load_rows(sp.XB, XLD, P.dgh1 + tb * 3 * H, 3 * H, 3 * H, nb);
    __syncthreads();
    {
      const int nq_e = (S.e1 - S.e0 + 3) / 4, nq_h = S.hs / 4;
      for (int q = warp; q < nq_e + nq_h; q += kWarps) {
        if (q < nq_e) {
          const int e = S.e0 + q * 4;
          const float v = gemv_quad(S2C_QUAD_PTRS(P.wt_ih1, 3 * H, e, S.e1), 3 * H, sp.XA, XLD, lane);
          const int ee = e + li;
          if (ee < S.e1 && lr < nb) {
            const size_t o = (tb + lr) * E + ee;
            P.du[o] = P.u[o] > 0.f ? v : 0.f;
          }
        } else {
          const int jl = (q - nq_e) * 4;
          const float v = gemv_quad(S2C_QUAD_PTRS(P.wt_hh1, 3 * H, S.j0 + jl, S.j1), 3 * H, sp.XB, XLD, lane);
          d1[(jl + li) * kRows + lr] += v;
        }
      }
    }
    cl_sync();
    // ---- B8: dh2 += W_tdh^T d u_pre (own units; consumed by this CTA's B1 of the previous word)
    load_rows(sp.XA, XLD, P.du + tb * E, E, E, nb);
    __syncthreads();
    for (int q = warp; q * 4 < S.hs; q += kWarps) {
      const int jl = q * 4;
      const float v = gemv_quad(S2C_QUAD_PTRS(P.wt_tdh, E, S.j0 + jl, S.j1), E, sp.XA, XLD, lane);
      d2[(jl + li) * kRows + lr] += v;
    }
    __syncthreads();
  }
  if (mc_on) {  // the accumulated d_mapped of the valid pairs, written once
    __syncthreads();
    for (int i = threadIdx.x; i < sp.pb[kRows] * S.hs; i += kThreads) {
      const int p = i / S.hs, jl = i - p * S.hs;
      if (sp.pair_r[p] < nb) P.d_mapped[((size_t)(rb + sp.pair_r[p]) * K + sp.pair_k[p]) * H + S.j0 + jl] += sp.dmacc[i];
    }
  }
  // d w_att: sum the per-(unit,row) partials over the 8 rows; one slot per cluster (summed by the caller)
  if (threadIdx.x < S.hs * kRows) {
    float v = dwatt;
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    if ((threadIdx.x & 7) == 0) P.d_watt[(size_t)cid * H + S.j0 + (threadIdx.x >> 3)] = v;
  }
}

template <int CL, bool BWD>
int launch_caption(const s2c_caption_params &P, cudaStream_t st, bool probe_only) {
  auto kern = BWD ? caption_bwd_kernel<CL> : caption_fwd_kernel<CL>;
  const int xld = BWD ? 3 * P.H : P.F + P.H;
  int level = 2;  // the highest cache level whose shared-memory plan fits
  while (level > 0 && plan_bytes(xld, P.H / CL, P.K, P.F, P.H, CL, BWD, level == 2 ? 3 : level) > 227 * 1024) --level;
  const size_t smem = plan_bytes(xld, P.H / CL, P.K, P.F, P.H, CL, BWD, level == 2 ? 3 : level);
  if (smem > 227 * 1024) return -1;
  S2C_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "caption smem attr");
  if (CL > 8) S2C_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1), "caption cluster attr");
  cudaLaunchConfig_t cfg = {};
  const int nclusters = (P.B + kRows - 1) / kRows;
  cfg.gridDim = dim3((unsigned)(nclusters * CL));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (probe_only) {
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess) { cudaGetLastError(); return -1; }
    return n > 0 ? 0 : -1;
  }
  S2C_CUDA(cudaLaunchKernelEx(&cfg, kern, P, level), BWD ? "caption_decode_bwd launch" : "caption_decode_fwd launch");
  return S2C_OK;
}

int check_params(const s2c_caption_params *P, const char *what) {
  S2C_REQUIRE(P != nullptr, "%s: null params", what);
  S2C_REQUIRE(P->B >= 1 && P->T >= 1 && P->K >= 1, "%s: bad sizes B=%d T=%d K=%d", what, P->B, P->T, P->K);
  S2C_REQUIRE(P->E % 4 == 0 && P->H % 64 == 0 && P->F % 4 == 0 && P->E >= 4 && P->F >= 4,
              "%s: E=%d, F=%d must be multiples of 4 and H=%d a multiple of 64", what, P->E, P->F, P->H);
  S2C_REQUIRE(P->ld_tdh % 4 == 0 && P->ld_tdh >= P->H, "%s: ld_tdh=%lld", what, P->ld_tdh);
  return S2C_OK;
}

// 16-CTA clusters when the device can co-schedule one for THIS shape (non-portable size), else 8.  The answer of the
// occupancy probe depends on the shared-memory plan, i.e. on (H, K, F): it is cached per shape, not per process.
template <bool BWD>
int dispatch(const s2c_caption_params &P, cudaStream_t st) {
  struct Probe { int H, K, F, use16; };
  static thread_local Probe cache[8];
  static thread_local int ncache = 0;
  int use16 = -1;
  for (int i = 0; i < ncache; ++i)
    if (cache[i].H == P.H && cache[i].K == P.K && cache[i].F == P.F) use16 = cache[i].use16;
  if (use16 < 0) {
    use16 = (P.H % (4 * 16) == 0 && launch_caption<16, BWD>(P, st, true) == 0) ? 1 : 0;
    cache[ncache % 8] = Probe{P.H, P.K, P.F, use16};
    ++ncache;
    if (ncache > 8) ncache = 8;
  }
  int rc = -1;
  if (use16) rc = launch_caption<16, BWD>(P, st, false);
  if (rc == -1) rc = launch_caption<8, BWD>(P, st, false);
  if (rc == -1) {
    s2c::set_error("%s: no shared-memory plan fits 227 KB for H=%d, K=%d, F=%d (8- and 16-CTA clusters tried)",
                   BWD ? "caption_decode_bwd" : "caption_decode_fwd", P.H, P.K, P.F);
    return S2C_ERR_UNSUPPORTED;
  }
  return rc;
}

}  // namespace
}  // namespace s2c

namespace s2c {
int caption_grid_launch(const s2c_caption_params &P, bool bwd, unsigned int *bar, cudaStream_t st);  // caption_grid.cu
}
extern "C" int s2c_caption_decode_fwd(const s2c_caption_params *P, void *stream) {
  if (int rc = check_params(P, "caption_decode_fwd")) return rc;
  S2C_REQUIRE(P->pre_word && P->pre_tgt && P->mapped && P->obj && P->valid && P->w_tdh && P->w_ih1 && P->w_hh1 &&
                  P->b_ih1 && P->b_hh1 && P->w_hidd && P->w_att && P->w_lang && P->b_lang && P->w_ih2 && P->w_hh2 &&
                  P->b_ih2 && P->b_hh2,
              "caption_decode_fwd: null input");
  S2C_REQUIRE(P->u && P->h1 && P->r1 && P->z1 && P->n1 && P->hn1 && P->q && P->probs && P->att && P->lang && P->r2 &&
                  P->z2 && P->n2 && P->hn2 && P->h2 && P->scores,
              "caption_decode_fwd: null output");
  if (P->grid_bar != nullptr) {
    const int rc = s2c::caption_grid_launch(*P, false, P->grid_bar, (cudaStream_t)stream);
    if (rc != -1) return rc;
  }
  return dispatch<false>(*P, (cudaStream_t)stream);
}

extern "C" int s2c_caption_decode_bwd(const s2c_caption_params *P, void *stream) {
  if (int rc = check_params(P, "caption_decode_bwd")) return rc;
  S2C_REQUIRE(P->mapped && P->obj && P->valid && P->w_att && P->u && P->h1 && P->r1 && P->z1 && P->n1 && P->hn1 &&
                  P->q && P->probs && P->lang && P->r2 && P->z2 && P->n2 && P->hn2 && P->h2,
              "caption_decode_bwd: null saved tensor");
  S2C_REQUIRE(P->wt_tdh && P->wt_ih1 && P->wt_hh1 && P->wt_hidd && P->wt_lang && P->wt_ih2 && P->wt_hh2,
              "caption_decode_bwd: null transposed weight");
  S2C_REQUIRE(P->d_h2 && P->dgi2 && P->dgh2 && P->dlang && P->datt && P->dq && P->dgi1 && P->dgh1 && P->du &&
                  P->d_mapped && P->d_obj && P->d_watt,
              "caption_decode_bwd: null gradient buffer");
  if (P->grid_bar != nullptr) {
    const int rc = s2c::caption_grid_launch(*P, true, P->grid_bar, (cudaStream_t)stream);
    if (rc != -1) return rc;
  }
  return dispatch<true>(*P, (cudaStream_t)stream);
}
