// Caption decoder recurrence as a PERSISTENT COOPERATIVE GRID (B <= 8 scenes): H/4 CTAs (128 for H = 512), each owning
// 4 hidden units, 4 embedding units and 4 feature units, with ITS rows of every weight matrix resident in shared memory
// for all T words (~105 KB forward, ~121 KB backward) -- the classic persistent-RNN layout.  Per word nothing but the
// B x {300,512} activation vectors moves: they are exchanged through L2 between grid-wide barriers (one atomic counter,
// ld.acquire spin, co-residency guaranteed by the cooperative launch).  Same arithmetic, same saved tensors and the same
// C ABI as the cluster kernels of caption.cu (which stream 12 MB of weights from L2 per word on 16 SMs and remain the
// path for B > 8 or when a cooperative launch is not possible).
//
// Stage structure per word (forward): S1 u | S2 GRU-1 | S3 q | S4a attention scores of this CTA's (scene, proposal) pair
// | S4b softmax, attended features, language MLP | S5 GRU-2 : 6 grid barriers.  Backward: 6 barriers (see caption.cu).
#define S2C_CAP_NT 256
#include "caption_common.cuh"

namespace s2c {
namespace {

constexpr int kGThreads = 256;
constexpr int kGWarps = kGThreads / 32;

__device__ __forceinline__ void grid_sync(unsigned int *counter, unsigned int &target, unsigned int nctas) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nctas;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (++spins > (1u << 25)) __trap();  // watchdog (~10 s): a lost CTA must fail loudly, not hang the GPU
    } while (v < target);
  }
  __syncthreads();
}

// nrows rows of a row-major matrix (row stride ld, K columns) starting at row0 -> shared memory, rows >= rowmax zero
__device__ __forceinline__ void copy_rows(float *dst, const float *W, size_t ld, int row0, int nrows, int rowmax, int K) {
  const int k4 = K >> 2;
  for (int i = threadIdx.x; i < nrows * k4; i += blockDim.x) {
    const int r = i / k4, c = (i - r * k4) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < rowmax) v = __ldg(reinterpret_cast<const float4 *>(W + (size_t)(row0 + r) * ld + c));
    *reinterpret_cast<float4 *>(dst + (size_t)r * K + c) = v;
  }
}
#define QUAD_SM(p, K) (p), (p) + (K), (p) + 2 * (K), (p) + 3 * (K)

// GRU cell element-wise part for this CTA's hs units (G holds [ih r,z,n | hh r,z,n] x hs x 8 rows)
__device__ __forceinline__ void gru_forward(const float *G, int hs, int j0, int H, int nb, const float *b_ih, const float *b_hh,
                                            const float *hprev_s, int xld, size_t tb, float *r_o, float *z_o, float *n_o,
                                            float *hn_o, float *h_o) {
  for (int i = threadIdx.x; i < hs * kRows; i += blockDim.x) {
    const int jl = i >> 3, r = i & 7, j = j0 + jl;
    if (r < nb) {
      const float gir = G[(0 * hs + jl) * kRows + r] + b_ih[j], giz = G[(1 * hs + jl) * kRows + r] + b_ih[H + j],
                  gin = G[(2 * hs + jl) * kRows + r] + b_ih[2 * H + j];
      const float ghr = G[(3 * hs + jl) * kRows + r] + b_hh[j], ghz = G[(4 * hs + jl) * kRows + r] + b_hh[H + j],
                  ghn = G[(5 * hs + jl) * kRows + r] + b_hh[2 * H + j];
      const float rg = sigmoidf_(gir + ghr), zg = sigmoidf_(giz + ghz), ng = tanhf(gin + rg * ghn);
      const float hp = hprev_s[r * xld + j];
      const size_t o = (tb + r) * H + j;
      r_o[o] = rg; z_o[o] = zg; n_o[o] = ng; hn_o[o] = ghn; h_o[o] = (1.f - zg) * ng + zg * hp;
    }
  }
}

// ================================================================== forward
__global__ void __launch_bounds__(kGThreads, 1)
caption_fwd_grid_kernel(const s2c_caption_params P, unsigned int *bar, const int wsm_off) {
  extern __shared__ __align__(16) float smem[];
  const int B = P.B, T = P.T, K = P.K, E = P.E, H = P.H, F = P.F;
  const int CL = gridDim.x, c = blockIdx.x;
  const int nb = B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Slices S = make_slices(c, CL, H, E, F);
  const int hs = S.hs;
  const int XLD = F + H;
  const SmemPlan sp = plan(smem, XLD, hs, K, F, H, CL, false, 3);
  build_valid_lists(sp, P.valid, P.obj, 0, nb, K, F);
  const int li = lane >> 3, lr = lane & 7;
  unsigned int target = 0;
  // ---- this CTA's weight rows -> shared memory
  const int nqe = (S.e1 - S.e0 + 3) / 4, nqh = hs / 4;
  float *w = smem + wsm_off;
  float *s_tdh = w;  w += (size_t)nqe * 4 * H;
  float *s_ih1 = w;  w += (size_t)3 * hs * E;
  float *s_hh1 = w;  w += (size_t)3 * hs * H;
  float *s_hidd = w; w += (size_t)hs * H;
  float *s_lang = w; w += (size_t)nqe * 4 * (F + H);
  float *s_ih2 = w;  w += (size_t)3 * hs * E;
  float *s_hh2 = w;
  copy_rows(s_tdh, P.w_tdh, P.ld_tdh, S.e0, nqe * 4, S.e1, H);
  copy_rows(s_lang, P.w_lang, F + H, S.e0, nqe * 4, S.e1, F + H);
  copy_rows(s_hidd, P.w_hidd, H, S.j0, hs, H, H);
  for (int g = 0; g < 3; ++g) {
    copy_rows(s_ih1 + (size_t)g * hs * E, P.w_ih1, E, g * H + S.j0, hs, 3 * H, E);
    copy_rows(s_hh1 + (size_t)g * hs * H, P.w_hh1, H, g * H + S.j0, hs, 3 * H, H);
    copy_rows(s_ih2 + (size_t)g * hs * E, P.w_ih2, E, g * H + S.j0, hs, 3 * H, E);
    copy_rows(s_hh2 + (size_t)g * hs * H, P.w_hh2, H, g * H + S.j0, hs, 3 * H, H);
  }
  // the (scene, proposal) pairs are dealt round-robin to the CTAs; each keeps the map_feat rows of its pairs
  const bool split = sp.pb[kRows] >= 0;
  const int n_own = split ? (sp.pb[kRows] - c + CL - 1) / CL : 0;
  if (split) {
    const int h4 = H >> 2;
    for (int i = threadIdx.x; i < n_own * h4; i += blockDim.x) {
      const int o = i / h4, hh = (i - o * h4) * 4;
      const int p = c + o * CL;
      *reinterpret_cast<float4 *>(sp.mcache + (size_t)o * H + hh) = __ldg(reinterpret_cast<const float4 *>(
          P.mapped + ((size_t)sp.pair_r[p] * K + sp.pair_k[p]) * H + hh));
    }
  }
  __syncthreads();

  for (int t = 0; t < T; ++t) {
    const size_t tb = (size_t)t * B, tb_prev = (size_t)(t - 1) * B;
    // ---- S1: u = relu(pre_word_t + pre_tgt + W_tdh h2)
    stamp(P.dbg_ts, c, t, 0);
    load_rows<4>(sp.XB, XLD, t > 0 ? P.h2 + tb_prev * H : nullptr, H, H, nb);
    __syncthreads();
    for (int q = warp; q < nqe; q += kGWarps) {
      const float v = gemv_quad<true>(QUAD_SM(s_tdh + (size_t)q * 4 * H, H), H, sp.XB, XLD, lane);
      const int ee = S.e0 + q * 4 + li;
      if (ee < S.e1 && lr < nb)
        P.u[(tb + lr) * E + ee] = fmaxf(P.pre_word[((size_t)lr * T + t) * E + ee] + P.pre_tgt[(size_t)lr * E + ee] + v, 0.f);
    }
    stamp(P.dbg_ts, c, t, 1);
    grid_sync(bar, target, CL);
    stamp(P.dbg_ts, c, t, 2);
    // ---- S2: GRU cell 1 on (u, h1_prev)
    load_rows<4>(sp.XA, XLD, P.u + tb * E, E, E, nb);
    load_rows<4>(sp.XB, XLD, t > 0 ? P.h1 + tb_prev * H : nullptr, H, H, nb);
    __syncthreads();
    for (int q = warp; q < 6 * nqh; q += kGWarps) {
      const int m = q / (3 * nqh), rem = q - m * 3 * nqh;  // rem = g*nqh + quad-in-gate: rows rem*4.. of the [3*hs] block
      const float v = m ? gemv_quad<true>(QUAD_SM(s_hh1 + (size_t)rem * 4 * H, H), H, sp.XB, XLD, lane)
                        : gemv_quad<true>(QUAD_SM(s_ih1 + (size_t)rem * 4 * E, E), E, sp.XA, XLD, lane);
      sp.G[(q * 4 + li) * kRows + lr] = v;
    }
    __syncthreads();
    gru_forward(sp.G, hs, S.j0, H, nb, P.b_ih1, P.b_hh1, sp.XB, XLD, tb, P.r1, P.z1, P.n1, P.hn1, P.h1);
    stamp(P.dbg_ts, c, t, 3);
    grid_sync(bar, target, CL);
    // ---- S3: q = W_hidd h1   (h1 goes to columns [F, F+H) of XA, where the language MLP expects it: x = [att ; h1])
    load_rows<4>(sp.XA + F, XLD, P.h1 + tb * H, H, H, nb);
    __syncthreads();
    for (int q = warp; q < nqh; q += kGWarps) {
      const float v = gemv_quad<true>(QUAD_SM(s_hidd + (size_t)q * 4 * H, H), H, sp.XA + F, XLD, lane);
      if (lr < nb) P.q[(tb + lr) * H + S.j0 + q * 4 + li] = v;
    }
    grid_sync(bar, target, CL);
    stamp(P.dbg_ts, c, t, 4);
    // ---- S4a: attention scores (q in XB)
    load_rows<4>(sp.XB, XLD, P.q + tb * H, H, H, nb);
    __syncthreads();
    if (split) {
      for (int o = warp; o < n_own; o += kGWarps) {
        const int p = c + o * CL, r = sp.pair_r[p];
        const float *mp = sp.mcache + (size_t)o * H;
        float s = 0.f;
        for (int h = lane * 4; h < H; h += 128) {
          const float4 m4 = *reinterpret_cast<const float4 *>(mp + h);
          const float4 qv = *reinterpret_cast<const float4 *>(sp.XB + r * XLD + h);
          const float4 wv = __ldg(reinterpret_cast<const float4 *>(P.w_att + h));
          s = fmaf(tanhf(m4.x + qv.x), wv.x, s); s = fmaf(tanhf(m4.y + qv.y), wv.y, s);
          s = fmaf(tanhf(m4.z + qv.z), wv.z, s); s = fmaf(tanhf(m4.w + qv.w), wv.w, s);
        }
#pragma unroll
        for (int o2 = 16; o2; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
        if (lane == 0 && r < nb) P.scores[(tb + r) * K + sp.pair_k[p]] = s;
      }
    } else {
      // too many valid proposals for the caches: the (scene, proposal) pairs are still dealt round-robin
      int base = 0;
      for (int r = 0; r < nb; ++r) {
        const int n = sp.nv[r];
        if (!sp.uniform[r]) {
          for (int i = warp; i < n; i += kGWarps) {
            if ((base + i) % CL != c) continue;
            const int k = sp.vk[r * K + i];
            const float *mp = P.mapped + ((size_t)r * K + k) * H;
            float s = 0.f;
            for (int h = lane * 4; h < H; h += 128) {
              const float4 m4 = __ldg(reinterpret_cast<const float4 *>(mp + h));
              const float4 qv = *reinterpret_cast<const float4 *>(sp.XB + r * XLD + h);
              const float4 wv = __ldg(reinterpret_cast<const float4 *>(P.w_att + h));
              s = fmaf(tanhf(m4.x + qv.x), wv.x, s); s = fmaf(tanhf(m4.y + qv.y), wv.y, s);
              s = fmaf(tanhf(m4.z + qv.z), wv.z, s); s = fmaf(tanhf(m4.w + qv.w), wv.w, s);
            }
#pragma unroll
            for (int o2 = 16; o2; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
            if (lane == 0) P.scores[(tb + r) * K + k] = s;
          }
        }
        base += n;
      }
    }
    grid_sync(bar, target, CL);
    // ---- S4b: softmax over the valid proposals, attended features, l = relu(W_lang [att ; h1] + b)
    if (split) {  // one (scene, proposal) pair per thread: a single L2 round trip
      for (int p = threadIdx.x; p < sp.pb[kRows]; p += blockDim.x) {
        const int r = sp.pair_r[p];
        if (r < nb) sp.sc[r * K + p - sp.pb[r]] = sp.uniform[r] ? 0.f : __ldcg(P.scores + (tb + r) * K + sp.pair_k[p]);
      }
    } else {
      for (int i = threadIdx.x; i < nb * K; i += blockDim.x) {
        const int r = i / K, ii = i - r * K;
        if (ii < sp.nv[r]) sp.sc[i] = sp.uniform[r] ? 0.f : __ldcg(P.scores + (tb + r) * K + sp.vk[i]);
      }
    }
    __syncthreads();
    if (warp < nb) {
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
    for (int i = threadIdx.x; i < kRows * F; i += blockDim.x) {
      const int r = i / F, f = i - r * F;
      float a = 0.f;
      if (r < nb) {
        const int n = sp.nv[r];
        for (int ii = 0; ii < n; ++ii) a = fmaf(sp.probs[r * K + sp.vk[r * K + ii]], obj_row(sp, P.obj, 0, r, ii, K, F)[f], a);
      }
      sp.XA[r * XLD + f] = a;  // x = [att ; h1]: h1 was placed at columns [F, F+H) in S3
      if (c == 0 && r < nb) P.att[tb * F + i] = a;
    }
    if (c == 0)
      for (int i = threadIdx.x; i < nb * K; i += blockDim.x) P.probs[tb * K + i] = sp.probs[i];
    __syncthreads();
    stamp(P.dbg_ts, c, t, 5);
    for (int q = warp; q < nqe; q += kGWarps) {
      const float v = gemv_quad<true>(QUAD_SM(s_lang + (size_t)q * 4 * (F + H), F + H), F + H, sp.XA, XLD, lane);
      const int ee = S.e0 + q * 4 + li;
      if (ee < S.e1 && lr < nb) P.lang[(tb + lr) * E + ee] = fmaxf(v + P.b_lang[ee], 0.f);
    }
    grid_sync(bar, target, CL);
    stamp(P.dbg_ts, c, t, 6);
    // ---- S5: GRU cell 2 on (l, h2_prev)
    load_rows<4>(sp.XA, XLD, P.lang + tb * E, E, E, nb);
    load_rows<4>(sp.XB, XLD, t > 0 ? P.h2 + tb_prev * H : nullptr, H, H, nb);
    __syncthreads();
    for (int q = warp; q < 6 * nqh; q += kGWarps) {
      const int m = q / (3 * nqh), rem = q - m * 3 * nqh;
      const float v = m ? gemv_quad<true>(QUAD_SM(s_hh2 + (size_t)rem * 4 * H, H), H, sp.XB, XLD, lane)
                        : gemv_quad<true>(QUAD_SM(s_ih2 + (size_t)rem * 4 * E, E), E, sp.XA, XLD, lane);
      sp.G[(q * 4 + li) * kRows + lr] = v;
    }
    __syncthreads();
    gru_forward(sp.G, hs, S.j0, H, nb, P.b_ih2, P.b_hh2, sp.XB, XLD, tb, P.r2, P.z2, P.n2, P.hn2, P.h2);
    grid_sync(bar, target, CL);
  }
}

// GRU cell element-wise backward for this CTA's units: consumes d (total dh of the units), leaves dh * z in d
__device__ __forceinline__ void gru_backward(float *d, const float *extra, int hs, int j0, int H, int nb, size_t tb,
                                             size_t tb_prev, bool first, const float *r_s, const float *z_s, const float *n_s,
                                             const float *hn_s, const float *h_s, float *dgi, float *dgh) {
  for (int i = threadIdx.x; i < hs * kRows; i += blockDim.x) {
    const int jl = i >> 3, r = i & 7, j = j0 + jl;
    if (r < nb) {
      const size_t o = (tb + r) * H + j;
      const float dh = d[i] + (extra ? extra[o] : 0.f);
      const float rg = r_s[o], zg = z_s[o], ng = n_s[o], hn = hn_s[o];
      const float hp = first ? 0.f : h_s[(tb_prev + r) * H + j];
      const float dn = dh * (1.f - zg), dz = dh * (hp - ng);
      d[i] = dh * zg;
      const float dnp = dn * (1.f - ng * ng), dzp = dz * zg * (1.f - zg), drp = dnp * hn * rg * (1.f - rg);
      const size_t g = (tb + r) * 3 * H + j;
      dgi[g] = drp; dgi[g + H] = dzp; dgi[g + 2 * H] = dnp;
      dgh[g] = drp; dgh[g + H] = dzp; dgh[g + 2 * H] = dnp * rg;
    }
  }
}

// ================================================================== backward
__global__ void __launch_bounds__(kGThreads, 1)
caption_bwd_grid_kernel(const s2c_caption_params P, unsigned int *bar, const int wsm_off) {
  extern __shared__ __align__(16) float smem[];
  const int B = P.B, T = P.T, K = P.K, E = P.E, H = P.H, F = P.F;
  const int CL = gridDim.x, c = blockIdx.x;
  const int nb = B;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Slices S = make_slices(c, CL, H, E, F);
  const int hs = S.hs;
  const int XLD = 3 * H;
  // one 8 x 3H activation buffer: XA and XB of a plan with row stride 3H/2 are contiguous
  const SmemPlan sp = plan(smem, XLD / 2, hs, K, F, H, CL, true, 2);
  float *X = sp.XA;
  float *d1 = sp.G, *d2 = sp.G + hs * kRows, *dsv = sp.sc;
  build_valid_lists(sp, P.valid, P.obj, 0, nb, K, F);
  for (int i = threadIdx.x; i < 2 * hs * kRows; i += blockDim.x) sp.G[i] = 0.f;
  const bool mc_on = sp.pb[kRows] >= 0;
  if (mc_on) {
    for (int i = threadIdx.x; i < sp.pb[kRows] * hs; i += blockDim.x) {
      const int p = i / hs, jl = i - p * hs;
      sp.mcache[i] = __ldg(P.mapped + ((size_t)sp.pair_r[p] * K + sp.pair_k[p]) * H + S.j0 + jl);
      sp.dmacc[i] = 0.f;
    }
  }
  // ---- this CTA's rows of the transposed weights -> shared memory
  const int nqe = (S.e1 - S.e0 + 3) / 4, nqf = (S.f1 - S.f0 + 3) / 4, nqh = hs / 4;
  float *w = smem + wsm_off;
  float *s_ih2 = w;   w += (size_t)nqe * 4 * 3 * H;   // rows e of W_ih2^T (E, 3H)
  float *s_hh2 = w;   w += (size_t)hs * 3 * H;        // rows j of W_hh2^T (H, 3H)
  float *s_langf = w; w += (size_t)nqf * 4 * E;       // rows f of W_lang^T (F+H, E)
  float *s_langh = w; w += (size_t)hs * E;            // rows F+j
  float *s_hidd = w;  w += (size_t)hs * H;            // rows j of W_hidd^T
  float *s_ih1 = w;   w += (size_t)nqe * 4 * 3 * H;
  float *s_hh1 = w;   w += (size_t)hs * 3 * H;
  float *s_tdh = w;                                    // rows j of W_tdh^T (H, E)
  copy_rows(s_ih2, P.wt_ih2, 3 * H, S.e0, nqe * 4, S.e1, 3 * H);
  copy_rows(s_ih1, P.wt_ih1, 3 * H, S.e0, nqe * 4, S.e1, 3 * H);
  copy_rows(s_hh2, P.wt_hh2, 3 * H, S.j0, hs, H, 3 * H);
  copy_rows(s_hh1, P.wt_hh1, 3 * H, S.j0, hs, H, 3 * H);
  copy_rows(s_langf, P.wt_lang, E, S.f0, nqf * 4, S.f1, E);
  copy_rows(s_langh, P.wt_lang, E, F + S.j0, hs, F + H, E);
  copy_rows(s_hidd, P.wt_hidd, H, S.j0, hs, H, H);
  copy_rows(s_tdh, P.wt_tdh, E, S.j0, hs, H, E);
  __syncthreads();
  const int li = lane >> 3, lr = lane & 7;
  float dwatt = 0.f;
  unsigned int target = 0;

  for (int t = T - 1; t >= 0; --t) {
    const size_t tb = (size_t)t * B, tb_prev = (size_t)(t - 1) * B;
    // ---- B1: GRU cell 2 backward (own units)
    gru_backward(d2, P.d_h2, hs, S.j0, H, nb, tb, tb_prev, t == 0, P.r2, P.z2, P.n2, P.hn2, P.h2, P.dgi2, P.dgh2);
    grid_sync(bar, target, CL);
    // ---- B2: d l_pre = (W_ih2^T dgi2) * [l > 0]  ;  dh2 += W_hh2^T dgh2 (own units)
    load_rows<4>(X, XLD, P.dgi2 + tb * 3 * H, 3 * H, 3 * H, nb);
    __syncthreads();
    for (int q = warp; q < nqe; q += kGWarps) {
      const float v = gemv_quad<true>(QUAD_SM(s_ih2 + (size_t)q * 4 * 3 * H, 3 * H), 3 * H, X, XLD, lane);
      const int ee = S.e0 + q * 4 + li;
      if (ee < S.e1 && lr < nb) {
        const size_t o = (tb + lr) * E + ee;
        P.dlang[o] = P.lang[o] > 0.f ? v : 0.f;
      }
    }
    __syncthreads();
    load_rows<4>(X, XLD, P.dgh2 + tb * 3 * H, 3 * H, 3 * H, nb);
    __syncthreads();
    for (int q = warp; q < nqh; q += kGWarps) {
      const float v = gemv_quad<true>(QUAD_SM(s_hh2 + (size_t)q * 4 * 3 * H, 3 * H), 3 * H, X, XLD, lane);
      d2[(q * 4 + li) * kRows + lr] += v;
    }
    grid_sync(bar, target, CL);
    // ---- B3: [d att ; d h1] = W_lang^T d l_pre
    load_rows<4>(X, XLD, P.dlang + tb * E, E, E, nb);
    __syncthreads();
    for (int q = warp; q < nqf + nqh; q += kGWarps) {
      if (q < nqf) {
        const float v = gemv_quad<true>(QUAD_SM(s_langf + (size_t)q * 4 * E, E), E, X, XLD, lane);
        const int ff = S.f0 + q * 4 + li;
        if (ff < S.f1 && lr < nb) P.datt[(tb + lr) * F + ff] = v;
      } else {
        const int qq = q - nqf;
        const float v = gemv_quad<true>(QUAD_SM(s_langh + (size_t)qq * 4 * E, E), E, X, XLD, lane);
        d1[(qq * 4 + li) * kRows + lr] += v;
      }
    }
    grid_sync(bar, target, CL);
    // ---- B4: attention backward
    load_rows<4>(sp.att, F, P.datt + tb * F, F, F, nb);
    load_flat(sp.probs, P.probs + tb * K, nb * K);
    __syncthreads();
    for (int r = 0; r < nb; ++r) {
      const int n = sp.nv[r];
      for (int i = warp; i < n; i += kGWarps) {
        const int k = sp.vk[r * K + i];
        const float *ob = P.obj + ((size_t)r * K + k) * F;
        float s = 0.f;
        for (int f = lane; f < F; f += 32) s = fmaf(sp.att[r * F + f], __ldg(ob + f), s);
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sp.sc[r * K + i] = s + (P.d_probs ? P.d_probs[(tb + r) * K + k] : 0.f);
      }
    }
    __syncthreads();
    if (warp < nb) {
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
    for (int i = threadIdx.x; i < hs * kRows; i += blockDim.x) {
      const int jl = i >> 3, r = i & 7, j = S.j0 + jl;
      if (r < nb) {
        const float qv = P.q[(tb + r) * H + j], wa = P.w_att[j];
        float dq = 0.f;
        if (!sp.uniform[r]) {
          const int n = sp.nv[r], p0 = mc_on ? sp.pb[r] : 0;
          for (int ii = 0; ii < n; ++ii) {
            const size_t mo = ((size_t)r * K + sp.vk[r * K + ii]) * H + j;
            const float cb = tanhf((mc_on ? sp.mcache[(p0 + ii) * hs + jl] : __ldg(P.mapped + mo)) + qv);
            const float ds = dsv[r * K + ii];
            const float dpre = ds * wa * (1.f - cb * cb);
            if (mc_on) sp.dmacc[(p0 + ii) * hs + jl] += dpre; else P.d_mapped[mo] += dpre;
            dq += dpre;
            dwatt = fmaf(ds, cb, dwatt);
          }
        }
        P.dq[(tb + r) * H + j] = dq;
      }
    }
    for (int r = 0; r < nb; ++r) {
      const int n = sp.nv[r], fw = S.f1 - S.f0;
      for (int i = threadIdx.x; i < n * fw; i += blockDim.x) {
        const int ii = i / fw, f = S.f0 + i - ii * fw;
        const int k = sp.vk[r * K + ii];
        P.d_obj[((size_t)r * K + k) * F + f] += sp.probs[r * K + k] * sp.att[r * F + f];
      }
    }
    grid_sync(bar, target, CL);
    // ---- B5: dh1 += W_hidd^T dq (own units) ; B6: GRU cell 1 backward (own units)
    load_rows<4>(X, XLD, P.dq + tb * H, H, H, nb);
    __syncthreads();
    for (int q = warp; q < nqh; q += kGWarps) {
      const float v = gemv_quad<true>(QUAD_SM(s_hidd + (size_t)q * 4 * H, H), H, X, XLD, lane);
      d1[(q * 4 + li) * kRows + lr] += v;
    }
    __syncthreads();
    gru_backward(d1, nullptr, hs, S.j0, H, nb, tb, tb_prev, t == 0, P.r1, P.z1, P.n1, P.hn1, P.h1, P.dgi1, P.dgh1);
    grid_sync(bar, target, CL);
    // ---- B7: d u_pre = (W_ih1^T dgi1) * [u > 0]  ;  dh1 += W_hh1^T dgh1 (own units)
    load_rows<4>(X, XLD, P.dgi1 + tb * 3 * H, 3 * H, 3 * H, nb);
    __syncthreads();
    for (int q = warp; q < nqe; q += kGWarps) {
      const float v = gemv_quad<true>(QUAD_SM(s_ih1 + (size_t)q * 4 * 3 * H, 3 * H), 3 * H, X, XLD, lane);
      const int ee = S.e0 + q * 4 + li;
      if (ee < S.e1 && lr < nb) {
        const size_t o = (tb + lr) * E + ee;
        P.du[o] = P.u[o] > 0.f ? v : 0.f;
      }
    }
    __syncthreads();
    load_rows<4>(X, XLD, P.dgh1 + tb * 3 * H, 3 * H, 3 * H, nb);
    __syncthreads();
    for (int q = warp; q < nqh; q += kGWarps) {
      const float v = gemv_quad<true>(QUAD_SM(s_hh1 + (size_t)q * 4 * 3 * H, 3 * H), 3 * H, X, XLD, lane);
      d1[(q * 4 + li) * kRows + lr] += v;
    }
    grid_sync(bar, target, CL);
    // ---- B8: dh2 += W_tdh^T d u_pre (own units; consumed by this CTA's B1 of the previous word)
    load_rows<4>(X, XLD, P.du + tb * E, E, E, nb);
    __syncthreads();
    for (int q = warp; q < nqh; q += kGWarps) {
      const float v = gemv_quad<true>(QUAD_SM(s_tdh + (size_t)q * 4 * E, E), E, X, XLD, lane);
      d2[(q * 4 + li) * kRows + lr] += v;
    }
    __syncthreads();
  }
  if (mc_on) {
    for (int i = threadIdx.x; i < sp.pb[kRows] * hs; i += blockDim.x) {
      const int p = i / hs, jl = i - p * hs;
      if (sp.pair_r[p] < nb) P.d_mapped[((size_t)sp.pair_r[p] * K + sp.pair_k[p]) * H + S.j0 + jl] += sp.dmacc[i];
    }
  }
  if (threadIdx.x < hs * kRows) {  // hs*8 = 32 threads = warp 0: sum the partials of the 8 rows
    float v = dwatt;
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    if ((threadIdx.x & 7) == 0) P.d_watt[S.j0 + (threadIdx.x >> 3)] = v;
  }
}

size_t grid_wsm_floats(bool bwd, int E, int H, int F, int CL) {
  const int hs = H / CL;
  const int es = ((E + CL - 1) / CL + 3) & ~3, fs = ((F + CL - 1) / CL + 3) & ~3;
  if (!bwd) return (size_t)es * H + (size_t)6 * hs * E + (size_t)6 * hs * H + (size_t)hs * H + (size_t)es * (F + H);
  return (size_t)2 * es * 3 * H + (size_t)2 * hs * 3 * H + (size_t)fs * E + (size_t)hs * E + (size_t)hs * H + (size_t)hs * E;
}

}  // namespace

// -> S2C_OK when the persistent grid ran; -1 when it does not apply (shape, shared memory, co-residency): the caller
// then uses the cluster kernels.  `bar`: one zero-initialised unsigned int of device memory.
int caption_grid_launch(const s2c_caption_params &P, bool bwd, unsigned int *bar, cudaStream_t st) {
  const int H = P.H;
  if (bar == nullptr || P.B > kRows || H % 32 != 0) return -1;
  const int CL = H / 4;  // 4 hidden units per CTA
  int dev = 0, sms = 0, coop = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  if (!coop || CL > sms || CL < 1) return -1;
  const int xld = bwd ? 3 * H / 2 : P.F + P.H;
  const int mask = bwd ? 2 : 3;
  const size_t pb = plan_bytes(xld, H / CL, P.K, P.F, H, CL, bwd, mask);
  const size_t plan_fl = (pb - 16) / sizeof(float);
  const size_t smem = (plan_fl + grid_wsm_floats(bwd, P.E, H, P.F, CL)) * sizeof(float) + 16;
  if (smem > 227 * 1024) return -1;
  const void *kern = bwd ? (const void *)caption_bwd_grid_kernel : (const void *)caption_fwd_grid_kernel;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kGThreads, smem) != cudaSuccess || per_sm * sms < CL) {
    cudaGetLastError();
    return -1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)CL);
  cfg.blockDim = dim3(kGThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const int wsm_off = (int)plan_fl;
  cudaError_t e = bwd ? cudaLaunchKernelEx(&cfg, caption_bwd_grid_kernel, P, bar, wsm_off)
                      : cudaLaunchKernelEx(&cfg, caption_fwd_grid_kernel, P, bar, wsm_off);
  if (e != cudaSuccess) return cuda_fail(e, bwd ? "caption_decode_bwd (grid) launch" : "caption_decode_fwd (grid) launch");
  return S2C_OK;
}

}  // namespace s2c
