// out (M, N) = A^T X over a SHORT reduction (R rows = T*B of the caption decoder, a few hundred): the weight gradients
// dW = dGates^T * Inputs after the decoder's backward recurrence (models/caption_module.py:428-500 under autograd).
// cuBLAS' SIMT heuristics pick split-K kernels with < 50 CTAs for these shapes (85 us for 0.2 GFLOP); a plain
// 64x64-tile kernel with one CTA per output tile is latency-bound at ~6 us.  Optionally also emits the column sums
// of A (the bias gradients) from the first tile column.
#include "s2c_common.cuh"

namespace s2c {
namespace {

constexpr int TM = 64, TN = 64, TR = 64;  // 64 reduction rows per stage: few, long stages (each pays one L2 round trip)

__global__ void __launch_bounds__(256)
gemm_tn_kernel(const float *__restrict__ A, long long lda, const float *__restrict__ X, long long ldx, int R, int M, int N,
               float *__restrict__ out, long long ldo, float *__restrict__ colsum) {
  __shared__ float As[TR][TM + 4], Xs[TR][TN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float cs[4] = {0.f, 0.f, 0.f, 0.f};
  // gridDim.z > 1: this CTA reduces rows [r_begin, r_end) and ADDS its partial to out / colsum (zero-filled by the host)
  const int rows_per_split = ((R + (int)gridDim.z - 1) / (int)gridDim.z + TR - 1) / TR * TR;
  const int r_begin = blockIdx.z * rows_per_split, r_end = min(R, r_begin + rows_per_split);
  for (int r0 = r_begin; r0 < r_end; r0 += TR) {
#pragma unroll 8
    for (int i = threadIdx.x; i < TR * TM; i += 256) {
      const int rr = i / TM, c = i - rr * TM;
      As[rr][c] = (r0 + rr < r_end && m0 + c < M) ? A[(size_t)(r0 + rr) * lda + m0 + c] : 0.f;
      Xs[rr][c] = (r0 + rr < r_end && n0 + c < N) ? X[(size_t)(r0 + rr) * ldx + n0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 16
    for (int rr = 0; rr < TR; ++rr) {
      const float4 a = *reinterpret_cast<const float4 *>(&As[rr][ty * 4]);
      const float4 x = *reinterpret_cast<const float4 *>(&Xs[rr][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        cs[i] += av[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], xv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) {
        if (gridDim.z == 1) out[(size_t)m * ldo + n] = acc[i][j];
        else atomicAdd(out + (size_t)m * ldo + n, acc[i][j]);
      }
    }
    if (colsum != nullptr && blockIdx.x == 0 && tx == 0) {
      if (gridDim.z == 1) colsum[m] = cs[i];
      else atomicAdd(colsum + m, cs[i]);
    }
  }
}

}  // namespace
}  // namespace s2c

extern "C" int s2c_gemm_tn(const float *A, long long lda, const float *X, long long ldx, int R, int M, int N, float *out,
                           long long ldo, float *colsum, void *stream) {
  using namespace s2c;
  S2C_REQUIRE(R >= 0 && M >= 0 && N >= 0 && lda >= M && ldx >= N && ldo >= N, "gemm_tn: bad sizes");
  if (M == 0 || N == 0) return S2C_OK;
  S2C_REQUIRE(A && X && out, "gemm_tn: null pointer");
  dim3 grid((unsigned)ceil_div(N, TN), (unsigned)ceil_div(M, TM));
  // few output tiles and a long reduction (map_feat's weight gradient: 16 tiles, 2048 rows): split the rows over
  // gridDim.z so about two CTAs per SM exist, each with at least two 64-row stages; partials meet through atomics
  const int tiles = (int)(grid.x * grid.y);
  int splits = ceil_div(2 * kNumSMs, tiles);
  if (splits > R / (2 * TR)) splits = R / (2 * TR);
  if (splits > 1) {
    grid.z = (unsigned)splits;
    cudaStream_t st = (cudaStream_t)stream;
    if (ldo == N) {
      S2C_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)M * N, st), "gemm_tn memset");
    } else {
      S2C_CUDA(cudaMemset2DAsync(out, sizeof(float) * (size_t)ldo, 0, sizeof(float) * (size_t)N, (size_t)M, st), "gemm_tn memset");
    }
    if (colsum) S2C_CUDA(cudaMemsetAsync(colsum, 0, sizeof(float) * (size_t)M, st), "gemm_tn memset");
  }
  gemm_tn_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, X, ldx, R, M, N, out, ldo, colsum);
  S2C_CHECK_LAUNCH("gemm_tn");
  return S2C_OK;
}

// C (M, N) = A (M, K) * B (K, N) [+ bias[n]] [relu] in plain fp32 FMAs, both operands as generic strided views:
//     A(m, k) = A[m * sam + k * sak],   B(k, n) = B[k * sbk + n * sbn]
// so the same kernel is the Linear forward (x W^T: B(k, n) = W[n * ldw + k]) and its input gradient (dY W:
// B(k, n) = W[k * ldw + n]) for the caption module's nn.Linear layers (models/caption_module.py:216-240: map_feat,
// the hoisted word / target terms of map_topdown, classifier), whose widths (300, 812, 3500) are not multiples of the
// tensor-core kernels' 64-column tiles.  A few hundred rows x a few thousand columns: one 64x128 tile per CTA, the
// library's SIMT sgemm kernels these replace take 28-52 us per call here.
namespace s2c {
namespace {

constexpr int GK = 32, GM = 64, GN = 128;  // 64 x 128 tile, 256 threads, 4 x 8 outputs per thread (32 FMAs per 3 LDS.128)

__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float *__restrict__ A, long long sam, long long sak, const float *__restrict__ B, long long sbk,
                 long long sbn, const float *__restrict__ bias, int relu, int M, int N, int K, float *__restrict__ C,
                 long long ldc) {
  __shared__ __align__(16) float As[GK][GM + 4], Bs[GK][GN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  const bool a_kfast = sak == 1, b_kfast = sbk == 1;  // which index is contiguous in memory: coalesce along it
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  // gridDim.z > 1: split K; this CTA reduces [k_begin, k_end) and ADDS its partial to C (zero-filled by the host)
  const int k_per_split = ((K + (int)gridDim.z - 1) / (int)gridDim.z + GK - 1) / GK * GK;
  const int k_begin = blockIdx.z * k_per_split, k_end = min(K, k_begin + k_per_split);
  // register double buffering: the global loads of stage s + 1 are in flight while stage s is multiplied out of shared
  // memory (every stage used to expose one full L2 / HBM round trip: 4 us of the 67 us classifier forward were math)
  constexpr int LA = GK * GM / 256, LB = GK * GN / 256;
  float ra[LA], rb[LB];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int u = 0; u < LA; ++u) {
      const int i = threadIdx.x + u * 256;
      const int ka = a_kfast ? (i % GK) : (i / GM), ma = a_kfast ? (i / GK) : (i % GM);
      ra[u] = (k0 + ka < k_end && m0 + ma < M) ? A[(long long)(m0 + ma) * sam + (long long)(k0 + ka) * sak] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < LB; ++u) {
      const int i = threadIdx.x + u * 256;
      const int kb = b_kfast ? (i % GK) : (i / GN), nb = b_kfast ? (i / GK) : (i % GN);
      rb[u] = (k0 + kb < k_end && n0 + nb < N) ? B[(long long)(k0 + kb) * sbk + (long long)(n0 + nb) * sbn] : 0.f;
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int u = 0; u < LA; ++u) {
      const int i = threadIdx.x + u * 256;
      const int ka = a_kfast ? (i % GK) : (i / GM), ma = a_kfast ? (i / GK) : (i % GM);
      As[ka][ma] = ra[u];
    }
#pragma unroll
    for (int u = 0; u < LB; ++u) {
      const int i = threadIdx.x + u * 256;
      const int kb = b_kfast ? (i % GK) : (i / GN), nb = b_kfast ? (i / GK) : (i % GN);
      Bs[kb][nb] = rb[u];
    }
  };
  if (k_begin < k_end) fetch(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += GK) {
    stage();
    __syncthreads();
    if (k0 + GK < k_end) fetch(k0 + GK);
#pragma unroll 8
    for (int kk = 0; kk < GK; ++kk) {
      const float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);        // columns tx*4 .. +3
      const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[kk][64 + tx * 4]);   // columns 64 + tx*4 .. +3
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j >> 2) * 64 + tx * 4 + (j & 3);
      if (n >= N) continue;
      float v = acc[i][j] + ((bias != nullptr && blockIdx.z == 0) ? bias[n] : 0.f);
      if (gridDim.z == 1) {
        if (relu) v = fmaxf(v, 0.f);
        C[(long long)m * ldc + n] = v;
      } else {
        atomicAdd(C + (long long)m * ldc + n, v);
      }
    }
  }
}

// N <= 8 output columns with A rows contiguous in k (the coordinate columns of a first layer's input gradient,
// lib/pointnet2/fused_mlp.py: (32 768 x 128) x (128 x 4)): one warp per row, lanes stride k, B in shared memory
__global__ void __launch_bounds__(256)
gemm_skinny_kernel(const float *__restrict__ A, long long sam, const float *__restrict__ B, long long sbk, long long sbn,
                   const float *__restrict__ bias, int M, int N, int K, float *__restrict__ C, long long ldc) {
  extern __shared__ float Bsm[];  // [K][8]
  for (int i = threadIdx.x; i < K * 8; i += 256) {
    const int k = i >> 3, n = i & 7;
    Bsm[i] = n < N ? B[(long long)k * sbk + (long long)n * sbn] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * 256) >> 5;
  for (long long m = warp; m < M; m += nwarps) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float *a = A + m * sam;
    for (int k = lane; k < K; k += 32) {
      const float v = __ldg(a + k);
      const float4 b0 = *reinterpret_cast<const float4 *>(Bsm + k * 8), b1 = *reinterpret_cast<const float4 *>(Bsm + k * 8 + 4);
      acc[0] = fmaf(v, b0.x, acc[0]); acc[1] = fmaf(v, b0.y, acc[1]); acc[2] = fmaf(v, b0.z, acc[2]); acc[3] = fmaf(v, b0.w, acc[3]);
      acc[4] = fmaf(v, b1.x, acc[4]); acc[5] = fmaf(v, b1.y, acc[5]); acc[6] = fmaf(v, b1.z, acc[6]); acc[7] = fmaf(v, b1.w, acc[7]);
    }
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
    if (lane < N) {
      float v = acc[0];
#pragma unroll
      for (int n = 1; n < 8; ++n) v = lane == n ? acc[n] : v;
      C[m * ldc + lane] = v + (bias != nullptr ? bias[lane] : 0.f);
    }
  }
}

}  // namespace
}  // namespace s2c

extern "C" int s2c_gemm(const float *A, long long sam, long long sak, const float *B, long long sbk, long long sbn,
                        const float *bias, int relu, int M, int N, int K, float *C, long long ldc, void *stream) {
  using namespace s2c;
  S2C_REQUIRE(M >= 0 && N >= 0 && K >= 0 && ldc >= N, "gemm: bad sizes");
  if (M == 0 || N == 0) return S2C_OK;
  S2C_REQUIRE(C && (K == 0 || (A && B)), "gemm: null pointer");
  if (N <= 8 && sak == 1 && !relu && K >= 1 && K <= 1024) {
    const long long want = ((long long)M + 7) / 8;
    const int blocks = (int)(want < 8LL * kNumSMs ? want : 8LL * kNumSMs);
    gemm_skinny_kernel<<<blocks, 256, (size_t)K * 8 * sizeof(float), (cudaStream_t)stream>>>(A, sam, B, sbk, sbn, bias, M, N, K, C, ldc);
    S2C_CHECK_LAUNCH("gemm (skinny)");
    return S2C_OK;
  }
  dim3 grid((unsigned)ceil_div(N, GN), (unsigned)ceil_div(M, GM));
  S2C_REQUIRE(grid.y <= 65535u, "gemm: M=%d exceeds 65535 row tiles", M);
  // few output tiles and a long reduction (the classifier's input gradient: 16 tiles, K = 3500): split K over
  // gridDim.z (about two CTAs per SM, at least four 32-wide stages each); partials meet through atomics
  const int tiles = (int)(grid.x * grid.y);
  int splits = (relu || 2 * tiles >= kNumSMs) ? 1 : ceil_div(2 * kNumSMs, tiles);  // (a large output pays for its atomics)
  if (splits > K / (4 * GK)) splits = K / (4 * GK);
  if (splits > 1) {
    grid.z = (unsigned)splits;
    cudaStream_t st = (cudaStream_t)stream;
    if (ldc == N) {
      S2C_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, st), "gemm memset");
    } else {
      S2C_CUDA(cudaMemset2DAsync(C, sizeof(float) * (size_t)ldc, 0, sizeof(float) * (size_t)N, (size_t)M, st), "gemm memset");
    }
  }
  gemm_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, sam, sak, B, sbk, sbn, bias, relu, M, N, K, C, ldc);
  S2C_CHECK_LAUNCH("gemm");
  return S2C_OK;
}

// Column sums of a tall matrix (bias gradients of the Linear / Conv1d layers): out[c] = sum_r A[r, c].
namespace s2c {
namespace {
__global__ void __launch_bounds__(256)
col_sum_kernel(const float *__restrict__ A, long long lda, long long R, int M, long long rows_per_cta, float *__restrict__ out) {
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = r0 + rows_per_cta < R ? r0 + rows_per_cta : R;
  for (int c = threadIdx.x; c < M; c += 256) {
    float s0 = 0.f, s1 = 0.f;
    long long r = r0;
    for (; r + 1 < r1; r += 2) { s0 += A[r * lda + c]; s1 += A[(r + 1) * lda + c]; }
    if (r < r1) s0 += A[r * lda + c];
    atomicAdd(out + c, s0 + s1);
  }
}
}  // namespace
}  // namespace s2c

extern "C" int s2c_col_sum(const float *A, long long lda, long long R, int M, float *out, void *stream) {
  using namespace s2c;
  S2C_REQUIRE(R >= 0 && M >= 0 && lda >= M, "col_sum: bad sizes");
  if (M == 0) return S2C_OK;
  S2C_REQUIRE(out, "col_sum: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  S2C_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)M, st), "col_sum memset");
  if (R == 0) return S2C_OK;
  S2C_REQUIRE(A, "col_sum: null pointer");
  const long long rows_per_cta = R / (2 * kNumSMs) + 1 > 32 ? R / (2 * kNumSMs) + 1 : 32;
  const unsigned grid = (unsigned)ceil_div_ll(R, rows_per_cta);
  col_sum_kernel<<<grid, 256, 0, st>>>(A, lda, R, M, rows_per_cta, out);
  S2C_CHECK_LAUNCH("col_sum");
  return S2C_OK;
}
