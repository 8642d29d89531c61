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
  for (int r0 = 0; r0 < R; r0 += TR) {
#pragma unroll 8
    for (int i = threadIdx.x; i < TR * TM; i += 256) {
      const int rr = i / TM, c = i - rr * TM;
      As[rr][c] = (r0 + rr < R && m0 + c < M) ? A[(size_t)(r0 + rr) * lda + m0 + c] : 0.f;
      Xs[rr][c] = (r0 + rr < R && n0 + c < N) ? X[(size_t)(r0 + rr) * ldx + n0 + c] : 0.f;
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
      if (n < N) out[(size_t)m * ldo + n] = acc[i][j];
    }
    if (colsum != nullptr && blockIdx.x == 0 && tx == 0) colsum[m] = cs[i];
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
  gemm_tn_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, X, ldx, R, M, N, out, ldo, colsum);
  S2C_CHECK_LAUNCH("gemm_tn");
  return S2C_OK;
}

// C (M, N) = A (M, K) * B (K, N) [+ bias[n]] [relu] in plain fp32 FMAs, both operands as generic strided views:
//     A(m, k) = A[m * sam + k * sak],   B(k, n) = B[k * sbk + n * sbn]
// so the same kernel is the Linear forward (x W^T: B(k, n) = W[n * ldw + k]) and its input gradient (dY W:
// B(k, n) = W[k * ldw + n]) for the caption module's nn.Linear layers (models/caption_module.py:216-240: map_feat,
// the hoisted word / target terms of map_topdown, classifier), whose widths (300, 812, 3500) are not multiples of the
// tensor-core kernels' 64-column tiles.  A few hundred rows x a few thousand columns: one 64x64 tile per CTA, the
// library's SIMT sgemm kernels these replace take 28-52 us per call here.
namespace s2c {
namespace {

constexpr int GK = 32;

__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float *__restrict__ A, long long sam, long long sak, const float *__restrict__ B, long long sbk,
                 long long sbn, const float *__restrict__ bias, int relu, int M, int N, int K, float *__restrict__ C,
                 long long ldc) {
  __shared__ float As[GK][TM + 4], Bs[GK][TN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const bool a_kfast = sak == 1, b_kfast = sbk == 1;  // which index is contiguous in memory: coalesce along it
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += GK) {
#pragma unroll 8
    for (int i = threadIdx.x; i < GK * TM; i += 256) {
      const int ka = a_kfast ? (i % GK) : (i / TM), ma = a_kfast ? (i / GK) : (i % TM);
      As[ka][ma] = (k0 + ka < K && m0 + ma < M) ? A[(long long)(m0 + ma) * sam + (long long)(k0 + ka) * sak] : 0.f;
      const int kb = b_kfast ? (i % GK) : (i / TN), nb = b_kfast ? (i / GK) : (i % TN);
      Bs[kb][nb] = (k0 + kb < K && n0 + nb < N) ? B[(long long)(k0 + kb) * sbk + (long long)(n0 + nb) * sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll 16
    for (int kk = 0; kk < GK; ++kk) {
      const float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
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
      if (n >= N) continue;
      float v = acc[i][j] + (bias != nullptr ? bias[n] : 0.f);
      if (relu) v = fmaxf(v, 0.f);
      C[(long long)m * ldc + n] = v;
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
  dim3 grid((unsigned)ceil_div(N, TN), (unsigned)ceil_div(M, TM));
  S2C_REQUIRE(grid.y <= 65535u, "gemm: M=%d exceeds 65535 row tiles", M);
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
