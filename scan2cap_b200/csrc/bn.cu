// BatchNorm coefficient kernels of the fused shared-MLP path (one launch per layer instead of ~20 element-wise
// framework kernels on 64..256-element vectors).
//
// The reference runs nn.BatchNorm2d after every 1x1 conv of SharedMLP (lib/pointnet2/pytorch_utils.py:88-120):
// training mode normalises with the batch statistics over B*M*ns positions and updates running_mean / running_var
// (momentum 0.1, unbiased variance).  In the fused path the GEMM epilogue already produced the float64 column sums
// (sum y, sum y^2); what is left per layer is O(channels) arithmetic:
//   forward : mean, invstd -> folded (scale, shift) for the next kernel's prologue, running-statistics update
//   backward: dY = a*g + b*y + c (batch-statistics backward is affine per channel), grad_gamma, grad_beta
#include "s2c_common.cuh"

namespace s2c {
namespace {

__global__ void bn_finalize_kernel(const double *__restrict__ s1, const double *__restrict__ s2, double R, int N,
                                   const float *__restrict__ gamma, const float *__restrict__ beta, double eps,
                                   float momentum, float one_minus_momentum, int use_batch, int update_running,
                                   float *__restrict__ running_mean, float *__restrict__ running_var,
                                   long long *__restrict__ num_batches_tracked, double *__restrict__ mean_out,
                                   double *__restrict__ invstd_out, float *__restrict__ scale_out,
                                   float *__restrict__ shift_out) {
  // ONE block (channels are looped over): every thread reads the old counter, the barrier orders the reads before the
  // increment.  momentum < 0 selects nn.BatchNorm's momentum=None mode: cumulative moving average with factor
  // 1 / num_batches_tracked (counter value AFTER this batch's increment), torch/nn/modules/batchnorm.py.
  const long long seen = (update_running && num_batches_tracked) ? *num_batches_tracked : 0;
  __syncthreads();
  if (threadIdx.x == 0 && update_running && num_batches_tracked) *num_batches_tracked = seen + 1;
  if (momentum < 0.f) {
    momentum = 1.0f / (float)(seen + 1);
    one_minus_momentum = 1.0f - momentum;
  }
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    double mean, var;
    if (use_batch) {
      mean = s1[c] / R;
      var = fmax(s2[c] / R - mean * mean, 0.0);
      if (update_running) {
        const double unbiased = var * (R / fmax(R - 1.0, 1.0));
        running_mean[c] = __fmaf_rn(momentum, (float)mean, __fmul_rn(running_mean[c], one_minus_momentum));
        running_var[c] = __fmaf_rn(momentum, (float)unbiased, __fmul_rn(running_var[c], one_minus_momentum));
      }
    } else {
      mean = (double)running_mean[c];
      var = (double)running_var[c];
    }
    const double invstd = 1.0 / sqrt(var + eps);
    const double scale = (double)gamma[c] * invstd;
    mean_out[c] = mean;
    invstd_out[c] = invstd;
    scale_out[c] = (float)scale;
    shift_out[c] = (float)((double)beta[c] - mean * scale);
  }
}

__global__ void bn_backward_coeffs_kernel(const double *__restrict__ sum_g, const double *__restrict__ sum_gy,
                                          const double *__restrict__ mean, const double *__restrict__ invstd,
                                          const float *__restrict__ gamma, double R, int N, int batch_stats,
                                          float *__restrict__ grad_gamma, float *__restrict__ grad_beta,
                                          float *__restrict__ a_out, float *__restrict__ b_out,
                                          float *__restrict__ c_out) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= N) return;
  const double sg = sum_g[ch], m = mean[ch], is = invstd[ch];
  const double sgx = (sum_gy[ch] - m * sg) * is;  // sum of g * xhat
  grad_gamma[ch] = (float)sgx;
  grad_beta[ch] = (float)sg;
  const double a = (double)gamma[ch] * is;
  double b = 0.0, c = 0.0;
  if (batch_stats) {
    b = -a * is * (sgx / R);
    c = -a * (sg / R) - b * m;
  }
  a_out[ch] = (float)a;
  b_out[ch] = (float)b;
  c_out[ch] = (float)c;
}

// ---- channels-last scatter-add: gradient of the grouped rows w.r.t. the point-major features / coordinates ------
// rows (B, T, ld) with the wanted channels at [c0, c0+C); idx (B, T); out (B, n, C) += scale * rows.
// One warp per row; lanes run over channels, so the atomics of a row hit one contiguous run of the point's feature
// vector (the (B,C,T) formulation of group_points_grad issues T*C isolated 4-byte atomics C*4 bytes apart).
template <int VEC>
__global__ void __launch_bounds__(256)
rows_scatter_add_kernel(const float *__restrict__ rows, long long ld, int c0, int C, const int *__restrict__ idx,
                        long long T, int n, float scale, float *__restrict__ out) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  rows += (size_t)b * T * ld + c0;
  idx += (size_t)b * T;
  out += (size_t)b * n * C;
  for (long long t = warp0; t < T; t += nwarps) {
    const int k = __ldg(idx + t);
    const float *g = rows + (size_t)t * ld;
    float *o = out + (size_t)k * C;
    if (VEC == 4) {
      for (int c = lane * 4; c < C; c += 128) {
        // (the source run starts at channel c0 of a row: 4-byte aligned only; the destination is 16-byte aligned)
        const float v0 = ld_stream(g + c), v1 = ld_stream(g + c + 1), v2 = ld_stream(g + c + 2), v3 = ld_stream(g + c + 3);
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o + c), "f"(v0 * scale), "f"(v1 * scale),
                     "f"(v2 * scale), "f"(v3 * scale)
                     : "memory");
      }
    } else {
      for (int c = lane; c < C; c += 32) atomicAdd(o + c, ld_stream(g + c) * scale);
    }
  }
}

}  // namespace
}  // namespace s2c

using namespace s2c;

extern "C" int s2c_bn_finalize(const double *sum, const double *sumsq, long long R, int N, const float *gamma,
                               const float *beta, double eps, double momentum, int use_batch_stats,
                               int update_running, float *running_mean, float *running_var,
                               long long *num_batches_tracked, double *mean, double *invstd, float *scale,
                               float *shift, void *stream) {
  S2C_REQUIRE(N >= 0 && R >= 1, "bn_finalize: bad sizes R=%lld N=%d", R, N);
  if (N == 0) return S2C_OK;
  S2C_REQUIRE(gamma && beta && mean && invstd && scale && shift, "bn_finalize: null pointer");
  S2C_REQUIRE(!use_batch_stats || (sum && sumsq), "bn_finalize: batch statistics requested without sums");
  S2C_REQUIRE((use_batch_stats && !update_running) || (running_mean && running_var),
              "bn_finalize: running statistics needed but null");
  S2C_REQUIRE(momentum >= 0.0 || !update_running || num_batches_tracked,
              "bn_finalize: cumulative average (momentum < 0) needs num_batches_tracked");
  bn_finalize_kernel<<<1, N <= 128 ? 128 : 256, 0, (cudaStream_t)stream>>>(
      sum, sumsq, (double)R, N, gamma, beta, eps, (float)momentum, (float)(1.0 - momentum), use_batch_stats ? 1 : 0,
      update_running ? 1 : 0, running_mean, running_var, num_batches_tracked, mean, invstd, scale, shift);
  S2C_CHECK_LAUNCH("bn_finalize");
  return S2C_OK;
}

extern "C" int s2c_bn_backward_coeffs(const double *sum_g, const double *sum_gy, const double *mean,
                                      const double *invstd, const float *gamma, long long R, int N, int batch_stats,
                                      float *grad_gamma, float *grad_beta, float *a, float *b, float *c,
                                      void *stream) {
  S2C_REQUIRE(N >= 0 && R >= 1, "bn_backward_coeffs: bad sizes R=%lld N=%d", R, N);
  if (N == 0) return S2C_OK;
  S2C_REQUIRE(sum_g && sum_gy && mean && invstd && gamma && grad_gamma && grad_beta && a && b && c,
              "bn_backward_coeffs: null pointer");
  bn_backward_coeffs_kernel<<<ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(
      sum_g, sum_gy, mean, invstd, gamma, (double)R, N, batch_stats ? 1 : 0, grad_gamma, grad_beta, a, b, c);
  S2C_CHECK_LAUNCH("bn_backward_coeffs");
  return S2C_OK;
}

extern "C" int s2c_group_rows_grad(const float *rows, long long ld, int c0, int C, const int *idx, int B,
                                   long long T, int n, float scale, float *out, void *stream) {
  S2C_REQUIRE(B >= 0 && C >= 0 && T >= 0 && n >= 0 && c0 >= 0 && ld >= c0 + C, "group_rows_grad: bad sizes");
  if (B == 0 || C == 0 || n == 0) return S2C_OK;
  S2C_REQUIRE(out, "group_rows_grad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  S2C_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * n * C, st), "group_rows_grad memset");
  if (T == 0) return S2C_OK;
  S2C_REQUIRE(rows && idx, "group_rows_grad: null pointer");
  S2C_REQUIRE(B <= 65535, "group_rows_grad: B too large");
  const int warps = 8;
  const long long want = ceil_div_ll(T, warps);
  const unsigned gx = (unsigned)(want < 8LL * kNumSMs ? want : 8LL * kNumSMs);
  dim3 grid(gx, (unsigned)B);
  const bool vec = (C % 4 == 0) && (((uintptr_t)out & 15) == 0);
  if (vec)
    rows_scatter_add_kernel<4><<<grid, warps * 32, 0, st>>>(rows, ld, c0, C, idx, T, n, scale, out);
  else
    rows_scatter_add_kernel<1><<<grid, warps * 32, 0, st>>>(rows, ld, c0, C, idx, T, n, scale, out);
  S2C_CHECK_LAUNCH("group_rows_grad");
  return S2C_OK;
}
