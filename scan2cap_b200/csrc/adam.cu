// Adam over ONE flat fp32 parameter / gradient / moment buffer (torch.optim.Adam semantics, lib/solver.py:293-300 +
// scripts/train.py:134 in the reference: optimizer.step() after loss.backward()).  The framework's multi-tensor Adam
// walks the 144 parameter tensors of CapNet in ~27 launches (0.43 ms per step at 4.4 M parameters); with parameters,
// gradients and moments each living in one contiguous buffer the update is one streaming pass (7 x 4 B per parameter).
#include "s2c_common.cuh"

namespace s2c {
namespace {

// hyper = [lr, beta1, beta2, eps, weight_decay] on the device (a captured graph must not bake the learning rate in);
// steps = the per-parameter-tensor step counters torch.optim.Adam keeps (all equal), incremented here;
// coef  = [lr / (1 - beta1^t), 1 / sqrt(1 - beta2^t)] for the update kernel
__global__ void adam_prepare_kernel(const float *__restrict__ hyper, float *__restrict__ steps, int nsteps, float *__restrict__ coef) {
  const double t = (double)steps[0] + 1.0;
  __syncthreads();   // every thread has read the old counter before anybody writes
  for (int i = threadIdx.x; i < nsteps; i += blockDim.x) steps[i] = (float)t;
  if (threadIdx.x == 0) {
    const double b1 = hyper[1], b2 = hyper[2];
    coef[0] = (float)((double)hyper[0] / (1.0 - pow(b1, t)));
    coef[1] = (float)(1.0 / sqrt(1.0 - pow(b2, t)));
  }
}

__global__ void adam_update_kernel(float4 *__restrict__ p, const float4 *__restrict__ g, float4 *__restrict__ m, float4 *__restrict__ v,
                                   long long n4, const float *__restrict__ hyper, const float *__restrict__ coef, float grad_scale) {
  const float b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4];
  const float step_size = coef[0], inv_bc2_sqrt = coef[1];
  const float omb1 = 1.f - b1, omb2 = 1.f - b2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    float *P = &pp.x, *G = &gg.x, *M = &mm.x, *V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float grad = G[k] * grad_scale;
      grad = fmaf(wd, P[k], grad);                       // grad + weight_decay * param
      M[k] = fmaf(grad - M[k], omb1, M[k]);              // exp_avg.lerp_(grad, 1 - beta1)
      V[k] = fmaf(omb2 * grad, grad, V[k] * b2);         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      const float denom = fmaf(sqrtf(V[k]), inv_bc2_sqrt, eps);
      P[k] = P[k] - step_size * (M[k] / denom);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

}  // namespace
}  // namespace s2c

extern "C" int s2c_adam_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, long long n,
                             const float *hyper, float *steps, int nsteps, float *coef, float grad_scale, void *stream) {
  using namespace s2c;
  S2C_REQUIRE(n >= 0 && (n & 3) == 0, "adam_step: n=%lld must be a multiple of 4 (pad the flat buffers)", n);
  S2C_REQUIRE(nsteps >= 1, "adam_step: need at least one step counter");
  if (n == 0) return S2C_OK;
  S2C_REQUIRE(params && grads && exp_avg && exp_avg_sq && hyper && steps && coef, "adam_step: null pointer");
  S2C_REQUIRE((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
              "adam_step: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  adam_prepare_kernel<<<1, 256, 0, st>>>(hyper, steps, nsteps, coef);
  S2C_CHECK_LAUNCH("adam_prepare");
  const long long n4 = n >> 2;
  const long long want = (n4 + 255) / 256;
  const unsigned grid = (unsigned)(want < 1 ? 1 : (want > 8LL * kNumSMs ? 8LL * kNumSMs : want));
  adam_update_kernel<<<grid, 256, 0, st>>>((float4 *)params, (const float4 *)grads, (float4 *)exp_avg, (float4 *)exp_avg_sq, n4,
                                           hyper, coef, grad_scale);
  S2C_CHECK_LAUNCH("adam_update");
  return S2C_OK;
}
