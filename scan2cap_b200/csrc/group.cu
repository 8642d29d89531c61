// gather_points / group_points and their gradients for sm_100a.
//
// Replaces (reference lib/pointnet2/_ext_src/src/):
//   gather_points_kernel       sampling_gpu.cu:8-30      gather_points_grad_kernel  sampling_gpu.cu:34-57
//   group_points_kernel        group_points_gpu.cu:8-28  group_points_grad_kernel   group_points_gpu.cu:43-64
// The reference launches one block per scene (grid = B) and, in group_points, lets adjacent threads
// write `nsample` floats apart while re-reading idx once per channel.  Here the grid covers
// (output element tiles) x (channel chunks) x B, adjacent threads own adjacent output elements
// (coalesced 128-byte stores), and each index is read once per channel chunk.
#include "s2c_common.cuh"

namespace s2c {
namespace {

constexpr int kThreads = 256;
constexpr int kChanChunk = 8;

// points (B,C,N), idx (B,T) -> out (B,C,T)          [T = m for gather, np*ns for group]
__global__ void __launch_bounds__(kThreads)
index_select_kernel(const float *__restrict__ points, const int *__restrict__ idx, int C, int N, long long T,
                    float *__restrict__ out) {
  const int b = blockIdx.z;
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (t >= T) return;
  const int c0 = blockIdx.y * kChanChunk;
  const int k = idx[(size_t)b * T + t];
  const float *p = points + ((size_t)b * C + c0) * N + k;
  float *o = out + ((size_t)b * C + c0) * T + t;
  const int cn = min(kChanChunk, C - c0);
  float v[kChanChunk];
#pragma unroll
  for (int c = 0; c < kChanChunk; ++c)
    if (c < cn) v[c] = __ldg(p + (size_t)c * N);
#pragma unroll
  for (int c = 0; c < kChanChunk; ++c)
    if (c < cn) st_stream(o + (size_t)c * T, v[c]);
}

// grad_out (B,C,T), idx (B,T) -> grad_points (B,C,N) += (float atomics; pre-zeroed by the caller)
__global__ void __launch_bounds__(kThreads)
index_scatter_add_kernel(const float *__restrict__ grad_out, const int *__restrict__ idx, int C, int N, long long T,
                         float *__restrict__ grad_points) {
  const int b = blockIdx.z;
  const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (t >= T) return;
  const int c0 = blockIdx.y * kChanChunk;
  const int k = idx[(size_t)b * T + t];
  const float *g = grad_out + ((size_t)b * C + c0) * T + t;
  float *p = grad_points + ((size_t)b * C + c0) * N + k;
  const int cn = min(kChanChunk, C - c0);
  float v[kChanChunk];
#pragma unroll
  for (int c = 0; c < kChanChunk; ++c)
    if (c < cn) v[c] = ld_stream(g + (size_t)c * T);
#pragma unroll
  for (int c = 0; c < kChanChunk; ++c)
    if (c < cn) atomicAdd(p + (size_t)c * N, v[c]);
}

int select(const float *points, const int *idx, int B, int C, int N, long long T, float *out, cudaStream_t st,
           const char *what) {
  if (B == 0 || C == 0 || T == 0) return S2C_OK;
  S2C_REQUIRE(points && idx && out, "%s: null pointer", what);
  S2C_REQUIRE(B <= 65535 && ceil_div(C, kChanChunk) <= 65535, "%s: B or C too large", what);
  dim3 grid((unsigned)ceil_div_ll(T, kThreads), (unsigned)ceil_div(C, kChanChunk), (unsigned)B);
  index_select_kernel<<<grid, kThreads, 0, st>>>(points, idx, C, N, T, out);
  S2C_CHECK_LAUNCH(what);
  return S2C_OK;
}

int scatter(const float *grad_out, const int *idx, int B, int C, int N, long long T, float *grad_points,
            cudaStream_t st, const char *what) {
  if (B == 0 || C == 0 || N == 0) return S2C_OK;
  S2C_REQUIRE(grad_points, "%s: null pointer", what);
  S2C_CUDA(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)B * C * N, st), what);
  if (T == 0) return S2C_OK;
  S2C_REQUIRE(grad_out && idx, "%s: null pointer", what);
  S2C_REQUIRE(B <= 65535 && ceil_div(C, kChanChunk) <= 65535, "%s: B or C too large", what);
  dim3 grid((unsigned)ceil_div_ll(T, kThreads), (unsigned)ceil_div(C, kChanChunk), (unsigned)B);
  index_scatter_add_kernel<<<grid, kThreads, 0, st>>>(grad_out, idx, C, N, T, grad_points);
  S2C_CHECK_LAUNCH(what);
  return S2C_OK;
}

}  // namespace
}  // namespace s2c

using namespace s2c;

extern "C" int s2c_gather_points(const float *points, const int *idx, int B, int C, int N, int m, float *out,
                                 void *stream) {
  S2C_REQUIRE(B >= 0 && C >= 0 && N >= 0 && m >= 0, "gather_points: negative size");
  return select(points, idx, B, C, N, m, out, (cudaStream_t)stream, "gather_points");
}

extern "C" int s2c_gather_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int m,
                                      float *grad_points, void *stream) {
  S2C_REQUIRE(B >= 0 && C >= 0 && N >= 0 && m >= 0, "gather_points_grad: negative size");
  return scatter(grad_out, idx, B, C, N, m, grad_points, (cudaStream_t)stream, "gather_points_grad");
}

extern "C" int s2c_group_points(const float *points, const int *idx, int B, int C, int N, int npoints, int nsample,
                                float *out, void *stream) {
  S2C_REQUIRE(B >= 0 && C >= 0 && N >= 0 && npoints >= 0 && nsample >= 0, "group_points: negative size");
  return select(points, idx, B, C, N, (long long)npoints * nsample, out, (cudaStream_t)stream, "group_points");
}

extern "C" int s2c_group_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int npoints,
                                     int nsample, float *grad_points, void *stream) {
  S2C_REQUIRE(B >= 0 && C >= 0 && N >= 0 && npoints >= 0 && nsample >= 0, "group_points_grad: negative size");
  return scatter(grad_out, idx, B, C, N, (long long)npoints * nsample, grad_points, (cudaStream_t)stream,
                 "group_points_grad");
}
