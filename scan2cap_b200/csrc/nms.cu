// Post-processing of benchmark/predict.py on the device (lib/ap_helper.py:40-178 parse_predictions, utils/nms.py):
//   * points_in_boxes_count: how many scene points lie inside every (axis-aligned: ScanNet boxes have heading 0,
//     data/scannet/model_util_scannet.py:130-134) predicted box -- the reference builds a scipy Delaunay hull per box
//     (model_util_scannet.py:13-22 in_hull) and loops over B*K boxes on the host;
//   * nms3d_samecls: the greedy class-aware 3-D NMS of utils/nms.py:110-150 (nms_3d_faster_samecls) in float64, one
//     CTA per scene, K sequential rounds inside the kernel instead of a numpy loop per scene.
#include "s2c_common.cuh"

namespace s2c {
namespace {

// one thread per (scene, box); the point loop is a broadcast read (all threads of a CTA walk the same scene's points)
__global__ void points_in_boxes_count_kernel(const float *__restrict__ xyz, long long xyz_ld, int N,
                                             const double *__restrict__ boxes, int K, int *__restrict__ count) {
  const int b = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const double *bx = boxes + ((size_t)b * K + k) * 6;
  const double lo0 = bx[0], lo1 = bx[1], lo2 = bx[2], hi0 = bx[3], hi1 = bx[4], hi2 = bx[5];
  const float *p = xyz + (size_t)b * N * xyz_ld;
  int c = 0;
  for (int i = 0; i < N; ++i) {
    const double x = (double)p[(size_t)i * xyz_ld], y = (double)p[(size_t)i * xyz_ld + 1], z = (double)p[(size_t)i * xyz_ld + 2];
    c += (x >= lo0 && x <= hi0 && y >= lo1 && y <= hi1 && z >= lo2 && z <= hi2) ? 1 : 0;
  }
  count[(size_t)b * K + k] = c;
}

constexpr int kMaxBoxes = 1024;

__global__ void __launch_bounds__(kMaxBoxes) nms3d_samecls_kernel(const double *__restrict__ boxes, const double *__restrict__ score,
                                                                  const long long *__restrict__ cls, const int *__restrict__ valid,
                                                                  int K, double thr, int old_type, int same_class_only,
                                                                  int *__restrict__ keep) {
  __shared__ int order[kMaxBoxes];       // box ids, best score first
  __shared__ unsigned char dead[kMaxBoxes];
  __shared__ int s_pick;
  const int b = blockIdx.x, j = threadIdx.x;
  boxes += (size_t)b * K * 6; score += (size_t)b * K; cls += (size_t)b * K; valid += (size_t)b * K; keep += (size_t)b * K;
  double x1 = 0, y1 = 0, z1 = 0, x2 = 0, y2 = 0, z2 = 0, sc = 0, area = 0;
  long long cj = 0;
  bool ok = false;
  if (j < K) {
    x1 = boxes[j * 6 + 0]; y1 = boxes[j * 6 + 1]; z1 = boxes[j * 6 + 2];
    x2 = boxes[j * 6 + 3]; y2 = boxes[j * 6 + 4]; z2 = boxes[j * 6 + 5];
    sc = score[j]; cj = cls[j]; ok = valid[j] != 0;
    area = (x2 - x1) * (y2 - y1) * (z2 - z1);
    keep[j] = 0;
    dead[j] = ok ? 0 : 1;
    // rank among the valid boxes: higher score first; equal scores: the LATER box first (np.argsort is taken from the
    // end, utils/nms.py:121-126; for equal keys a stable sort keeps the later index last, i.e. picked first)
    int rank = 0;
    for (int q = 0; q < K; ++q) {
      if (!valid[q] || q == j) continue;
      const double sq = score[q];
      rank += (sq > sc || (sq == sc && q > j)) ? 1 : 0;
    }
    if (ok) order[rank] = j;
  }
  __syncthreads();
  int nvalid = 0;
  for (int q = 0; q < K; ++q) nvalid += valid[q] ? 1 : 0;   // uniform
  for (int r = 0; r < nvalid; ++r) {
    if (j == 0) s_pick = dead[order[r]] ? -1 : order[r];
    __syncthreads();
    const int i = s_pick;
    if (i >= 0) {
      if (j == i) { keep[i] = 1; dead[i] = 1; }
      else if (j < K && !dead[j]) {
        const double *bi = boxes + i * 6;
        const double xx1 = fmax(bi[0], x1), yy1 = fmax(bi[1], y1), zz1 = fmax(bi[2], z1);
        const double xx2 = fmin(bi[3], x2), yy2 = fmin(bi[4], y2), zz2 = fmin(bi[5], z2);
        const double l = fmax(0.0, xx2 - xx1), w = fmax(0.0, yy2 - yy1), h = fmax(0.0, zz2 - zz1);
        const double inter = l * w * h;
        const double ai = (bi[3] - bi[0]) * (bi[4] - bi[1]) * (bi[5] - bi[2]);
        double o = old_type ? inter / area : inter / (ai + area - inter + 1e-8);
        if (same_class_only && cls[i] != cj) o = 0.0;
        if (o > thr) dead[j] = 1;
      }
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace s2c

extern "C" int s2c_points_in_boxes_count(const float *xyz, long long xyz_ld, int B, int N, const double *boxes, int K,
                                         int *count, void *stream) {
  using namespace s2c;
  S2C_REQUIRE(B >= 0 && N >= 0 && K >= 0 && xyz_ld >= 3, "points_in_boxes_count: bad sizes");
  if (B == 0 || K == 0) return S2C_OK;
  S2C_REQUIRE(xyz && boxes && count, "points_in_boxes_count: null pointer");
  S2C_REQUIRE(B <= 65535, "points_in_boxes_count: B too large");
  points_in_boxes_count_kernel<<<dim3((unsigned)ceil_div(K, 64), (unsigned)B), 64, 0, (cudaStream_t)stream>>>(xyz, xyz_ld, N, boxes, K, count);
  S2C_CHECK_LAUNCH("points_in_boxes_count");
  return S2C_OK;
}

extern "C" int s2c_nms3d(const double *boxes, const double *score, const long long *cls, const int *valid, int B, int K,
                         double iou_threshold, int old_type, int same_class_only, int *keep, void *stream) {
  using namespace s2c;
  S2C_REQUIRE(B >= 0 && K >= 1 && K <= kMaxBoxes, "nms3d: K=%d outside [1,%d]", K, kMaxBoxes);
  if (B == 0) return S2C_OK;
  S2C_REQUIRE(boxes && score && cls && valid && keep, "nms3d: null pointer");
  const int threads = ((K + 31) / 32) * 32;
  nms3d_samecls_kernel<<<B, threads, 0, (cudaStream_t)stream>>>(boxes, score, cls, valid, K, iou_threshold, old_type,
                                                               same_class_only, keep);
  S2C_CHECK_LAUNCH("nms3d");
  return S2C_OK;
}
