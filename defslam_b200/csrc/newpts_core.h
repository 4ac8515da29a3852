/*
 * newpts_core.h -- occupancy test and placement of new map points.
 *
 * DefLocalMapping::CreateNewMapPoints (Modules/Mapping/DefLocalMapping.cc:240-347) and
 * ::needNewTemplate (:355-403) decide with an image-sized mask + cv::filter2D box filter whether a
 * keypoint without a map point lies next to one that has.  The filtered mask is non-zero at pixel
 * (x, y) iff some marked pixel (mx, my) satisfies, per axis, reflect101(c + d) == m for an offset
 * d in [-a, k-1-a] (k = cols/20, a = k/2: filter2D's default anchor).  That predicate is evaluated
 * here directly against the marked keypoints: no image, no filter.
 */
#ifndef DS_NEWPTS_CORE_H_
#define DS_NEWPTS_CORE_H_

#include "ds_common.h"

namespace ds {

/* does the 1-D window [c-a, c+k-1-a], folded back into [0, n) by BORDER_REFLECT_101
 * (cv::borderInterpolate: p < 0 -> -p, p >= n -> 2(n-1) - p), contain m?  Requires k <= n, so
 * one reflection is enough. */
DS_FN bool window_hits(int c, int m, int n, int k, int a) {
  const int lo = c - a, hi = c + k - 1 - a;
  if (m >= lo && m <= hi) return true;
  if (m >= 1 && -m >= lo) return true;                    /* images of lo..-1 are 1..-lo        */
  if (m <= n - 2 && 2 * (n - 1) - m <= hi) return true;   /* images of n..hi are 2(n-1)-hi..n-2 */
  return false;
}

/* fp32 multiply / add with one rounding each (never contracted into an FMA) */
DS_FN float mul_rn(float a, float b) {
#if DS_CUDA
  return __fmul_rn(a, b);
#else
  volatile float t = a * b;
  return t;
#endif
}
DS_FN float add_rn(float a, float b) {
#if DS_CUDA
  return __fadd_rn(a, b);
#else
  volatile float t = a + b;
  return t;
#endif
}

/* x3w = (Twc [x3c; 1])(0..2) as OpenCV computes a CV_32F 4x4 * 4x1 product (cv::gemm's small-
 * matrix path): fp32 products summed left to right in fp32, no FMA -- checked bit for bit against
 * cv2.gemm (tests/golden/newpts_cv.npz) */
DS_FN void surface_point_to_world(const float *Twc, const float *x3c, float *out) {
  for (int r = 0; r < 3; r++) {
    float s = mul_rn(Twc[4 * r], x3c[0]);
    s = add_rn(s, mul_rn(Twc[4 * r + 1], x3c[1]));
    s = add_rn(s, mul_rn(Twc[4 * r + 2], x3c[2]));
    s = add_rn(s, mul_rn(Twc[4 * r + 3], 1.0f));
    out[r] = s;
  }
}

}  // namespace ds
#endif
