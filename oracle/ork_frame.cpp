// ork_frame.cpp — ORACLE (test infrastructure): the per-frame glue between extractor and matcher.
//
//   Frame::isInFrustum(MapPoint*, viewingCosLimit), Nleft == -1 branch        src/Frame.cc:571-650
//   MapPoint::PredictScale / Get{Max,Min}DistanceInvariance                   src/MapPoint.cc:566-610
//   Frame::UndistortKeyPoints -> cv::undistortPoints(mat, mat, K, D, Mat(), K) src/Frame.cc:874-924
//
// cv::undistortPoints lives in OpenCV (not in the reference tree): restated from its published algorithm (5 fixed-point
// iterations of the Brown-Conrady inverse in double, `icdist < 0` guard) and pinned against Python cv2 4.13
// (tests/test_oracle_frame.py); OpenCV 3.x, which the reference names, runs the same 5 iterations without the guard.
// cv::Mat float algebra is evaluated in fp32 left to right, cv::norm / Mat::dot accumulate in double (normL2_32f,
// dotProd_32f), PredictScale's log in double — the same conventions as oracle/ork_matcher.cpp.
#include <cmath>
#include "ork.h"

extern "C" {

// Out, per MapPoint: in_view (mbTrackInView); proj_x/proj_y (mTrackProjX/Y: -1 when the point is behind the camera
// or outside the image, else the projection even if a later test fails, as in the reference); proj_xr, depth, level,
// view_cos are written only where in_view = 1 (the reference leaves the stale values in place otherwise,
// SURVEY.md App. B #23).  Returns the number of points in view.
int ork_is_in_frustum(const orbx_camera* cam, const float* Rcw, const float* tcw, const float* Ow, float min_x, float max_x,
                      float min_y, float max_y, float viewing_cos_limit, int nlevels, float log_scale_factor, int nmp,
                      const float* xw, const float* mp_max_dist, const float* mp_min_dist, const float* mp_normal,
                      uint8_t* in_view, float* proj_x, float* proj_y, float* proj_xr, float* depth, int32_t* level,
                      float* view_cos) {
  int n = 0;
  for (int i = 0; i < nmp; ++i) {
    in_view[i] = 0;
    proj_x[i] = -1;
    proj_y[i] = -1;
    const float X = xw[3 * i], Y = xw[3 * i + 1], Z = xw[3 * i + 2];
    const float xc = Rcw[0] * X + Rcw[1] * Y + Rcw[2] * Z + tcw[0];
    const float yc = Rcw[3] * X + Rcw[4] * Y + Rcw[5] * Z + tcw[1];
    const float zc = Rcw[6] * X + Rcw[7] * Y + Rcw[8] * Z + tcw[2];
    const float pcDist = (float)std::sqrt((double)xc * xc + (double)yc * yc + (double)zc * zc);
    const float invz = 1.0f / zc;
    if (zc < 0.0f) continue;
    const float u = cam->fx * xc / zc + cam->cx, v = cam->fy * yc / zc + cam->cy;
    if (u < min_x || u > max_x) continue;
    if (v < min_y || v > max_y) continue;
    proj_x[i] = u;
    proj_y[i] = v;
    const float maxDistance = 1.2f * mp_max_dist[i], minDistance = 0.8f * mp_min_dist[i];
    const float PO[3] = {X - Ow[0], Y - Ow[1], Z - Ow[2]};
    const float dist = (float)std::sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);
    if (dist < minDistance || dist > maxDistance) continue;
    const double dot = (double)PO[0] * mp_normal[3 * i] + (double)PO[1] * mp_normal[3 * i + 1] + (double)PO[2] * mp_normal[3 * i + 2];
    const float viewCos = (float)(dot / dist);
    if (viewCos < viewing_cos_limit) continue;
    const float ratio = mp_max_dist[i] / dist;
    int lvl = (int)std::ceil(std::log((double)ratio) / (double)log_scale_factor);
    if (lvl < 0) lvl = 0; else if (lvl >= nlevels) lvl = nlevels - 1;
    in_view[i] = 1;
    proj_xr[i] = u - cam->bf * invz;
    depth[i] = pcDist;
    level[i] = lvl;
    view_cos[i] = viewCos;
    ++n;
  }
  return n;
}

// cv::undistortPoints(src, dst, K, D, noArray(), P = K) on n points; dist = (k1, k2, p1, p2[, k3]); ndist = 4 or 5.
// The reference skips the call when k1 == 0 (src/Frame.cc:877-881): then the keypoints are copied.
int ork_cv_undistort_points(const float* xy, int n, const orbx_camera* cam, const float* dist, int ndist, float* out_xy);
int ork_undistort_points(const float* xy, int n, const orbx_camera* cam, const float* dist, int ndist, float* out_xy) {
  if (ndist < 4) return ORBX_EINVAL;
  if (dist[0] == 0.0f) {
    for (int i = 0; i < 2 * n; ++i) out_xy[i] = xy[i];
    return ORBX_OK;
  }
  return ork_cv_undistort_points(xy, n, cam, dist, ndist, out_xy);
}

// cv::undistortPoints itself (no shortcut): what oracle/ref_stub's stand-in calls when the reference's Frame.cc does
int ork_cv_undistort_points(const float* xy, int n, const orbx_camera* cam, const float* dist, int ndist, float* out_xy) {
  if (ndist < 4) return ORBX_EINVAL;
  double k[5] = {dist[0], dist[1], dist[2], dist[3], ndist > 4 ? (double)dist[4] : 0.0};
  const double fx = cam->fx, fy = cam->fy, cx = cam->cx, cy = cam->cy;
  const double ifx = 1. / fx, ify = 1. / fy;
  for (int i = 0; i < n; ++i) {
    const double u = xy[2 * i], v = xy[2 * i + 1];
    double x = (u - cx) * ifx, y = (v - cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; ++j) {
      const double r2 = x * x + y * y;
      const double icdist = (1 + ((0 * r2 + 0) * r2 + 0) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
      if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
      const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + 0 * r2 + 0 * r2 * r2;
      const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + 0 * r2 + 0 * r2 * r2;
      x = (x0 - deltaX) * icdist;
      y = (y0 - deltaY) * icdist;
    }
    const double xx = fx * x + 0 * y + cx, yy = 0 * x + fy * y + cy, ww = 1. / (0 * x + 0 * y + 1.0);
    out_xy[2 * i] = (float)(xx * ww);
    out_xy[2 * i + 1] = (float)(yy * ww);
  }
  return ORBX_OK;
}

}  // extern "C"
