// orbx_frame.cu — the per-frame glue between extractor and matcher on the device:
//   K18 frustum_kernel   : Frame::isInFrustum(MapPoint*, viewingCosLimit) for a whole local map at once
//                          (src/Frame.cc:571-650, Nleft == -1; MapPoint::PredictScale src/MapPoint.cc:578-610)
//   K19 undistort_kernel : Frame::UndistortKeyPoints = cv::undistortPoints(mat, mat, K, mDistCoef, Mat(), mK)
//                          (src/Frame.cc:874-924): 5 fixed-point iterations of the Brown-Conrady inverse in double
// Both are one thread per element, fully independent; fp32 expressions are written without FMA contraction and the
// double-precision parts follow OpenCV's expression order (compiled with --fmad=false).
#include "orbx_match.cuh"

struct FrustumArgs {
  float R[9], t[3], Ow[3];
  float fx, fy, cx, cy, bf;
  float minX, maxX, minY, maxY, cosLimit, logScale;
  int nlevels, n;
  const float *xw, *maxDist, *minDist, *normal;
  uint8_t* inView;
  float *projX, *projY, *projXR, *depth, *viewCos;
  int* level;
  int* count;
};

__global__ void __launch_bounds__(128) frustum_kernel(FrustumArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  A.inView[i] = 0;
  A.projX[i] = -1.f;
  A.projY[i] = -1.f;
  const float X = A.xw[3 * i], Y = A.xw[3 * i + 1], Z = A.xw[3 * i + 2];
  const float xc = A.R[0] * X + A.R[1] * Y + A.R[2] * Z + A.t[0];
  const float yc = A.R[3] * X + A.R[4] * Y + A.R[5] * Z + A.t[1];
  const float zc = A.R[6] * X + A.R[7] * Y + A.R[8] * Z + A.t[2];
  const float pcDist = (float)sqrt((double)xc * (double)xc + (double)yc * (double)yc + (double)zc * (double)zc);
  const float invz = 1.0f / zc;
  if (zc < 0.0f) return;
  const float u = A.fx * xc / zc + A.cx, v = A.fy * yc / zc + A.cy;
  if (u < A.minX || u > A.maxX) return;
  if (v < A.minY || v > A.maxY) return;
  A.projX[i] = u;
  A.projY[i] = v;
  const float maxDistance = 1.2f * A.maxDist[i], minDistance = 0.8f * A.minDist[i];
  const float P0 = X - A.Ow[0], P1 = Y - A.Ow[1], P2 = Z - A.Ow[2];
  const float dist = (float)sqrt((double)P0 * (double)P0 + (double)P1 * (double)P1 + (double)P2 * (double)P2);
  if (dist < minDistance || dist > maxDistance) return;
  const double dot = (double)P0 * (double)A.normal[3 * i] + (double)P1 * (double)A.normal[3 * i + 1] + (double)P2 * (double)A.normal[3 * i + 2];
  const float viewCos = (float)(dot / (double)dist);
  if (viewCos < A.cosLimit) return;
  const float ratio = A.maxDist[i] / dist;
  int lvl = (int)ceil(log((double)ratio) / (double)A.logScale);
  if (lvl < 0) lvl = 0; else if (lvl >= A.nlevels) lvl = A.nlevels - 1;
  A.inView[i] = 1;
  A.projXR[i] = u - A.bf * invz;
  A.depth[i] = pcDist;
  A.level[i] = lvl;
  A.viewCos[i] = viewCos;
  atomicAdd(A.count, 1);
}

__global__ void __launch_bounds__(128) undistort_kernel(const float2* __restrict__ xy, int n, double fx, double fy, double cx, double cy,
                                                        double k1, double k2, double p1, double p2, double k3, float2* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 p = xy[i];
  const double u = p.x, v = p.y;
  const double ifx = 1. / fx, ify = 1. / fy;
  double x = (u - cx) * ifx, y = (v - cy) * ify;
  const double x0 = x, y0 = y;
#pragma unroll 1
  for (int j = 0; j < 5; ++j) {
    const double r2 = x * x + y * y;
    const double icdist = 1.0 / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);   // numerator (1 + ((k7 r2 + k6) r2 + k5) r2) == 1 exactly
    if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
    const double deltaX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
    const double deltaY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
    x = (x0 - deltaX) * icdist;
    y = (y0 - deltaY) * icdist;
  }
  out[i] = make_float2((float)(fx * x + cx), (float)(fy * y + cy));
}

extern "C" {

int orbx_is_in_frustum(orbx_ctx* ctx, const orbx_camera* cam, const float* Rcw, const float* tcw, const float* Ow, float min_x,
                       float max_x, float min_y, float max_y, float viewing_cos_limit, int nlevels, float log_scale_factor,
                       int nmp, const float* xw, const float* mp_max_dist, const float* mp_min_dist, const float* mp_normal,
                       uint8_t* in_view, float* proj_x, float* proj_y, float* proj_xr, float* depth, int32_t* level,
                       float* view_cos, int32_t* n_in_view) {
  if (!ctx || !cam || !Rcw || !tcw || !Ow || nmp < 0 || nlevels < 1 || !n_in_view) return ORBX_EINVAL;
  *n_in_view = 0;
  if (nmp == 0) return ORBX_OK;
  if (!xw || !mp_max_dist || !mp_min_dist || !mp_normal || !in_view || !proj_x || !proj_y || !proj_xr || !depth || !level || !view_cos)
    return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DevScope S(ctx, st);
  FrustumArgs A;
  for (int i = 0; i < 9; ++i) A.R[i] = Rcw[i];
  for (int i = 0; i < 3; ++i) { A.t[i] = tcw[i]; A.Ow[i] = Ow[i]; }
  A.fx = cam->fx; A.fy = cam->fy; A.cx = cam->cx; A.cy = cam->cy; A.bf = cam->bf;
  A.minX = min_x; A.maxX = max_x; A.minY = min_y; A.maxY = max_y;
  A.cosLimit = viewing_cos_limit;
  A.logScale = log_scale_factor;
  A.nlevels = nlevels;
  A.n = nmp;
  A.xw = S.upload(xw, (size_t)3 * nmp);
  A.maxDist = S.upload(mp_max_dist, nmp);
  A.minDist = S.upload(mp_min_dist, nmp);
  A.normal = S.upload(mp_normal, (size_t)3 * nmp);
  A.inView = S.alloc<uint8_t>(nmp);
  A.projX = S.alloc<float>(nmp); A.projY = S.alloc<float>(nmp);
  // the four "written only when in view" outputs start from the caller's values, like the stale MapPoint fields
  A.projXR = S.upload(proj_xr, nmp); A.depth = S.upload(depth, nmp); A.viewCos = S.upload(view_cos, nmp);
  A.level = S.upload(level, nmp);
  A.count = S.alloc<int>(1);
  if (S.failed) return ORBX_ECUDA;
  ORBX_CUDA(cudaMemsetAsync(A.count, 0, sizeof(int), st));
  frustum_kernel<<<div_up(nmp, 128), 128, 0, st>>>(A);
  ORBX_LAUNCH(ctx);
  ORBX_CUDA(cudaGetLastError());
  S.download(in_view, A.inView, (size_t)nmp);
  S.download(proj_x, A.projX, (size_t)nmp);
  S.download(proj_y, A.projY, (size_t)nmp);
  S.download(proj_xr, A.projXR, (size_t)nmp);
  S.download(depth, A.depth, (size_t)nmp);
  S.download(view_cos, A.viewCos, (size_t)nmp);
  S.download(level, A.level, (size_t)nmp);
  S.download(n_in_view, A.count, (size_t)1);
  return S.finish();
}

int orbx_undistort_keypoints(orbx_ctx* ctx, const float* xy, int n, const orbx_camera* cam, const float* dist_coef, int n_dist,
                             float* out_xy) {
  if (!ctx || n < 0 || !cam || !dist_coef || n_dist < 4 || (n > 0 && (!xy || !out_xy))) return ORBX_EINVAL;
  if (n == 0) return ORBX_OK;
  if (dist_coef[0] == 0.0f) {                       // mvKeysUn = mvKeys (src/Frame.cc:877-881)
    if (out_xy != xy) memcpy(out_xy, xy, sizeof(float) * 2 * (size_t)n);
    return ORBX_OK;
  }
  ORBX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DevScope S(ctx, st);
  const float2* d_in = reinterpret_cast<const float2*>(S.upload(xy, (size_t)2 * n));
  float2* d_out = reinterpret_cast<float2*>(S.alloc<float>((size_t)2 * n));
  if (S.failed) return ORBX_ECUDA;
  undistort_kernel<<<div_up(n, 128), 128, 0, st>>>(d_in, n, (double)cam->fx, (double)cam->fy, (double)cam->cx, (double)cam->cy,
                                                   (double)dist_coef[0], (double)dist_coef[1], (double)dist_coef[2],
                                                   (double)dist_coef[3], n_dist > 4 ? (double)dist_coef[4] : 0.0, d_out);
  ORBX_LAUNCH(ctx);
  ORBX_CUDA(cudaGetLastError());
  S.download(out_xy, reinterpret_cast<const float*>(d_out), (size_t)2 * n);
  return S.finish();
}

}  // extern "C"
