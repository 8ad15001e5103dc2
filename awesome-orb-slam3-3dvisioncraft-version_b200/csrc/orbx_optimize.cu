// orbx_optimize.cu — sm_100a PoseOptimization and LocalBundleAdjustment behind include/orbx.h.
//
// Replaces (reference paths): Optimizer::PoseOptimization src/Optimizer.cc:907-1272, the numeric core of
// Optimizer::LocalBundleAdjustment src/Optimizer.cc:1958-2352, and the g2o code under them
// (Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-185, core/block_solver.hpp:354-610,
// core/base_{unary,binary}_edge.hpp, core/robust_kernel_impl.cpp:66-91, types/se3quat.h,
// types/types_six_dof_expmap.cpp, src/OptimizableTypes.cpp).
//
// B200 design: these are tiny, latency-bound fp64 problems (E ~ 150-500 edges for a pose, ~18 k edges /
// 20 keyframes / 3000 points for a local BA), so a whole optimisation — graph-free: flat edge arrays, the
// LM controller, the 6x6 solve or the Schur complement + dense LDL^T, the chi2 classification — runs
// inside ONE thread block with zero host round trips; independent streams batch across blocks/SMs.
// Reductions use fixed-shape trees, so results are run-to-run deterministic.  Compiled with --fmad=false so
// every expression rounds exactly as in the CPU path (only summation order differs).
#include <algorithm>
#include <vector>
#include <cooperative_groups.h>
#include "orbx_match.cuh"

namespace cg = cooperative_groups;

// ------------------------------------------------------------------------------------
// SE3 (unit quaternion + translation), fp64 — same expressions as g2o::SE3Quat / Eigen
// ------------------------------------------------------------------------------------
struct Quat { double x, y, z, w; };
struct SE3d { Quat r; double t[3]; };

__device__ __forceinline__ Quat quat_from_R(const double* R) {
  Quat q;
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (R[7] - R[5]) * t;
    q.y = (R[2] - R[6]) * t;
    q.z = (R[3] - R[1]) * t;
  } else {
    // Eigen's branch on the largest diagonal element, written out per case (no dynamically indexed arrays:
    // those would live in local memory)
    int i = 0;
    double rii = R[0];
    if (R[4] > rii) { i = 1; rii = R[4]; }
    if (R[8] > rii) i = 2;
    if (i == 0) {
      t = sqrt(R[0] - R[4] - R[8] + 1.0);
      q.x = 0.5 * t;
      t = 0.5 / t;
      q.w = (R[7] - R[5]) * t;
      q.y = (R[3] + R[1]) * t;
      q.z = (R[6] + R[2]) * t;
    } else if (i == 1) {
      t = sqrt(R[4] - R[8] - R[0] + 1.0);
      q.y = 0.5 * t;
      t = 0.5 / t;
      q.w = (R[2] - R[6]) * t;
      q.z = (R[7] + R[5]) * t;
      q.x = (R[1] + R[3]) * t;
    } else {
      t = sqrt(R[8] - R[0] - R[4] + 1.0);
      q.z = 0.5 * t;
      t = 0.5 / t;
      q.w = (R[3] - R[1]) * t;
      q.x = (R[2] + R[6]) * t;
      q.y = (R[5] + R[7]) * t;
    }
  }
  return q;
}
__device__ __forceinline__ void quat_normalize(Quat& q) {
  if (q.w < 0) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
  const double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}
__device__ __forceinline__ Quat quat_mul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
__device__ __forceinline__ void quat_rot(const Quat& q, const double* v, double* out) {
  double uv0 = q.y * v[2] - q.z * v[1], uv1 = q.z * v[0] - q.x * v[2], uv2 = q.x * v[1] - q.y * v[0];
  uv0 += uv0; uv1 += uv1; uv2 += uv2;
  out[0] = v[0] + q.w * uv0 + (q.y * uv2 - q.z * uv1);
  out[1] = v[1] + q.w * uv1 + (q.z * uv0 - q.x * uv2);
  out[2] = v[2] + q.w * uv2 + (q.x * uv1 - q.y * uv0);
}
__device__ __forceinline__ void quat_to_R(const Quat& q, double* R) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
__device__ __forceinline__ SE3d se3_from_Tcw(const float* T) {   // Converter::toSE3Quat
  double R[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i * 3 + j] = (double)T[i * 4 + j];
  SE3d s;
  s.r = quat_from_R(R);
  quat_normalize(s.r);
  for (int i = 0; i < 3; ++i) s.t[i] = (double)T[i * 4 + 3];
  return s;
}
__device__ __forceinline__ void se3_to_Tcw(const SE3d& s, float* T) {   // Converter::toCvMat(SE3Quat)
  double R[9];
  quat_to_R(s.r, R);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T[i * 4 + j] = (float)R[i * 3 + j];
    T[i * 4 + 3] = (float)s.t[i];
  }
  T[12] = T[13] = T[14] = 0.f;
  T[15] = 1.f;
}
__device__ __forceinline__ void se3_map(const SE3d& s, const double* x, double* out) {
  quat_rot(s.r, x, out);
  out[0] += s.t[0]; out[1] += s.t[1]; out[2] += s.t[2];
}
__device__ __forceinline__ SE3d se3_mul(const SE3d& a, const SE3d& b) {
  SE3d r = a;
  double rt[3];
  quat_rot(a.r, b.t, rt);
  r.t[0] += rt[0]; r.t[1] += rt[1]; r.t[2] += rt[2];
  r.r = quat_mul(a.r, b.r);
  quat_normalize(r.r);
  return r;
}
__device__ SE3d se3_exp(const double* u) {   // SE3Quat::exp, u = [omega, upsilon]
  const double* om = u;
  const double* up = u + 3;
  const double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
  const double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
  double O2[9], R[9], V[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) O2[i * 3 + j] = O[i * 3] * O[j] + O[i * 3 + 1] * O[3 + j] + O[i * 3 + 2] * O[6 + j];
  if (theta < 0.00001) {
#pragma unroll
    for (int i = 0; i < 9; ++i) { R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + O[i] + O2[i]; V[i] = R[i]; }
  } else {
    const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta);
    const double c = (theta - sin(theta)) / (theta * theta * theta);   // pow(theta, 3)
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const double id = (i % 4 == 0) ? 1.0 : 0.0;
      R[i] = id + a * O[i] + b * O2[i];
      V[i] = id + b * O[i] + c * O2[i];
    }
  }
  SE3d s;
  s.r = quat_from_R(R);
  quat_normalize(s.r);
#pragma unroll
  for (int i = 0; i < 3; ++i) s.t[i] = V[i * 3] * up[0] + V[i * 3 + 1] * up[1] + V[i * 3 + 2] * up[2];
  return s;
}

// RobustKernelHuber (delta is the reference's float sqrt(5.991|7.815) widened; dsqr is a float member)
struct HuberD { double delta; float dsqr; };
__device__ __forceinline__ HuberD huber_make(float d) {
  HuberD h;
  h.delta = (double)d;
  h.dsqr = (float)((double)d * (double)d);
  return h;
}
__device__ __forceinline__ double huber_rho(const HuberD& h, double e, double& w) {
  if (e <= (double)h.dsqr) { w = 1.0; return e; }
  const double sqrte = sqrt(e);
  w = h.delta / sqrte;
  return 2 * sqrte * h.delta - (double)h.dsqr;
}

// projection residual of one observation (mono 2-D or stereo 3-D); p = camera-frame point
__device__ __forceinline__ void reproj_error(const double* p, const float* ob, bool stereo, float fx, float fy, float cx,
                                             float cy, float bf, double* out) {
  if (!stereo) {
    out[0] = (double)ob[0] - ((double)fx * p[0] / p[2] + (double)cx);
    out[1] = (double)ob[1] - ((double)fy * p[1] / p[2] + (double)cy);
    out[2] = 0;
  } else {
    const float invz = (float)(1.0 / p[2]);   // float invz: types_six_dof_expmap.cpp:191,340
    const double u = p[0] * (double)invz * (double)fx + (double)cx;
    const double v = p[1] * (double)invz * (double)fy + (double)cy;
    out[0] = (double)ob[0] - u;
    out[1] = (double)ob[1] - v;
    out[2] = (double)ob[2] - (u - (double)bf * (double)invz);
  }
}

// d err / d pose (Dx6): EdgeSE3ProjectXYZOnlyPose / EdgeSE3ProjectXYZ (mono) and the stereo variants.
// `binary` selects the z-division form of the LocalBA edges (x*y/z_2 ...) over the only-pose form (x*y*invz_2 ...).
__device__ __forceinline__ void jac_pose(const double* p, bool stereo, bool binary, double fx, double fy, double bf, double* J) {
  const double x = p[0], y = p[1], z = p[2];
  if (!stereo) {
    // mono: -(projectJac * [ -[p]x | I ])
    const double j00 = fx / z, j02 = -fx * x / (z * z), j11 = fy / z, j12 = -fy * y / (z * z);
    const double S0[6] = {0, z, -y, 1, 0, 0}, S1[6] = {-z, 0, x, 0, 1, 0}, S2[6] = {y, -x, 0, 0, 0, 1};
    if (!binary) {
      for (int c = 0; c < 6; ++c) {
        J[c] = -(j00 * S0[c] + j02 * S2[c]);
        J[6 + c] = -(j11 * S1[c] + j12 * S2[c]);
      }
    } else {
      const double n00 = -(fx / z), n02 = fx * x / (z * z), n11 = -(fy / z), n12 = fy * y / (z * z);
      for (int c = 0; c < 6; ++c) {
        J[c] = n00 * S0[c] + n02 * S2[c];
        J[6 + c] = n11 * S1[c] + n12 * S2[c];
      }
    }
  } else if (!binary) {
    const double invz = 1.0 / z, invz2 = invz * invz;
    J[0] = x * y * invz2 * fx; J[1] = -(1 + (x * x * invz2)) * fx; J[2] = y * invz * fx;
    J[3] = -invz * fx; J[4] = 0; J[5] = x * invz2 * fx;
    J[6] = (1 + y * y * invz2) * fy; J[7] = -x * y * invz2 * fy; J[8] = -x * invz * fy;
    J[9] = 0; J[10] = -invz * fy; J[11] = y * invz2 * fy;
    J[12] = J[0] - bf * y * invz2; J[13] = J[1] + bf * x * invz2; J[14] = J[2];
    J[15] = J[3]; J[16] = 0; J[17] = J[5] - bf * invz2;
  } else {
    const double z2 = z * z;
    J[0] = x * y / z2 * fx; J[1] = -(1 + (x * x / z2)) * fx; J[2] = y / z * fx;
    J[3] = -1. / z * fx; J[4] = 0; J[5] = x / z2 * fx;
    J[6] = (1 + y * y / z2) * fy; J[7] = -x * y / z2 * fy; J[8] = -x / z * fy;
    J[9] = 0; J[10] = -1. / z * fy; J[11] = y / z2 * fy;
    J[12] = J[0] - bf * y / z2; J[13] = J[1] + bf * x / z2; J[14] = J[2];
    J[15] = J[3]; J[16] = 0; J[17] = J[5] - bf / z2;
  }
}

// Block-wide sum of NV doubles per thread; every thread returns the same totals (fixed-shape tree:
// lane butterfly, then warps in index order).  s_red must hold (blockDim/32)*NV doubles.
template <int NV>
__device__ __forceinline__ void block_sum(double* v, double* s_red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  __syncthreads();   // protect s_red from the previous use
  if (lane == 0)
    for (int k = 0; k < NV; ++k) s_red[wid * NV + k] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double s = 0;
    for (int w = 0; w < nw; ++w) s += s_red[w * NV + k];
    v[k] = s;
  }
}

// First two stages of block_sum only: lane butterfly, then one partial per warp into s_red[wid*NV + k].
template <int NV>
__device__ __forceinline__ void block_partials(double* v, double* s_red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  __syncthreads();   // protect s_red from the previous use
  if (lane == 0)
    for (int k = 0; k < NV; ++k) s_red[wid * NV + k] = v[k];
  __syncthreads();
}

// Eigen::LDLT-like 6x6 solve with diagonal pivoting (largest |diagonal|); false on a negative pivot.
__device__ bool ldlt6_solve(const double* Ain, const double* b, double* x) {
  double A[36];
  int perm[6];
  for (int i = 0; i < 36; ++i) A[i] = Ain[i];
  for (int i = 0; i < 6; ++i) perm[i] = i;
  bool positive = true;
  for (int k = 0; k < 6; ++k) {
    int p = k;
    double best = fabs(A[k * 6 + k]);
    for (int i = k + 1; i < 6; ++i)
      if (fabs(A[i * 6 + i]) > best) { best = fabs(A[i * 6 + i]); p = i; }
    if (p != k) {
      for (int j = 0; j < 6; ++j) { const double t = A[k * 6 + j]; A[k * 6 + j] = A[p * 6 + j]; A[p * 6 + j] = t; }
      for (int i = 0; i < 6; ++i) { const double t = A[i * 6 + k]; A[i * 6 + k] = A[i * 6 + p]; A[i * 6 + p] = t; }
      const int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
    }
    const double d = A[k * 6 + k];
    if (d < 0) positive = false;
    if (d == 0) continue;
    for (int i = k + 1; i < 6; ++i) A[i * 6 + k] /= d;
    for (int i = k + 1; i < 6; ++i)
      for (int j = k + 1; j <= i; ++j) {
        A[i * 6 + j] -= A[i * 6 + k] * d * A[j * 6 + k];
        A[j * 6 + i] = A[i * 6 + j];
      }
  }
  if (!positive) return false;
  double y[6];
  for (int i = 0; i < 6; ++i) y[i] = b[perm[i]];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < i; ++j) y[i] -= A[i * 6 + j] * y[j];
  for (int i = 0; i < 6; ++i) { const double d = A[i * 6 + i]; y[i] = (d != 0) ? y[i] / d : 0.0; }
  for (int i = 5; i >= 0; --i)
    for (int j = 5; j > i; --j) y[i] -= A[j * 6 + i] * y[j];
  for (int i = 0; i < 6; ++i) x[perm[i]] = y[i];
  return true;
}

// Warp-cooperative version of ldlt6_solve on shared memory (A: 36 doubles, overwritten; perm/y scratch).  Same
// arithmetic as the serial routine, element for element: step k picks the largest |diagonal| (first on ties),
// swaps row/column k<->p, scales column k by the pivot and applies A[i][j] -= A[i][k]*d*A[j][k] to the trailing
// lower triangle (mirrored); forward substitution subtracts columns in increasing j, backward in decreasing j.
// All 32 lanes of ONE warp must call; returns the same value on every lane.
__device__ bool ldlt6_solve_warp(double* A, const double* b, double* x, int* perm, double* y) {
  const int lane = threadIdx.x & 31;
  if (lane < 6) perm[lane] = lane;
  __syncwarp();
  bool positive = true;
  for (int k = 0; k < 6; ++k) {
    // pivot: first index >= k with the largest |A[i][i]|
    int p = k;
    double best = fabs(A[k * 6 + k]);
    for (int i = k + 1; i < 6; ++i) {
      const double v = fabs(A[i * 6 + i]);
      if (v > best) { best = v; p = i; }
    }
    if (p != k) {   // (uniform: every lane computed the same p)
      double r0 = 0, r1 = 0;
      if (lane < 6) { r0 = A[k * 6 + lane]; r1 = A[p * 6 + lane]; }
      __syncwarp();
      if (lane < 6) { A[k * 6 + lane] = r1; A[p * 6 + lane] = r0; }
      __syncwarp();
      if (lane < 6) { r0 = A[lane * 6 + k]; r1 = A[lane * 6 + p]; }
      __syncwarp();
      if (lane < 6) { A[lane * 6 + k] = r1; A[lane * 6 + p] = r0; }
      if (lane == 0) { const int t = perm[k]; perm[k] = perm[p]; perm[p] = t; }
      __syncwarp();
    }
    const double d = A[k * 6 + k];
    if (d < 0) positive = false;
    if (d == 0) continue;
    if (lane > k && lane < 6) A[lane * 6 + k] /= d;
    __syncwarp();
    // trailing lower triangle: lane -> (i, j) with k < j <= i < 6  (at most 15 pairs)
    {
      const int m = 5 - k;                       // size of the trailing block
      int i = -1, j = -1;
      if (lane < m * (m + 1) / 2) {
        int r = 0, rem = lane;
        while (rem > r) { rem -= r + 1; ++r; }   // row r of the packed triangle, column rem
        i = k + 1 + r;
        j = k + 1 + rem;
      }
      double v = 0;
      if (i >= 0) v = A[i * 6 + j] - A[i * 6 + k] * d * A[j * 6 + k];
      __syncwarp();
      if (i >= 0) { A[i * 6 + j] = v; A[j * 6 + i] = v; }
      __syncwarp();
    }
  }
  if (!positive) return false;
  if (lane < 6) y[lane] = b[perm[lane]];
  __syncwarp();
  for (int j = 0; j < 6; ++j) {                  // forward: y[i] -= L[i][j] * y[j], j ascending
    if (lane > j && lane < 6) y[lane] -= A[lane * 6 + j] * y[j];
    __syncwarp();
  }
  if (lane < 6) { const double d = A[lane * 6 + lane]; y[lane] = (d != 0) ? y[lane] / d : 0.0; }
  __syncwarp();
  for (int j = 5; j >= 0; --j) {                 // backward: y[i] -= L[j][i] * y[j], j descending
    if (lane < j) y[lane] -= A[j * 6 + lane] * y[j];
    __syncwarp();
  }
  if (lane < 6) x[perm[lane]] = y[lane];
  __syncwarp();
  return true;
}

// Register/shuffle version of ldlt6_solve for ONE warp, no shared-memory round trips: lane i (< 6) keeps row i of the
// symmetric matrix in registers.  Same arithmetic as the serial routine, element for element: the pivot is the first
// largest |diagonal|, rows are exchanged between lanes and columns inside each lane, the trailing entry (i, j) is
// A[i][j] - (A[max][k]*d)*A[min][k] exactly as the serial code computes the lower entry and mirrors it.  Lane k also
// keeps the scaled column k as its row tail (L^T), so the backward substitution reads registers only.
// A: 36 doubles (symmetric, row-major, lambda already on the diagonal), b: 6, x: 6 (shared or global memory).
// All 32 lanes of the warp must call; every lane returns the same value.
template <int N>
__device__ bool ldlt_solve_shfl(const double* A, const double* b, double* x) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int li = lane < N ? lane : N - 1;        // lanes >= N shadow lane N-1 (never a shuffle source)
  double a[N];
#pragma unroll
  for (int j = 0; j < N; ++j) a[j] = A[li * N + j];
  int perm = li;
  bool positive = true;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double dg = a[0];
#pragma unroll
    for (int j = 1; j < N; ++j)
      if (li == j) dg = a[j];
    int p = k;
    double best = fabs(__shfl_sync(FULL, dg, k));
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      const double v = fabs(__shfl_sync(FULL, dg, i));
      if (v > best) { best = v; p = i; }
    }
    if (p != k) {                                 // warp-uniform
      const int src = lane == k ? p : (lane == p ? k : lane);
#pragma unroll
      for (int j = 0; j < N; ++j) a[j] = __shfl_sync(FULL, a[j], src);
      perm = __shfl_sync(FULL, perm, src);
      const double t = a[k];
      double ap = t;
#pragma unroll
      for (int j = k + 1; j < N; ++j)
        if (p == j) { ap = a[j]; a[j] = t; }
      a[k] = ap;
    }
    const double d = __shfl_sync(FULL, a[k], k);
    if (d < 0) positive = false;
    if (d == 0) continue;
    double lik = a[k];
    if (li > k) { lik = a[k] / d; a[k] = lik; }
#pragma unroll
    for (int j = k + 1; j < N; ++j) {
      const double ljk = __shfl_sync(FULL, lik, j);
      if (li > k) {
        const double hi = j <= li ? lik : ljk, lo = j <= li ? ljk : lik;   // (A[max(i,j)][k]*d)*A[min(i,j)][k]
        a[j] = a[j] - hi * d * lo;
      } else if (li == k) {
        a[j] = ljk;                               // row k keeps L^T
      }
    }
  }
  if (!positive) return false;
  double y = b[perm];
#pragma unroll
  for (int j = 0; j < N - 1; ++j) {               // forward: y[i] -= L[i][j]*y[j], j ascending
    const double yj = __shfl_sync(FULL, y, j);
    if (li > j) y -= a[j] * yj;
  }
  {
    double dg = a[0];
#pragma unroll
    for (int j = 1; j < N; ++j)
      if (li == j) dg = a[j];
    y = (dg != 0) ? y / dg : 0.0;
  }
#pragma unroll
  for (int j = N - 1; j > 0; --j) {               // backward: y[i] -= L[j][i]*y[j], j descending
    const double yj = __shfl_sync(FULL, y, j);
    if (li < j) y -= a[j] * yj;
  }
  if (lane < N) x[perm] = y;
  __syncwarp();
  return true;
}

__device__ __forceinline__ bool ldlt6_solve_shfl(const double* A, const double* b, double* x) { return ldlt_solve_shfl<6>(A, b, x); }

// The same factorisation for 6 < N <= 32 with the matrix in shared memory (leading dimension LD, odd => conflict-free
// column accesses; destroyed): lane i owns row i, computes the lower entries (i, j <= i) of the trailing update
// A[i][j] - (A[i][k]*d)*A[j][k] and mirrors them, exactly the serial routine's arithmetic; the pivot is the first largest
// |diagonal| (butterfly arg-max with index tie-break).  One warp; all 32 lanes must call; every lane returns the same.
template <int N, int LD>
__device__ bool ldlt_solve_smem(double* A, const double* b, double* x) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const bool row = lane < N;
  int perm = lane;
  bool positive = true;
  __syncwarp();
  for (int k = 0; k < N; ++k) {
    // first largest |diagonal| among rows >= k: |x| bit patterns order like unsigned integers, so two 32-bit warp
    // maxima (high word, then low word among the lanes that tie on the high word) and a find-first-set do it
    const bool cand = row && lane >= k;
    const unsigned long long key = cand ? (unsigned long long)__double_as_longlong(fabs(A[lane * LD + lane])) : 0ull;
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_max_sync(FULL, hi);
    const unsigned mlo = __reduce_max_sync(FULL, hi == mhi ? lo : 0u);
    const unsigned win = __ballot_sync(FULL, cand && hi == mhi && lo == mlo);
    const int p = __ffs(win) - 1;
    if (p != k) {                                 // warp-uniform
      if (row) { const double t = A[k * LD + lane]; A[k * LD + lane] = A[p * LD + lane]; A[p * LD + lane] = t; }
      __syncwarp();
      if (row) { const double t = A[lane * LD + k]; A[lane * LD + k] = A[lane * LD + p]; A[lane * LD + p] = t; }
      const int pk = __shfl_sync(FULL, perm, k), pp = __shfl_sync(FULL, perm, p);
      if (lane == k) perm = pp; else if (lane == p) perm = pk;
      __syncwarp();
    }
    const double d = A[k * LD + k];
    if (d < 0) positive = false;
    if (d == 0) continue;
    const bool below = row && lane > k;
    double lik = 0;
    if (below) { lik = A[lane * LD + k] / d; A[lane * LD + k] = lik; }
    __syncwarp();
    if (below) {
      const double lid = lik * d;
      int j = k + 1;
      for (; j + 3 <= lane; j += 4) {             // loads first: the stores below cannot alias them, the compiler cannot know
        const double a0 = A[lane * LD + j], a1 = A[lane * LD + j + 1], a2 = A[lane * LD + j + 2], a3 = A[lane * LD + j + 3];
        const double c0 = A[j * LD + k], c1 = A[(j + 1) * LD + k], c2 = A[(j + 2) * LD + k], c3 = A[(j + 3) * LD + k];
        const double n0 = a0 - lid * c0, n1 = a1 - lid * c1, n2 = a2 - lid * c2, n3 = a3 - lid * c3;
        A[lane * LD + j] = n0; A[lane * LD + j + 1] = n1; A[lane * LD + j + 2] = n2; A[lane * LD + j + 3] = n3;
        A[j * LD + lane] = n0; A[(j + 1) * LD + lane] = n1; A[(j + 2) * LD + lane] = n2; A[(j + 3) * LD + lane] = n3;
      }
      for (; j <= lane; ++j) {
        const double nv = A[lane * LD + j] - lid * A[j * LD + k];
        A[lane * LD + j] = nv;
        A[j * LD + lane] = nv;
      }
    }
    __syncwarp();
  }
  if (!positive) return false;
  double y = row ? b[perm] : 0.0;
  for (int j = 0; j < N - 1; ++j) {               // forward: y[i] -= L[i][j]*y[j], j ascending
    const double yj = __shfl_sync(FULL, y, j);
    if (row && lane > j) y -= A[lane * LD + j] * yj;
  }
  if (row) { const double dg = A[lane * LD + lane]; y = (dg != 0) ? y / dg : 0.0; }
  for (int j = N - 1; j > 0; --j) {               // backward: y[i] -= L[j][i]*y[j], j descending
    const double yj = __shfl_sync(FULL, y, j);
    if (lane < j) y -= A[j * LD + lane] * yj;
  }
  if (row) x[perm] = y;
  __syncwarp();
  return true;
}

// =====================================================================================
// K11  PoseOptimization: one CTA per frame
// =====================================================================================
// PO_VT "virtual threads" fix the summation order of every reduction (edge e belongs to virtual thread e % PO_VT; lane
// butterfly inside each virtual warp, then the virtual warps in index order), PO_NT real threads execute them: with
// PO_NT = 128 a CTA would hold half the registers (measured: slower overall, profiles/r02x_pose_footprint.md; 256 is used).
#define PO_VT 256
// Speculative damping trials.  When a trial is rejected, g2o multiplies lambda by ni and doubles ni
// (optimization_algorithm_levenberg.cpp:120-128): the lambdas of the NEXT rejections are known in advance.  Every round
// ends with a chain of ~5 rejected trials whose matrices all differ (lambda crosses 40 orders of magnitude in five steps),
// each costing a serial 6x6 solve + exp on one warp while seven wait.  From the second solved trial of an iteration on,
// warps 0..PO_NSPEC-1 therefore solve the next PO_NSPEC lambdas of the chain side by side; a later trial whose lambda is
// found among them (exact comparison) skips its solve.  Same arithmetic, same decisions, fewer serial phases.
#define PO_NSPEC_MAX 5

// lane butterfly of NV doubles, then lane 0 stores the virtual warp's partials (no barriers: the caller places them)
template <int NV>
__device__ __forceinline__ void warp_partials(double* v, double* s_red, int vwarp) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < NV; ++k) s_red[vwarp * NV + k] = v[k];
}

struct PoseOptArgs {
  const int* edgeOfs;     // [P+1] contiguous slices ...
  const int* edgeStart;   // ... or, when non-null, [P] slice starts with
  const int* edgeCount;   //     [P] slice lengths (batched pipelines with fixed per-problem capacity)
  const float* xw;        // [Etot][3]
  const float* obs;       // [Etot][3]
  const float* invSigma2; // [Etot]
  float fx, fy, cx, cy, bf;
  float* Tcw;             // [P][16]
  uint8_t* outlier;       // [Etot]
  int* nInliers;          // [P]
  int* iters;             // [P][4]
  double* err;            // [Etot][3] scratch: residual of the last evaluated state
  int profile;            // 1: accumulate per-phase SM cycles into g_po_prof (orbx_debug_pose_opt_profile)
};

// per-phase cycle totals over all CTAs (thread 0 of each): 0 build pass, 1 reduction + unpack, 2 solve + exp (warp 0),
// 3 trial residual pass, 4 trial reduction, 5 accept/reject logic incl. replayed trials, 6 chi2 classification,
// 7 whole kernel; counts: 8 builds, 9 solved trials, 10 replayed trials, 11 CTAs
__device__ unsigned long long g_po_prof[16];
#define PO_TICK(k) do { if (A.profile && tid == 0) { const long long t1_ = clock64(); atomicAdd(&g_po_prof[k], (unsigned long long)(t1_ - t0)); t0 = t1_; } } while (0)
#define PO_COUNT(k) do { if (A.profile && tid == 0) atomicAdd(&g_po_prof[k], 1ull); } while (0)

// structural zeros of the only-pose Jacobians (row d, column i): d(u)/d(ty), d(v)/d(tx), d(ur)/d(ty)
__device__ __forceinline__ constexpr bool po_jzero(int d, int i) { return (d == 0 && i == 4) || (d == 1 && i == 3) || (d == 2 && i == 4); }

template <int PO_NT>
__global__ void __launch_bounds__(PO_NT, 512 / PO_NT) pose_opt_kernel(const PoseOptArgs A) {
  constexpr int PO_NSPEC = PO_NT / 32 < PO_NSPEC_MAX ? PO_NT / 32 : PO_NSPEC_MAX;
  static_assert(PO_VT % PO_NT == 0 && PO_NT % 32 == 0 && PO_NT >= 64, "pose_opt_kernel thread mapping");
  const int prob = blockIdx.x, tid = threadIdx.x;
  long long t0 = A.profile ? clock64() : 0ll;
  const long long tStart = t0;
  const int e0 = A.edgeStart ? A.edgeStart[prob] : A.edgeOfs[prob];
  const int E = A.edgeStart ? A.edgeCount[prob] : A.edgeOfs[prob + 1] - e0;
  __shared__ SE3d s_est, s_init;
  __shared__ double s_red[(PO_VT / 32) * 28];
  __shared__ double s_H[36], s_b[6], s_tot[28];
  // damping-trial candidates: slot w holds the solve for the lambda reached after w further rejections (PO_NSPEC below)
  __shared__ SE3d s_cTrial[PO_NSPEC_MAX];
  __shared__ double s_cA[PO_NSPEC_MAX][36], s_cX[PO_NSPEC_MAX][6], s_cLambda[PO_NSPEC_MAX];
  __shared__ int s_cOk[PO_NSPEC_MAX];
  const float* xw = A.xw + 3 * (size_t)e0;
  const float* obs = A.obs + 3 * (size_t)e0;
  const float* isg = A.invSigma2 + e0;
  uint8_t* outlier = A.outlier + e0;
  double* err = A.err + 3 * (size_t)e0;
  int* iters = A.iters + 4 * prob;
  if (tid < 4) iters[tid] = 0;
  if (E < 3) {   // nInitialCorrespondences<3: return 0, pose untouched (src/Optimizer.cc:1134-1135)
    if (tid == 0) A.nInliers[prob] = 0;
    for (int e = tid; e < E; e += PO_NT) outlier[e] = 0;
    return;
  }
  if (tid == 0) s_init = se3_from_Tcw(A.Tcw + 16 * prob);
  for (int e = tid; e < E; e += PO_NT) outlier[e] = 0;
  __syncthreads();
  const HuberD hMono = huber_make(sqrtf(5.991f)), hStereo = huber_make(sqrtf(7.815f));
  const double fx = A.fx, fy = A.fy, bf = A.bf;
  bool robust = true;
  int nBadFinal = 0;

  for (int round = 0; round < 4; ++round) {
    if (tid == 0) s_est = s_init;      // vSE3->setEstimate(toSE3Quat(pFrame->mTcw)) every round
    __syncthreads();
    // ---------------- optimize(10): modified g2o Levenberg-Marquardt ----------------
    double lambda = -1, ni = 2;
    int nBad = 0, cj = 0;
    bool ok = true;
    for (int it = 0; it < 10 && ok; ++it) {
      // computeActiveErrors + activeRobustChi2 + buildSystem in one pass over the active edges
      double acc[28];
      const SE3d est = s_est;
      PO_TICK(5);
      __syncthreads();                            // s_red is free (previous reduction read by everyone)
      for (int v = tid; v < PO_VT; v += PO_NT) {
#pragma unroll
      for (int k = 0; k < 28; ++k) acc[k] = 0;
      for (int e = v; e < E; e += PO_VT) {
        if (outlier[e]) continue;                 // level-1 edges are not active
        const double X[3] = {(double)xw[3 * e], (double)xw[3 * e + 1], (double)xw[3 * e + 2]};
        double p[3], r[3], J[18];
        se3_map(est, X, p);
        const bool st = obs[3 * e + 2] >= 0;
        reproj_error(p, obs + 3 * e, st, A.fx, A.fy, A.cx, A.cy, A.bf, r);
        err[3 * e] = r[0]; err[3 * e + 1] = r[1]; err[3 * e + 2] = r[2];
        const double om = (double)isg[e];
        const double chi = r[0] * om * r[0] + r[1] * om * r[1] + (st ? r[2] * om * r[2] : 0.0);
        double w = 1.0;
        if (robust) acc[27] += huber_rho(st ? hStereo : hMono, chi, w);
        else acc[27] += chi;
        jac_pose(p, st, false, fx, fy, bf, J);
        const int D = st ? 3 : 2;
        // J[0][4], J[1][3] and J[2][4] are exact zeros for both edge types (jac_pose): their products are +-0 and adding
        // them never changes a sum that started at +0, so those terms are left out (same bits, 30 % fewer multiply-adds)
        int idx = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          double s = 0;
#pragma unroll
          for (int d = 0; d < 3; ++d)
            if (d < D && !po_jzero(d, i)) s += J[d * 6 + i] * om * r[d];
          acc[21 + i] -= w * s;
#pragma unroll
          for (int j = i; j < 6; ++j) {
            double a = 0;
#pragma unroll
            for (int d = 0; d < 3; ++d)
              if (d < D && !po_jzero(d, i) && !po_jzero(d, j)) a += J[d * 6 + i] * (w * om) * J[d * 6 + j];
            acc[idx++] += a;
          }
        }
      }
      // fixed-shape tree: lane butterfly per virtual warp, then the virtual warps in index order, the cross-warp stage
      // done once by 28 threads
      warp_partials<28>(acc, s_red, v >> 5);
      }
      PO_TICK(0);
      PO_COUNT(8);
      __syncthreads();
      if (tid < 28) {
        double s = 0;
        for (int w = 0; w < PO_VT / 32; ++w) s += s_red[w * 28 + tid];
        s_tot[tid] = s;
      }
      __syncthreads();
      double currentChi = s_tot[27];
      const double iniChi = currentChi;
      if (tid < 36) {                             // unpack the 21 upper-triangle sums into the symmetric 6x6
        const int i = tid / 6, j = tid - 6 * i;
        const int lo = min(i, j), hi = max(i, j);
        s_H[tid] = s_tot[6 * lo - (lo * (lo - 1)) / 2 + (hi - lo)];
      } else if (tid < 42) {
        s_b[tid - 36] = s_tot[21 + tid - 36];
      }
      if (it == 0) {
        double m = 0;
        { int idx = 0; for (int i = 0; i < 6; ++i) { m = fmax(fabs(s_tot[idx]), m); idx += 6 - i; } }
        lambda = 1e-50 * m;   // computeLambdaInit with _tau = 1e-50
        ni = 2;
        nBad = 0;
      }
      __syncthreads();
      PO_TICK(1);
      double rho = 0;
      int qmax = 0;
      // The reference's damping policy (tau = 1e-50, up to 100 trials, optimization_algorithm_levenberg.cpp:47-51)
      // ends every round with ~19 rejected trials whose lambda is still far below one ulp of the diagonal.  A trial
      // whose damped matrix H + lambda*I is bit-identical to the last SOLVED trial's reproduces that trial exactly
      // (same update, same estimate, same residuals, same chi2), so it is replayed from the saved results instead of
      // being solved again; only rho's denominator, which depends on lambda itself, is recomputed.
      bool haveTrial = false, lastScalePos = false;
      double lastLambda = 0, trialChi = 0;
      int nCand = 0, cur = 0;
      do {
        bool same = haveTrial;
        if (same) {
#pragma unroll
          for (int i = 0; i < 6; ++i) same = same && (s_H[7 * i] + lambda == s_H[7 * i] + lastLambda);
        }
        if (!same) {
          PO_TICK(5);
          int hit = -1;
          for (int c = 0; c < nCand; ++c)
            if (s_cLambda[c] == lambda) hit = c;
          if (hit < 0) {
            PO_COUNT(9);
            __syncthreads();                      // every thread is done with the previous candidates
            const int ns = haveTrial ? PO_NSPEC : 1;   // the first trial of an iteration is usually accepted: no speculation
            const int wid = tid >> 5, lane = tid & 31;
            if (wid < ns) {                       // warp w: 6x6 solve for the lambda after w more rejections, then lane 0 applies the update
              double lam = lambda, nn = ni;
              for (int k = 0; k < wid; ++k) { lam *= nn; nn *= 2; }
              double* cA = s_cA[wid];
              double* cX = s_cX[wid];
              for (int i = lane; i < 36; i += 32) cA[i] = s_H[i] + ((i % 7 == 0) ? lam : 0.0);
              if (lane < 6) cX[lane] = 0;
              __syncwarp();
              const bool okSolve = ldlt6_solve_shfl(cA, s_b, cX);
              if (lane == 0) {
                s_cOk[wid] = okSolve ? 1 : 0;
                s_cLambda[wid] = lam;
                double x[6];
                for (int i = 0; i < 6; ++i) x[i] = cX[i];
                s_cTrial[wid] = se3_mul(se3_exp(x), s_est);   // push(); oplus: exp(update) * estimate
              }
            }
            __syncthreads();
            nCand = ns;
            hit = 0;
          } else {
            __syncthreads();                      // the previous trial's readers are done with s_red (compute-sanitizer racecheck)
          }
          cur = hit;
          PO_TICK(2);
          const SE3d trial = s_cTrial[cur];
          double chi[1];
          for (int v = tid; v < PO_VT; v += PO_NT) {
          chi[0] = 0;
          for (int e = v; e < E; e += PO_VT) {
            if (outlier[e]) continue;
            const double X[3] = {(double)xw[3 * e], (double)xw[3 * e + 1], (double)xw[3 * e + 2]};
            double p[3], r[3];
            se3_map(trial, X, p);
            const bool st = obs[3 * e + 2] >= 0;
            reproj_error(p, obs + 3 * e, st, A.fx, A.fy, A.cx, A.cy, A.bf, r);
            err[3 * e] = r[0]; err[3 * e + 1] = r[1]; err[3 * e + 2] = r[2];
            const double om = (double)isg[e];
            const double c = r[0] * om * r[0] + r[1] * om * r[1] + (st ? r[2] * om * r[2] : 0.0);
            double w;
            chi[0] += robust ? huber_rho(st ? hStereo : hMono, c, w) : c;
          }
          warp_partials<1>(chi, s_red, v >> 5);
          }
          PO_TICK(3);
          __syncthreads();
          {
            double t = 0;
            for (int w = 0; w < PO_VT / 32; ++w) t += s_red[w];
            chi[0] = t;
          }
          PO_TICK(4);
          trialChi = chi[0];
          lastLambda = lambda;
          haveTrial = true;
        } else {
          PO_COUNT(10);
        }
        double tempChi = trialChi;
        if (!s_cOk[cur]) tempChi = 1.7976931348623157e308;
        rho = currentChi - tempChi;
        if (same && rho < 0 && lastScalePos) {
          // replayed trial that made things worse: rho's denominator sum_j x_j (lambda x_j + b_j) + 1e-3 was positive for the
          // smaller lambda of the solved trial and only grows with lambda (x, b unchanged, x.b = x^T (H + lambda I) x > 0), so
          // rho stays negative: reject without the sum and the division
          lambda *= ni;
          ni *= 2;
          ++qmax;
          continue;
        }
        double scale = 0;
        for (int j = 0; j < 6; ++j) scale += s_cX[cur][j] * (lambda * s_cX[cur][j] + s_b[j]);
        scale += 1e-3;
        lastScalePos = scale > 0;
        rho /= scale;
        if (rho > 0 && isfinite(tempChi)) {
          const double tr = 2 * rho - 1;
          double alpha = 1. - tr * tr * tr;   // pow(2*rho-1, 3)
          alpha = fmin(alpha, 2. / 3.);
          const double scaleFactor = fmax(1. / 3., alpha);
          lambda *= scaleFactor;
          ni = 2;
          currentChi = tempChi;
          __syncthreads();                        // every thread has read s_est / s_trial of this trial
          if (tid == 0) s_est = s_cTrial[cur];    // keep the update (discardTop)
          __syncthreads();
        } else {
          lambda *= ni;                           // pop(): s_est was never overwritten
          ni *= 2;
        }
        ++qmax;
      } while (rho < 0 && qmax < 100);
      ++cj;
      if (qmax == 100 || rho == 0) { ok = false; continue; }
      if ((iniChi - currentChi) * 1e3 < iniChi) ++nBad; else nBad = 0;
      if (nBad >= 3) ok = false;
    }
    if (tid == 0) iters[round] = cj;
    PO_TICK(5);
    // ---------------- chi2 classification (src/Optimizer.cc:1157-1255) ----------------
    const SE3d est = s_est;
    double bad[1] = {0};
    for (int e = tid; e < E; e += PO_NT) {
      const bool st = obs[3 * e + 2] >= 0;
      if (outlier[e]) {                           // e->computeError() for edges that sat out
        const double X[3] = {(double)xw[3 * e], (double)xw[3 * e + 1], (double)xw[3 * e + 2]};
        double p[3], r[3];
        se3_map(est, X, p);
        reproj_error(p, obs + 3 * e, st, A.fx, A.fy, A.cx, A.cy, A.bf, r);
        err[3 * e] = r[0]; err[3 * e + 1] = r[1]; err[3 * e + 2] = r[2];
      }
      const double om = (double)isg[e];
      const double* r = err + 3 * e;
      const float chi2 = (float)(r[0] * om * r[0] + r[1] * om * r[1] + (st ? r[2] * om * r[2] : 0.0));
      const bool isBad = chi2 > (st ? 7.815f : 5.991f);
      outlier[e] = isBad ? 1 : 0;
      bad[0] += isBad ? 1.0 : 0.0;
    }
    block_sum<1>(bad, s_red);
    nBadFinal = (int)bad[0];
    PO_TICK(6);
    if (round == 2) robust = false;               // e->setRobustKernel(0)
    if (E < 10) break;                            // optimizer.edges().size()<10
  }
  if (tid == 0) {
    se3_to_Tcw(s_est, A.Tcw + 16 * prob);
    A.nInliers[prob] = E - nBadFinal;
    if (A.profile) {
      atomicAdd(&g_po_prof[7], (unsigned long long)(clock64() - tStart));
      atomicAdd(&g_po_prof[11], 1ull);
    }
  }
}

static int g_po_profile = 0;   // test / tuning hook, off in production

// =====================================================================================
// K12  LocalBundleAdjustment: one CTA (1024 threads) per problem; all state in global/L2 scratch
// =====================================================================================
#define LBA_NT 1024

struct LbaArgs {
  int K, M, E, nFree, n;      // n = 6*nFree
  float* kfT;                 // [K][16] in/out
  const uint8_t* kfFixed;
  float* mpXyz;               // [M][3] in/out
  const int *ekf, *emp;
  const float *obs, *invSigma2;
  float fx, fy, cx, cy, bf;
  double lambdaInit;
  const volatile uint8_t* stop;   // device-visible (mapped host) or null
  // graph indices (host-built)
  const int* hidx;            // [K] index among free poses or -1
  const int* ptOfs;           // [M+1] edges of a point whose pose is free (for the Schur complement)
  const int* ptEdges;
  const int* ptAllOfs;        // [M+1] all edges of a point
  const int* ptAllEdges;
  const int* kfOfs;           // [nFree+1] edges of a free pose
  const int* kfEdges;
  const int* obsEdge;         // [M][nFree] edge id of (point, free pose) or -1
  // scratch (doubles unless noted)
  SE3d *pose, *poseBak;       // [K]
  double *pt, *ptBak;         // [M][3]
  double* err;                // [E][3]
  double *Ji, *Jj;            // [E][9], [E][18]
  double *wom, *omr;          // [E], [E][3]
  double *Hpl, *BD;           // [E][18] (6x3), B*Dinv
  double *Hll, *Dinv;         // [M][9]
  double* Hpp;                // [nFree][36]
  double *b, *x;              // [n + 3M]
  double *S, *bs;             // [n][n], [n]
  double* db;                 // [M][3]
  // cooperative (multi-CTA) kernel only
  double* tmp;                // [max(E, n + 3M)] per-item terms of the canonical sums
  double* part;               // [3][32] per-warp partial sums
  double* gmax;               // [1] max |diagonal|
  int* gflag;                 // [4] stop flag / solver status broadcast
  unsigned long long* prof;   // [16] optional per-phase nanosecond totals (ORBX_LBA_PROFILE=1), else null
  const int* blkOfs;          // [nBlocks + 1] offsets of the Schur blocks' edge-pair lists
  int* pairE2;                // [blkOfs[nBlocks]]
  // outputs
  uint8_t* edgeBad;
  int* iters;                 // [2]
  int* status;
};

__device__ __forceinline__ bool lba_stop(const LbaArgs& A, int* s_flag) {
  __syncthreads();
  if (threadIdx.x == 0) *s_flag = (A.stop && *A.stop) ? 1 : 0;
  __syncthreads();
  return *s_flag != 0;
}

// residuals + robust chi2 of every edge at the current estimate
__device__ double lba_errors_chi(const LbaArgs& A, const HuberD& hM, const HuberD& hS, double* s_red) {
  double chi[1] = {0};
  for (int e = threadIdx.x; e < A.E; e += LBA_NT) {
    double p[3], r[3];
    se3_map(A.pose[A.ekf[e]], A.pt + 3 * (size_t)A.emp[e], p);
    const bool st = A.obs[3 * e + 2] >= 0;
    reproj_error(p, A.obs + 3 * e, st, A.fx, A.fy, A.cx, A.cy, A.bf, r);
    A.err[3 * e] = r[0]; A.err[3 * e + 1] = r[1]; A.err[3 * e + 2] = r[2];
    const double om = (double)A.invSigma2[e];
    const double c = r[0] * om * r[0] + r[1] * om * r[1] + (st ? r[2] * om * r[2] : 0.0);
    double w;
    chi[0] += huber_rho(st ? hS : hM, c, w);
  }
  block_sum<1>(chi, s_red);
  return chi[0];
}

static __device__ void lba_single_cta(const LbaArgs& A) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, NW = LBA_NT / 32;
  __shared__ double s_red[(LBA_NT / 32) * 2];
  __shared__ int s_flag;
  __shared__ double s_scal[4];
  const HuberD hM = huber_make(sqrtf(5.991f)), hS = huber_make(sqrtf(7.815f));
  const double fx = A.fx, fy = A.fy, bf = A.bf;
  const int n = A.n, NX = A.n + 3 * A.M;

  if (tid == 0) { A.iters[0] = A.iters[1] = 0; *A.status = 0; }
  for (int e = tid; e < A.E; e += LBA_NT) A.edgeBad[e] = 0;
  if (lba_stop(A, &s_flag)) {   // if(pbStopFlag) if(*pbStopFlag) return;  (src/Optimizer.cc:2195-2197)
    if (tid == 0) *A.status = 1;
    return;
  }
  for (int k = tid; k < A.K; k += LBA_NT) A.pose[k] = se3_from_Tcw(A.kfT + 16 * k);
  for (int i = tid; i < 3 * A.M; i += LBA_NT) A.pt[i] = (double)A.mpXyz[i];
  __syncthreads();

  for (int call = 0; call < 2; ++call) {
    const int iterations = call == 0 ? 5 : 10;
    if (call == 1 && lba_stop(A, &s_flag)) break;   // bDoMore
    double lambda = -1, ni = 2;
    int nBad = 0, cj = 0;
    bool ok = true;
    for (int it = 0; it < iterations && ok; ++it) {
      if (lba_stop(A, &s_flag)) break;
      // ---- computeActiveErrors / activeRobustChi2 ----
      double currentChi = lba_errors_chi(A, hM, hS, s_red);
      const double iniChi = currentChi;
      // ---- buildSystem: per-edge Jacobians, then deterministic gathers per point / per pose ----
      for (int e = tid; e < A.E; e += LBA_NT) {
        const int k = A.ekf[e];
        double p[3], R[9];
        se3_map(A.pose[k], A.pt + 3 * (size_t)A.emp[e], p);
        quat_to_R(A.pose[k].r, R);
        const bool st = A.obs[3 * e + 2] >= 0;
        const int D = st ? 3 : 2;
        double* Ji = A.Ji + 9 * (size_t)e;
        double* Jj = A.Jj + 18 * (size_t)e;
        const double x = p[0], y = p[1], z = p[2];
        if (!st) {
          const double j00 = -(fx / z), j02 = fx * x / (z * z), j11 = -(fy / z), j12 = fy * y / (z * z);
          for (int c = 0; c < 3; ++c) {
            Ji[c] = j00 * R[c] + j02 * R[6 + c];
            Ji[3 + c] = j11 * R[3 + c] + j12 * R[6 + c];
            Ji[6 + c] = 0;
          }
        } else {
          const double z2 = z * z;
          for (int c = 0; c < 3; ++c) {
            Ji[c] = -fx * R[c] / z + fx * x * R[6 + c] / z2;
            Ji[3 + c] = -fy * R[3 + c] / z + fy * y * R[6 + c] / z2;
            Ji[6 + c] = Ji[c] - bf * R[6 + c] / z2;
          }
        }
        jac_pose(p, st, true, fx, fy, bf, Jj);
        if (!st) for (int c = 12; c < 18; ++c) Jj[c] = 0;
        const double om = (double)A.invSigma2[e];
        const double* r = A.err + 3 * e;
        const double c2 = r[0] * om * r[0] + r[1] * om * r[1] + (st ? r[2] * om * r[2] : 0.0);
        double w;
        huber_rho(st ? hS : hM, c2, w);
        for (int d = 0; d < 3; ++d) A.omr[3 * e + d] = (d < D) ? -om * r[d] * w : 0.0;
        const double wo = w * om;
        A.wom[e] = wo;
        // pose-landmark block of this observation: Jj^T (w Omega) Ji   (6x3)
        double* hpl = A.Hpl + 18 * (size_t)e;
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 3; ++j) {
            double a = 0;
            for (int d = 0; d < D; ++d) a += Jj[d * 6 + i] * wo * Ji[d * 3 + j];
            hpl[i * 3 + j] = a;
          }
      }
      __syncthreads();
      // points: Hll, b_l (every edge of the point, fixed poses included)
      for (int m = tid; m < A.M; m += LBA_NT) {
        double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, bl[3] = {0, 0, 0};
        for (int q = A.ptAllOfs[m]; q < A.ptAllOfs[m + 1]; ++q) {
          const int e = A.ptAllEdges[q];
          const int D = A.obs[3 * e + 2] >= 0 ? 3 : 2;
          const double* Ji = A.Ji + 9 * (size_t)e;
          const double wo = A.wom[e];
          const double* omr = A.omr + 3 * e;
          for (int i = 0; i < 3; ++i) {
            double s = 0;
            for (int d = 0; d < D; ++d) s += Ji[d * 3 + i] * omr[d];
            bl[i] += s;
            for (int j = 0; j < 3; ++j) {
              double a = 0;
              for (int d = 0; d < D; ++d) a += Ji[d * 3 + i] * wo * Ji[d * 3 + j];
              H[i * 3 + j] += a;
            }
          }
        }
        for (int i = 0; i < 9; ++i) A.Hll[9 * (size_t)m + i] = H[i];
        for (int i = 0; i < 3; ++i) A.b[n + 3 * m + i] = bl[i];
      }
      // free poses: Hpp (upper 21) + b_p, one warp per pose
      for (int hk = wid; hk < A.nFree; hk += NW) {
        double acc[27];
#pragma unroll
        for (int i = 0; i < 27; ++i) acc[i] = 0;
        for (int q = A.kfOfs[hk] + lane; q < A.kfOfs[hk + 1]; q += 32) {
          const int e = A.kfEdges[q];
          const int D = A.obs[3 * e + 2] >= 0 ? 3 : 2;
          const double* Jj = A.Jj + 18 * (size_t)e;
          const double wo = A.wom[e];
          const double* omr = A.omr + 3 * e;
          int idx = 0;
          for (int i = 0; i < 6; ++i) {
            double s = 0;
            for (int d = 0; d < D; ++d) s += Jj[d * 6 + i] * omr[d];
            acc[21 + i] += s;
            for (int j = i; j < 6; ++j) {
              double a = 0;
              for (int d = 0; d < D; ++d) a += Jj[d * 6 + i] * wo * Jj[d * 6 + j];
              acc[idx++] += a;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 27; ++i)
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
        if (lane == 0) {
          int idx = 0;
          for (int i = 0; i < 6; ++i)
            for (int j = i; j < 6; ++j) {
              A.Hpp[36 * (size_t)hk + i * 6 + j] = acc[idx];
              A.Hpp[36 * (size_t)hk + j * 6 + i] = acc[idx];
              ++idx;
            }
          for (int i = 0; i < 6; ++i) A.b[6 * hk + i] = acc[21 + i];
        }
      }
      __syncthreads();
      if (it == 0) {
        if (A.lambdaInit > 0) lambda = A.lambdaInit;
        else {
          double m[1] = {0};   // max |diag|: reduce with max (emulated through block_sum-free tree below)
          for (int i = tid; i < A.nFree * 6; i += LBA_NT) m[0] = fmax(m[0], fabs(A.Hpp[36 * (size_t)(i / 6) + (i % 6) * 7]));
          for (int i = tid; i < A.M * 3; i += LBA_NT) m[0] = fmax(m[0], fabs(A.Hll[9 * (size_t)(i / 3) + (i % 3) * 4]));
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) m[0] = fmax(m[0], __shfl_xor_sync(0xffffffffu, m[0], o));
          __syncthreads();
          if (lane == 0) s_red[wid] = m[0];
          __syncthreads();
          double mm = 0;
          for (int w = 0; w < NW; ++w) mm = fmax(mm, s_red[w]);
          __syncthreads();
          lambda = 1e-50 * mm;
        }
        ni = 2;
        nBad = 0;
      }
      double rho = 0;
      int qmax = 0;
      bool stopped = false;
      do {
        // ---- push ----
        for (int k = tid; k < A.K; k += LBA_NT) A.poseBak[k] = A.pose[k];
        for (int i = tid; i < 3 * A.M; i += LBA_NT) A.ptBak[i] = A.pt[i];
        // ---- solve with Schur complement (block_solver.hpp:354-486) ----
        // landmarks: Dinv = (Hll + lambda I)^-1 ; db = Dinv b_l
        for (int m = tid; m < A.M; m += LBA_NT) {
          double Dm[9];
          for (int i = 0; i < 9; ++i) Dm[i] = A.Hll[9 * (size_t)m + i];
          Dm[0] += lambda; Dm[4] += lambda; Dm[8] += lambda;
          const double c00 = Dm[4] * Dm[8] - Dm[5] * Dm[7], c01 = Dm[5] * Dm[6] - Dm[3] * Dm[8], c02 = Dm[3] * Dm[7] - Dm[4] * Dm[6];
          const double det = Dm[0] * c00 + Dm[1] * c01 + Dm[2] * c02, id = 1.0 / det;
          double* Di = A.Dinv + 9 * (size_t)m;
          Di[0] = c00 * id; Di[1] = (Dm[2] * Dm[7] - Dm[1] * Dm[8]) * id; Di[2] = (Dm[1] * Dm[5] - Dm[2] * Dm[4]) * id;
          Di[3] = c01 * id; Di[4] = (Dm[0] * Dm[8] - Dm[2] * Dm[6]) * id; Di[5] = (Dm[2] * Dm[3] - Dm[0] * Dm[5]) * id;
          Di[6] = c02 * id; Di[7] = (Dm[1] * Dm[6] - Dm[0] * Dm[7]) * id; Di[8] = (Dm[0] * Dm[4] - Dm[1] * Dm[3]) * id;
          const double* bl = A.b + n + 3 * m;
          for (int i = 0; i < 3; ++i) A.db[3 * m + i] = Di[i * 3] * bl[0] + Di[i * 3 + 1] * bl[1] + Di[i * 3 + 2] * bl[2];
        }
        __syncthreads();
        // BD[e] = B_e Dinv_m for edges of free poses
        for (int e = tid; e < A.E; e += LBA_NT) {
          if (A.hidx[A.ekf[e]] < 0) continue;
          const double* B = A.Hpl + 18 * (size_t)e;
          const double* Di = A.Dinv + 9 * (size_t)A.emp[e];
          double* BD = A.BD + 18 * (size_t)e;
          for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 3; ++j) BD[i * 3 + j] = B[i * 3] * Di[j] + B[i * 3 + 1] * Di[3 + j] + B[i * 3 + 2] * Di[6 + j];
        }
        __syncthreads();
        // reduced system: one warp per (free pose i, free pose j >= i) block, lanes over pose i's edges
        const int nBlocks = A.nFree * (A.nFree + 1) / 2;
        for (int blk = wid; blk < nBlocks; blk += NW) {
          int bi = 0, rem = blk;
          while (rem >= A.nFree - bi) { rem -= A.nFree - bi; ++bi; }
          const int bj = bi + rem;
          double acc[36];
#pragma unroll
          for (int i = 0; i < 36; ++i) acc[i] = 0;
          for (int q = A.kfOfs[bi] + lane; q < A.kfOfs[bi + 1]; q += 32) {
            const int e1 = A.kfEdges[q];
            const int e2 = A.obsEdge[(size_t)A.emp[e1] * A.nFree + bj];
            if (e2 < 0) continue;
            const double* BD = A.BD + 18 * (size_t)e1;
            const double* B2 = A.Hpl + 18 * (size_t)e2;
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
              for (int j = 0; j < 6; ++j)
                acc[i * 6 + j] += BD[i * 3] * B2[j * 3] + BD[i * 3 + 1] * B2[j * 3 + 1] + BD[i * 3 + 2] * B2[j * 3 + 2];
          }
#pragma unroll
          for (int i = 0; i < 36; ++i)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
          if (lane == 0) {
            for (int i = 0; i < 6; ++i)
              for (int j = 0; j < 6; ++j) {
                double v = -acc[i * 6 + j];
                if (bi == bj) v += A.Hpp[36 * (size_t)bi + i * 6 + j] + (i == j ? lambda : 0.0);
                A.S[(size_t)(6 * bi + i) * n + 6 * bj + j] = v;
                A.S[(size_t)(6 * bj + j) * n + 6 * bi + i] = v;
              }
          }
        }
        // b_schur = b_p - sum_e B_e db_m, one warp per free pose
        for (int hk = wid; hk < A.nFree; hk += NW) {
          double acc[6] = {0, 0, 0, 0, 0, 0};
          for (int q = A.kfOfs[hk] + lane; q < A.kfOfs[hk + 1]; q += 32) {
            const int e = A.kfEdges[q];
            const double* B = A.Hpl + 18 * (size_t)e;
            const double* db = A.db + 3 * (size_t)A.emp[e];
            for (int i = 0; i < 6; ++i) acc[i] += B[i * 3] * db[0] + B[i * 3 + 1] * db[1] + B[i * 3 + 2] * db[2];
          }
#pragma unroll
          for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
          if (lane == 0)
            for (int i = 0; i < 6; ++i) A.bs[6 * hk + i] = A.b[6 * hk + i] - acc[i];
        }
        __syncthreads();
        // dense LDL^T of S in place (lower), un-pivoted (stand-in for SimplicialLDLT); fails on a zero pivot
        if (tid == 0) s_flag = 1;
        __syncthreads();
        for (int k = 0; k < n; ++k) {
          const double d = A.S[(size_t)k * n + k];
          if (d == 0 || !isfinite(d)) { if (tid == 0) s_flag = 0; break; }
          for (int i = k + 1 + tid; i < n; i += LBA_NT) A.S[(size_t)i * n + k] /= d;
          __syncthreads();
          const int mrem = n - k - 1;
          for (int t = tid; t < mrem * mrem; t += LBA_NT) {
            const int i = k + 1 + t / mrem, j = k + 1 + t % mrem;
            if (j <= i) A.S[(size_t)i * n + j] -= A.S[(size_t)i * n + k] * d * A.S[(size_t)j * n + k];
          }
          __syncthreads();
        }
        __syncthreads();
        const bool solved = s_flag != 0;
        for (int i = tid; i < NX; i += LBA_NT) A.x[i] = 0;
        __syncthreads();
        if (solved && n > 0) {
          // forward (column sweep keeps the j-ascending subtraction order), diagonal, backward (j descending)
          for (int i = tid; i < n; i += LBA_NT) A.x[i] = A.bs[i];
          __syncthreads();
          for (int j = 0; j < n; ++j) {
            const double yj = A.x[j];
            for (int i = j + 1 + tid; i < n; i += LBA_NT) A.x[i] -= A.S[(size_t)i * n + j] * yj;
            __syncthreads();
          }
          for (int i = tid; i < n; i += LBA_NT) A.x[i] /= A.S[(size_t)i * n + i];
          __syncthreads();
          for (int j = n - 1; j >= 0; --j) {
            const double yj = A.x[j];
            for (int i = tid; i < j; i += LBA_NT) A.x[i] -= A.S[(size_t)j * n + i] * yj;
            __syncthreads();
          }
          // landmarks: x_l = Dinv (b_l - sum_e B_e^T x_p)
          for (int m = tid; m < A.M; m += LBA_NT) {
            double cl[3] = {A.b[n + 3 * m], A.b[n + 3 * m + 1], A.b[n + 3 * m + 2]};
            for (int q = A.ptOfs[m]; q < A.ptOfs[m + 1]; ++q) {
              const int e = A.ptEdges[q];
              const int hk = A.hidx[A.ekf[e]];
              const double* B = A.Hpl + 18 * (size_t)e;
              for (int j = 0; j < 3; ++j)
                for (int i = 0; i < 6; ++i) cl[j] -= B[i * 3 + j] * A.x[6 * hk + i];
            }
            const double* Di = A.Dinv + 9 * (size_t)m;
            for (int i = 0; i < 3; ++i) A.x[n + 3 * m + i] = Di[i * 3] * cl[0] + Di[i * 3 + 1] * cl[1] + Di[i * 3 + 2] * cl[2];
          }
        }
        __syncthreads();
        // ---- update (oplus) ----
        for (int k = tid; k < A.K; k += LBA_NT)
          if (A.hidx[k] >= 0) A.pose[k] = se3_mul(se3_exp(A.x + 6 * A.hidx[k]), A.pose[k]);
        for (int i = tid; i < 3 * A.M; i += LBA_NT) A.pt[i] += A.x[n + i];
        __syncthreads();
        double tempChi = lba_errors_chi(A, hM, hS, s_red);
        if (!solved) tempChi = 1.7976931348623157e308;
        rho = currentChi - tempChi;
        double sc[1] = {0};
        for (int j = tid; j < NX; j += LBA_NT) sc[0] += A.x[j] * (lambda * A.x[j] + A.b[j]);
        block_sum<1>(sc, s_red);
        const double scale = sc[0] + 1e-3;
        rho /= scale;
        if (rho > 0 && isfinite(tempChi)) {
          const double tr = 2 * rho - 1;
          double alpha = 1. - tr * tr * tr;   // pow(2*rho-1, 3)
          alpha = fmin(alpha, 2. / 3.);
          const double scaleFactor = fmax(1. / 3., alpha);
          lambda *= scaleFactor;
          ni = 2;
          currentChi = tempChi;
        } else {
          lambda *= ni;
          ni *= 2;
          for (int k = tid; k < A.K; k += LBA_NT) A.pose[k] = A.poseBak[k];   // pop
          for (int i = tid; i < 3 * A.M; i += LBA_NT) A.pt[i] = A.ptBak[i];
        }
        __syncthreads();
        ++qmax;
        stopped = lba_stop(A, &s_flag);
      } while (rho < 0 && qmax < 100 && !stopped);
      ++cj;
      if (qmax == 100 || rho == 0) { ok = false; continue; }
      if ((iniChi - currentChi) * 1e3 < iniChi) ++nBad; else nBad = 0;
      if (nBad >= 3) ok = false;
    }
    if (tid == 0) A.iters[call] = cj;
  }
  __syncthreads();
  // ---- final chi2 / depth test on the residuals of the last evaluated state (src/Optimizer.cc:2295-2352) ----
  double bad[1] = {0};
  for (int e = tid; e < A.E; e += LBA_NT) {
    const bool st = A.obs[3 * e + 2] >= 0;
    const double om = (double)A.invSigma2[e];
    const double* r = A.err + 3 * e;
    const double c = r[0] * om * r[0] + r[1] * om * r[1] + (st ? r[2] * om * r[2] : 0.0);
    double p[3];
    se3_map(A.pose[A.ekf[e]], A.pt + 3 * (size_t)A.emp[e], p);
    const bool isBad = c > (st ? 7.815 : 5.991) || !(p[2] > 0.0);
    A.edgeBad[e] = isBad ? 1 : 0;
    bad[0] += isBad ? 1.0 : 0.0;
  }
  block_sum<1>(bad, s_red);
  if (bad[0] >= A.E * 0.5) {
    if (tid == 0) *A.status = 2;
    return;
  }
  for (int k = tid; k < A.K; k += LBA_NT)
    if (!A.kfFixed[k]) se3_to_Tcw(A.pose[k], A.kfT + 16 * k);
  for (int i = tid; i < 3 * A.M; i += LBA_NT) A.mpXyz[i] = (float)A.pt[i];
  (void)s_scal;
}

__global__ void __launch_bounds__(LBA_NT) lba_kernel(const LbaArgs A) { lba_single_cta(A); }

// Many problems per launch, one 1024-thread CTA each (the keyframe-rate step of S independent streams): the problem's
// argument block is staged in shared memory once, then the CTA runs exactly the single-problem code, so every result is
// bit-identical to orbx_local_ba's single-CTA path.
__global__ void __launch_bounds__(LBA_NT, 1) lba_batch_kernel(const LbaArgs* __restrict__ args) {
  __shared__ LbaArgs sA;
  {
    const int* src = reinterpret_cast<const int*>(args + blockIdx.x);
    int* dst = reinterpret_cast<int*>(&sA);
    for (int i = threadIdx.x; i < (int)(sizeof(LbaArgs) / sizeof(int)); i += LBA_NT) dst[i] = src[i];
  }
  __syncthreads();
  lba_single_cta(sA);
}

// =====================================================================================
// K12b  LocalBundleAdjustment, cooperative multi-CTA form: the same phases as lba_kernel, spread over the whole GPU
// (one 256-thread CTA per SM, grid-wide barriers between phases).  Every sum keeps lba_kernel's canonical order:
//   * per-pose / per-block accumulations are warp-local (32 accumulators + butterfly): any warp of the grid can own one;
//   * the three NT = 1024 sums (robust chi2, the LM scale, the bad-edge count) are evaluated from per-item terms by the
//     first 1024 threads of the grid exactly as 1024 threads of one CTA would (item i -> accumulator i % 1024, butterfly
//     per 32, the 32 group sums added in order);
//   * the dense LDL^T of the reduced camera system and its substitutions run in CTA 0, in shared memory when the system
//     fits (n <= LBC_MAX_SMEM_N), element for element as before.
// =====================================================================================
#define LBC_NT 256
#define LBC_MAX_SMEM_N 150

__device__ __forceinline__ unsigned long long lbc_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// phase timer (thread 0 of the grid only, when profiling is on): adds the time since `t0` to slot `k`, restarts t0
#define LBC_TICK(k) do { if (A.prof && gt == 0) { const unsigned long long t1_ = lbc_now(); A.prof[k] += t1_ - t0; t0 = t1_; } } while (0)

// canonical NT = 1024 sum of v[0..count): returns the same value on every thread of the grid
__device__ double lbc_sum1024(cg::grid_group& grid, const double* v, int count, double* part) {
  const int gt = blockIdx.x * LBC_NT + threadIdx.x;
  grid.sync();                                        // the terms are complete
  if (gt < 1024) {
    double s = 0;
    for (int i = gt; i < count; i += 1024) s += v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((gt & 31) == 0) part[gt >> 5] = s;
  }
  grid.sync();
  double tot = 0;
  for (int w = 0; w < 32; ++w) tot += part[w];
  return tot;
}

// residuals + per-edge robust chi2 terms of every edge at the current estimate, then their canonical sum
__device__ double lbc_errors_chi(cg::grid_group& grid, const LbaArgs& A, const HuberD& hM, const HuberD& hS) {
  const int gt = blockIdx.x * LBC_NT + threadIdx.x, GS = gridDim.x * LBC_NT;
  for (int e = gt; e < A.E; e += GS) {
    double p[3], r[3];
    se3_map(A.pose[A.ekf[e]], A.pt + 3 * (size_t)A.emp[e], p);
    const bool st = A.obs[3 * e + 2] >= 0;
    reproj_error(p, A.obs + 3 * e, st, A.fx, A.fy, A.cx, A.cy, A.bf, r);
    A.err[3 * e] = r[0]; A.err[3 * e + 1] = r[1]; A.err[3 * e + 2] = r[2];
    const double om = (double)A.invSigma2[e];
    const double c = r[0] * om * r[0] + r[1] * om * r[1] + (st ? r[2] * om * r[2] : 0.0);
    double w;
    A.tmp[e] = huber_rho(st ? hS : hM, c, w);
  }
  return lbc_sum1024(grid, A.tmp, A.E, A.part);
}

__device__ bool lbc_stop(cg::grid_group& grid, const LbaArgs& A, int& seq) {
  const int slot = seq & 1;
  ++seq;
  if (blockIdx.x == 0 && threadIdx.x == 0) A.gflag[slot] = (A.stop && *A.stop) ? 1 : 0;
  grid.sync();
  return A.gflag[slot] != 0;
}

__global__ void __launch_bounds__(LBC_NT) lba_coop_kernel(const LbaArgs A) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) double s_S[];       // CTA 0: reduced system + right-hand side when n <= LBC_MAX_SMEM_N
  __shared__ int s_flag;
  const int tid = threadIdx.x, lane = tid & 31;
  const int gt = blockIdx.x * LBC_NT + tid, GS = gridDim.x * LBC_NT;
  const int gw = gt >> 5, GW = GS >> 5;
  const HuberD hM = huber_make(sqrtf(5.991f)), hS = huber_make(sqrtf(7.815f));
  const double fx = A.fx, fy = A.fy, bf = A.bf;
  const int n = A.n, NX = A.n + 3 * A.M;
  int stopSeq = 0;
  unsigned long long t0 = A.prof ? lbc_now() : 0ull;

  if (gt == 0) { A.iters[0] = A.iters[1] = 0; *A.status = 0; }
  for (int e = gt; e < A.E; e += GS) A.edgeBad[e] = 0;
  if (lbc_stop(grid, A, stopSeq)) {   // if(pbStopFlag) if(*pbStopFlag) return;  (src/Optimizer.cc:2195-2197)
    if (gt == 0) *A.status = 1;
    return;
  }
  for (int k = gt; k < A.K; k += GS) A.pose[k] = se3_from_Tcw(A.kfT + 16 * k);
  for (int i = gt; i < 3 * A.M; i += GS) A.pt[i] = (double)A.mpXyz[i];
  {   // static structure of the Schur product: for block (bi, bj) and the q-th edge of pose bi, the edge of the same point in bj
    const int nBlocks0 = A.nFree * (A.nFree + 1) / 2;
    for (int blk = gw; blk < nBlocks0; blk += GW) {
      int bi = 0, rem = blk;
      while (rem >= A.nFree - bi) { rem -= A.nFree - bi; ++bi; }
      const int bj = bi + rem;
      int* pr = A.pairE2 + A.blkOfs[blk] - A.kfOfs[bi];
      for (int q = A.kfOfs[bi] + lane; q < A.kfOfs[bi + 1]; q += 32) pr[q] = A.obsEdge[(size_t)A.emp[A.kfEdges[q]] * A.nFree + bj];
    }
  }
  grid.sync();

  for (int call = 0; call < 2; ++call) {
    const int iterations = call == 0 ? 5 : 10;
    if (call == 1 && lbc_stop(grid, A, stopSeq)) break;   // bDoMore
    double lambda = -1, ni = 2;
    int nBad = 0, cj = 0;
    bool ok = true;
    for (int it = 0; it < iterations && ok; ++it) {
      if (lbc_stop(grid, A, stopSeq)) break;
      // ---- computeActiveErrors / activeRobustChi2 ----
      LBC_TICK(0);
      double currentChi = lbc_errors_chi(grid, A, hM, hS);
      const double iniChi = currentChi;
      LBC_TICK(1);
      // ---- buildSystem: per-edge Jacobians, then deterministic gathers per point / per pose ----
      for (int e = gt; e < A.E; e += GS) {
        const int k = A.ekf[e];
        double p[3], R[9];
        se3_map(A.pose[k], A.pt + 3 * (size_t)A.emp[e], p);
        quat_to_R(A.pose[k].r, R);
        const bool st = A.obs[3 * e + 2] >= 0;
        const int D = st ? 3 : 2;
        double* Ji = A.Ji + 9 * (size_t)e;
        double* Jj = A.Jj + 18 * (size_t)e;
        const double x = p[0], y = p[1], z = p[2];
        if (!st) {
          const double j00 = -(fx / z), j02 = fx * x / (z * z), j11 = -(fy / z), j12 = fy * y / (z * z);
          for (int c = 0; c < 3; ++c) {
            Ji[c] = j00 * R[c] + j02 * R[6 + c];
            Ji[3 + c] = j11 * R[3 + c] + j12 * R[6 + c];
            Ji[6 + c] = 0;
          }
        } else {
          const double z2 = z * z;
          for (int c = 0; c < 3; ++c) {
            Ji[c] = -fx * R[c] / z + fx * x * R[6 + c] / z2;
            Ji[3 + c] = -fy * R[3 + c] / z + fy * y * R[6 + c] / z2;
            Ji[6 + c] = Ji[c] - bf * R[6 + c] / z2;
          }
        }
        jac_pose(p, st, true, fx, fy, bf, Jj);
        if (!st) for (int c = 12; c < 18; ++c) Jj[c] = 0;
        const double om = (double)A.invSigma2[e];
        const double* r = A.err + 3 * e;
        const double c2 = r[0] * om * r[0] + r[1] * om * r[1] + (st ? r[2] * om * r[2] : 0.0);
        double w;
        huber_rho(st ? hS : hM, c2, w);
        for (int d = 0; d < 3; ++d) A.omr[3 * e + d] = (d < D) ? -om * r[d] * w : 0.0;
        const double wo = w * om;
        A.wom[e] = wo;
        double* hpl = A.Hpl + 18 * (size_t)e;
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 3; ++j) {
            double a = 0;
            for (int d = 0; d < D; ++d) a += Jj[d * 6 + i] * wo * Ji[d * 3 + j];
            hpl[i * 3 + j] = a;
          }
      }
      grid.sync();
      LBC_TICK(2);
      // points: Hll, b_l (every edge of the point, fixed poses included)
      for (int m = gt; m < A.M; m += GS) {
        double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, bl[3] = {0, 0, 0};
        for (int q = A.ptAllOfs[m]; q < A.ptAllOfs[m + 1]; ++q) {
          const int e = A.ptAllEdges[q];
          const int D = A.obs[3 * e + 2] >= 0 ? 3 : 2;
          const double* Ji = A.Ji + 9 * (size_t)e;
          const double wo = A.wom[e];
          const double* omr = A.omr + 3 * e;
          for (int i = 0; i < 3; ++i) {
            double s = 0;
            for (int d = 0; d < D; ++d) s += Ji[d * 3 + i] * omr[d];
            bl[i] += s;
            for (int j = 0; j < 3; ++j) {
              double a = 0;
              for (int d = 0; d < D; ++d) a += Ji[d * 3 + i] * wo * Ji[d * 3 + j];
              H[i * 3 + j] += a;
            }
          }
        }
        for (int i = 0; i < 9; ++i) A.Hll[9 * (size_t)m + i] = H[i];
        for (int i = 0; i < 3; ++i) A.b[n + 3 * m + i] = bl[i];
      }
      // free poses: Hpp (upper 21) + b_p.  Every one of the 27 sums of a pose keeps lba_kernel's order (the pose's edges
      // dealt round-robin to 32 lanes, butterfly), but each sum gets its own warp: 27 short loops instead of one long one
      for (int item = gw; item < A.nFree * 27; item += GW) {
        const int hk = item / 27, v = item - 27 * hk;
        int vi = 0, vj = 0;
        if (v < 21) { int r = v; while (r >= 6 - vi) { r -= 6 - vi; ++vi; } vj = vi + r; }
        else vi = v - 21;
        double acc = 0;
        // (a mono edge has Jj[12..17] = 0 and omr[2] = 0, so its third term is an exact +0: the loop needs no branch on D)
        if (v < 21) {
#pragma unroll 4
          for (int q = A.kfOfs[hk] + lane; q < A.kfOfs[hk + 1]; q += 32) {
            const int e = A.kfEdges[q];
            const double* Jj = A.Jj + 18 * (size_t)e;
            const double wo = A.wom[e];
            double t = 0;
#pragma unroll
            for (int d = 0; d < 3; ++d) t += Jj[d * 6 + vi] * wo * Jj[d * 6 + vj];
            acc += t;
          }
        } else {
#pragma unroll 4
          for (int q = A.kfOfs[hk] + lane; q < A.kfOfs[hk + 1]; q += 32) {
            const int e = A.kfEdges[q];
            const double* Jj = A.Jj + 18 * (size_t)e;
            const double* omr = A.omr + 3 * e;
            double t = 0;
#pragma unroll
            for (int d = 0; d < 3; ++d) t += Jj[d * 6 + vi] * omr[d];
            acc += t;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
          if (v < 21) {
            A.Hpp[36 * (size_t)hk + vi * 6 + vj] = acc;
            A.Hpp[36 * (size_t)hk + vj * 6 + vi] = acc;
          } else {
            A.b[6 * hk + vi] = acc;
          }
        }
      }
      if (gt == 0) *A.gmax = 0.0;
      grid.sync();
      LBC_TICK(3);
      if (it == 0) {
        if (A.lambdaInit > 0) lambda = A.lambdaInit;
        else {
          double m = 0;   // max |diag| (order-free): doubles >= 0 order like their bit patterns
          for (int i = gt; i < A.nFree * 6; i += GS) m = fmax(m, fabs(A.Hpp[36 * (size_t)(i / 6) + (i % 6) * 7]));
          for (int i = gt; i < A.M * 3; i += GS) m = fmax(m, fabs(A.Hll[9 * (size_t)(i / 3) + (i % 3) * 4]));
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
          if (lane == 0 && m > 0) atomicMax(reinterpret_cast<unsigned long long*>(A.gmax), (unsigned long long)__double_as_longlong(m));
          grid.sync();
          lambda = 1e-50 * *A.gmax;
        }
        ni = 2;
        nBad = 0;
      }
      double rho = 0;
      int qmax = 0;
      bool stopped = false;
      do {
        // ---- push ----
        for (int k = gt; k < A.K; k += GS) A.poseBak[k] = A.pose[k];
        for (int i = gt; i < 3 * A.M; i += GS) A.ptBak[i] = A.pt[i];
        // ---- solve with Schur complement (block_solver.hpp:354-486) ----
        for (int m = gt; m < A.M; m += GS) {
          double Dm[9];
          for (int i = 0; i < 9; ++i) Dm[i] = A.Hll[9 * (size_t)m + i];
          Dm[0] += lambda; Dm[4] += lambda; Dm[8] += lambda;
          const double c00 = Dm[4] * Dm[8] - Dm[5] * Dm[7], c01 = Dm[5] * Dm[6] - Dm[3] * Dm[8], c02 = Dm[3] * Dm[7] - Dm[4] * Dm[6];
          const double det = Dm[0] * c00 + Dm[1] * c01 + Dm[2] * c02, id = 1.0 / det;
          double* Di = A.Dinv + 9 * (size_t)m;
          Di[0] = c00 * id; Di[1] = (Dm[2] * Dm[7] - Dm[1] * Dm[8]) * id; Di[2] = (Dm[1] * Dm[5] - Dm[2] * Dm[4]) * id;
          Di[3] = c01 * id; Di[4] = (Dm[0] * Dm[8] - Dm[2] * Dm[6]) * id; Di[5] = (Dm[2] * Dm[3] - Dm[0] * Dm[5]) * id;
          Di[6] = c02 * id; Di[7] = (Dm[1] * Dm[6] - Dm[0] * Dm[7]) * id; Di[8] = (Dm[0] * Dm[4] - Dm[1] * Dm[3]) * id;
          const double* bl = A.b + n + 3 * m;
          for (int i = 0; i < 3; ++i) A.db[3 * m + i] = Di[i * 3] * bl[0] + Di[i * 3 + 1] * bl[1] + Di[i * 3 + 2] * bl[2];
        }
        grid.sync();
        LBC_TICK(4);
        // BD[e] = B_e Dinv_m for edges of free poses
        for (int e = gt; e < A.E; e += GS) {
          if (A.hidx[A.ekf[e]] < 0) continue;
          const double* B = A.Hpl + 18 * (size_t)e;
          const double* Di = A.Dinv + 9 * (size_t)A.emp[e];
          double* BD = A.BD + 18 * (size_t)e;
          for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 3; ++j) BD[i * 3 + j] = B[i * 3] * Di[j] + B[i * 3 + 1] * Di[3 + j] + B[i * 3 + 2] * Di[6 + j];
        }
        grid.sync();
        LBC_TICK(5);
        // reduced system: one warp per (free pose i, free pose j >= i) block, lanes over pose i's edges
        const int nBlocks = A.nFree * (A.nFree + 1) / 2;
        // (one warp per ROW of a 6x6 block: the same per-lane sums and butterfly as lba_kernel, six short loops per block)
        for (int item = gw; item < nBlocks * 6; item += GW) {
          const int blk = item / 6, i = item - 6 * blk;
          int bi = 0, rem = blk;
          while (rem >= A.nFree - bi) { rem -= A.nFree - bi; ++bi; }
          const int bj = bi + rem;
          double acc[6] = {0, 0, 0, 0, 0, 0};
          const int* pr = A.pairE2 + A.blkOfs[blk] - A.kfOfs[bi];
#pragma unroll 2
          for (int q = A.kfOfs[bi] + lane; q < A.kfOfs[bi + 1]; q += 32) {
            const int e2 = pr[q];                       // the observation of the same point in pose bj (or -1), precomputed
            if (e2 < 0) continue;
            const int e1 = A.kfEdges[q];
            const double* BD = A.BD + 18 * (size_t)e1 + 3 * i;
            const double* B2 = A.Hpl + 18 * (size_t)e2;
            const double d0 = BD[0], d1 = BD[1], d2 = BD[2];
#pragma unroll
            for (int j = 0; j < 6; ++j) acc[j] += d0 * B2[j * 3] + d1 * B2[j * 3 + 1] + d2 * B2[j * 3 + 2];
          }
#pragma unroll
          for (int j = 0; j < 6; ++j)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
          if (lane == 0) {
            // lba_kernel writes (i,j) and its mirror for every (i,j) in row-major order; inside a diagonal block the later
            // write wins, i.e. both positions end up with the value of the pair whose FIRST index is the larger one
            const int jEnd = bi == bj ? i : 5;
            for (int j = 0; j <= jEnd; ++j) {
              double v = -acc[j];
              if (bi == bj) v += A.Hpp[36 * (size_t)bi + i * 6 + j] + (i == j ? lambda : 0.0);
              A.S[(size_t)(6 * bi + i) * n + 6 * bj + j] = v;
              A.S[(size_t)(6 * bj + j) * n + 6 * bi + i] = v;
            }
          }
        }
        // b_schur = b_p - sum_e B_e db_m, one warp per free pose (warps from the far end of the grid)
        for (int hk = GW - 1 - gw; hk < A.nFree; hk += GW) {
          double acc[6] = {0, 0, 0, 0, 0, 0};
          for (int q = A.kfOfs[hk] + lane; q < A.kfOfs[hk + 1]; q += 32) {
            const int e = A.kfEdges[q];
            const double* B = A.Hpl + 18 * (size_t)e;
            const double* db = A.db + 3 * (size_t)A.emp[e];
            for (int i = 0; i < 6; ++i) acc[i] += B[i * 3] * db[0] + B[i * 3 + 1] * db[1] + B[i * 3 + 2] * db[2];
          }
#pragma unroll
          for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
          if (lane == 0)
            for (int i = 0; i < 6; ++i) A.bs[6 * hk + i] = A.b[6 * hk + i] - acc[i];
        }
        for (int i = gt; i < NX; i += GS) A.x[i] = 0;
        grid.sync();
        LBC_TICK(6);
        // ---- CTA 0: dense LDL^T of S (lower, un-pivoted; stand-in for SimplicialLDLT) + substitutions ----
        if (blockIdx.x == 0) {
          const bool inSmem = n <= LBC_MAX_SMEM_N;
          double* Sm = inSmem ? s_S : A.S;
          if (inSmem) {
            for (int i = tid; i < n * n; i += LBC_NT) Sm[i] = A.S[i];
          }
          double* xs = inSmem ? (s_S + (size_t)n * n) : A.x;     // x_p lives next to the factor
          for (int i = tid; i < n; i += LBC_NT) xs[i] = A.bs[i];
          if (tid == 0) s_flag = 1;
          __syncthreads();
          // right-looking LDL^T; the forward substitution rides along (after column k is scaled, y_k is final and
          // y_i -= L[i][k] * y_k is the k-th term of y_i, in the same ascending order as a separate sweep)
          const int wid = tid >> 5;
          for (int k = 0; k < n; ++k) {
            const double d = Sm[(size_t)k * n + k];
            if (d == 0 || !isfinite(d)) { if (tid == 0) s_flag = 0; break; }
            for (int i = k + 1 + tid; i < n; i += LBC_NT) Sm[(size_t)i * n + k] /= d;
            __syncthreads();
            const double yk = xs[k];
            for (int i = k + 1 + wid; i < n; i += LBC_NT / 32) {      // one warp per row, lanes over the row's lower part
              const double lik = Sm[(size_t)i * n + k];
              for (int j = k + 1 + lane; j <= i; j += 32) Sm[(size_t)i * n + j] -= lik * d * Sm[(size_t)j * n + k];
              if (lane == 0) xs[i] -= lik * yk;
            }
            __syncthreads();
          }
          __syncthreads();
          const bool solvedLocal = s_flag != 0;
          if (tid == 0) A.gflag[2] = solvedLocal ? 1 : 0;
          if (!solvedLocal && !inSmem)
            for (int i = tid; i < n; i += LBC_NT) A.x[i] = 0;    // (the global x doubled as y: restore the "not solved" zeros)
          if (solvedLocal && n > 0) {
            for (int i = tid; i < n; i += LBC_NT) xs[i] /= Sm[(size_t)i * n + i];
            __syncthreads();
            for (int j = n - 1; j >= 0; --j) {
              const double yj = xs[j];
              for (int i = tid; i < j; i += LBC_NT) xs[i] -= Sm[(size_t)j * n + i] * yj;
              __syncthreads();
            }
            if (inSmem)
              for (int i = tid; i < n; i += LBC_NT) A.x[i] = xs[i];
          }
        }
        grid.sync();
        LBC_TICK(7);
        const bool solved = A.gflag[2] != 0;
        if (solved && n > 0) {
          // landmarks: x_l = Dinv (b_l - sum_e B_e^T x_p)
          for (int m = gt; m < A.M; m += GS) {
            double cl[3] = {A.b[n + 3 * m], A.b[n + 3 * m + 1], A.b[n + 3 * m + 2]};
            for (int q = A.ptOfs[m]; q < A.ptOfs[m + 1]; ++q) {
              const int e = A.ptEdges[q];
              const int hk = A.hidx[A.ekf[e]];
              const double* B = A.Hpl + 18 * (size_t)e;
              for (int j = 0; j < 3; ++j)
                for (int i = 0; i < 6; ++i) cl[j] -= B[i * 3 + j] * A.x[6 * hk + i];
            }
            const double* Di = A.Dinv + 9 * (size_t)m;
            for (int i = 0; i < 3; ++i) A.x[n + 3 * m + i] = Di[i * 3] * cl[0] + Di[i * 3 + 1] * cl[1] + Di[i * 3 + 2] * cl[2];
          }
        }
        grid.sync();
        LBC_TICK(8);
        // ---- update (oplus) ----
        for (int k = gt; k < A.K; k += GS)
          if (A.hidx[k] >= 0) A.pose[k] = se3_mul(se3_exp(A.x + 6 * A.hidx[k]), A.pose[k]);
        for (int i = gt; i < 3 * A.M; i += GS) A.pt[i] += A.x[n + i];
        grid.sync();
        LBC_TICK(9);
        double tempChi = lbc_errors_chi(grid, A, hM, hS);
        LBC_TICK(10);
        if (!solved) tempChi = 1.7976931348623157e308;
        rho = currentChi - tempChi;
        for (int j = gt; j < NX; j += GS) A.tmp[j] = A.x[j] * (lambda * A.x[j] + A.b[j]);
        const double scale = lbc_sum1024(grid, A.tmp, NX, A.part + 32) + 1e-3;
        LBC_TICK(11);
        rho /= scale;
        if (rho > 0 && isfinite(tempChi)) {
          const double tr = 2 * rho - 1;
          double alpha = 1. - tr * tr * tr;   // pow(2*rho-1, 3)
          alpha = fmin(alpha, 2. / 3.);
          const double scaleFactor = fmax(1. / 3., alpha);
          lambda *= scaleFactor;
          ni = 2;
          currentChi = tempChi;
        } else {
          lambda *= ni;
          ni *= 2;
          for (int k = gt; k < A.K; k += GS) A.pose[k] = A.poseBak[k];   // pop
          for (int i = gt; i < 3 * A.M; i += GS) A.pt[i] = A.ptBak[i];
        }
        ++qmax;
        stopped = lbc_stop(grid, A, stopSeq);       // (grid-wide barrier: the pop is complete)
        LBC_TICK(12);
        if (A.prof && gt == 0) A.prof[15] += 1;
      } while (rho < 0 && qmax < 100 && !stopped);
      ++cj;
      if (qmax == 100 || rho == 0) { ok = false; continue; }
      if ((iniChi - currentChi) * 1e3 < iniChi) ++nBad; else nBad = 0;
      if (nBad >= 3) ok = false;
    }
    if (gt == 0) A.iters[call] = cj;
  }
  grid.sync();
  // ---- final chi2 / depth test on the residuals of the last evaluated state (src/Optimizer.cc:2295-2352) ----
  for (int e = gt; e < A.E; e += GS) {
    const bool st = A.obs[3 * e + 2] >= 0;
    const double om = (double)A.invSigma2[e];
    const double* r = A.err + 3 * e;
    const double c = r[0] * om * r[0] + r[1] * om * r[1] + (st ? r[2] * om * r[2] : 0.0);
    double p[3];
    se3_map(A.pose[A.ekf[e]], A.pt + 3 * (size_t)A.emp[e], p);
    const bool isBad = c > (st ? 7.815 : 5.991) || !(p[2] > 0.0);
    A.edgeBad[e] = isBad ? 1 : 0;
    A.tmp[e] = isBad ? 1.0 : 0.0;
  }
  const double nBadEdges = lbc_sum1024(grid, A.tmp, A.E, A.part + 64);
  if (nBadEdges >= A.E * 0.5) {
    if (gt == 0) *A.status = 2;
    return;
  }
  for (int k = gt; k < A.K; k += GS)
    if (!A.kfFixed[k]) se3_to_Tcw(A.pose[k], A.kfT + 16 * k);
  for (int i = gt; i < 3 * A.M; i += GS) A.mpXyz[i] = (float)A.pt[i];
}

// internal (orbx_track.cu): P problems with fixed-capacity slices
int orbx_launch_pose_opt_slices(orbx_ctx* ctx, cudaStream_t st, int P, const int* d_start, const int* d_count, const float* d_xw,
                                const float* d_obs, const float* d_isg, const orbx_camera* cam, float* d_Tcw, uint8_t* d_outlier,
                                int* d_ninl, int* d_iters, double* d_scratch) {
  PoseOptArgs A;
  A.edgeOfs = nullptr;
  A.edgeStart = d_start;
  A.edgeCount = d_count;
  A.xw = d_xw;
  A.obs = d_obs;
  A.invSigma2 = d_isg;
  A.fx = cam->fx; A.fy = cam->fy; A.cx = cam->cx; A.cy = cam->cy; A.bf = cam->bf;
  A.Tcw = d_Tcw;
  A.outlier = d_outlier;
  A.nInliers = d_ninl;
  A.iters = d_iters;
  A.err = d_scratch;
  A.profile = g_po_profile;
  // Many-stream tracker: a problem's CTA holds 32 K registers for ~0.4 ms while most of its warps wait, and the extractor's
  // CTAs of the next step (other stream) only get what is left.  Both ways of shrinking that footprint were measured and
  // lost (profiles/r02x_pose_footprint.md): 128 threads per problem, and one problem per SM.
  pose_opt_kernel<256><<<P, 256, 0, st>>>(A);
  ORBX_LAUNCH(ctx);
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}

// =====================================================================================
// K20  PoseInertialOptimizationLastKeyFrame (src/Optimizer.cc:7665-8066, src/G2oTypes.cc:170-220,385-407,
//      496-520,730-812,995-1090): one CTA per problem.  15 unknowns (pose 6, velocity 3, gyro bias 3, acc bias 3),
//      Gauss-Newton 4 x 10 iterations, visual only-pose edges reduced with the canonical 256-way tree of K11, the
//      inertial / random-walk edges added by 81 threads with a fixed per-element summation order (k ascending), 15x15 pivoted
//      LDL^T in registers + shuffles (one warp), ImuCamPose::Update by one thread.  Conventions for the two points where
//      the reference is not reproducible (re-orthonormalisation, ExpSO3's float SVD): DESIGN.md §7.
// =====================================================================================
#define PIO_NT 288
struct PioArgs {
  int E;
  const float *xw, *obs, *invSigma2;
  const uint8_t* closePt;
  float fx, fy, cx, cy, bf;
  double Rcb[9], tcb[3], Rbc[9], tbc[3], Rcw0[9], tcw0[3];
  double state[21], kf[21];          // Rwb, twb, v, bg, ba
  double dR[9], dV[3], dP[3], dt;
  double infoI[81], infoG[9], infoA[9];
  int recInit;
  uint8_t* outlier;                  // [E]
  double* err;                       // [3E] scratch
  double* outState;                  // [21]
  double* H15;                       // [225]
  int* nRet;
  int* iters;                        // [4]
};

__device__ __forceinline__ void d_m3_mul(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
__device__ __forceinline__ void d_m3_t(const double* A, double* T) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) T[i * 3 + j] = A[j * 3 + i];
}
__device__ __forceinline__ void d_m3_v(const double* A, const double* v, double* o) {
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}
__device__ __forceinline__ void d_orthonormalize(double* R) {
  Quat q = quat_from_R(R);
  quat_normalize(q);
  quat_to_R(q, R);
}
__device__ void d_exp_so3(const double* w, double* R) {
  const double x = w[0], y = w[1], z = w[2];
  const double d2 = x * x + y * y + z * z, d = sqrt(d2);
  const double W[9] = {0.0, -z, y, z, 0.0, -x, -y, x, 0.0};
  double W2[9];
  d_m3_mul(W, W, W2);
  if (d < 1e-5) {
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + W[i] + 0.5 * W2[i];
  } else {
    const double a = sin(d) / d, b = (1.0 - cos(d)) / d2;
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + W[i] * a + W2[i] * b;
  }
  d_orthonormalize(R);
}
__device__ void d_log_so3(const double* R, double* w) {
  const double tr = R[0] + R[4] + R[8];
  w[0] = (R[7] - R[5]) / 2; w[1] = (R[2] - R[6]) / 2; w[2] = (R[3] - R[1]) / 2;
  const double costheta = (tr - 1.0) * 0.5;
  if (costheta > 1 || costheta < -1) return;
  const double theta = acos(costheta), sn = sin(theta);
  if (fabs(sn) < 1e-5) return;
#pragma unroll
  for (int i = 0; i < 3; ++i) w[i] = theta * w[i] / sn;
}
__device__ void d_inv_right_jac(const double* v, double* J) {
  const double x = v[0], y = v[1], z = v[2];
  const double d2 = x * x + y * y + z * z, d = sqrt(d2);
  const double W[9] = {0.0, -z, y, z, 0.0, -x, -y, x, 0.0};
  if (d < 1e-5) {
#pragma unroll
    for (int i = 0; i < 9; ++i) J[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  double W2[9];
  d_m3_mul(W, W, W2);
  const double c = 1.0 / d2 - (1.0 + cos(d)) / (2.0 * d * sin(d));
#pragma unroll
  for (int i = 0; i < 9; ++i) J[i] = ((i % 4 == 0) ? 1.0 : 0.0) + W[i] / 2 + W2[i] * c;
}

struct PioShared {
  double Rwb[9], twb[3], Rcw[9], tcw[3], v[3], bg[3], ba[3];
  double H[225], b[15], x[15], tot[36];
  double e9[9], J[81], OJ[81], Oe[9];
  double red[(PIO_NT / 32) * 36];
  int its, ok;
};

// residual of a visual edge at the shared camera pose
template <class AT, class ST>
__device__ __forceinline__ void pio_edge_error(const AT& A, const ST& S, int e, bool st, double* out, double* Xc) {
  const double X[3] = {(double)A.xw[3 * e], (double)A.xw[3 * e + 1], (double)A.xw[3 * e + 2]};
  d_m3_v(S.Rcw, X, Xc);
  Xc[0] += S.tcw[0]; Xc[1] += S.tcw[1]; Xc[2] += S.tcw[2];
  const double u = (double)A.fx * Xc[0] / Xc[2] + (double)A.cx, v = (double)A.fy * Xc[1] / Xc[2] + (double)A.cy;
  out[0] = (double)A.obs[3 * e] - u;
  out[1] = (double)A.obs[3 * e + 1] - v;
  out[2] = 0;
  if (st) { const double invZ = 1 / Xc[2]; out[2] = (double)A.obs[3 * e + 2] - (u - (double)A.bf * invZ); }
}
template <class AT>
__device__ __forceinline__ void pio_edge_jacobian(const AT& A, const double* Xc, bool st, double* J) {
  double Xb[3];
  d_m3_v(A.Rbc, Xc, Xb);
  Xb[0] += A.tbc[0]; Xb[1] += A.tbc[1]; Xb[2] += A.tbc[2];
  double pj[9] = {(double)A.fx / Xc[2], 0.0, -(double)A.fx * Xc[0] / (Xc[2] * Xc[2]),
                  0.0, (double)A.fy / Xc[2], -(double)A.fy * Xc[1] / (Xc[2] * Xc[2]), 0, 0, 0};
  if (st) {
    const double inv_z2 = 1.0 / (Xc[2] * Xc[2]);
    pj[6] = pj[0]; pj[7] = pj[1]; pj[8] = pj[2] + (double)A.bf * inv_z2;
  }
  double PR[9];
  d_m3_mul(pj, A.Rcb, PR);
  const double x_ = Xb[0], y_ = Xb[1], z_ = Xb[2];
  const double Sd[18] = {0.0, z_, -y_, 1.0, 0.0, 0.0, -z_, 0.0, x_, 0.0, 1.0, 0.0, y_, -x_, 0.0, 0.0, 0.0, 1.0};
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) J[r * 6 + c] = PR[r * 3] * Sd[c] + PR[r * 3 + 1] * Sd[6 + c] + PR[r * 3 + 2] * Sd[12 + c];
}
// EdgeInertial error + Jacobian (pose 2: columns 0-5, velocity 2: columns 6-8) by one thread into shared memory
__device__ void pio_inertial(const PioArgs& A, PioShared& S, bool withError) {
  const double* Rwb1 = A.kf;
  double Rbw1[9], dRt[9], T1[9], eR[9], er[3], invJr[9], RR[9];
  d_m3_t(Rwb1, Rbw1);
  d_m3_t(A.dR, dRt);
  d_m3_mul(dRt, Rbw1, T1);
  d_m3_mul(T1, S.Rwb, eR);
  d_log_so3(eR, er);
  if (withError) {
    double a[3], c[3];
    S.e9[0] = er[0]; S.e9[1] = er[1]; S.e9[2] = er[2];
    for (int i = 0; i < 3; ++i) a[i] = S.v[i] - A.kf[12 + i] - (i == 2 ? -9.81 : 0.0) * A.dt;
    d_m3_v(Rbw1, a, c);
    for (int i = 0; i < 3; ++i) S.e9[3 + i] = c[i] - A.dV[i];
    for (int i = 0; i < 3; ++i) a[i] = S.twb[i] - A.kf[9 + i] - A.kf[12 + i] * A.dt - (i == 2 ? -9.81 : 0.0) * A.dt * A.dt / 2;
    d_m3_v(Rbw1, a, c);
    for (int i = 0; i < 3; ++i) S.e9[6 + i] = c[i] - A.dP[i];
  }
  d_inv_right_jac(er, invJr);
  d_m3_mul(Rbw1, S.Rwb, RR);
  for (int i = 0; i < 81; ++i) S.J[i] = 0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      S.J[r * 9 + c] = invJr[r * 3 + c];
      S.J[(6 + r) * 9 + 3 + c] = RR[r * 3 + c];
      S.J[(3 + r) * 9 + 6 + c] = Rbw1[r * 3 + c];
    }
}

__global__ void __launch_bounds__(PIO_NT, 2) pose_inertial_kernel(const PioArgs* __restrict__ args) {
  const PioArgs& A = args[blockIdx.x];
  __shared__ PioShared S;
  const int tid = threadIdx.x, E = A.E;
  const bool edgeThread = tid < 256;     // warps 0-7: the visual edges; warp 8 linearises the inertial edge beside them
  const float* isg = A.invSigma2;
  uint8_t* outlier = A.outlier;
  double* err = A.err;
  if (tid < 9) { S.Rwb[tid] = A.state[tid]; S.Rcw[tid] = A.Rcw0[tid]; }
  if (tid < 3) { S.twb[tid] = A.state[9 + tid]; S.v[tid] = A.state[12 + tid]; S.bg[tid] = A.state[15 + tid]; S.ba[tid] = A.state[18 + tid]; S.tcw[tid] = A.tcw0[tid]; }
  if (tid < 15) S.x[tid] = 0;
  if (tid == 0) S.its = 0;
  if (tid < 4) A.iters[tid] = 0;
  for (int e = edgeThread ? tid : E; e < E; e += 256) outlier[e] = 0;
  __syncthreads();
  const HuberD hMono = huber_make(sqrtf(5.991f)), hStereo = huber_make(sqrtf(7.815f));
  const float chi2Mono[4] = {12.f, 7.5f, 5.991f, 5.991f}, chi2Stereo[4] = {15.6f, 9.8f, 7.815f, 7.815f};
  bool robust = true;
  int nBad = 0, nInliers = 0;
  for (int round = 0; round < 4; ++round) {
    int cj = 0;
    bool ok = true;
    for (int it = 0; it < 10 && ok; ++it) {
      // ---- computeActiveErrors + buildSystem of the visual edges (canonical 256-way tree) ----
      double acc[27];
#pragma unroll
      for (int k = 0; k < 27; ++k) acc[k] = 0;
      if (tid == 256) pio_inertial(A, S, true);
      for (int e = edgeThread ? tid : E; e < E; e += 256) {
        if (outlier[e]) continue;
        const bool st = A.obs[3 * e + 2] >= 0;
        double r[3], Xc[3], J[18];
        pio_edge_error(A, S, e, st, r, Xc);
        err[3 * e] = r[0]; err[3 * e + 1] = r[1]; err[3 * e + 2] = r[2];
        pio_edge_jacobian(A, Xc, st, J);
        const int D = st ? 3 : 2;
        const double om = (double)isg[e];
        double w = 1.0;
        if (robust) huber_rho(st ? hStereo : hMono, r[0] * om * r[0] + r[1] * om * r[1] + (st ? r[2] * om * r[2] : 0.0), w);
        int idx = 0;
        for (int i = 0; i < 6; ++i) {
          double sg = 0;
          for (int d = 0; d < D; ++d) sg += J[d * 6 + i] * om * r[d];
          acc[21 + i] -= w * sg;
          for (int j = i; j < 6; ++j) {
            double a = 0;
            for (int d = 0; d < D; ++d) a += J[d * 6 + i] * (w * om) * J[d * 6 + j];
            acc[idx++] += a;
          }
        }
      }
      block_partials<27>(acc, S.red);
      if (tid < 27) {
        double s = 0;
        for (int w = 0; w < 8; ++w) s += S.red[w * 27 + tid];
        S.tot[tid] = s;
      }
      __syncthreads();
      // ---- assemble H (15x15), b ----
      if (tid < 225) S.H[tid] = 0;
      if (tid < 15) S.b[tid] = 0;
      __syncthreads();
      if (tid < 36) {
        const int i = tid / 6, j = tid - 6 * i, lo = min(i, j), hi = max(i, j);
        S.H[i * 15 + j] = S.tot[6 * lo - (lo * (lo - 1)) / 2 + (hi - lo)];
      } else if (tid < 42) {
        S.b[tid - 36] = S.tot[21 + tid - 36];
      }
      if (tid >= 64 && tid < 64 + 81) {             // Omega * J, Omega * e (per element, k ascending)
        const int r = (tid - 64) / 9, c = (tid - 64) - 9 * r;
        double t = 0;
        for (int k = 0; k < 9; ++k) t += A.infoI[r * 9 + k] * S.J[k * 9 + c];
        S.OJ[r * 9 + c] = t;
        if (c == 0) {
          double s2 = 0;
          for (int k = 0; k < 9; ++k) s2 += A.infoI[r * 9 + k] * S.e9[k];
          S.Oe[r] = s2;
        }
      }
      __syncthreads();
      if (tid < 81) {                               // H += J^T (Omega J), b -= J^T (Omega e)
        const int i = tid / 9, j = tid - 9 * i;
        double t = 0;
        for (int k = 0; k < 9; ++k) t += S.J[k * 9 + i] * S.OJ[k * 9 + j];
        S.H[i * 15 + j] += t;
        if (j == 0) {
          double s2 = 0;
          for (int k = 0; k < 9; ++k) s2 += S.J[k * 9 + i] * S.Oe[k];
          S.b[i] -= s2;
        }
      } else if (tid >= 96 && tid < 99) {           // random-walk edges: error = bias - bias_kf, Jacobian I
        const int i = tid - 96;
        double sg = 0, sa = 0;
        for (int k = 0; k < 3; ++k) { sg += A.infoG[i * 3 + k] * (S.bg[k] - A.kf[15 + k]); sa += A.infoA[i * 3 + k] * (S.ba[k] - A.kf[18 + k]); }
        S.b[9 + i] -= sg;
        S.b[12 + i] -= sa;
        for (int j = 0; j < 3; ++j) { S.H[(9 + i) * 15 + 9 + j] += A.infoG[i * 3 + j]; S.H[(12 + i) * 15 + 12 + j] += A.infoA[i * 3 + j]; }
      }
      __syncthreads();
      // ---- solve (LinearSolverDense: pivoted LDL^T; a failure leaves x as it was) + update ----
      if (tid < 32) {
        const bool okSolve = ldlt_solve_smem<15, 15>(S.H, S.b, S.x);
        if (tid == 0) {
          S.ok = okSolve ? 1 : 0;
          double d[3], E3[9], Rn[9];
          d_m3_v(S.Rwb, S.x + 3, d);
          for (int i = 0; i < 3; ++i) S.twb[i] += d[i];
          d_exp_so3(S.x, E3);
          d_m3_mul(S.Rwb, E3, Rn);
          for (int i = 0; i < 9; ++i) S.Rwb[i] = Rn[i];
          if (++S.its >= 3) { d_orthonormalize(S.Rwb); S.its = 0; }
          double Rbw[9], tbw[3];
          d_m3_t(S.Rwb, Rbw);
          d_m3_v(Rbw, S.twb, tbw);
          for (int i = 0; i < 3; ++i) tbw[i] = -tbw[i];
          d_m3_mul(A.Rcb, Rbw, S.Rcw);
          d_m3_v(A.Rcb, tbw, S.tcw);
          for (int i = 0; i < 3; ++i) { S.tcw[i] += A.tcb[i]; S.v[i] += S.x[6 + i]; S.bg[i] += S.x[9 + i]; S.ba[i] += S.x[12 + i]; }
        }
      }
      __syncthreads();
      ok = S.ok != 0;
      ++cj;
    }
    if (tid == 0) A.iters[round] = cj;
    // ---- chi2 classification (src/Optimizer.cc:7900-7975) ----
    const float chi2close = 1.5f * chi2Mono[round];
    double cnt[2] = {0, 0};
    for (int e = edgeThread ? tid : E; e < E; e += 256) {
      const bool st = A.obs[3 * e + 2] >= 0;
      if (outlier[e]) {
        double r[3], Xc[3];
        pio_edge_error(A, S, e, st, r, Xc);
        err[3 * e] = r[0]; err[3 * e + 1] = r[1]; err[3 * e + 2] = r[2];
      }
      const double om = (double)isg[e];
      const double* r = err + 3 * e;
      const float chi2 = (float)(r[0] * om * r[0] + r[1] * om * r[1] + (st ? r[2] * om * r[2] : 0.0));
      bool bad;
      if (!st) {
        const bool bClose = A.closePt[e] != 0;
        const bool depthPos = (S.Rcw[6] * (double)A.xw[3 * e] + S.Rcw[7] * (double)A.xw[3 * e + 1] + S.Rcw[8] * (double)A.xw[3 * e + 2] + S.tcw[2]) > 0.0;
        bad = (chi2 > chi2Mono[round] && !bClose) || (bClose && chi2 > chi2close) || !depthPos;
      } else {
        bad = chi2 > chi2Stereo[round];
      }
      outlier[e] = bad ? 1 : 0;
      cnt[0] += bad ? 1.0 : 0.0;
      cnt[1] += bad ? 0.0 : 1.0;
    }
    block_sum<2>(cnt, S.red);
    nBad = (int)cnt[0];
    nInliers = (int)cnt[1];
    if (round == 2) robust = false;
    if (E + 3 < 10) break;
  }
  __syncthreads();
  if (nInliers < 30 && !A.recInit) {               // recovery (:7990-8020)
    double cnt[1] = {0};
    for (int e = edgeThread ? tid : E; e < E; e += 256) {
      const bool st = A.obs[3 * e + 2] >= 0;
      double r[3], Xc[3];
      pio_edge_error(A, S, e, st, r, Xc);
      err[3 * e] = r[0]; err[3 * e + 1] = r[1]; err[3 * e + 2] = r[2];
      const double om = (double)isg[e];
      const double c2 = r[0] * om * r[0] + r[1] * om * r[1] + (st ? r[2] * om * r[2] : 0.0);
      if (c2 < (double)(st ? 24.f : 18.f)) outlier[e] = 0; else cnt[0] += 1.0;
    }
    block_sum<1>(cnt, S.red);
    nBad = (int)cnt[0];
  }
  __syncthreads();
  // ---- outputs: state, prior Hessian ----
  if (tid < 9) A.outState[tid] = S.Rwb[tid];
  if (tid < 3) { A.outState[9 + tid] = S.twb[tid]; A.outState[12 + tid] = S.v[tid]; A.outState[15 + tid] = S.bg[tid]; A.outState[18 + tid] = S.ba[tid]; }
  if (tid == 0) *A.nRet = E - nBad;
  double a36[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) a36[k] = 0;
  if (tid == 256) pio_inertial(A, S, false);
  for (int e = edgeThread ? tid : E; e < E; e += 256) {
    if (outlier[e]) continue;
    const bool st = A.obs[3 * e + 2] >= 0;
    double r[3], Xc[3], J[18];
    pio_edge_error(A, S, e, st, r, Xc);
    pio_edge_jacobian(A, Xc, st, J);
    const int D = st ? 3 : 2;
    const double om = (double)isg[e];
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) {
        double t = 0;
        for (int d = 0; d < D; ++d) t += J[d * 6 + i] * om * J[d * 6 + j];
        a36[i * 6 + j] += t;
      }
  }
  block_partials<36>(a36, S.red);
  if (tid < 36) {
    double s = 0;
    for (int w = 0; w < 8; ++w) s += S.red[w * 36 + tid];
    S.tot[tid] = s;
  }
  __syncthreads();
  if (tid < 225) S.H[tid] = 0;
  __syncthreads();
  if (tid >= 64 && tid < 64 + 81) {
    const int r = (tid - 64) / 9, c = (tid - 64) - 9 * r;
    double t = 0;
    for (int k = 0; k < 9; ++k) t += A.infoI[r * 9 + k] * S.J[k * 9 + c];
    S.OJ[r * 9 + c] = t;
  }
  __syncthreads();
  if (tid < 81) {
    const int i = tid / 9, j = tid - 9 * i;
    double t = 0;
    for (int k = 0; k < 9; ++k) t += S.J[k * 9 + i] * S.OJ[k * 9 + j];
    S.H[i * 15 + j] += t;
  } else if (tid >= 96 && tid < 105) {
    const int i = (tid - 96) / 3, j = (tid - 96) - 3 * i;
    S.H[(9 + i) * 15 + 9 + j] += A.infoG[i * 3 + j];
    S.H[(12 + i) * 15 + 12 + j] += A.infoA[i * 3 + j];
  }
  __syncthreads();
  if (tid < 36) { const int i = tid / 6, j = tid - 6 * i; S.H[i * 15 + j] += S.tot[tid]; }
  __syncthreads();
  if (tid < 225) A.H15[tid] = S.H[tid];
}

// =====================================================================================
// K21  PoseInertialOptimizationLastFrame (second function; src/Optimizer.cc:8068-8603, EdgeInertial with all
//      six vertices free src/G2oTypes.cc:730-812, EdgePriorPoseImu :941-981, bias-corrected deltas src/ImuTypes.cc:367-394,
//      Optimizer::Marginalize :5366-5450): one CTA per problem, 30 unknowns (g2o's vertex-id order: frame 0-14, previous
//      frame 15-29).  Warps 0-7 own the visual edges (the canonical 256-way tree), warp 8 linearises the inertial edge and
//      warp 9 the prior edge at the same time; the 30x30 system is assembled edge by edge in a fixed order, solved by one
//      warp (pivoted LDL^T, lane i = row i) and applied to the two body states by two threads.  The final 30x30 Hessian is
//      reduced to the frame's 15x15 prior with a warp-parallel cyclic Jacobi eigen-solver (pseudo-inverse, 1e-6 threshold).
// =====================================================================================
#define PLF_NT 320
#define PLF_LD 31     // leading dimension of the 30x30 system in shared memory (odd: conflict-free columns)
struct PlfArgs {
  int E;
  const float *xw, *obs, *invSigma2;
  const uint8_t* closePt;
  float fx, fy, cx, cy, bf;
  double Rcb[9], tcb[3], Rbc[9], tbc[3], Rcw0[9], tcw0[3];
  double state[21], prev[21];        // Rwb, twb, v, bg, ba
  double dR0[9], dV0[3], dP0[3], dt, JRg[9], JVg[9], JVa[9], JPg[9], JPa[9], bpre[6];
  double infoI[81], infoG[9], infoA[9];
  double prior[21], Hp[225];
  int recInit;
  uint8_t* outlier;
  double* err;
  double* outState;
  double* H15;
  int* nRet;
  int* iters;
  unsigned long long* prof;   // [8] optional per-phase nanosecond totals (ORBX_PIO_PROFILE=1), else null
};
#define PLF_TICK(n) do { if (A.prof && tid == 0) { const unsigned long long t1_ = lbc_now(); A.prof[n] += t1_ - t0; t0 = t1_; } } while (0)
struct PlfShared {
  double cur[21], prev[21], Rcw[9], tcw[3];
  double H[30 * PLF_LD], b[30], x[30], tot[36];
  double e9[9], J[216], OJ[216], Oe[9];
  double e15[15], Jp[225], OJp[225], Oep[15], wPrior;
  double mA[225], mV[225], mInv[225], mT[225];
  double red[(PLF_NT / 32) * 36];
  int itsCur, itsPrev, ok;
};

__device__ void d_right_jac(const double* v, double* J) {
  const double x = v[0], y = v[1], z = v[2];
  const double d2 = x * x + y * y + z * z, d = sqrt(d2);
  const double W[9] = {0.0, -z, y, z, 0.0, -x, -y, x, 0.0};
  if (d < 1e-5) {
#pragma unroll
    for (int i = 0; i < 9; ++i) J[i] = (i % 4 == 0) ? 1.0 : 0.0;
    return;
  }
  double W2[9];
  d_m3_mul(W, W, W2);
  const double a = (1.0 - cos(d)) / d2, b = (d - sin(d)) / (d2 * d);
#pragma unroll
  for (int i = 0; i < 9; ++i) J[i] = ((i % 4 == 0) ? 1.0 : 0.0) - W[i] * a + W2[i] * b;
}
__device__ __forceinline__ void d_skew(const double* v, double* S) {
  S[0] = 0; S[1] = -v[2]; S[2] = v[1]; S[3] = v[2]; S[4] = 0; S[5] = -v[0]; S[6] = -v[1]; S[7] = v[0]; S[8] = 0;
}
// body state (Rwb 0-8, twb 9-11, v 12-14, bg 15-17, ba 18-20) (+)= x[15]: ImuCamPose::Update + plain additions
__device__ void d_body_update(double* B, int& its, const double* x) {
  double d[3], E3[9], Rn[9];
  d_m3_v(B, x + 3, d);
  for (int i = 0; i < 3; ++i) B[9 + i] += d[i];
  d_exp_so3(x, E3);
  d_m3_mul(B, E3, Rn);
  for (int i = 0; i < 9; ++i) B[i] = Rn[i];
  if (++its >= 3) { d_orthonormalize(B); its = 0; }
  for (int i = 0; i < 9; ++i) B[12 + i] += x[6 + i];
}
// EdgeInertial::computeError + linearizeOplus (one thread): S.e9, S.J[9][24] in the edge's vertex order
__device__ void plf_inertial(const PlfArgs& A, PlfShared& S) {
  const double* cur = S.cur;
  const double* prev = S.prev;
  double dbg[3], dba[3], w[3], Ew[9], dRc[9], dV[3], dP[3], t1[3], t2[3];
  for (int i = 0; i < 3; ++i) { dbg[i] = prev[15 + i] - A.bpre[i]; dba[i] = prev[18 + i] - A.bpre[3 + i]; }
  d_m3_v(A.JRg, dbg, w);
  d_exp_so3(w, Ew);
  d_m3_mul(A.dR0, Ew, dRc);
  d_orthonormalize(dRc);
  d_m3_v(A.JVg, dbg, t1); d_m3_v(A.JVa, dba, t2);
  for (int i = 0; i < 3; ++i) dV[i] = A.dV0[i] + t1[i] + t2[i];
  d_m3_v(A.JPg, dbg, t1); d_m3_v(A.JPa, dba, t2);
  for (int i = 0; i < 3; ++i) dP[i] = A.dP0[i] + t1[i] + t2[i];
  const double dt = A.dt;
  double Rbw1[9], dRt[9], T1[9], eR[9], er[3], a[3], cv[3], cp[3];
  d_m3_t(prev, Rbw1);
  d_m3_t(dRc, dRt);
  d_m3_mul(dRt, Rbw1, T1);
  d_m3_mul(T1, cur, eR);
  d_log_so3(eR, er);
  for (int i = 0; i < 3; ++i) a[i] = cur[12 + i] - prev[12 + i] - (i == 2 ? -9.81 : 0.0) * dt;
  d_m3_v(Rbw1, a, cv);
  for (int i = 0; i < 3; ++i) a[i] = cur[9 + i] - prev[9 + i] - prev[12 + i] * dt - (i == 2 ? -9.81 : 0.0) * dt * dt / 2;
  d_m3_v(Rbw1, a, cp);
  for (int i = 0; i < 3; ++i) { S.e9[i] = er[i]; S.e9[3 + i] = cv[i] - dV[i]; S.e9[6 + i] = cp[i] - dP[i]; }
  double invJr[9], Rbw2[9], M1[9], M2[9], Sv[9], Sp[9], eRt[9], RJ[9], RR[9], M3[9], M4[9], cp2[3];
  d_inv_right_jac(er, invJr);
  d_m3_t(cur, Rbw2);
  d_m3_mul(invJr, Rbw2, M1);
  d_m3_mul(M1, prev, M2);
  d_skew(cv, Sv);
  for (int i = 0; i < 3; ++i) a[i] = cur[9 + i] - prev[9 + i] - prev[12 + i] * dt - 0.5 * (i == 2 ? -9.81 : 0.0) * dt * dt;
  d_m3_v(Rbw1, a, cp2);
  d_skew(cp2, Sp);
  d_m3_t(eR, eRt);
  d_right_jac(w, RJ);
  d_m3_mul(invJr, eRt, M1);
  d_m3_mul(M1, RJ, M3);
  d_m3_mul(M3, A.JRg, M4);
  d_m3_mul(Rbw1, cur, RR);
  double* J = S.J;
  for (int i = 0; i < 216; ++i) J[i] = 0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      const int k = r * 3 + c;
      J[r * 24 + c] = -M2[k];
      J[(3 + r) * 24 + c] = Sv[k];
      J[(6 + r) * 24 + c] = Sp[k];
      J[(6 + r) * 24 + 3 + c] = r == c ? -1.0 : 0.0;
      J[(3 + r) * 24 + 6 + c] = -Rbw1[k];
      J[(6 + r) * 24 + 6 + c] = -Rbw1[k] * dt;
      J[r * 24 + 9 + c] = -M4[k];
      J[(3 + r) * 24 + 9 + c] = -A.JVg[k];
      J[(6 + r) * 24 + 9 + c] = -A.JPg[k];
      J[(3 + r) * 24 + 12 + c] = -A.JVa[k];
      J[(6 + r) * 24 + 12 + c] = -A.JPa[k];
      J[r * 24 + 15 + c] = invJr[k];
      J[(6 + r) * 24 + 18 + c] = RR[k];
      J[(3 + r) * 24 + 21 + c] = Rbw1[k];
    }
}
// EdgePriorPoseImu::computeError + linearizeOplus + Huber weight (one thread): S.e15, S.Jp, S.wPrior
__device__ void plf_prior(const PlfArgs& A, PlfShared& S, bool robust) {
  const double* prev = S.prev;
  double pRt[9], eR[9], er[3], d[3], et[3], invJr[9];
  d_m3_t(A.prior, pRt);
  d_m3_mul(pRt, prev, eR);
  d_log_so3(eR, er);
  for (int i = 0; i < 3; ++i) d[i] = prev[9 + i] - A.prior[9 + i];
  d_m3_v(pRt, d, et);
  for (int i = 0; i < 3; ++i) {
    S.e15[i] = er[i]; S.e15[3 + i] = et[i]; S.e15[6 + i] = prev[12 + i] - A.prior[12 + i];
    S.e15[9 + i] = prev[15 + i] - A.prior[15 + i]; S.e15[12 + i] = prev[18 + i] - A.prior[18 + i];
  }
  d_inv_right_jac(er, invJr);
  for (int i = 0; i < 225; ++i) S.Jp[i] = 0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) { S.Jp[r * 15 + c] = invJr[r * 3 + c]; S.Jp[(3 + r) * 15 + 3 + c] = eR[r * 3 + c]; }
  for (int i = 6; i < 15; ++i) S.Jp[i * 15 + i] = 1.0;
  double w = 1.0;
  if (robust) {
    double chi = 0;
    for (int r = 0; r < 15; ++r) {
      double s = 0;
      for (int k = 0; k < 15; ++k) s += A.Hp[r * 15 + k] * S.e15[k];
      chi += S.e15[r] * s;
    }
    huber_rho(huber_make(5.0f), chi, w);
  }
  S.wPrior = w;
}
// Assembly of the inertial, random-walk and prior edges into S.H (ld PLF_LD) / S.b, edge by edge in a fixed order.  `fin`:
// the reference's final ordering (previous frame first), no gradient, no robust weight.
__device__ void plf_assemble(const PlfArgs& A, PlfShared& S, bool fin) {
  const int tid = threadIdx.x;
  const int oI = fin ? 0 : 15;    // inertial local column c -> (c < 15 ? oI + c : oC + c - 15)
  const int oC = fin ? 15 : 0;
  for (int t = tid; t < 216 + 9 + 225 + 15; t += PLF_NT) {
    if (t < 216) {
      const int r = t / 24, c = t - 24 * r;
      double v = 0;
      for (int k = 0; k < 9; ++k) v += (1.0 * A.infoI[r * 9 + k]) * S.J[k * 24 + c];
      S.OJ[t] = v;
    } else if (t < 225) {
      const int r = t - 216;
      double v = 0;
      for (int k = 0; k < 9; ++k) v += (1.0 * A.infoI[r * 9 + k]) * S.e9[k];
      S.Oe[r] = v;
    } else if (t < 450) {
      const int u = t - 225, r = u / 15, c = u - 15 * r;
      const double w = S.wPrior;
      double v = 0;
      for (int k = 0; k < 15; ++k) v += (w * A.Hp[r * 15 + k]) * S.Jp[k * 15 + c];
      S.OJp[u] = v;
    } else {
      const int r = t - 450;
      const double w = S.wPrior;
      double v = 0;
      for (int k = 0; k < 15; ++k) v += (w * A.Hp[r * 15 + k]) * S.e15[k];
      S.Oep[r] = v;
    }
  }
  __syncthreads();
  // EdgeInertial
  for (int t = tid; t < 576 + 24; t += PLF_NT) {
    if (t < 576) {
      const int i = t / 24, j = t - 24 * i;
      double v = 0;
      for (int k = 0; k < 9; ++k) v += S.J[k * 24 + i] * S.OJ[k * 24 + j];
      const int gi = i < 15 ? oI + i : oC + i - 15, gj = j < 15 ? oI + j : oC + j - 15;
      S.H[gi * PLF_LD + gj] += v;
    } else if (!fin) {
      const int i = t - 576;
      double v = 0;
      for (int k = 0; k < 9; ++k) v += S.J[k * 24 + i] * S.Oe[k];
      S.b[i < 15 ? oI + i : oC + i - 15] -= v;
    }
  }
  __syncthreads();
  // EdgeGyroRW / EdgeAccRW: e = bias(frame) - bias(previous), Jacobians (-I, +I)
  if (tid < 9) {
    const int i = tid / 3, j = tid - 3 * i;
    const double gI = A.infoG[tid], aI = A.infoA[tid];
    const int c = oC, p = oI;
    S.H[(c + 9 + i) * PLF_LD + c + 9 + j] += gI;  S.H[(p + 9 + i) * PLF_LD + p + 9 + j] += gI;
    S.H[(c + 9 + i) * PLF_LD + p + 9 + j] -= gI;  S.H[(p + 9 + i) * PLF_LD + c + 9 + j] -= gI;
    S.H[(c + 12 + i) * PLF_LD + c + 12 + j] += aI; S.H[(p + 12 + i) * PLF_LD + p + 12 + j] += aI;
    S.H[(c + 12 + i) * PLF_LD + p + 12 + j] -= aI; S.H[(p + 12 + i) * PLF_LD + c + 12 + j] -= aI;
  } else if (tid >= 32 && tid < 35 && !fin) {
    const int i = tid - 32;
    double sg = 0, sa = 0;
    for (int k = 0; k < 3; ++k) {
      sg += A.infoG[i * 3 + k] * (S.cur[15 + k] - S.prev[15 + k]);
      sa += A.infoA[i * 3 + k] * (S.cur[18 + k] - S.prev[18 + k]);
    }
    S.b[9 + i] -= sg;  S.b[24 + i] += sg;
    S.b[12 + i] -= sa; S.b[27 + i] += sa;
  }
  __syncthreads();
  // EdgePriorPoseImu
  for (int t = tid; t < 225 + 15; t += PLF_NT) {
    if (t < 225) {
      const int i = t / 15, j = t - 15 * i;
      double v = 0;
      for (int k = 0; k < 15; ++k) v += S.Jp[k * 15 + i] * S.OJp[k * 15 + j];
      S.H[(oI + i) * PLF_LD + oI + j] += v;
    } else if (!fin) {
      const int i = t - 225;
      double v = 0;
      for (int k = 0; k < 15; ++k) v += S.Jp[k * 15 + i] * S.Oep[k];
      S.b[oI + i] -= v;
    }
  }
  __syncthreads();
}
// warp 0: cyclic Jacobi on S.mA (15x15, symmetric) -> eigenvalues on the diagonal, eigenvectors in S.mV's columns
__device__ void plf_jacobi15(PlfShared& S) {
  const int k = threadIdx.x;   // lane
  double* Am = S.mA;
  double* V = S.mV;
  if (k < 15) for (int j = 0; j < 15; ++j) V[k * 15 + j] = k == j ? 1.0 : 0.0;
  __syncwarp();
  for (int s = 0; s < 16; ++s) {
    int nrot = 0;
    for (int p = 0; p < 14; ++p)
      for (int q = p + 1; q < 15; ++q) {
        const double apq = Am[p * 15 + q], app = Am[p * 15 + p], aqq = Am[q * 15 + q];
        if (fabs(apq) <= 1e-20 * (fabs(app) + fabs(aqq))) continue;   // (uniform across the warp)
        ++nrot;
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        __syncwarp();
        if (k < 15) {
          const double akp = Am[k * 15 + p], akq = Am[k * 15 + q];
          Am[k * 15 + p] = c * akp - sn * akq;
          Am[k * 15 + q] = sn * akp + c * akq;
        }
        __syncwarp();
        if (k < 15) {
          const double apk = Am[p * 15 + k], aqk = Am[q * 15 + k];
          Am[p * 15 + k] = c * apk - sn * aqk;
          Am[q * 15 + k] = sn * apk + c * aqk;
          const double vkp = V[k * 15 + p], vkq = V[k * 15 + q];
          V[k * 15 + p] = c * vkp - sn * vkq;
          V[k * 15 + q] = sn * vkp + c * vkq;
        }
        __syncwarp();
      }
    if (!nrot) break;
  }
}

__global__ void __launch_bounds__(PLF_NT, 2) pose_inertial_lf_kernel(const PlfArgs* __restrict__ args) {
  const PlfArgs& A = args[blockIdx.x];
  __shared__ PlfShared S;
  const int tid = threadIdx.x, E = A.E;
  const bool edgeThread = tid < 256;
  const float* isg = A.invSigma2;
  uint8_t* outlier = A.outlier;
  double* err = A.err;
  if (tid < 21) { S.cur[tid] = A.state[tid]; S.prev[tid] = A.prev[tid]; }
  if (tid < 9) S.Rcw[tid] = A.Rcw0[tid];
  if (tid < 3) S.tcw[tid] = A.tcw0[tid];
  if (tid < 30) S.x[tid] = 0;
  if (tid == 0) { S.itsCur = 0; S.itsPrev = 0; }
  if (tid < 4) A.iters[tid] = 0;
  for (int e = tid; e < E; e += PLF_NT) outlier[e] = 0;
  __syncthreads();
  const HuberD hMono = huber_make(sqrtf(5.991f)), hStereo = huber_make(sqrtf(7.815f));
  const float chi2Mono[4] = {5.991f, 5.991f, 5.991f, 5.991f}, chi2Stereo[4] = {15.6f, 9.8f, 7.815f, 7.815f};
  bool robust = true;
  int nBad = 0, nInliers = 0;
  unsigned long long t0 = A.prof ? lbc_now() : 0ull;
  for (int round = 0; round < 4; ++round) {
    int cj = 0;
    bool ok = true;
    for (int it = 0; it < 10 && ok; ++it) {
      double acc[27];
#pragma unroll
      for (int k = 0; k < 27; ++k) acc[k] = 0;
      if (tid == 256) plf_inertial(A, S);
      if (tid == 288) plf_prior(A, S, true);
      if (edgeThread)
        for (int e = tid; e < E; e += 256) {
          if (outlier[e]) continue;
          const bool st = A.obs[3 * e + 2] >= 0;
          double r[3], Xc[3], J[18];
          pio_edge_error(A, S, e, st, r, Xc);
          err[3 * e] = r[0]; err[3 * e + 1] = r[1]; err[3 * e + 2] = r[2];
          pio_edge_jacobian(A, Xc, st, J);
          const int D = st ? 3 : 2;
          const double om = (double)isg[e];
          double w = 1.0;
          if (robust) huber_rho(st ? hStereo : hMono, r[0] * om * r[0] + r[1] * om * r[1] + (st ? r[2] * om * r[2] : 0.0), w);
          int idx = 0;
          for (int i = 0; i < 6; ++i) {
            double sg = 0;
            for (int d = 0; d < D; ++d) sg += J[d * 6 + i] * om * r[d];
            acc[21 + i] -= w * sg;
            for (int j = i; j < 6; ++j) {
              double a = 0;
              for (int d = 0; d < D; ++d) a += J[d * 6 + i] * (w * om) * J[d * 6 + j];
              acc[idx++] += a;
            }
          }
        }
      PLF_TICK(0);
      block_partials<27>(acc, S.red);
      PLF_TICK(1);
      if (tid < 27) {
        double s = 0;
        for (int w = 0; w < 8; ++w) s += S.red[w * 27 + tid];
        S.tot[tid] = s;
      }
      for (int t = tid; t < 30 * PLF_LD; t += PLF_NT) S.H[t] = 0;
      if (tid >= 288 && tid < 318) S.b[tid - 288] = 0;
      __syncthreads();
      if (tid < 36) {
        const int i = tid / 6, j = tid - 6 * i, lo = min(i, j), hi = max(i, j);
        S.H[i * PLF_LD + j] = S.tot[6 * lo - (lo * (lo - 1)) / 2 + (hi - lo)];
      } else if (tid < 42) {
        S.b[tid - 36] = S.tot[21 + tid - 36];
      }
      __syncthreads();
      PLF_TICK(2);
      plf_assemble(A, S, false);
      PLF_TICK(3);
      if (tid < 32) {
        const bool okSolve = ldlt_solve_smem<30, PLF_LD>(S.H, S.b, S.x);
        if (tid == 0) S.ok = okSolve ? 1 : 0;
      }
      PLF_TICK(4);
      __syncthreads();
      if (tid == 0) {
        d_body_update(S.cur, S.itsCur, S.x);
        double Rbw[9], tbw[3];
        d_m3_t(S.cur, Rbw);
        d_m3_v(Rbw, S.cur + 9, tbw);
        for (int i = 0; i < 3; ++i) tbw[i] = -tbw[i];
        d_m3_mul(A.Rcb, Rbw, S.Rcw);
        d_m3_v(A.Rcb, tbw, S.tcw);
        for (int i = 0; i < 3; ++i) S.tcw[i] += A.tcb[i];
      } else if (tid == 32) {
        d_body_update(S.prev, S.itsPrev, S.x + 15);
      }
      __syncthreads();
      PLF_TICK(5);
      ok = S.ok != 0;
      ++cj;
    }
    if (tid == 0) A.iters[round] = cj;
    const float chi2close = 1.5f * chi2Mono[round];
    double cnt[2] = {0, 0};
    if (edgeThread)
      for (int e = tid; e < E; e += 256) {
        const bool st = A.obs[3 * e + 2] >= 0;
        if (outlier[e]) {
          double r[3], Xc[3];
          pio_edge_error(A, S, e, st, r, Xc);
          err[3 * e] = r[0]; err[3 * e + 1] = r[1]; err[3 * e + 2] = r[2];
        }
        const double om = (double)isg[e];
        const double* r = err + 3 * e;
        const float chi2 = (float)(r[0] * om * r[0] + r[1] * om * r[1] + (st ? r[2] * om * r[2] : 0.0));
        bool bad;
        if (!st) {
          const bool bClose = A.closePt[e] != 0;
          const bool depthPos = (S.Rcw[6] * (double)A.xw[3 * e] + S.Rcw[7] * (double)A.xw[3 * e + 1] + S.Rcw[8] * (double)A.xw[3 * e + 2] + S.tcw[2]) > 0.0;
          bad = (chi2 > chi2Mono[round] && !bClose) || (bClose && chi2 > chi2close) || !depthPos;
        } else {
          bad = chi2 > chi2Stereo[round];
        }
        outlier[e] = bad ? 1 : 0;
        cnt[0] += bad ? 1.0 : 0.0;
        cnt[1] += bad ? 0.0 : 1.0;
      }
    block_sum<2>(cnt, S.red);
    nBad = (int)cnt[0];
    nInliers = (int)cnt[1];
    if (round == 2) robust = false;
    if (E + 4 < 10) break;
  }
  __syncthreads();
  if (nInliers < 30 && !A.recInit) {
    double cnt[1] = {0};
    if (edgeThread)
      for (int e = tid; e < E; e += 256) {
        const bool st = A.obs[3 * e + 2] >= 0;
        double r[3], Xc[3];
        pio_edge_error(A, S, e, st, r, Xc);
        err[3 * e] = r[0]; err[3 * e + 1] = r[1]; err[3 * e + 2] = r[2];
        const double om = (double)isg[e];
        const double c2 = r[0] * om * r[0] + r[1] * om * r[1] + (st ? r[2] * om * r[2] : 0.0);
        if (c2 < (double)(st ? 24.f : 18.f)) outlier[e] = 0; else cnt[0] += 1.0;
      }
    block_sum<1>(cnt, S.red);
    nBad = (int)cnt[0];
  }
  __syncthreads();
  PLF_TICK(6);
  if (tid < 21) A.outState[tid] = S.cur[tid];
  if (tid == 0) *A.nRet = E - nBad;
  // ---- 30x30 Hessian in the reference's order (previous frame 0-14, frame 15-29) at the final estimates ----
  double a36[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) a36[k] = 0;
  if (tid == 256) plf_inertial(A, S);
  if (tid == 288) plf_prior(A, S, false);
  if (edgeThread)
    for (int e = tid; e < E; e += 256) {
      if (outlier[e]) continue;
      const bool st = A.obs[3 * e + 2] >= 0;
      double r[3], Xc[3], J[18];
      pio_edge_error(A, S, e, st, r, Xc);
      pio_edge_jacobian(A, Xc, st, J);
      const int D = st ? 3 : 2;
      const double om = (double)isg[e];
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
          double t = 0;
          for (int d = 0; d < D; ++d) t += J[d * 6 + i] * om * J[d * 6 + j];
          a36[i * 6 + j] += t;
        }
    }
  block_partials<36>(a36, S.red);
  if (tid < 36) {
    double s = 0;
    for (int w = 0; w < 8; ++w) s += S.red[w * 36 + tid];
    S.tot[tid] = s;
  }
  for (int t = tid; t < 30 * PLF_LD; t += PLF_NT) S.H[t] = 0;
  __syncthreads();
  plf_assemble(A, S, true);
  if (tid < 36) { const int i = tid / 6, j = tid - 6 * i; S.H[(15 + i) * PLF_LD + 15 + j] += S.tot[tid]; }
  __syncthreads();
  // ---- Marginalize(H, 0, 14): Hcc - Hcp * pinv(Hpp) * Hpc ----
  if (tid < 225) { const int i = tid / 15, j = tid - 15 * i; S.mA[tid] = 0.5 * (S.H[i * PLF_LD + j] + S.H[j * PLF_LD + i]); }
  __syncthreads();
  if (tid < 32) plf_jacobi15(S);
  __syncthreads();
  if (tid < 225) {
    const int i = tid / 15, j = tid - 15 * i;
    double s = 0;
    for (int k = 0; k < 15; ++k) {
      const double ev = S.mA[k * 15 + k];
      const double iv = fabs(ev) > 1e-6 ? 1.0 / ev : 0.0;
      s += S.mV[i * 15 + k] * iv * S.mV[j * 15 + k];
    }
    S.mInv[tid] = s;
  }
  __syncthreads();
  if (tid < 225) {
    const int i = tid / 15, k = tid - 15 * i;
    double t = 0;
    for (int l = 0; l < 15; ++l) t += S.H[(15 + i) * PLF_LD + l] * S.mInv[l * 15 + k];
    S.mT[tid] = t;
  }
  __syncthreads();
  if (tid < 225) {
    const int i = tid / 15, j = tid - 15 * i;
    double s = 0;
    for (int k = 0; k < 15; ++k) s += S.mT[i * 15 + k] * S.H[k * PLF_LD + 15 + j];
    A.H15[tid] = S.H[(15 + i) * PLF_LD + 15 + j] - s;
  }
  PLF_TICK(7);
}

// =====================================================================================
// K20/K21 driven from the tracker (SURVEY.md §8 f3 + e; src/Tracking.cc:2466-2490): the per-problem argument blocks are
// filled ON THE DEVICE from the tracker's buffers -- the frame's body state is derived from the pose the first
// PoseOptimization left (Frame::GetImuRotation / GetImuPosition, src/Frame.cc:534-554, float cv::Mat arithmetic:
// ((a0*b0 + a1*b1) + a2*b2) for plain products, double accumulation where an operand is a lazy transpose) -- and the optimised state is written back as the
// frame's Tcw the way Frame::SetImuPoseVelocity does it (src/Frame.cc:520-530: double -> float, Tbw, Tcb * Tbw).
// =====================================================================================
template <class ARGS>
__device__ void inertial_fill_common(ARGS& A, const OrbxInertialSlices& I, int s) {
  const int o = I.estart[s];
  A.E = I.ecount[s];
  A.xw = I.exw + 3 * (size_t)o; A.obs = I.eobs + 3 * (size_t)o; A.invSigma2 = I.eisg + o; A.closePt = I.eclose + o;
  A.fx = I.fx; A.fy = I.fy; A.cx = I.cx; A.cy = I.cy; A.bf = I.bf;
  const float* T = I.T1 + 16 * (size_t)s;
  const float* Tcb = I.Tcb;
  float Rwc[9], tcw[3], Ow[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      A.Rcb[i * 3 + j] = Tcb[i * 4 + j]; A.Rbc[j * 3 + i] = Tcb[i * 4 + j]; A.Rcw0[i * 3 + j] = T[i * 4 + j];
      Rwc[j * 3 + i] = T[i * 4 + j];
    }
    A.tcb[i] = Tcb[i * 4 + 3]; A.tbc[i] = I.Tbc[i * 4 + 3]; A.tcw0[i] = T[i * 4 + 3];
    tcw[i] = T[i * 4 + 3];
  }
  for (int i = 0; i < 3; ++i)   // mOw = -mRcw.t()*mtcw: gemm's general path (transposed operand), double accumulation
    Ow[i] = (float)(-((double)Rwc[i * 3] * (double)tcw[0] + (double)Rwc[i * 3 + 1] * (double)tcw[1] + (double)Rwc[i * 3 + 2] * (double)tcw[2]));
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)                                  // GetImuRotation: mRwc * Rcb
      A.state[i * 3 + j] = (double)((Rwc[i * 3] * Tcb[j] + Rwc[i * 3 + 1] * Tcb[4 + j]) + Rwc[i * 3 + 2] * Tcb[8 + j]);
    // GetImuPosition: mRwc * tcb + mOw
    A.state[9 + i] = (double)(((Rwc[i * 3] * Tcb[3] + Rwc[i * 3 + 1] * Tcb[7]) + Rwc[i * 3 + 2] * Tcb[11]) + Ow[i]);
    A.state[12 + i] = (double)I.vel[3 * (size_t)s + i];
  }
  for (int i = 0; i < 6; ++i) A.state[15 + i] = (double)I.bias[6 * (size_t)s + i];
  for (int i = 0; i < 81; ++i) A.infoI[i] = I.infoI[81 * (size_t)s + i];
  for (int i = 0; i < 9; ++i) { A.infoG[i] = I.infoG[9 * (size_t)s + i]; A.infoA[i] = I.infoA[9 * (size_t)s + i]; }
  A.recInit = I.recInit;
  A.outlier = I.eoutlier + o;
  A.err = I.err + 3 * (size_t)o;
  A.outState = I.stateOut + 21 * (size_t)s;
  A.H15 = I.H15 + 225 * (size_t)s;
  A.nRet = I.nRet + s;
  A.iters = I.iters + 4 * (size_t)s;
}

__global__ void inertial_fill_kf_kernel(const OrbxInertialSlices I, PioArgs* args) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= I.S) return;
  PioArgs& A = args[s];
  inertial_fill_common(A, I, s);
  const double* pr = I.preint + 16 * (size_t)s;
  for (int i = 0; i < 21; ++i) A.kf[i] = I.ref[21 * (size_t)s + i];
  for (int i = 0; i < 9; ++i) A.dR[i] = pr[i];
  for (int i = 0; i < 3; ++i) { A.dV[i] = pr[9 + i]; A.dP[i] = pr[12 + i]; }
  A.dt = pr[15];
}

__global__ void inertial_fill_lf_kernel(const OrbxInertialSlices I, PlfArgs* args) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= I.S) return;
  PlfArgs& A = args[s];
  inertial_fill_common(A, I, s);
  const double* pr = I.preint + 16 * (size_t)s;
  const double* pj = I.preintJac + 45 * (size_t)s;
  for (int i = 0; i < 21; ++i) { A.prev[i] = I.ref[21 * (size_t)s + i]; A.prior[i] = I.priorState[21 * (size_t)s + i]; }
  for (int i = 0; i < 9; ++i) {
    A.dR0[i] = pr[i];
    A.JRg[i] = pj[i]; A.JVg[i] = pj[9 + i]; A.JVa[i] = pj[18 + i]; A.JPg[i] = pj[27 + i]; A.JPa[i] = pj[36 + i];
  }
  for (int i = 0; i < 3; ++i) { A.dV0[i] = pr[9 + i]; A.dP0[i] = pr[12 + i]; }
  A.dt = pr[15];
  for (int i = 0; i < 6; ++i) A.bpre[i] = I.preintBias[6 * (size_t)s + i];
  for (int i = 0; i < 225; ++i) A.Hp[i] = I.priorH[225 * (size_t)s + i];
  A.prof = nullptr;
}

// Frame::SetImuPoseVelocity (src/Frame.cc:520-530) on the optimised state: Tcw = Tcb * [Rwb^T | -Rwb^T twb] in float
__global__ void inertial_finish_kernel(const OrbxInertialSlices I) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= I.S) return;
  const double* st = I.stateOut + 21 * (size_t)s;
  float Tbw[16];
  float twb[3];
  for (int i = 0; i < 3; ++i) twb[i] = (float)st[9 + i];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Tbw[i * 4 + j] = (float)st[j * 3 + i];                  // Rbw = Rwb.t()
  for (int i = 0; i < 3; ++i) Tbw[i * 4 + 3] = -((Tbw[i * 4] * twb[0] + Tbw[i * 4 + 1] * twb[1]) + Tbw[i * 4 + 2] * twb[2]);
  Tbw[12] = Tbw[13] = Tbw[14] = 0.f;
  Tbw[15] = 1.f;
  float* T = I.T2 + 16 * (size_t)s;
  const float* Tcb = I.Tcb;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      T[i * 4 + j] = ((Tcb[i * 4] * Tbw[j] + Tcb[i * 4 + 1] * Tbw[4 + j]) + Tcb[i * 4 + 2] * Tbw[8 + j]) + Tcb[i * 4 + 3] * Tbw[12 + j];
}

size_t orbx_inertial_args_bytes(int S) { return (size_t)S * (sizeof(PioArgs) > sizeof(PlfArgs) ? sizeof(PioArgs) : sizeof(PlfArgs)); }

// internal (orbx_track.cu): S problems on fixed-capacity edge slices, everything device-resident
int orbx_launch_pose_inertial_slices(orbx_ctx* ctx, cudaStream_t st, const OrbxInertialSlices& I, void* d_args) {
  if (I.mode != 1 && I.mode != 2) return ORBX_EINVAL;
  const int nb = (I.S + 63) / 64;
  if (I.mode == 1) {
    inertial_fill_kf_kernel<<<nb, 64, 0, st>>>(I, (PioArgs*)d_args);
    ORBX_LAUNCH(ctx);
    pose_inertial_kernel<<<I.S, PIO_NT, 0, st>>>((const PioArgs*)d_args);
    ORBX_LAUNCH(ctx);
  } else {
    inertial_fill_lf_kernel<<<nb, 64, 0, st>>>(I, (PlfArgs*)d_args);
    ORBX_LAUNCH(ctx);
    pose_inertial_lf_kernel<<<I.S, PLF_NT, 0, st>>>((const PlfArgs*)d_args);
    ORBX_LAUNCH(ctx);
  }
  inertial_finish_kernel<<<nb, 64, 0, st>>>(I);
  ORBX_LAUNCH(ctx);
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}

// =====================================================================================
// host entry points
// =====================================================================================
extern "C" {

// Tuning hook: enable = 1 switches the per-phase cycle counters of pose_opt_kernel on (and clears them), 0 off;
// out (may be null) receives the 16 totals accumulated so far (layout: g_po_prof above).
int orbx_debug_pose_opt_profile(orbx_ctx* ctx, int enable, unsigned long long* out) {
  if (!ctx) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(ctx->device));
  ORBX_CUDA(cudaDeviceSynchronize());
  if (out) ORBX_CUDA(cudaMemcpyFromSymbol(out, g_po_prof, sizeof(unsigned long long) * 16));
  if (enable) {
    unsigned long long z[16] = {0};
    ORBX_CUDA(cudaMemcpyToSymbol(g_po_prof, z, sizeof z));
  }
  g_po_profile = enable ? 1 : 0;
  return ORBX_OK;
}

int orbx_pose_optimization_batch_device(orbx_ctx* ctx, int P, const int32_t* d_edge_ofs, const float* d_xw,
                                        const float* d_obs, const float* d_inv_sigma2, const orbx_camera* cam,
                                        float* d_Tcw, uint8_t* d_outlier, int32_t* d_n_inliers, int32_t* d_iters,
                                        double* d_scratch) {
  if (!ctx || P < 1 || !d_edge_ofs || !d_xw || !d_obs || !d_inv_sigma2 || !cam || !d_Tcw || !d_outlier || !d_n_inliers ||
      !d_iters || !d_scratch)
    return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(ctx->device));
  PoseOptArgs A;
  A.edgeOfs = d_edge_ofs;
  A.edgeStart = nullptr;
  A.edgeCount = nullptr;
  A.xw = d_xw;
  A.obs = d_obs;
  A.invSigma2 = d_inv_sigma2;
  A.fx = cam->fx; A.fy = cam->fy; A.cx = cam->cx; A.cy = cam->cy; A.bf = cam->bf;
  A.Tcw = d_Tcw;
  A.outlier = d_outlier;
  A.nInliers = d_n_inliers;
  A.iters = d_iters;
  A.err = d_scratch;
  A.profile = g_po_profile;
  pose_opt_kernel<256><<<P, 256, 0, ctx->stream>>>(A);
  ORBX_LAUNCH(ctx);
  ORBX_CUDA(cudaGetLastError());
  return ORBX_OK;
}

int orbx_pose_optimization(orbx_ctx* ctx, int n_edges, const float* xw, const float* obs, const float* inv_sigma2,
                           const orbx_camera* cam, float* Tcw, uint8_t* outlier, int32_t* n_inliers, int32_t* iters) {
  if (!ctx || n_edges < 0 || !cam || !Tcw || !n_inliers || !iters) return ORBX_EINVAL;
  if (n_edges > 0 && (!xw || !obs || !inv_sigma2 || !outlier)) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DevScope S(ctx, st);
  const int ofs[2] = {0, n_edges};
  int* d_ofs = S.upload(ofs, 2);
  float* d_xw = S.upload(xw, (size_t)3 * n_edges);
  float* d_obs = S.upload(obs, (size_t)3 * n_edges);
  float* d_isg = S.upload(inv_sigma2, n_edges);
  float* d_T = S.upload(Tcw, 16);
  uint8_t* d_out = S.alloc<uint8_t>(n_edges);
  int* d_res = S.alloc<int>(5);
  double* d_scr = S.alloc<double>((size_t)3 * n_edges);
  if (S.failed) return ORBX_ECUDA;
  int rc = orbx_pose_optimization_batch_device(ctx, 1, d_ofs, d_xw, d_obs, d_isg, cam, d_T, d_out, d_res, d_res + 1, d_scr);
  if (rc != ORBX_OK) return rc;
  int h[5];
  ORBX_CUDA(cudaMemcpyAsync(h, d_res, sizeof h, cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(cudaMemcpyAsync(Tcw, d_T, 16 * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (n_edges > 0) ORBX_CUDA(cudaMemcpyAsync(outlier, d_out, n_edges, cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(cudaStreamSynchronize(st));
  *n_inliers = h[0];
  for (int i = 0; i < 4; ++i) iters[i] = h[1 + i];
  return ORBX_OK;
}

int orbx_local_ba(orbx_ctx* ctx, int n_kf, float* kf_Tcw, const uint8_t* kf_fixed, int n_mp, float* mp_xyz, int n_edges,
                  const int32_t* e_kf, const int32_t* e_mp, const float* e_obs, const float* e_inv_sigma2,
                  const orbx_camera* cam, double lambda_init, const volatile uint8_t* stop_flag, uint8_t* edge_bad,
                  int32_t* iters, int32_t* status) {
  if (!ctx || n_kf < 1 || !kf_Tcw || !kf_fixed || n_mp < 1 || !mp_xyz || n_edges < 1 || !e_kf || !e_mp || !e_obs ||
      !e_inv_sigma2 || !cam || !edge_bad || !iters || !status)
    return ORBX_EINVAL;
  for (int e = 0; e < n_edges; ++e)
    if (e_kf[e] < 0 || e_kf[e] >= n_kf || e_mp[e] < 0 || e_mp[e] >= n_mp) {
      orbx_set_error("orbx_local_ba: edge %d references vertex out of range", e);
      return ORBX_EINVAL;
    }
  ORBX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DevScope S(ctx, st);
  LbaArgs A;
  A.K = n_kf; A.M = n_mp; A.E = n_edges;
  // --- graph indices (the reference builds the same adjacency inside g2o's buildStructure) ---
  std::vector<int> hidx(n_kf, -1);
  int nFree = 0;
  for (int k = 0; k < n_kf; ++k)
    if (!kf_fixed[k]) hidx[k] = nFree++;
  A.nFree = nFree;
  A.n = 6 * nFree;
  std::vector<int> ptAllOfs(n_mp + 1, 0), ptOfs(n_mp + 1, 0), kfOfs(nFree + 1, 0);
  for (int e = 0; e < n_edges; ++e) {
    ++ptAllOfs[e_mp[e] + 1];
    if (hidx[e_kf[e]] >= 0) { ++ptOfs[e_mp[e] + 1]; ++kfOfs[hidx[e_kf[e]] + 1]; }
  }
  for (int m = 0; m < n_mp; ++m) { ptAllOfs[m + 1] += ptAllOfs[m]; ptOfs[m + 1] += ptOfs[m]; }
  for (int k = 0; k < nFree; ++k) kfOfs[k + 1] += kfOfs[k];
  std::vector<int> ptAllEdges(std::max(n_edges, 1)), ptEdges(std::max(ptOfs[n_mp], 1)), kfEdges(std::max(kfOfs[nFree], 1));
  std::vector<int> obsEdge((size_t)n_mp * std::max(nFree, 1), -1);
  {
    std::vector<int> a(ptAllOfs.begin(), ptAllOfs.end() - 1), b(ptOfs.begin(), ptOfs.end() - 1), c(kfOfs.begin(), kfOfs.end() - 1);
    for (int e = 0; e < n_edges; ++e) {
      ptAllEdges[a[e_mp[e]]++] = e;
      const int hk = hidx[e_kf[e]];
      if (hk >= 0) {
        ptEdges[b[e_mp[e]]++] = e;
        kfEdges[c[hk]++] = e;
        obsEdge[(size_t)e_mp[e] * nFree + hk] = e;   // a (KeyFrame, MapPoint) pair has one observation
      }
    }
  }
  A.kfT = S.upload(kf_Tcw, (size_t)16 * n_kf);
  A.kfFixed = S.upload(kf_fixed, n_kf);
  A.mpXyz = S.upload(mp_xyz, (size_t)3 * n_mp);
  A.ekf = S.upload(e_kf, n_edges);
  A.emp = S.upload(e_mp, n_edges);
  A.obs = S.upload(e_obs, (size_t)3 * n_edges);
  A.invSigma2 = S.upload(e_inv_sigma2, n_edges);
  A.fx = cam->fx; A.fy = cam->fy; A.cx = cam->cx; A.cy = cam->cy; A.bf = cam->bf;
  A.lambdaInit = lambda_init;
  A.hidx = S.upload(hidx.data(), n_kf);
  A.ptOfs = S.upload(ptOfs.data(), n_mp + 1);
  A.ptEdges = S.upload(ptEdges.data(), ptEdges.size());
  A.ptAllOfs = S.upload(ptAllOfs.data(), n_mp + 1);
  A.ptAllEdges = S.upload(ptAllEdges.data(), ptAllEdges.size());
  A.kfOfs = S.upload(kfOfs.data(), nFree + 1);
  A.kfEdges = S.upload(kfEdges.data(), kfEdges.size());
  A.obsEdge = S.upload(obsEdge.data(), obsEdge.size());
  A.pose = S.alloc<SE3d>(n_kf);
  A.poseBak = S.alloc<SE3d>(n_kf);
  A.pt = S.alloc<double>((size_t)3 * n_mp);
  A.ptBak = S.alloc<double>((size_t)3 * n_mp);
  A.err = S.alloc<double>((size_t)3 * n_edges);
  A.Ji = S.alloc<double>((size_t)9 * n_edges);
  A.Jj = S.alloc<double>((size_t)18 * n_edges);
  A.wom = S.alloc<double>(n_edges);
  A.omr = S.alloc<double>((size_t)3 * n_edges);
  A.Hpl = S.alloc<double>((size_t)18 * n_edges);
  A.BD = S.alloc<double>((size_t)18 * n_edges);
  A.Hll = S.alloc<double>((size_t)9 * n_mp);
  A.Dinv = S.alloc<double>((size_t)9 * n_mp);
  A.Hpp = S.alloc<double>((size_t)36 * std::max(nFree, 1));
  A.b = S.alloc<double>((size_t)A.n + 3 * n_mp);
  A.x = S.alloc<double>((size_t)A.n + 3 * n_mp);
  A.S = S.alloc<double>((size_t)std::max(A.n, 1) * std::max(A.n, 1));
  A.bs = S.alloc<double>(std::max(A.n, 1));
  A.db = S.alloc<double>((size_t)3 * n_mp);
  A.tmp = S.alloc<double>(std::max((size_t)n_edges, (size_t)A.n + 3 * n_mp));
  A.part = S.alloc<double>(96);
  A.gmax = S.alloc<double>(1);
  A.gflag = S.alloc<int>(4);
  {
    const int nBlocks = nFree * (nFree + 1) / 2;
    std::vector<int> blkOfs(nBlocks + 1, 0);
    int blk = 0;
    for (int bi = 0; bi < nFree; ++bi)
      for (int bj = bi; bj < nFree; ++bj, ++blk) blkOfs[blk + 1] = blkOfs[blk] + (kfOfs[bi + 1] - kfOfs[bi]);
    A.blkOfs = S.upload(blkOfs.data(), blkOfs.size());
    A.pairE2 = S.alloc<int>(std::max(blkOfs[nBlocks], 1));
  }
  A.prof = nullptr;
  {
    const char* pe = getenv("ORBX_LBA_PROFILE");
    if (pe && pe[0] == '1') {
      A.prof = S.alloc<unsigned long long>(16);
      if (A.prof) cudaMemsetAsync(A.prof, 0, 16 * sizeof(unsigned long long), st);
    }
  }
  A.edgeBad = S.alloc<uint8_t>(n_edges);
  int* d_res = S.alloc<int>(3);
  A.iters = d_res;
  A.status = d_res + 2;
  // stop flag: a mapped pinned byte the kernel polls; the host forwards *stop_flag into it while waiting
  uint8_t* h_flag = nullptr;
  uint8_t* d_flag = nullptr;
  if (stop_flag) {
    ORBX_CUDA(cudaHostAlloc(&h_flag, 64, cudaHostAllocMapped));
    *h_flag = *stop_flag ? 1 : 0;
    ORBX_CUDA(cudaHostGetDevicePointer(&d_flag, h_flag, 0));
  }
  A.stop = d_flag;
  if (S.failed) { if (h_flag) cudaFreeHost(h_flag); return ORBX_ECUDA; }
  // cooperative multi-CTA kernel (one CTA per SM) whenever the device can co-schedule it; the single-CTA kernel
  // otherwise (ORBX_LBA_SINGLE_CTA=1 forces it, for A/B timing)
  bool launched = false;
  {
    static int coopOk = -1, maxCtasPerSm = 0, smemConfigured = 0;
    const size_t smem = A.n <= LBC_MAX_SMEM_N ? sizeof(double) * ((size_t)A.n * A.n + A.n) : 0;
    if (coopOk < 0) {
      int v = 0;
      cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, ctx->device);
      const char* env = getenv("ORBX_LBA_SINGLE_CTA");
      coopOk = (v && !(env && env[0] == '1')) ? 1 : 0;
    }
    if (coopOk == 1) {
      if ((int)smem > smemConfigured) {
        const int want = (int)(sizeof(double) * ((size_t)LBC_MAX_SMEM_N * LBC_MAX_SMEM_N + LBC_MAX_SMEM_N));
        if (cudaFuncSetAttribute(lba_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, want) == cudaSuccess) smemConfigured = want;
      }
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxCtasPerSm, lba_coop_kernel, LBC_NT, smem) == cudaSuccess && maxCtasPerSm >= 1 &&
          ctx->sm_count * LBC_NT >= 1024) {
        void* kargs[] = {(void*)&A};
        cudaError_t ce = cudaLaunchCooperativeKernel((void*)lba_coop_kernel, dim3(ctx->sm_count), dim3(LBC_NT), kargs, smem, st);
        if (ce == cudaSuccess) launched = true;
        else cudaGetLastError();
      }
    }
  }
  if (!launched) lba_kernel<<<1, LBA_NT, 0, st>>>(A);
  ORBX_LAUNCH(ctx);
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) {
    if (h_flag) cudaFreeHost(h_flag);
    orbx_set_error("orbx_local_ba: launch failed: %s", cudaGetErrorString(le));
    return ORBX_ECUDA;
  }
  int h[3];
  cudaEvent_t done;
  ORBX_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
  ORBX_CUDA(cudaMemcpyAsync(h, d_res, sizeof h, cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(cudaMemcpyAsync(edge_bad, A.edgeBad, n_edges, cudaMemcpyDeviceToHost, st));
  ORBX_CUDA(cudaEventRecord(done, st));
  if (stop_flag) {
    while (cudaEventQuery(done) == cudaErrorNotReady)
      if (*stop_flag) *(volatile uint8_t*)h_flag = 1;   // forward Tracking's InterruptBA() to the device
  }
  cudaError_t e2 = cudaEventSynchronize(done);
  cudaEventDestroy(done);
  if (e2 != cudaSuccess) {
    if (h_flag) cudaFreeHost(h_flag);
    orbx_set_error("orbx_local_ba: %s", cudaGetErrorString(e2));
    return ORBX_ECUDA;
  }
  iters[0] = h[0];
  iters[1] = h[1];
  *status = h[2];
  if (A.prof) {
    unsigned long long hp[16];
    if (cudaMemcpy(hp, A.prof, sizeof hp, cudaMemcpyDeviceToHost) == cudaSuccess) {
      static const char* name[13] = {"stop poll", "errors+chi2", "edge Jacobians", "Hll+Hpp", "Dinv", "B*Dinv", "Schur+b", "LDLT+subst (CTA 0)",
                                     "landmark x", "oplus", "trial errors+chi2", "scale", "accept/pop+stop"};
      fprintf(stderr, "orbx_local_ba phases (us, %llu trials):", hp[15]);
      for (int i = 0; i < 13; ++i) fprintf(stderr, " %s=%.0f", name[i], hp[i] * 1e-3);
      fprintf(stderr, "\n");
    }
  }
  if (h[2] == 0) {
    ORBX_CUDA(cudaMemcpyAsync(kf_Tcw, A.kfT, sizeof(float) * 16 * n_kf, cudaMemcpyDeviceToHost, st));
    ORBX_CUDA(cudaMemcpyAsync(mp_xyz, A.mpXyz, sizeof(float) * 3 * n_mp, cudaMemcpyDeviceToHost, st));
    ORBX_CUDA(cudaStreamSynchronize(st));
  }
  if (h_flag) cudaFreeHost(h_flag);
  return ORBX_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Prepared many-problem LocalBundleAdjustment (include/orbx.h: orbx_lba_batch_*): P independent problems, one CTA each,
// one launch.  prepare() builds every problem's adjacency exactly as orbx_local_ba does, lays inputs, indices and
// scratch out in ONE device pool and keeps a pristine copy of the in/out arrays; run() restores them (two device copies)
// and launches; fetch() synchronises and downloads poses, points, edge flags, iteration counts and status.
// ------------------------------------------------------------------------------------------------------------------
struct orbx_lba_batch {
  orbx_ctx* ctx = nullptr;
  int device = 0;               // kept separately: destroy() must not dereference a context that may be gone already
  int P = 0;
  uint8_t* pool = nullptr;
  size_t poolBytes = 0;
  LbaArgs* dArgs = nullptr;
  float *dKfT = nullptr, *dKfT0 = nullptr, *dMp = nullptr, *dMp0 = nullptr;   // all problems back to back
  size_t kfFloats = 0, mpFloats = 0;
  uint8_t* dBad = nullptr;
  size_t badBytes = 0;
  int* dRes = nullptr;                                                         // [P][3] iters[2], status
  std::vector<size_t> kfOfs, mpOfs, badOfs;
  std::vector<int> nKf, nMp, nE;
  cudaStream_t lastStream = nullptr;
};

orbx_lba_batch* orbx_lba_batch_prepare(orbx_ctx* ctx, int P, const orbx_lba_problem* pr, const orbx_camera* cam) {
  if (!ctx || P < 1 || !pr || !cam) {
    orbx_set_error("orbx_lba_batch_prepare: invalid argument");
    return nullptr;
  }
  for (int p = 0; p < P; ++p) {
    const orbx_lba_problem& Q = pr[p];
    if (Q.n_kf < 1 || Q.n_mp < 1 || Q.n_edges < 1 || !Q.kf_Tcw || !Q.kf_fixed || !Q.mp_xyz || !Q.e_kf || !Q.e_mp || !Q.e_obs || !Q.e_inv_sigma2) {
      orbx_set_error("orbx_lba_batch_prepare: problem %d is incomplete", p);
      return nullptr;
    }
    for (int e = 0; e < Q.n_edges; ++e)
      if (Q.e_kf[e] < 0 || Q.e_kf[e] >= Q.n_kf || Q.e_mp[e] < 0 || Q.e_mp[e] >= Q.n_mp) {
        orbx_set_error("orbx_lba_batch_prepare: problem %d, edge %d references a vertex out of range", p, e);
        return nullptr;
      }
  }
  if (cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
  orbx_lba_batch* L = new orbx_lba_batch();
  L->ctx = ctx; L->device = ctx->device; L->P = P;
  L->kfOfs.assign(P + 1, 0); L->mpOfs.assign(P + 1, 0); L->badOfs.assign(P + 1, 0);
  L->nKf.resize(P); L->nMp.resize(P); L->nE.resize(P);
  for (int p = 0; p < P; ++p) {
    L->nKf[p] = pr[p].n_kf; L->nMp[p] = pr[p].n_mp; L->nE[p] = pr[p].n_edges;
    L->kfOfs[p + 1] = L->kfOfs[p] + (size_t)16 * pr[p].n_kf;
    L->mpOfs[p + 1] = L->mpOfs[p] + (size_t)3 * pr[p].n_mp;
    L->badOfs[p + 1] = L->badOfs[p] + (((size_t)pr[p].n_edges + 15) & ~(size_t)15);
  }
  L->kfFloats = L->kfOfs[P]; L->mpFloats = L->mpOfs[P]; L->badBytes = L->badOfs[P];
  PoolBuilder B;
  // in/out arrays of all problems, back to back, pristine copies first
  const size_t oKf0 = B.reserve(sizeof(float) * L->kfFloats), oMp0 = B.reserve(sizeof(float) * L->mpFloats);
  for (int p = 0; p < P; ++p) {
    memcpy(B.h.data() + oKf0 + sizeof(float) * L->kfOfs[p], pr[p].kf_Tcw, sizeof(float) * 16 * (size_t)pr[p].n_kf);
    memcpy(B.h.data() + oMp0 + sizeof(float) * L->mpOfs[p], pr[p].mp_xyz, sizeof(float) * 3 * (size_t)pr[p].n_mp);
  }
  const size_t oKf = B.reserve(sizeof(float) * L->kfFloats), oMp = B.reserve(sizeof(float) * L->mpFloats);
  const size_t oBad = B.reserve(L->badBytes), oRes = B.reserve(sizeof(int) * 3 * (size_t)P);
  struct Ofs { size_t fixed, ekf, emp, obs, isg, hidx, ptOfs, ptEdges, ptAllOfs, ptAllEdges, kfOfs, kfEdges, obsEdge, pose, poseBak, pt, ptBak,
                      err, Ji, Jj, wom, omr, Hpl, BD, Hll, Dinv, Hpp, b, x, S, bs, db; int nFree; };
  std::vector<Ofs> O(P);
  for (int p = 0; p < P; ++p) {
    const orbx_lba_problem& Q = pr[p];
    const int n_kf = Q.n_kf, n_mp = Q.n_mp, n_edges = Q.n_edges;
    // --- graph indices: identical to orbx_local_ba (the reference builds the same adjacency in g2o's buildStructure) ---
    std::vector<int> hidx(n_kf, -1);
    int nFree = 0;
    for (int k = 0; k < n_kf; ++k)
      if (!Q.kf_fixed[k]) hidx[k] = nFree++;
    std::vector<int> ptAllOfs(n_mp + 1, 0), ptOfs(n_mp + 1, 0), kfOfs(nFree + 1, 0);
    for (int e = 0; e < n_edges; ++e) {
      ++ptAllOfs[Q.e_mp[e] + 1];
      if (hidx[Q.e_kf[e]] >= 0) { ++ptOfs[Q.e_mp[e] + 1]; ++kfOfs[hidx[Q.e_kf[e]] + 1]; }
    }
    for (int m = 0; m < n_mp; ++m) { ptAllOfs[m + 1] += ptAllOfs[m]; ptOfs[m + 1] += ptOfs[m]; }
    for (int k = 0; k < nFree; ++k) kfOfs[k + 1] += kfOfs[k];
    std::vector<int> ptAllEdges(std::max(n_edges, 1)), ptEdges(std::max(ptOfs[n_mp], 1)), kfEdges(std::max(kfOfs[nFree], 1));
    std::vector<int> obsEdge((size_t)n_mp * std::max(nFree, 1), -1);
    {
      std::vector<int> a(ptAllOfs.begin(), ptAllOfs.end() - 1), b(ptOfs.begin(), ptOfs.end() - 1), c(kfOfs.begin(), kfOfs.end() - 1);
      for (int e = 0; e < n_edges; ++e) {
        ptAllEdges[a[Q.e_mp[e]]++] = e;
        const int hk = hidx[Q.e_kf[e]];
        if (hk >= 0) {
          ptEdges[b[Q.e_mp[e]]++] = e;
          kfEdges[c[hk]++] = e;
          obsEdge[(size_t)Q.e_mp[e] * nFree + hk] = e;
        }
      }
    }
    Ofs& o = O[p];
    o.nFree = nFree;
    const int n = 6 * nFree;
    o.fixed = B.add(Q.kf_fixed, n_kf);
    o.ekf = B.add(Q.e_kf, sizeof(int) * (size_t)n_edges);
    o.emp = B.add(Q.e_mp, sizeof(int) * (size_t)n_edges);
    o.obs = B.add(Q.e_obs, sizeof(float) * 3 * (size_t)n_edges);
    o.isg = B.add(Q.e_inv_sigma2, sizeof(float) * (size_t)n_edges);
    o.hidx = B.addv(hidx); o.ptOfs = B.addv(ptOfs); o.ptEdges = B.addv(ptEdges); o.ptAllOfs = B.addv(ptAllOfs);
    o.ptAllEdges = B.addv(ptAllEdges); o.kfOfs = B.addv(kfOfs); o.kfEdges = B.addv(kfEdges); o.obsEdge = B.addv(obsEdge);
    const size_t E = (size_t)n_edges, M = (size_t)n_mp, D = sizeof(double);
    o.pose = B.reserve(sizeof(SE3d) * n_kf); o.poseBak = B.reserve(sizeof(SE3d) * n_kf);
    o.pt = B.reserve(D * 3 * M); o.ptBak = B.reserve(D * 3 * M);
    o.err = B.reserve(D * 3 * E); o.Ji = B.reserve(D * 9 * E); o.Jj = B.reserve(D * 18 * E); o.wom = B.reserve(D * E); o.omr = B.reserve(D * 3 * E);
    o.Hpl = B.reserve(D * 18 * E); o.BD = B.reserve(D * 18 * E); o.Hll = B.reserve(D * 9 * M); o.Dinv = B.reserve(D * 9 * M);
    o.Hpp = B.reserve(D * 36 * (size_t)std::max(nFree, 1));
    o.b = B.reserve(D * ((size_t)n + 3 * M)); o.x = B.reserve(D * ((size_t)n + 3 * M));
    o.S = B.reserve(D * (size_t)std::max(n, 1) * std::max(n, 1)); o.bs = B.reserve(D * (size_t)std::max(n, 1)); o.db = B.reserve(D * 3 * M);
  }
  const size_t oArgs = B.reserve(sizeof(LbaArgs) * (size_t)P);
  if (cudaMalloc(&L->pool, B.size()) != cudaSuccess) {
    orbx_set_error("orbx_lba_batch_prepare: cudaMalloc(%zu bytes) failed", B.size());
    cudaGetLastError();
    delete L;
    return nullptr;
  }
  L->poolBytes = B.size();
  uint8_t* base = L->pool;
  std::vector<LbaArgs> A(P);
  for (int p = 0; p < P; ++p) {
    const orbx_lba_problem& Q = pr[p];
    const Ofs& o = O[p];
    LbaArgs& a = A[p];
    memset(&a, 0, sizeof a);
    a.K = Q.n_kf; a.M = Q.n_mp; a.E = Q.n_edges; a.nFree = o.nFree; a.n = 6 * o.nFree;
    a.kfT = (float*)(base + oKf) + L->kfOfs[p];
    a.kfFixed = base + o.fixed;
    a.mpXyz = (float*)(base + oMp) + L->mpOfs[p];
    a.ekf = (const int*)(base + o.ekf); a.emp = (const int*)(base + o.emp);
    a.obs = (const float*)(base + o.obs); a.invSigma2 = (const float*)(base + o.isg);
    a.fx = cam->fx; a.fy = cam->fy; a.cx = cam->cx; a.cy = cam->cy; a.bf = cam->bf;
    a.lambdaInit = Q.lambda_init;
    a.stop = nullptr;
    a.hidx = (const int*)(base + o.hidx); a.ptOfs = (const int*)(base + o.ptOfs); a.ptEdges = (const int*)(base + o.ptEdges);
    a.ptAllOfs = (const int*)(base + o.ptAllOfs); a.ptAllEdges = (const int*)(base + o.ptAllEdges);
    a.kfOfs = (const int*)(base + o.kfOfs); a.kfEdges = (const int*)(base + o.kfEdges); a.obsEdge = (const int*)(base + o.obsEdge);
    a.pose = (SE3d*)(base + o.pose); a.poseBak = (SE3d*)(base + o.poseBak);
    a.pt = (double*)(base + o.pt); a.ptBak = (double*)(base + o.ptBak);
    a.err = (double*)(base + o.err); a.Ji = (double*)(base + o.Ji); a.Jj = (double*)(base + o.Jj); a.wom = (double*)(base + o.wom);
    a.omr = (double*)(base + o.omr); a.Hpl = (double*)(base + o.Hpl); a.BD = (double*)(base + o.BD); a.Hll = (double*)(base + o.Hll);
    a.Dinv = (double*)(base + o.Dinv); a.Hpp = (double*)(base + o.Hpp); a.b = (double*)(base + o.b); a.x = (double*)(base + o.x);
    a.S = (double*)(base + o.S); a.bs = (double*)(base + o.bs); a.db = (double*)(base + o.db);
    a.edgeBad = base + oBad + L->badOfs[p];
    a.iters = (int*)(base + oRes) + 3 * p;
    a.status = (int*)(base + oRes) + 3 * p + 2;
  }
  memcpy(B.h.data() + oArgs, A.data(), sizeof(LbaArgs) * (size_t)P);
  // (pageable H2D on the legacy stream returns once the data is staged: wait for the DMA, the plan runs on non-blocking streams)
  if (cudaMemcpy(L->pool, B.h.data(), B.h.size(), cudaMemcpyHostToDevice) != cudaSuccess || cudaStreamSynchronize(0) != cudaSuccess) {
    orbx_set_error("orbx_lba_batch_prepare: upload failed");
    cudaGetLastError();
    cudaFree(L->pool);
    delete L;
    return nullptr;
  }
  L->dArgs = (LbaArgs*)(base + oArgs);
  L->dKfT0 = (float*)(base + oKf0); L->dMp0 = (float*)(base + oMp0);
  L->dKfT = (float*)(base + oKf); L->dMp = (float*)(base + oMp);
  L->dBad = base + oBad;
  L->dRes = (int*)(base + oRes);
  return L;
}

size_t orbx_lba_batch_device_bytes(const orbx_lba_batch* L) { return L ? L->poolBytes : 0; }

int orbx_lba_batch_run(orbx_lba_batch* L, void* cuda_stream) {
  if (!L) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(L->ctx->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : L->ctx->stream;
  ORBX_CUDA(cudaMemcpyAsync(L->dKfT, L->dKfT0, sizeof(float) * L->kfFloats, cudaMemcpyDeviceToDevice, st));
  ORBX_CUDA(cudaMemcpyAsync(L->dMp, L->dMp0, sizeof(float) * L->mpFloats, cudaMemcpyDeviceToDevice, st));
  lba_batch_kernel<<<L->P, LBA_NT, 0, st>>>(L->dArgs);
  ORBX_LAUNCH(L->ctx);
  ORBX_CUDA(cudaGetLastError());
  L->lastStream = st;
  return ORBX_OK;
}

int orbx_lba_batch_fetch(orbx_lba_batch* L, orbx_lba_problem* pr) {
  if (!L || !pr) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(L->ctx->device));
  ORBX_CUDA(cudaStreamSynchronize(L->lastStream ? L->lastStream : L->ctx->stream));
  std::vector<float> kf(L->kfFloats), mp(L->mpFloats);
  std::vector<uint8_t> bad(L->badBytes);
  std::vector<int> res(3 * (size_t)L->P);
  ORBX_CUDA(cudaMemcpy(kf.data(), L->dKfT, sizeof(float) * L->kfFloats, cudaMemcpyDeviceToHost));
  ORBX_CUDA(cudaMemcpy(mp.data(), L->dMp, sizeof(float) * L->mpFloats, cudaMemcpyDeviceToHost));
  ORBX_CUDA(cudaMemcpy(bad.data(), L->dBad, L->badBytes, cudaMemcpyDeviceToHost));
  ORBX_CUDA(cudaMemcpy(res.data(), L->dRes, sizeof(int) * res.size(), cudaMemcpyDeviceToHost));
  for (int p = 0; p < L->P; ++p) {
    pr[p].iters[0] = res[3 * (size_t)p];
    pr[p].iters[1] = res[3 * (size_t)p + 1];
    pr[p].status = res[3 * (size_t)p + 2];
    if (pr[p].edge_bad) memcpy(pr[p].edge_bad, bad.data() + L->badOfs[p], (size_t)L->nE[p]);
    if (pr[p].status == 0) {   // an aborted optimisation leaves the caller's poses and points as they were
      if (pr[p].kf_Tcw) memcpy(pr[p].kf_Tcw, kf.data() + L->kfOfs[p], sizeof(float) * 16 * (size_t)L->nKf[p]);
      if (pr[p].mp_xyz) memcpy(pr[p].mp_xyz, mp.data() + L->mpOfs[p], sizeof(float) * 3 * (size_t)L->nMp[p]);
    }
  }
  return ORBX_OK;
}

void orbx_lba_batch_destroy(orbx_lba_batch* L) {
  if (!L) return;
  cudaSetDevice(L->device);
  cudaDeviceSynchronize();   // the stream of the last run may belong to an object that is gone already
  cudaFree(L->pool);
  delete L;
}

// Many-problem form (one CTA per problem, one launch); the single call below is the P = 1 case.
int orbx_pose_inertial_optimization_last_keyframe_batch(orbx_ctx* ctx, int P, const int32_t* edge_ofs, const float* xw, const float* obs,
                                                        const float* inv_sigma2, const uint8_t* close_pt, const orbx_camera* cam,
                                                        const float* Tcw, const float* Tcb, const float* Tbc, double* state,
                                                        const double* kf_state, const double* preint, const double* info_inertial,
                                                        const double* info_gyro, const double* info_acc, int rec_init, uint8_t* outlier,
                                                        double* H15, int32_t* n_ret, int32_t* iters) {
  if (!ctx || P < 0 || !edge_ofs || !cam || !Tcw || !Tcb || !Tbc || !state || !kf_state || !preint || !info_inertial || !info_gyro ||
      !info_acc || !H15 || !n_ret || !iters)
    return ORBX_EINVAL;
  if (P == 0) return ORBX_OK;
  const int total = edge_ofs[P];
  if (edge_ofs[0] != 0 || total < 0) return ORBX_EINVAL;
  for (int p = 0; p < P; ++p) if (edge_ofs[p + 1] < edge_ofs[p]) return ORBX_EINVAL;
  if (total > 0 && (!xw || !obs || !inv_sigma2 || !close_pt || !outlier)) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DevScope S(ctx, st);
  const float* d_xw = S.upload(xw, (size_t)3 * total);
  const float* d_obs = S.upload(obs, (size_t)3 * total);
  const float* d_isg = S.upload(inv_sigma2, total);
  const uint8_t* d_close = S.upload(close_pt, total);
  uint8_t* d_outlier = S.alloc<uint8_t>(total);
  double* d_err = S.alloc<double>((size_t)3 * total);
  double* d_state = S.alloc<double>((size_t)21 * P);
  double* d_H = S.alloc<double>((size_t)225 * P);
  int* d_res = S.alloc<int>((size_t)5 * P);
  std::vector<PioArgs> args(P);
  for (int p = 0; p < P; ++p) {
    PioArgs& A = args[p];
    const int o = edge_ofs[p];
    A.E = edge_ofs[p + 1] - o;
    A.xw = d_xw + 3 * (size_t)o; A.obs = d_obs + 3 * (size_t)o; A.invSigma2 = d_isg + o; A.closePt = d_close + o;
    A.fx = cam->fx; A.fy = cam->fy; A.cx = cam->cx; A.cy = cam->cy; A.bf = cam->bf;
    const float* T = Tcw + 16 * (size_t)p;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) { A.Rcb[i * 3 + j] = Tcb[i * 4 + j]; A.Rbc[j * 3 + i] = Tcb[i * 4 + j]; A.Rcw0[i * 3 + j] = T[i * 4 + j]; }
      A.tcb[i] = Tcb[i * 4 + 3]; A.tbc[i] = Tbc[i * 4 + 3]; A.tcw0[i] = T[i * 4 + 3];
    }
    memcpy(A.state, state + 21 * (size_t)p, sizeof A.state);
    memcpy(A.kf, kf_state + 21 * (size_t)p, sizeof A.kf);
    const double* pr = preint + 16 * (size_t)p;
    memcpy(A.dR, pr, 72); memcpy(A.dV, pr + 9, 24); memcpy(A.dP, pr + 12, 24);
    A.dt = pr[15];
    memcpy(A.infoI, info_inertial + 81 * (size_t)p, sizeof A.infoI);
    memcpy(A.infoG, info_gyro + 9 * (size_t)p, sizeof A.infoG);
    memcpy(A.infoA, info_acc + 9 * (size_t)p, sizeof A.infoA);
    A.recInit = rec_init;
    A.outlier = d_outlier + o;
    A.err = d_err + 3 * (size_t)o;
    A.outState = d_state + 21 * (size_t)p;
    A.H15 = d_H + 225 * (size_t)p;
    A.nRet = d_res + 5 * (size_t)p;
    A.iters = d_res + 5 * (size_t)p + 1;
  }
  PioArgs* dA = S.upload(args.data(), (size_t)P);
  if (S.failed) return ORBX_ECUDA;
  pose_inertial_kernel<<<P, PIO_NT, 0, st>>>(dA);
  ORBX_LAUNCH(ctx);
  ORBX_CUDA(cudaGetLastError());
  std::vector<int32_t> res((size_t)5 * P, 0);
  S.download(state, (const double*)d_state, (size_t)21 * P);
  S.download(H15, (const double*)d_H, (size_t)225 * P);
  S.download(res.data(), (const int32_t*)d_res, (size_t)5 * P);
  if (total > 0) S.download(outlier, (const uint8_t*)d_outlier, (size_t)total);
  int rc = S.finish();
  if (rc != ORBX_OK) return rc;
  for (int p = 0; p < P; ++p) {
    n_ret[p] = res[5 * (size_t)p];
    for (int i = 0; i < 4; ++i) iters[4 * (size_t)p + i] = res[5 * (size_t)p + 1 + i];
  }
  return ORBX_OK;
}

int orbx_pose_inertial_optimization_last_keyframe(orbx_ctx* ctx, int n_edges, const float* xw, const float* obs, const float* inv_sigma2,
                                                  const uint8_t* close_pt, const orbx_camera* cam, const float* Tcw, const float* Tcb,
                                                  const float* Tbc, double* state, const double* kf_state, const double* preint,
                                                  const double* info_inertial, const double* info_gyro, const double* info_acc,
                                                  int rec_init, uint8_t* outlier, double* H15, int32_t* n_ret, int32_t* iters) {
  if (n_edges < 0) return ORBX_EINVAL;
  const int32_t ofs[2] = {0, n_edges};
  return orbx_pose_inertial_optimization_last_keyframe_batch(ctx, 1, ofs, xw, obs, inv_sigma2, close_pt, cam, Tcw, Tcb, Tbc, state, kf_state,
                                                             preint, info_inertial, info_gyro, info_acc, rec_init, outlier, H15, n_ret,
                                                             iters);
}

// Many-problem form of orbx_pose_inertial_optimization_last_frame (one CTA per problem, one launch): P independent streams'
// frames.  edge_ofs[P+1] delimits each problem's slice of xw/obs/inv_sigma2/close_pt/outlier; every other per-problem array
// is the single-call argument with a leading [P] dimension; Tcb/Tbc/cam are shared.
int orbx_pose_inertial_optimization_last_frame_batch(orbx_ctx* ctx, int P, const int32_t* edge_ofs, const float* xw, const float* obs,
                                                     const float* inv_sigma2, const uint8_t* close_pt, const orbx_camera* cam,
                                                     const float* Tcw, const float* Tcb, const float* Tbc, double* state,
                                                     const double* prev_state, const double* preint, const double* preint_jac,
                                                     const double* preint_bias, const double* info_inertial, const double* info_gyro,
                                                     const double* info_acc, const double* prior_state, const double* prior_H,
                                                     int rec_init, uint8_t* outlier, double* H15, int32_t* n_ret, int32_t* iters) {
  if (!ctx || P < 0 || !edge_ofs || !cam || !Tcw || !Tcb || !Tbc || !state || !prev_state || !preint || !preint_jac || !preint_bias ||
      !info_inertial || !info_gyro || !info_acc || !prior_state || !prior_H || !H15 || !n_ret || !iters)
    return ORBX_EINVAL;
  if (P == 0) return ORBX_OK;
  const int total = edge_ofs[P];
  if (edge_ofs[0] != 0 || total < 0) return ORBX_EINVAL;
  for (int p = 0; p < P; ++p) if (edge_ofs[p + 1] < edge_ofs[p]) return ORBX_EINVAL;
  if (total > 0 && (!xw || !obs || !inv_sigma2 || !close_pt || !outlier)) return ORBX_EINVAL;
  ORBX_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  DevScope S(ctx, st);
  const float* d_xw = S.upload(xw, (size_t)3 * total);
  const float* d_obs = S.upload(obs, (size_t)3 * total);
  const float* d_isg = S.upload(inv_sigma2, total);
  const uint8_t* d_close = S.upload(close_pt, total);
  uint8_t* d_outlier = S.alloc<uint8_t>(total);
  double* d_err = S.alloc<double>((size_t)3 * total);
  double* d_state = S.alloc<double>((size_t)21 * P);
  double* d_H = S.alloc<double>((size_t)225 * P);
  int* d_res = S.alloc<int>((size_t)5 * P);
  std::vector<PlfArgs> args(P);
  for (int p = 0; p < P; ++p) {
    PlfArgs& A = args[p];
    const int o = edge_ofs[p];
    A.E = edge_ofs[p + 1] - o;
    A.xw = d_xw + 3 * (size_t)o; A.obs = d_obs + 3 * (size_t)o; A.invSigma2 = d_isg + o; A.closePt = d_close + o;
    A.fx = cam->fx; A.fy = cam->fy; A.cx = cam->cx; A.cy = cam->cy; A.bf = cam->bf;
    const float* T = Tcw + 16 * (size_t)p;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) { A.Rcb[i * 3 + j] = Tcb[i * 4 + j]; A.Rbc[j * 3 + i] = Tcb[i * 4 + j]; A.Rcw0[i * 3 + j] = T[i * 4 + j]; }
      A.tcb[i] = Tcb[i * 4 + 3]; A.tbc[i] = Tbc[i * 4 + 3]; A.tcw0[i] = T[i * 4 + 3];
    }
    memcpy(A.state, state + 21 * (size_t)p, sizeof A.state);
    memcpy(A.prev, prev_state + 21 * (size_t)p, sizeof A.prev);
    const double* pr = preint + 16 * (size_t)p;
    memcpy(A.dR0, pr, 72); memcpy(A.dV0, pr + 9, 24); memcpy(A.dP0, pr + 12, 24);
    A.dt = pr[15];
    const double* pj = preint_jac + 45 * (size_t)p;
    memcpy(A.JRg, pj, 72); memcpy(A.JVg, pj + 9, 72); memcpy(A.JVa, pj + 18, 72); memcpy(A.JPg, pj + 27, 72); memcpy(A.JPa, pj + 36, 72);
    memcpy(A.bpre, preint_bias + 6 * (size_t)p, 48);
    memcpy(A.infoI, info_inertial + 81 * (size_t)p, sizeof A.infoI);
    memcpy(A.infoG, info_gyro + 9 * (size_t)p, sizeof A.infoG);
    memcpy(A.infoA, info_acc + 9 * (size_t)p, sizeof A.infoA);
    memcpy(A.prior, prior_state + 21 * (size_t)p, sizeof A.prior);
    memcpy(A.Hp, prior_H + 225 * (size_t)p, sizeof A.Hp);
    A.recInit = rec_init;
    A.outlier = d_outlier + o;
    A.err = d_err + 3 * (size_t)o;
    A.outState = d_state + 21 * (size_t)p;
    A.H15 = d_H + 225 * (size_t)p;
    A.nRet = d_res + 5 * (size_t)p;
    A.iters = d_res + 5 * (size_t)p + 1;
    A.prof = nullptr;
  }
  {
    const char* pe = getenv("ORBX_PIO_PROFILE");                   // phase timers of problem 0 (single calls)
    if (pe && pe[0] == '1' && P == 1) {
      args[0].prof = S.alloc<unsigned long long>(8);
      if (args[0].prof) cudaMemsetAsync(args[0].prof, 0, 8 * sizeof(unsigned long long), st);
    }
  }
  PlfArgs* dA = S.upload(args.data(), (size_t)P);
  if (S.failed) return ORBX_ECUDA;
  pose_inertial_lf_kernel<<<P, PLF_NT, 0, st>>>(dA);
  ORBX_LAUNCH(ctx);
  ORBX_CUDA(cudaGetLastError());
  std::vector<int32_t> res((size_t)5 * P, 0);
  S.download(state, (const double*)d_state, (size_t)21 * P);
  S.download(H15, (const double*)d_H, (size_t)225 * P);
  S.download(res.data(), (const int32_t*)d_res, (size_t)5 * P);
  if (total > 0) S.download(outlier, (const uint8_t*)d_outlier, (size_t)total);
  int rc = S.finish();
  if (rc != ORBX_OK) return rc;
  if (args[0].prof) {
    unsigned long long hp[8];
    if (cudaMemcpy(hp, args[0].prof, sizeof hp, cudaMemcpyDeviceToHost) == cudaSuccess)
      fprintf(stderr, "[orbx pio-lf] us: edges+inertial %.1f partials %.1f fill %.1f assemble %.1f solve %.1f update %.1f | classify/tail %.1f final-H+marginalise %.1f\n",
              hp[0] / 1e3, hp[1] / 1e3, hp[2] / 1e3, hp[3] / 1e3, hp[4] / 1e3, hp[5] / 1e3, hp[6] / 1e3, hp[7] / 1e3);
  }
  for (int p = 0; p < P; ++p) {
    n_ret[p] = res[5 * (size_t)p];
    for (int i = 0; i < 4; ++i) iters[4 * (size_t)p + i] = res[5 * (size_t)p + 1 + i];
  }
  return ORBX_OK;
}

int orbx_pose_inertial_optimization_last_frame(orbx_ctx* ctx, int n_edges, const float* xw, const float* obs, const float* inv_sigma2,
                                               const uint8_t* close_pt, const orbx_camera* cam, const float* Tcw, const float* Tcb,
                                               const float* Tbc, double* state, const double* prev_state, const double* preint,
                                               const double* preint_jac, const double* preint_bias, const double* info_inertial,
                                               const double* info_gyro, const double* info_acc, const double* prior_state,
                                               const double* prior_H, int rec_init, uint8_t* outlier, double* H15, int32_t* n_ret,
                                               int32_t* iters) {
  if (n_edges < 0) return ORBX_EINVAL;
  const int32_t ofs[2] = {0, n_edges};
  return orbx_pose_inertial_optimization_last_frame_batch(ctx, 1, ofs, xw, obs, inv_sigma2, close_pt, cam, Tcw, Tcb, Tbc, state, prev_state,
                                                          preint, preint_jac, preint_bias, info_inertial, info_gyro, info_acc, prior_state,
                                                          prior_H, rec_init, outlier, H15, n_ret, iters);
}

}  // extern "C"
