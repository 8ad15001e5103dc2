"""CPU: the float matrix algebra of the OpenCV stand-in (oracle/ref_stub/cvstub.cpp), which the reference's ORBmatcher.cc
runs on inside oracle/_ref, is what cv2 4.13 computes: cv::gemm's small-matrix fp32 path (inner dimension 2..4 equal to a
side of the result: a0*b0 + a1*b1 + ... left to right in float) and its double-accumulating general path.  The image
primitives of the stand-in are the oracle's, pinned to cv2 by tests/test_oracle_primitives.py."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include "cvstub.h"
extern "C" void stub_gemm(const float* a, int ar, int ac, const float* b, int bc, float* c) {
  cv::Mat A(ar, ac, CV_32F, (void*)a), B(ac, bc, CV_32F, (void*)b);
  cv::Mat Cm = A * B;
  for (int i = 0; i < ar; ++i) for (int j = 0; j < bc; ++j) c[i * bc + j] = Cm.at<float>(i, j);
}
// flags: 1 = A.t()*B, 2 = A*B.t(), 3 = A.t()*B.t(); neg: the transposed operand is negated first (-A.t()*B)
extern "C" void stub_gemm_t(const float* a, int ar, int ac, const float* b, int br, int bc, int flags, int neg, float* c) {
  cv::Mat A(ar, ac, CV_32F, (void*)a), B(br, bc, CV_32F, (void*)b);
  cv::Mat Cm;
  if (flags == 1) Cm = neg ? cv::Mat(-A.t() * B) : cv::Mat(A.t() * B);
  else if (flags == 2) Cm = neg ? cv::Mat(A * (-B.t())) : cv::Mat(A * B.t());
  else Cm = A.t() * B.t();
  for (int i = 0; i < Cm.rows; ++i) for (int j = 0; j < Cm.cols; ++j) c[i * Cm.cols + j] = Cm.at<float>(i, j);
}
extern "C" void stub_inv3(const float* a, float* c) {
  cv::Mat A(3, 3, CV_32F, (void*)a);
  cv::Mat I = A.inv();
  for (int i = 0; i < 9; ++i) c[i] = I.at<float>(i / 3, i % 3);
}
extern "C" double stub_norm(const float* a, int n) { return cv::norm(cv::Mat(n, 1, CV_32F, (void*)a)); }
extern "C" double stub_dot(const float* a, const float* b, int n) { return cv::Mat(n, 1, CV_32F, (void*)a).dot(cv::Mat(n, 1, CV_32F, (void*)b)); }
'''


@pytest.fixture(scope="module")
def stub(tmp_path_factory):
    d = tmp_path_factory.mktemp("stub")
    (d / "t.cpp").write_text(SRC)
    so = str(d / "libstubtest.so")
    st = os.path.join(ROOT, "oracle", "ref_stub")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++14", "-fPIC", "-shared", "-I", st, "-o", so, str(d / "t.cpp"),
                           os.path.join(st, "cvstub.cpp"), os.path.join(ROOT, "oracle", "ork_primitives.cpp"),
                           os.path.join(ROOT, "oracle", "ork_frame.cpp")])
    L = C.CDLL(so)
    L.stub_norm.restype = C.c_double
    L.stub_dot.restype = C.c_double
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("shape", [(3, 3, 3), (3, 3, 1), (4, 4, 4), (3, 3, 4), (1, 3, 3), (4, 4, 1), (2, 2, 2), (5, 5, 5), (3, 8, 3),
                                   (6, 3, 6), (3, 5, 1)])
def test_gemm_equals_cv2(stub, shape):
    rng = np.random.default_rng(sum(shape))
    for _ in range(200):
        A = rng.normal(size=shape[:2]).astype(np.float32) * np.float32(10 ** rng.uniform(-2, 3))
        B = rng.normal(size=shape[1:]).astype(np.float32)
        Cm = np.zeros((shape[0], shape[2]), np.float32)
        stub.stub_gemm(_p(A), shape[0], shape[1], _p(B), shape[2], _p(Cm))
        assert np.array_equal(Cm, cv2.gemm(A, B, 1, None, 0)), shape


def test_transposed_products_equal_cv2(stub):
    """A.t()*B, A*B.t() (cv::MatExpr -> gemm with GEMM_1_T / GEMM_2_T): never the small-matrix fp32 path, e.g.
    Frame::UpdatePoseMatrices' mOw = -mRcw.t()*mtcw and SearchForTriangulation's R12 = R1w*R2w.t()."""
    rng = np.random.default_rng(11)
    differs_from_fp32 = 0
    for _ in range(300):
        A, B, v = (rng.normal(size=(3, 3)).astype(np.float32), rng.normal(size=(3, 3)).astype(np.float32),
                   rng.normal(size=(3, 1)).astype(np.float32))
        out = np.zeros((3, 3), np.float32)
        stub.stub_gemm_t(_p(A), 3, 3, _p(B), 3, 3, 1, 0, _p(out))
        assert np.array_equal(out, cv2.gemm(A, B, 1, None, 0, flags=cv2.GEMM_1_T))
        stub.stub_gemm_t(_p(A), 3, 3, _p(B), 3, 3, 2, 0, _p(out))
        want = cv2.gemm(A, B, 1, None, 0, flags=cv2.GEMM_2_T)
        assert np.array_equal(out, want)
        differs_from_fp32 += not np.array_equal(want, cv2.gemm(A, np.ascontiguousarray(B.T), 1, None, 0))
        stub.stub_gemm_t(_p(A), 3, 3, _p(B), 3, 3, 3, 0, _p(out))
        assert np.array_equal(out, cv2.gemm(A, B, 1, None, 0, flags=cv2.GEMM_1_T | cv2.GEMM_2_T))
        o3 = np.zeros((3, 1), np.float32)
        stub.stub_gemm_t(_p(A), 3, 3, _p(v), 3, 1, 1, 1, _p(o3))
        assert np.array_equal(o3, cv2.gemm(A, v, -1, None, 0, flags=cv2.GEMM_1_T))
        stub.stub_gemm_t(_p(A), 3, 3, _p(B), 3, 3, 2, 1, _p(out))
        assert np.array_equal(out, cv2.gemm(A, B, -1, None, 0, flags=cv2.GEMM_2_T))
    assert differs_from_fp32 > 100      # the distinction matters: most random 3x3 products differ in some element


def test_invert_norm_dot_equal_cv2(stub):
    rng = np.random.default_rng(5)
    for _ in range(300):
        A = rng.normal(size=(3, 3)).astype(np.float32)
        I = np.zeros((3, 3), np.float32)
        stub.stub_inv3(_p(A), _p(I))
        assert np.array_equal(I, cv2.invert(A)[1])
        v, w = rng.normal(size=3).astype(np.float32) * 7, rng.normal(size=3).astype(np.float32)
        assert stub.stub_norm(_p(v), 3) == cv2.norm(v.reshape(3, 1))
        assert stub.stub_dot(_p(v), _p(w), 3) == float(np.dot(v.astype(np.float64), w.astype(np.float64))) or \
            abs(stub.stub_dot(_p(v), _p(w), 3) - v.reshape(1, 3).astype(np.float64) @ w.astype(np.float64)) < 1e-15
