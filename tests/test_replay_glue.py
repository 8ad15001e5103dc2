"""The float32 glue of the inertial tracking chain (tests/replay_reference.py: imu_state_from_pose, pose_from_imu_state)
against the cv::Mat expressions it restates (src/Frame.cc:520-554), evaluated with cv2's own gemm."""
import numpy as np
import pytest

import scenarios as sc
from replay_reference import imu_state_from_pose, pose_from_imu_state

cv2 = pytest.importorskip("cv2")
F32 = np.float32


def _rig(rng):
    T = sc.se3_matrix(sc.rot_small(rng, 20.0), rng.uniform(-1, 1, 3)).astype(F32)
    Tbc = sc.se3_matrix(sc.rot_small(rng, 80.0), rng.uniform(-0.1, 0.1, 3))
    return T, np.linalg.inv(Tbc).astype(F32)


def test_imu_state_from_pose_is_cv_gemm():
    rng = np.random.default_rng(1)
    for _ in range(200):
        T, Tcb = _rig(rng)
        Rcw, tcw = np.ascontiguousarray(T[:3, :3]), np.ascontiguousarray(T[:3, 3:4])
        Rwc = np.ascontiguousarray(Rcw.T)
        Ow = cv2.gemm(Rcw, tcw, -1.0, None, 0.0, flags=cv2.GEMM_1_T)                       # mOw = -mRcw.t()*mtcw
        twb = cv2.gemm(Rwc, np.ascontiguousarray(Tcb[:3, 3:4]), 1.0, Ow, 1.0)               # mRwc*tcb + mOw
        Rwb = cv2.gemm(Rwc, np.ascontiguousarray(Tcb[:3, :3]), 1.0, None, 0.0)              # mRwc*Rcb
        st = imu_state_from_pose(T, Tcb, np.zeros(3), np.zeros(6))
        assert np.array_equal(st[:9].astype(F32).reshape(3, 3), Rwb)
        assert np.array_equal(st[9:12].astype(F32), twb.ravel())


def test_pose_from_imu_state_is_cv_gemm():
    rng = np.random.default_rng(2)
    for _ in range(200):
        T, Tcb = _rig(rng)
        Rwb = sc.rot_small(rng, 30.0)
        twb = rng.uniform(-2, 2, 3)
        state = np.concatenate([Rwb.ravel(), twb, np.zeros(9)])
        Rbw = np.ascontiguousarray(Rwb.astype(F32).T)
        tbw = cv2.gemm(Rbw, twb.astype(F32).reshape(3, 1), -1.0, None, 0.0)                 # -Rbw*twb
        Tbw = np.eye(4, dtype=F32)
        Tbw[:3, :3] = Rbw
        Tbw[:3, 3] = tbw.ravel()
        want = cv2.gemm(Tcb, Tbw, 1.0, None, 0.0)                                           # mImuCalib.Tcb*Tbw
        assert np.array_equal(pose_from_imu_state(state, Tcb), want)
