"""ctypes mirrors of the POD structs in include/orbx.h (shared by the product binding and, for the
struct layouts only, by the test-side oracle binding)."""
import ctypes as C
import numpy as np


class FrameDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("kps", C.c_void_p), ("desc", C.c_void_p), ("uright", C.c_void_p),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float)]


class Camera(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("bf", C.c_float), ("b", C.c_float)]


class TriProblem(C.Structure):
    """orbx_tri_problem (include/orbx.h): one SearchForTriangulation call of a prepared batch."""
    _fields_ = [("kf1", C.c_void_p), ("kf2", C.c_void_p), ("has_mp1", C.c_void_p), ("has_mp2", C.c_void_p),
                ("nn1", C.c_int32), ("fv1_node", C.c_void_p), ("fv1_off", C.c_void_p), ("fv1_idx", C.c_void_p),
                ("nn2", C.c_int32), ("fv2_node", C.c_void_p), ("fv2_off", C.c_void_p), ("fv2_idx", C.c_void_p),
                ("cam1", C.c_void_p), ("cam2", C.c_void_p), ("R1w", C.c_void_p), ("t1w", C.c_void_p), ("R2w", C.c_void_p),
                ("t2w", C.c_void_p), ("only_stereo", C.c_int32), ("coarse", C.c_int32), ("match12", C.c_void_p),
                ("nmatches", C.c_int32)]


class LbaProblem(C.Structure):
    """orbx_lba_problem (include/orbx.h): one LocalBundleAdjustment of a prepared batch."""
    _fields_ = [("n_kf", C.c_int32), ("n_mp", C.c_int32), ("n_edges", C.c_int32), ("kf_Tcw", C.c_void_p), ("kf_fixed", C.c_void_p),
                ("mp_xyz", C.c_void_p), ("e_kf", C.c_void_p), ("e_mp", C.c_void_p), ("e_obs", C.c_void_p), ("e_inv_sigma2", C.c_void_p),
                ("lambda_init", C.c_double), ("edge_bad", C.c_void_p), ("iters", C.c_int32 * 2), ("status", C.c_int32)]


class TrackMap(C.Structure):
    """orbx_track_map (include/orbx.h): the flat local map of every stream for one step."""
    _fields_ = [("m_cap", C.c_int32), ("n_map", C.c_void_p), ("xw", C.c_void_p), ("desc", C.c_void_p), ("last_flags", C.c_void_p),
                ("last_octave", C.c_void_p), ("last_angle", C.c_void_p), ("map_flags", C.c_void_p), ("max_dist", C.c_void_p),
                ("min_dist", C.c_void_p), ("normal", C.c_void_p), ("log_scale_factor", C.c_float)]

    FIELDS = (("n_map", np.int32), ("xw", np.float32), ("desc", np.uint8), ("last_flags", np.uint8), ("last_octave", np.int32),
              ("last_angle", np.float32), ("map_flags", np.uint8), ("max_dist", np.float32), ("min_dist", np.float32),
              ("normal", np.float32))


class TrackImu(C.Structure):
    """orbx_track_imu (include/orbx.h): inputs of the visual-inertial second pose optimisation of a tracker step."""
    _fields_ = [("mode", C.c_int32), ("rec_init", C.c_int32), ("Tcb", C.c_void_p), ("Tbc", C.c_void_p), ("velocity", C.c_void_p),
                ("bias", C.c_void_p), ("ref_state", C.c_void_p), ("preint", C.c_void_p), ("preint_jac", C.c_void_p),
                ("preint_bias", C.c_void_p), ("info_inertial", C.c_void_p), ("info_gyro", C.c_void_p), ("info_acc", C.c_void_p),
                ("prior_state", C.c_void_p), ("prior_H", C.c_void_p)]

    # (name, dtype, trailing shape after the leading [S]; None = shared by the rig)
    FIELDS = (("Tcb", np.float32, None), ("Tbc", np.float32, None), ("velocity", np.float32, (3,)), ("bias", np.float32, (6,)),
              ("ref_state", np.float64, (21,)), ("preint", np.float64, (16,)), ("preint_jac", np.float64, (45,)),
              ("preint_bias", np.float64, (6,)), ("info_inertial", np.float64, (81,)), ("info_gyro", np.float64, (9,)),
              ("info_acc", np.float64, (9,)), ("prior_state", np.float64, (21,)), ("prior_H", np.float64, (225,)))


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def c32(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype)


class Frame:
    """Flat view of a Frame / KeyFrame (mvKeysUn, mDescriptors, mvuRight, image bounds)."""

    def __init__(self, kps, desc, uright=None, bounds=(0.0, 0.0, 752.0, 480.0)):
        self.kps = np.ascontiguousarray(kps)
        self.desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        self.uright = None if uright is None else np.ascontiguousarray(uright, np.float32)
        self.bounds = tuple(float(b) for b in bounds)
        assert len(self.kps) == len(self.desc)
        self.n = len(self.kps)
        self.c = FrameDesc(self.n, self.kps.ctypes.data, self.desc.ctypes.data,
                           None if self.uright is None else self.uright.ctypes.data, *self.bounds)

    def ref(self):
        return C.byref(self.c)


def make_camera(fx=458.654, fy=457.296, cx=367.215, cy=248.375, bf=47.9):
    return Camera(fx, fy, cx, cy, bf, bf / fx)
