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
