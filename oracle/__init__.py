"""ctypes view of the CPU ORACLE (oracle/libork.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package; the product (awesome-orb-slam3-3dvisioncraft-version_b200/) never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4")])


def build(force=False):
    so = os.path.join(_HERE, "libork.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h", ".inc"))]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(map(os.path.getmtime, srcs)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libork.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libork.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        L = _LIB
        L.ork_extractor_create.restype = C.c_void_p
        L.ork_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.ork_extractor_destroy.argtypes = [C.c_void_p]
        L.ork_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ork_extractor_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.ork_extractor_umax.argtypes = [C.c_void_p, C.c_void_p]
        L.ork_pyramid_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ork_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ork_resize_linear_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                           C.c_int, C.c_int]
        L.ork_fast9_16.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                   C.c_int]
        L.ork_gaussian_blur7.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.ork_fast_atan2.restype = C.c_float
        L.ork_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.ork_fast_atan2_array.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ork_distribute_octree.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int]
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().ork_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def fast(img, threshold, nms=True):
    """-> int32 array [n,3] of (x, y, score) in raster order (cv::FAST semantics)."""
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size
    out = np.empty((max(cap, 1), 3), np.int32)
    n = lib().ork_fast9_16(_p(img), img.shape[1], img.shape[0], img.strides[0], threshold, int(nms),
                           _p(out), cap)
    return out[:n].copy()


def gaussian_blur7(img):
    img = np.ascontiguousarray(img, np.uint8)
    dst = np.empty_like(img)
    lib().ork_gaussian_blur7(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(dst), dst.strides[0])
    return dst


def fast_atan2(y, x):
    y = np.ascontiguousarray(y, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(y)
    lib().ork_fast_atan2_array(_p(y), _p(x), _p(out), y.size)
    return out


def distribute_octree(x, y, resp, minX, maxX, minY, maxY, N):
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    resp = np.ascontiguousarray(resp, np.float32)
    cap = max(len(x), 1)
    ox, oy, orr = (np.empty(cap, np.float32) for _ in range(3))
    n = lib().ork_distribute_octree(_p(x), _p(y), _p(resp), len(x), minX, maxX, minY, maxY, N, _p(ox),
                                    _p(oy), _p(orr), cap)
    return ox[:n].copy(), oy[:n].copy(), orr[:n].copy()


class Extractor:
    """Oracle counterpart of ORBextractor (src/ORBextractor.cc)."""

    def __init__(self, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7):
        self.h = lib().ork_extractor_create(nfeatures, scale, nlevels, ini_th, min_th)
        if not self.h:
            raise ValueError("bad extractor parameters")
        self.nfeatures, self.nlevels = nfeatures, nlevels
        t = [np.empty(nlevels, np.float32) for _ in range(4)]
        nf = np.empty(nlevels, np.int32)
        lib().ork_extractor_tables(self.h, _p(t[0]), _p(t[1]), _p(t[2]), _p(t[3]), _p(nf))
        self.scale, self.inv_scale, self.sigma2, self.inv_sigma2 = t
        self.features_per_level = nf
        um = np.empty(16, np.int32)
        lib().ork_extractor_umax(self.h, _p(um))
        self.umax = um

    def __del__(self):
        if getattr(self, "h", None):
            lib().ork_extractor_destroy(self.h)
            self.h = None

    def __call__(self, img, lap=(0, 0)):
        """-> (status, keypoints[KP_DTYPE], desc[n,32] u8, monoIndex)"""
        if img is None or img.size == 0:
            return -1, np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8), 0
        img = np.ascontiguousarray(img, np.uint8)
        cap = self.nfeatures + 64 * self.nlevels
        kps = np.empty(cap, KP_DTYPE)
        desc = np.empty((cap, 32), np.uint8)
        n = C.c_int(0)
        mono = C.c_int(0)
        rc = lib().ork_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], lap[0], lap[1],
                               _p(kps), _p(desc), cap, C.byref(n), C.byref(mono))
        return rc, kps[:n.value].copy(), desc[:n.value].copy(), mono.value

    def pyramid_level(self, level):
        w, h = C.c_int(0), C.c_int(0)
        lib().ork_pyramid_level(self.h, level, None, 0, C.byref(w), C.byref(h))
        out = np.empty((h.value, w.value), np.uint8)
        lib().ork_pyramid_level(self.h, level, _p(out), w.value, C.byref(w), C.byref(h))
        return out

    def candidates(self, level):
        n = C.c_int(0)
        cap = 1 << 20
        xy = np.empty((cap, 2), np.int16)
        sc = np.empty(cap, np.uint8)
        rc = lib().ork_candidates(self.h, level, _p(xy), _p(sc), cap, C.byref(n))
        assert rc == 0
        return xy[:n.value].copy(), sc[:n.value].copy()
