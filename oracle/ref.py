"""ctypes view of oracle/_ref/: the reference's OWN hot-path translation units, compiled unmodified against the OpenCV
stand-in of oracle/ref_stub/ (oracle/Makefile, target `_ref`).  TEST INFRASTRUCTURE ONLY: its one job is to pin the
oracle restatement (oracle/libork.so) to reference source.  Only tests/ and tools/ import it; the product never does.

The libraries are built in this container, where /root/reference is mounted, and travel to the GPU box as prebuilt
files; nothing here reads /root/reference at run time.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("ORBX_REFERENCE_ROOT", "/root/reference")
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4")])
_LIBS = {}


def build():
    """Compile oracle/_ref from the reference sources where they lie (no-op when the tree is not mounted)."""
    if not os.path.isdir(os.path.join(REF_ROOT, "src")):
        return False
    subprocess.check_call(["make", "-C", _HERE, "-s", "_ref", "REF=" + REF_ROOT])
    return True


def available(name="libref_extractor.so"):
    return os.path.exists(os.path.join(_HERE, "_ref", name))


def _lib(name):
    if name not in _LIBS:
        path = os.path.join(_HERE, "_ref", name)
        if not os.path.exists(path):
            build()
        _LIBS[name] = C.CDLL(path)
    return _LIBS[name]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Extractor:
    """ORB_SLAM3::ORBextractor of the reference (src/ORBextractor.cc, unmodified).

    variant: "bump"   monotonic operator new: the quadtree's pointer tie-break (src/ORBextractor.cc:682) becomes
                      "later-created node = larger address", the rule the oracle and the device follow
             "malloc" glibc's allocator decides, as in a reference binary (history dependent)
             "nofma"  like bump, reference TU compiled -O2 -ffp-contract=off instead of the reference's -O3 defaults
    """
    _SO = {"bump": "libref_extractor.so", "malloc": "libref_extractor_malloc.so", "nofma": "libref_extractor_nofma.so"}

    def __init__(self, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7, variant="bump"):
        L = self.L = _lib(self._SO[variant])
        L.ref_extractor_create.restype = C.c_void_p
        L.ref_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.ref_extractor_destroy.argtypes = [C.c_void_p]
        L.ref_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_int, C.c_void_p]
        L.ref_extractor_tables.argtypes = [C.c_void_p] * 5
        L.ref_pyramid_level.argtypes = [C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                        C.c_int, C.c_void_p, C.c_void_p]
        L.ref_features_per_level.argtypes = [C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
        assert L.ref_alloc_mode() == (0 if variant == "malloc" else 1)
        self.nfeatures, self.nlevels, self.scale_factor = nfeatures, nlevels, scale
        self.h = L.ref_extractor_create(nfeatures, scale, nlevels, ini_th, min_th)
        t = [np.empty(nlevels, np.float32) for _ in range(4)]
        L.ref_extractor_tables(self.h, *[_p(a) for a in t])
        self.scale, self.inv_scale, self.sigma2, self.inv_sigma2 = t
        nf, um = np.empty(nlevels, np.int32), np.empty(16, np.int32)
        L.ref_features_per_level(nfeatures, scale, nlevels, _p(nf), _p(um))
        self.features_per_level, self.umax = nf, um

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_extractor_destroy(self.h)
            self.h = None

    def __call__(self, img, lap=(0, 0)):
        """-> (status, keypoints[KP_DTYPE], desc[n,32] u8, monoIndex): the oracle wrapper's convention"""
        if img is None or img.size == 0:
            m = C.c_int(0)
            n = self.L.ref_extract(self.h, None, 0, 0, 0, lap[0], lap[1], None, None, 0, C.byref(m))
            return (-1 if n < 0 else 0), np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8), 0
        img = np.ascontiguousarray(img, np.uint8)
        cap = self.nfeatures + 64 * self.nlevels + 512
        kps, desc, mono = np.empty(cap, KP_DTYPE), np.empty((cap, 32), np.uint8), C.c_int(0)
        n = self.L.ref_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], lap[0], lap[1], _p(kps),
                               _p(desc), cap, C.byref(mono))
        assert 0 <= n <= cap
        return 0, kps[:n].copy(), desc[:n].copy(), mono.value

    def pyramid_level(self, img, level):
        img = np.ascontiguousarray(img, np.uint8)
        w, h = C.c_int(0), C.c_int(0)
        self.L.ref_pyramid_level(self.nlevels, self.scale_factor, _p(img), img.shape[1], img.shape[0], img.strides[0],
                                 level, None, 0, C.byref(w), C.byref(h))
        out = np.empty((h.value, w.value), np.uint8)
        self.L.ref_pyramid_level(self.nlevels, self.scale_factor, _p(img), img.shape[1], img.shape[0], img.strides[0],
                                 level, _p(out), w.value, C.byref(w), C.byref(h))
        return out


class Vocabulary:
    """ORBVocabulary of the reference = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> (unmodified)."""

    def __init__(self, path, binary=True):
        L = self.L = _lib("libref_dbow2.so")
        L.ref_voc_load.restype = C.c_void_p
        L.ref_voc_load.argtypes = [C.c_char_p, C.c_int]
        L.ref_voc_destroy.argtypes = [C.c_void_p]
        L.ref_voc_info.argtypes = [C.c_void_p] * 6
        L.ref_voc_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_voc_score.restype = C.c_double
        L.ref_voc_score.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_forb_distance.argtypes = [C.c_void_p, C.c_void_p]
        self.h = L.ref_voc_load(str(path).encode(), int(binary))
        if not self.h:
            raise RuntimeError("reference vocabulary: cannot load %s" % path)
        v = (C.c_int * 5)()
        L.ref_voc_info(self.h, *[C.byref(v, 4 * k) for k in range(5)])
        self.k, self.L_, self.n_words, self.scoring, self.weighting = [int(x) for x in v]

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_voc_destroy(self.h)
            self.h = None

    def transform(self, desc, levelsup=4):
        """-> dict(bow_word, bow_value, fv_node, fv_off, fv_idx) in the oracle wrapper's layout"""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        cap = max(n, 1)
        bw, bv = np.zeros(cap, np.uint32), np.zeros(cap, np.float64)
        fn, fo, fi = np.zeros(cap, np.uint32), np.zeros(cap + 1, np.int32), np.zeros(cap, np.uint32)
        nb, nn = C.c_int(0), C.c_int(0)
        rc = self.L.ref_voc_transform(self.h, _p(desc), n, levelsup, _p(bw), _p(bv), cap, C.byref(nb), _p(fn), _p(fo), cap,
                                      C.byref(nn), _p(fi), cap)
        assert rc == 0
        nb, nn = nb.value, nn.value
        return dict(bow_word=bw[:nb].astype(np.int32), bow_value=bv[:nb].copy(), fv_node=fn[:nn].astype(np.int32),
                    fv_off=fo[:nn + 1].copy(), fv_idx=fi[:fo[nn]].astype(np.int32))

    def score(self, a, b):
        ia, va = np.ascontiguousarray(a[0], np.uint32), np.ascontiguousarray(a[1], np.float64)
        ib, vb = np.ascontiguousarray(b[0], np.uint32), np.ascontiguousarray(b[1], np.float64)
        return self.L.ref_voc_score(self.h, _p(ia), _p(va), len(ia), _p(ib), _p(vb), len(ib))

    def distance(self, a, b):
        a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
        return self.L.ref_forb_distance(_p(a), _p(b))


# ---------------------------------------------------------------------------------------------------------------------
# ORBmatcher of the reference (src/ORBmatcher.cc, unmodified) behind the flat argument lists of the oracle / C ABI.
# `F`, `Cur`, `KF1` ... are orbx.Frame objects (POD mirrors of orbx_frame_desc), `cam` an orbx camera struct.
# ---------------------------------------------------------------------------------------------------------------------
def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dt)


def _pp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _ml():
    return _lib("libref_matcher.so")


def descriptor_distance(a, b):
    a, b = _c(a, np.uint8), _c(b, np.uint8)
    f = _ml().ref_descriptor_distance
    f.argtypes = [C.c_void_p, C.c_void_p]
    return f(_pp(a), _pp(b))


def three_maxima(counts):
    counts = _c(counts, np.int32)
    out = np.zeros(3, np.int32)
    f = _ml().ref_three_maxima
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    f(_pp(counts), len(counts), _pp(out))
    return tuple(int(v) for v in out)


def search_by_projection_map(F, kp_blocked, projX, projY, projXR, level, viewCos, mpDesc, flags, th, nnratio, scaleFactors):
    """-> (nmatches, cur_mp[n]): final F.mvpMapPoints as MapPoint index (-1 none, -2 a keypoint blocked at entry)"""
    nq = len(projX)
    a = [_c(kp_blocked, np.uint8), _c(projX, np.float32), _c(projY, np.float32), _c(projXR, np.float32), _c(level, np.int32),
         _c(viewCos, np.float32), _c(mpDesc, np.uint8), _c(flags, np.uint8)]
    if a[0] is None:
        a[0] = np.zeros(F.n, np.uint8)
    if a[3] is None:
        a[3] = np.zeros(nq, np.float32)
    sf = _c(scaleFactors, np.float32)
    cur = np.full(max(F.n, 1), -1, np.int32)
    nm = C.c_int(0)
    f = _ml().ref_search_by_projection_map
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 7 + [C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    f(F.ref(), _pp(a[0]), nq, *[_pp(v) for v in a[1:]], float(th), float(nnratio), _pp(sf), len(sf), _pp(cur), C.byref(nm))
    return nm.value, cur[:F.n]


def search_by_projection_frame(Cur, cur_blocked, cam, Tcw_cur, Tcw_last, flags, xw, octave, angle, mpDesc, th, bMono, checkOri,
                               scaleFactors):
    """-> (nmatches, cur_match[n]): final CurrentFrame.mvpMapPoints as last-frame index (-1 none, -2 blocked at entry)"""
    nq = len(flags)
    blk = _c(cur_blocked, np.uint8)
    if blk is None:
        blk = np.zeros(Cur.n, np.uint8)
    a = [_c(Tcw_cur, np.float32), _c(Tcw_last, np.float32)]
    b = [_c(flags, np.uint8), _c(xw, np.float32), _c(octave, np.int32), _c(angle, np.float32), _c(mpDesc, np.uint8)]
    sf = _c(scaleFactors, np.float32)
    cur = np.full(max(Cur.n, 1), -1, np.int32)
    nm = C.c_int(0)
    f = _ml().ref_search_by_projection_frame
    f.argtypes = [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 5 + [C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 2
    f(Cur.ref(), _pp(blk), C.byref(cam), _pp(a[0]), _pp(a[1]), nq, *[_pp(v) for v in b], float(th), int(bMono), int(checkOri),
      _pp(sf), len(sf), _pp(cur), C.byref(nm))
    return nm.value, cur[:Cur.n]


def search_for_triangulation(KF1, KF2, has1, has2, fv1, fv2, cam1, cam2, R1w, t1w, R2w, t2w, sigma2, scaleFactors,
                             bOnlyStereo=False, bCoarse=False, checkOri=True):
    m12 = np.full(max(KF1.n, 1), -1, np.int32)
    nm = C.c_int(0)
    f1 = [_c(v, np.int32) for v in fv1]
    f2 = [_c(v, np.int32) for v in fv2]
    a = [_c(has1, np.uint8), _c(has2, np.uint8)]
    g = [_c(R1w, np.float32), _c(t1w, np.float32), _c(R2w, np.float32), _c(t2w, np.float32), _c(sigma2, np.float32),
         _c(scaleFactors, np.float32)]
    f = _ml().ref_search_for_triangulation
    f.argtypes = [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_void_p] * 8 + [C.c_int] * 4 + \
                 [C.c_void_p, C.c_void_p]
    f(KF1.ref(), KF2.ref(), _pp(a[0]), _pp(a[1]), len(f1[0]), _pp(f1[0]), _pp(f1[1]), _pp(f1[2]), len(f2[0]), _pp(f2[0]), _pp(f2[1]),
      _pp(f2[2]), C.byref(cam1), C.byref(cam2), *[_pp(v) for v in g], len(g[5]), int(bOnlyStereo), int(bCoarse), int(checkOri),
      _pp(m12), C.byref(nm))
    return nm.value, m12[:KF1.n]


def search_by_bow(kf, frame, kf_has_mp, fv_kf, fv_f, nnratio=0.7, check_orientation=True):
    has = np.ascontiguousarray(kf_has_mp, np.uint8)
    kn, ko, ki = [np.ascontiguousarray(a, np.int32) for a in fv_kf]
    fn, fo, fi = [np.ascontiguousarray(a, np.int32) for a in fv_f]
    out = np.full(max(frame.n, 1), -1, np.int32)
    nm = C.c_int32(0)
    f = _ml().ref_search_by_bow
    f.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    f(kf.ref(), frame.ref(), _p(has), len(kn), _p(kn), _p(ko), _p(ki), len(fn), _p(fn), _p(fo), _p(fi), nnratio,
      int(check_orientation), _p(out), C.byref(nm))
    return nm.value, out[:frame.n]


def fuse(kf, cam, Rcw, tcw, Ow, flags, xw, max_dist, min_dist, normal, mp_desc, th, scale_factors, inv_level_sigma2, log_scale_factor):
    f32 = lambda a: np.ascontiguousarray(a, np.float32)   # noqa: E731
    Rcw, tcw, Ow, xw, max_dist, min_dist, normal = map(f32, (Rcw, tcw, Ow, xw, max_dist, min_dist, normal))
    sf, isg = f32(scale_factors), f32(inv_level_sigma2)
    flags = np.ascontiguousarray(flags, np.uint8)
    mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
    n = len(flags)
    out = np.full(max(n, 1), -1, np.int32)
    nf = C.c_int32(0)
    f = _ml().ref_fuse
    f.argtypes = [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 6 + [C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    f(kf.ref(), C.byref(cam), _p(Rcw), _p(tcw), _p(Ow), n, _p(flags), _p(xw), _p(max_dist), _p(min_dist), _p(normal), _p(mp_desc), th,
      _p(sf), _p(isg), len(sf), log_scale_factor, _p(out), C.byref(nf))
    return nf.value, out[:n]


# ------------------------------------------------------------------------------------------------------------------
# ORB_SLAM3::Frame of the reference (src/Frame.cc + include/Frame.h, unmodified; oracle/ref_stub/frame_prelude.h)
# ------------------------------------------------------------------------------------------------------------------
_KP = np.dtype([("x", np.float32), ("y", np.float32), ("size", np.float32), ("angle", np.float32), ("response", np.float32),
                ("octave", np.int32)])


class Frame:
    """A real reference Frame built by its stereo (imgR given) or monocular constructor: extraction by the reference's
    ORBextractor on two threads, UndistortKeyPoints, ComputeStereoMatches, AssignFeaturesToGrid."""

    def __init__(self, imgL, imgR, cam, bf, th_depth=40.0, dist=(0, 0, 0, 0), nfeatures=1000, scale=1.2, nlevels=8, ini_th=20,
                 min_th=7, Tcb=None, fma=False):
        """fma=True: the TU built with the reference's own Release flags (-O3, -ffp-contract=fast) instead of one rounding
        per operation"""
        L = self.L = _lib("libref_frame_fma.so" if fma else "libref_frame.so")
        _ml()                                             # make sure libref_matcher.so is built
        L.ref_frame_init.argtypes = [C.c_char_p]
        assert L.ref_frame_init(os.path.join(_HERE, "_ref", "libref_matcher.so").encode()) == 0
        L.ref_frame_create.restype = C.c_void_p
        L.ref_frame_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                       C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
        for f in ("ref_frame_destroy", "ref_frame_n", "ref_frame_n_right"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.ref_frame_mb.restype = C.c_float
        L.ref_frame_mb.argtypes = [C.c_void_p]
        L.ref_frame_log_scale_factor.restype = C.c_float
        L.ref_frame_log_scale_factor.argtypes = [C.c_void_p]
        L.ref_frame_bounds.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_frame_keys.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_frame_stereo_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_frame_features_in_area.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.ref_frame_set_pose.argtypes = [C.c_void_p] * 5
        L.ref_frame_set_imu_pose.argtypes = [C.c_void_p] * 4
        L.ref_frame_is_in_frustum.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_float] + [C.c_void_p] * 7
        imgL = np.ascontiguousarray(imgL, np.uint8)
        imgR = None if imgR is None else np.ascontiguousarray(imgR, np.uint8)
        h, w = imgL.shape
        d = np.ascontiguousarray(dist, np.float32)
        T = None if Tcb is None else np.ascontiguousarray(Tcb, np.float32)
        self.h = L.ref_frame_create(_p(imgL), None if imgR is None else _p(imgR), w, h, w, nfeatures, scale, nlevels, ini_th, min_th,
                                    cam.fx, cam.fy, cam.cx, cam.cy, _p(d), bf, th_depth, None if T is None else _p(T))
        self.n, self.n_right = L.ref_frame_n(self.h), L.ref_frame_n_right(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.ref_frame_destroy(self.h)
            self.h = None

    __del__ = close

    @property
    def mb(self):
        return float(self.L.ref_frame_mb(self.h))

    @property
    def log_scale_factor(self):
        return float(self.L.ref_frame_log_scale_factor(self.h))

    @property
    def bounds(self):
        b = np.zeros(4, np.float32)
        self.L.ref_frame_bounds(self.h, _p(b))
        return b                                          # minX, minY, maxX, maxY

    def keys(self, which=0):
        """which: 0 mvKeys (+ mDescriptors), 1 mvKeysRight (+ mDescriptorsRight), 2 mvKeysUn"""
        n = self.n_right if which == 1 else self.n
        k = np.zeros(max(n, 1), _KP)
        d = np.zeros((max(n, 1), 32), np.uint8)
        self.L.ref_frame_keys(self.h, which, _p(k), _p(d) if which != 2 else None)
        return k[:n], d[:n]

    def stereo_matches(self):
        ur, dp = np.zeros(max(self.n, 1), np.float32), np.zeros(max(self.n, 1), np.float32)
        self.L.ref_frame_stereo_matches(self.h, _p(ur), _p(dp))
        return ur[:self.n], dp[:self.n]

    def features_in_area(self, x, y, r, min_level=-1, max_level=-1, cap=4096):
        out = np.zeros(cap, np.int32)
        n = self.L.ref_frame_features_in_area(self.h, float(x), float(y), float(r), int(min_level), int(max_level), _p(out), cap)
        assert n <= cap
        return out[:n]

    def set_pose(self, Tcw):
        T = np.ascontiguousarray(Tcw, np.float32)
        Ow, R, t = np.zeros(3, np.float32), np.zeros(9, np.float32), np.zeros(3, np.float32)
        self.L.ref_frame_set_pose(self.h, _p(T), _p(Ow), _p(R), _p(t))
        return Ow, R.reshape(3, 3), t

    def set_imu_pose(self, Rwb, twb):
        R, t = np.ascontiguousarray(Rwb, np.float32), np.ascontiguousarray(twb, np.float32)
        T = np.zeros(16, np.float32)
        self.L.ref_frame_set_imu_pose(self.h, _p(R), _p(t), _p(T))
        return T.reshape(4, 4)

    def is_in_frustum(self, xw, max_dist, min_dist, normal, cos_limit=0.5):
        n = len(max_dist)
        f32 = lambda a: np.ascontiguousarray(a, np.float32)   # noqa: E731
        xw, max_dist, min_dist, normal = map(f32, (xw, max_dist, min_dist, normal))
        out = dict(in_view=np.zeros(n, np.uint8), proj_x=np.zeros(n, np.float32), proj_y=np.zeros(n, np.float32),
                   proj_xr=np.zeros(n, np.float32), depth=np.zeros(n, np.float32), level=np.zeros(n, np.int32),
                   view_cos=np.zeros(n, np.float32))
        out["n"] = self.L.ref_frame_is_in_frustum(self.h, n, _p(xw), _p(max_dist), _p(min_dist), _p(normal), cos_limit,
                                                  _p(out["in_view"]), _p(out["proj_x"]), _p(out["proj_y"]), _p(out["proj_xr"]),
                                                  _p(out["depth"]), _p(out["level"]), _p(out["view_cos"]))
        return out
