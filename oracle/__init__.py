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
        L.ork_search_by_bow.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
        L.ork_fuse.argtypes = [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 6 + [C.c_float, C.c_void_p, C.c_void_p, C.c_int,
                                                                                  C.c_float, C.c_void_p, C.c_void_p]
        L.ork_is_in_frustum.argtypes = [C.c_void_p] * 4 + [C.c_float] * 5 + [C.c_int, C.c_float, C.c_int] + [C.c_void_p] * 11
        L.ork_undistort_points.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ork_pose_inertial_opt_last_kf.argtypes = [C.c_int] + [C.c_void_p] * 14 + [C.c_int] + [C.c_void_p] * 4
        L.ork_pose_inertial_opt_last_frame.argtypes = [C.c_int] + [C.c_void_p] * 18 + [C.c_int] + [C.c_void_p] * 4
        L.ork_inertial_debug.argtypes = [C.c_void_p] * 5
        L.ork_voc_from_memory.restype = C.c_void_p
        L.ork_voc_from_memory.argtypes = [C.c_void_p, C.c_size_t]
        L.ork_voc_load.restype = C.c_void_p
        L.ork_voc_load.argtypes = [C.c_char_p]
        L.ork_voc_destroy.argtypes = [C.c_void_p]
        L.ork_voc_info.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.ork_voc_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 9
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


# ------------------------------------------------------------------------------------------------
# matchers (oracle/ork_matcher.cpp).  `Frame` / `Camera` are the POD mirrors of include/orbx.h.
# ------------------------------------------------------------------------------------------------
def _frame_types():
    import sys
    pkg = os.path.join(os.path.dirname(_HERE), "awesome-orb-slam3-3dvisioncraft-version_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    from orbx import abi
    return abi


def _c(a, dt):
    return None if a is None else np.ascontiguousarray(a, dt)


def _pp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def descriptor_distance(a, b):
    a, b = _c(a, np.uint8), _c(b, np.uint8)
    f = lib().ork_descriptor_distance
    f.argtypes = [C.c_void_p, C.c_void_p]
    return f(_pp(a), _pp(b))


def features_in_area(F, x, y, r, minLevel, maxLevel, cap=256):
    nq = len(x)
    out = np.full((nq, cap), -1, np.int32)
    n = np.zeros(nq, np.int32)
    a = [_c(x, np.float32), _c(y, np.float32), _c(r, np.float32), _c(minLevel, np.int32), _c(maxLevel, np.int32)]
    f = lib().ork_features_in_area
    f.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p]
    f(F.ref(), nq, *[_pp(v) for v in a], _pp(out), cap, _pp(n))
    return out, n


def stereo_match(pyrL, pyrR, kpL, descL, kpR, descR, scale, inv_scale, bf, b):
    """pyrL/pyrR: lists of un-bordered uint8 level images."""
    L = len(pyrL)
    pl = [np.ascontiguousarray(p, np.uint8) for p in pyrL]
    pr = [np.ascontiguousarray(p, np.uint8) for p in pyrR]
    PL = (C.c_void_p * L)(*[p.ctypes.data for p in pl])
    PR = (C.c_void_p * L)(*[p.ctypes.data for p in pr])
    lw = np.array([p.shape[1] for p in pl], np.int32)
    lh = np.array([p.shape[0] for p in pl], np.int32)
    kpL, kpR = np.ascontiguousarray(kpL), np.ascontiguousarray(kpR)
    descL, descR = _c(descL, np.uint8), _c(descR, np.uint8)
    sc, isc = _c(scale, np.float32), _c(inv_scale, np.float32)
    ur = np.empty(len(kpL), np.float32)
    dp = np.empty(len(kpL), np.float32)
    f = lib().ork_stereo_match
    f.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float,
                                     C.c_float, C.c_void_p, C.c_void_p]
    f(PL, PR, _pp(lw), _pp(lh), _pp(kpL), _pp(descL), len(kpL), _pp(kpR), _pp(descR), len(kpR), _pp(sc), _pp(isc),
      float(bf), float(b), _pp(ur), _pp(dp))
    return ur, dp


def search_by_projection_map(F, kp_blocked, projX, projY, projXR, level, viewCos, mpDesc, flags, th, nnratio,
                             scaleFactors):
    nq = len(projX)
    best = np.full(nq, -1, np.int32)
    nm = C.c_int(0)
    a = [_c(kp_blocked, np.uint8), _c(projX, np.float32), _c(projY, np.float32), _c(projXR, np.float32),
         _c(level, np.int32), _c(viewCos, np.float32), _c(mpDesc, np.uint8), _c(flags, np.uint8)]
    if a[0] is None:
        a[0] = np.zeros(F.n, np.uint8)
    if a[3] is None:
        a[3] = np.zeros(nq, np.float32)
    sf = _c(scaleFactors, np.float32)
    f = lib().ork_search_by_projection_map
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 7 + [C.c_float, C.c_float, C.c_void_p, C.c_int,
                                                                        C.c_void_p, C.c_void_p]
    f(F.ref(), _pp(a[0]), nq, *[_pp(v) for v in a[1:]], float(th), float(nnratio), _pp(sf), len(sf), _pp(best),
      C.byref(nm))
    return nm.value, best


def search_by_projection_frame(Cur, cur_blocked, cam, Tcw_cur, Tcw_last, flags, xw, octave, angle, mpDesc, th, bMono,
                               checkOri, scaleFactors):
    nq = len(flags)
    match = np.full(nq, -1, np.int32)
    kept = np.zeros(nq, np.uint8)
    cur_match = np.full(Cur.n, -1, np.int32)
    nm = C.c_int(0)
    blk = _c(cur_blocked, np.uint8)
    if blk is None:
        blk = np.zeros(Cur.n, np.uint8)
    a = [_c(Tcw_cur, np.float32), _c(Tcw_last, np.float32)]
    b = [_c(flags, np.uint8), _c(xw, np.float32), _c(octave, np.int32), _c(angle, np.float32), _c(mpDesc, np.uint8)]
    sf = _c(scaleFactors, np.float32)
    f = lib().ork_search_by_projection_frame
    f.argtypes = [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 5 + [C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int] + \
                 [C.c_void_p] * 4
    f(Cur.ref(), _pp(blk), C.byref(cam), _pp(a[0]), _pp(a[1]), nq, *[_pp(v) for v in b], float(th), int(bMono),
      int(checkOri), _pp(sf), len(sf), _pp(match), _pp(kept), _pp(cur_match), C.byref(nm))
    return nm.value, match, kept, cur_match


def search_for_triangulation(KF1, KF2, has1, has2, fv1, fv2, cam1, cam2, R1w, t1w, R2w, t2w, sigma2, scaleFactors,
                             bOnlyStereo=False, bCoarse=False, checkOri=True):
    m12 = np.full(KF1.n, -1, np.int32)
    nm = C.c_int(0)
    f1 = [_c(v, np.int32) for v in fv1]
    f2 = [_c(v, np.int32) for v in fv2]
    a = [_c(has1, np.uint8), _c(has2, np.uint8)]
    g = [_c(R1w, np.float32), _c(t1w, np.float32), _c(R2w, np.float32), _c(t2w, np.float32), _c(sigma2, np.float32),
         _c(scaleFactors, np.float32)]
    f = lib().ork_search_for_triangulation
    f.argtypes = [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_void_p] * 8 + \
                 [C.c_int] * 4 + [C.c_void_p, C.c_void_p]
    f(KF1.ref(), KF2.ref(), _pp(a[0]), _pp(a[1]), len(f1[0]), _pp(f1[0]), _pp(f1[1]), _pp(f1[2]), len(f2[0]),
      _pp(f2[0]), _pp(f2[1]), _pp(f2[2]), C.byref(cam1), C.byref(cam2), *[_pp(v) for v in g], len(g[5]),
      int(bOnlyStereo), int(bCoarse), int(checkOri), _pp(m12), C.byref(nm))
    return nm.value, m12


# ------------------------------------------------------------------------------------------------
# optimisers (oracle/ork_optimizer.cpp)
# ------------------------------------------------------------------------------------------------
def pose_optimization(xw, obs, inv_sigma2, cam, Tcw):
    """-> (Tcw_out[4,4] f32, outlier[E] u8, nInliers, iters[4])"""
    xw, obs, inv_sigma2 = _c(xw, np.float32), _c(obs, np.float32), _c(inv_sigma2, np.float32)
    T = np.array(Tcw, np.float32).reshape(4, 4).copy()
    E = len(inv_sigma2)
    outl = np.zeros(max(E, 1), np.uint8)
    nin = C.c_int(0)
    iters = np.zeros(4, np.int32)
    f = lib().ork_pose_optimization
    f.argtypes = [C.c_int] + [C.c_void_p] * 8
    f(E, _pp(xw), _pp(obs), _pp(inv_sigma2), C.byref(cam), _pp(T), _pp(outl), C.byref(nin), _pp(iters))
    return T, outl[:E].copy(), nin.value, iters


def local_ba(kf_T, kf_fixed, mp_xyz, e_kf, e_mp, e_obs, e_inv_sigma2, cam, lambda_init=0.0, stop=None):
    """-> (kf_T_out, mp_xyz_out, edge_bad, iters[2], status)"""
    T = np.array(kf_T, np.float32).reshape(-1, 16).copy()
    X = np.array(mp_xyz, np.float32).reshape(-1, 3).copy()
    fixed = _c(kf_fixed, np.uint8)
    ekf, emp = _c(e_kf, np.int32), _c(e_mp, np.int32)
    obs, isg = _c(e_obs, np.float32), _c(e_inv_sigma2, np.float32)
    E = len(ekf)
    bad = np.zeros(max(E, 1), np.uint8)
    iters = np.zeros(2, np.int32)
    status = C.c_int(0)
    st = _c(stop, np.uint8) if stop is not None else None
    f = lib().ork_local_ba
    f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 5 + \
                 [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    f(len(T), _pp(T), _pp(fixed), len(X), _pp(X), E, _pp(ekf), _pp(emp), _pp(obs), _pp(isg), C.byref(cam),
      float(lambda_init), _pp(st), _pp(bad), _pp(iters), C.byref(status))
    return T.reshape(-1, 4, 4), X, bad[:E].copy(), iters, status.value


class Vocabulary:
    """Oracle DBoW2 vocabulary (oracle/ork_vocabulary.cpp)."""

    def __init__(self, source):
        L = lib()
        if isinstance(source, (bytes, bytearray, memoryview, np.ndarray)):
            buf = np.frombuffer(bytes(source), np.uint8)
            self.h = L.ork_voc_from_memory(_p(buf), buf.size)
        else:
            self.h = L.ork_voc_load(str(source).encode())
        if not self.h:
            raise RuntimeError("oracle vocabulary: cannot parse")
        v = (C.c_int * 6)()
        L.ork_voc_info(self.h, *[C.byref(v, 4 * k) for k in range(6)])
        self.k, self.L, self.n_nodes, self.n_words, self.scoring, self.weighting = [int(x) for x in v]

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().ork_voc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def transform(self, desc, levelsup=4):
        """-> dict(word_id, node_id per feature; bow_word, bow_value; fv_node, fv_off, fv_idx)"""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        wid, nid = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
        bw, bv = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float64)
        fn, fo, fi = np.zeros(max(n, 1), np.int32), np.zeros(n + 1, np.int32), np.zeros(max(n, 1), np.int32)
        nb, nn = C.c_int32(0), C.c_int32(0)
        rc = lib().ork_voc_transform(self.h, _p(desc), n, levelsup, _p(wid), _p(nid), _p(bw), _p(bv), C.byref(nb), _p(fn),
                                     _p(fo), _p(fi), C.byref(nn))
        assert rc == 0
        nb, nn = nb.value, nn.value
        return dict(word_id=wid[:n].copy(), node_id=nid[:n].copy(), bow_word=bw[:nb].copy(), bow_value=bv[:nb].copy(),
                    fv_node=fn[:nn].copy(), fv_off=fo[:nn + 1].copy(), fv_idx=fi[:fo[nn]].copy())


def search_by_bow(kf, frame, kf_has_mp, fv_kf, fv_f, nnratio=0.7, check_orientation=True):
    """oracle ORBmatcher::SearchByBoW(pKF, F, ...) -> (nmatches, match_f)"""
    has = np.ascontiguousarray(kf_has_mp, np.uint8)
    kn, ko, ki = [np.ascontiguousarray(a, np.int32) for a in fv_kf]
    fn, fo, fi = [np.ascontiguousarray(a, np.int32) for a in fv_f]
    out = np.full(max(frame.n, 1), -1, np.int32)
    nm = C.c_int32(0)
    rc = lib().ork_search_by_bow(kf.ref(), frame.ref(), _p(has), len(kn), _p(kn), _p(ko), _p(ki), len(fn), _p(fn), _p(fo), _p(fi),
                                 nnratio, int(check_orientation), _p(out), C.byref(nm))
    assert rc == 0
    return nm.value, out[:frame.n]


def fuse(kf, cam, Rcw, tcw, Ow, flags, xw, max_dist, min_dist, normal, mp_desc, th, scale_factors, inv_level_sigma2,
         log_scale_factor):
    """oracle search half of ORBmatcher::Fuse -> (nFused, best_idx)"""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)   # noqa: E731
    Rcw, tcw, Ow, xw, max_dist, min_dist, normal = map(f32, (Rcw, tcw, Ow, xw, max_dist, min_dist, normal))
    sf, isg = f32(scale_factors), f32(inv_level_sigma2)
    flags = np.ascontiguousarray(flags, np.uint8)
    mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
    n = len(flags)
    out = np.full(max(n, 1), -1, np.int32)
    nf = C.c_int32(0)
    rc = lib().ork_fuse(kf.ref(), C.byref(cam), _p(Rcw), _p(tcw), _p(Ow), n, _p(flags), _p(xw), _p(max_dist), _p(min_dist),
                        _p(normal), _p(mp_desc), th, _p(sf), _p(isg), len(sf), log_scale_factor, _p(out), C.byref(nf))
    assert rc == 0
    return nf.value, out[:n]


def is_in_frustum(cam, Rcw, tcw, Ow, bounds, cos_limit, nlevels, log_scale_factor, xw, max_dist, min_dist, normal, stale=None):
    """oracle Frame::isInFrustum batch -> dict of track fields + n"""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)   # noqa: E731
    Rcw, tcw, Ow, xw, max_dist, min_dist, normal = map(f32, (Rcw, tcw, Ow, xw, max_dist, min_dist, normal))
    n = len(max_dist)
    st = stale or {}
    out = dict(in_view=np.zeros(max(n, 1), np.uint8), proj_x=np.zeros(max(n, 1), np.float32), proj_y=np.zeros(max(n, 1), np.float32),
               proj_xr=f32(st.get("proj_xr", np.zeros(max(n, 1)))).copy(), depth=f32(st.get("depth", np.zeros(max(n, 1)))).copy(),
               level=np.ascontiguousarray(st.get("level", np.zeros(max(n, 1))), np.int32).copy(),
               view_cos=f32(st.get("view_cos", np.zeros(max(n, 1)))).copy())
    cnt = lib().ork_is_in_frustum(C.byref(cam), _p(Rcw), _p(tcw), _p(Ow), bounds[0], bounds[1], bounds[2], bounds[3], cos_limit,
                                  nlevels, log_scale_factor, n, _p(xw), _p(max_dist), _p(min_dist), _p(normal), _p(out["in_view"]),
                                  _p(out["proj_x"]), _p(out["proj_y"]), _p(out["proj_xr"]), _p(out["depth"]), _p(out["level"]),
                                  _p(out["view_cos"]))
    out = {k: v[:n] for k, v in out.items()}
    out["n"] = cnt
    return out


def undistort_points(xy, cam, dist_coef):
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    d = np.ascontiguousarray(dist_coef, np.float32).ravel()
    out = np.zeros_like(xy)
    rc = lib().ork_undistort_points(_p(xy), len(xy), C.byref(cam), _p(d), len(d), _p(out))
    assert rc == 0
    return out


def pose_inertial_optimization_last_keyframe(s, cam, rec_init=False):
    """oracle Optimizer::PoseInertialOptimizationLastKeyFrame on a scenario dict (tests/scenarios.inertial_scenario).
    -> dict(state[21], outlier[E], H[15,15], n, iters[4])"""
    E = len(s["isg"])
    f32 = lambda a: np.ascontiguousarray(a, np.float32)   # noqa: E731
    f64 = lambda a: np.ascontiguousarray(a, np.float64)   # noqa: E731
    xw, obs, isg, Tcw, Tcb, Tbc = map(f32, (s["xw"], s["obs"], s["isg"], s["Tcw"], s["Tcb"], s["Tbc"]))
    close = np.ascontiguousarray(s["close"], np.uint8)
    state, kf, pre, iI, iG, iA = map(f64, (np.array(s["state"]).copy(), s["kf"], s["preint"], s["infoI"], s["infoG"], s["infoA"]))
    outlier = np.zeros(max(E, 1), np.uint8)
    H = np.zeros(225, np.float64)
    n = C.c_int(0)
    iters = np.zeros(4, np.int32)
    rc = lib().ork_pose_inertial_opt_last_kf(E, _p(xw), _p(obs), _p(isg), _p(close), C.byref(cam), _p(Tcw), _p(Tcb), _p(Tbc), _p(state),
                                             _p(kf), _p(pre), _p(iI), _p(iG), _p(iA), int(rec_init), _p(outlier), _p(H), C.byref(n),
                                             _p(iters))
    assert rc == 0
    return dict(state=state, outlier=outlier[:E], H=H.reshape(15, 15), n=n.value, iters=iters)


def pose_inertial_optimization_last_frame(s, cam, rec_init=False):
    """oracle Optimizer::PoseInertialOptimizationLastFrame on tests/scenarios.inertial_lf_scenario.
    -> dict(state[21], outlier[E], H[15,15] (previous frame marginalised out), n, iters[4])"""
    E = len(s["isg"])
    f32 = lambda a: np.ascontiguousarray(a, np.float32)   # noqa: E731
    f64 = lambda a: np.ascontiguousarray(a, np.float64)   # noqa: E731
    xw, obs, isg, Tcw, Tcb, Tbc = map(f32, (s["xw"], s["obs"], s["isg"], s["Tcw"], s["Tcb"], s["Tbc"]))
    close = np.ascontiguousarray(s["close"], np.uint8)
    state = f64(np.array(s["state"]).copy())
    prev, pre, pj, pb, iI, iG, iA, ps, pH = map(f64, (s["prev"], s["preint"], s["preint_jac"], s["preint_bias"], s["infoI"], s["infoG"],
                                                       s["infoA"], s["prior_state"], s["prior_H"]))
    outlier = np.zeros(max(E, 1), np.uint8)
    H = np.zeros(225, np.float64)
    n = C.c_int(0)
    iters = np.zeros(4, np.int32)
    rc = lib().ork_pose_inertial_opt_last_frame(E, _p(xw), _p(obs), _p(isg), _p(close), C.byref(cam), _p(Tcw), _p(Tcb), _p(Tbc),
                                                _p(state), _p(prev), _p(pre), _p(pj), _p(pb), _p(iI), _p(iG), _p(iA), _p(ps), _p(pH),
                                                int(rec_init), _p(outlier), _p(H), C.byref(n), _p(iters))
    assert rc == 0
    return dict(state=state, outlier=outlier[:E], H=H.reshape(15, 15), n=n.value, iters=iters)


def inertial_lf_debug(state, prev, s):
    f64 = lambda a: np.ascontiguousarray(a, np.float64)   # noqa: E731
    state, prev, pre, pj, pb, ps = map(f64, (state, prev, s["preint"], s["preint_jac"], s["preint_bias"], s["prior_state"]))
    e9, J, e15, Jp = np.zeros(9), np.zeros(216), np.zeros(15), np.zeros(225)
    lib().ork_inertial_lf_debug(_p(state), _p(prev), _p(pre), _p(pj), _p(pb), _p(ps), _p(e9), _p(J), _p(e15), _p(Jp))
    return e9, J.reshape(9, 24), e15, Jp.reshape(15, 15)


def jacobi_eig(A):
    A = np.ascontiguousarray(A, np.float64).copy()
    n = len(A)
    V = np.zeros((n, n))
    lib().ork_jacobi_eig(n, _p(A), _p(V))
    return np.diag(A).copy(), V, A


def marginalize_prev(H30):
    H30 = np.ascontiguousarray(H30, np.float64)
    out = np.zeros((15, 15))
    lib().ork_marginalize_prev(_p(H30), _p(out))
    return out


def inertial_debug(state, kf, preint):
    f64 = lambda a: np.ascontiguousarray(a, np.float64)   # noqa: E731
    state, kf, preint = map(f64, (state, kf, preint))
    e9, J = np.zeros(9), np.zeros(81)
    lib().ork_inertial_debug(_p(state), _p(kf), _p(preint), _p(e9), _p(J))
    return e9, J.reshape(9, 9)
