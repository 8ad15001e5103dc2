"""ctypes binding of the C ABI in include/orbx.h."""
import ctypes as C
import os
import re
import subprocess
import numpy as np
from . import abi
from .abi import Frame, Camera, make_camera, ptr as _ptr, c32 as _c32  # noqa: F401

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ROOT = os.path.dirname(_PKG)
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4")])


class OrbxError(RuntimeError):
    pass


def lib_path():
    return os.path.join(_PKG, "liborbx.so")


def build_library(force=False):
    """nvcc-compile every kernel for sm_100a into liborbx.so (in-tree)."""
    args = ["make", "-C", _PKG, "-s"]
    if force:
        args.append("-B")
    subprocess.check_call(args + ["liborbx.so"])
    return lib_path()


def declared_symbols():
    """Every function include/orbx.h declares (used by the CPU-side ABI test)."""
    txt = open(os.path.join(_ROOT, "include", "orbx.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(orbx_[a-z0-9_]+)\s*\(", txt)))


def load_library():
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise OrbxError("liborbx.so is not built (run __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(p)
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    L.orbx_abi_version.restype = i
    L.orbx_last_error.restype = C.c_char_p
    L.orbx_create.restype = vp
    L.orbx_create.argtypes = [i]
    L.orbx_destroy.argtypes = [vp]
    L.orbx_stream.restype = vp
    L.orbx_stream.argtypes = [vp]
    L.orbx_synchronize.argtypes = [vp]
    L.orbx_launch_count.restype = C.c_uint64
    L.orbx_launch_count.argtypes = [vp]
    L.orbx_host_alloc.restype = vp
    L.orbx_host_alloc.argtypes = [C.c_size_t]
    L.orbx_host_free.argtypes = [vp]
    L.orbx_extractor_create.restype = vp
    L.orbx_extractor_create.argtypes = [vp, i, f, i, i, i, i, i, i]
    L.orbx_extractor_destroy.argtypes = [vp]
    L.orbx_extractor_levels.argtypes = [vp]
    L.orbx_extractor_scale_tables.argtypes = [vp, vp, vp, vp, vp]
    L.orbx_extractor_features_per_level.argtypes = [vp, vp]
    L.orbx_extractor_max_keypoints.argtypes = [vp]
    L.orbx_extractor_stream.restype = vp
    L.orbx_extractor_stream.argtypes = [vp]
    L.orbx_extract.argtypes = [vp, vp, i, i, i, i, i, vp, vp, i, vp, vp]
    L.orbx_extract_batch.argtypes = [vp, i, vp, i, i, i, i, i, vp, vp, i, vp, vp]
    L.orbx_extract_batch_device.argtypes = [vp, i, vp, i, i, i, i, i, vp, vp, i, vp, vp]
    L.orbx_extractor_set_profiling.argtypes = [vp, i]
    L.orbx_extractor_stage_ms.argtypes = [vp, vp, vp]
    L.orbx_descriptor_distance.argtypes = [vp, vp]
    L.orbx_features_in_area.argtypes = [vp, vp, i, vp, vp, vp, vp, vp, vp, i, vp]
    L.orbx_stereo_match.argtypes = [vp, vp, i, vp, i, vp, vp, i, vp, vp, i, f, f, vp, vp]
    L.orbx_search_by_projection_map.argtypes = [vp, vp, vp, i, vp, vp, vp, vp, vp, vp, vp, f, f, vp, i, vp, vp]
    L.orbx_search_by_projection_frame.argtypes = [vp, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, f, i, i, vp, i,
                                                  vp, vp, vp, vp]
    L.orbx_search_for_triangulation.argtypes = [vp, vp, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, vp, vp, vp, vp,
                                                vp, vp, vp, vp, i, i, i, i, vp, vp]
    L.orbx_pose_optimization.argtypes = [vp, i, vp, vp, vp, vp, vp, vp, vp, vp]
    L.orbx_pose_optimization_batch_device.argtypes = [vp, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.orbx_local_ba.argtypes = [vp, i, vp, vp, i, vp, i, vp, vp, vp, vp, vp, C.c_double, vp, vp, vp, vp]
    L.orbx_frame_upload.restype = vp
    L.orbx_frame_upload.argtypes = [vp, vp]
    L.orbx_frame_release.argtypes = [vp]
    L.orbx_frame_count.argtypes = [vp]
    L.orbx_tri_batch_prepare.restype = vp
    L.orbx_tri_batch_prepare.argtypes = [vp, i, vp, vp, vp, i, i]
    L.orbx_tri_batch_run.argtypes = [vp, vp]
    L.orbx_tri_batch_fetch.argtypes = [vp, vp]
    L.orbx_tri_batch_destroy.argtypes = [vp]
    L.orbx_lba_batch_prepare.restype = vp
    L.orbx_lba_batch_prepare.argtypes = [vp, i, vp, vp]
    L.orbx_lba_batch_run.argtypes = [vp, vp]
    L.orbx_lba_batch_fetch.argtypes = [vp, vp]
    L.orbx_lba_batch_destroy.argtypes = [vp]
    L.orbx_lba_batch_device_bytes.restype = C.c_size_t
    L.orbx_lba_batch_device_bytes.argtypes = [vp]
    L.orbx_pose_inertial_optimization_last_keyframe.argtypes = [vp, i] + [vp] * 14 + [i] + [vp] * 4
    L.orbx_pose_inertial_optimization_last_frame.argtypes = [vp, i] + [vp] * 18 + [i] + [vp] * 4
    L.orbx_pose_inertial_optimization_last_frame_batch.argtypes = [vp, i] + [vp] * 19 + [i] + [vp] * 4
    L.orbx_pose_inertial_optimization_last_keyframe_batch.argtypes = [vp, i] + [vp] * 15 + [i] + [vp] * 4
    L.orbx_tracker_create.restype = vp
    L.orbx_tracker_create.argtypes = [vp, vp, i, vp, f, f, f]
    L.orbx_tracker_create_mono.restype = vp
    L.orbx_tracker_create_mono.argtypes = [vp, vp, i, vp, f, f, f]
    L.orbx_tracker_images_per_stream.argtypes = [vp]
    L.orbx_tracker_destroy.argtypes = [vp]
    L.orbx_tracker_step_device.argtypes = [vp, vp, i, i, i, vp, vp, vp, vp]
    L.orbx_tracker_step.argtypes = [vp, vp, i, i, i, vp, vp, vp, vp]
    L.orbx_tracker_submit.argtypes = [vp, vp, i, i, i, vp, vp]
    L.orbx_tracker_collect.argtypes = [vp, vp, vp]
    L.orbx_tracker_set_overlap.argtypes = [vp, i]
    L.orbx_tracker_result_stream.restype = vp
    L.orbx_tracker_result_stream.argtypes = [vp]
    L.orbx_tracker_synchronize.argtypes = [vp]
    L.orbx_tracker_set_keyframe_work.argtypes = [vp, vp, vp, i]
    L.orbx_tracker_map_capacity.argtypes = [vp]
    L.orbx_tracker_map_bytes.restype = C.c_size_t
    L.orbx_tracker_map_bytes.argtypes = [vp]
    L.orbx_tracker_set_map.argtypes = [vp, vp]
    L.orbx_tracker_upload_map.argtypes = [vp, vp]
    L.orbx_tracker_set_graph.argtypes = [vp, i]
    L.orbx_tracker_graph_launches.restype = C.c_longlong
    L.orbx_tracker_graph_launches.argtypes = [vp]
    L.orbx_tracker_set_inertial.argtypes = [vp, vp]
    L.orbx_tracker_upload_inertial.argtypes = [vp, vp]
    L.orbx_tracker_inertial_result.argtypes = [vp, vp, vp]
    L.orbx_tracker_inertial_state_dev.restype = vp
    L.orbx_tracker_inertial_state_dev.argtypes = [vp]
    L.orbx_tracker_inertial_hessian_dev.restype = vp
    L.orbx_tracker_inertial_hessian_dev.argtypes = [vp]
    L.orbx_tracker_set_chain.argtypes = [vp, i, vp]
    L.orbx_tracker_keyframe_stream.restype = vp
    L.orbx_tracker_keyframe_stream.argtypes = [vp]
    L.orbx_tracker_keyframe_runs.restype = C.c_longlong
    L.orbx_tracker_keyframe_runs.argtypes = [vp]
    L.orbx_tracker_set_profiling.argtypes = [vp, i]
    L.orbx_tracker_stage_ms.argtypes = [vp, vp]
    L.orbx_pyramid_level.argtypes = [vp, i, i, vp, i, vp, vp]
    L.orbx_search_by_bow.argtypes = [vp, vp, vp, vp, i, vp, vp, vp, i, vp, vp, vp, f, i, vp, vp]
    L.orbx_fuse.argtypes = [vp, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, vp, f, vp, vp, i, f, vp, vp]
    L.orbx_is_in_frustum.argtypes = [vp, vp, vp, vp, vp, f, f, f, f, f, i, f, i] + [vp] * 12
    L.orbx_undistort_keypoints.argtypes = [vp, vp, i, vp, vp, i, vp]
    L.orbx_vocabulary_load.restype = vp
    L.orbx_vocabulary_load.argtypes = [vp, C.c_char_p]
    L.orbx_vocabulary_from_memory.restype = vp
    L.orbx_vocabulary_from_memory.argtypes = [vp, vp, C.c_size_t]
    L.orbx_vocabulary_destroy.argtypes = [vp]
    L.orbx_vocabulary_info.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.orbx_vocabulary_transform.argtypes = [vp, vp, i, i, vp, vp, vp, vp, vp, vp, vp]
    L.orbx_vocabulary_transform_batch_device.argtypes = [vp, i, vp, vp, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.orbx_debug_candidates.argtypes = [vp, i, i, vp, vp, i, vp]
    L.orbx_debug_blur_level.argtypes = [vp, i, i, vp, i, vp, vp]
    _LIB = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _check(rc, what):
    if rc != 0:
        msg = load_library().orbx_last_error().decode(errors="replace")
        raise OrbxError("%s failed with status %d: %s" % (what, rc, msg))


def host_array(shape, dtype=np.uint8):
    """numpy array backed by page-locked memory (orbx_host_alloc): host-pointer entry points DMA from it directly.
    The memory lives until the process exits (the finaliser frees it when the array is collected)."""
    import weakref
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = load_library().orbx_host_alloc(n)
    if not p:
        raise OrbxError("orbx_host_alloc(%d) failed" % n)
    buf = (C.c_uint8 * n).from_address(p)
    a = np.frombuffer(buf, dtype=dtype).reshape(shape)
    weakref.finalize(buf, load_library().orbx_host_free, p)
    return a


class Context:
    """orbx_ctx: one per (process, device)."""

    def __init__(self, device=0):
        L = load_library()
        self.h = L.orbx_create(device)
        if not self.h:
            raise OrbxError("orbx_create(%d): %s" % (device, L.orbx_last_error().decode(errors="replace")))
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            load_library().orbx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        _check(load_library().orbx_synchronize(self.h), "orbx_synchronize")

    @property
    def launches(self):
        return int(load_library().orbx_launch_count(self.h))


class ORBextractor:
    """Mirror of ORB_SLAM3::ORBextractor (include/ORBextractor.h:43-109).

    __call__(image, lapping) -> (monoIndex_or_-1, keypoints, descriptors) follows operator()
    (src/ORBextractor.cc:1074-1156): an empty image returns -1 and empty outputs.
    """

    def __init__(self, ctx, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7,
                 max_w=752, max_h=480, max_batch=1):
        L = load_library()
        self.ctx = ctx
        self.h = L.orbx_extractor_create(ctx.h, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_w,
                                         max_h, max_batch)
        if not self.h:
            raise OrbxError("orbx_extractor_create: " + L.orbx_last_error().decode(errors="replace"))
        self.nfeatures, self.nlevels, self.max_batch = nfeatures, nlevels, max_batch
        t = [np.empty(nlevels, np.float32) for _ in range(4)]
        L.orbx_extractor_scale_tables(self.h, *[_p(a) for a in t])
        self.mvScaleFactor, self.mvInvScaleFactor, self.mvLevelSigma2, self.mvInvLevelSigma2 = t
        nf = np.empty(nlevels, np.int32)
        L.orbx_extractor_features_per_level(self.h, _p(nf))
        self.mnFeaturesPerLevel = nf
        self.cap = L.orbx_extractor_max_keypoints(self.h)
        self.stream = L.orbx_extractor_stream(self.h)

    def close(self):
        if getattr(self, "h", None):
            load_library().orbx_extractor_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # reference getters (include/ORBextractor.h:59-81)
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactors(self):
        return self.mvScaleFactor

    def GetInverseScaleFactors(self):
        return self.mvInvScaleFactor

    def GetScaleSigmaSquares(self):
        return self.mvLevelSigma2

    def GetInverseScaleSigmaSquares(self):
        return self.mvInvLevelSigma2

    def __call__(self, image, vLappingArea=(0, 0)):
        if image is None or image.size == 0:
            return -1, np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8)
        image = np.ascontiguousarray(image, np.uint8)
        kps = np.empty(self.cap, KP_DTYPE)
        desc = np.empty((self.cap, 32), np.uint8)
        n, mono = C.c_int(0), C.c_int(0)
        rc = load_library().orbx_extract(self.h, _p(image), image.shape[1], image.shape[0], image.strides[0],
                                         int(vLappingArea[0]), int(vLappingArea[1]), _p(kps), _p(desc), self.cap,
                                         C.byref(n), C.byref(mono))
        if rc == -1:
            return -1, np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8)
        _check(rc, "orbx_extract")
        return mono.value, kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images, vLappingArea=(0, 0)):
        """images: sequence of equally sized uint8 arrays -> list of (monoIndex, kps, desc)."""
        imgs = [np.ascontiguousarray(im, np.uint8) for im in images]
        B = len(imgs)
        h, w = imgs[0].shape
        assert all(im.shape == (h, w) for im in imgs)
        ptrs = (C.c_void_p * B)(*[im.ctypes.data for im in imgs])
        kps = np.empty((B, self.cap), KP_DTYPE)
        desc = np.empty((B, self.cap, 32), np.uint8)
        n = np.zeros(B, np.int32)
        mono = np.zeros(B, np.int32)
        rc = load_library().orbx_extract_batch(self.h, B, ptrs, w, h, w, int(vLappingArea[0]), int(vLappingArea[1]),
                                               _p(kps), _p(desc), self.cap, _p(n), _p(mono))
        _check(rc, "orbx_extract_batch")
        return [(int(mono[b]), kps[b, :n[b]].copy(), desc[b, :n[b]].copy()) for b in range(B)]

    def extract_batch_device(self, d_imgs_ptr, B, w, h, stride, d_kps_ptr, d_desc_ptr, cap, d_n_ptr, d_mono_ptr,
                             vLappingArea=(0, 0)):
        rc = load_library().orbx_extract_batch_device(self.h, B, d_imgs_ptr, w, h, stride, int(vLappingArea[0]),
                                                      int(vLappingArea[1]), d_kps_ptr, d_desc_ptr, cap, d_n_ptr,
                                                      d_mono_ptr)
        _check(rc, "orbx_extract_batch_device")

    STAGES = ("pyramid", "fast", "blur", "quadtree", "describe")

    def set_profiling(self, on=True):
        _check(load_library().orbx_extractor_set_profiling(self.h, int(on)), "orbx_extractor_set_profiling")

    def stage_ms(self):
        ms = np.zeros(len(self.STAGES), np.float32)
        ln = np.zeros(len(self.STAGES), np.int32)
        _check(load_library().orbx_extractor_stage_ms(self.h, _p(ms), _p(ln)), "orbx_extractor_stage_ms")
        return ms, ln

    def pyramid_level(self, level, b=0):
        """mvImagePyramid[level] of image b of the last call."""
        w, h = C.c_int(0), C.c_int(0)
        _check(load_library().orbx_pyramid_level(self.h, b, level, None, 0, C.byref(w), C.byref(h)),
               "orbx_pyramid_level")
        out = np.empty((h.value, w.value), np.uint8)
        _check(load_library().orbx_pyramid_level(self.h, b, level, _p(out), w.value, C.byref(w), C.byref(h)),
               "orbx_pyramid_level")
        return out

    def debug_blur_level(self, level, b=0):
        """The blurred copy of mvImagePyramid[level] that the descriptors are sampled from (test hook)."""
        w, h = C.c_int(0), C.c_int(0)
        _check(load_library().orbx_debug_blur_level(self.h, b, level, None, 0, C.byref(w), C.byref(h)), "orbx_debug_blur_level")
        out = np.empty((h.value, w.value), np.uint8)
        _check(load_library().orbx_debug_blur_level(self.h, b, level, _p(out), w.value, C.byref(w), C.byref(h)),
               "orbx_debug_blur_level")
        return out

    def debug_candidates(self, level, b=0):
        cap = 1 << 20
        xy = np.empty((cap, 2), np.int16)
        sc = np.empty(cap, np.uint8)
        n = C.c_int(0)
        _check(load_library().orbx_debug_candidates(self.h, b, level, _p(xy), _p(sc), cap, C.byref(n)),
               "orbx_debug_candidates")
        return xy[:n.value].copy(), sc[:n.value].copy()


class ORBmatcher:
    """Mirror of ORB_SLAM3::ORBmatcher (include/ORBmatcher.h:35-108) for the hot-path overloads, on flat
    arrays.  Results come back as index arrays; the C++ shim scatters them into Frame::mvpMapPoints."""

    TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30

    def __init__(self, ctx, nnratio=0.6, checkOri=True):
        self.ctx, self.mfNNratio, self.mbCheckOrientation = ctx, float(nnratio), bool(checkOri)

    @staticmethod
    def DescriptorDistance(a, b):
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        return load_library().orbx_descriptor_distance(_p(a), _p(b))

    def SearchByProjectionMap(self, F, kp_blocked, projX, projY, projXR, level, viewCos, mpDesc, flags, th,
                              scaleFactors):
        nq = len(projX)
        best = np.full(nq, -1, np.int32)
        nm = C.c_int(0)
        a = [_c32(kp_blocked, np.uint8), _c32(projX, np.float32), _c32(projY, np.float32), _c32(projXR, np.float32),
             _c32(level, np.int32), _c32(viewCos, np.float32), _c32(mpDesc, np.uint8), _c32(flags, np.uint8)]
        sf = _c32(scaleFactors, np.float32)
        rc = load_library().orbx_search_by_projection_map(self.ctx.h, F.ref(), _ptr(a[0]), nq, _ptr(a[1]), _ptr(a[2]),
                                                          _ptr(a[3]), _ptr(a[4]), _ptr(a[5]), _ptr(a[6]), _ptr(a[7]),
                                                          float(th), self.mfNNratio, _ptr(sf), len(sf), _p(best),
                                                          C.byref(nm))
        _check(rc, "orbx_search_by_projection_map")
        return nm.value, best

    def SearchByProjectionFrame(self, Cur, cur_blocked, cam, Tcw_cur, Tcw_last, flags, xw, octave, angle, mpDesc, th,
                                bMono, scaleFactors):
        nq = len(flags)
        match = np.full(nq, -1, np.int32)
        kept = np.zeros(nq, np.uint8)
        cur_match = np.full(Cur.n, -1, np.int32)
        nm = C.c_int(0)
        a = [_c32(cur_blocked, np.uint8), _c32(Tcw_cur, np.float32), _c32(Tcw_last, np.float32),
             _c32(flags, np.uint8), _c32(xw, np.float32), _c32(octave, np.int32), _c32(angle, np.float32),
             _c32(mpDesc, np.uint8)]
        sf = _c32(scaleFactors, np.float32)
        rc = load_library().orbx_search_by_projection_frame(self.ctx.h, Cur.ref(), _ptr(a[0]), C.byref(cam), _ptr(a[1]),
                                                            _ptr(a[2]), nq, _ptr(a[3]), _ptr(a[4]), _ptr(a[5]),
                                                            _ptr(a[6]), _ptr(a[7]), float(th), int(bMono),
                                                            int(self.mbCheckOrientation), _ptr(sf), len(sf), _p(match),
                                                            _p(kept), _p(cur_match), C.byref(nm))
        _check(rc, "orbx_search_by_projection_frame")
        return nm.value, match, kept, cur_match

    def SearchForTriangulation(self, KF1, KF2, has1, has2, fv1, fv2, cam1, cam2, R1w, t1w, R2w, t2w, sigma2,
                               scaleFactors, bOnlyStereo=False, bCoarse=False):
        """fv = (node_ids, offsets, indices) CSR of a DBoW2::FeatureVector."""
        m12 = np.full(KF1.n, -1, np.int32)
        nm = C.c_int(0)
        f1 = [_c32(v, np.int32) for v in fv1]
        f2 = [_c32(v, np.int32) for v in fv2]
        a = [_c32(has1, np.uint8), _c32(has2, np.uint8), _c32(R1w, np.float32), _c32(t1w, np.float32),
             _c32(R2w, np.float32), _c32(t2w, np.float32), _c32(sigma2, np.float32), _c32(scaleFactors, np.float32)]
        rc = load_library().orbx_search_for_triangulation(
            self.ctx.h, KF1.ref(), KF2.ref(), _ptr(a[0]), _ptr(a[1]), len(f1[0]), _ptr(f1[0]), _ptr(f1[1]), _ptr(f1[2]),
            len(f2[0]), _ptr(f2[0]), _ptr(f2[1]), _ptr(f2[2]), C.byref(cam1), C.byref(cam2), _ptr(a[2]), _ptr(a[3]),
            _ptr(a[4]), _ptr(a[5]), _ptr(a[6]), _ptr(a[7]), len(a[7]), int(bOnlyStereo), int(bCoarse),
            int(self.mbCheckOrientation), _p(m12), C.byref(nm))
        _check(rc, "orbx_search_for_triangulation")
        return nm.value, m12


def features_in_area(ctx, F, x, y, r, minLevel, maxLevel, cap=256):
    nq = len(x)
    out = np.full((nq, cap), -1, np.int32)
    n = np.zeros(nq, np.int32)
    a = [_c32(x, np.float32), _c32(y, np.float32), _c32(r, np.float32), _c32(minLevel, np.int32),
         _c32(maxLevel, np.int32)]
    _check(load_library().orbx_features_in_area(ctx.h, F.ref(), nq, *[_ptr(v) for v in a], _p(out), cap, _p(n)),
           "orbx_features_in_area")
    return out, n


def stereo_match(ctx, extL, bL, extR, bR, kpL, descL, kpR, descR, bf, b):
    """Frame::ComputeStereoMatches -> (mvuRight, mvDepth)."""
    kpL, kpR = np.ascontiguousarray(kpL), np.ascontiguousarray(kpR)
    descL, descR = np.ascontiguousarray(descL, np.uint8), np.ascontiguousarray(descR, np.uint8)
    ur = np.empty(len(kpL), np.float32)
    dp = np.empty(len(kpL), np.float32)
    _check(load_library().orbx_stereo_match(ctx.h, extL.h, bL, extR.h, bR, _p(kpL), _p(descL), len(kpL), _p(kpR),
                                            _p(descR), len(kpR), float(bf), float(b), _p(ur), _p(dp)),
           "orbx_stereo_match")
    return ur, dp


class ResidentFrame:
    """A Frame / KeyFrame kept on the device between calls (orbx_frame_upload): keypoints, descriptors, mvuRight and the
    64x48 grid cross the boundary once; later matcher calls that receive the same `Frame` object use the resident copy."""

    def __init__(self, ctx, frame):
        self.ctx, self.frame = ctx, frame            # keeps the host arrays (the lookup key) alive
        self.h = load_library().orbx_frame_upload(ctx.h, frame.ref())
        if not self.h:
            raise OrbxError("orbx_frame_upload: " + load_library().orbx_last_error().decode(errors="replace"))

    def release(self):
        if getattr(self, "h", None):
            load_library().orbx_frame_release(self.h)
            self.h = None

    __del__ = release


class TriangulationBatch:
    """Prepared plan of many ORBmatcher::SearchForTriangulation calls (orbx_tri_batch_*, include/orbx.h): the
    CreateNewMapPoints searches of the keyframe-rate step of many streams, one launch sequence.

    problems: list of dicts with the arguments of ORBmatcher.SearchForTriangulation
              (KF1, KF2, has1, has2, fv1, fv2, cam1, cam2, R1w, t1w, R2w, t2w, only_stereo, coarse)."""

    def __init__(self, ctx, problems, sigma2, scaleFactors, check_orientation=True):
        self.ctx, self.Q = ctx, len(problems)
        self._keep = []
        self.P = (abi.TriProblem * self.Q)()
        self.match = []
        for q, pr in enumerate(problems):
            f1 = [_c32(v, np.int32) for v in pr["fv1"]]
            f2 = [_c32(v, np.int32) for v in pr["fv2"]]
            a = [_c32(pr["has1"], np.uint8), _c32(pr["has2"], np.uint8), _c32(pr["R1w"], np.float32), _c32(pr["t1w"], np.float32),
                 _c32(pr["R2w"], np.float32), _c32(pr["t2w"], np.float32)]
            m = np.full(max(pr["KF1"].n, 1), -1, np.int32)
            self._keep += [f1, f2, a, pr["KF1"], pr["KF2"], pr["cam1"], pr["cam2"]]
            self.match.append(m)
            P = self.P[q]
            P.kf1, P.kf2 = C.addressof(pr["KF1"].c), C.addressof(pr["KF2"].c)
            P.has_mp1, P.has_mp2 = a[0].ctypes.data, a[1].ctypes.data
            P.nn1, P.fv1_node, P.fv1_off, P.fv1_idx = len(f1[0]), f1[0].ctypes.data, f1[1].ctypes.data, f1[2].ctypes.data
            P.nn2, P.fv2_node, P.fv2_off, P.fv2_idx = len(f2[0]), f2[0].ctypes.data, f2[1].ctypes.data, f2[2].ctypes.data
            P.cam1, P.cam2 = C.addressof(pr["cam1"]), C.addressof(pr["cam2"])
            P.R1w, P.t1w, P.R2w, P.t2w = a[2].ctypes.data, a[3].ctypes.data, a[4].ctypes.data, a[5].ctypes.data
            P.only_stereo, P.coarse = int(pr.get("only_stereo", False)), int(pr.get("coarse", False))
            P.match12 = m.ctypes.data
        sg, sf = _c32(sigma2, np.float32), _c32(scaleFactors, np.float32)
        self.h = load_library().orbx_tri_batch_prepare(ctx.h, self.Q, C.byref(self.P), _p(sg), _p(sf), len(sf), int(check_orientation))
        if not self.h:
            raise OrbxError("orbx_tri_batch_prepare: " + load_library().orbx_last_error().decode(errors="replace"))

    def run(self, stream=None):
        _check(load_library().orbx_tri_batch_run(self.h, stream), "orbx_tri_batch_run")

    def fetch(self):
        """-> list of (nmatches, match12[n1])"""
        _check(load_library().orbx_tri_batch_fetch(self.h, C.byref(self.P)), "orbx_tri_batch_fetch")
        return [(int(self.P[q].nmatches), self.match[q][:self._keep[7 * q + 3].n].copy()) for q in range(self.Q)]

    def close(self):
        if getattr(self, "h", None):
            load_library().orbx_tri_batch_destroy(self.h)
            self.h = None

    __del__ = close


class LocalBABatch:
    """Prepared plan of many Optimizer::LocalBundleAdjustment problems (orbx_lba_batch_*, include/orbx.h), one CTA per
    problem in one launch.  problems: list of dicts (kf_T, kf_fixed, mp_xyz, e_kf, e_mp, e_obs, e_inv_sigma2[, lambda_init])."""

    def __init__(self, ctx, problems, cam):
        self.ctx, self.n = ctx, len(problems)
        self.P = (abi.LbaProblem * self.n)()
        self._keep, self.out = [cam], []
        for q, pr in enumerate(problems):
            T = np.array(pr["kf_T"], np.float32).reshape(-1, 16).copy()
            X = np.array(pr["mp_xyz"], np.float32).reshape(-1, 3).copy()
            a = [_c32(pr["kf_fixed"], np.uint8), _c32(pr["e_kf"], np.int32), _c32(pr["e_mp"], np.int32), _c32(pr["e_obs"], np.float32),
                 _c32(pr["e_inv_sigma2"], np.float32)]
            bad = np.zeros(max(len(a[1]), 1), np.uint8)
            self._keep.append(a)
            self.out.append((T, X, bad))
            P = self.P[q]
            P.n_kf, P.n_mp, P.n_edges = len(T), len(X), len(a[1])
            P.kf_Tcw, P.kf_fixed, P.mp_xyz = T.ctypes.data, a[0].ctypes.data, X.ctypes.data
            P.e_kf, P.e_mp, P.e_obs, P.e_inv_sigma2 = a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data, a[4].ctypes.data
            P.lambda_init = float(pr.get("lambda_init", 0.0))
            P.edge_bad = bad.ctypes.data
        self.h = load_library().orbx_lba_batch_prepare(ctx.h, self.n, C.byref(self.P), C.byref(cam))
        if not self.h:
            raise OrbxError("orbx_lba_batch_prepare: " + load_library().orbx_last_error().decode(errors="replace"))
        self.device_bytes = load_library().orbx_lba_batch_device_bytes(self.h)

    def run(self, stream=None):
        _check(load_library().orbx_lba_batch_run(self.h, stream), "orbx_lba_batch_run")

    def fetch(self):
        """-> list of (kf_T[K,4,4], mp_xyz[M,3], edge_bad[E], iters[2], status), like Optimizer.LocalBundleAdjustment"""
        _check(load_library().orbx_lba_batch_fetch(self.h, C.byref(self.P)), "orbx_lba_batch_fetch")
        res = []
        for q in range(self.n):
            T, X, bad = self.out[q]
            res.append((T.reshape(-1, 4, 4).copy(), X.copy(), bad[:self.P[q].n_edges].copy(),
                        np.array([self.P[q].iters[0], self.P[q].iters[1]], np.int32), int(self.P[q].status)))
        return res

    def close(self):
        if getattr(self, "h", None):
            load_library().orbx_lba_batch_destroy(self.h)
            self.h = None

    __del__ = close


class Optimizer:
    """Mirror of the static ORB_SLAM3::Optimizer entry points on the hot path (include/Optimizer.h:58,62,64,65)."""

    def __init__(self, ctx):
        self.ctx = ctx

    def PoseOptimization(self, xw, obs, inv_sigma2, cam, Tcw):
        """-> (Tcw_out[4,4], mvbOutlier[E], return value, LM iterations per round[4])"""
        xw, obs, isg = _c32(xw, np.float32), _c32(obs, np.float32), _c32(inv_sigma2, np.float32)
        T = np.array(Tcw, np.float32).reshape(4, 4).copy()
        E = len(isg)
        outl = np.zeros(max(E, 1), np.uint8)
        nin = C.c_int(0)
        iters = np.zeros(4, np.int32)
        rc = load_library().orbx_pose_optimization(self.ctx.h, E, _ptr(xw), _ptr(obs), _ptr(isg), C.byref(cam), _p(T),
                                                   _p(outl), C.byref(nin), _p(iters))
        _check(rc, "orbx_pose_optimization")
        return T, outl[:E].copy(), nin.value, iters

    def LocalBundleAdjustment(self, kf_T, kf_fixed, mp_xyz, e_kf, e_mp, e_obs, e_inv_sigma2, cam, lambda_init=0.0,
                              stop=None):
        """-> (kf_T_out[K,4,4], mp_xyz_out[M,3], edge_bad[E], iters[2], status)"""
        T = np.array(kf_T, np.float32).reshape(-1, 16).copy()
        X = np.array(mp_xyz, np.float32).reshape(-1, 3).copy()
        fixed = _c32(kf_fixed, np.uint8)
        ekf, emp = _c32(e_kf, np.int32), _c32(e_mp, np.int32)
        obs, isg = _c32(e_obs, np.float32), _c32(e_inv_sigma2, np.float32)
        E = len(ekf)
        bad = np.zeros(max(E, 1), np.uint8)
        iters = np.zeros(2, np.int32)
        status = C.c_int(0)
        st = _c32(stop, np.uint8) if stop is not None else None
        rc = load_library().orbx_local_ba(self.ctx.h, len(T), _p(T), _ptr(fixed), len(X), _p(X), E, _ptr(ekf), _ptr(emp),
                                          _ptr(obs), _ptr(isg), C.byref(cam), float(lambda_init), _ptr(st), _p(bad),
                                          _p(iters), C.byref(status))
        _check(rc, "orbx_local_ba")
        return T.reshape(-1, 4, 4), X, bad[:E].copy(), iters, status.value

    def PoseInertialOptimizationLastKeyFrame(self, xw, obs, inv_sigma2, close_pt, cam, Tcw, Tcb, Tbc, state, kf_state,
                                             preint, info_inertial, info_gyro, info_acc, rec_init=False):
        """include/Optimizer.h:64.  -> dict(state[21], outlier[E], H[15,15], n, iters[4]); argument layout: include/orbx.h"""
        f64 = lambda a, n: np.ascontiguousarray(np.asarray(a, np.float64).reshape(-1)[:n])   # noqa: E731
        xw, obs, isg = _c32(xw, np.float32), _c32(obs, np.float32), _c32(inv_sigma2, np.float32)
        close = _c32(close_pt, np.uint8)
        T, Tcb, Tbc = (np.ascontiguousarray(np.asarray(a, np.float32).reshape(4, 4)) for a in (Tcw, Tcb, Tbc))
        st = f64(state, 21).copy()
        kf, pre, iI, iG, iA = f64(kf_state, 21), f64(preint, 16), f64(info_inertial, 81), f64(info_gyro, 9), f64(info_acc, 9)
        E = len(isg)
        outl = np.zeros(max(E, 1), np.uint8)
        H = np.zeros(225, np.float64)
        n = C.c_int(0)
        iters = np.zeros(4, np.int32)
        rc = load_library().orbx_pose_inertial_optimization_last_keyframe(
            self.ctx.h, E, _ptr(xw), _ptr(obs), _ptr(isg), _ptr(close), C.byref(cam), _p(T), _p(Tcb), _p(Tbc), _p(st), _p(kf),
            _p(pre), _p(iI), _p(iG), _p(iA), int(rec_init), _p(outl), _p(H), C.byref(n), _p(iters))
        _check(rc, "orbx_pose_inertial_optimization_last_keyframe")
        return dict(state=st, outlier=outl[:E].copy(), H=H.reshape(15, 15), n=n.value, iters=iters)

    def PoseInertialOptimizationLastFrame(self, xw, obs, inv_sigma2, close_pt, cam, Tcw, Tcb, Tbc, state, prev_state, preint,
                                          preint_jac, preint_bias, info_inertial, info_gyro, info_acc, prior_state, prior_H,
                                          rec_init=False):
        """include/Optimizer.h:65.  -> dict(state[21], outlier[E], H[15,15], n, iters[4]); argument layout: include/orbx.h"""
        f64 = lambda a, n: np.ascontiguousarray(np.asarray(a, np.float64).reshape(-1)[:n])   # noqa: E731
        xw, obs, isg = _c32(xw, np.float32), _c32(obs, np.float32), _c32(inv_sigma2, np.float32)
        close = _c32(close_pt, np.uint8)
        T, Tcb, Tbc = (np.ascontiguousarray(np.asarray(a, np.float32).reshape(4, 4)) for a in (Tcw, Tcb, Tbc))
        st = f64(state, 21).copy()
        pv, pre, pj, pb = f64(prev_state, 21), f64(preint, 16), f64(preint_jac, 45), f64(preint_bias, 6)
        iI, iG, iA, ps, pH = f64(info_inertial, 81), f64(info_gyro, 9), f64(info_acc, 9), f64(prior_state, 21), f64(prior_H, 225)
        E = len(isg)
        outl = np.zeros(max(E, 1), np.uint8)
        H = np.zeros(225, np.float64)
        n = C.c_int(0)
        iters = np.zeros(4, np.int32)
        rc = load_library().orbx_pose_inertial_optimization_last_frame(
            self.ctx.h, E, _ptr(xw), _ptr(obs), _ptr(isg), _ptr(close), C.byref(cam), _p(T), _p(Tcb), _p(Tbc), _p(st), _p(pv),
            _p(pre), _p(pj), _p(pb), _p(iI), _p(iG), _p(iA), _p(ps), _p(pH), int(rec_init), _p(outl), _p(H), C.byref(n), _p(iters))
        _check(rc, "orbx_pose_inertial_optimization_last_frame")
        return dict(state=st, outlier=outl[:E].copy(), H=H.reshape(15, 15), n=n.value, iters=iters)

    @staticmethod
    def pack_inertial_lf(problems):
        """Concatenate a list of single-call argument dicts into the arrays of the many-stream entry point (reusable)."""
        P = len(problems)
        cat32 = lambda k, w: np.ascontiguousarray(np.concatenate([np.asarray(q[k], np.float32).reshape(-1, w) for q in problems]))  # noqa: E731
        st64 = lambda k, n: np.ascontiguousarray(np.stack([np.asarray(q[k], np.float64).reshape(-1)[:n] for q in problems]))     # noqa: E731
        ofs = np.zeros(P + 1, np.int32)
        ofs[1:] = np.cumsum([len(q["isg"]) for q in problems])
        return dict(P=P, ofs=ofs, xw=cat32("xw", 3), obs=cat32("obs", 3), isg=cat32("isg", 1),
                    close=np.ascontiguousarray(np.concatenate([np.asarray(q["close"], np.uint8).reshape(-1) for q in problems])),
                    Tcw=np.ascontiguousarray(np.stack([np.asarray(q["Tcw"], np.float32).reshape(16) for q in problems])),
                    Tcb=np.ascontiguousarray(np.asarray(problems[0]["Tcb"], np.float32).reshape(4, 4)),
                    Tbc=np.ascontiguousarray(np.asarray(problems[0]["Tbc"], np.float32).reshape(4, 4)),
                    state=st64("state", 21), prev=st64("prev", 21), preint=st64("preint", 16), preint_jac=st64("preint_jac", 45),
                    preint_bias=st64("preint_bias", 6), infoI=st64("infoI", 81), infoG=st64("infoG", 9), infoA=st64("infoA", 9),
                    prior_state=st64("prior_state", 21), prior_H=st64("prior_H", 225))

    def PoseInertialOptimizationLastKeyFrameBatch(self, problems, cam, rec_init=False):
        """Many-stream form of PoseInertialOptimizationLastKeyFrame; `problems`: list of dicts (xw, obs, isg, close, Tcw, Tcb, Tbc,
        state, kf, preint, infoI, infoG, infoA).  -> list of dict(state, outlier, H, n, iters), identical to single calls."""
        P = len(problems)
        cat32 = lambda k, w: np.ascontiguousarray(np.concatenate([np.asarray(q[k], np.float32).reshape(-1, w) for q in problems]))  # noqa: E731
        st64 = lambda k, n: np.ascontiguousarray(np.stack([np.asarray(q[k], np.float64).reshape(-1)[:n] for q in problems]))     # noqa: E731
        ofs = np.zeros(P + 1, np.int32)
        ofs[1:] = np.cumsum([len(q["isg"]) for q in problems])
        xw, obs, isg = cat32("xw", 3), cat32("obs", 3), cat32("isg", 1)
        close = np.ascontiguousarray(np.concatenate([np.asarray(q["close"], np.uint8).reshape(-1) for q in problems]))
        Tcw = np.ascontiguousarray(np.stack([np.asarray(q["Tcw"], np.float32).reshape(16) for q in problems]))
        Tcb = np.ascontiguousarray(np.asarray(problems[0]["Tcb"], np.float32).reshape(4, 4))
        Tbc = np.ascontiguousarray(np.asarray(problems[0]["Tbc"], np.float32).reshape(4, 4))
        st = st64("state", 21).copy()
        kf, pre, iI, iG, iA = st64("kf", 21), st64("preint", 16), st64("infoI", 81), st64("infoG", 9), st64("infoA", 9)
        total = int(ofs[-1])
        outl = np.zeros(max(total, 1), np.uint8)
        H = np.zeros((P, 225), np.float64)
        n = np.zeros(P, np.int32)
        iters = np.zeros((P, 4), np.int32)
        rc = load_library().orbx_pose_inertial_optimization_last_keyframe_batch(
            self.ctx.h, P, _p(ofs), _ptr(xw), _ptr(obs), _ptr(isg), _ptr(close), C.byref(cam), _p(Tcw), _p(Tcb), _p(Tbc), _p(st),
            _p(kf), _p(pre), _p(iI), _p(iG), _p(iA), int(rec_init), _p(outl), _p(H), _p(n), _p(iters))
        _check(rc, "orbx_pose_inertial_optimization_last_keyframe_batch")
        return [dict(state=st[p].copy(), outlier=outl[ofs[p]:ofs[p + 1]].copy(), H=H[p].reshape(15, 15), n=int(n[p]), iters=iters[p].copy())
                for p in range(P)]

    def PoseInertialOptimizationLastFrameBatch(self, problems, cam, rec_init=False):
        """Many-stream form: `problems` = list of dicts with the keys of the single call (xw, obs, isg, close, Tcw, state, prev,
        preint, preint_jac, preint_bias, infoI, infoG, infoA, prior_state, prior_H; Tcb/Tbc taken from the first), or the
        result of pack_inertial_lf().  -> list of dict(state, outlier, H, n, iters), identical to single calls."""
        k = problems if isinstance(problems, dict) else self.pack_inertial_lf(problems)
        P, ofs = k["P"], k["ofs"]
        st = k["state"].copy()
        total = int(ofs[-1])
        outl = np.zeros(max(total, 1), np.uint8)
        H = np.zeros((P, 225), np.float64)
        n = np.zeros(P, np.int32)
        iters = np.zeros((P, 4), np.int32)
        rc = load_library().orbx_pose_inertial_optimization_last_frame_batch(
            self.ctx.h, P, _p(ofs), _ptr(k["xw"]), _ptr(k["obs"]), _ptr(k["isg"]), _ptr(k["close"]), C.byref(cam), _p(k["Tcw"]),
            _p(k["Tcb"]), _p(k["Tbc"]), _p(st), _p(k["prev"]), _p(k["preint"]), _p(k["preint_jac"]), _p(k["preint_bias"]),
            _p(k["infoI"]), _p(k["infoG"]), _p(k["infoA"]), _p(k["prior_state"]), _p(k["prior_H"]), int(rec_init), _p(outl), _p(H),
            _p(n), _p(iters))
        _check(rc, "orbx_pose_inertial_optimization_last_frame_batch")
        return [dict(state=st[p].copy(), outlier=outl[ofs[p]:ofs[p + 1]].copy(), H=H[p].reshape(15, 15), n=int(n[p]), iters=iters[p].copy())
                for p in range(P)]


class _PreparedImages:
    """The `const uint8_t* const*` argument of the tracker's host entry points, built once for a list of images."""

    def __init__(self, images):
        self.keep = [np.ascontiguousarray(im, np.uint8) for im in images]
        self.n = len(self.keep)
        self.h, self.w = self.keep[0].shape
        self.ptrs = (C.c_void_p * self.n)(*[im.ctypes.data for im in self.keep])


def prepare_images(images):
    return _PreparedImages(images)


class Tracker:
    """Many-stream tracking replay (orbx_tracker): S stereo streams advance one frame per step()."""

    STAGES = ("extract", "stereo_match", "search_last_frame", "pose_opt_1", "search_local_map", "pose_opt_2")
    STATS = ("nL", "nR", "nStereo", "matches_frame", "inliers_1", "matches_map", "inliers_2", "lm_iters")

    def __init__(self, ctx, ext, S, cam, th_frame=None, th_map=1.0, nnratio_map=0.8, mono=False):
        """mono=True: one image per stream, no stereo matching, th_frame defaults to 15 (src/Tracking.cc:2364-2368) and
        the map must be given (set_map / upload_map) before the first step."""
        self.ctx, self.ext, self.S, self.cam = ctx, ext, S, cam
        self.ips = 1 if mono else 2
        if th_frame is None:
            th_frame = 15.0 if mono else 7.0
        create = load_library().orbx_tracker_create_mono if mono else load_library().orbx_tracker_create
        self.h = create(ctx.h, ext.h, S, C.byref(cam), th_frame, th_map, nnratio_map)
        if not self.h:
            raise OrbxError("orbx_tracker_create: " + load_library().orbx_last_error().decode(errors="replace"))

    def close(self):
        if getattr(self, "h", None):
            load_library().orbx_tracker_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, images, Tcw_true, Tcw_prior):
        """images: [L0, R0, L1, R1, ...] (monocular tracker: [I0, I1, ...]) equally sized uint8 arrays (host)
        -> (Tcw_out[S,4,4], stats[S,8])"""
        imgs = [np.ascontiguousarray(im, np.uint8) for im in images]
        assert len(imgs) == self.ips * self.S
        h, w = imgs[0].shape
        ptrs = (C.c_void_p * len(imgs))(*[im.ctypes.data for im in imgs])
        Tt = np.ascontiguousarray(Tcw_true, np.float32).reshape(self.S, 16)
        Tp = np.ascontiguousarray(Tcw_prior, np.float32).reshape(self.S, 16)
        out = np.empty((self.S, 16), np.float32)
        stats = np.zeros((self.S, 8), np.int32)
        _check(load_library().orbx_tracker_step(self.h, ptrs, w, h, w, _p(Tt), _p(Tp), _p(out), _p(stats)),
               "orbx_tracker_step")
        return out.reshape(self.S, 4, 4), stats

    def submit(self, images, Tcw_true, Tcw_prior):
        """Asynchronous step(): enqueue only (orbx_tracker_submit).  `images` must stay alive until collect()."""
        if not isinstance(images, _PreparedImages):
            images = _PreparedImages(images)
        assert images.n == self.ips * self.S
        Tt = np.ascontiguousarray(Tcw_true, np.float32).reshape(self.S, 16)
        Tp = np.ascontiguousarray(Tcw_prior, np.float32).reshape(self.S, 16)
        _check(load_library().orbx_tracker_submit(self.h, images.ptrs, images.w, images.h, images.w, _p(Tt), _p(Tp)),
               "orbx_tracker_submit")

    def collect(self):
        out = np.empty((self.S, 16), np.float32)
        stats = np.zeros((self.S, 8), np.int32)
        _check(load_library().orbx_tracker_collect(self.h, _p(out), _p(stats)), "orbx_tracker_collect")
        return out.reshape(self.S, 4, 4), stats

    def step_device(self, d_imgs, w, h, stride, d_true, d_prior, d_out, d_stats):
        _check(load_library().orbx_tracker_step_device(self.h, d_imgs, w, h, stride, d_true, d_prior, d_out, d_stats),
               "orbx_tracker_step_device")

    def set_overlap(self, on=True):
        _check(load_library().orbx_tracker_set_overlap(self.h, int(on)), "orbx_tracker_set_overlap")

    @property
    def result_stream(self):
        return load_library().orbx_tracker_result_stream(self.h)

    @property
    def map_capacity(self):
        return load_library().orbx_tracker_map_capacity(self.h)

    @property
    def map_bytes(self):
        return load_library().orbx_tracker_map_bytes(self.h)

    def _track_map(self, fields, log_scale_factor):
        m = abi.TrackMap()
        m.m_cap = self.map_capacity
        m.log_scale_factor = float(log_scale_factor)
        for name, _ in abi.TrackMap.FIELDS:
            setattr(m, name, fields[name])
        return m

    def set_map(self, device_ptrs, log_scale_factor=0.0):
        """device_ptrs: dict name -> device address of the arrays of orbx_track_map (stride map_capacity); None detaches."""
        if device_ptrs is None:
            _check(load_library().orbx_tracker_set_map(self.h, None), "orbx_tracker_set_map")
            return
        m = self._track_map(device_ptrs, log_scale_factor)
        _check(load_library().orbx_tracker_set_map(self.h, C.byref(m)), "orbx_tracker_set_map")

    def upload_map(self, host_arrays, log_scale_factor=0.0):
        """host_arrays: dict name -> numpy array ([S, map_capacity(, 3|32)], dtypes of abi.TrackMap.FIELDS); copied to
        the device for the NEXT step (call right before step / submit)."""
        keep = {}
        for name, dt in abi.TrackMap.FIELDS:
            a = host_arrays[name]
            assert a.dtype == dt and a.flags["C_CONTIGUOUS"], name
            keep[name] = a.ctypes.data
        self._map_keep = host_arrays
        m = self._track_map(keep, log_scale_factor)
        _check(load_library().orbx_tracker_upload_map(self.h, C.byref(m)), "orbx_tracker_upload_map")

    def set_graph(self, on=True):
        """Replay a step whose arguments repeat as ONE CUDA graph launch (single-frame latency path)."""
        _check(load_library().orbx_tracker_set_graph(self.h, int(bool(on))), "orbx_tracker_set_graph")

    @property
    def graph_launches(self):
        return int(load_library().orbx_tracker_graph_launches(self.h))

    # ---- visual-inertial TrackLocalMap (orbx_track_imu) ----
    def _track_imu(self, mode, ptrs, rec_init):
        m = abi.TrackImu()
        m.mode, m.rec_init = int(mode), int(bool(rec_init))
        for name, _, _ in abi.TrackImu.FIELDS:
            setattr(m, name, ptrs.get(name))
        return m

    def set_inertial(self, mode, device_ptrs=None, rec_init=False):
        """mode 0 / None: back to the visual PoseOptimization; 1: PoseInertialOptimizationLastKeyFrame, 2: ...LastFrame.
        device_ptrs: dict name -> device address (abi.TrackImu.FIELDS)."""
        if not mode:
            _check(load_library().orbx_tracker_set_inertial(self.h, None), "orbx_tracker_set_inertial")
            return
        m = self._track_imu(mode, device_ptrs, rec_init)
        _check(load_library().orbx_tracker_set_inertial(self.h, C.byref(m)), "orbx_tracker_set_inertial")

    def upload_inertial(self, mode, host_arrays, rec_init=False):
        """host_arrays: dict name -> numpy array (abi.TrackImu.FIELDS; mode 2 may leave ref_state / prior_state / prior_H out
        to chain on what the previous inertial step left on the device).  Copied for the NEXT step."""
        keep, ptrs = {}, {}
        for name, dt, shape in abi.TrackImu.FIELDS:
            a = host_arrays.get(name)
            if a is None:
                continue
            a = np.ascontiguousarray(a, dt)
            assert a.size == (16 if shape is None else self.S * int(np.prod(shape))), (name, a.shape)
            keep[name] = a
            ptrs[name] = a.ctypes.data
        self._imu_keep = keep
        m = self._track_imu(mode, ptrs, rec_init)
        _check(load_library().orbx_tracker_upload_inertial(self.h, C.byref(m)), "orbx_tracker_upload_inertial")

    def inertial_result(self):
        """-> (state[S,21], H15[S,15,15]) of the last inertial step"""
        st = np.zeros((self.S, 21), np.float64)
        H = np.zeros((self.S, 225), np.float64)
        _check(load_library().orbx_tracker_inertial_result(self.h, _p(st), _p(H)), "orbx_tracker_inertial_result")
        return st, H.reshape(self.S, 15, 15)

    @property
    def inertial_state_dev(self):
        return load_library().orbx_tracker_inertial_state_dev(self.h)

    @property
    def inertial_hessian_dev(self):
        return load_library().orbx_tracker_inertial_hessian_dev(self.h)

    def set_chain(self, enable, d_Tcw_init=None):
        """Motion-model chaining: Tcw_prior of the following steps is the relative motion; d_Tcw_init = device [S,16]."""
        _check(load_library().orbx_tracker_set_chain(self.h, int(bool(enable)), d_Tcw_init), "orbx_tracker_set_chain")

    def set_keyframe_work(self, tri_batch, lba_batch, period):
        """Every `period`-th step enqueue the keyframe-rate plans (TriangulationBatch, LocalBABatch) on a third stream."""
        self._kf = (tri_batch, lba_batch)
        _check(load_library().orbx_tracker_set_keyframe_work(self.h, tri_batch.h if tri_batch else None,
                                                             lba_batch.h if lba_batch else None, int(period)),
               "orbx_tracker_set_keyframe_work")

    @property
    def keyframe_stream(self):
        return load_library().orbx_tracker_keyframe_stream(self.h)

    @property
    def keyframe_runs(self):
        return load_library().orbx_tracker_keyframe_runs(self.h)

    def synchronize(self):
        _check(load_library().orbx_tracker_synchronize(self.h), "orbx_tracker_synchronize")

    def set_profiling(self, on=True):
        _check(load_library().orbx_tracker_set_profiling(self.h, int(on)), "orbx_tracker_set_profiling")

    def stage_ms(self):
        ms = np.zeros(len(self.STAGES), np.float32)
        _check(load_library().orbx_tracker_stage_ms(self.h, _p(ms)), "orbx_tracker_stage_ms")
        return ms


class ORBVocabulary:
    """DBoW2 ORBVocabulary on the device (include/ORBVocabulary.h; Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h).
    `source` is a path to a DBoW2 binary vocabulary (the reference's Vocabulary/ORBvoc.bin) or its bytes."""

    def __init__(self, ctx, source):
        L = load_library()
        self.ctx = ctx
        if isinstance(source, (bytes, bytearray, memoryview, np.ndarray)):
            buf = np.frombuffer(bytes(source), np.uint8)
            self.h = L.orbx_vocabulary_from_memory(ctx.h, _p(buf), buf.size)
        else:
            self.h = L.orbx_vocabulary_load(ctx.h, str(source).encode())
        if not self.h:
            raise OrbxError("orbx_vocabulary: " + L.orbx_last_error().decode(errors="replace"))
        v = (C.c_int * 6)()
        L.orbx_vocabulary_info(self.h, *[C.byref(v, 4 * k) for k in range(6)])
        self.k, self.L, self.n_nodes, self.n_words, self.scoring, self.weighting = [int(x) for x in v]

    def __del__(self):
        try:
            if getattr(self, "h", None):
                load_library().orbx_vocabulary_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def transform(self, desc, levelsup=4):
        """-> (bow_word int32[nb], bow_value float64[nb], fv_node int32[nn], fv_off int32[nn+1], fv_idx int32[...])"""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        bw, bv = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float64)
        fn, fo, fi = np.zeros(max(n, 1), np.int32), np.zeros(n + 1, np.int32), np.zeros(max(n, 1), np.int32)
        nb, nn = C.c_int32(0), C.c_int32(0)
        _check(load_library().orbx_vocabulary_transform(self.h, _p(desc), n, levelsup, _p(bw), _p(bv), C.byref(nb), _p(fn),
                                                        _p(fo), _p(fi), C.byref(nn)), "orbx_vocabulary_transform")
        nb, nn = nb.value, nn.value
        return bw[:nb].copy(), bv[:nb].copy(), fn[:nn].copy(), fo[:nn + 1].copy(), fi[:fo[nn]].copy()


def search_by_bow(ctx, kf, frame, kf_has_mp, fv_kf, fv_f, nnratio=0.7, check_orientation=True):
    """ORBmatcher(nnratio, checkOri).SearchByBoW(pKF, F, vpMapPointMatches) (src/ORBmatcher.cc:323-591).
    fv_* = (node, off, idx) CSR triples.  -> (nmatches, match_f[frame.n])"""
    has = np.ascontiguousarray(kf_has_mp, np.uint8)
    kn, ko, ki = [np.ascontiguousarray(a, np.int32) for a in fv_kf]
    fn, fo, fi = [np.ascontiguousarray(a, np.int32) for a in fv_f]
    out = np.full(max(frame.n, 1), -1, np.int32)
    nm = C.c_int32(0)
    _check(load_library().orbx_search_by_bow(ctx.h, kf.ref(), frame.ref(), _p(has), len(kn), _p(kn), _p(ko), _p(ki), len(fn),
                                             _p(fn), _p(fo), _p(fi), nnratio, int(check_orientation), _p(out), C.byref(nm)),
           "orbx_search_by_bow")
    return nm.value, out[:frame.n]


def fuse(ctx, kf, cam, Rcw, tcw, Ow, flags, xw, max_dist, min_dist, normal, mp_desc, th, scale_factors, inv_level_sigma2,
         log_scale_factor):
    """Search half of ORBmatcher::Fuse(pKF, vpMapPoints, th) (src/ORBmatcher.cc:1630-1883). -> (nFused, best_idx[nmp])"""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)   # noqa: E731
    Rcw, tcw, Ow, xw, max_dist, min_dist, normal = map(f32, (Rcw, tcw, Ow, xw, max_dist, min_dist, normal))
    sf, isg = f32(scale_factors), f32(inv_level_sigma2)
    flags = np.ascontiguousarray(flags, np.uint8)
    mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
    n = len(flags)
    out = np.full(max(n, 1), -1, np.int32)
    nf = C.c_int32(0)
    _check(load_library().orbx_fuse(ctx.h, kf.ref(), C.byref(cam), _p(Rcw), _p(tcw), _p(Ow), n, _p(flags), _p(xw), _p(max_dist),
                                    _p(min_dist), _p(normal), _p(mp_desc), th, _p(sf), _p(isg), len(sf), log_scale_factor,
                                    _p(out), C.byref(nf)), "orbx_fuse")
    return nf.value, out[:n]


def is_in_frustum(ctx, cam, Rcw, tcw, Ow, bounds, cos_limit, nlevels, log_scale_factor, xw, max_dist, min_dist, normal,
                  stale=None):
    """Frame::isInFrustum over a whole local map (src/Frame.cc:571-650).  `stale` = dict(proj_xr, depth, level, view_cos)
    of previous values (kept where the point is not in view).  -> dict of the MapPoint track fields + n."""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)   # noqa: E731
    Rcw, tcw, Ow, xw, max_dist, min_dist, normal = map(f32, (Rcw, tcw, Ow, xw, max_dist, min_dist, normal))
    n = len(max_dist)
    st = stale or {}
    out = dict(in_view=np.zeros(max(n, 1), np.uint8), proj_x=np.zeros(max(n, 1), np.float32), proj_y=np.zeros(max(n, 1), np.float32),
               proj_xr=f32(st.get("proj_xr", np.zeros(max(n, 1)))).copy(), depth=f32(st.get("depth", np.zeros(max(n, 1)))).copy(),
               level=np.ascontiguousarray(st.get("level", np.zeros(max(n, 1))), np.int32).copy(),
               view_cos=f32(st.get("view_cos", np.zeros(max(n, 1)))).copy())
    cnt = C.c_int32(0)
    _check(load_library().orbx_is_in_frustum(ctx.h, C.byref(cam), _p(Rcw), _p(tcw), _p(Ow), bounds[0], bounds[1], bounds[2], bounds[3],
                                             cos_limit, nlevels, log_scale_factor, n, _p(xw), _p(max_dist), _p(min_dist), _p(normal),
                                             _p(out["in_view"]), _p(out["proj_x"]), _p(out["proj_y"]), _p(out["proj_xr"]),
                                             _p(out["depth"]), _p(out["level"]), _p(out["view_cos"]), C.byref(cnt)),
           "orbx_is_in_frustum")
    out = {k: v[:n] for k, v in out.items()}
    out["n"] = cnt.value
    return out


def undistort_keypoints(ctx, xy, cam, dist_coef):
    """Frame::UndistortKeyPoints (src/Frame.cc:874-924) -> float32 [n, 2]"""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    d = np.ascontiguousarray(dist_coef, np.float32).ravel()
    out = np.zeros_like(xy)
    _check(load_library().orbx_undistort_keypoints(ctx.h, _p(xy), len(xy), C.byref(cam), _p(d), len(d), _p(out)),
           "orbx_undistort_keypoints")
    return out
