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
