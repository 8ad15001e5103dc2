"""Second, independent restatement of ORBextractor::operator() (src/ORBextractor.cc:1074-1156) in
Python, composed from the *real* OpenCV primitives of cv2 4.13 (cv2.resize, per-cell
cv2.FastFeatureDetector, cv2.GaussianBlur, cv2.fastAtan2) plus a pure-Python quadtree.

Purpose: pin the C++ oracle (oracle/) end to end.  The oracle restates the OpenCV primitives as
integer formulas; this module calls OpenCV itself, so agreement of the two on whole images pins
both the primitives *in composition* (cell tiling, ROI semantics, threshold fallback) and the
octree/ordering logic (two implementations written separately).  Slow; used on a few images only,
and to mint tests/golden/*.npz (tools/make_golden.py).
"""
import math
import numpy as np
import cv2

EDGE = 19
HALF_PATCH = 15


def f32(x):
    return np.float32(x)


def cv_round(x):
    return int(np.rint(x))  # round-half-even


def tables(nfeatures, scale_factor, nlevels):
    sf = float(np.float32(scale_factor))  # double member initialised from a float
    scale = [np.float32(1.0)]
    for i in range(1, nlevels):
        scale.append(np.float32(float(scale[-1]) * sf))
    inv = [np.float32(1.0) / s for s in scale]
    factor = np.float32(1.0 / sf)
    ndes = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(math.pow(float(factor), float(nlevels))))
    nf, tot = [], 0
    for _ in range(nlevels - 1):
        nf.append(cv_round(ndes))
        tot += nf[-1]
        ndes = f32(ndes * factor)
    nf.append(max(nfeatures - tot, 0))
    umax = [0] * (HALF_PATCH + 1)
    vmax = int(math.floor(float(f32(HALF_PATCH) * f32(math.sqrt(2.0)) / f32(2) + f32(1))))
    vmin = int(math.ceil(float(f32(HALF_PATCH) * f32(math.sqrt(2.0)) / f32(2))))
    for v in range(vmax + 1):
        umax[v] = cv_round(math.sqrt(HALF_PATCH * HALF_PATCH - v * v))
    v0 = 0
    for v in range(HALF_PATCH, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return scale, inv, nf, umax


class _Node:
    __slots__ = ("x0", "x1", "y0", "y1", "keys", "no_more", "seq")

    def __init__(self, x0, x1, y0, y1):
        self.x0, self.x1, self.y0, self.y1 = x0, x1, y0, y1
        self.keys, self.no_more, self.seq = [], False, 0


def _divide(n, K):
    hx = int(math.ceil(float(f32(n.x1 - n.x0) / f32(2))))
    hy = int(math.ceil(float(f32(n.y1 - n.y0) / f32(2))))
    mx, my = n.x0 + hx, n.y0 + hy
    ch = [_Node(n.x0, mx, n.y0, my), _Node(mx, n.x1, n.y0, my), _Node(n.x0, mx, my, n.y1),
          _Node(mx, n.x1, my, n.y1)]
    for k in n.keys:
        x, y = K[k][0], K[k][1]
        ch[(0 if x < mx else 1) + (0 if y < my else 2)].keys.append(k)
    for c in ch:
        if len(c.keys) == 1:
            c.no_more = True
    return ch


def distribute_octree(K, minX, maxX, minY, maxY, N):
    """K: list of (x, y, response).  Returns indices into K in the reference's output order.
    A Python list plays std::list: index 0 is the front."""
    n_ini = int(math.floor(float(f32(maxX - minX) / f32(maxY - minY)) + 0.5))
    hX = f32(maxX - minX) / f32(n_ini)
    seq = 0
    L = []
    for i in range(n_ini):
        n = _Node(int(hX * f32(i)), int(hX * f32(i + 1)), 0, maxY - minY)
        n.seq = seq
        seq += 1
        L.append(n)
    ini = list(L)
    for k, kp in enumerate(K):
        ini[int(f32(kp[0]) / hX)].keys.append(k)
    for n in L:
        if len(n.keys) == 1:
            n.no_more = True
    L = [n for n in L if n.keys]
    finish = False
    while not finish:
        prev_size = len(L)
        n_to_expand = 0
        pending = []
        front = []  # nodes pushed to the front during this pass, in creation order
        keep = []
        for n in L:
            if n.no_more:
                keep.append(n)
                continue
            for c in _divide(n, K):
                if c.keys:
                    c.seq = seq
                    seq += 1
                    front.append(c)
                    if len(c.keys) > 1:
                        n_to_expand += 1
                        pending.append(c)
        L = front[::-1] + keep
        if len(L) >= N or len(L) == prev_size:
            finish = True
        elif len(L) + n_to_expand * 3 > N:
            while not finish:
                prev_size = len(L)
                prev = sorted(pending, key=lambda c: (len(c.keys), c.seq))
                pending = []
                for n in reversed(prev):
                    new = []
                    for c in _divide(n, K):
                        if c.keys:
                            c.seq = seq
                            seq += 1
                            new.append(c)
                            if len(c.keys) > 1:
                                pending.append(c)
                    L = new[::-1] + [m for m in L if m is not n]
                    if len(L) >= N:
                        break
                if len(L) >= N or len(L) == prev_size:
                    finish = True
    out = []
    for n in L:
        best = n.keys[0]
        for k in n.keys[1:]:
            if K[k][2] > K[best][2]:
                best = k
        out.append(best)
    return out


def ic_angle(img, x, y, umax):
    c = img.astype(np.int64)
    m10 = sum(u * int(c[y, x + u]) for u in range(-HALF_PATCH, HALF_PATCH + 1))
    m01 = 0
    for v in range(1, HALF_PATCH + 1):
        d = umax[v]
        us = np.arange(-d, d + 1)
        p, m = c[y + v, x + us], c[y - v, x + us]
        m01 += v * int((p - m).sum())
        m10 += int((us * (p + m)).sum())
    return np.float32(cv2.fastAtan2(float(np.float32(m01)), float(np.float32(m10))))


def descriptor(blur, x, y, angle_deg, pattern):
    factor_pi = np.float32(math.pi / 180.0)
    ang = np.float32(angle_deg) * factor_pi
    a, b = np.float32(math.cos(float(ang))), np.float32(math.sin(float(ang)))
    px, py = pattern[:, 0].astype(np.float32), pattern[:, 1].astype(np.float32)
    rr = np.rint(px * b + py * a).astype(np.int64)  # separate fp32 mul/add: no FMA
    cc = np.rint(px * a - py * b).astype(np.int64)
    vals = blur[y + rr, x + cc].astype(np.int32)
    bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
    return np.packbits(bits.reshape(32, 8), axis=1, bitorder="little").ravel()


def load_pattern(path):
    txt = "".join(l for l in open(path) if not l.lstrip().startswith("//"))
    v = np.array([int(t) for t in txt.replace("\n", "").split(",") if t.strip()], np.int32)
    assert v.size == 1024
    return v.reshape(512, 2)


def extract(img, pattern, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, lap=(0, 0),
            return_debug=False):
    cv2.setNumThreads(1)
    scale, inv, nfl, umax = tables(nfeatures, scale_factor, nlevels)
    H, W = img.shape
    pyr = [img]
    for l in range(1, nlevels):
        sz = (cv_round(f32(W) * inv[l]), cv_round(f32(H) * inv[l]))
        pyr.append(cv2.resize(pyr[-1], sz, interpolation=cv2.INTER_LINEAR))
    det = {t: cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True,
                                             type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16) for t in (ini_th, min_th)}
    all_kp, cands = [], []
    for l in range(nlevels):
        im = pyr[l]
        minB = EDGE - 3
        maxBX, maxBY = im.shape[1] - EDGE + 3, im.shape[0] - EDGE + 3
        width, height = f32(maxBX - minB), f32(maxBY - minB)
        nC, nR = int(width / f32(30)), int(height / f32(30))
        wC, hC = int(math.ceil(float(width / f32(nC)))), int(math.ceil(float(height / f32(nR))))
        K = []
        for i in range(nR):
            iy = minB + i * hC
            my = iy + hC + 6
            if iy >= maxBY - 3:
                continue
            my = min(my, maxBY)
            for j in range(nC):
                ix = minB + j * wC
                mx = ix + wC + 6
                if ix >= maxBX - 6:
                    continue
                mx = min(mx, maxBX)
                roi = im[iy:my, ix:mx]
                kps = det[ini_th].detect(roi)
                if not kps:
                    kps = det[min_th].detect(roi)
                for p in kps:
                    K.append((p.pt[0] + j * wC, p.pt[1] + i * hC, p.response))
        cands.append(K)
        sel = distribute_octree(K, minB, maxBX, minB, maxBY, nfl[l])
        size = float(int(f32(31) * scale[l]))
        kl = []
        for k in sel:
            x, y = int(K[k][0]) + minB, int(K[k][1]) + minB
            kl.append([x, y, size, ic_angle(im, x, y, umax), K[k][2], l])
        all_kp.append(kl)
    n = sum(len(k) for k in all_kp)
    out_k = np.zeros(n, dtype=[("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                               ("response", "<f4"), ("octave", "<i4")])
    out_d = np.zeros((n, 32), np.uint8)
    mono, stereo = 0, n - 1
    for l in range(nlevels):
        if not all_kp[l]:
            continue
        blur = cv2.GaussianBlur(pyr[l].copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        for (x, y, size, ang, resp, octv) in all_kp[l]:
            d = descriptor(blur, x, y, ang, pattern)
            fx, fy = f32(x), f32(y)
            if l != 0:
                fx, fy = fx * scale[l], fy * scale[l]
            if fx >= lap[0] and fx <= lap[1]:
                slot = stereo
                stereo -= 1
            else:
                slot = mono
                mono += 1
            out_k[slot] = (fx, fy, size, ang, resp, octv)
            out_d[slot] = d
    if return_debug:
        return out_k, out_d, mono, pyr, cands
    return out_k, out_d, mono
