"""CPU: oracle SearchByBoW / Fuse against straight Python restatements of src/ORBmatcher.cc:323-591
and :1630-1883 that share no code with the oracle."""
import numpy as np
import pytest

import scenarios as sc
from orbx import abi

F32 = np.float32


def ham(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def rot_bin(a, b):
    rot = F32(a) - F32(b)
    if rot < 0:
        rot = F32(rot + F32(360.0))
    v = float(F32(rot * F32(1.0 / 30)))
    b_ = int(np.floor(v + 0.5))          # round() half away from zero, v >= 0
    return 0 if b_ == 30 else b_


def three_maxima(h):
    m1 = m2 = m3 = 0
    i1 = i2 = i3 = -1
    for i, s in enumerate(h):
        if s > m1:
            m3, m2, m1, i3, i2, i1 = m2, m1, s, i2, i1, i
        elif s > m2:
            m3, m2, i3, i2 = m2, s, i2, i
        elif s > m3:
            m3, i3 = s, i
    if F32(m2) < F32(0.1) * F32(m1):
        i2 = i3 = -1
    elif F32(m3) < F32(0.1) * F32(m1):
        i3 = -1
    return i1, i2, i3


def bow_python(s, nnratio, check_ori):
    (kn, ko, ki), (fn, fo, fi) = s["fvK"], s["fvF"]
    fmap = {int(n): (int(fo[j]), int(fo[j + 1])) for j, n in enumerate(fn)}
    match = np.full(len(s["kF"]), -1, np.int32)
    hist = [[] for _ in range(30)]
    n = 0
    for a, node in enumerate(kn):
        if int(node) not in fmap:
            continue
        fb, fe = fmap[int(node)]
        for iK in range(ko[a], ko[a + 1]):
            r = int(ki[iK])
            if not s["has"][r]:
                continue
            b1, b2, bi = 256, 256, -1
            for iF in range(fb, fe):
                j = int(fi[iF])
                if match[j] >= 0:
                    continue
                d = ham(s["dK"][r], s["dF"][j])
                if d < b1:
                    b2, b1, bi = b1, d, j
                elif d < b2:
                    b2 = d
            if b1 <= 50 and F32(b1) < F32(nnratio) * F32(b2):
                match[bi] = r
                if check_ori:
                    hist[rot_bin(s["kK"]["angle"][r], s["kF"]["angle"][bi])].append(bi)
                n += 1
    if check_ori:
        keep = three_maxima([len(h) for h in hist])
        for i in range(30):
            if i in keep:
                continue
            for j in hist[i]:
                match[j] = -1
                n -= 1
    return n, match


@pytest.mark.parametrize("seed,nnratio,ori", [(1, 0.7, True), (2, 0.9, True), (3, 0.6, False), (4, 0.75, True)])
def test_search_by_bow_matches_python(ork, seed, nnratio, ori):
    s = sc.bow_scenario(seed)
    KF, F = abi.Frame(s["kK"], s["dK"]), abi.Frame(s["kF"], s["dF"])
    n, m = ork.search_by_bow(KF, F, s["has"], s["fvK"], s["fvF"], nnratio, ori)
    pn, pm = bow_python(s, nnratio, ori)
    assert n == pn and np.array_equal(m, pm)
    assert n == int((m >= 0).sum()) and n > 40
    got = m[m >= 0]
    assert len(np.unique(got)) <= len(got) and np.all(s["has"][got] == 1)


def test_search_by_bow_edge_cases(ork):
    s = sc.bow_scenario(9, n_kf=50, n_f=40, nwords=6)
    KF, F = abi.Frame(s["kK"], s["dK"]), abi.Frame(s["kF"], s["dF"])
    empty = (np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32))
    assert ork.search_by_bow(KF, F, s["has"], empty, s["fvF"])[0] == 0
    assert ork.search_by_bow(KF, F, s["has"], s["fvK"], empty)[0] == 0
    assert ork.search_by_bow(KF, F, np.zeros_like(s["has"]), s["fvK"], s["fvF"])[0] == 0


def fuse_python(s):
    k, ur = s["kK"], s["ur"]
    R, t, Ow = s["R"], s["t"], s["Ow"]
    fx, fy, cx, cy, bf = map(F32, (sc.FX, sc.FY, sc.CX, sc.CY, sc.BF))
    wInv, hInv = F32(64) / F32(752), F32(48) / F32(480)
    gx = np.floor((k["x"] * wInv).astype(F32) + F32(0.5)).astype(int)
    gy = np.floor((k["y"] * hInv).astype(F32) + F32(0.5)).astype(int)
    best = np.full(len(s["flags"]), -1, np.int32)
    n = 0
    for i in range(len(s["flags"])):
        if not (s["flags"][i] & 1):
            continue
        X, Y, Z = s["xw"][i]
        xc = F32(F32(F32(R[0] * X) + F32(R[1] * Y)) + F32(R[2] * Z)) + t[0]
        yc = F32(F32(F32(R[3] * X) + F32(R[4] * Y)) + F32(R[5] * Z)) + t[1]
        zc = F32(F32(F32(R[6] * X) + F32(R[7] * Y)) + F32(R[8] * Z)) + t[2]
        xc, yc, zc = F32(xc), F32(yc), F32(zc)
        if zc < 0:
            continue
        with np.errstate(divide="ignore", invalid="ignore"):
            invz = F32(1) / zc
            u = F32(F32(F32(fx * xc) / zc) + cx)
            v = F32(F32(F32(fy * yc) / zc) + cy)
        if not (u >= 0 and u < 752 and v >= 0 and v < 480):
            continue
        urp = F32(u - F32(bf * invz))
        maxD, minD = F32(F32(1.2) * s["maxd"][i]), F32(F32(0.8) * s["mind"][i])
        PO = [F32(X - Ow[0]), F32(Y - Ow[1]), F32(Z - Ow[2])]
        d3 = F32(np.sqrt(float(PO[0]) * float(PO[0]) + float(PO[1]) * float(PO[1]) + float(PO[2]) * float(PO[2])))
        if d3 < minD or d3 > maxD:
            continue
        nr = s["normal"][i]
        dot = float(PO[0]) * float(nr[0]) + float(PO[1]) * float(nr[1]) + float(PO[2]) * float(nr[2])
        if dot < 0.5 * float(d3):
            continue
        ratio = F32(s["maxd"][i] / d3)
        lvl = int(np.ceil(np.log(float(ratio)) / float(F32(s["log_sf"]))))
        lvl = 0 if lvl < 0 else (7 if lvl >= 8 else lvl)
        rad = F32(F32(s["th"]) * s["scale"][lvl])
        c = np.flatnonzero((np.abs(k["x"] - u) < rad) & (np.abs(k["y"] - v) < rad) & (gx < 64) & (gy < 48))
        c = sorted(c.tolist(), key=lambda j: (gx[j], gy[j], j))
        bd, bi = 256, -1
        for j in c:
            kl = int(k["octave"][j])
            if kl < lvl - 1 or kl > lvl:
                continue
            ex, ey = F32(u - k["x"][j]), F32(v - k["y"][j])
            if ur is not None and ur[j] >= 0:
                er = F32(urp - ur[j])
                e2 = F32(F32(F32(ex * ex) + F32(ey * ey)) + F32(er * er))
                if float(F32(e2 * s["inv_sigma2"][kl])) > 7.8:
                    continue
            else:
                e2 = F32(F32(ex * ex) + F32(ey * ey))
                if float(F32(e2 * s["inv_sigma2"][kl])) > 5.99:
                    continue
            d = ham(s["desc"][i], s["dK"][j])
            if d < bd:
                bd, bi = d, j
        if bd <= 50:
            best[i] = bi
            n += 1
    return n, best


@pytest.mark.parametrize("seed,stereo,th", [(1, True, 3.0), (2, False, 3.0), (3, True, 4.0), (4, False, 2.5)])
def test_fuse_matches_python(ork, seed, stereo, th):
    s = sc.fuse_scenario(seed, stereo=stereo, th=th)
    KF = abi.Frame(s["kK"], s["dK"], s["ur"])
    cam = abi.make_camera()
    n, b = ork.fuse(KF, cam, s["R"], s["t"], s["Ow"], s["flags"], s["xw"], s["maxd"], s["mind"], s["normal"], s["desc"],
                    s["th"], s["scale"], s["inv_sigma2"], s["log_sf"])
    pn, pb = fuse_python(s)
    assert n == pn and np.array_equal(b, pb)
    assert n == int((b >= 0).sum()) and n > 50
