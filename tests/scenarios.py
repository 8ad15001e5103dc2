"""Seeded synthetic matcher / optimiser scenarios, shared by CPU and GPU tests.
Real descriptors come from the extractor run on synthetic images and from tests/golden/orbvoc_sample.npy
(4096 genuine ORB descriptors sampled from the reference's Vocabulary/ORBvoc.bin)."""
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FX, FY, CX, CY, BF = 458.654, 457.296, 367.215, 248.375, 47.9   # Examples/Monocular/EuRoC.yaml:9-12


def orbvoc():
    return np.load(os.path.join(HERE, "golden", "orbvoc_sample.npy"))


def flip_bits(rng, desc, nflip):
    d = np.unpackbits(desc.copy())
    pos = rng.choice(256, size=nflip, replace=False)
    d[pos] ^= 1
    return np.packbits(d)


def extract_frame(ork, img, nfeatures=1000, lap=(0, 0)):
    ex = ork.Extractor(nfeatures)
    rc, k, d, m = ex(img, lap)
    assert rc == 0
    return ex, k, d


def sbp_map_scenario(seed, kps, desc, uright=None, nq=1500, th=1.0):
    rng = np.random.default_rng(seed)
    voc = orbvoc()
    n = len(kps)
    scale = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    projX, projY, projXR = np.zeros(nq, np.float32), np.zeros(nq, np.float32), np.zeros(nq, np.float32)
    level, viewCos = np.zeros(nq, np.int32), np.zeros(nq, np.float32)
    mpDesc, flags = np.zeros((nq, 32), np.uint8), np.zeros(nq, np.uint8)
    for q in range(nq):
        if rng.random() < 0.7 and n > 0:
            i = int(rng.integers(n))
            mpDesc[q] = flip_bits(rng, desc[i], int(rng.integers(0, 41)))
            projX[q] = kps["x"][i] + rng.normal(0, 2.0)
            projY[q] = kps["y"][i] + rng.normal(0, 2.0)
            level[q] = int(np.clip(kps["octave"][i] + rng.choice([0, 0, 0, 1]), 0, 7))
            if uright is not None and uright[i] > 0:
                projXR[q] = uright[i] + rng.normal(0, 1.5)
            else:
                projXR[q] = projX[q] - rng.uniform(3, 40)
        else:
            mpDesc[q] = voc[int(rng.integers(len(voc)))]
            projX[q], projY[q] = rng.uniform(0, 752), rng.uniform(0, 480)
            level[q] = int(rng.integers(0, 8))
            projXR[q] = projX[q] - rng.uniform(3, 40)
        viewCos[q] = rng.choice([0.9995, 0.9985, 0.95, 0.998])
        flags[q] = (1 if rng.random() < 0.9 else 0) | (2 if rng.random() < 0.85 else 0)
    blocked = (rng.random(n) < 0.05).astype(np.uint8)
    return dict(kp_blocked=blocked, projX=projX, projY=projY, projXR=projXR, level=level, viewCos=viewCos,
                mpDesc=mpDesc, flags=flags, th=th, scaleFactors=scale)


def rot_small(rng, deg):
    w = rng.normal(0, 1, 3)
    w = w / np.linalg.norm(w) * np.deg2rad(deg)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def sbp_frame_scenario(seed, kps, desc, uright=None, depth=None, rot_deg=0.6, trans=0.03, forward=0.0):
    """Last frame == the given keypoints observed at identity pose; current pose slightly perturbed."""
    rng = np.random.default_rng(seed)
    n = len(kps)
    z = np.where((depth is not None) & (np.asarray(depth if depth is not None else np.zeros(n)) > 0),
                 depth if depth is not None else 0, rng.uniform(1.5, 12.0, n)).astype(np.float64)
    X = np.stack([(kps["x"] - CX) / FX * z, (kps["y"] - CY) / FY * z, z], 1)
    Tl = np.eye(4, dtype=np.float32)
    R = rot_small(rng, rot_deg)
    t = rng.normal(0, trans, 3)
    t[2] += forward
    Tc = np.eye(4)
    Tc[:3, :3], Tc[:3, 3] = R, t
    flags = ((rng.random(n) < 0.7).astype(np.uint8)) | ((rng.random(n) < 0.9).astype(np.uint8) << 1)
    voc = orbvoc()
    mpDesc = np.stack([flip_bits(rng, desc[i], int(rng.integers(0, 31))) if rng.random() < 0.9
                       else voc[int(rng.integers(len(voc)))] for i in range(n)]) if n else np.zeros((0, 32), np.uint8)
    angle = (kps["angle"] + np.where(rng.random(n) < 0.85, rng.normal(3.0, 2.0, n), rng.uniform(0, 360, n))) % 360.0
    blocked = (rng.random(n) < 0.03).astype(np.uint8)
    return dict(cur_blocked=blocked, Tcw_cur=Tc.astype(np.float32), Tcw_last=Tl, flags=flags,
                xw=X.astype(np.float32), octave=kps["octave"].astype(np.int32), angle=angle.astype(np.float32),
                mpDesc=mpDesc, scaleFactors=(np.float32(1.2) ** np.arange(8)).astype(np.float32))


def feature_vector(desc, words, id_stride=7):
    """Stand-in for DBoW2::FeatureVector: bucket = nearest of `words` (Hamming); CSR (node ids, offsets, idx)."""
    if len(desc) == 0:
        return np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32)
    d = np.unpackbits(desc, axis=1).astype(np.int16)
    w = np.unpackbits(words, axis=1).astype(np.int16)
    ham = (d[:, None, :] != w[None, :, :]).sum(2)
    node = ham.argmin(1)
    ids = np.unique(node)
    off, idx = [0], []
    for nid in ids:
        members = np.flatnonzero(node == nid)
        idx.extend(members.tolist())
        off.append(len(idx))
    return (ids * id_stride + 3).astype(np.int32), np.array(off, np.int32), np.array(idx, np.int32)


def tri_scenario(seed, kps, desc, uright=None, baseline=0.25, rot_deg=2.0, nwords=48):
    """KF1 = given keypoints at identity; KF2 = their reprojection under a known relative pose (+noise,
    +distractors, shuffled)."""
    rng = np.random.default_rng(seed)
    voc = orbvoc()
    n = len(kps)
    z = rng.uniform(2.0, 12.0, n)
    X = np.stack([(kps["x"] - CX) / FX * z, (kps["y"] - CY) / FY * z, z], 1)
    R2 = rot_small(rng, rot_deg)
    t2 = np.array([-baseline, rng.normal(0, 0.02), rng.normal(0, 0.05)])
    Xc2 = X @ R2.T + t2
    u2 = FX * Xc2[:, 0] / Xc2[:, 2] + CX + rng.normal(0, 0.6, n)
    v2 = FY * Xc2[:, 1] / Xc2[:, 2] + CY + rng.normal(0, 0.6, n)
    ok = (u2 > 20) & (u2 < 732) & (v2 > 20) & (v2 < 460) & (rng.random(n) < 0.85)
    k2 = kps[ok].copy()
    k2["x"], k2["y"] = u2[ok].astype(np.float32), v2[ok].astype(np.float32)
    k2["angle"] = ((k2["angle"] + np.where(rng.random(len(k2)) < 0.9, rng.normal(-4, 2, len(k2)),
                                            rng.uniform(0, 360, len(k2)))) % 360).astype(np.float32)
    d2 = np.stack([flip_bits(rng, d, int(rng.integers(0, 36))) for d in desc[ok]]) if ok.any() else np.zeros((0, 32), np.uint8)
    nd = 150
    kd = np.zeros(nd, kps.dtype)
    kd["x"], kd["y"] = rng.uniform(20, 732, nd), rng.uniform(20, 460, nd)
    kd["octave"], kd["angle"] = rng.integers(0, 8, nd), rng.uniform(0, 360, nd)
    dd = voc[rng.integers(0, len(voc), nd)]
    k2 = np.concatenate([k2, kd])
    d2 = np.concatenate([d2, dd])
    perm = rng.permutation(len(k2))
    k2, d2 = k2[perm], np.ascontiguousarray(d2[perm])
    words = voc[rng.choice(len(voc), nwords, replace=False)]
    ur1 = uright
    ur2 = np.where(rng.random(len(k2)) < 0.3, k2["x"] - rng.uniform(2, 30, len(k2)), -1).astype(np.float32) \
        if uright is not None else None
    return dict(k1=kps, d1=desc, ur1=ur1, k2=k2, d2=d2, ur2=ur2,
                has1=(rng.random(n) < 0.3).astype(np.uint8), has2=(rng.random(len(k2)) < 0.3).astype(np.uint8),
                fv1=feature_vector(desc, words), fv2=feature_vector(d2, words),
                R1w=np.eye(3, dtype=np.float32).ravel(), t1w=np.zeros(3, np.float32),
                R2w=R2.astype(np.float32).ravel(), t2w=t2.astype(np.float32),
                sigma2=((np.float32(1.2) ** np.arange(8)) ** 2).astype(np.float32),
                scaleFactors=(np.float32(1.2) ** np.arange(8)).astype(np.float32))
