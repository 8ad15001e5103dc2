"""Synthetic DBoW2 vocabularies in the reference's binary format (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:
1442-1478) plus a numpy/pure-Python restatement of `transform` that shares no code with the oracle or the device path.
Test infrastructure only."""
import struct
import numpy as np


def make_vocabulary(seed, k=6, L=4, scoring=0, weighting=0, p_early_leaf=0.08, p_stop=0.1, ragged=True):
    """-> bytes.  Tree built depth-first like DBoW2's HKmeansStep (children of a node get consecutive ids); child
    descriptors are the parent's with ~40 random bit flips, so descents are meaningful and ties do occur."""
    rng = np.random.default_rng(seed)
    nodes = []   # (parent, desc, weight, leaf)

    def grow(parent_id, parent_desc, depth):
        nch = k if not ragged else int(rng.integers(2, k + 1))
        ids = []
        for _ in range(nch):
            d = parent_desc.copy()
            flips = rng.integers(0, 256, int(rng.integers(20, 60)))
            for b in flips:
                d[b >> 3] ^= np.uint8(1 << (b & 7))
            leaf = depth == L or (depth >= 2 and rng.random() < p_early_leaf)
            w = 0.0
            if leaf:
                w = 0.0 if rng.random() < p_stop else float(np.float32(rng.uniform(0.5, 9.7)))
            nodes.append([parent_id, d, w, leaf])
            ids.append(len(nodes))          # node ids start at 1
        for cid in ids:
            if not nodes[cid - 1][3]:
                grow(cid, nodes[cid - 1][1], depth + 1)

    # DBoW2 creates all k children first, then recurses into each: reproduce that id order
    grow(0, rng.integers(0, 256, 32).astype(np.uint8), 1)
    out = [struct.pack("<IIiiii", len(nodes) + 1, 41, k, L, scoring, weighting)]
    for parent, d, w, leaf in nodes:
        out.append(struct.pack("<i", parent) + d.tobytes() + struct.pack("<f", w) + (b"\x01" if leaf else b"\x00"))
    return b"".join(out)


def parse(vocbytes):
    nb, sz, k, L, sc, we = struct.unpack("<IIiiii", vocbytes[:24])
    rec = np.frombuffer(vocbytes[24:], np.uint8)
    rec = rec[:len(rec) // sz * sz].reshape(-1, sz)
    parent = rec[:, :4].copy().view("<i4").ravel()
    desc = rec[:, 4:36].copy()
    weight = rec[:, 36:40].copy().view("<f4").ravel().astype(np.float64)
    leaf = rec[:, 40] != 0
    n = len(rec) + 1
    children = [[] for _ in range(n)]
    for i, p in enumerate(parent):
        children[p].append(i + 1)
    word = np.full(n, -1)
    word[1:][leaf] = np.arange(int(leaf.sum()))
    D = np.zeros((n, 32), np.uint8)
    D[1:] = desc
    W = np.zeros(n)
    W[1:] = weight
    return dict(k=k, L=L, scoring=sc, weighting=we, children=children, desc=D, weight=W, word=word, n=n)


_POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def brute_transform(V, feats, levelsup):
    """transform(features, BowVector, FeatureVector, levelsup): TemplatedVocabulary.h:1140-1219 / :1231-1271."""
    bow, fv = {}, {}
    words, nids = [], []
    tf = V["weighting"] in (0, 1)
    must, l2 = V["scoring"] != 5, V["scoring"] == 1
    nid_level = V["L"] - levelsup
    for i, f in enumerate(np.asarray(feats, np.uint8).reshape(-1, 32)):
        cur, level, nid = 0, 0, 0
        while V["children"][cur]:
            level += 1
            ch = V["children"][cur]
            d = _POP[V["desc"][ch] ^ f].sum(1)
            cur = ch[int(np.argmin(d))]          # argmin returns the first minimum, like `d < best_d`
            if level == nid_level:
                nid = cur
        w = float(V["weight"][cur])
        words.append(int(V["word"][cur]) if w > 0 else -1)
        nids.append(nid)
        if w > 0:
            wid = int(V["word"][cur])
            if tf:
                bow[wid] = bow[wid] + w if wid in bow else w
            elif wid not in bow:
                bow[wid] = w
            fv.setdefault(nid, []).append(i)
    keys = sorted(bow)
    vals = [bow[kk] for kk in keys]
    if tf and keys and not must:
        vals = [v / float(len(keys)) for v in vals]
    if must:
        norm = 0.0
        if not l2:
            for v in vals:
                norm += abs(v)
        else:
            for v in vals:
                norm += v * v
            norm = float(np.sqrt(norm))
        if norm > 0.0:
            vals = [v / norm for v in vals]
    fkeys = sorted(fv)
    off = [0]
    idx = []
    for kk in fkeys:
        idx += fv[kk]
        off.append(len(idx))
    return dict(word_id=np.array(words, np.int32), node_id=np.array(nids, np.int32), bow_word=np.array(keys, np.int32),
                bow_value=np.array(vals, np.float64), fv_node=np.array(fkeys, np.int32), fv_off=np.array(off, np.int32),
                fv_idx=np.array(idx, np.int32))


def query_descriptors(V, seed, n, max_flips=60):
    """n descriptors near random leaves of the tree (0..max_flips bit flips) — realistic descents with ties."""
    rng = np.random.default_rng(seed)
    leaves = np.flatnonzero(V["word"] >= 0)
    out = V["desc"][rng.choice(leaves, n)].copy()
    for r in range(n):
        for b in rng.integers(0, 256, int(rng.integers(0, max_flips + 1))):
            out[r, b >> 3] ^= np.uint8(1 << (b & 7))
    return out
