"""CPU tests of the oracle's DBoW2 vocabulary against an independent numpy restatement, on synthetic
trees in the reference's binary format and — when the reference tree is mounted — on its real Vocabulary/ORBvoc.bin."""
import os
import numpy as np
import pytest

import oracle
import voc_util as vu

REAL = "/root/reference/Vocabulary/ORBvoc.bin"
KEYS = ("bow_word", "bow_value", "fv_node", "fv_off", "fv_idx", "word_id", "node_id")


def _same(a, b):
    for k in KEYS:
        assert a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k   # float64 values bit for bit


@pytest.mark.parametrize("seed,k,L,levelsup,scoring,weighting", [
    (1, 6, 4, 2, 0, 0), (2, 10, 3, 1, 0, 0), (3, 4, 5, 4, 0, 0), (4, 5, 4, 4, 1, 1), (5, 6, 3, 5, 5, 0),
    (6, 6, 4, 2, 2, 2), (7, 3, 6, 4, 0, 3), (8, 6, 4, 2, 5, 3)])
def test_oracle_matches_brute_force(seed, k, L, levelsup, scoring, weighting):
    vb = vu.make_vocabulary(seed, k, L, scoring, weighting)
    V = vu.parse(vb)
    ov = oracle.Vocabulary(vb)
    assert (ov.k, ov.L, ov.n_nodes, ov.scoring, ov.weighting) == (k, L, V["n"], scoring, weighting)
    assert ov.n_words == int((V["word"] >= 0).sum())
    for n in (0, 1, 37, 400):
        q = vu.query_descriptors(V, 100 + seed, n)
        _same(ov.transform(q, levelsup), vu.brute_transform(V, q, levelsup))


def test_bow_vector_properties():
    vb = vu.make_vocabulary(11, 8, 4)
    V = vu.parse(vb)
    r = oracle.Vocabulary(vb).transform(vu.query_descriptors(V, 5, 600), 2)
    assert np.all(np.diff(r["bow_word"]) > 0) and np.all(np.diff(r["fv_node"]) > 0)
    assert abs(r["bow_value"].sum() - 1.0) < 1e-12                       # L1-normalised
    kept = np.flatnonzero(r["word_id"] >= 0)
    assert sorted(r["fv_idx"].tolist()) == kept.tolist()                 # every non-stopped feature exactly once
    for a, b in zip(r["fv_off"][:-1], r["fv_off"][1:]):
        assert np.all(np.diff(r["fv_idx"][a:b]) > 0)                     # ascending inside a node
    assert len(kept) < 600                                               # the generator plants stopped words


@pytest.mark.skipif(not os.path.exists(REAL), reason="reference vocabulary not mounted on this machine")
def test_real_orbvoc_against_brute_force_and_golden():
    vb = open(REAL, "rb").read()
    V = vu.parse(vb)
    ov = oracle.Vocabulary(REAL)
    assert (ov.k, ov.L, ov.n_nodes, ov.n_words, ov.scoring, ov.weighting) == (10, 6, 1082074, 971815, 0, 0)
    sample = np.load(os.path.join(os.path.dirname(__file__), "golden", "orbvoc_sample.npy"))
    rng = np.random.default_rng(0)
    q = sample[rng.choice(len(sample), 300, replace=False)].copy()
    for r in range(100, 300):                                            # 200 of them perturbed by 1..50 bit flips
        for b in rng.integers(0, 256, int(rng.integers(1, 51))):
            q[r, b >> 3] ^= np.uint8(1 << (b & 7))
    got = ov.transform(q, 4)
    _same(got, vu.brute_transform(V, q, 4))
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "orbvoc_transform_golden.npz"))
    assert np.array_equal(gold["desc"], q)
    for k in KEYS:
        assert np.array_equal(gold[k], got[k]), k
