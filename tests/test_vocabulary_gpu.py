"""GPU parity of the device DBoW2 vocabulary: orbx_vocabulary_transform == oracle, bit for bit
(word ids, node ids, float64 BowVector values), on synthetic trees in the reference's binary format."""
import numpy as np
import pytest

import oracle
import orbx
import voc_util as vu

pytestmark = pytest.mark.gpu
KEYS = ("bow_word", "bow_value", "fv_node", "fv_off", "fv_idx")


@pytest.mark.parametrize("seed,k,L,levelsup,scoring,weighting", [
    (1, 6, 4, 2, 0, 0), (2, 10, 3, 1, 0, 0), (3, 4, 5, 4, 0, 0), (4, 5, 4, 4, 1, 1), (5, 6, 3, 5, 5, 0),
    (6, 6, 4, 2, 2, 2), (7, 3, 6, 4, 0, 3), (8, 6, 4, 2, 5, 3), (9, 10, 4, 2, 0, 0)])
def test_transform_matches_oracle(ctx, seed, k, L, levelsup, scoring, weighting):
    vb = vu.make_vocabulary(seed, k, L, scoring, weighting)
    V = vu.parse(vb)
    ov = oracle.Vocabulary(vb)
    dv = orbx.ORBVocabulary(ctx, vb)
    assert (dv.k, dv.L, dv.n_nodes, dv.n_words, dv.scoring, dv.weighting) == (ov.k, ov.L, ov.n_nodes, ov.n_words, ov.scoring, ov.weighting)
    for n in (0, 1, 33, 1000, 5000):
        q = vu.query_descriptors(V, 100 + seed, n)
        want = ov.transform(q, levelsup)
        got = dv.transform(q, levelsup)
        for name, g in zip(KEYS, got):
            assert g.shape == want[name].shape and np.array_equal(g, want[name]), (name, n)


def test_transform_feeds_search_for_triangulation_layout(ctx):
    """The FeatureVector CSR is exactly the layout orbx_search_for_triangulation consumes: ascending node ids, every
    kept feature once, ascending indices inside a node."""
    vb = vu.make_vocabulary(21, 10, 5)
    V = vu.parse(vb)
    q = vu.query_descriptors(V, 7, 1200)
    _, bv, fn, fo, fi = orbx.ORBVocabulary(ctx, vb).transform(q, 3)
    assert np.all(np.diff(fn) > 0) and fo[0] == 0 and fo[-1] == len(fi)
    assert len(np.unique(fi)) == len(fi)
    for a, b in zip(fo[:-1], fo[1:]):
        assert np.all(np.diff(fi[a:b]) > 0)
    assert abs(bv.sum() - 1.0) < 1e-12


def test_batch_device_matches_single(ctx):
    import torch
    vb = vu.make_vocabulary(31, 8, 4)
    V = vu.parse(vb)
    dv = orbx.ORBVocabulary(ctx, vb)
    F, cap = 5, 700
    ns = [700, 0, 123, 699, 1]
    desc = np.zeros((F, cap, 32), np.uint8)
    for f, n in enumerate(ns):
        desc[f, :n] = vu.query_descriptors(V, 50 + f, n)
    t = lambda a: torch.from_numpy(a).cuda()   # noqa: E731
    d_desc, d_n = t(desc), t(np.array(ns, np.int32))
    i32 = lambda *s: torch.zeros(s, dtype=torch.int32, device="cuda")   # noqa: E731
    leaf, node, bw, fn, fi = i32(F, cap), i32(F, cap), i32(F, cap), i32(F, cap), i32(F, cap)
    fo, nb, nn = i32(F, cap + 1), i32(F), i32(F)
    bv = torch.zeros((F, cap), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()   # the zero-fills above run on torch's stream, the transform on the context's
    L = orbx.load_library()
    rc = L.orbx_vocabulary_transform_batch_device(dv.h, F, d_desc.data_ptr(), d_n.data_ptr(), cap, 2, leaf.data_ptr(),
                                                  node.data_ptr(), bw.data_ptr(), bv.data_ptr(), nb.data_ptr(), fn.data_ptr(),
                                                  fo.data_ptr(), fi.data_ptr(), nn.data_ptr())
    assert rc == 0
    ctx.synchronize()
    torch.cuda.synchronize()
    for f, n in enumerate(ns):
        w_bw, w_bv, w_fn, w_fo, w_fi = dv.transform(desc[f, :n], 2)
        k, m = int(nb[f]), int(nn[f])
        assert k == len(w_bw) and m == len(w_fn)
        assert np.array_equal(bw[f, :k].cpu().numpy(), w_bw) and np.array_equal(bv[f, :k].cpu().numpy(), w_bv)
        assert np.array_equal(fn[f, :m].cpu().numpy(), w_fn) and np.array_equal(fo[f, :m + 1].cpu().numpy(), w_fo)
        assert np.array_equal(fi[f, :len(w_fi)].cpu().numpy(), w_fi)
