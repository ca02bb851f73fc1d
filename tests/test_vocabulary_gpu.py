"""GPU parity of DBoW2 TemplatedVocabulary<FORB>::transform (TemplatedVocabulary.h:1127-1256): word id,
weight and FeatureVector node id per feature bit-exact against the oracle; BowVector / FeatureVector composition."""
import numpy as np
import pytest

from synth import synth_voc_features, synth_vocabulary

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,L,levelsup,ragged", [(10, 4, 2, True), (10, 3, 4, False), (7, 5, 3, True), (10, 6, 4, False)])
def test_bow_transform(pkg, oracle, k, L, levelsup, ragged):
    voc = synth_vocabulary(k=k, L=min(L, 4) if k == 10 and L == 6 else L, seed=3 + L, ragged=ragged)
    if k == 10 and L == 6:  # deeper than the tree: node id stays at its level, leaves are reached earlier
        voc["L"] = 6
    feats = synth_voc_features(voc, n=2000, seed=5 + k)
    v = pkg.ORBVocabulary(voc["k"], voc["L"], voc["parent"], voc["leaf_flag"], voc["desc"], voc["weight"])
    word, w, node = v.transform_features(feats, levelsup)
    nw, oword, ow, onode = oracle.bow_transform(voc, feats, levelsup)
    assert np.array_equal(word, oword) and np.array_equal(w, ow) and np.array_equal(node, onode)
    assert nw == int(voc["leaf_flag"].sum())
    assert len(np.unique(word)) > 50


def test_bow_vector_composition(pkg, oracle):
    voc = synth_vocabulary(k=10, L=4, seed=11)
    feats = synth_voc_features(voc, n=1500, seed=2)
    v = pkg.ORBVocabulary(voc["k"], voc["L"], voc["parent"], voc["leaf_flag"], voc["desc"], voc["weight"])
    bow, fv = v.transform(feats, levelsup=2)
    _, oword, ow, onode = oracle.bow_transform(voc, feats, 2)
    assert abs(sum(bow.values()) - 1.0) < 1e-12  # L1-normalised
    kept = ow > 0
    assert set(bow) == set(oword[kept].tolist())
    assert sum(len(x) for x in fv.values()) == int(kept.sum())
    for nid, idxs in fv.items():
        assert idxs == sorted(idxs) and all(onode[i] == nid for i in idxs)


def test_bow_transform_feeds_search_by_bow(pkg, oracle):
    """The node ids of the transform are what SearchByBoW consumes: a matcher run on transform output equals the oracle."""
    voc = synth_vocabulary(k=10, L=3, seed=21, ragged=False)
    A = synth_voc_features(voc, n=1200, seed=31)
    rng = np.random.RandomState(1)
    B = A[rng.permutation(len(A))].copy()
    for i in range(len(B)):
        for b in rng.choice(256, size=int(rng.randint(0, 25)), replace=False):
            B[i, b >> 3] ^= np.uint8(1 << (b & 7))
    v = pkg.ORBVocabulary(voc["k"], voc["L"], voc["parent"], voc["leaf_flag"], voc["desc"], voc["weight"])
    _, _, nA = v.transform_features(A, levelsup=2)
    _, _, nB = v.transform_features(B, levelsup=2)
    vA = np.ones(len(A), np.uint8)
    aA = rng.uniform(0, 360, len(A)).astype(np.float32)
    aB = rng.uniform(0, 360, len(B)).astype(np.float32)
    m = pkg.ORBmatcher(0.7, False)
    n, match = m.SearchByBoW(A, nA, vA, aA, B, nB, aB)
    on, om = oracle.search_by_bow(A, nA, vA, aA, B, nB, aB, nnratio=0.7, check_ori=False)
    assert n == on and np.array_equal(match, om) and n > 200
