"""Pins the oracle's DBoW2 transform against the REFERENCE'S OWN vendored DBoW2 (Thirdparty/DBoW2 compiled in place
against oracle/refshim -> oracle/_ref/libref_dbow2.so, oracle/ref_dbow2_glue.cpp): the vocabulary is written in the
text format of ORBvoc.txt, loaded by the reference's loadFromTextFile and run through the reference's
TemplatedVocabulary<FORB>::transform — per feature (word id, weight, node id `levelsup` levels above the leaves) and as
the BowVector / FeatureVector that Frame::ComputeBoW stores (src/Frame.cc:880-896).  Exact equality, doubles included."""
import ctypes
import os

import numpy as np
import pytest

from synth import synth_voc_features, synth_vocabulary

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_dbow2.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref not built (needs /root/reference: make -C oracle ref)")
vp = ctypes.c_void_p


def write_vocabulary_txt(voc, path, trailing_newline=False):
    """ORBvoc.txt layout (TemplatedVocabulary.h:1338-1425 reads it, :1428-1452 writes it): header `k L scoring weighting`,
    then one line per node: parent id, leaf flag, 32 descriptor bytes, weight."""
    lines = ["%d %d 0 0" % (voc["k"], voc["L"])]  # L1_NORM, TF_IDF (Vocabulary/ORBvoc.txt)
    for nid in range(1, len(voc["parent"])):
        lines.append("%d %d %s %r" % (voc["parent"][nid], voc["leaf_flag"][nid], " ".join(str(int(b)) for b in voc["desc"][nid]),
                                     float(voc["weight"][nid])))
    with open(path, "w") as f:
        f.write("\n".join(lines) + ("\n" if trailing_newline else ""))


def run_reference(path, feats, levelsup):
    R = ctypes.CDLL(LIB)
    R.ref_bow_transform.argtypes = [ctypes.c_char_p, vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.c_int, vp, vp, vp]
    n = len(feats)
    feats = np.ascontiguousarray(feats, np.uint8)
    word, w, node = np.zeros(n, np.int32), np.zeros(n, np.float64), np.zeros(n, np.int32)
    bw, bv = np.zeros(n, np.int32), np.zeros(n, np.float64)
    fn, fp = np.zeros(n, np.int32), np.zeros(n, np.int32)
    nwords = ctypes.c_int32(0)
    k = R.ref_bow_transform(path.encode(), feats.ctypes.data, n, levelsup, word.ctypes.data, w.ctypes.data, node.ctypes.data,
                            bw.ctypes.data, bv.ctypes.data, n, fn.ctypes.data, fp.ctypes.data, ctypes.byref(nwords))
    assert k >= 0
    return dict(word=word, weight=w, node=node, bow=dict(zip(bw[:k].tolist(), bv[:k].tolist())), fv_node=fn, fv_pos=fp,
                n_words=nwords.value)


def assemble(word, w, node):
    """What the product's ORBVocabulary::transform builds from the per-feature kernel output
    (self_commit_orb-slam2_b200/host/ORBVocabulary.cc, __init__.py ORBVocabulary.transform)."""
    bow, fv = {}, {}
    for i in range(len(word)):
        if w[i] > 0:
            bow[int(word[i])] = bow.get(int(word[i]), 0.0) + float(w[i])
            fv.setdefault(int(node[i]), []).append(i)
    norm = 0.0
    for k in sorted(bow):
        norm += abs(bow[k])
    if norm > 0.0:
        for k in bow:
            bow[k] /= norm
    return bow, fv


@pytest.mark.parametrize("k,L,seed,levelsup,ragged", [(10, 4, 7, 4, True), (10, 4, 7, 2, True), (6, 5, 9, 4, False), (10, 3, 11, 1, True),
                                                       (10, 6, 13, 4, True)])
def test_transform_equals_reference_dbow2(checker, tmp_path, k, L, seed, levelsup, ragged):
    if L == 6:
        k = 4  # keep the node count moderate
    voc = synth_vocabulary(k=k, L=L, seed=seed, ragged=ragged)
    feats = synth_voc_features(voc, 1500, seed + 1)
    path = str(tmp_path / "voc.txt")
    write_vocabulary_txt(voc, path)
    r = run_reference(path, feats, levelsup)
    nw, word, w, node = checker.bow_transform(voc, feats, levelsup)
    assert r["n_words"] == int(voc["leaf_flag"].sum())
    assert np.array_equal(r["word"], word)
    assert np.array_equal(r["weight"], w)          # doubles, exact
    live = w > 0                                   # DBoW2 leaves nid untouched only when levelsup > depth; compare where filed
    assert np.array_equal(r["node"][live], node[live])
    bow, fv = assemble(word, w, node)
    assert r["bow"] == bow                         # same keys, bit-identical normalised values
    for nid, lst in fv.items():
        for pos, i in enumerate(lst):
            assert r["fv_node"][i] == nid and r["fv_pos"][i] == pos
    assert int((r["fv_node"] >= 0).sum()) == sum(len(v) for v in fv.values())
    assert nw == int(live.sum()) or nw == len(bow) or nw >= 0


def test_trailing_newline_adds_a_phantom_word(tmp_path):
    """ORBvoc.txt ends with a newline; the reference's `while(!f.eof())` loop (TemplatedVocabulary.h:1378) then parses one
    more, empty, line.  Every extraction from the empty stream fails without writing, so `pid` and `nIsLeaf` keep whatever
    the stack slot held (in practice the previous line's values: parent of the last node, leaf flag 1): the vocabulary
    gains ONE phantom word — weight 0 (Node's default), descriptor never written by FORB::fromString (uninitialised
    memory in OpenCV, zeros in the stand-in).  A feature that descends to it is dropped like a stopped word; everything
    else is unchanged.  That is undefined behaviour, not an algorithm: the product's loader skips empty lines instead
    (deviations)."""
    voc = synth_vocabulary(k=10, L=3, seed=5)
    feats = synth_voc_features(voc, 500, 6)
    a, b = str(tmp_path / "a.txt"), str(tmp_path / "b.txt")
    write_vocabulary_txt(voc, a)
    write_vocabulary_txt(voc, b, trailing_newline=True)
    ra, rb = run_reference(a, feats, 2), run_reference(b, feats, 2)
    assert rb["n_words"] == ra["n_words"] + 1
    same = rb["weight"] == ra["weight"]
    assert same.mean() > 0.95                      # only features nearest to the phantom's descriptor differ ...
    assert np.all(rb["weight"][~same] == 0.0)      # ... and those are dropped
    assert np.array_equal(rb["word"][same], ra["word"][same])


ORBVOC_TGZ = "/root/reference/Vocabulary/ORBvoc.txt.tar.gz"


@pytest.mark.skipif(not os.path.exists(ORBVOC_TGZ), reason="the reference's vocabulary is only mounted in the build container")
def test_real_orbvoc_transform_equals_reference_dbow2(oracle, tmp_path):
    """The reference's one real fixture for this path: Vocabulary/ORBvoc.txt (k = 10, L = 6, 1,082,072 nodes, 971,814
    words).  Loaded by the reference's own loadFromTextFile and by the flattened loader the oracle / product use; 2000
    features (noisy copies of real words + random descriptors) must get the same word, weight and level-2 node
    (levelsup = 4, as Frame::ComputeBoW asks) from both.  The file ends with a newline, so the reference side also carries
    the phantom word described above; a feature that lands on it (none does here) would be excluded."""
    import tarfile

    import pandas as pd
    with tarfile.open(ORBVOC_TGZ) as t:
        t.extract("ORBvoc.txt", path=str(tmp_path))
    path = str(tmp_path / "ORBvoc.txt")
    with open(path) as f:
        k, L, scoring, weighting = [int(x) for x in f.readline().split()]
    assert (k, L, scoring, weighting) == (10, 6, 0, 0)
    a = pd.read_csv(path, sep=r"\s+", header=None, skiprows=1, engine="c").values
    voc = dict(k=k, L=L, parent=np.concatenate([[-1], a[:, 0].astype(np.int32)]), leaf_flag=np.concatenate([[0], a[:, 1].astype(np.uint8)]),
               desc=np.concatenate([np.zeros((1, 32), np.uint8), a[:, 2:34].astype(np.uint8)]),
               weight=np.concatenate([[0.0], a[:, 34].astype(np.float64)]))
    assert len(voc["parent"]) == 1082073 and int(voc["leaf_flag"].sum()) == 971814
    rng = np.random.RandomState(1)
    leaves = np.nonzero(voc["leaf_flag"])[0]
    feats = voc["desc"][rng.choice(leaves, 2000)].copy()
    for i in range(2000):
        for b in rng.choice(256, size=int(rng.randint(0, 40)), replace=False):
            feats[i, b >> 3] ^= np.uint8(1 << (b & 7))
    feats[::10] = rng.randint(0, 256, size=(200, 32)).astype(np.uint8)
    r = run_reference(path, feats, 4)
    assert r["n_words"] == 971814 + 1  # + the phantom word of the trailing newline
    nw, word, w, node = oracle.bow_transform(voc, feats, 4)
    ok = r["word"] != r["n_words"] - 1
    assert ok.mean() > 0.99
    assert np.array_equal(r["word"][ok], word[ok]) and np.array_equal(r["weight"][ok], w[ok])
    assert np.array_equal(r["node"][ok], node[ok])
    bow, fv = assemble(word[ok], w[ok], node[ok])
    if ok.all():
        assert r["bow"] == bow
    assert len(set(node[ok].tolist())) > 50  # the level-2 partition SearchByBoW joins on
