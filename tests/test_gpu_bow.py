"""Parity of the device-side bag-of-words transform (Frame::ComputeBoW -> DBoW2 TemplatedVocabulary::transform,
src/Frame.cc:803-810) through the C ABI against the oracle.  Bar: word ids, node ids, weights, the normalised
BowVector values (doubles) and the FeatureVector are bit-exact."""
import numpy as np
import pytest

import common
import orc
import pysdyn
import scenario

pytestmark = pytest.mark.gpu


def same(got, ref):
    for key in ("word", "node", "bow_ids", "fv_nodes", "fv_offset", "fv_index"):
        assert np.array_equal(got[key], ref[key]), key
    assert np.array_equal(got["weight"].view(np.uint64), ref["weight"].view(np.uint64))
    assert np.array_equal(got["bow_values"].view(np.uint64), ref["bow_values"].view(np.uint64))


@pytest.mark.parametrize("k,L,levelsup,ragged", [(10, 4, 2, 0.0), (10, 3, 4, 0.0), (7, 5, 4, 0.2), (3, 6, 1, 0.3)])
def test_transform_host_descriptors(k, L, levelsup, ragged, tmp_path):
    ex = pysdyn.Extractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480)
    kp, d = ex(common.frame("tum", 0))
    parent, leaf, desc, weight = scenario.synthetic_vocabulary(k, L, seed=k, ragged=ragged, base_desc=d[0])
    feats = np.concatenate([d, scenario.flip_bits(desc[-200:], np.random.default_rng(1), np.full(200, 2))])
    ref = orc.Vocabulary(parent, leaf, desc, weight, k, L).transform(feats, levelsup)
    voc = pysdyn.Vocabulary(parent, leaf, desc, weight, k, L)
    assert voc.nnodes == len(parent) and voc.nwords == int(leaf.sum())
    same(pysdyn.bow_transform(ex, voc, feats, levelsup), ref)
    assert len(ref["bow_ids"]) > 20 and len(ref["fv_nodes"]) >= 1
    # the same vocabulary through the ORBvoc.txt text format
    path = tmp_path / "voc.txt"
    scenario.write_vocabulary_text(path, parent, leaf, desc, weight, k, L)
    voc2 = pysdyn.Vocabulary(path=path)
    assert (voc2.nnodes, voc2.nwords, voc2.k, voc2.L) == (voc.nnodes, voc.nwords, k, L)
    same(pysdyn.bow_transform(ex, voc2, feats, levelsup), ref)
    assert len(pysdyn.bow_transform(ex, voc, feats[:0], levelsup)["bow_ids"]) == 0
    voc.close(); voc2.close(); ex.close()


def test_transform_device_resident_batch():
    import torch
    cfg, B = "kitti", 3
    W, H, _, nf, ini, mn = common.CONFIGS[cfg]
    frames = np.stack([common.frame(cfg, i) for i in range(B)])
    ex = pysdyn.Extractor(nf, 1.2, 8, ini, mn, max_width=W, max_height=H, max_batch=B)
    dfr = torch.from_numpy(frames).cuda()
    ex.extract_batch_device(dfr.data_ptr(), B, W * H, W, H, W)
    kps, desc, counts = ex.fetch(B)
    parent, leaf, vdesc, weight = scenario.synthetic_vocabulary(10, 4, seed=2, base_desc=desc[0, 0])
    voc = pysdyn.Vocabulary(parent, leaf, vdesc, weight, 10, 4)
    ovoc = orc.Vocabulary(parent, leaf, vdesc, weight, 10, 4)
    pysdyn.bow_transform_device(ex, voc, B, 4)
    word, w, node = pysdyn.bow_fetch(ex, B)
    for f in range(B):
        n = counts[f]
        ref = ovoc.transform(desc[f, :n], 4)
        got = dict(word=word[f, :n], weight=w[f, :n], node=node[f, :n])
        got.update(pysdyn.bow_assemble(word[f, :n], w[f, :n], node[f, :n]))
        same(got, ref)
    voc.close(); ex.close()
