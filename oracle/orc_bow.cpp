/* ORACLE — TEST INFRASTRUCTURE ONLY.  DBoW2 vocabulary tree and TemplatedVocabulary::transform as Frame::ComputeBoW
 * uses them (reference: src/Frame.cc:803-810; Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1258, 1338-1425;
 * BowVector.cpp:36-86; FeatureVector.cpp; FORB.cpp:81-101), restated with the same containers (std::map, std::vector).
 * DBoW2 is vendored in the reference tree but needs OpenCV C++ to compile, so it cannot be built here; parity pin:
 * tests/test_oracle_bow.py (independent pure-Python re-statement on synthetic trees, incl. ragged ones and stopped words). */
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

namespace orc {

struct VocNode { int parent = 0; std::vector<unsigned> children; uint8_t desc[32]; double weight = 0; unsigned word_id = 0; };

struct Vocabulary {
    int k = 0, L = 0;
    std::vector<VocNode> nodes;
    int nwords = 0;
};

/* FORB::distance — the bit-twiddling population count of FORB.cpp:81-101 */
static int forb_distance(const uint8_t* a, const uint8_t* b)
{
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t x, y;
        std::memcpy(&x, a + 4 * i, 4); std::memcpy(&y, b + 4 * i, 4);
        unsigned v = x ^ y;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

/* loadFromTextFile's tree construction (:1376-1420) from already-parsed rows */
Vocabulary* vocabulary_build(int nnodes, const int32_t* parent, const uint8_t* isLeaf, const uint8_t* desc, const double* weight, int k, int L)
{
    Vocabulary* V = new Vocabulary();
    V->k = k; V->L = L;
    V->nodes.resize(1);
    for (int nid = 1; nid < nnodes; ++nid) {
        V->nodes.resize(V->nodes.size() + 1);
        V->nodes[nid].parent = parent[nid];
        V->nodes[parent[nid]].children.push_back(nid);
        std::memcpy(V->nodes[nid].desc, desc + 32 * (size_t)nid, 32);
        V->nodes[nid].weight = weight[nid];
        if (isLeaf[nid] > 0) V->nodes[nid].word_id = V->nwords++;
    }
    return V;
}

/* transform(feature, word_id, weight, nid, levelsup), :1210-1258 */
static void transform_one(const Vocabulary& V, const uint8_t* feature, unsigned& word_id, double& weight, unsigned* nid, int levelsup)
{
    const int nid_level = V.L - levelsup;
    if (nid_level <= 0 && nid) *nid = 0;
    unsigned final_id = 0;
    int current_level = 0;
    do {
        ++current_level;
        const std::vector<unsigned>& nodes = V.nodes[final_id].children;
        final_id = nodes[0];
        double best_d = forb_distance(feature, V.nodes[final_id].desc);
        for (size_t j = 1; j < nodes.size(); ++j) {
            const unsigned id = nodes[j];
            const double d = forb_distance(feature, V.nodes[id].desc);
            if (d < best_d) { best_d = d; final_id = id; }
        }
        if (nid && current_level == nid_level) *nid = final_id;
    } while (!V.nodes[final_id].children.empty());
    word_id = V.nodes[final_id].word_id;
    weight = V.nodes[final_id].weight;
}

/* transform(features, BowVector&, FeatureVector&, levelsup) with TF_IDF weighting and L1 scoring, :1127-1195 */
void vocabulary_transform(const Vocabulary& V, const uint8_t* features, int n, int levelsup,
                          std::map<unsigned, double>& v, std::map<unsigned, std::vector<unsigned>>& fv,
                          unsigned* wordOut, double* weightOut, unsigned* nodeOut)
{
    v.clear(); fv.clear();
    for (int i = 0; i < n; ++i) {
        unsigned id = 0, nid = 0; double w = 0;
        transform_one(V, features + 32 * (size_t)i, id, w, &nid, levelsup);
        if (wordOut) { wordOut[i] = id; weightOut[i] = w; nodeOut[i] = nid; }
        if (w > 0) {
            /* BowVector::addWeight, BowVector.cpp:36-48 */
            auto vit = v.lower_bound(id);
            if (vit != v.end() && !(v.key_comp()(id, vit->first))) vit->second += w;
            else v.insert(vit, std::make_pair(id, w));
            /* FeatureVector::addFeature */
            auto fit = fv.lower_bound(nid);
            if (fit != fv.end() && fit->first == nid) fit->second.push_back(i);
            else { fit = fv.insert(fit, std::make_pair(nid, std::vector<unsigned>())); fit->second.push_back(i); }
        }
    }
    /* must normalise with L1 (L1Scoring::mustNormalize); BowVector::normalize, BowVector.cpp:64-86 */
    double norm = 0.0;
    for (auto it = v.begin(); it != v.end(); ++it) norm += std::fabs(it->second);
    if (norm > 0.0)
        for (auto it = v.begin(); it != v.end(); ++it) it->second /= norm;
}

}  // namespace orc

extern "C" {

void* orc_vocab_build(int nnodes, const int32_t* parent, const uint8_t* isLeaf, const uint8_t* desc, const double* weight, int k, int L)
{ return orc::vocabulary_build(nnodes, parent, isLeaf, desc, weight, k, L); }
void orc_vocab_free(void* v) { delete (orc::Vocabulary*)v; }

int orc_bow_transform(void* voc, const uint8_t* features, int n, int levelsup, unsigned* wordId, double* weight, unsigned* nodeId,
                      unsigned* bowIds, double* bowValues, int* nWords, unsigned* fvNodes, int* fvOffset, unsigned* fvIndex, int* nFvNodes)
{
    std::map<unsigned, double> v; std::map<unsigned, std::vector<unsigned>> fv;
    orc::vocabulary_transform(*(orc::Vocabulary*)voc, features, n, levelsup, v, fv, wordId, weight, nodeId);
    int k = 0;
    for (auto& kv : v) { bowIds[k] = kv.first; bowValues[k] = kv.second; ++k; }
    *nWords = k;
    int m = 0, pos = 0;
    fvOffset[0] = 0;
    for (auto& kv : fv) { fvNodes[m] = kv.first; for (unsigned i : kv.second) fvIndex[pos++] = i; fvOffset[++m] = pos; }
    *nFvNodes = m;
    return 0;
}

}
