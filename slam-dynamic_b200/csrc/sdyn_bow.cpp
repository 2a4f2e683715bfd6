/* C-ABI of the bag-of-words transform: vocabulary objects, Frame::ComputeBoW's descent on the device and the host
 * assembly of DBoW2's BowVector / FeatureVector (include/sdyn.h, "Frame::ComputeBoW"). */
#include "bow_internal.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

struct sdyn_vocab {
    int device = 0, nnodes = 0, nwords = 0, k = 0, L = 0;
    sdyn::VocabView view{};
    void* block = nullptr;
    std::string err;
};

namespace sdyn {

int api_fail(sdyn_ctx* c, int code, const std::string& msg);

struct BowState {
    int B = 0, cap = 0;
    uint32_t* dWord = nullptr; double* dWeight = nullptr; uint32_t* dNode = nullptr; uint8_t* dDescIn = nullptr; int descCap = 0;
};

void free_bow_state(sdyn_ctx* c)
{
    BowState* s = static_cast<BowState*>(c->bow);
    if (!s) return;
    cudaFree(s->dWord); cudaFree(s->dWeight); cudaFree(s->dNode); cudaFree(s->dDescIn);
    delete s;
    c->bow = nullptr;
}

static int ensure_bow_state(sdyn_ctx* c, int hostN)
{
    BowState* s = static_cast<BowState*>(c->bow);
    if (!s) {
        s = new BowState();
        s->B = c->maxBatch; s->cap = c->maxKp;
        c->bow = s;
    }
    const size_t need = std::max((size_t)s->B * s->cap, (size_t)std::max(hostN, 0));
    static_assert(sizeof(double) == 8, "double");
    if (!s->dWord || (size_t)s->descCap < need) {
        cudaStreamSynchronize(c->stream);
        cudaFree(s->dWord); cudaFree(s->dWeight); cudaFree(s->dNode); cudaFree(s->dDescIn);
        s->dWord = nullptr; s->dWeight = nullptr; s->dNode = nullptr; s->dDescIn = nullptr; s->descCap = 0;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&s->dWord), need * 4);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&s->dWeight), need * 8);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&s->dNode), need * 4);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&s->dDescIn), need * 32);
        if (e != cudaSuccess) { free_bow_state(c); return api_fail(c, SDYN_ERR_NOMEM, std::string("bow state: ") + cudaGetErrorString(e)); }
        s->descCap = (int)need;
    }
    return SDYN_OK;
}

}  // namespace sdyn

using namespace sdyn;

#define BCU(c, call)                                                                              \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) return api_fail((c), SDYN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

extern "C" {

int sdyn_vocab_create(int device, int nnodes, const int32_t* parent, const uint8_t* isLeaf, const uint8_t* desc,
                      const double* weight, int k, int L, sdyn_vocab** out)
{
    if (!out) return SDYN_ERR_ARG;
    *out = nullptr;
    if (nnodes < 2 || !parent || !isLeaf || !desc || !weight || L < 1 || L > 10) return SDYN_ERR_ARG;
    /* children lists in node-id order = the push_back order of loadFromTextFile (TemplatedVocabulary.h:1389-1391) */
    std::vector<int32_t> off((size_t)nnodes + 1, 0);
    for (int i = 1; i < nnodes; ++i) {
        if (parent[i] < 0 || parent[i] >= i) return SDYN_ERR_ARG;            /* a parent precedes its children in the file */
        ++off[parent[i] + 1];
    }
    for (int i = 0; i < nnodes; ++i) off[i + 1] += off[i];
    std::vector<uint32_t> idx((size_t)nnodes - 1), word((size_t)nnodes, 0);
    std::vector<int32_t> cur(off.begin(), off.end() - 1);
    int nwords = 0;
    for (int i = 1; i < nnodes; ++i) {
        idx[cur[parent[i]]++] = (uint32_t)i;
        if (isLeaf[i]) word[i] = (uint32_t)nwords++;                         /* :1409-1416 */
    }
    if (off[1] == 0) return SDYN_ERR_ARG;                                    /* a root without children: empty vocabulary */
    if (cudaSetDevice(device) != cudaSuccess) return SDYN_ERR_CUDA;
    sdyn_vocab* v = new sdyn_vocab();
    v->device = device; v->nnodes = nnodes; v->nwords = nwords; v->k = k; v->L = L;
    const size_t bOff = 0, bIdx = (off.size() * 4 + 255) / 256 * 256, bDesc = bIdx + (idx.size() * 4 + 255) / 256 * 256;
    const size_t bW = bDesc + ((size_t)nnodes * 32 + 255) / 256 * 256, bWord = bW + ((size_t)nnodes * 8 + 255) / 256 * 256;
    const size_t total = bWord + (size_t)nnodes * 4;
    cudaError_t e = cudaMalloc(&v->block, total);
    uint8_t* b = static_cast<uint8_t*>(v->block);
    if (e == cudaSuccess) e = cudaMemcpy(b + bOff, off.data(), off.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(b + bIdx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(b + bDesc, desc, (size_t)nnodes * 32, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(b + bW, weight, (size_t)nnodes * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(b + bWord, word.data(), (size_t)nnodes * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(v->block); delete v; return e == cudaErrorMemoryAllocation ? SDYN_ERR_NOMEM : SDYN_ERR_CUDA; }
    v->view.nnodes = nnodes; v->view.L = L;
    v->view.childOff = reinterpret_cast<const int32_t*>(b + bOff); v->view.childIdx = reinterpret_cast<const uint32_t*>(b + bIdx);
    v->view.desc = b + bDesc; v->view.weight = reinterpret_cast<const double*>(b + bW);
    v->view.wordOf = reinterpret_cast<const uint32_t*>(b + bWord);
    *out = v;
    return SDYN_OK;
}

int sdyn_vocab_load_text(int device, const char* path, sdyn_vocab** out)
{
    if (!out || !path) return SDYN_ERR_ARG;
    *out = nullptr;
    std::ifstream f(path);
    if (!f.good()) return SDYN_ERR_ARG;
    std::string s;
    std::getline(f, s);
    int k = 0, L = 0, n1 = -1, n2 = -1;
    { std::stringstream ss(s); ss >> k >> L >> n1 >> n2; }
    /* the range check of TemplatedVocabulary.h:1359; only L1_NORM scoring + TF_IDF weighting (ORBvoc) are supported */
    if (k < 0 || k > 20 || L < 1 || L > 10 || n1 != 0 || n2 != 0) return SDYN_ERR_ARG;
    std::vector<int32_t> parent(1, 0); std::vector<uint8_t> leaf(1, 0), desc(32, 0); std::vector<double> w(1, 0.0);
    while (std::getline(f, s)) {
        if (s.empty()) continue;
        std::stringstream ss(s);
        int pid = 0, isLeaf = 0;
        ss >> pid >> isLeaf;
        uint8_t d[32];
        for (int i = 0; i < 32; ++i) { int e = 0; ss >> e; d[i] = (uint8_t)e; }     /* FORB::fromString: 32 decimal bytes */
        double weight = 0;
        ss >> weight;
        if (ss.fail()) return SDYN_ERR_ARG;
        parent.push_back(pid); leaf.push_back(isLeaf > 0); desc.insert(desc.end(), d, d + 32); w.push_back(weight);
    }
    return sdyn_vocab_create(device, (int)parent.size(), parent.data(), leaf.data(), desc.data(), w.data(), k, L, out);
}

int sdyn_vocab_destroy(sdyn_vocab* v)
{
    if (!v) return SDYN_OK;
    cudaSetDevice(v->device);
    cudaFree(v->block);
    delete v;
    return SDYN_OK;
}

int sdyn_vocab_info(const sdyn_vocab* v, int32_t info[4])
{
    if (!v || !info) return SDYN_ERR_ARG;
    info[0] = v->nnodes; info[1] = v->nwords; info[2] = v->k; info[3] = v->L;
    return SDYN_OK;
}

int sdyn_bow_transform_device(sdyn_ctx* c, const sdyn_vocab* voc, int nframes, int levelsup, void* stream)
{
    if (!c) return SDYN_ERR_ARG;
    if (!voc || nframes < 1 || nframes > c->maxBatch || voc->device != c->device)
        return api_fail(c, SDYN_ERR_ARG, "sdyn_bow_transform_device: bad argument");
    BCU(c, cudaSetDevice(c->device));
    int rc = ensure_bow_state(c, 0);
    if (rc != SDYN_OK) return rc;
    BowState* s = static_cast<BowState*>(c->bow);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    StageTimer tm(c, st, SDYN_STAGE_BOW);
    BCU(c, launch_bow_descend(voc->view, c->dDesc, c->dCount, c->maxKp, nframes, levelsup, s->dWord, s->dWeight, s->dNode, st));
    c->launches += 1;
    return SDYN_OK;
}

int sdyn_bow_fetch(sdyn_ctx* c, int nframes, uint32_t* wordId, double* weight, uint32_t* nodeId, int cap, void* stream)
{
    if (!c) return SDYN_ERR_ARG;
    BowState* s = static_cast<BowState*>(c->bow);
    if (!s || nframes < 1 || nframes > s->B || cap < 0) return api_fail(c, SDYN_ERR_ARG, "sdyn_bow_fetch: bad argument");
    BCU(c, cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const size_t m = (size_t)std::min(cap, s->cap);
    if (wordId) BCU(c, cudaMemcpy2DAsync(wordId, (size_t)cap * 4, s->dWord, (size_t)s->cap * 4, m * 4, nframes, cudaMemcpyDeviceToHost, st));
    if (weight) BCU(c, cudaMemcpy2DAsync(weight, (size_t)cap * 8, s->dWeight, (size_t)s->cap * 8, m * 8, nframes, cudaMemcpyDeviceToHost, st));
    if (nodeId) BCU(c, cudaMemcpy2DAsync(nodeId, (size_t)cap * 4, s->dNode, (size_t)s->cap * 4, m * 4, nframes, cudaMemcpyDeviceToHost, st));
    BCU(c, cudaStreamSynchronize(st));
    return SDYN_OK;
}

int sdyn_bow_transform(sdyn_ctx* c, const sdyn_vocab* voc, const uint8_t* desc, int n, int levelsup, uint32_t* wordId,
                       double* weight, uint32_t* nodeId)
{
    if (!c) return SDYN_ERR_ARG;
    if (!voc || n < 0 || (n > 0 && (!desc || !wordId || !weight || !nodeId)) || voc->device != c->device)
        return api_fail(c, SDYN_ERR_ARG, "sdyn_bow_transform: bad argument");
    if (n == 0) return SDYN_OK;
    BCU(c, cudaSetDevice(c->device));
    int rc = ensure_bow_state(c, n);
    if (rc != SDYN_OK) return rc;
    BowState* s = static_cast<BowState*>(c->bow);
    BCU(c, cudaMemcpyAsync(s->dDescIn, desc, (size_t)n * 32, cudaMemcpyHostToDevice, c->stream));
    BCU(c, launch_bow_descend(voc->view, s->dDescIn, nullptr, n, 1, levelsup, s->dWord, s->dWeight, s->dNode, c->stream));
    c->launches += 1;
    BCU(c, cudaMemcpyAsync(wordId, s->dWord, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    BCU(c, cudaMemcpyAsync(weight, s->dWeight, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    BCU(c, cudaMemcpyAsync(nodeId, s->dNode, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    BCU(c, cudaStreamSynchronize(c->stream));
    return SDYN_OK;
}

int sdyn_bow_assemble(const uint32_t* wordId, const double* weight, const uint32_t* nodeId, int n, uint32_t* bowIds,
                      double* bowValues, int* nWords, uint32_t* fvNodes, int32_t* fvOffset, uint32_t* fvIndex, int* nFvNodes)
{
    if (n < 0 || !nWords || !nFvNodes || (n > 0 && (!wordId || !weight || !nodeId || !bowIds || !bowValues || !fvNodes || !fvOffset || !fvIndex)))
        return SDYN_ERR_ARG;
    /* TF_IDF weighting: v.addWeight(id, w) and fv.addFeature(nid, i) for every feature with w > 0, in feature order
     * (TemplatedVocabulary.h:1147-1164); L1 scoring normalises (:1194, BowVector.cpp:64-86) */
    std::map<uint32_t, double> v;
    std::map<uint32_t, std::vector<uint32_t>> fv;
    for (int i = 0; i < n; ++i) {
        if (!(weight[i] > 0)) continue;
        auto it = v.lower_bound(wordId[i]);
        if (it != v.end() && it->first == wordId[i]) it->second += weight[i];
        else v.insert(it, std::make_pair(wordId[i], weight[i]));
        fv[nodeId[i]].push_back((uint32_t)i);
    }
    double norm = 0.0;
    for (auto& kv : v) norm += std::fabs(kv.second);
    if (norm > 0.0)
        for (auto& kv : v) kv.second /= norm;
    int k = 0;
    for (auto& kv : v) { bowIds[k] = kv.first; bowValues[k] = kv.second; ++k; }
    *nWords = k;
    int m = 0, pos = 0;
    if (n > 0) fvOffset[0] = 0;
    for (auto& kv : fv) {
        fvNodes[m] = kv.first;
        for (uint32_t idx : kv.second) fvIndex[pos++] = idx;
        fvOffset[++m] = pos;
    }
    *nFvNodes = m;
    return SDYN_OK;
}

}  // extern "C"
