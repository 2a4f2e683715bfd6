// Standalone check of the TMA wrappers in csrc/tma.h: loads a 160 x 38 byte box from a 3-D byte tensor.
#include "../../slam-dynamic_b200/csrc/tma.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace sdyn;

struct Maps { CUtensorMap m[16]; };

__global__ void __launch_bounds__(256) probe_global(const CUtensorMap* maps, int level, int x, int y, int z, uint8_t* out)
{
    __shared__ __align__(128) uint8_t px[38 * 160];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" :: "l"(maps + level) : "memory");
        mbar_expect_tx(&bar, 38 * 160);
        tma_load_3d(px, maps + level, x, y, z, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < 38 * 160; i += 256) out[i] = px[i];
}

__global__ void __launch_bounds__(256) probe_nowait(const __grid_constant__ Maps maps, int level, int x, int y, int z, uint8_t* out)
{
    /* barrier + expect_tx only, no TMA: isolates the mbarrier wrappers (completes via a plain arrive with tx = 0) */
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, 0);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    out[threadIdx.x] = (uint8_t)(level + x + y + z + maps.m[0].opaque[0]);
}

__global__ void __launch_bounds__(256) probe(const __grid_constant__ Maps maps, int level, int x, int y, int z, uint8_t* out)
{
    __shared__ __align__(128) uint8_t px[38 * 160];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, 38 * 160);
        tma_load_3d(px, &maps.m[level], x, y, z, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < 38 * 160; i += 256) out[i] = px[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char** argv)
{
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int pitch = 1312, rows = 414, frames = 3;
    const size_t frameBytes = 1024 * 1024;
    std::vector<uint8_t> h(frameBytes * frames);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 2654435761u >> 24);
    uint8_t* d; CK(cudaMalloc(&d, h.size())); CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
    uint8_t* dout; CK(cudaMalloc(&dout, 38 * 160));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    Maps maps; memset(&maps, 0, sizeof maps);
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)frames};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frameBytes};
    const cuuint32_t box[3] = {160, 38, 1}, estr[3] = {1, 1, 1};
    for (int l = 0; l < 3; ++l) {
        CUresult r = ((EncodeTiledFn)fn)(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d + 256 * l, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode level %d -> %d\n", l, (int)r);
    }
    CUtensorMap* dmaps; CK(cudaMalloc(&dmaps, sizeof maps)); CK(cudaMemcpy(dmaps, &maps, sizeof maps, cudaMemcpyHostToDevice));
    printf("variant %d\n", variant);
    const int cases[][4] = {{0, 43, 31, 0}, {1, 45, 61, 2}, {2, 1300, 400, 1}, {0, 7, 0, 1}};
    for (auto& c : cases) {
        if (variant == 0) probe<<<1, 256>>>(maps, c[0], c[1], c[2], c[3], dout);
        else if (variant == 1) probe_global<<<1, 256>>>(dmaps, c[0], c[1], c[2], c[3], dout);
        else probe_nowait<<<1, 256>>>(maps, c[0], c[1], c[2], c[3], dout);
        cudaError_t e = cudaDeviceSynchronize();
        printf("case level %d x %d y %d z %d: %s\n", c[0], c[1], c[2], c[3], cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<uint8_t> o(38 * 160);
        CK(cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int r = 0; r < 38; ++r)
            for (int k = 0; k < 160; ++k) {
                const int xx = c[1] + k, yy = c[2] + r;
                const uint8_t want = (xx < pitch && yy < rows) ? h[256 * c[0] + (size_t)c[3] * frameBytes + (size_t)yy * pitch + xx] : 0;
                bad += o[r * 160 + k] != want;
            }
        printf("  mismatches: %d\n", bad);
    }
    return 0;
}
