// Variants of the TMA load to find which form the box accepts.
#include "../../slam-dynamic_b200/csrc/tma.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using namespace sdyn;
struct Maps { CUtensorMap m[16]; };

__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, int x, int y, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}

__global__ void __launch_bounds__(256) probe(const __grid_constant__ Maps maps, int mode, int bytes, int x, int y, int z, const uint8_t* src, uint8_t* out, int level, int stat)
{
    extern __shared__ __align__(1024) uint8_t pxd[];
    __shared__ __align__(128) uint8_t pxs[38 * 160];
    uint8_t* px = stat ? pxs : pxd;
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, bytes);
        if (mode == 0) tma_load_3d(px, &maps.m[level], x, y, z, &bar);
        else if (mode == 1) tma_load_2d(px, &maps.m[1], x, y, &bar);
        else asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                          :: "r"(smem_u32(px)), "l"(src), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bytes; i += 256) out[i] = px[i];
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
    uint8_t* dout; CK(cudaMalloc(&dout, 65536));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    Maps maps; memset(&maps, 0, sizeof maps);
    int bw = 160, bh = 38, mode = 0;
    CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    if (variant == 1) { bw = 128; bh = 32; }
    if (variant == 2) { bw = 64; bh = 8; }
    if (variant == 3) l2 = CU_TENSOR_MAP_L2_PROMOTION_NONE;
    if (variant == 4) mode = 1;
    if (variant == 5) { mode = 1; bw = 128; bh = 32; }
    if (variant == 6) mode = 2;
    if (variant == 7) { bw = 256; bh = 16; }
    if (variant == 8) { bw = 16; bh = 38; }
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)frames};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frameBytes};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, estr[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&maps.m[0], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = ((EncodeTiledFn)fn)(&maps.m[1], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d: box %dx%d mode %d encode %d %d\n", variant, bw, bh, mode, (int)r, (int)r2);
    const int bytes = bw * bh, x = variant == 11 ? 43 : variant == 12 ? 44 : variant == 13 ? 40 : 48, y = 31, z = mode == 0 ? 1 : 0;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    if (variant == 9) maps.m[2] = maps.m[0];
    probe<<<1, 256, 16384>>>(maps, mode, bytes, x, y, z, d + 4096, dout, variant == 9 ? 2 : 0, variant == 10);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  run: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint8_t> o(bytes);
    CK(cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int rr = 0; rr < bh; ++rr)
        for (int k = 0; k < bw; ++k) {
            uint8_t want = mode == 2 ? h[4096 + rr * bw + k] : h[(size_t)z * frameBytes + (size_t)(y + rr) * pitch + x + k];
            bad += o[rr * bw + k] != want;
        }
    printf("  mismatches: %d\n", bad);
    return 0;
}
