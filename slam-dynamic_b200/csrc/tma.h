/* Tensor Memory Accelerator plumbing (sm_100a): tiles of the pyramid are brought into shared memory by ONE
 * cp.async.bulk.tensor issued by one thread and signalled on an mbarrier, instead of a per-thread load / shift / store
 * loop.  The tensor maps (one per pyramid level: x = byte in the padded row, y = bordered row, z = frame) are encoded
 * on the host with cuTensorMapEncodeTiled and passed to the kernel as a __grid_constant__ parameter; out-of-range parts
 * of a box are zero-filled by the hardware, so edge tiles need no clamping. */
#pragma once
#include "../../include/sdyn.h"
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace sdyn {

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(phase) : "memory");
}

/* 3-D tile load: box (set in the tensor map) at element coordinates (x, y, z) -> dst (128-byte aligned shared memory) */
__device__ __forceinline__ void tma_load_3d(void* dst, const void* tmap, int x, int y, int z, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}
#endif

/* one tensor map per pyramid level, passed by value as a kernel parameter */
struct LevelMaps { CUtensorMap m[SDYN_MAX_LEVELS]; };

/* the tensor maps of a context (host copies, re-encoded when the geometry changes).  The x coordinate of a box must be a
 * multiple of 16 bytes (the TMA unit faults otherwise): the blur tiles (128-column grid), the patch origins of the
 * descriptor stage and the resize source rectangles are 16-byte aligned by construction, and the FAST box starts at the
 * aligned column at or before the tile (the tile grid keeps the remainder a multiple of 4). */
struct TmaMaps {
    LevelMaps blurTile;     /* over the pyramid: kBlurStageW x kBlurStageH (tile + 7x7 halo) */
    LevelMaps fastTile;     /* over the pyramid: kFastStageW x kFastStageH (tile + NMS halo + ring) */
    LevelMaps orientPatch;  /* over the pyramid: kPatchPitch x 31 rows (IC_Angle disc) */
    LevelMaps descPatch;    /* over the blurred pyramid: kPatchPitch x 37 rows (rotated pattern reach) */
    LevelMaps resizeSrc;    /* m[l], l >= 1: over level l-1 of the pyramid, kResizePitch x rsRows(l) (source rectangle of a tile) */
};

/* host: encodes the maps of boxW x boxH x 1 byte boxes over a pyramid buffer of maxBatch frames */
struct Geom;
cudaError_t encode_level_maps(const Geom& g, const uint8_t* dBuffer, int maxBatch, int boxW, int boxH, LevelMaps* out,
                              const char** why);
/* m[l] = boxes of kResizePitch x L[l].rsRows bytes over level l-1 (levels whose window is wider keep a zeroed map) */
cudaError_t encode_resize_maps(const Geom& g, const uint8_t* dBuffer, int maxBatch, LevelMaps* out, const char** why);

}  // namespace sdyn
