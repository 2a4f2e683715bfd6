// Measures the B200's sustained __popc (POPC.B32) issue rate: the denominator of the matcher's "popc pipe" roofline.
// 8 independent accumulators per thread, 4096 iterations, every SM saturated.  Prints G popc/s.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned* out, unsigned seed, int iters)
{
    unsigned a0 = threadIdx.x ^ seed, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u, a4 = a0 * 11u, a5 = a0 * 13u, a6 = a0 * 17u, a7 = a0 * 19u;
    unsigned s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;
    for (int i = 0; i < iters; ++i) {
        s0 += __popc(a0 ^ s1); s1 += __popc(a1 ^ s2); s2 += __popc(a2 ^ s3); s3 += __popc(a3 ^ s4);
        s4 += __popc(a4 ^ s5); s5 += __popc(a5 ^ s6); s6 += __popc(a6 ^ s7); s7 += __popc(a7 ^ s0);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3 + s4 + s5 + s6 + s7;
}
int main()
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    unsigned* d; cudaMalloc(&d, (size_t)blocks * threads * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int w = 0; w < 3; ++w) k<<<blocks, threads>>>(d, w, iters);
    cudaEventRecord(a);
    const int reps = 10;
    for (int r = 0; r < reps; ++r) k<<<blocks, threads>>>(d, r, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    const double ops = (double)reps * blocks * threads * iters * 8;
    printf("{\"sms\": %d, \"popc_gops\": %.1f, \"per_sm_per_clk_at_1965MHz\": %.2f}\n", sms, ops / ms / 1e6, ops / ms / 1e6 / sms / 1.965);
    return 0;
}
