// occ_ubench.cu -- how much of the butterfly peak survives at the occupancies the NTT kernels actually run at?
// A register-resident loop of the library's own butterfly (16 elements per thread, 4 stages of 8 independent
// butterflies, like one register round of pass 2) is launched with 128-thread CTAs and a dynamic shared-memory
// request that caps the resident CTAs per SM.  Development tool, not part of the library.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../mkhe_kklss_b200/csrc -o occ_ubench occ_ubench.cu
#include <cstdio>
#include <cstdlib>
#include "mkhe_arith.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// EXTRA independent ALU instructions (LOP3 / IADD3 alternating) per butterfly: do ALU-pipe instructions delay the multiplier pipe?
template <int EXTRA>
__global__ void __launch_bounds__(128, 4) k_extra(u64 *sink, ModC m, const ulonglong2 *tw, int iters) {
    extern __shared__ unsigned char sm[];
    u64 v[16];
    u32 e[8];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = (u64)(threadIdx.x * 16 + k + blockIdx.x) % m.q;
#pragma unroll
    for (int k = 0; k < 8; k++) e[k] = threadIdx.x * 7 + k;
    const NttC c = nttc(m);
    ulonglong2 w[4];
#pragma unroll
    for (int k = 0; k < 4; k++) w[k] = tw[k];
    const u32 z = (u32)m.q;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int b = 3; b >= 0; b--) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int g = i >> b, j = i & ((1 << b) - 1);
                bf_fwd(v[(g << (b + 1)) + j], v[(g << (b + 1)) + (1 << b) + j], w[b].x, w[b].y, c);
#pragma unroll
                for (int x = 0; x < EXTRA; x++) {
                    if (x & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(e[(i + x) & 7]) : "r"(z), "r"(e[(i + x + 3) & 7]));
                    else asm volatile("add.u32 %0, %0, %1;" : "+r"(e[(i + x) & 7]) : "r"(z));
                }
            }
        }
    }
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) s ^= v[k];
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= e[k];
    if (s == 0x123456789abcdefull) sink[0] = s + sm[0];
}

template <int MAXREG_CTAS>
__global__ void __launch_bounds__(128, MAXREG_CTAS) k_round(u64 *sink, ModC m, const ulonglong2 *tw, int iters) {
    extern __shared__ unsigned char sm[];
    u64 v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = (u64)(threadIdx.x * 16 + k + blockIdx.x) % m.q;
    const NttC c = nttc(m);
    ulonglong2 w[4];
#pragma unroll
    for (int k = 0; k < 4; k++) w[k] = tw[k];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int b = 3; b >= 0; b--) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int g = i >> b, j = i & ((1 << b) - 1);
                bf_fwd(v[(g << (b + 1)) + j], v[(g << (b + 1)) + (1 << b) + j], w[b].x, w[b].y, c);
            }
        }
    }
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) s ^= v[k];
    if (s == 0x123456789abcdefull) sink[0] = s + sm[0];
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const int sms = prop.multiProcessorCount;
    u64 *sink; CK(cudaMalloc(&sink, 64));
    ulonglong2 htw[4], *tw; CK(cudaMalloc(&tw, sizeof htw));
    ModC m; memset(&m, 0, sizeof m);
    m.q = 0x3fffffffd60001ull; m.nq = 0 - m.q;
    for (int k = 0; k < 4; k++) { u64 w = (0x123456789abcdef1ull * (k + 1)) % m.q; htw[k].x = w; htw[k].y = (u64)((((unsigned __int128)w) << 64) / m.q); }
    CK(cudaMemcpy(tw, htw, sizeof htw, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_round<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_round<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaFuncAttributes fa4, fa8; CK(cudaFuncGetAttributes(&fa4, k_round<4>)); CK(cudaFuncGetAttributes(&fa8, k_round<8>));
    printf("device %s, %d SMs, %.0f MHz; regs: minb4 %d, minb8 %d\n", prop.name, sms, clk_khz / 1e3, fa4.numRegs, fa8.numRegs);
    const int iters = 1500;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int ctas_list[] = {1, 2, 3, 4, 5, 6, 8};
    for (int variant = 0; variant < 2; variant++)
    for (int ci = 0; ci < 7; ci++) {
        const int ctas = ctas_list[ci];
        if (variant == 0 && ctas > 4) continue;
        const size_t smem = (size_t)(200 * 1024 / ctas) & ~(size_t)1023;
        const int blocks = sms * ctas * 4;
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaEventRecord(e0));
            if (variant == 0) k_round<4><<<blocks, 128, smem>>>(sink, m, tw, iters);
            else k_round<8><<<blocks, 128, smem>>>(sink, m, tw, iters);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep && ms < best) best = ms;
        }
        CK(cudaGetLastError());
        const double bfs = (double)blocks * 128 * iters * 32.0;
        const double cyc = best * 1e-3 * (clk_khz * 1e3) * sms * 4 / (bfs / 32);
        printf("minb%d  %d CTAs/SM = %2d warps/SM: %8.3f ms  %.3e bf/s  %6.2f SMSP-cycles per warp-butterfly\n", variant ? 8 : 4, ctas, ctas * 4, best, bfs / (best * 1e-3), cyc);
    }
#define RUN_EXTRA(E) { \
        const int blocks = sms * 16; float best = 1e30f; \
        for (int rep = 0; rep < 4; rep++) { CK(cudaEventRecord(e0)); k_extra<E><<<blocks, 128, 48 * 1024>>>(sink, m, tw, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); \
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep && ms < best) best = ms; } \
        CK(cudaGetLastError()); \
        const double bfs = (double)blocks * 128 * iters * 32.0; \
        printf("extra ALU ops per butterfly %2d: %6.2f SMSP-cycles per warp-butterfly (16 warps/SM)\n", E, best * 1e-3 * (clk_khz * 1e3) * sms * 4 / (bfs / 32)); }
    RUN_EXTRA(0) RUN_EXTRA(1) RUN_EXTRA(2) RUN_EXTRA(4) RUN_EXTRA(6) RUN_EXTRA(8) RUN_EXTRA(12)
    return 0;
}
