// cuda_emu.h -- DEVELOPMENT TOOL, NOT A PRODUCT PATH.
//
// There is no GPU in the build container, and a gpurun round trip costs minutes.  This shim lets the
// *unmodified* kernel sources under mkhe_kklss_b200/csrc/ be compiled with g++ (-DMKHE_EMU) and
// executed block by block on the CPU, so index arithmetic and synchronisation structure can be
// debugged before any GPU time is spent.  Each CUDA thread of a block runs as a ucontext coroutine;
// __syncthreads() yields to a round-robin scheduler; blocks are spread over OpenMP threads.
//
// The emulated library (tools/emu/_build/libmkhe_emu.so) is loaded ONLY by tests/test_emu_kernels.py
// through its own loader.  mkhe_kklss_b200/_lib.py never looks for it: the product fails loudly when
// the CUDA library or a GPU is missing.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <ucontext.h>
#include <vector>
#include <functional>

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu { unsigned x, y, z; };
struct ulonglong2 { unsigned long long x, y; };
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { ulonglong2 r; r.x = x; r.y = y; return r; }

extern thread_local uint3_emu threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;
extern thread_local unsigned char *emu_smem;           // dynamic shared memory of the running block
void emu_syncthreads();
void emu_syncwarp();
void emu_barrier(int id, int nthreads);
void emu_mbar_init(unsigned long long *bar, int count);
void emu_mbar_arrive(unsigned long long *bar);
void emu_mbar_expect_tx(unsigned long long *bar, unsigned bytes);
void emu_mbar_complete_tx(unsigned long long *bar, unsigned bytes);
void emu_mbar_wait(unsigned long long *bar, unsigned parity);

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __constant__
#define __syncthreads() emu_syncthreads()
#define EMU_SHARED_DECL(type, name) type *name = reinterpret_cast<type *>(emu_smem)

static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
static inline double __ull2double_rn(unsigned long long x) { return (double)x; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline unsigned long long __double2ull_rz(double x) { return (unsigned long long)x; }
static inline long long __double_as_longlong(double x) { long long r; memcpy(&r, &x, 8); return r; }
static inline double __longlong_as_double(long long x) { double r; memcpy(&r, &x, 8); return r; }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
static inline int atomicAdd(int *p, int v) { const int o = *p; *p = o + v; return o; }      // fibers are cooperative

// ---- runtime API subset -------------------------------------------------------------------------
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef struct emu_event { double t; } *cudaEvent_t;
typedef void *cudaGraphExec_t;      // the graph cache of mkhe_api.cu is compiled out under emulation; only the type of its entries is needed
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaStreamNonBlocking = 1 };
static inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
enum { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) { *v = 6; return 0; }    // a small "device": persistent grids stay cheap to emulate
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeAsync(void *p, cudaStream_t) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < h; r++) memmove((char *)d + r * dp, (const char *)s + r * sp, w);
    return 0;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
double emu_now_ms();
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event{0}; return 0; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new emu_event{0}; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = emu_now_ms(); return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return 0; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }

void emu_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
#define MKHE_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu_launch(dim3(grid), dim3(block), (smem), [&]() { kernel(__VA_ARGS__); })
