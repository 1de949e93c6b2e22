// bfly_ubench.cu -- integer-pipe microbenchmarks for sm_100a (development tool, not part of the library).
//   (1) raw issue rates of the instructions a 64-bit modular butterfly is made of
//       (IMAD.WIDE.U32, IMAD lo, IMAD.HI, IADD3, mixes), and
//   (2) complete Harvey/Shoup forward butterflies in several formulations, each checked against a
//       128-bit reference, so that the NTT kernels use the formulation that is fastest on the B200.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o bfly_ubench bfly_ubench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// ---------------------------------------------------------------- raw instruction loops
// 8 independent chains per thread, ITER x 8 x UNR instructions
#define RAW_KERNEL(name, BODY)                                                            \
    __global__ void __launch_bounds__(256) name(u64 *sink, u32 a0, u32 b0, int iters, u64 sentinel) {   \
        u32 a = a0 + threadIdx.x, b = b0 | 1;                                             \
        u64 c[8];                                                                         \
        u32 d[8], e[8];                                                                   \
        for (int k = 0; k < 8; k++) { c[k] = threadIdx.x * 8 + k; d[k] = threadIdx.x + k; e[k] = d[k] * 3 + 1; } \
        for (int it = 0; it < iters; it++) {                                              \
            _Pragma("unroll") for (int r = 0; r < 4; r++) {                               \
                _Pragma("unroll") for (int k = 0; k < 8; k++) { BODY }                    \
            }                                                                             \
        }                                                                                 \
        u64 s = 0;                                                                        \
        for (int k = 0; k < 8; k++) s ^= c[k] ^ d[k] ^ e[k];                                     \
        if (s == sentinel) sink[0] = s;                                                   \
    }

RAW_KERNEL(raw_wide, asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[k]) : "r"((u32)c[k]), "r"(b));)
RAW_KERNEL(raw_lo, asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(d[k]) : "r"(d[k]), "r"(b));)
RAW_KERNEL(raw_hi, asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(d[k]) : "r"(d[k]), "r"(b));)
RAW_KERNEL(raw_add, asm volatile("add.u32 %0, %0, %1;" : "+r"(d[k]) : "r"(d[k] ^ b));)
RAW_KERNEL(raw_add64, asm volatile("add.u64 %0, %0, %1;" : "+l"(c[k]) : "l"(c[k] ^ a));)
RAW_KERNEL(raw_lop, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(d[k]) : "r"(d[k]), "r"(d[k] + 1));)
RAW_KERNEL(raw_wide_add, asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[k]) : "r"((u32)c[k]), "r"(b)); asm volatile("add.u32 %0, %0, %1;" : "+r"(d[k]) : "r"(d[k] ^ b));)
RAW_KERNEL(raw_wide_2add, asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[k]) : "r"((u32)c[k]), "r"(b)); asm volatile("add.u32 %0, %0, %1;" : "+r"(d[k]) : "r"(d[k] ^ b)); asm volatile("add.u32 %0, %0, %1;" : "+r"(e[k]) : "r"(e[k] ^ a));)
RAW_KERNEL(raw_lo_add, asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(d[k]) : "r"(d[k]), "r"(b)); asm volatile("add.u32 %0, %0, %1;" : "+r"(e[k]) : "r"(e[k] ^ b));)
RAW_KERNEL(raw_lo_2add, asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(d[k]) : "r"(d[k]), "r"(b)); asm volatile("add.u32 %0, %0, %1;" : "+r"(e[k]) : "r"(e[k] ^ b)); asm volatile("add.u32 %0, %0, %1;" : "+r"(e[k]) : "r"(e[k] ^ a));)
RAW_KERNEL(raw_wide_lo, asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[k]) : "r"((u32)c[k]), "r"(b)); asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(d[k]) : "r"(d[k]), "r"(b));)
RAW_KERNEL(raw_wide_2lo, asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[k]) : "r"((u32)c[k]), "r"(b)); asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(d[k]) : "r"(d[k]), "r"(b)); asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(e[k]) : "r"(e[k]), "r"(a));)
RAW_KERNEL(raw_mul64lo, c[k] = c[k] * (c[k] | 1) + 1;)
RAW_KERNEL(raw_mul64hi, c[k] = __umul64hi(c[k], c[k] | 0x8000000000000000ull) + k + 1;)

// ---------------------------------------------------------------- butterflies
struct Mod { u64 q, w, wsh; };

// V0: the library's current C formulation, values in [0,4q)
struct BfV0 {
    static constexpr bool INV = false;
    static constexpr int LAZY = 4;
    __device__ __forceinline__ static void bf(u64 &X, u64 &Y, u64 w, u64 wsh, u64 q) {
        const u64 twoq = 2 * q;
        u64 x = X >= twoq ? X - twoq : X;
        u64 h = __umul64hi(Y, wsh);
        u64 t = Y * w - h * q;
        X = x + t;
        Y = x - t + twoq;
    }
};

__device__ __forceinline__ u32 lo32(u64 x) { return (u32)x; }
__device__ __forceinline__ u32 hi32(u64 x) { return (u32)(x >> 32); }
__device__ __forceinline__ u64 mk64(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }
__device__ __forceinline__ u64 madwide(u32 a, u32 b, u64 c) { u64 r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mulwide(u32 a, u32 b) { u64 r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ u32 madlo(u32 a, u32 b, u32 c) { u32 r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ u32 mulhi32(u32 a, u32 b) { u32 r; asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

// T = Y*w - Qh*q  (mod 2^64) with nq = -q mod 2^64: 2 wide + 4 lo multiply-adds, no separate additions
__device__ __forceinline__ u64 shoup_tail(u64 Y, u64 Q, u64 w, u64 nq) {
    u64 t = mulwide(lo32(Y), lo32(w));
    t = madwide(lo32(Q), lo32(nq), t);
    u32 th = hi32(t);
    th = madlo(lo32(Y), hi32(w), th);
    th = madlo(hi32(Y), lo32(w), th);
    th = madlo(lo32(Q), hi32(nq), th);
    th = madlo(hi32(Q), lo32(nq), th);
    return mk64(lo32(t), th);
}

// V1: exact quotient (compiler's mulhi), fused tail; values in [0,4q)
struct BfV1 {
    static constexpr bool INV = false;
    static constexpr int LAZY = 4;
    __device__ __forceinline__ static void bf(u64 &X, u64 &Y, u64 w, u64 wsh, u64 q) {
        const u64 twoq = 2 * q;
        u64 x = X >= twoq ? X - twoq : X;
        u64 h = __umul64hi(Y, wsh);
        u64 t = shoup_tail(Y, h, w, 0 - q);
        X = x + t;
        Y = x - t + twoq;
    }
};

// V2: truncated quotient Q' = y1*s1 + hi32(y1*s0) + hi32(y0*s1) in [Q-2, Q]  =>  t in [0,4q);
//     values kept in [0,8q) (q < 2^60), conditional subtraction of 4q
struct BfV2 {
    static constexpr bool INV = false;
    static constexpr int LAZY = 8;
    __device__ __forceinline__ static void bf(u64 &X, u64 &Y, u64 w, u64 wsh, u64 q) {
        const u64 fourq = 4 * q;
        u64 x = X >= fourq ? X - fourq : X;
        u32 y0 = lo32(Y), y1 = hi32(Y), s0 = lo32(wsh), s1 = hi32(wsh);
        u64 a = mulwide(y1, s0);
        u64 b = mulwide(y0, s1);
        u64 h = madwide(y1, s1, (u64)hi32(a)) + hi32(b);
        u64 t = shoup_tail(Y, h, w, 0 - q);
        X = x + t;
        Y = x - t + fourq;
    }
};
// V2h: same with mul.hi for the cross terms
struct BfV2h {
    static constexpr bool INV = false;
    static constexpr int LAZY = 8;
    __device__ __forceinline__ static void bf(u64 &X, u64 &Y, u64 w, u64 wsh, u64 q) {
        const u64 fourq = 4 * q;
        u64 x = X >= fourq ? X - fourq : X;
        u32 y0 = lo32(Y), y1 = hi32(Y), s0 = lo32(wsh), s1 = hi32(wsh);
        u32 a1 = mulhi32(y1, s0);
        u32 b1 = mulhi32(y0, s1);
        u64 h = madwide(y1, s1, (u64)a1) + b1;
        u64 t = shoup_tail(Y, h, w, 0 - q);
        X = x + t;
        Y = x - t + fourq;
    }
};
// V3: V2 without any conditional subtraction (valid while (8 + 4*stages) q < 2^64, i.e. q < 2^57 for 15 stages);
//     here the benchmark loop reduces once per 12 butterflies to stay in range
struct BfV3 {
    static constexpr bool INV = false;
    static constexpr int LAZY = 0;
    __device__ __forceinline__ static void bf(u64 &X, u64 &Y, u64 w, u64 wsh, u64 q) {
        const u64 fourq = 4 * q;
        u32 y0 = lo32(Y), y1 = hi32(Y), s0 = lo32(wsh), s1 = hi32(wsh);
        u64 a = mulwide(y1, s0);
        u64 b = mulwide(y0, s1);
        u64 h = madwide(y1, s1, (u64)hi32(a)) + hi32(b);
        u64 t = shoup_tail(Y, h, w, 0 - q);
        u64 x = X;
        X = x + t;
        Y = x - t + fourq;
    }
};
// V4: quotient from the top words only: Q'' = y1*s1 + hi32(y1*s0)  (2 wide) -- error <= 1 + y0*s1/2^64 < 2^32 ... too lossy in
//     general, but with Y pre-shifted it is the classic single-word Barrett; kept as an instruction-count probe only.
struct BfV4 {
    static constexpr bool INV = false;
    static constexpr int LAZY = -1;
    __device__ __forceinline__ static void bf(u64 &X, u64 &Y, u64 w, u64 wsh, u64 q) {
        const u64 fourq = 4 * q;
        u64 x = X >= fourq ? X - fourq : X;
        u32 y1 = hi32(Y), s0 = lo32(wsh), s1 = hi32(wsh);
        u64 a = mulwide(y1, s0);
        u64 h = madwide(y1, s1, (u64)hi32(a));
        u64 t = shoup_tail(Y, h, w, 0 - q);
        X = x + t;
        Y = x - t + fourq;
    }
};

// V5: truncated quotient via mul.hi, fused tail, NO conditional subtraction (q < 2^57: 61q < 2^63 after 15 stages)
struct BfV5 {
    static constexpr bool INV = false;
    static constexpr int LAZY = 0;
    __device__ __forceinline__ static void bf(u64 &X, u64 &Y, u64 w, u64 wsh, u64 q) {
        const u64 fourq = 4 * q;
        u32 y0 = lo32(Y), y1 = hi32(Y), s0 = lo32(wsh), s1 = hi32(wsh);
        u32 a1 = mulhi32(y1, s0);
        u32 b1 = mulhi32(y0, s1);
        u64 h = madwide(y1, s1, (u64)a1) + b1;
        u64 t = shoup_tail(Y, h, w, 0 - q);
        u64 x = X;
        X = x + t;
        Y = x - t + fourq;
    }
};
// V6: V5 + top-bit conditional subtraction of 8q: values in [0, 2^63 + 4q), any q < 2^60
struct BfV6 {
    static constexpr bool INV = false;
    static constexpr int LAZY = 8;
    __device__ __forceinline__ static void bf(u64 &X, u64 &Y, u64 w, u64 wsh, u64 q) {
        const u64 fourq = 4 * q, eightq = 8 * q;
        u64 x = X;
        if ((long long)x < 0) x -= eightq;
        u32 y0 = lo32(Y), y1 = hi32(Y), s0 = lo32(wsh), s1 = hi32(wsh);
        u32 a1 = mulhi32(y1, s0);
        u32 b1 = mulhi32(y0, s1);
        u64 h = madwide(y1, s1, (u64)a1) + b1;
        u64 t = shoup_tail(Y, h, w, 0 - q);
        X = x + t;
        Y = x - t + fourq;
    }
};
// V7: as V6 but the conditional subtraction compares the high word only (threshold 4q rounded up to 2^32)
struct BfV7 {
    static constexpr bool INV = false;
    static constexpr int LAZY = 8;
    __device__ __forceinline__ static void bf(u64 &X, u64 &Y, u64 w, u64 wsh, u64 q) {
        const u64 fourq = 4 * q, fiveq = 5 * q;
        u64 x = X;
        if (hi32(x) > hi32(fourq)) x -= fourq;
        u32 y0 = lo32(Y), y1 = hi32(Y), s0 = lo32(wsh), s1 = hi32(wsh);
        u32 a1 = mulhi32(y1, s0);
        u32 b1 = mulhi32(y0, s1);
        u64 h = madwide(y1, s1, (u64)a1) + b1;
        u64 t = shoup_tail(Y, h, w, 0 - q);
        X = x + t;
        Y = x - t + fiveq;
    }
};
// I0: the library's inverse (Gentleman-Sande) butterfly;  I1: fused tail + truncated quotient + high-word compare
struct BfI0 {
    static constexpr bool INV = true;
    static constexpr int LAZY = 2;
    __device__ __forceinline__ static void bf(u64 &X, u64 &Y, u64 w, u64 wsh, u64 q) {
        const u64 twoq = 2 * q;
        u64 s = X + Y;
        u64 t = X - Y + twoq;
        X = s >= twoq ? s - twoq : s;
        u64 h = __umul64hi(t, wsh);
        Y = t * w - h * q;
    }
};
struct BfI1 {
    static constexpr bool INV = true;
    static constexpr int LAZY = 4;
    __device__ __forceinline__ static void bf(u64 &X, u64 &Y, u64 w, u64 wsh, u64 q) {
        const u64 fourq = 4 * q, fiveq = 5 * q;
        u64 s = X + Y;
        u64 d = X - Y + fiveq;
        if (hi32(s) > hi32(fourq)) s -= fourq;
        X = s;
        u32 y0 = lo32(d), y1 = hi32(d), s0 = lo32(wsh), s1 = hi32(wsh);
        u32 a1 = mulhi32(y1, s0);
        u32 b1 = mulhi32(y0, s1);
        u64 h = madwide(y1, s1, (u64)a1) + b1;
        Y = shoup_tail(d, h, w, 0 - q);
    }
};

// V8: everything on explicit 32-bit halves in PTX (carry chains spelled out) -- top-bit conditional subtraction of 8q
__device__ __forceinline__ void bf_ptx_fwd(u64 &X, u64 &Y, u64 w, u64 wsh, u64 nq, u64 fourq, u64 eightq, bool reduce) {
    u32 x0 = lo32(X), x1 = hi32(X), y0 = lo32(Y), y1 = hi32(Y);
    u32 X0, X1, Y0, Y1;
    if (reduce) {
        asm("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %1, 0;\n\t@p sub.cc.u32 %0, %0, %2;\n\t@p subc.u32 %1, %1, %3;\n\t}"
            : "+r"(x0), "+r"(x1) : "r"(lo32(eightq)), "r"(hi32(eightq)));
    }
    asm("{\n\t.reg .u32 a1, b1, c1, t0, t1, h0, h1;\n\t.reg .u64 h, t, c;\n\t"
        "mul.hi.u32 a1, %5, %8;\n\t"          // hi32(y1*s0)
        "mul.hi.u32 b1, %4, %9;\n\t"          // hi32(y0*s1)
        "add.cc.u32 a1, a1, b1;\n\t"
        "addc.u32 c1, 0, 0;\n\t"
        "mov.b64 c, {a1, c1};\n\t"
        "mad.wide.u32 h, %5, %9, c;\n\t"      // y1*s1 + ...
        "mov.b64 {h0, h1}, h;\n\t"
        "mul.wide.u32 t, %4, %6;\n\t"         // y0*w0
        "mad.wide.u32 t, h0, %10, t;\n\t"     // + h0*nq0
        "mov.b64 {t0, t1}, t;\n\t"
        "mad.lo.u32 t1, %4, %7, t1;\n\t"      // y0*w1
        "mad.lo.u32 t1, %5, %6, t1;\n\t"      // y1*w0
        "mad.lo.u32 t1, h0, %11, t1;\n\t"     // h0*nq1
        "mad.lo.u32 t1, h1, %10, t1;\n\t"     // h1*nq0
        "add.cc.u32 %0, %12, t0;\n\t"
        "addc.u32 %1, %13, t1;\n\t"
        "sub.cc.u32 %2, %12, t0;\n\t"
        "subc.u32 %3, %13, t1;\n\t"
        "add.cc.u32 %2, %2, %14;\n\t"
        "addc.u32 %3, %3, %15;\n\t}"
        : "=&r"(X0), "=&r"(X1), "=&r"(Y0), "=&r"(Y1)
        : "r"(y0), "r"(y1), "r"(lo32(w)), "r"(hi32(w)), "r"(lo32(wsh)), "r"(hi32(wsh)), "r"(lo32(nq)), "r"(hi32(nq)),
          "r"(x0), "r"(x1), "r"(lo32(fourq)), "r"(hi32(fourq)));
    X = mk64(X0, X1);
    Y = mk64(Y0, Y1);
}
struct BfV8 {
    static constexpr bool INV = false;
    static constexpr int LAZY = 8;
    __device__ __forceinline__ static void bf(u64 &X, u64 &Y, u64 w, u64 wsh, u64 q) { bf_ptx_fwd(X, Y, w, wsh, 0 - q, 4 * q, 8 * q, true); }
};
struct BfV9 {   // no conditional subtraction
    static constexpr bool INV = false;
    static constexpr int LAZY = 0;
    __device__ __forceinline__ static void bf(u64 &X, u64 &Y, u64 w, u64 wsh, u64 q) { bf_ptx_fwd(X, Y, w, wsh, 0 - q, 4 * q, 8 * q, false); }
};

template <class BF, bool REDUCE>
__global__ void __launch_bounds__(256) k_bf(u64 *sink, Mod m, int iters, u64 *verify) {
    u64 v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = ((u64)(threadIdx.x * 8 + k + blockIdx.x) * 0x9E3779B97F4A7C15ull) % m.q;
    const u64 w = m.w, wsh = m.wsh, q = m.q;
    const u64 mu = (~0ull) / q;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 4; k++) BF::bf(v[k], v[k + 4], w, wsh, q);
#pragma unroll
        for (int g = 0; g < 2; g++) { BF::bf(v[4 * g], v[4 * g + 2], w, wsh, q); BF::bf(v[4 * g + 1], v[4 * g + 3], w, wsh, q); }
#pragma unroll
        for (int g = 0; g < 4; g++) BF::bf(v[2 * g], v[2 * g + 1], w, wsh, q);
        if (BF::LAZY == 0 && REDUCE) {
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = v[k] - __umul64hi(v[k], mu) * q;
        }
    }
    if (verify && blockIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 8; k++) verify[threadIdx.x * 8 + k] = v[k] % q;
    }
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= v[k];
    if (s == 0x123456789abcdefull) sink[0] = s;
}

// exact reference of the same loop
__global__ void k_ref(Mod m, int iters, u64 *verify, bool inv) {
    u64 v[8];
    for (int k = 0; k < 8; k++) v[k] = ((u64)(threadIdx.x * 8 + k + blockIdx.x) * 0x9E3779B97F4A7C15ull) % m.q;
    auto bf = [&](u64 &X, u64 &Y) {
        if (inv) {
            u64 s = (X + Y) % m.q, d = (X + m.q - Y) % m.q;
            X = s;
            Y = (u64)(((unsigned __int128)d * m.w) % m.q);
            return;
        }
        u64 t = (u64)(((unsigned __int128)Y * m.w) % m.q);
        u64 x = X;
        X = (x + t) % m.q;
        Y = (x + m.q - t) % m.q;
    };
    for (int it = 0; it < iters; it++) {
        for (int k = 0; k < 4; k++) bf(v[k], v[k + 4]);
        for (int g = 0; g < 2; g++) { bf(v[4 * g], v[4 * g + 2]); bf(v[4 * g + 1], v[4 * g + 3]); }
        for (int g = 0; g < 4; g++) bf(v[2 * g], v[2 * g + 1]);
    }
    for (int k = 0; k < 8; k++) verify[threadIdx.x * 8 + k] = v[k];
}

template <class F>
static double time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, max clock %.0f MHz\n", prop.name, sms, clk_khz / 1e3);
    u64 *sink, *ver0, *ver1;
    CK(cudaMalloc(&sink, 64)); CK(cudaMalloc(&ver0, 2048 * 8)); CK(cudaMalloc(&ver1, 2048 * 8));
    const int blocks = sms * 8, iters = 4096;

#define RUN_RAW(name, ninst)                                                                                  \
    {                                                                                                         \
        double ms = time_ms([&] { name<<<blocks, 256>>>(sink, 12345u, 6789u, iters, 0x123456789abcdefull); }, 5);                     \
        double winst = (double)blocks * 8 /*warps*/ * iters * 32.0 * (ninst);                                  \
        double per_smsp_cyc = ms * 1e-3 * (clk_khz * 1e3) * sms * 4 / winst;                                   \
        printf("%-16s %8.3f ms  %6.2f SMSP-cycles per warp-instruction-group (%d instr)  [at max clock]\n", #name, ms, per_smsp_cyc, ninst); \
    }
    RUN_RAW(raw_wide, 1);
    RUN_RAW(raw_lo, 1);
    RUN_RAW(raw_hi, 1);
    RUN_RAW(raw_add, 1);
    RUN_RAW(raw_add64, 1);
    RUN_RAW(raw_lop, 1);
    RUN_RAW(raw_wide_add, 2);
    RUN_RAW(raw_wide_2add, 3);
    RUN_RAW(raw_lo_add, 2);
    RUN_RAW(raw_lo_2add, 3);
    RUN_RAW(raw_wide_lo, 2);
    RUN_RAW(raw_wide_2lo, 3);
    RUN_RAW(raw_mul64lo, 1);
    RUN_RAW(raw_mul64hi, 1);

    const u64 qs[3] = {0xfffffffff6a0001ull, 0x3fffffffd60001ull, 0x800000020001ull};
    for (int qi = 0; qi < 3; qi++) {
        Mod m; m.q = qs[qi];
        m.w = 0x123456789abcdef1ull % m.q;
        m.wsh = (u64)((((unsigned __int128)m.w) << 64) / m.q);
        const int vit = 37;
        static u64 h0[2][2048], h1[2048];
        for (int inv = 0; inv < 2; inv++) {
            k_ref<<<1, 256>>>(m, vit, ver0, inv != 0);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h0[inv], ver0, sizeof(h0[inv]), cudaMemcpyDeviceToHost));
        }
#define RUN_BF(BF)                                                                                            \
        {                                                                                                     \
            k_bf<BF, true><<<1, 256>>>(sink, m, vit, ver1);                                                         \
            CK(cudaDeviceSynchronize());                                                                      \
            CK(cudaMemcpy(h1, ver1, sizeof(h1), cudaMemcpyDeviceToHost));                                     \
            int bad = 0;                                                                                      \
            for (int i = 0; i < 2048; i++) bad += (h0[BF::INV][i] != h1[i]);                                           \
            double ms = time_ms([&] { k_bf<BF, false><<<blocks, 256>>>(sink, m, iters, nullptr); }, 5);              \
            double bfs = (double)blocks * 256 * iters * 12.0;                                                 \
            double cyc = ms * 1e-3 * (clk_khz * 1e3) * sms * 4 / (bfs / 32);                                  \
            printf("q=%016llx %-6s %8.3f ms  %.3e butterflies/s  %6.2f SMSP-cycles per warp-butterfly  mismatches=%d%s\n", \
                   m.q, #BF, ms, bfs / (ms * 1e-3), cyc, bad, BF::LAZY < 0 ? " (probe only, not exact)" : "");  \
        }
        RUN_BF(BfV0);
        RUN_BF(BfV1);
        RUN_BF(BfV2);
        RUN_BF(BfV2h);
        RUN_BF(BfV3);
        RUN_BF(BfV4);
        RUN_BF(BfV5);
        RUN_BF(BfV6);
        RUN_BF(BfV7);
        RUN_BF(BfV8);
        RUN_BF(BfV9);
        RUN_BF(BfI0);
        RUN_BF(BfI1);
    }
    return 0;
}
