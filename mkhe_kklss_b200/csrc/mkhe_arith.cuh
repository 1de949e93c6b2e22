// mkhe_arith.cuh -- 64-bit modular arithmetic for sm_100a (no tensor cores: these are modular integer
// transforms).  B200 has no 64-bit integer multiplier; every 64x64 product below lowers to IMAD.WIDE
// chains, which is why the NTT stages are integer-pipe bound.
//
// Conventions: q < 2^60 (every prime of the reference's parameter sets), so 4q < 2^62.
//   Shoup pair (w, wsh): wsh = floor(w * 2^64 / q);  shoup_lazy(x) = x*w - floor(x*wsh/2^64)*q in [0,2q) for ANY x < 2^64.
//   Montgomery: R = 2^64, qinv = q^-1 mod 2^64 (lattigo MredParams);  mred(x,y) = x*y/R mod q, canonical.
#pragma once
#include <stdint.h>

#ifdef MKHE_EMU
#include "cuda_emu.h"
#define MKHE_SMEM(name) unsigned char *name = emu_smem
#else
#include <cuda_runtime.h>
#define MKHE_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define MKHE_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

typedef unsigned long long u64;
typedef unsigned int u32;

#define MKHE_MAX_PARTIES_K 66     // entries of a kernel-argument pointer list (64 parties + component "0" + 1)

#ifdef MKHE_EMU
__device__ __forceinline__ int mkhe_clz(u32 x) { return __builtin_clz(x); }
#else
__device__ __forceinline__ int mkhe_clz(u32 x) { return __clz((int)x); }
#endif

// per-modulus constants, one entry per prime (Q limbs, then P limbs, then QMul limbs)
struct ModC {
    u64 q, qinv;              // modulus, q^-1 mod 2^64
    u64 mu;                   // floor(2^64 / q): Barrett constant for 64-bit inputs
    u64 r2;                   // 2^128 mod q  (MForm(x) = mred(x, r2))
    u64 ninv, ninv_sh;        // N^-1 mod q as a Shoup pair
    u64 w1ninv, w1ninv_sh;    // psi^-(N/2)... = invtw[1] * N^-1 as a Shoup pair (last inverse stage)
    double qd;                // (double) q, for the fp64 overflow estimate v (basis_extension.go:548)
    u64 pad;
};

__device__ __forceinline__ u64 mulhi(u64 a, u64 b) { return __umul64hi(a, b); }

// x*w mod q in [0,2q); valid for any 64-bit x
__device__ __forceinline__ u64 shoup_lazy(u64 x, u64 w, u64 wsh, u64 q) {
    u64 h = mulhi(x, wsh);
    return x * w - h * q;
}
// x mod q in [0,2q) for any 64-bit x
__device__ __forceinline__ u64 barrett_lazy(u64 x, u64 q, u64 mu) { return x - mulhi(x, mu) * q; }
__device__ __forceinline__ u64 csub(u64 x, u64 q) { return x >= q ? x - q : x; }
// Montgomery reduction of (hi:lo) < q*2^64 -> canonical [0,q)
__device__ __forceinline__ u64 mont_reduce(u64 hi, u64 lo, u64 q, u64 qinv) {
    u64 t = mulhi(lo * qinv, q);
    u64 r = hi - t;
    return hi < t ? r + q : r;
}
__device__ __forceinline__ u64 mred(u64 x, u64 y, u64 q, u64 qinv) {
    return mont_reduce(mulhi(x, y), x * y, q, qinv);
}
// 128-bit multiply-accumulate (acc += a*b)
__device__ __forceinline__ void mac128(u64 &hi, u64 &lo, u64 a, u64 b) {
    u64 pl = a * b, ph = mulhi(a, b);
    lo += pl;
    hi += ph + (lo < pl ? 1ull : 0ull);
}

// Harvey forward (Cooley-Tukey) butterfly: X,Y in [0,4q) -> [0,4q)
__device__ __forceinline__ void bf_fwd(u64 &X, u64 &Y, u64 w, u64 wsh, u64 q, u64 twoq) {
    u64 x = X >= twoq ? X - twoq : X;
    u64 t = shoup_lazy(Y, w, wsh, q);
    X = x + t;
    Y = x - t + twoq;
}
// Harvey inverse (Gentleman-Sande) butterfly: X,Y in [0,2q) -> [0,2q)
__device__ __forceinline__ void bf_inv(u64 &X, u64 &Y, u64 w, u64 wsh, u64 q, u64 twoq) {
    u64 s = X + Y;
    u64 t = X - Y + twoq;
    X = s >= twoq ? s - twoq : s;
    Y = shoup_lazy(t, w, wsh, q);
}
