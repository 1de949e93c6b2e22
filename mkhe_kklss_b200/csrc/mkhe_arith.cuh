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

// 16-byte asynchronous global -> shared copy (LDGSTS), 32-byte global store (STG.256, one full sector per lane)
#ifdef MKHE_EMU
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) { *reinterpret_cast<ulonglong2 *>(smem) = *reinterpret_cast<const ulonglong2 *>(gmem); }
__device__ __forceinline__ void cp_async_wait_all() {}
__device__ __forceinline__ void prefetch_l2(const void *) {}
// TMA bulk copies complete at issue under emulation; the mbarrier keeps its transaction count and phase, so a waiter
// really waits for the thread that issues the copy
__device__ __forceinline__ void mbar_init(u64 *bar, int count) { emu_mbar_init(bar, count); }
__device__ __forceinline__ void mbar_arrive(u64 *bar) { emu_mbar_arrive(bar); }
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void fence_proxy_async_smem() {}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) { emu_mbar_expect_tx(bar, bytes); }
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) { emu_mbar_wait(bar, parity); }
__device__ __forceinline__ void tma_load_1d(void *smem, const void *gmem, u32 bytes, u64 *bar) { memcpy(smem, gmem, bytes); emu_mbar_complete_tx(bar, bytes); }
__device__ __forceinline__ void named_sync(int id, int nthreads) { emu_barrier(id, nthreads); }
__device__ __forceinline__ u64 l2_policy_evict_first() { return 0; }
__device__ __forceinline__ u64 l2_policy_evict_last() { return 0; }
__device__ __forceinline__ void tma_load_1d_hint(void *smem, const void *gmem, u32 bytes, u64 *bar, u64) { memcpy(smem, gmem, bytes); emu_mbar_complete_tx(bar, bytes); }
__device__ __forceinline__ void st_global_v4(u64 *p, u64 a, u64 b, u64 c, u64 d) { p[0] = a; p[1] = b; p[2] = c; p[3] = d; }
#else
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// mbarrier + TMA (1-D bulk copy global -> shared, UBLKCP): one thread arms the barrier with the byte count and issues the
// copies, the consumers wait on the barrier's phase parity.
__device__ __forceinline__ void mbar_init(u64 *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
// consumer -> producer: one arrival on an "empty" barrier.  When the buffer is refilled by the TMA engine (async proxy) the
// arrival must be preceded by fence_proxy_async() (MKHE_PRE_RELEASE in mkhe_kernels.cuh): the .release of the arrive orders
// generic-proxy accesses among themselves, not against an async-proxy write (measured, see there).
__device__ __forceinline__ void mbar_arrive(u64 *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// orders this thread's earlier generic-proxy accesses (ld / st) before later async-proxy (TMA) accesses of the same bytes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// the same, restricted to this CTA's shared memory (does not wait for the thread's global stores)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity) : "memory");
}
// barrier among `nthreads` threads of the CTA (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void tma_load_1d(void *smem, const void *gmem, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
// L2 eviction priorities for the bulk copies: a stream that is read exactly once (keys, hoisted forms) should not push out what
// other CTAs are about to read again (an operand common to several product groups, the twiddle tables)
__device__ __forceinline__ u64 l2_policy_evict_first() { u64 p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ u64 l2_policy_evict_last() { u64 p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ void tma_load_1d_hint(void *smem, const void *gmem, u32 bytes, u64 *bar, u64 policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void st_global_v4(u64 *p, u64 a, u64 b, u64 c, u64 d) {
    asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
#endif

// loads of data that an EARLIER KERNEL of the same stream wrote: through L2 (ld.global.cg), never from this SM's L1
#ifdef MKHE_EMU
template <class T> __device__ __forceinline__ T ld_cg(const T *p) { return *p; }
#else
__device__ __forceinline__ u64 ld_cg(const u64 *p) { return __ldcg(p); }
__device__ __forceinline__ ulonglong2 ld_cg(const ulonglong2 *p) { return __ldcg(p); }
#endif

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
    u64 nq;                   // -q mod 2^64
    u32 big;                  // q >= 2^57: forward butterflies keep values in [0,8q) with a conditional subtraction
    u32 bshift;               // bitlen(q) - 2                      } canon(): 32-bit Barrett quotient,
    u32 mu32;                 // floor(2^(bitlen(q)+30) / q)        } valid for any 64-bit input (q >= 2^36)
    u32 pad;
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


// acc += a*b on a 128-bit accumulator; the compiler's __int128 lowering chains IMAD.WIDE with carries (about 12 instructions
// per product, against 19 for mul.lo + mul.hi + a compare-and-select carry)
typedef unsigned __int128 u128;
__device__ __forceinline__ void mac128w(u128 &acc, u64 a, u64 b) { acc += (u128)a * b; }
__device__ __forceinline__ u64 hi64(u128 x) { return (u64)(x >> 64); }
__device__ __forceinline__ u64 lo64(u128 x) { return (u64)x; }

// ------------------------------------------------------------------------------------------------
// 32-bit building blocks.  Measured on B200 (tools/ubench, profiles/r01b_ubench_raw.txt): IMAD.WIDE.U32 and
// IMAD.HI.U32 occupy the FMA pipe for 4 cycles per warp, a 32-bit IMAD for 2, IADD3/LOP3/SEL run on the ALU pipe
// (2 cycles) concurrently.  The butterfly below needs 5 wide + 4 narrow multiply-adds = 28 pipe cycles; the
// compiler's generic 64-bit code (mul.hi.u64 + 2 mul.lo.u64 + a 64-bit compare/select) needs 44.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 lo32(u64 x) { return (u32)x; }
__device__ __forceinline__ u32 hi32(u64 x) { return (u32)(x >> 32); }
__device__ __forceinline__ u64 mk64(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }
#ifdef MKHE_EMU
__device__ __forceinline__ u64 mulwide(u32 a, u32 b) { return (u64)a * b; }
__device__ __forceinline__ u64 madwide(u32 a, u32 b, u64 c) { return (u64)a * b + c; }
__device__ __forceinline__ u32 madlo(u32 a, u32 b, u32 c) { return a * b + c; }
__device__ __forceinline__ u32 mulhi32(u32 a, u32 b) { return (u32)(((u64)a * b) >> 32); }
#else
__device__ __forceinline__ u64 mulwide(u32 a, u32 b) { u64 r; asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ u64 madwide(u32 a, u32 b, u64 c) { u64 r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }
__device__ __forceinline__ u32 madlo(u32 a, u32 b, u32 c) { u32 r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ u32 mulhi32(u32 a, u32 b) { u32 r; asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
#endif

// x*w mod q for ANY 64-bit x, result in [0,4q):  the Shoup quotient floor(x*wsh / 2^64) is computed without the
// low x low partial product and without carries between the cross terms, so it is short by at most 2;
// the remainder x*w - Q*q is evaluated mod 2^64 as x*w + Q*nq with nq = -q (2 wide + 4 narrow multiply-adds).
__device__ __forceinline__ u64 shoup4(u64 x, u64 w, u64 wsh, u64 nq) {
    const u32 x0 = lo32(x), x1 = hi32(x), s0 = lo32(wsh), s1 = hi32(wsh);
    const u64 a = mulwide(x1, s0);
    const u64 b = mulwide(x0, s1);
    const u64 h = madwide(x1, s1, (u64)hi32(a)) + hi32(b);
    u64 t = mulwide(x0, lo32(w));
    t = madwide(lo32(h), lo32(nq), t);
    u32 th = hi32(t);
    th = madlo(x0, hi32(w), th);
    th = madlo(x1, lo32(w), th);
    th = madlo(lo32(h), hi32(nq), th);
    th = madlo(hi32(h), lo32(nq), th);
    return mk64(lo32(t), th);
}

// constants of one modulus as the NTT butterflies use them
struct NttC {
    u64 q, nq, fourq;
    u64 zero;     // 0 at run time, opaque to the compiler: a third addend turns `x + t` into IADD3 / IADD3.X with two carries,
                  // which ptxas cannot move to the multiplier pipe as IMAD.X (the pipe that bounds the butterflies)
};
__device__ __forceinline__ NttC nttc(const ModC &m) { NttC c; c.q = m.q; c.nq = m.nq; c.fourq = 4 * m.q; c.zero = m.pad; return c; }

// forward (Cooley-Tukey) butterfly without any reduction: the Shoup product lands in [0,4q) for ANY 64-bit Y, so a stage
// adds at most 4q to the magnitude of both outputs.
//   q < 2^57        : inputs below 2^60 stay below 2^60 + 64q < 2^64 through all (<= 16) stages.
//   2^57 <= q < 2^60: the callers keep values below 16q <= 2^64 with fwd_sweep() before every second stage.
__device__ __forceinline__ void bf_fwd(u64 &X, u64 &Y, u64 w, u64 wsh, const NttC &c) {
    const u64 x = X;
    const u64 t = shoup4(Y, w, wsh, c.nq);
    X = x + t + c.zero;
    Y = x - t + c.fourq;
}
// [0,16q) -> [0,8q)
__device__ __forceinline__ u64 fwd_sweep(u64 x, const NttC &c) {
    const u64 e = 2 * c.fourq;
    return x - (x >= e ? e : 0ull) + c.zero;
}
// inverse (Gentleman-Sande) butterfly: X,Y in [0,4q) -> [0,4q)   (8q < 2^63 for every supported q)
__device__ __forceinline__ void bf_inv(u64 &X, u64 &Y, u64 w, u64 wsh, const NttC &c) {
    const u64 s = X + Y + c.zero;
    const u64 d = X - Y + c.fourq;
    X = s - (s >= c.fourq ? c.fourq : 0ull) + c.zero;
    Y = shoup4(d, w, wsh, c.nq);
}
// any 64-bit v -> canonical [0,q): 32-bit Barrett quotient.  With b = bitlen(q) >= 40: vh = floor(v / 2^(b-2)) < 2^26,
// mu32 = floor(2^(b+30)/q); the estimate floor(vh*mu32 / 2^32) is never above floor(v/q) and short by at most 1
// (error terms vh/2^32 + mu32/2^32 < 2^-6 + 1/2), so one conditional subtraction finishes the job.
__device__ __forceinline__ u64 canon(u64 v, const ModC &m) {
    const u32 vh = (u32)(v >> m.bshift);
    const u32 qa = mulhi32(vh, m.mu32);
    u64 p = mulwide(qa, lo32(m.q));
    p = mk64(lo32(p), madlo(qa, hi32(m.q), hi32(p)));
    const u64 r = v - p;
    const u64 t = r - m.q;
    return (long long)t < 0 ? r : t;
}
