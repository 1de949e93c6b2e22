// mkhe_kernels.cuh -- sm_100a kernels of the MKHE-KKLSS hot path.
//
// NTT organisation (N = 2^logN, 12 <= logN <= 16).  lattigo's forward NTT is an in-place Cooley-Tukey
// DIT with in-order input and bit-reversed output; stage s (1-based) pairs positions j and
// j + N/2^s and uses the twiddle psi[2^(s-1) + (j >> (logN-s+1))] (lattigo ring/ntt.go; the tables
// are public fields read at mkrlwe/basis_extension.go:263-264).  The same dataflow is split into
//   pass 1: the first S1 = logN-11 stages.  They act on 2048 independent "columns" (positions with
//           equal j mod 2048) and every column uses the same 2^S1-1 twiddles.  A thread owns one
//           column = 2^S1 elements (stride 2048) and does the whole pass in registers: no shared
//           memory, no barrier; a warp touches 256 contiguous bytes per access.
//   pass 2: the last 11 stages = N/2048 independent contiguous tiles of 2048 elements, each with its
//           own 2047 twiddles.  A CTA of 128 threads owns a tile, 16 elements per thread, three
//           register rounds (4 + 4 + 3 stages) with two XOR-swizzled shared-memory exchanges.
// The inverse transform mirrors this (Gentleman-Sande, pass A = tiles, pass B = columns with N^-1
// folded into the last stage).  Positions are never permuted, so key material in lattigo's ordering
// is used as uploaded.  Butterfly arithmetic: mkhe_arith.cuh.
#pragma once
#include "mkhe_arith.cuh"
#include "../../include/mkhe_prng.h"

#define MKHE_TILE 2048
// Releasing a TMA landing buffer: the consumer's shared-memory reads are generic-proxy accesses, the refill is an async-proxy
// write.  mbarrier.arrive (.release) alone does NOT keep a read that is merely in flight ahead of that write on B200: measured,
// 23-33 of 6400 back-to-back MulRelin calls had one 2048-coefficient tile wrong without this fence (L2 hints off; about 1 in 3000
// with them), 0 of 6400 with it (tools/gpu_race_ab.sh, profiles/r02m_release_fence.txt).  fence.proxy.async orders the thread's
// earlier generic-proxy accesses before later async-proxy ones; it sits before every arrive on an "empty" barrier.
#ifndef MKHE_RELEASE_FENCE                 // development: A/B builds override it (0 = none, 1 = fence.proxy.async, 2 = .shared::cta)
#define MKHE_RELEASE_FENCE 2
#endif
#if MKHE_RELEASE_FENCE == 0
#define MKHE_PRE_RELEASE()
#elif MKHE_RELEASE_FENCE == 1
#define MKHE_PRE_RELEASE() fence_proxy_async()
#else
#define MKHE_PRE_RELEASE() fence_proxy_async_smem()
#endif
#define MKHE_NTT_THREADS 128     // NTT kernels: 16 elements per thread
#define MKHE_THREADS 256         // element-wise kernels
#define MKHE_MAX_RANKS 8          // ranks of a multi-GPU team
#define MKHE_MAX_SLOTS 40        // limb slots per launch list (nQ + nP <= 40: PN16QP1761 has 34 + 4)

// exchange buffer: the three register layouts of a tile (see tile_fwd) meet in one padded buffer with two
// address maps, so that every access is conflict-free and its address is a per-thread base plus a compile-time offset:
//   map 1 (layouts A <-> B): element idx at idx + 8*(idx >> 7)        (rows of 128 padded to 136)
//   map 2 (layouts B <-> C): element idx at idx + (idx >> 4)          (groups of 16 padded to 17)
// Both maps send the 512 elements [512w, 512w+512) that warp w owns in layouts B and C to the same slots
// [544w, 544w+544), so the B <-> C exchange only needs a warp-level barrier.
#define MKHE_XBUF 2176           // u64 slots of the exchange buffer (2048 + 128 padding)
#ifdef MKHE_EMU
#define MKHE_SYNCWARP() emu_syncwarp()
#else
#define MKHE_SYNCWARP() __syncwarp()
#endif

struct PtrList { u64 *p[MKHE_MAX_PARTIES_K]; };

// ------------------------------------------------------------------------------------------------
// twiddles of one pass-2 / pass-A tile in shared memory.  Stage "d" = butterfly distance 2^d in
// tile-local index space (0..10); level lv = 10-d has 2^lv twiddles, heap index 2^lv + g where
// g = local_idx >> (d+1).  Global table index = 2^(S1+lv) + (tile << lv) + g.
//   levels 0..3 (round A, warp-uniform reads): heap order.
//   levels 4..7 (round B, thread t = hi*8+low reads entries hi*2^(lv-4) + j): stored as [j][hi].
//   levels 8..10 (round C, thread t reads its private entries t*2^(lv-7) + j): stored as [j][t].
// so every read is (per-thread base) + (compile-time offset) and a warp reads consecutive 16-byte entries.
// ------------------------------------------------------------------------------------------------
// The staged image of every (modulus, tile) is built once per context (k_tile_twiddles) so that a CTA fetches its 32 KiB
// with ONE TMA bulk copy:  tiled[(mod * ntiles + tile) * 2048 + pos].
__device__ __forceinline__ int tile_twiddle_pos(int h) {
    const int lv = 31 - mkhe_clz((u32)h);
    const int g = h - (1 << lv);
    if (lv >= 8) return (1 << lv) + (g & ((1 << (lv - 7)) - 1)) * MKHE_NTT_THREADS + (g >> (lv - 7));
    if (lv >= 4) return (1 << lv) + (g & ((1 << (lv - 4)) - 1)) * 16 + (g >> (lv - 4));
    return h;
}
// grid = (ntiles, nmods), 256 threads
__global__ void __launch_bounds__(MKHE_THREADS) k_tile_twiddles(const ulonglong2 *tab, ulonglong2 *tiled, int logN) {
    const long N = 1L << logN;
    const int S1 = logN - 11, ntiles = (int)(N / MKHE_TILE), tile = blockIdx.x, mi = blockIdx.y;
    const ulonglong2 *src = tab + (long)mi * N;
    ulonglong2 *dst = tiled + ((long)mi * ntiles + tile) * MKHE_TILE;
    for (int h = threadIdx.x; h < MKHE_TILE; h += MKHE_THREADS) {
        if (h == 0) { dst[0] = make_ulonglong2(0, 0); continue; }
        const int lv = 31 - mkhe_clz((u32)h);
        dst[tile_twiddle_pos(h)] = src[((size_t)1 << (S1 + lv)) + ((size_t)tile << lv) + (h - (1 << lv))];
    }
}
struct TwShared {          // accessor used by the tile rounds; b = the register-index bit of the stage
    const ulonglong2 *sA, *sB, *sC;     // round A: 15 warp-uniform entries (broadcast reads)
    __device__ __forceinline__ TwShared(const ulonglong2 *base, int tid)
        : sA(base), sB(base + (tid >> 3)), sC(base + tid) {}
    __device__ __forceinline__ ulonglong2 A(int b, int g) const { return sA[(1 << (3 - b)) + g]; }
    __device__ __forceinline__ ulonglong2 B(int b, int g) const { return sB[(1 << (7 - b)) + g * 16]; }
    __device__ __forceinline__ ulonglong2 C(int b, int g) const { return sC[(1 << (10 - b)) + g * MKHE_NTT_THREADS]; }
};
struct TwGlobal {          // same interface straight from the global table (single-use tiles)
    const ulonglong2 *tA, *tB, *tC;
    int S1, tile;
    __device__ __forceinline__ TwGlobal(const ulonglong2 *tab, int S1_, int tile_) : tA(tab), tB(tab), tC(tab), S1(S1_), tile(tile_) {}
    __device__ __forceinline__ ulonglong2 at(int lv, size_t g) const { return __ldg(tA + ((size_t)1 << (S1 + lv)) + ((size_t)tile << lv) + g); }
    __device__ __forceinline__ ulonglong2 A(int b, int g) const { return at(3 - b, g); }
    __device__ __forceinline__ ulonglong2 B(int b, int g) const { return at(7 - b, ((size_t)(threadIdx.x >> 3) << (3 - b)) + g); }
    __device__ __forceinline__ ulonglong2 C(int b, int g) const { return at(10 - b, ((size_t)threadIdx.x << (3 - b)) + g); }
};

struct TwGlobalT {         // TwGlobal for a group that is not the first 128 threads of its CTA (explicit thread index)
    const ulonglong2 *t;
    int S1, tile, tid;
    __device__ __forceinline__ TwGlobalT(const ulonglong2 *tab, int S1_, int tile_, int tid_) : t(tab), S1(S1_), tile(tile_), tid(tid_) {}
    __device__ __forceinline__ ulonglong2 at(int lv, size_t g) const { return __ldg(t + ((size_t)1 << (S1 + lv)) + ((size_t)tile << lv) + g); }
    __device__ __forceinline__ ulonglong2 A(int b, int g) const { return at(3 - b, g); }
    __device__ __forceinline__ ulonglong2 B(int b, int g) const { return at(7 - b, ((size_t)(tid >> 3) << (3 - b)) + g); }
    __device__ __forceinline__ ulonglong2 C(int b, int g) const { return at(10 - b, ((size_t)tid << (3 - b)) + g); }
};

// ------------------------------------------------------------------------------------------------
// tile rounds.  Thread t = hi*8 + low holds 16 elements; their tile-local indices are
//   layout A:  k*128 + t                 (stages d = 10..7 are register-local, global access coalesced)
//   layout B:  (hi*16 + k)*8 + low       (stages d = 6..3)
//   layout C:  t*16 + k                  (stages d = 2..0; 16 contiguous elements per thread)
// ------------------------------------------------------------------------------------------------
struct XAddr {             // per-thread bases into the exchange buffer
    u64 *a1, *b1, *c2;
    __device__ __forceinline__ XAddr(u64 *buf, int tid) {
        const int hi = tid >> 3, low = tid & 7;
        a1 = buf + tid;                    // map 1, layout A: + k*136
        b1 = buf + hi * 136 + low;         // map 1, layout B: + k*8 ;  map 2, layout B: + k*8 + (k>>1)
        c2 = buf + tid * 17;               // map 2, layout C: + k
    }
};
#define MKHE_STAGE(BF, TWEXPR, NB)                                                                     \
    _Pragma("unroll") for (int i = 0; i < 8; i++) {                                                    \
        const int g = i >> b, j = i & ((1 << b) - 1);                                                  \
        const ulonglong2 w = TWEXPR;                                                                   \
        BF(v[(g << (b + 1)) + j], v[(g << (b + 1)) + (1 << b) + j], w.x, w.y, c);                      \
    }
#define MKHE_SWEEP16()                                                                                 \
    if (big) { _Pragma("unroll") for (int k = 0; k < 16; k++) v[k] = fwd_sweep(v[k], c); }

// the caller guarantees (block barrier) that nobody still reads the buffer when tile_fwd / tile_inv starts.
// big (2^57 <= q < 2^60, CTA-uniform): inputs below 16q, a sweep to [0,8q) before every second stage keeps them there.
// Round A works on registers only; rounds B and C start with the stores into the exchange buffer, so a caller can put
// one barrier between the two halves that covers both "everybody has consumed its input" and "everybody has left the
// previous instance's exchange buffer".
template <class TW>
__device__ __forceinline__ void tile_fwd_A(u64 v[16], const TW &tw, const NttC &c, const bool big) {
#pragma unroll
    for (int b = 3; b >= 0; b--) {
        if (b & 1) MKHE_SWEEP16()
        MKHE_STAGE(bf_fwd, tw.A(b, g), 0)
    }
}
template <class TW>
__device__ __forceinline__ void tile_fwd_BC(u64 v[16], const XAddr &x, const TW &tw, const NttC &c, const bool big, const int bar_id) {
#pragma unroll
    for (int k = 0; k < 16; k++) x.a1[k * 136] = v[k];
    named_sync(bar_id, MKHE_NTT_THREADS);
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = x.b1[k * 8];
#pragma unroll
    for (int b = 3; b >= 0; b--) {
        if (b & 1) MKHE_SWEEP16()
        MKHE_STAGE(bf_fwd, tw.B(b, g), 0)
    }
    MKHE_SYNCWARP();
#pragma unroll
    for (int k = 0; k < 16; k++) x.b1[k * 8 + (k >> 1)] = v[k];
    MKHE_SYNCWARP();
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = x.c2[k];
#pragma unroll
    for (int b = 2; b >= 0; b--) {
        if (!(b & 1)) MKHE_SWEEP16()
        MKHE_STAGE(bf_fwd, tw.C(b, g), 0)
    }
}
// inverse: layout C in, layout A out; values stay in [0,4q)
template <class TW>
__device__ __forceinline__ void tile_inv(u64 v[16], const XAddr &x, const TW &tw, const NttC &c, const int bar_id) {
#pragma unroll
    for (int b = 0; b <= 2; b++) { MKHE_STAGE(bf_inv, tw.C(b, g), 0) }
#pragma unroll
    for (int k = 0; k < 16; k++) x.c2[k] = v[k];
    MKHE_SYNCWARP();
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = x.b1[k * 8 + (k >> 1)];
#pragma unroll
    for (int b = 0; b <= 3; b++) { MKHE_STAGE(bf_inv, tw.B(b, g), 0) }
    MKHE_SYNCWARP();
#pragma unroll
    for (int k = 0; k < 16; k++) x.b1[k * 8] = v[k];
    named_sync(bar_id, MKHE_NTT_THREADS);
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = x.a1[k * 136];
#pragma unroll
    for (int b = 0; b <= 3; b++) { MKHE_STAGE(bf_inv, tw.A(b, g), 0) }
}

// ------------------------------------------------------------------------------------------------
// column stages in registers: a thread owns column `col`, element k at position k*2048 + col,
// E = 2^S1 elements.  Stage s (1..S1) pairs k and k + E/2^s with twiddle index 2^(s-1) + (k >> (S1-s+1)),
// the same for every column (warp-uniform loads).
// ------------------------------------------------------------------------------------------------
// big: inputs below 8q, outputs below 16q (sweep before the 3rd and 5th stage)
template <int S1>
__device__ __forceinline__ void cols_fwd(u64 *v, const ulonglong2 *tw, const NttC &c, const bool big) {
#pragma unroll
    for (int b = S1 - 1; b >= 0; b--) {
        if (b < S1 - 1 && ((S1 - 1 - b) & 1) == 0) {
            if (big) {
#pragma unroll
                for (int k = 0; k < (1 << S1); k++) v[k] = fwd_sweep(v[k], c);
            }
        }
#pragma unroll
        for (int i = 0; i < (1 << (S1 - 1)); i++) {
            const int g = i >> b, j = i & ((1 << b) - 1);
            const ulonglong2 w = __ldg(tw + (1 << (S1 - 1 - b)) + g);
            bf_fwd(v[(g << (b + 1)) + j], v[(g << (b + 1)) + (1 << b) + j], w.x, w.y, c);
        }
    }
}
// inverse column stages: inputs in [0,4q); the last stage (twiddle index 1) also multiplies by N^-1
// and returns canonical values.
template <int S1>
__device__ __forceinline__ void cols_inv(u64 *v, const ulonglong2 *tw, const NttC &c, const ModC &m) {
#pragma unroll
    for (int b = 0; b < S1 - 1; b++) {
#pragma unroll
        for (int i = 0; i < (1 << (S1 - 1)); i++) {
            const int g = i >> b, j = i & ((1 << b) - 1);
            const ulonglong2 w = __ldg(tw + (1 << (S1 - 1 - b)) + g);
            bf_inv(v[(g << (b + 1)) + j], v[(g << (b + 1)) + (1 << b) + j], w.x, w.y, c);
        }
    }
    constexpr int H = 1 << (S1 - 1);
#pragma unroll
    for (int j = 0; j < H; j++) {
        const u64 s = v[j] + v[j + H];                         // < 8q
        const u64 d = v[j] - v[j + H] + c.fourq;               // < 8q
        v[j] = csub(shoup_lazy(s, m.ninv, m.ninv_sh, m.q), m.q);
        v[j + H] = csub(shoup_lazy(d, m.w1ninv, m.w1ninv_sh, m.q), m.q);
    }
}

// ------------------------------------------------------------------------------------------------
// K1'  decompose + ModUp (alpha = 1: digit broadcast) + first S1 NTT stages
//   replaces Decomposer.DecomposeAndSplit copy branch + the head of ringQ/ringP.NTTLvl
//   (mkrlwe/basis_extension.go:443-451, mkrlwe/keyswitch.go:21-31).
//   grid = (2048/128 column groups * slot groups, ndigits, npolys).  The digit limb (values < 2^60) is read
//   once per slot group and fed UNREDUCED into every target limb: the first butterfly's Shoup product reduces
//   the multiplied arm, the lazy ranges of bf_fwd absorb the other one, so no Barrett step is needed.
// ------------------------------------------------------------------------------------------------
struct BcastArgs {
    PtrList in;            // per poly: coefficient-domain poly
    PtrList out;           // per poly: swk-shaped buffer [digit][Dmax][N]
    int in_limb0;          // source limb of digit d = in_limb0 + d * in_limb_stride (BFV's second half starts at nQ; alpha > 1: stride alpha)
    int in_limb_stride;
    int digit0;            // first digit of this launch
    int dmax;              // limb slots per digit in the output
    int nslots;
    int slot_groups;       // the slot list is split over blockIdx.x / 16
    int slots[MKHE_MAX_SLOTS];   // modulus index == limb slot of each target limb
    int logN;
};

#ifndef MKHE_BCAST_MINB
#define MKHE_BCAST_MINB 5
#endif
template <int S1>
__global__ void __launch_bounds__(MKHE_NTT_THREADS, (S1 <= 4 ? MKHE_BCAST_MINB : 2)) k_bcast_ntt_pass1(BcastArgs a, const ModC *mods, const ulonglong2 *twf) {
    constexpr int E = 1 << S1;
    const int col = (blockIdx.x & 15) * MKHE_NTT_THREADS + threadIdx.x, sg = blockIdx.x >> 4;
    const int digit = a.digit0 + blockIdx.y, poly = blockIdx.z;
    const long N = 1L << a.logN;
    const u64 *src = a.in.p[poly] + (long)(a.in_limb0 + digit * a.in_limb_stride) * N + col;
    u64 *dst0 = a.out.p[poly] + (long)digit * a.dmax * N + col;
    const int per = (a.nslots + a.slot_groups - 1) / a.slot_groups;
    const int s_end = (sg + 1) * per < a.nslots ? (sg + 1) * per : a.nslots;
    for (int s = sg * per; s < s_end; s++) {
        const int mi = a.slots[s];
        const ModC m = mods[mi];
        const NttC c = nttc(m);
        const ulonglong2 *tw = twf + (long)mi * N;
        // the digit column is re-read for every target limb (L1 hits after the first) instead of being held in registers:
        // the kernel is integer bound and the registers buy a fifth resident CTA
        u64 v[E];
#pragma unroll
        for (int k = 0; k < E; k++) v[k] = __ldg(src + (long)k * MKHE_TILE);
        if (m.big) {       // [0,8q) is required: digits of another limb may exceed it only for lazy 60-bit limbs; reduce to be safe
#pragma unroll
            for (int k = 0; k < E; k++) v[k] = barrett_lazy(v[k], m.q, m.mu);
        }
        cols_fwd<S1>(v, tw, c, m.big != 0);
        u64 *dst = dst0 + (long)mi * N;
#pragma unroll
        for (int k = 0; k < E; k++) dst[(long)k * MKHE_TILE] = v[k];
    }
}

// ------------------------------------------------------------------------------------------------
// K1'' Decomposer.DecomposeAndSplit for alpha > 1 (mkrlwe/basis_extension.go:454-534): the nsrc limbs of digit d are
//   lifted exactly to every Q limb <= level and every P limb: y_i = x_i (Q_d/q_i)^-1 mod q_i, v = trunc(sum fl(y_i)/fl(q_i))
//   (fp64, same operation order), out_j = multSum's lazy value.  The reference overwrites the digit's own limbs with their
//   multSum values as well (:525 starts at alpha*beta); every limb is written the same way here.  Coefficient domain in,
//   coefficient domain out (the plain forward NTT follows).  grid = (N/256, ndigits, npolys)
// ------------------------------------------------------------------------------------------------
#define MKHE_LIFT_MAX_SRC 4
struct LiftTable {            // constants of one (digit, nsrc) pair; device resident
    int nsrc, src_limb0;
    u64 qoverqiinvqi[MKHE_LIFT_MAX_SRC];                       // Montgomery form
    u64 qoverqimodp[MKHE_MAX_SLOTS][MKHE_LIFT_MAX_SRC];        // [target modulus index][i], Montgomery form
    u64 vtimesqmodp[MKHE_MAX_SLOTS][MKHE_LIFT_MAX_SRC + 1];
};
struct LiftArgs {
    PtrList in;            // per poly: coefficient-domain poly
    PtrList out;           // per poly: swk-shaped buffer
    int digit0;            // first digit of this launch
    int table0;            // index of digit0's table in the table array; tables of consecutive digits are table_stride apart
    int table_stride;
    int last_digit, last_table;      // the (possibly partial) last digit uses table `last_table`
    int dmax, nslots;
    int slots[MKHE_MAX_SLOTS];
    int logN;
};
__global__ void __launch_bounds__(MKHE_THREADS) k_decomp_lift(LiftArgs a, const LiftTable *tabs, const ModC *mods) {
    const long N = 1L << a.logN;
    const long x = (long)blockIdx.x * MKHE_THREADS + threadIdx.x;
    const int digit = a.digit0 + blockIdx.y, poly = blockIdx.z;
    const LiftTable &tab = tabs[digit == a.last_digit ? a.last_table : a.table0 + (digit - a.digit0) * a.table_stride];
    u64 y[MKHE_LIFT_MAX_SRC];
    double vi = 0.0;
#pragma unroll
    for (int i = 0; i < MKHE_LIFT_MAX_SRC; i++) {
        if (i < tab.nsrc) {
            const ModC &m = mods[tab.src_limb0 + i];
            y[i] = mred(a.in.p[poly][(long)(tab.src_limb0 + i) * N + x], tab.qoverqiinvqi[i], m.q, m.qinv);
            vi = __dadd_rn(vi, __ddiv_rn(__ull2double_rn(y[i]), m.qd));
        } else {
            y[i] = 0;
        }
    }
    const u64 v = __double2ull_rz(vi);
    u64 *dst = a.out.p[poly] + (long)digit * a.dmax * N + x;
    for (int s = 0; s < a.nslots; s++) {
        const int mi = a.slots[s];
        const ModC &m = mods[mi];
        u64 rlo = 0, rhi = 0;
#pragma unroll
        for (int i = 0; i < MKHE_LIFT_MAX_SRC; i++)
            if (i < tab.nsrc) mac128(rhi, rlo, y[i], tab.qoverqimodp[mi][i]);
        const u64 hhi = mulhi(rlo * m.qinv, m.q);
        dst[(long)mi * N] = rhi - hhi + m.q + tab.vtimesqmodp[mi][v];
    }
}

// ------------------------------------------------------------------------------------------------
// generic limb lists for the plain transforms: one launch handles `nlimbs` limbs of `npolys` polys
// ------------------------------------------------------------------------------------------------
struct LimbArgs {
    PtrList in, out;       // per poly base pointers (entry y of the list lives at + slot_of[y]*N)
    int nlimbs;
    int slot_of[MKHE_MAX_SLOTS];
    int mod_of_limb[MKHE_MAX_SLOTS];
    int logN;
};

// K1 pass 1 (plain): grid = (16, nlimbs, npolys).  Inputs below 2^60 (canonical, "q" from rotations,
// BFV's lazy multSum limbs) are consumed as they are.
template <int S1>
__global__ void __launch_bounds__(MKHE_NTT_THREADS) k_ntt_pass1(LimbArgs a, const ModC *mods, const ulonglong2 *twf) {
    constexpr int E = 1 << S1;
    const int col = blockIdx.x * MKHE_NTT_THREADS + threadIdx.x, limb = blockIdx.y, poly = blockIdx.z;
    const long N = 1L << a.logN;
    const int mi = a.mod_of_limb[limb];
    const ModC m = mods[mi];
    const NttC c = nttc(m);
    const u64 *src = a.in.p[poly] + (long)a.slot_of[limb] * N + col;
    u64 *dst = a.out.p[poly] + (long)a.slot_of[limb] * N + col;
    u64 v[E];
#pragma unroll
    for (int k = 0; k < E; k++) v[k] = src[(long)k * MKHE_TILE];
    cols_fwd<S1>(v, twf + (long)mi * N, c, m.big != 0);
#pragma unroll
    for (int k = 0; k < E; k++) dst[(long)k * MKHE_TILE] = v[k];
}

// ------------------------------------------------------------------------------------------------
struct Pass2Args {
    PtrList buf;
    int count;                       // instances per poly (digits); instance i lives at buf[i / count] + (i % count) * inst_stride
    int ninst;                       // npolys * count
    long inst_stride;
    int nslots;
    int slots[MKHE_MAX_SLOTS];       // limb slot (offset slot*N)
    int mods[MKHE_MAX_SLOTS];        // modulus index of that slot
    int wslot[MKHE_MAX_SLOTS];       // cost of one tile of that slot (16 = a modulus below 2^57; the 59/60-bit ones sweep and cost more)
    long wstart[MKHE_MAX_SLOTS + 1]; // cumulative cost before slot s; wstart[nslots] = total
    int lazy_out;                    // store the un-reduced outputs (below 2^64, congruent): for forms that only feed the 128-bit MACs
    int logN;
#ifdef MKHE_P2_TIMING
    u64 *timing;                     // development builds: per CTA {start ns, end ns, SM id, tiles}
#endif
};
#ifndef MKHE_P2_GROUPS
#define MKHE_P2_GROUPS 4
#endif
#define MKHE_P2_THREADS (MKHE_P2_GROUPS * MKHE_NTT_THREADS)
#define MKHE_P2_GROUP_BYTES (MKHE_XBUF * 8 + MKHE_TILE * 8)
#define MKHE_P2_NBARS (2 * (1 + MKHE_P2_GROUPS))          // tw_full, in_full[groups], tw_empty, in_empty[groups]
#define MKHE_P2_SMEM (MKHE_TILE * 16 + MKHE_P2_GROUPS * MKHE_P2_GROUP_BYTES + 8 * MKHE_P2_NBARS + 16 + 16 * MKHE_P2_GROUPS)

// the work of a launch is the list of tiles (slot, tile, instance), instance fastest; CTA c of G owns the stretch whose cumulative
// cost lies in [c, c+1) * total / G.  The boundary is the first list index whose cumulative cost reaches the target.
__device__ __forceinline__ long p2_boundary(const Pass2Args &a, long target, long per_slot) {
    int s = 0;
    while (s + 1 < a.nslots && a.wstart[s + 1] <= target) s++;
    long o = (target - a.wstart[s] + a.wslot[s] - 1) / a.wslot[s];
    if (o > per_slot) o = per_slot;
    return (long)s * per_slot + o;
}
struct P2Mail { int off, pad; u64 *ptr; };       // a group's next tile, published by its thread 0

// K1 pass 2: last 11 stages on contiguous tiles, canonical output, in place.
//   Persistent: ONE CTA per SM (two co-resident CTAs do not share an SM fairly: the older one is served first, measured
//   2.9 vs 4.0 us per tile, and the SM is half empty once it leaves), MKHE_P2_GROUPS groups of 128 threads.  Every CTA owns an
//   equal-cost stretch of the tile list, so there is no tail across SMs and the start-up is paid once.  Within the stretch the
//   groups take tiles from a shared counter (a group that the warp schedulers favour simply transforms more tiles).  The tiles
//   of one (slot, tile) pair are consecutive: the pair's 2047 twiddles arrive once by TMA; a group's next 16 KiB input tile is
//   prefetched by TMA into its landing buffer while the current one is transformed.
//   Both TMA targets are recycled through a full / empty mbarrier pair: the readers arrive on the "empty" barrier after their last
//   shared-memory read of the buffer (release), the issuing thread waits for that phase (acquire) before it re-arms the "full"
//   barrier and issues the next cp.async.bulk -- a CTA barrier alone does not order reads that are merely issued before an
//   async-proxy write.
#ifdef MKHE_P2_MAXNREG
__global__ void __maxnreg__(MKHE_P2_MAXNREG) k_ntt_pass2(
#else
__global__ void __launch_bounds__(MKHE_P2_THREADS, 1) k_ntt_pass2(
#endif
    Pass2Args a, const ModC *mods, const ulonglong2 *tiled) {
    MKHE_SMEM(smraw);
    const int tid = threadIdx.x & (MKHE_NTT_THREADS - 1), grp = threadIdx.x / MKHE_NTT_THREADS;
    ulonglong2 *stw = reinterpret_cast<ulonglong2 *>(smraw);                                          // 32 KiB twiddles
    u64 *xbuf = reinterpret_cast<u64 *>(smraw + MKHE_TILE * 16 + grp * MKHE_P2_GROUP_BYTES);           // 17 KiB exchange
    u64 *inbuf = xbuf + MKHE_XBUF;                                                                    // 16 KiB landing
    unsigned char *tail = smraw + MKHE_TILE * 16 + MKHE_P2_GROUPS * MKHE_P2_GROUP_BYTES;
    u64 *bars = reinterpret_cast<u64 *>(tail);                    // [0] twiddles full, [1 + g] input tile of group g full
    u64 *tw_empty = bars + 1 + MKHE_P2_GROUPS, *in_empty = tw_empty + 1;
    int *ctr = reinterpret_cast<int *>(tail + 8 * MKHE_P2_NBARS);
    P2Mail *mail = reinterpret_cast<P2Mail *>(tail + 8 * MKHE_P2_NBARS + 16);
    const long N = 1L << a.logN;
    const int ntiles = (int)(N / MKHE_TILE);
    const int per_slot = ntiles * a.ninst;          // the tile list has at most 40 * 32 * 64 * 40 entries: 32-bit indices (no 64-bit divisions)
    const long total = a.wstart[a.nslots];
    const int w_begin = (int)p2_boundary(a, total * blockIdx.x / gridDim.x, per_slot);
    const int w_end = (int)p2_boundary(a, total * (blockIdx.x + 1) / gridDim.x, per_slot);
    const int nwork = w_end - w_begin;
    auto tile_ptr = [&](int w) -> u64 * {
        const unsigned sidx = (unsigned)w / (unsigned)per_slot;
        const unsigned r = (unsigned)w - sidx * (unsigned)per_slot;
        const unsigned tile = r / (unsigned)a.ninst, inst = r - tile * (unsigned)a.ninst;
        const unsigned poly = inst / (unsigned)a.count;
        return a.buf.p[poly] + (long)(inst - poly * a.count) * a.inst_stride + (long)a.slots[sidx] * N + (long)tile * MKHE_TILE;
    };
    if (threadIdx.x == 0) {
        for (int b = 0; b <= MKHE_P2_GROUPS; b++) mbar_init(&bars[b], 1);
        mbar_init(tw_empty, MKHE_P2_THREADS);
        for (int b = 0; b < MKHE_P2_GROUPS; b++) mbar_init(&in_empty[b], MKHE_NTT_THREADS);
        *ctr = MKHE_P2_GROUPS;                      // group g starts with tile g of the stretch
    }
#ifdef MKHE_P2_TIMING
    if (threadIdx.x == 0 && a.timing) {
        u64 t; u32 sm;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        a.timing[4 * blockIdx.x] = t; a.timing[4 * blockIdx.x + 2] = sm; a.timing[4 * blockIdx.x + 3] = (u64)nwork;
    }
#endif
    __syncthreads();
    int cur = grp;                                  // offset of the group's current tile within the stretch
    u64 *cur_ptr = nullptr;
    if (cur < nwork) {
        cur_ptr = tile_ptr(w_begin + cur);
        if (tid == 0) {
            mbar_expect_tx(&bars[1 + grp], MKHE_TILE * 8);
            tma_load_1d(inbuf, cur_ptr, MKHE_TILE * 8, &bars[1 + grp]);
        }
    }
    const TwShared tw(stw, tid);
    const XAddr x(xbuf, tid);
    u32 tw_phase = 0, in_phase = 0;
    u32 npairs = 0;                                 // (slot, tile) pairs this CTA has finished
    for (int seg = w_begin; seg < w_end;) {         // CTA-uniform walk over the (slot, tile) pairs of the stretch
        const int sidx = seg / per_slot;
        const int tile = (seg - sidx * per_slot) / a.ninst;
        const int pair_end = sidx * per_slot + (tile + 1) * a.ninst;
        seg = pair_end < w_end ? pair_end : w_end;
        const int seg_end = seg - w_begin;
        const int mi = a.mods[sidx];
        const ModC m = mods[mi];
        const NttC c = nttc(m);
        const bool big = m.big != 0;
        if (threadIdx.x == 0) {
            // every thread of the CTA has read the previous pair's twiddles for the last time (they are about to be overwritten)
            if (npairs) mbar_wait(tw_empty, (npairs - 1) & 1);
            mbar_expect_tx(&bars[0], MKHE_TILE * 16);
            tma_load_1d(stw, tiled + ((long)mi * ntiles + tile) * MKHE_TILE, MKHE_TILE * 16, &bars[0]);
        }
        mbar_wait(&bars[0], tw_phase);
        tw_phase ^= 1;
        while (cur < seg_end) {
            mbar_wait(&bars[1 + grp], in_phase);
            u64 v[16];
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = inbuf[k * 128 + tid];
            MKHE_PRE_RELEASE();
            mbar_arrive(&in_empty[grp]);                    // this thread's last read of the landing buffer
            tile_fwd_A(v, tw, c, big);
            // everybody of the group has left the previous tile's exchange buffer and mail slot
            named_sync(1 + grp, MKHE_NTT_THREADS);
            if (tid == 0) {
                const int nxt = atomicAdd(ctr, 1);
                u64 *np = nullptr;
                if (nxt < nwork) {                  // the next tile may belong to the next pair: its data does not depend on twiddles
                    np = tile_ptr(w_begin + nxt);
                    mbar_wait(&in_empty[grp], in_phase);    // all 128 readers have arrived: the buffer may be overwritten
                    mbar_expect_tx(&bars[1 + grp], MKHE_TILE * 8);
                    tma_load_1d(inbuf, np, MKHE_TILE * 8, &bars[1 + grp]);
                }
                mail[grp].off = nxt;
                mail[grp].ptr = np;
            }
            in_phase ^= 1;
            tile_fwd_BC(v, x, tw, c, big, 1 + grp);         // its barrier publishes the mail slot
            u64 *o = cur_ptr + tid * 16;
            cur = mail[grp].off;
            cur_ptr = mail[grp].ptr;
            if (a.lazy_out) {
#pragma unroll
                for (int k = 0; k < 4; k++) st_global_v4(o + 4 * k, v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    st_global_v4(o + 4 * k, canon(v[4 * k], m), canon(v[4 * k + 1], m), canon(v[4 * k + 2], m), canon(v[4 * k + 3], m));
            }
        }
        MKHE_PRE_RELEASE();
        mbar_arrive(tw_empty);                              // this thread is done with the pair's twiddles
        npairs++;
    }
#ifdef MKHE_P2_TIMING
    __syncthreads();
    if (threadIdx.x == 0 && a.timing) {
        u64 t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.timing[4 * blockIdx.x + 1] = t;
    }
#endif
}

// ------------------------------------------------------------------------------------------------
// K3+K2  digit multiply-accumulate FUSED with the inverse pass A:  acc_g = sum_{set} sum_{i<beta} S[set][i] (.) P_g[set][i]
//   (MulCoeffsMontgomery[AndAdd]Lvl loops of mkrlwe/keyswitch_hoisted.go:24-32; two sets for mkbfv/keyswitch_hoisted.go:20-30),
//   one Montgomery reduction, then the first 11 Gentleman-Sande stages of InvNTT[Lazy]Lvl (:34-35) on the accumulator tile while
//   it is still on chip -- the integer work of the transform hides under a kernel that is bound by the key / hoisted-form stream.
//
//   Persistent, ONE CTA per SM, warp specialised:
//     * G consumer groups of 128 threads: group g owns product g of the unit (shared operand S against its private operand P_g).
//     * one producer warp (one elected lane) that only issues TMA bulk copies.
//   A unit = (group of G products, limb slot, tile of 2048 coefficients).  Units are dealt round robin over the CTAs with the
//   product group varying fastest, so the CTAs that stream a common operand (the CRS entry u of every party's products, a
//   rotation's a, x or y of MulRelin) touch the same tile of it at the same time and it comes from DRAM once.
//   Stages of the ring: (G+1) boxes of 16 KiB = the 2048-coefficient tile of S and of every P_g for one term (set, digit).
//   Synchronisation is a full/empty mbarrier pair per stage:
//     producer: wait empty[st] (phase parity) -> mbarrier.arrive.expect_tx full[st] -> cp.async.bulk x (G+1)
//     consumer: wait full[st] -> ld.shared of the stage + MACs -> mbarrier.arrive empty[st] (release: after its last read)
//   The 2047 inverse twiddles of the unit's (modulus, tile) travel the same way (tw_full / tw_empty), requested after the unit's
//   last term so that the producer never waits for pass A of the previous unit while it could prefetch.
//   Output: pass-A values (< 4q) of limb `slot` of out[g], as k_intt_passA leaves them for the ModDown kernels.
// ------------------------------------------------------------------------------------------------
#define MKHE_MI_GROUPS 32                                       // product groups per launch
#define MKHE_MI_BOX (MKHE_TILE * 8)                              // bytes of one operand tile
// MKHE_MI_TWG (A/B builds): the G = 2 kernel reads the unit's inverse twiddles straight from the global table (L2 resident) instead
// of staging them in 32 KiB of shared memory, which buys a fourth ring stage (three stages = 144 KiB in flight per SM instead of 96)
#ifndef MKHE_MI_TWG
#define MKHE_MI_TWG 0
#endif
#define MKHE_MI_TWGLOBAL(G) (MKHE_MI_TWG && (G) == 2)
#define MKHE_MI_STAGES(G) ((G) == 1 ? 4 : (MKHE_MI_TWG ? 4 : 3))
#define MKHE_MI_THREADS(G) (2 * (G) * MKHE_NTT_THREADS + 32)
#define MKHE_MI_NBARS(G) (2 * MKHE_MI_STAGES(G) + 2 + 2 * (G))
#define MKHE_MI_SMEM(G) ((MKHE_MI_TWGLOBAL(G) ? 0 : MKHE_TILE * 16) + MKHE_MI_STAGES(G) * ((G) + 1) * MKHE_MI_BOX + (G) * MKHE_XBUF * 8 + 8 * MKHE_MI_NBARS(G))
struct MacInttArgs {
    const u64 *shared[2][MKHE_MI_GROUPS];
    const u64 *priv[2][MKHE_MI_GROUPS * 2];
    u64 *out[MKHE_MI_GROUPS * 2];
    int nsets, beta, ngroups;
    long digit_stride;       // dmax * N
    int nslots;
    int slots[MKHE_MAX_SLOTS];
    int mods[MKHE_MAX_SLOTS];
    u64 galEl;               // != 0 (RotateHoisted): the columns of every tile are stored permuted, x -> x * galEl mod 2048 (see k_moddown_Q)
    unsigned shared_multi;   // bit grp: the group's shared operand is also streamed by another group of the launch
    u64 priv_multi;          // bit 2 grp + g: the same for the group's private operand g
    int l2_hints;            // use the two masks for L2 eviction priorities (evict_last for common operands, evict_first for the rest)
    int logN;
};
template <int G>
__global__ void __launch_bounds__(MKHE_MI_THREADS(G), 1) k_mac_intt(MacInttArgs a, const ModC *mods, const ulonglong2 *tiled_inv, const ulonglong2 *twi_plain) {
    MKHE_SMEM(smraw);
    constexpr int NS = MKHE_MI_STAGES(G), BOX = MKHE_MI_BOX, STAGE = (G + 1) * BOX, NG = G * MKHE_NTT_THREADS;
    constexpr bool TWG = MKHE_MI_TWGLOBAL(G);          // twiddles from the global table: no staging buffer, no tw_full / tw_empty traffic
    ulonglong2 *stw = reinterpret_cast<ulonglong2 *>(smraw);                                   // 32 KiB inverse twiddles of the unit
    unsigned char *ring = smraw + (TWG ? 0 : MKHE_TILE * 16);
    unsigned char *xb = ring + NS * STAGE;
    u64 *full = reinterpret_cast<u64 *>(xb + G * MKHE_XBUF * 8), *empty = full + NS, *tw_full = empty + NS, *tw_empty = tw_full + 1;
    u64 *ho_full = tw_empty + 1, *ho_empty = ho_full + G;        // hand-over of a product's accumulator tile: MAC group -> transform group
    const long N = 1L << a.logN;
    const int ntiles = (int)(N / MKHE_TILE);
    const int nterms = a.nsets * a.beta;
    const int nunits = a.nslots * ntiles * a.ngroups;
    if (threadIdx.x == 0) {
        for (int st = 0; st < NS; st++) { mbar_init(&full[st], 1); mbar_init(&empty[st], NG); }
        mbar_init(tw_full, 1);
        mbar_init(tw_empty, NG);
        for (int g = 0; g < G; g++) { mbar_init(&ho_full[g], MKHE_NTT_THREADS); mbar_init(&ho_empty[g], MKHE_NTT_THREADS); }
    }
    __syncthreads();
    const int tid = threadIdx.x & (MKHE_NTT_THREADS - 1);
    if (threadIdx.x >= 2 * NG) {
        // ---- producer warp: one lane issues every copy of the CTA
        if (threadIdx.x != 2 * NG) return;
        const u64 pol_once = l2_policy_evict_first(), pol_again = l2_policy_evict_last();
        unsigned T = 0, nu = 0;
        for (int u = blockIdx.x; u < nunits; u += gridDim.x, nu++) {
            const int grp = u % a.ngroups, r = u / a.ngroups, tile = r % ntiles, sidx = r / ntiles;
            const long off = (long)a.slots[sidx] * N + (long)tile * MKHE_TILE;
            const u64 pol_s = (a.shared_multi >> grp) & 1 ? pol_again : pol_once;
            u64 pol_p[G];
#pragma unroll
            for (int g = 0; g < G; g++) pol_p[g] = (a.priv_multi >> (2 * grp + g)) & 1 ? pol_again : pol_once;
            for (int term = 0; term < nterms; term++, T++) {
                const int st = T % NS, t = term / a.beta, i = term - t * a.beta;
                mbar_wait(&empty[st], ((T / NS) & 1) ^ 1);                  // passes at once for the first NS fills
                const long d = (long)i * a.digit_stride + off;
                unsigned char *dst = ring + st * STAGE;
                mbar_expect_tx(&full[st], STAGE);
                if (a.l2_hints) {
                    tma_load_1d_hint(dst, a.shared[t][grp] + d, BOX, &full[st], pol_s);
#pragma unroll
                    for (int g = 0; g < G; g++) tma_load_1d_hint(dst + (1 + g) * BOX, a.priv[t][grp * 2 + g] + d, BOX, &full[st], pol_p[g]);
                } else {
                    tma_load_1d(dst, a.shared[t][grp] + d, BOX, &full[st]);
#pragma unroll
                    for (int g = 0; g < G; g++) tma_load_1d(dst + (1 + g) * BOX, a.priv[t][grp * 2 + g] + d, BOX, &full[st]);
                }
            }
            // the unit's twiddles are needed only after its last term; the previous unit's pass A (the last reader of the
            // twiddle buffer) ran beside this unit's terms and is long over, so this wait does not block
            if (!TWG) {
                mbar_wait(tw_empty, (nu & 1) ^ 1);
                mbar_expect_tx(tw_full, MKHE_TILE * 16);
                tma_load_1d(stw, tiled_inv + ((long)a.mods[sidx] * ntiles + tile) * MKHE_TILE, MKHE_TILE * 16, tw_full);
            }
        }
        return;
    }
    if (threadIdx.x < NG) {
        // ---- multiply-accumulate groups: group g accumulates product g of every unit and hands the reduced tile over
        const int g = threadIdx.x / MKHE_NTT_THREADS;
        u64 *xbuf = reinterpret_cast<u64 *>(xb) + g * MKHE_XBUF;
        unsigned T = 0, nu = 0;
        for (int u = blockIdx.x; u < nunits; u += gridDim.x, nu++) {
            const int sidx = (u / a.ngroups) / ntiles;
            const ModC m = mods[a.mods[sidx]];
            u128 acc[16];
#pragma unroll
            for (int k = 0; k < 16; k++) acc[k] = 0;
            for (int term = 0; term < nterms; term++, T++) {
                const int st = T % NS;
                mbar_wait(&full[st], (T / NS) & 1);
                const ulonglong2 *sS = reinterpret_cast<const ulonglong2 *>(ring + st * STAGE) + tid;
                const ulonglong2 *sP = sS + (1 + g) * (MKHE_TILE / 2);
#pragma unroll
                for (int j = 0; j < 8; j++) {                    // coefficients 2 (128 j + tid) and + 1: conflict-free 16-byte reads
                    const ulonglong2 sv = sS[j * MKHE_NTT_THREADS], pv = sP[j * MKHE_NTT_THREADS];
                    mac128w(acc[2 * j], sv.x, pv.x);
                    mac128w(acc[2 * j + 1], sv.y, pv.y);
                }
                MKHE_PRE_RELEASE();
                mbar_arrive(&empty[st]);                         // after this thread's last read of the stage
            }
            // one reduction per coefficient (the same canonical value as the reference's reduce-every-term loop), stored where the
            // transform's layout C expects it (thread t holds elements 16 t + k; exchange map 2 of the tile rounds)
            mbar_wait(&ho_empty[g], (nu & 1) ^ 1);               // the transform group has left the exchange buffer (previous unit)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int idx = 2 * (j * MKHE_NTT_THREADS + tid);
                u64 *dstp = xbuf + idx + (idx >> 4);
                dstp[0] = mont_reduce(csub(barrett_lazy(hi64(acc[2 * j]), m.q, m.mu), m.q), lo64(acc[2 * j]), m.q, m.qinv);
                dstp[1] = mont_reduce(csub(barrett_lazy(hi64(acc[2 * j + 1]), m.q, m.mu), m.q), lo64(acc[2 * j + 1]), m.q, m.qinv);
            }
            mbar_arrive(&ho_full[g]);
        }
        return;
    }
    // ---- transform groups: inverse pass A of product g's tile, beside the multiply-accumulates of the next unit
    {
        const int g = (threadIdx.x - NG) / MKHE_NTT_THREADS;
        u64 *xbuf = reinterpret_cast<u64 *>(xb) + g * MKHE_XBUF;
        const XAddr x(xbuf, tid);
        const TwShared tw(stw, tid);
        unsigned nu = 0;
        for (int u = blockIdx.x; u < nunits; u += gridDim.x, nu++) {
            const int grp = u % a.ngroups, r = u / a.ngroups, tile = r % ntiles, sidx = r / ntiles;
            const ModC m = mods[a.mods[sidx]];
            u64 v[16];
            mbar_wait(&ho_full[g], nu & 1);
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = x.c2[k];
            if (TWG) {
                const TwGlobalT twg(twi_plain + (long)a.mods[sidx] * N, a.logN - 11, tile, tid);
                tile_inv(v, x, twg, nttc(m), 1 + g);
            } else {
                mbar_wait(tw_full, nu & 1);
                tile_inv(v, x, tw, nttc(m), 1 + g);
                MKHE_PRE_RELEASE();
                mbar_arrive(tw_empty);                           // after this thread's last twiddle read
            }
            if (a.galEl) {
                // Rotation: X -> X^galEl sends coefficient i = x + 2048 row to (x galEl mod 2048) + 2048 row' (+ a sign): columns go
                // to columns.  The column part of the permutation is applied here, on chip, so that the ModDown kernel reads AND
                // writes whole columns contiguously; pass B does not care (every column uses the same twiddles).
                named_sync(1 + g, MKHE_NTT_THREADS);             // every thread has read its pass-A values from the buffer
#pragma unroll
                for (int k = 0; k < 16; k++) xbuf[((u64)(k * 128 + tid) * a.galEl) & (MKHE_TILE - 1)] = v[k];
                named_sync(1 + g, MKHE_NTT_THREADS);
#pragma unroll
                for (int k = 0; k < 16; k++) v[k] = xbuf[k * 128 + tid];
            }
            mbar_arrive(&ho_empty[g]);                           // ... and its last access to the exchange buffer
            u64 *o = a.out[grp * 2 + g] + (long)a.slots[sidx] * N + (long)tile * MKHE_TILE;
#pragma unroll
            for (int k = 0; k < 16; k++) o[k * 128 + tid] = v[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2  inverse pass A on contiguous tiles (head of ringQ.InvNTT[Lazy]Lvl): inputs < 4q in NTT order, outputs in
//   [0,4q) still needing pass B.  grid = (tiles, nslots, nbatch).
// ------------------------------------------------------------------------------------------------
struct InvAArgs {
    PtrList in;              // per batch input, limb slots of stride N
    PtrList out;             // per batch output
    int nbatch;
    int nslots;
    int slots[MKHE_MAX_SLOTS];
    int mods[MKHE_MAX_SLOTS];
    int logN;
};
#define MKHE_PA_GROUPS 2
#define MKHE_PA_THREADS (MKHE_PA_GROUPS * MKHE_NTT_THREADS)
#define MKHE_PA_SMEM (MKHE_TILE * 16 + MKHE_PA_GROUPS * MKHE_XBUF * 8 + 16)

// grid = (tiles, nslots, ceil(nbatch / 2)).  A CTA owns one (tile, limb) and two batch entries, one per 128-thread group: the
// tile's 2047 inverse twiddles arrive by ONE TMA bulk copy (shared by both groups) while the inputs are read from global memory.
__global__ void __launch_bounds__(MKHE_PA_THREADS, 3) k_intt_passA(InvAArgs a, const ModC *mods, const ulonglong2 *tiled_inv) {
    MKHE_SMEM(smraw);
    const int tid = threadIdx.x & (MKHE_NTT_THREADS - 1), grp = threadIdx.x / MKHE_NTT_THREADS;
    ulonglong2 *stw = reinterpret_cast<ulonglong2 *>(smraw);
    u64 *xbuf = reinterpret_cast<u64 *>(smraw + MKHE_TILE * 16 + grp * MKHE_XBUF * 8);
    u64 *bar = reinterpret_cast<u64 *>(smraw + MKHE_TILE * 16 + MKHE_PA_GROUPS * MKHE_XBUF * 8);
    const int tile = blockIdx.x, b = blockIdx.z * MKHE_PA_GROUPS + grp;
    const int slot = a.slots[blockIdx.y], mi = a.mods[blockIdx.y];
    const long N = 1L << a.logN;
    const int ntiles = (int)(N / MKHE_TILE);
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, MKHE_TILE * 16);
        tma_load_1d(stw, tiled_inv + ((long)mi * ntiles + tile) * MKHE_TILE, MKHE_TILE * 16, bar);
    }
    if (b >= a.nbatch) return;
    const ModC m = mods[mi];
    const XAddr x(xbuf, tid);
    u64 v[16];
    const ulonglong2 *p2 = reinterpret_cast<const ulonglong2 *>(a.in.p[b] + (long)slot * N + (long)tile * MKHE_TILE + tid * 16);
#pragma unroll
    for (int k = 0; k < 8; k++) { ulonglong2 xx = ld_cg(p2 + k); v[2 * k] = xx.x; v[2 * k + 1] = xx.y; }
    const TwShared tw(stw, tid);
    mbar_wait(bar, 0);
    tile_inv(v, x, tw, nttc(m), 1 + grp);
    u64 *o = a.out.p[b] + (long)slot * N + (long)tile * MKHE_TILE;
#pragma unroll
    for (int k = 0; k < 16; k++) o[k * 128 + tid] = v[k];
}

// K2 pass B: remaining S1 inverse stages on columns, N^-1, canonical output (in place capable).  grid = (16, nlimbs, npolys)
template <int S1>
__global__ void __launch_bounds__(MKHE_NTT_THREADS) k_intt_passB(LimbArgs a, const ModC *mods, const ulonglong2 *twi) {
    constexpr int E = 1 << S1;
    const int col = blockIdx.x * MKHE_NTT_THREADS + threadIdx.x, limb = blockIdx.y, poly = blockIdx.z;
    const long N = 1L << a.logN;
    const int mi = a.mod_of_limb[limb];
    const ModC m = mods[mi];
    const u64 *src = a.in.p[poly] + (long)a.slot_of[limb] * N + col;
    u64 *dst = a.out.p[poly] + (long)a.slot_of[limb] * N + col;
    u64 v[E];
#pragma unroll
    for (int k = 0; k < E; k++) v[k] = src[(long)k * MKHE_TILE];
    cols_inv<S1>(v, twi + (long)mi * N, nttc(m), m);
#pragma unroll
    for (int k = 0; k < E; k++) dst[(long)k * MKHE_TILE] = v[k];
}

// ------------------------------------------------------------------------------------------------
// K3' x / y generation: x_i = MForm( sum_id MRed(key_id[i], h_id[i]) )  for every digit i and limb
//   (mkrlwe/keyswitch_hoisted.go:79-117).  MForm(MRed(S)) = S mod q, so the kernel accumulates the raw
//   128-bit products and reduces once.  grid.x covers beta*nslots*N/ (256*2) element pairs.
// ------------------------------------------------------------------------------------------------
struct MacPartiesArgs {
    PtrList key, hst;       // per party
    u64 *out;               // swk-shaped
    int nparties, beta, dmax;
    int nslots;
    int slots[MKHE_MAX_SLOTS];
    int mods[MKHE_MAX_SLOTS];
    int logN;
};
__global__ void __launch_bounds__(MKHE_THREADS) k_mac_parties(MacPartiesArgs a, const ModC *mods) {
    const long N = 1L << a.logN;
    const int digit = blockIdx.z, slot = a.slots[blockIdx.y], mi = a.mods[blockIdx.y];
    const ModC m = mods[mi];
    const long off = ((long)digit * a.dmax + slot) * N + ((long)blockIdx.x * MKHE_THREADS + threadIdx.x) * 2;
    u128 a0 = 0, a1 = 0;
    for (int t = 0; t < a.nparties; t++) {
        ulonglong2 kk = *reinterpret_cast<const ulonglong2 *>(a.key.p[t] + off);
        ulonglong2 hh = *reinterpret_cast<const ulonglong2 *>(a.hst.p[t] + off);
        mac128w(a0, kk.x, hh.x);
        mac128w(a1, kk.y, hh.y);
        if ((t & 7) == 7) {                    // keep the high words below q: 8 more products of < 2^124 fit
            a0 = ((u128)csub(barrett_lazy(hi64(a0), m.q, m.mu), m.q) << 64) | lo64(a0);
            a1 = ((u128)csub(barrett_lazy(hi64(a1), m.q, m.mu), m.q) << 64) | lo64(a1);
        }
    }
    const u64 hi0 = csub(barrett_lazy(hi64(a0), m.q, m.mu), m.q), lo0 = lo64(a0);
    const u64 hi1 = csub(barrett_lazy(hi64(a1), m.q, m.mu), m.q), lo1 = lo64(a1);
    u64 r0 = mred(mont_reduce(hi0, lo0, m.q, m.qinv), m.r2, m.q, m.qinv);
    u64 r1 = mred(mont_reduce(hi1, lo1, m.q, m.qinv), m.r2, m.q, m.qinv);
    *reinterpret_cast<ulonglong2 *>(a.out + off) = make_ulonglong2(r0, r1);
}

// ------------------------------------------------------------------------------------------------
// K3'' party-sharded x / y generation fused with the exchange: the multiply-accumulate of k_mac_parties, but every
//   result goes STRAIGHT into the memory of the rank that owns its position (peer stores over NVLink / NVSwitch; cudaIpc-mapped
//   buffers) instead of a local buffer that a collective would then move.  A (digit, limb) row of N coefficients is split into
//   nranks segments of N / nranks; rank r owns segment r of every row.  On the owner, stage[src][which][digit][slot][segment]
//   receives the partial of rank `src`.  k_reduce_gather then sums the nranks partials of the rank's own segments, reduces mod q
//   (the partials are canonical residues in Montgomery form: MForm is linear) and stores the result into the x||y buffers of ALL
//   ranks.  Between the two kernels and after the second one the host enqueues a tiny all-reduce as a stream-ordered barrier.
//   grid as k_mac_parties.
// ------------------------------------------------------------------------------------------------
struct P2PArgs {
    u64 *peer_stage[MKHE_MAX_RANKS];     // stage buffer of every rank (index = owner)
    u64 *peer_xy[MKHE_MAX_RANKS];        // x||y buffer of every rank
    int nranks, rank;
    int which;                           // 0 = x, 1 = y
    long swk_elems;                      // beta_max * dmax * N
};
__global__ void __launch_bounds__(MKHE_THREADS) k_mac_parties_scatter(MacPartiesArgs a, P2PArgs p, const ModC *mods) {
    const long N = 1L << a.logN;
    const int digit = blockIdx.z, slot = a.slots[blockIdx.y], mi = a.mods[blockIdx.y];
    const ModC m = mods[mi];
    const long c = ((long)blockIdx.x * MKHE_THREADS + threadIdx.x) * 2;
    const long off = ((long)digit * a.dmax + slot) * N + c;
    u128 a0 = 0, a1 = 0;
    for (int t = 0; t < a.nparties; t++) {
        ulonglong2 kk = *reinterpret_cast<const ulonglong2 *>(a.key.p[t] + off);
        ulonglong2 hh = *reinterpret_cast<const ulonglong2 *>(a.hst.p[t] + off);
        mac128w(a0, kk.x, hh.x);
        mac128w(a1, kk.y, hh.y);
        if ((t & 7) == 7) {                    // keep the high words below q: 8 more products of < 2^124 fit
            a0 = ((u128)csub(barrett_lazy(hi64(a0), m.q, m.mu), m.q) << 64) | lo64(a0);
            a1 = ((u128)csub(barrett_lazy(hi64(a1), m.q, m.mu), m.q) << 64) | lo64(a1);
        }
    }
    const u64 hi0 = csub(barrett_lazy(hi64(a0), m.q, m.mu), m.q), lo0 = lo64(a0);
    const u64 hi1 = csub(barrett_lazy(hi64(a1), m.q, m.mu), m.q), lo1 = lo64(a1);
    const u64 r0 = mred(mont_reduce(hi0, lo0, m.q, m.qinv), m.r2, m.q, m.qinv);
    const u64 r1 = mred(mont_reduce(hi1, lo1, m.q, m.qinv), m.r2, m.q, m.qinv);
    const long seg = N / p.nranks;
    const int owner = (int)(c / seg);
    const long row = ((long)p.which * (p.swk_elems / N) + (long)digit * a.dmax + slot);     // row index in x||y
    u64 *dst = p.peer_stage[owner] + ((long)p.rank * 2 * (p.swk_elems / N) + row) * seg + (c - (long)owner * seg);
    *reinterpret_cast<ulonglong2 *>(dst) = make_ulonglong2(r0, r1);
}
// grid = (seg / (256 * 2), nslots, 2 * beta): blockIdx.z = which * beta + digit
struct ReduceGatherArgs {
    int beta, dmax, nslots;
    int slots[MKHE_MAX_SLOTS];
    int mods[MKHE_MAX_SLOTS];
    int logN;
};
__global__ void __launch_bounds__(MKHE_THREADS) k_reduce_gather(ReduceGatherArgs a, P2PArgs p, const ModC *mods) {
    const long N = 1L << a.logN;
    const int which = blockIdx.z / a.beta, digit = blockIdx.z - which * a.beta;
    const int slot = a.slots[blockIdx.y], mi = a.mods[blockIdx.y];
    const ModC m = mods[mi];
    const long seg = N / p.nranks;
    const long cs = ((long)blockIdx.x * MKHE_THREADS + threadIdx.x) * 2;          // within the segment
    const long rows = p.swk_elems / N;
    const long row = (long)which * rows + (long)digit * a.dmax + slot;
    u64 s0 = 0, s1 = 0;
    for (int src = 0; src < p.nranks; src++) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p.peer_stage[p.rank] + ((long)src * 2 * rows + row) * seg + cs);
        s0 += v.x;                     // <= 8 canonical residues below 2^60: no overflow
        s1 += v.y;
    }
    const ulonglong2 r = make_ulonglong2(csub(barrett_lazy(s0, m.q, m.mu), m.q), csub(barrett_lazy(s1, m.q, m.mu), m.q));
    const long off = row * N + (long)p.rank * seg + cs;
    for (int dst = 0; dst < p.nranks; dst++) *reinterpret_cast<ulonglong2 *>(p.peer_xy[dst] + off) = r;
}

// ------------------------------------------------------------------------------------------------
// K4 / K8  exact RNS basis conversion (modUpExact = reconstructRNS + multSum,
//   mkrlwe/basis_extension.go:337-357,537-646) and the ModDown combine (:203-229).
//   One thread per coefficient.  The fp64 estimate of the overflow count v is reproduced operation by
//   operation (RN convert, RN divide, RN sequential add, truncate).
// ------------------------------------------------------------------------------------------------
#define MKHE_CONV_MAX 40
struct ConvTable {            // device-resident constants of one (source basis -> target basis) pair
    int n1, n2;
    int src_mod[MKHE_CONV_MAX];             // modulus index of source limb i
    int dst_mod[MKHE_CONV_MAX];             // modulus index of target limb j
    u64 qoverqiinvqi[MKHE_CONV_MAX];        // (Q/q_i)^-1 mod q_i, Montgomery
    u64 qoverqimodp[MKHE_CONV_MAX][MKHE_CONV_MAX];   // [j][i] Q/q_i mod p_j, Montgomery
    u64 vtimesqmodp[MKHE_CONV_MAX][MKHE_CONV_MAX + 1];   // [j][v] (-v*Q) mod p_j
    u64 moddown[MKHE_CONV_MAX];             // [j] p_j - MForm(prod src^-1 mod p_j)   (ModDown only)
    // the key switch's ModDownQPtoQ in closed form (n1 <= 4 special primes): with S = prod src,
    //   MRed(lift + 2 p_j - x, moddown[j]) = (x - lift) S^-1 = x S^-1 - sum_i y_i src_i^-1 + v   (mod p_j),
    // because lift = sum_i y_i (S / src_i) - v S.  Shoup pairs of S^-1 and of -(src_i^-1) mod p_j:
    u64 md_sinv[MKHE_CONV_MAX][2];
    u64 md_nsrcinv[MKHE_CONV_MAX][MKHE_LIFT_MAX_SRC][2];
};
enum { CONV_MODUP = 0, CONV_MODDOWN = 1 };
struct ConvArgs {
    PtrList src;        // per batch: source limbs at src_limb0 + i
    PtrList x;          // CONV_MODDOWN: per batch, the limbs being scaled (target basis) at x_limb0 + j
    PtrList dst;        // per batch: target limbs at dst_limb0 + j
    PtrList acc;        // optional per batch accumulate target (AddLvl): dst = cred(acc + value)
    int src_limb0, x_limb0, dst_limb0;
    int n2_used;        // number of target limbs actually produced (level+1 for ModDownQPtoQ)
    int has_acc;
    int x_is_zero;      // mkbfv Rescale: the P part is zero (MulScalar(.,0,.), mkbfv/basis_extension.go:89)
    int logN;
};
template <int MODE>
__global__ void __launch_bounds__(MKHE_THREADS) k_conv(ConvArgs a, const ConvTable *tabp, const ModC *mods) {
    const ConvTable &tab = *tabp;
    const long N = 1L << a.logN;
    const long x = (long)blockIdx.x * MKHE_THREADS + threadIdx.x;
    const int b = blockIdx.y;
    if (x >= N) return;
    u64 y[MKHE_CONV_MAX];
    double vi = 0.0;
    const u64 *src = a.src.p[b] + (long)a.src_limb0 * N + x;
#pragma unroll 1
    for (int i = 0; i < tab.n1; i++) {
        const ModC &m = mods[tab.src_mod[i]];
        u64 yi = mred(src[(long)i * N], tab.qoverqiinvqi[i], m.q, m.qinv);
        y[i] = yi;
        vi = __dadd_rn(vi, __ddiv_rn(__ull2double_rn(yi), m.qd));
    }
    const u64 v = __double2ull_rz(vi);
#pragma unroll 1
    for (int j = 0; j < a.n2_used; j++) {
        const ModC &m = mods[tab.dst_mod[j]];
        u64 rlo = 0, rhi = 0;
        for (int i = 0; i < tab.n1; i++) mac128(rhi, rlo, y[i], tab.qoverqimodp[j][i]);
        u64 hhi = mulhi(rlo * m.qinv, m.q);
        u64 res = rhi - hhi + m.q + tab.vtimesqmodp[j][v];          // lazy, exactly multSum's value
        if (MODE == CONV_MODDOWN) {
            u64 xv = a.x_is_zero ? 0 : a.x.p[b][(long)(a.x_limb0 + j) * N + x];
            res = mred(res + 2 * m.q - xv, tab.moddown[j], m.q, m.qinv);
            if (a.has_acc) res = csub(a.acc.p[b][(long)(a.dst_limb0 + j) * N + x] + res, m.q);
        }
        a.dst.p[b][(long)(a.dst_limb0 + j) * N + x] = res;
    }
}

// ------------------------------------------------------------------------------------------------
// Team of ranks (limb-sharded ops): every rank owns a set of limb slots of every key, hoisted form and accumulator.
//   Each rank has one block of "team memory" that all ranks map (CUDA IPC across processes, plain pointers inside one process):
//     u64 [0, 8)  arrival flags of the barrier (flag r = last epoch rank r has reached)      [8] status word (1 = barrier timed out)
//     then the exchange areas (element offsets from the base are the same on every rank): P parts of the key-switch accumulators,
//     the polys that have to be complete on every rank before they are decomposed again, the gathered result.
//   Producers write their limbs straight into EVERY rank's copy (peer stores over NVLink / NVSwitch) from the epilogue of the
//   kernel that computes them; k_team_barrier then orders the consumers behind all producers.
// ------------------------------------------------------------------------------------------------
#define MKHE_TEAM_FLAGS 16                 // u64 words reserved at the base of the team memory
struct TeamArgs {
    u64 *peer[MKHE_MAX_RANKS];             // base of rank r's team memory as mapped by this rank
    int nranks, rank;
};
#ifdef MKHE_EMU
__device__ __forceinline__ void st_release_sys(u64 *p, u64 v) { *p = v; }
__device__ __forceinline__ u64 ld_acquire_sys(const u64 *p) { return *p; }
__device__ __forceinline__ u64 global_timer_ns() { return 0; }
__device__ __forceinline__ void fence_system() {}
#else
__device__ __forceinline__ void st_release_sys(u64 *p, u64 v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_acquire_sys(const u64 *p) { u64 v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ u64 global_timer_ns() { u64 t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void fence_system() { __threadfence_system(); }
#endif
struct GatherArgs {
    PtrList src;             // per poly (local)
    long dst_off[MKHE_MAX_PARTIES_K];      // per poly: element offset of its image in the team memory
    int nslots;
    int slots[MKHE_MAX_SLOTS];
    int logN;
};
// limbs `slots` of src[b] -> dst[b] on this device (the second half of mkhe_team_allgather).  grid = (N/512, nslots, npolys)
__global__ void __launch_bounds__(MKHE_THREADS) k_copy_limbs(GatherArgs a, PtrList dst) {
    const long N = 1L << a.logN;
    const long off = (long)a.slots[blockIdx.y] * N + ((long)blockIdx.x * MKHE_THREADS + threadIdx.x) * 2;
    *reinterpret_cast<ulonglong2 *>(dst.p[blockIdx.z] + off) = *reinterpret_cast<const ulonglong2 *>(a.src.p[blockIdx.z] + off);
}
// Barrier among the ranks, in stream order: everything this rank enqueued before it (peer stores included) is complete and
// visible before its flag is raised on the peers; the kernel returns when every rank has raised its flag here.  One warp; lane r
// talks to rank r.  A rank that never arrives (a crashed peer) ends the wait after `timeout_ns` with the status word set instead
// of hanging the device.
__global__ void __launch_bounds__(32) k_team_barrier(TeamArgs t, u64 epoch, u64 timeout_ns) {
    const int lane = threadIdx.x;
    fence_system();
    if (lane < t.nranks) st_release_sys(t.peer[lane] + t.rank, epoch);
    if (lane < t.nranks) {
        const u64 *mine = t.peer[t.rank] + lane;
        const u64 t0 = global_timer_ns();
        while (ld_acquire_sys(mine) < epoch) {
#ifdef MKHE_EMU
            break;                         // the emulator runs one rank at a time: only single-rank teams get here
#else
            if (global_timer_ns() - t0 > timeout_ns) { t.peer[t.rank][MKHE_MAX_RANKS] = 1; break; }
            __nanosleep(64);
#endif
        }
    }
    fence_system();
}
// The barrier as it is normally issued: this kernel only SIGNALS (one system-scope release-add on every peer's arrival counter, after
// a system fence that orders the rank's earlier peer stores), the wait is a stream memory operation on the rank's own counter
// (cuStreamWaitValue64, >= epoch * (nranks - 1), with a flush of outstanding remote writes) -- no thread spins on an SM, so the
// barrier does not depend on kernels of different streams being co-resident.  Word MKHE_TEAM_COUNTER of the team memory.
#define MKHE_TEAM_COUNTER 9
#ifdef MKHE_EMU
__device__ __forceinline__ void red_release_sys_add(u64 *p, u64 v) { *p += v; }
#else
__device__ __forceinline__ void red_release_sys_add(u64 *p, u64 v) { asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
#endif
__global__ void __launch_bounds__(32) k_team_signal(TeamArgs t) {
    const int lane = threadIdx.x;
    fence_system();
    if (lane < t.nranks && lane != t.rank) red_release_sys_add(t.peer[lane] + MKHE_TEAM_COUNTER, 1);
}
// result gather: limb `slot` of every listed poly goes to the same place of every rank's gather area.  grid = (N/512, nslots, npolys)
__global__ void __launch_bounds__(MKHE_THREADS) k_team_gather(GatherArgs a, TeamArgs t) {
    const long N = 1L << a.logN;
    const long off = (long)a.slots[blockIdx.y] * N + ((long)blockIdx.x * MKHE_THREADS + threadIdx.x) * 2;
    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(a.src.p[blockIdx.z] + off);
    for (int r = 0; r < t.nranks; r++) *reinterpret_cast<ulonglong2 *>(t.peer[r] + a.dst_off[blockIdx.z] + off) = v;
}

// ------------------------------------------------------------------------------------------------
// K2+K4  tail of the key switch: InvNTT pass B fused with ModDownQPtoQ, the accumulation into the target and (for rotations)
//   the automorphism  (mkrlwe/keyswitch_hoisted.go:34-39,217-245, basis_extension.go:192-232 with the P -> Q lift of
//   :337-357,537-646).  Works on the per-product QP accumulators after pass A; every accumulator has nP spare limb slots.
//   k_moddown_P : per product and P limb i: pass B, then y_i = x_i * (P/p_i)^-1 mod p_i in place of limb nQ+i, and the term
//                 fl(fl(y_i) / fl(p_i)) of the fp64 overflow estimate into spare slot i (reconstructRNS adds exactly these
//                 terms, in limb order, RN).  grid = (16, nproducts, nP)
//   k_moddown_Q : per target poly and Q limb j: for every product of the target: pass B of limb j, v = trunc(sum of the terms),
//                 lift of the P part (multSum, lazy value reproduced), (lift - x) * -P^-1, exact adds onto `src` (the target's
//                 current contents, another poly, or zero); the result is stored through X -> X^galEl with the reference's
//                 unreduced negation when galEl != 0.  grid = (16, level+1, ntargets)
// ------------------------------------------------------------------------------------------------
#define MKHE_MD_TARGETS 66
#define MKHE_MD_PRODUCTS 132
struct ModDownPArgs {
    u64 *acc[MKHE_MD_PRODUCTS];
    int np_limbs, p_slot0, vslot;
    int nplist, plist[4];                  // the P limbs this launch handles (blockIdx.z indexes the list): all of them, or a rank's share
    long pp_off[MKHE_MD_PRODUCTS];         // team mode: element offset of the product's P part (y_0.., then the fp64 terms) in the team memory
    TeamArgs team;                         // nranks = 0: single GPU, results stay in the accumulator's own P and spare slots
    int logN;
};
template <int S1>
__global__ void __launch_bounds__(MKHE_NTT_THREADS) k_moddown_P(ModDownPArgs a, const ConvTable *tabp, const ModC *mods, const ulonglong2 *twi) {
    constexpr int E = 1 << S1;
    const ConvTable &tab = *tabp;
    const long N = 1L << a.logN;
    const int col = blockIdx.x * MKHE_NTT_THREADS + threadIdx.x, i = a.plist[blockIdx.z];
    u64 *base = a.acc[blockIdx.y] + col;
    const int mi = tab.src_mod[i];
    const ModC m = mods[mi];
    u64 *p = base + (long)(a.p_slot0 + i) * N;
    u64 v[E];
#pragma unroll
    for (int k = 0; k < E; k++) v[k] = ld_cg(p + (long)k * MKHE_TILE);
    cols_inv<S1>(v, twi + (long)mi * N, nttc(m), m);
    const u64 f = tab.qoverqiinvqi[i];
    if (a.team.nranks == 0) {
        u64 *vp = base + (long)(a.vslot + i) * N;
#pragma unroll
        for (int k = 0; k < E; k++) {
            const u64 y = mred(v[k], f, m.q, m.qinv);
            p[(long)k * MKHE_TILE] = y;
            vp[(long)k * MKHE_TILE] = (u64)__double_as_longlong(__ddiv_rn(__ull2double_rn(y), m.qd));
        }
    } else {
        // limb sharding: this rank owns P limb i; every rank's ModDown of its Q limbs needs it -> peer stores into every rank's
        // copy of the product's P part (the all-gather of SURVEY 8e (2), fused into the kernel that produces the data)
        const long o = a.pp_off[blockIdx.y] + col;
#pragma unroll
        for (int k = 0; k < E; k++) {
            // only y travels (half the bytes of the exchange that loads the P owners' links most); the receivers redo the fp64
            // division, operation for operation
            const u64 y = mred(v[k], f, m.q, m.qinv);
            for (int r = 0; r < a.team.nranks; r++) a.team.peer[r][o + (long)i * N + (long)k * MKHE_TILE] = y;
        }
    }
}

// limb-sharded runs: after the exchange every rank holds y_0 .. y_{nP-1} of every product; the overflow estimate
//   v = trunc(sum_i fl(fl(y_i) / fl(p_i)))  (reconstructRNS, basis_extension.go:548-556: limb order, RN adds, truncation)
// is the same for every Q limb, so it is computed ONCE per coefficient here instead of in every (limb, product) CTA of
// k_moddown_Q (the fp64 divisions were half of that kernel's time on a team).  Stored as a u64 in the first fp64-term slot of the
// product's P part (pp + nP N).  grid = (N/256, nproducts)
struct ModDownOvArgs {
    u64 *pp[MKHE_MD_PRODUCTS];
    int np_limbs;
    int p_mod0;              // modulus index of the first special prime
    int logN;
};
__global__ void __launch_bounds__(MKHE_THREADS) k_moddown_ov(ModDownOvArgs a, const ModC *mods) {
    const long N = 1L << a.logN;
    const long j = (long)blockIdx.x * MKHE_THREADS + threadIdx.x;
    u64 *pp = a.pp[blockIdx.y];
    double vi = 0.0;
    for (int i = 0; i < a.np_limbs; i++)
        vi = __dadd_rn(vi, __ddiv_rn(__ull2double_rn(ld_cg(pp + (long)i * N + j)), mods[a.p_mod0 + i].qd));
    pp[(long)a.np_limbs * N + j] = __double2ull_rz(vi);
}

struct ModDownQArgs {
    u64 *dst[MKHE_MD_TARGETS];
    const u64 *src[MKHE_MD_TARGETS];       // the sum starts from this poly (AddLvl onto the target or onto another poly); nullptr = zero
    int first[MKHE_MD_TARGETS + 1];        // products of target t: acc[first[t]] .. acc[first[t+1]-1]
    int split[MKHE_MD_TARGETS];            // 1, 2, 4 or 8: the products of target t are dealt over `split` thread lanes of a CTA
    const u64 *acc[MKHE_MD_PRODUCTS];
    const u64 *pp[MKHE_MD_PRODUCTS];       // P part of product s: y_i at pp + i N, fp64 term i at pp + (nP + i) N (the accumulator's own P and
                                           // spare slots, or the rank's copy in the team memory)
    int nqlist, qlist[MKHE_MAX_SLOTS];     // the Q limbs this launch handles (blockIdx.y indexes the list)
    long dst_team_off[MKHE_MD_TARGETS];    // >= 0: the target has to be complete on every rank (it is decomposed again): stored into
                                           // every rank's team memory at this element offset instead of dst
    TeamArgs team;
    int np_limbs;
    u64 galEl, galInv;                     // != 0 (RotateHoisted): the result goes through X -> X^galEl; the accumulators arrive with
                                           // their columns already permuted (k_mac_intt); galInv = galEl^-1 mod 2N
    int logN;
};
#ifndef MKHE_MDQ_MINB
#define MKHE_MDQ_MINB 4               // resident CTAs per SM the register allocation of k_moddown_Q aims at (A/B builds override it)
#endif
#define MKHE_MDQ_SMEM(S1) (((size_t)8 << (S1)) * MKHE_NTT_THREADS)
// grid = (16 * max split, level+1, ntargets).  A target with several products (c_0 of a rotation or of the relinearisation: one
// product per party) would otherwise be one long serial chain per thread while the single-product targets finish early: its
// products are dealt over `split` lanes of the CTA (thread = (lane, column), 128 / split columns per CTA), every lane sums its
// share and the partial sums meet in shared memory (exact modular adds, any order).
template <int S1, int NP, bool TEAM>
__global__ void __launch_bounds__(MKHE_NTT_THREADS, (S1 <= 4 ? MKHE_MDQ_MINB : 2)) k_moddown_Q(ModDownQArgs a, const ConvTable *tabp, const ModC *mods, const ulonglong2 *twi) {
    constexpr int E = 1 << S1, HB = E < 8 ? E : 8;
    MKHE_SMEM(smraw);                          // E * 128 * 8 bytes
    u64 (*rowbuf)[MKHE_NTT_THREADS] = reinterpret_cast<u64 (*)[MKHE_NTT_THREADS]>(smraw);
    const ConvTable &tab = *tabp;
    const long N = 1L << a.logN;
    const int j = a.qlist[blockIdx.y], t = blockIdx.z;
    const int PP = a.split[t], cols_per = MKHE_NTT_THREADS / PP;
    if ((int)blockIdx.x >= (MKHE_TILE / MKHE_NTT_THREADS) * PP) return;
    const int lane = threadIdx.x / cols_per, cin = threadIdx.x - lane * cols_per;
    const int col = blockIdx.x * cols_per + cin;
    const ModC m = mods[tab.dst_mod[j]];          // == modulus j
    const NttC c = nttc(m);
    // ModDownQPtoQ in closed form (see ConvTable): d = x P^-1 - sum_i y_i p_i^-1 + v mod q_j, three lazy Shoup products and one
    // canonical reduction instead of multSum's 128-bit accumulation, its Montgomery fold and a full MRed -- the same canonical
    // residue the reference stores (basis_extension.go:203-229)
    u64 sinv[2], nsi[NP][2];
    sinv[0] = tab.md_sinv[j][0]; sinv[1] = tab.md_sinv[j][1];
#pragma unroll
    for (int i = 0; i < NP; i++) { nsi[i][0] = tab.md_nsrcinv[j][i][0]; nsi[i][1] = tab.md_nsrcinv[j][i][1]; }
    // with an automorphism this thread works on stored column `col` = original column xo of the products (pass B treats every
    // column alike); the polynomial the sum starts from is not permuted and is read at the original positions
    const int xo = a.galEl ? (int)(((u64)col * a.galInv) & (MKHE_TILE - 1)) : col;
    u64 r[E];
    if (a.src[t] && lane == 0) {
        const u64 *sp = a.src[t] + (long)j * N + xo;
#pragma unroll
        for (int k = 0; k < E; k++) r[k] = ld_cg(sp + (long)k * MKHE_TILE);
    } else {
#pragma unroll
        for (int k = 0; k < E; k++) r[k] = 0;
    }
#pragma unroll 1
    for (int s = a.first[t] + lane; s < a.first[t + 1]; s += PP) {
        const u64 *src = a.acc[s] + col;
        const u64 *pps = a.pp[s] + col;
        u64 v[E];
#pragma unroll
        for (int k = 0; k < E; k++) v[k] = ld_cg(src + (long)j * N + (long)k * MKHE_TILE);
        cols_inv<S1>(v, twi + (long)tab.dst_mod[j] * N, c, m);
#pragma unroll
        for (int h = 0; h < E; h += HB) {
            u64 y[NP][HB], ov[HB];
#pragma unroll
            for (int k = 0; k < HB; k++) {
                if (TEAM) {                       // the overflow estimate was made once per coefficient by k_moddown_ov
#pragma unroll
                    for (int i = 0; i < NP; i++) y[i][k] = ld_cg(pps + (long)i * N + (long)(h + k) * MKHE_TILE);
                    ov[k] = ld_cg(pps + (long)a.np_limbs * N + (long)(h + k) * MKHE_TILE);
                } else {
                    double vi = 0.0;              // reconstructRNS: vi += fl(y_i)/fl(p_i), limb order, RN; then truncation
#pragma unroll
                    for (int i = 0; i < NP; i++) {
                        y[i][k] = ld_cg(pps + (long)i * N + (long)(h + k) * MKHE_TILE);
                        vi = __dadd_rn(vi, __longlong_as_double((long long)ld_cg(pps + (long)(a.np_limbs + i) * N + (long)(h + k) * MKHE_TILE)));
                    }
                    ov[k] = __double2ull_rz(vi);
                }
            }
#pragma unroll
            for (int k = 0; k < HB; k++) {
                u64 d = shoup4(v[h + k], sinv[0], sinv[1], c.nq) + ov[k];        // each product below 4q: the sum stays below 2^64
#pragma unroll
                for (int i = 0; i < NP; i++) {
                    d += shoup4(y[i][k], nsi[i][0], nsi[i][1], c.nq);
                    if (NP > 2 && i == 1) d = canon(d, m);      // at most three lazy products (12q < 2^64) between reductions
                }
                r[h + k] = csub(r[h + k] + canon(d, m), m.q);
            }
        }
    }
    if (PP > 1) {                              // the lanes' partial sums meet in lane 0 (CTA-uniform branch)
        if (lane > 0) {
#pragma unroll
            for (int k = 0; k < E; k++) rowbuf[k][threadIdx.x] = r[k];
        }
        __syncthreads();
        if (lane == 0) {
#pragma unroll 1
            for (int l = 1; l < PP; l++) {
#pragma unroll
                for (int k = 0; k < E; k++) r[k] = csub(r[k] + rowbuf[k][l * cols_per + cin], m.q);   // r[k] may be a non-canonical start value (q): one subtraction, like AddLvl
            }
        }
        __syncthreads();
    }
    u64 *dst = a.dst[t] + (long)j * N;
    if (TEAM && a.dst_team_off[t] >= 0) {      // limb sharding: limb j of this poly goes to every rank (peer stores)
        // The CTA's E x cols_per block is staged in shared memory and leaves as 32-byte stores (4 coefficients per thread and
        // store): an NVLink write carries the same bytes in a quarter of the requests of the 8-byte-per-thread form.
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < E; k++) rowbuf[k][cin] = r[k];
        }
        __syncthreads();
        const int per_row = cols_per >> 2;       // cols_per is 128, 64, 32 or 16
        const long o = a.dst_team_off[t] + (long)j * N + (long)blockIdx.x * cols_per;
        for (int c = threadIdx.x; c < E * per_row; c += MKHE_NTT_THREADS) {
            const int row = c / per_row, q4 = (c - row * per_row) << 2;
            const u64 v0 = rowbuf[row][q4], v1 = rowbuf[row][q4 + 1], v2 = rowbuf[row][q4 + 2], v3 = rowbuf[row][q4 + 3];
            const long e = o + (long)row * MKHE_TILE + q4;
            for (int rk = 0; rk < a.team.nranks; rk++) st_global_v4(a.team.peer[rk] + e, v0, v1, v2, v3);
        }
    } else if (a.galEl == 0) {
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < E; k++) dst[col + (long)k * MKHE_TILE] = r[k];
        }
    } else {
        // keyswitch_hoisted.go:217-245: out[(i*galEl) & (N-1)] = ((i*galEl) >> logN) & 1 ? q - c : c   (c = 0 is stored as q).
        // i = xo + 2048 k lands in column (xo galEl mod 2048) = col, row ((i galEl) >> 11) mod E: the rows of the column are permuted
        // through shared memory so that every store of the warp is contiguous.
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < E; k++) {
                const u64 raw = (u64)(xo + k * MKHE_TILE) * a.galEl;
                rowbuf[(raw >> 11) & (E - 1)][cin] = ((raw >> a.logN) & 1) ? m.q - r[k] : r[k];
            }
        }
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < E; k++) dst[col + (long)k * MKHE_TILE] = rowbuf[k][cin];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// element-wise kernels
// ------------------------------------------------------------------------------------------------
// K5 tensor products in the NTT domain (mkrlwe/keyswitch_hoisted.go:119-140): all operands canonical;
//   MForm + MulCoeffsMontgomery = plain product mod q.
//   out_0 = A0*B0 ; out_t = [has0] B0*A_t + [has1] A0*B_t
struct TensorArgs {
    const u64 *A0, *B0;           // NTT(op0_0), NTT(op1_0)
    PtrList A, B;                 // per output party: NTT(op0_id) / NTT(op1_id) or nullptr
    long strideA, strideB;        // distance between consecutive limbs of A.p[t] / B.p[t]: N for a poly, (dmax + 1) * N when the
                                  // operand is read from the diagonal of the party's freshly hoisted form (digit i, limb i)
    PtrList out;                  // [0] = component "0", [1+t]
    int nout;                     // parties in the output
    int nlimbs;
    int limbs[MKHE_MAX_SLOTS];    // the Q limbs of this launch (all of them up to the level, or a rank's share); limb == modulus index
    int logN;
};
__global__ void __launch_bounds__(MKHE_THREADS) k_tensor(TensorArgs a, const ModC *mods) {
    const long N = 1L << a.logN;
    const int limb = a.limbs[blockIdx.y];
    const ModC m = mods[limb];
    const long x = (long)blockIdx.x * MKHE_THREADS + threadIdx.x;
    const long off = (long)limb * N + x;
    const long offA = (long)limb * a.strideA + x, offB = (long)limb * a.strideB + x;
    const u64 a0 = mred(a.A0[off], m.r2, m.q, m.qinv);      // MForm
    const u64 b0 = mred(a.B0[off], m.r2, m.q, m.qinv);
    a.out.p[0][off] = mred(a0, a.B0[off], m.q, m.qinv);
    for (int t = 0; t < a.nout; t++) {
        u64 r = 0;
        if (a.A.p[t]) r = mred(b0, a.A.p[t][offA], m.q, m.qinv);
        if (a.B.p[t]) r = csub(r + mred(a0, a.B.p[t][offB], m.q, m.qinv), m.q);
        a.out.p[1 + t][off] = r;
    }
}

// ringQ.AddLvl / SubLvl
template <bool SUB>
__global__ void __launch_bounds__(MKHE_THREADS) k_addsub(const u64 *x, const u64 *y, u64 *out, LimbArgs a, const ModC *mods) {
    const long N = 1L << a.logN;
    const int limb = blockIdx.y;
    const u64 q = mods[a.mod_of_limb[limb]].q;
    const long off = (long)a.slot_of[limb] * N + (long)blockIdx.x * MKHE_THREADS + threadIdx.x;
    out[off] = SUB ? csub(x[off] + q - y[off], q) : csub(x[off] + y[off], q);
}

// x mod q for sums of <= 8 canonical residues (after the party-sharded all-reduce).  grid = (N/256, nlimbs, nbufs)
__global__ void __launch_bounds__(MKHE_THREADS) k_reduce(LimbArgs a, const ModC *mods) {
    const long N = 1L << a.logN;
    const int limb = blockIdx.y, b = blockIdx.z;
    const ModC m = mods[a.mod_of_limb[limb]];
    const long off = (long)a.slot_of[limb] * N + (long)blockIdx.x * MKHE_THREADS + threadIdx.x;
    a.out.p[b][off] = csub(barrett_lazy(a.in.p[b][off], m.q, m.mu), m.q);
}

// K6 DivRoundByLastModulusLvl (lattigo ring/scaling.go; call site mkckks/evaluator.go:388).
//   Writes (x_l + h) mod q_l back into the input's last limb like lattigo does.
struct RescaleArgs {
    PtrList in, out;
    int level;                       // last limb index l
    u64 rescale[MKHE_MAX_SLOTS];     // q_i - MForm(q_l^-1 mod q_i)
    u64 halfneg[MKHE_MAX_SLOTS];     // q_i - (h mod q_i)
    int logN;
};
__global__ void __launch_bounds__(MKHE_THREADS) k_rescale(RescaleArgs a, const ModC *mods) {
    const long N = 1L << a.logN;
    const long x = (long)blockIdx.x * MKHE_THREADS + threadIdx.x;
    const int b = blockIdx.y;
    const u64 ql = mods[a.level].q;
    u64 *last = a.in.p[b] + (long)a.level * N + x;
    const u64 xl = csub(*last + ((ql - 1) >> 1), ql);
    for (int i = 0; i < a.level; i++) {
        const ModC &m = mods[i];
        u64 t = csub(barrett_lazy(xl, m.q, m.mu), m.q);
        u64 xi = a.in.p[b][(long)i * N + x];
        a.out.p[b][(long)i * N + x] = mred(t + a.halfneg[i] + 2 * m.q - xi, a.rescale[i], m.q, m.qinv);
    }
    *last = xl;
}

// K7 coefficient-domain automorphism X -> X^galEl with the reference's unreduced negation
//   (mkrlwe/keyswitch_hoisted.go:217-245: q - c, so c = 0 is stored as q).  grid = (N/256, nlimbs, npolys)
__global__ void __launch_bounds__(MKHE_THREADS) k_automorph(LimbArgs a, u64 galEl, const ModC *mods) {
    const long N = 1L << a.logN;
    const int limb = blockIdx.y, poly = blockIdx.z;
    const u64 q = mods[a.mod_of_limb[limb]].q;
    const u64 i = (u64)blockIdx.x * MKHE_THREADS + threadIdx.x;
    const u64 raw = i * galEl;
    const u64 idx = raw & (u64)(N - 1);
    const u64 neg = (raw >> a.logN) & 1;
    const u64 c = a.in.p[poly][(long)a.slot_of[limb] * N + i];
    a.out.p[poly][(long)a.slot_of[limb] * N + idx] = neg ? q - c : c;
}

// MulCoeffsMontgomery by a per-limb constant (mkbfv Rescale's mFormQMul, MulScalar by t in Quantize)
struct ScaleArgs {
    PtrList in, out;
    int nlimbs;
    int mod_of_limb[MKHE_MAX_SLOTS];
    u64 cmont[MKHE_MAX_SLOTS];       // constant in Montgomery form
    int logN;
};
__global__ void __launch_bounds__(MKHE_THREADS) k_scale(ScaleArgs a, const ModC *mods) {
    const long N = 1L << a.logN;
    const int limb = blockIdx.y, b = blockIdx.z;
    const ModC m = mods[a.mod_of_limb[limb]];
    const long off = (long)limb * N + (long)blockIdx.x * MKHE_THREADS + threadIdx.x;
    a.out.p[b][off] = mred(a.in.p[b][off], a.cmont[limb], m.q, m.qinv);
}

// MultByConst (mkckks/evaluator.go:117-198): coefficients [0, N/2) times c_first, [N/2, N) times c_second (Montgomery form);
// ringQ.NegLvl: q - x, unreduced (0 -> q).  grid = (N/256, nlimbs, npolys)
struct ConstArgs {
    PtrList in, out;
    int nlimbs;
    int mod_of_limb[MKHE_MAX_SLOTS];
    u64 c_first[MKHE_MAX_SLOTS], c_second[MKHE_MAX_SLOTS];
    int logN;
};
__global__ void __launch_bounds__(MKHE_THREADS) k_mul_const(ConstArgs a, const ModC *mods) {
    const long N = 1L << a.logN;
    const int limb = blockIdx.y, b = blockIdx.z;
    const ModC m = mods[a.mod_of_limb[limb]];
    const long j = (long)blockIdx.x * MKHE_THREADS + threadIdx.x;
    const long off = (long)limb * N + j;
    a.out.p[b][off] = mred(a.in.p[b][off], j < (N >> 1) ? a.c_first[limb] : a.c_second[limb], m.q, m.qinv);
}
__global__ void __launch_bounds__(MKHE_THREADS) k_neg(LimbArgs a, const ModC *mods) {
    const long N = 1L << a.logN;
    const int limb = blockIdx.y, b = blockIdx.z;
    const u64 q = mods[a.mod_of_limb[limb]].q;
    const long off = (long)a.slot_of[limb] * N + (long)blockIdx.x * MKHE_THREADS + threadIdx.x;
    a.out.p[b][off] = q - a.in.p[b][off];
}

// Decryptor (mkrlwe/decryptor.go:26-66): MulCoeffsMontgomeryLvl of NTT(c_id) by the secret key (NTT domain, Montgomery form), and
// the final accumulation c_0 + sum_id InvNTT(...) with ReduceLvl (canonical output).  grid = (N/256, nlimbs[, npolys])
__global__ void __launch_bounds__(MKHE_THREADS) k_mul_mont(LimbArgs a, PtrList other, const ModC *mods) {
    const long N = 1L << a.logN;
    const int limb = blockIdx.y, b = blockIdx.z;
    const ModC m = mods[a.mod_of_limb[limb]];
    const long off = (long)a.slot_of[limb] * N + (long)blockIdx.x * MKHE_THREADS + threadIdx.x;
    a.out.p[b][off] = mred(a.in.p[b][off], other.p[b][off], m.q, m.qinv);
}
__global__ void __launch_bounds__(MKHE_THREADS) k_decrypt_sum(LimbArgs a, const u64 *c0, int nparts, u64 *pt, const ModC *mods) {
    const long N = 1L << a.logN;
    const int limb = blockIdx.y;
    const ModC m = mods[a.mod_of_limb[limb]];
    const long off = (long)a.slot_of[limb] * N + (long)blockIdx.x * MKHE_THREADS + threadIdx.x;
    u64 acc = c0[off];                                   // may be non-canonical (q from a rotation): AddLvl = one conditional subtraction
    for (int t = 0; t < nparts; t++) acc = csub(acc + a.in.p[t][off], m.q);
    pt[off] = csub(barrett_lazy(acc, m.q, m.mu), m.q);   // ReduceLvl
}

// BFV tensor products in ring R (mkbfv/keyswitch_hoisted.go:144-181): out = A*B (+ C*D), operands may be lazy (<4q)
__global__ void __launch_bounds__(MKHE_THREADS) k_mul2(const u64 *A, const u64 *B, const u64 *Cc, const u64 *D, u64 *out,
                                                        LimbArgs a, const ModC *mods) {
    const long N = 1L << a.logN;
    const int limb = blockIdx.y;
    const ModC m = mods[a.mod_of_limb[limb]];
    const long off = (long)a.slot_of[limb] * N + (long)blockIdx.x * MKHE_THREADS + threadIdx.x;
    u64 r = mred(mred(A[off], m.r2, m.q, m.qinv), B[off], m.q, m.qinv);
    if (Cc) r = csub(r + mred(mred(Cc[off], m.r2, m.q, m.qinv), D[off], m.q, m.qinv), m.q);
    out[off] = r;
}

// ------------------------------------------------------------------------------------------------
// Key generation and encryption on the device (mkrlwe/keygen.go:44-327, encryptor.go:55-118).
// Randomness: the counter-based samplers of include/mkhe_prng.h -- one thread per coefficient, any order, the CPU oracle draws
// the same values.
// ------------------------------------------------------------------------------------------------
enum { SAMPLE_TERNARY = 0, SAMPLE_GAUSS = 1, SAMPLE_UNIFORM = 2 };
struct SampleArgs {
    u64 *out;                // instance i at out + i * inst_stride, limb slot s at + s * N
    long inst_stride;
    u64 seed, stream0;       // small samplers: instance i = stream0 + i * stream_stride; uniform: (instance i, slot s) = stream0 + i * stream_stride + s
    long stream_stride;
    int kind;
    int mform;               // uniform: store MForm(x) (the CRS: params.go:54-56)
    u64 thr53;               // ternary: floor(P(0) * 2^53)
    const u64 *add_to;       // optional (encryption: ReadAndAddLvl): out = CRed(add_to + sample) on the same layout
    const u64 *add_pt;       // optional second addend (the plaintext)
    int nslots;
    int slots[MKHE_MAX_SLOTS];
    int logN;
};
// grid = (N/256, ninstances).  Small samplers: one draw per coefficient, written to every listed limb as ring.Sampler.Read +
// ExtendBasisSmallNormAndCenter do (v >= 0 ? v : q + v).
__global__ void __launch_bounds__(MKHE_THREADS) k_sample(SampleArgs a, const ModC *mods) {
    const long N = 1L << a.logN;
    const u64 j = (u64)blockIdx.x * MKHE_THREADS + threadIdx.x;
    const int inst = blockIdx.y;
    u64 *o = a.out + (long)inst * a.inst_stride + j;
    if (a.kind == SAMPLE_UNIFORM) {
        for (int s = 0; s < a.nslots; s++) {
            const int slot = a.slots[s];
            const ModC m = mods[slot];
            const u64 x = mkhe_sample_uniform(a.seed, a.stream0 + (u64)inst * (u64)a.stream_stride + (u64)slot, j, m.q);
            o[(long)slot * N] = a.mform ? mred(x, m.r2, m.q, m.qinv) : x;
        }
        return;
    }
    const int v = a.kind == SAMPLE_TERNARY ? mkhe_sample_ternary(a.seed, a.stream0 + (u64)inst * (u64)a.stream_stride, j, a.thr53)
                                           : mkhe_sample_gaussian(a.seed, a.stream0 + (u64)inst * (u64)a.stream_stride, j);
    for (int s = 0; s < a.nslots; s++) {
        const int slot = a.slots[s];
        const u64 q = mods[slot].q;
        u64 x = v >= 0 ? (u64)v : q - (u64)(-v);
        const long off = (long)inst * a.inst_stride + (long)slot * N + (long)j;
        if (a.add_to) x = csub(a.add_to[off] + x, q);
        if (a.add_pt) x = csub(x + a.add_pt[off], q);
        o[(long)slot * N] = x;
    }
}
// one digit of a key:  out = [neg] ( [MForm] e  [+ G_digit s_gad]  [-/+ MRed(a, s_mul)] )
//   (GenSwitchingKey keygen.go:269-327, the relinearisation key's b / d / v :161-186, rotation / conjugation keys :212-264,
//   the public key :98-108; mkbfv's GenBFVSwitchingKey mkbfv/keygen.go:91-162).  The gadget constant G_digit mod q_slot is P on
//   the digit's own limbs and 0 elsewhere (pmont), or comes from a table (`gad`, BFV).  grid = (N/256, nslots, ndigits)
struct KeyFmaArgs {
    const u64 *e;            // [digit][dmax][N]: NTT(e_i), not in Montgomery form
    const u64 *a;            // [digit][dmax][N]: CRS operand (NTT, Montgomery) or nullptr
    const u64 *s_mul;        // [dmax][N]: the secret multiplied with a (NTT, Montgomery)
    const u64 *s_gad;        // [dmax][N]: the secret of the gadget term (NTT, Montgomery) or nullptr
    const u64 *gad;          // [digit][dmax]: MForm(G_digit mod q_slot), or nullptr = the P-on-own-limbs gadget of GenSwitchingKey
    u64 *out;
    int mform_e;             // e -> MForm(e) first (everything but the public key)
    int mul_mode;            // 0 none, 1: -= MRed(a, s_mul) (MulCoeffsMontgomeryAndSub), 2: += (..AndAdd)
    int neg;                 // NegLvl at the end: q - x, unreduced like lattigo
    int alpha, nQ, dmax;
    u64 pmont[MKHE_MAX_SLOTS];   // MForm(P mod q_j), Q limbs (MulScalarBigintLvl by P, keygen.go:289)
    int nslots;
    int slots[MKHE_MAX_SLOTS];
    int logN;
};
__global__ void __launch_bounds__(MKHE_THREADS) k_key_fma(KeyFmaArgs a, const ModC *mods) {
    const long N = 1L << a.logN;
    const int slot = a.slots[blockIdx.y], digit = blockIdx.z;
    const ModC m = mods[slot];
    const long j = (long)blockIdx.x * MKHE_THREADS + threadIdx.x;
    const long off = ((long)digit * a.dmax + slot) * N + j;
    u64 x = a.e[off];
    if (a.mform_e) x = mred(x, m.r2, m.q, m.qinv);
    if (a.s_gad) {
        if (a.gad) x = csub(x + mred(a.s_gad[(long)slot * N + j], a.gad[digit * a.dmax + slot], m.q, m.qinv), m.q);
        else if (slot < a.nQ && slot / a.alpha == digit) x = csub(x + mred(a.s_gad[(long)slot * N + j], a.pmont[slot], m.q, m.qinv), m.q);
    }
    if (a.mul_mode) {
        const u64 t = mred(a.a[off], a.s_mul[(long)slot * N + j], m.q, m.qinv);
        x = a.mul_mode == 1 ? csub(x + m.q - t, m.q) : csub(x + t, m.q);
    }
    if (a.neg) x = m.q - x;
    a.out[off] = x;
}
// ring.PermuteNTTWithIndexLvl with ring.PermuteNTTIndex(galEl, N) (keygen.go:212-214,243-245): out[i] = in[index(i)].  grid = (N/256, nlimbs)
__global__ void __launch_bounds__(MKHE_THREADS) k_permute_ntt(const u64 *in, u64 *out, u64 galEl, int nlimbs, int logN) {
    const long N = 1L << logN;
    const u64 i = (u64)blockIdx.x * MKHE_THREADS + threadIdx.x, mask = ((u64)N << 1) - 1;
    auto brev = [&](u64 x) { u64 r = 0; for (int b = 0; b < logN; b++) r |= ((x >> b) & 1) << (logN - 1 - b); return r; };
    const u64 t1 = 2 * brev(i) + 1, t2 = (((galEl * t1) & mask) - 1) >> 1;
    const u64 idx = brev(t2);
    const int limb = blockIdx.y;
    out[(long)limb * N + (long)i] = in[(long)limb * N + (long)idx];
}

// development: checksum of a buffer (race hunting: identical ops must produce identical intermediate buffers)
__global__ void __launch_bounds__(MKHE_THREADS) k_checksum(const u64 *p, size_t n, u64 *out) {
    u64 s = 0;
    for (size_t i = (size_t)blockIdx.x * MKHE_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * MKHE_THREADS) s += ld_cg(p + i) * (u64)(2 * i + 1);
    atomicAdd(out, s);
}

// register-resident butterfly throughput probe (integer-pipe roofline denominator)
__global__ void __launch_bounds__(MKHE_THREADS) k_bfly_peak(u64 *sink, ModC m, int iters) {
    u64 v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = (u64)(threadIdx.x * 8 + k + blockIdx.x) % m.q;
    u64 w = m.ninv, wsh = m.ninv_sh;
    const NttC c = nttc(m);
    for (int it = 0; it < iters; it++) {          // throughput probe only: values may wrap, the instruction stream is what counts
#pragma unroll
        for (int k = 0; k < 4; k++) bf_fwd(v[k], v[k + 4], w, wsh, c);
#pragma unroll
        for (int g = 0; g < 2; g++) { bf_fwd(v[4 * g], v[4 * g + 2], w, wsh, c); bf_fwd(v[4 * g + 1], v[4 * g + 3], w, wsh, c); }
#pragma unroll
        for (int g = 0; g < 4; g++) bf_fwd(v[2 * g], v[2 * g + 1], w, wsh, c);
    }
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= v[k];
    if (s == 0x123456789abcdefull) sink[0] = s;
}
