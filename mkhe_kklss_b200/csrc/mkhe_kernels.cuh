// mkhe_kernels.cuh -- sm_100a kernels of the MKHE-KKLSS hot path.
//
// NTT organisation (N = 2^logN, 12 <= logN <= 16).  lattigo's forward NTT is an in-place Cooley-Tukey
// DIT with in-order input and bit-reversed output; stage s (1-based) pairs positions j and
// j + N/2^s and uses the twiddle psi[2^(s-1) + (j >> (logN-s+1))] (lattigo ring/ntt.go; the tables
// are public fields read at mkrlwe/basis_extension.go:263-264).  The same dataflow is split into
//   pass 1: the first S1 = logN-9 stages.  They act on N2 = 512 independent "columns" (positions with
//           equal j mod 512) and every column uses the same 2^S1-1 twiddles.  A CTA owns a tile of
//           2^S1 rows x W = 2^(11-S1) adjacent columns (2048 elements, rows are W*8-byte contiguous
//           segments of HBM).
//   pass 2: the last 9 stages = 2^S1 independent contiguous blocks of 512 elements, each with its own
//           511 twiddles.  A CTA owns 4 adjacent blocks (2048 contiguous elements, 16 KiB).
// Inside a CTA every thread (256 of them) holds 8 elements in registers and performs radix-8 rounds
// (3 stages, 12 Harvey/Shoup butterflies, values lazily kept in [0,4q)); between rounds the tile is
// re-distributed through XOR-swizzled shared memory.  The inverse transform mirrors this
// (Gentleman-Sande, pass A = contiguous, pass B = columns, N^-1 folded into the last stage).
// Positions are never permuted, so key material in lattigo's ordering is used as uploaded.
#pragma once
#include "mkhe_arith.cuh"

#define MKHE_TILE 2048
#define MKHE_THREADS 256
#define MKHE_MAX_SLOTS 34        // limb slots per launch list (nQ + nP <= 34)

// conflict-free tile addressing: bank bits 0..3 (u64 units) are XORed with index bits 3..6
__device__ __forceinline__ int swz(int idx) { return idx ^ ((idx >> 3) & 15); }

struct PtrList { u64 *p[MKHE_MAX_PARTIES_K]; };

// ------------------------------------------------------------------------------------------------
// twiddle accessors.  Stage "d" = the butterfly distance is 2^d in tile-local index space (0..10);
// lg = local group index = local_idx >> (d+1).
//   pass 1 (columns): table index = 2^(10-d) + lg
//   pass 2 (contiguous): table index = 2^(S1+8-d) + (tile << (10-d)) + lg
// ------------------------------------------------------------------------------------------------
struct TwGlobal {
    const ulonglong2 *tab;
    int shift, tile;
    __device__ __forceinline__ ulonglong2 get(int d, int lg) const {
        return tab[(1u << (10 - d + shift)) + ((u32)tile << (10 - d)) + (u32)lg];
    }
};
// heap layout (levels 2^2 .. 2^10 of one pass-2 tile); entry h lives at h ^ ((h >> 3) & 3) so that the per-thread
// runs of 2 / 4 consecutive 16-byte entries read in the last round spread over all 8 bank groups
__device__ __forceinline__ int twz(int h) { return h ^ ((h >> 3) & 3); }
struct TwShared {
    const ulonglong2 *s;
    __device__ __forceinline__ ulonglong2 get(int d, int lg) const { return s[twz((1 << (10 - d)) + lg)]; }
};
__device__ __forceinline__ void load_tile_twiddles(ulonglong2 *s, const ulonglong2 *tab, int S1, int tile) {
    for (int i = threadIdx.x + 4; i < MKHE_TILE; i += MKHE_THREADS) {
        u32 h = (u32)i;
        int lv = 31 - mkhe_clz(h);
        u32 lg = h - (1u << lv);
        s[twz(i)] = tab[(1u << (lv + S1 - 2)) + ((u32)tile << lv) + lg];
    }
}

// ------------------------------------------------------------------------------------------------
// radix-8 register rounds on the window of local index bits [e, e+3); element k of the thread sits
// at local index (hi << (e+3)) | (k << e) | low.  NS = number of stages performed (the TOP NS bits).
// ------------------------------------------------------------------------------------------------
template <int NS, class TW>
__device__ __forceinline__ void round_fwd(u64 v[8], int e, int hi, const TW &tw, u64 q, u64 twoq) {
    {   // distance 2^(e+2)
        ulonglong2 w = tw.get(e + 2, hi);
#pragma unroll
        for (int k = 0; k < 4; k++) bf_fwd(v[k], v[k + 4], w.x, w.y, q, twoq);
    }
    if (NS >= 2) {  // distance 2^(e+1)
#pragma unroll
        for (int g = 0; g < 2; g++) {
            ulonglong2 w = tw.get(e + 1, 2 * hi + g);
            bf_fwd(v[4 * g + 0], v[4 * g + 2], w.x, w.y, q, twoq);
            bf_fwd(v[4 * g + 1], v[4 * g + 3], w.x, w.y, q, twoq);
        }
    }
    if (NS >= 3) {  // distance 2^e
#pragma unroll
        for (int g = 0; g < 4; g++) {
            ulonglong2 w = tw.get(e, 4 * hi + g);
            bf_fwd(v[2 * g], v[2 * g + 1], w.x, w.y, q, twoq);
        }
    }
}
// inverse: stages in the opposite order (lowest distance first).  If LAST, the top stage (distance
// 2^(e+2), which must be the final stage of the whole transform) also multiplies by N^-1.
template <int NS, bool LAST, class TW>
__device__ __forceinline__ void round_inv(u64 v[8], int e, int hi, const TW &tw, const ModC &m) {
    const u64 q = m.q, twoq = 2 * m.q;
    if (NS >= 3) {
#pragma unroll
        for (int g = 0; g < 4; g++) {
            ulonglong2 w = tw.get(e, 4 * hi + g);
            bf_inv(v[2 * g], v[2 * g + 1], w.x, w.y, q, twoq);
        }
    }
    if (NS >= 2) {
#pragma unroll
        for (int g = 0; g < 2; g++) {
            ulonglong2 w = tw.get(e + 1, 2 * hi + g);
            bf_inv(v[4 * g + 0], v[4 * g + 2], w.x, w.y, q, twoq);
            bf_inv(v[4 * g + 1], v[4 * g + 3], w.x, w.y, q, twoq);
        }
    }
    if (!LAST) {
        ulonglong2 w = tw.get(e + 2, hi);
#pragma unroll
        for (int k = 0; k < 4; k++) bf_inv(v[k], v[k + 4], w.x, w.y, q, twoq);
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            u64 s = v[k] + v[k + 4];
            u64 t = v[k] - v[k + 4] + twoq;
            v[k] = csub(shoup_lazy(s, m.ninv, m.ninv_sh, q), q);
            v[k + 4] = csub(shoup_lazy(t, m.w1ninv, m.w1ninv_sh, q), q);
        }
    }
}

// move the tile from the window-e_from distribution to the window-e_to distribution through smem
__device__ __forceinline__ void exchange(u64 v[8], u64 *sm, int e_from, int e_to) {
    const int tid = threadIdx.x;
    {
        int low = tid & ((1 << e_from) - 1), hi = tid >> e_from;
        int base = (hi << (e_from + 3)) | low;
#pragma unroll
        for (int k = 0; k < 8; k++) sm[swz(base | (k << e_from))] = v[k];
    }
    __syncthreads();
    {
        int low = tid & ((1 << e_to) - 1), hi = tid >> e_to;
        int base = (hi << (e_to + 3)) | low;
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = sm[swz(base | (k << e_to))];
    }
}

// ------------------------------------------------------------------------------------------------
// pass-1 / pass-B tile <-> global index:  local idx = (row << WB) | col,  WB = 11 - S1
// ------------------------------------------------------------------------------------------------
template <int S1>
__device__ __forceinline__ long col_tile_offset(int idx, int tile) {
    constexpr int WB = 11 - S1;
    int row = idx >> WB, col = idx & ((1 << WB) - 1);
    return (long)row * 512 + ((long)tile << WB) + col;
}

// forward column stages on registers (+ smem exchanges).  On entry v is in the window-8 distribution,
// on exit in the window-(11-S1) distribution.
template <int S1>
__device__ __forceinline__ void cols_fwd(u64 v[8], u64 *sm, const TwGlobal &tw, u64 q) {
    constexpr int R = S1 % 3;
    const int tid = threadIdx.x;
    const u64 twoq = 2 * q;
    int e = 8;
    if (R == 1) round_fwd<1>(v, 8, tid >> 8, tw, q, twoq);
    else if (R == 2) round_fwd<2>(v, 8, tid >> 8, tw, q, twoq);
    else round_fwd<3>(v, 8, tid >> 8, tw, q, twoq);
    int done = (R == 0) ? 3 : R;
#pragma unroll
    for (; done < S1; done += 3) {
        int en = 11 - done - 3;
        exchange(v, sm, e, en);
        e = en;
        round_fwd<3>(v, e, tid >> e, tw, q, twoq);
        if (done + 3 < S1) __syncthreads();
    }
}
// inverse column stages: entry distribution window-(11-S1), exit window-8; final stage applies N^-1
// and produces canonical values.
template <int S1>
__device__ __forceinline__ void cols_inv(u64 v[8], u64 *sm, const TwGlobal &tw, const ModC &m) {
    constexpr int R = S1 % 3;
    constexpr int FULL = S1 / 3;               // number of full rounds
    const int tid = threadIdx.x;
    int e = 11 - S1;
    if (R == 0) {
#pragma unroll
        for (int r = 0; r < FULL; r++) {
            if (r == FULL - 1) round_inv<3, true>(v, e, tid >> e, tw, m);
            else {
                round_inv<3, false>(v, e, tid >> e, tw, m);
                exchange(v, sm, e, e + 3);
                __syncthreads();
                e += 3;
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < FULL; r++) {
            round_inv<3, false>(v, e, tid >> e, tw, m);
            int en = (r == FULL - 1) ? 8 : e + 3;
            exchange(v, sm, e, en);
            __syncthreads();
            e = en;
        }
        if (R == 1) round_inv<1, true>(v, 8, tid >> 8, tw, m);
        else round_inv<2, true>(v, 8, tid >> 8, tw, m);
    }
}

// ------------------------------------------------------------------------------------------------
// K1'  decompose + ModUp (alpha = 1: digit broadcast) + first S1 NTT stages
//   replaces Decomposer.DecomposeAndSplit copy branch + the head of ringQ/ringP.NTTLvl
//   (mkrlwe/basis_extension.go:443-451, mkrlwe/keyswitch.go:21-31).
//   grid = (N/2048 column tiles, ndigits, npolys).  The digit limb is read ONCE and reduced into
//   every target limb slot (Barrett, then NTT stages), so HBM sees 1 read and D writes per digit.
// ------------------------------------------------------------------------------------------------
struct BcastArgs {
    PtrList in;            // per poly: coefficient-domain poly
    PtrList out;           // per poly: swk-shaped buffer [digit][Dmax][N]
    int in_limb0;          // first source limb (BFV's second half uses nQ)
    int dmax;              // limb slots per digit in the output
    int nslots;
    int slots[MKHE_MAX_SLOTS];   // modulus index == limb slot of each target limb
    int logN;
};

template <int S1>
__global__ void __launch_bounds__(MKHE_THREADS) k_bcast_ntt_pass1(BcastArgs a, const ModC *mods, const ulonglong2 *twf) {
    MKHE_SMEM(smraw);
    u64 *sm = reinterpret_cast<u64 *>(smraw);
    const int tid = threadIdx.x, tile = blockIdx.x, digit = blockIdx.y, poly = blockIdx.z;
    const long N = 1L << a.logN;
    const u64 *src = a.in.p[poly] + (long)(a.in_limb0 + digit) * N;
    u64 *dst0 = a.out.p[poly] + (long)digit * a.dmax * N;
    u64 raw[8];
#pragma unroll
    for (int k = 0; k < 8; k++) raw[k] = src[col_tile_offset<S1>((k << 8) | tid, tile)];
    constexpr int EL = 11 - S1;
    const int low = tid & ((1 << EL) - 1), hi = tid >> EL;
    for (int s = 0; s < a.nslots; s++) {
        const int mi = a.slots[s];
        const ModC m = mods[mi];
        TwGlobal tw{twf + (long)mi * N, 0, 0};
        u64 v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = barrett_lazy(raw[k], m.q, m.mu);
        cols_fwd<S1>(v, sm, tw, m.q);
        u64 *dst = dst0 + (long)mi * N;
#pragma unroll
        for (int k = 0; k < 8; k++) dst[col_tile_offset<S1>((hi << (EL + 3)) | (k << EL) | low, tile)] = v[k];
        if (S1 > 3) __syncthreads();
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

// K1 pass 1 (plain): grid = (tiles, nlimbs, npolys).  Inputs < 4q (canonical, "q" from rotations, or
// BFV's lazy multSum limbs) are consumed as they are.
template <int S1>
__global__ void __launch_bounds__(MKHE_THREADS) k_ntt_pass1(LimbArgs a, const ModC *mods, const ulonglong2 *twf) {
    MKHE_SMEM(smraw);
    u64 *sm = reinterpret_cast<u64 *>(smraw);
    const int tid = threadIdx.x, tile = blockIdx.x, limb = blockIdx.y, poly = blockIdx.z;
    const long N = 1L << a.logN;
    const int mi = a.mod_of_limb[limb];
    const ModC m = mods[mi];
    const u64 *src = a.in.p[poly] + (long)a.slot_of[limb] * N;
    u64 *dst = a.out.p[poly] + (long)a.slot_of[limb] * N;
    TwGlobal tw{twf + (long)mi * N, 0, 0};
    u64 v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = src[col_tile_offset<S1>((k << 8) | tid, tile)];
    cols_fwd<S1>(v, sm, tw, m.q);
    constexpr int EL = 11 - S1;
    const int low = tid & ((1 << EL) - 1), hi = tid >> EL;
#pragma unroll
    for (int k = 0; k < 8; k++) dst[col_tile_offset<S1>((hi << (EL + 3)) | (k << EL) | low, tile)] = v[k];
}

// ------------------------------------------------------------------------------------------------
// K1 pass 2: last 9 stages on contiguous tiles, canonical output.
//   grid = (tiles, nlimbs(slots), npolys); the CTA loops over `count` instances that share the limb
//   (the digits of a hoisted form) so the tile's 2044 twiddles are staged in shared memory once.
//   data pointer of instance i = base[poly] + i*inst_stride + slot*N + tile*2048   (in place)
// ------------------------------------------------------------------------------------------------
struct Pass2Args {
    PtrList buf;
    int count;
    long inst_stride;
    int nslots;
    int slots[MKHE_MAX_SLOTS];       // limb slot (offset slot*N)
    int mods[MKHE_MAX_SLOTS];        // modulus index of that slot
    int logN;
};

__device__ __forceinline__ void contig_fwd(u64 v[8], u64 *sm, const TwShared &tw, u64 q) {
    const int tid = threadIdx.x;
    const u64 twoq = 2 * q;
    round_fwd<3>(v, 6, tid >> 6, tw, q, twoq);
    exchange(v, sm, 6, 3);
    round_fwd<3>(v, 3, tid >> 3, tw, q, twoq);
    __syncthreads();
    exchange(v, sm, 3, 0);
    round_fwd<3>(v, 0, tid, tw, q, twoq);
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = csub(csub(v[k], twoq), q);
}
template <class TW>
__device__ __forceinline__ void contig_inv(u64 v[8], u64 *sm, const TW &tw, const ModC &m) {
    const int tid = threadIdx.x;
    round_inv<3, false>(v, 0, tid, tw, m);
    exchange(v, sm, 0, 3);
    round_inv<3, false>(v, 3, tid >> 3, tw, m);
    __syncthreads();
    exchange(v, sm, 3, 6);
    round_inv<3, false>(v, 6, tid >> 6, tw, m);
}

__global__ void __launch_bounds__(MKHE_THREADS) k_ntt_pass2(Pass2Args a, const ModC *mods, const ulonglong2 *twf) {
    MKHE_SMEM(smraw);
    u64 *sm = reinterpret_cast<u64 *>(smraw);                            // 16 KiB exchange tile
    ulonglong2 *stw = reinterpret_cast<ulonglong2 *>(smraw + MKHE_TILE * 8);   // 32 KiB twiddles
    const int tid = threadIdx.x, tile = blockIdx.x, poly = blockIdx.z;
    const int slot = a.slots[blockIdx.y], mi = a.mods[blockIdx.y];
    const long N = 1L << a.logN;
    const ModC m = mods[mi];
    load_tile_twiddles(stw, twf + (long)mi * N, a.logN - 9, tile);
    __syncthreads();
    TwShared tw{stw};
    u64 *base = a.buf.p[poly] + (long)slot * N + (long)tile * MKHE_TILE;
    for (int i = 0; i < a.count; i++) {
        u64 *p = base + (long)i * a.inst_stride;
        u64 v[8];
        const int low = tid & 63, hi = tid >> 6;
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = p[(hi << 9) | (k << 6) | low];
        contig_fwd(v, sm, tw, m.q);
        ulonglong2 *o = reinterpret_cast<ulonglong2 *>(p + tid * 8);
#pragma unroll
        for (int k = 0; k < 4; k++) o[k] = make_ulonglong2(v[2 * k], v[2 * k + 1]);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// K2/K3  inverse pass A on contiguous tiles, optionally fed by the key multiply-accumulate
//   SRC_LOAD: v = in[...]                                     (ringQ.InvNTTLvl head)
//   SRC_MAC : v = sum_t sum_{i<beta} key_t[i] (.) h_t[i]      (MulCoeffsMontgomery[AndAdd]Lvl loops,
//             mkrlwe/keyswitch_hoisted.go:24-32; nsets = 2 for mkbfv/keyswitch_hoisted.go:20-30)
//             accumulated in 128 bits, one Montgomery reduction at the end (same canonical value).
//   grid = (tiles, nslots, nbatch).  Output limb `slot` of out[batch] (coefficient-order positions, still
//   needing pass B).
// ------------------------------------------------------------------------------------------------
struct InvAArgs {
    PtrList in;              // SRC_LOAD: per batch input poly
    PtrList key[2];          // SRC_MAC : per batch key   (swk-shaped, Montgomery form)
    PtrList hst[2];          // SRC_MAC : per batch hoisted (swk-shaped)
    PtrList out;             // per batch output, limb slots of stride N
    int nsets, beta;
    long digit_stride;       // dmax * N
    int nslots;
    int slots[MKHE_MAX_SLOTS];
    int mods[MKHE_MAX_SLOTS];
    int out_slots[MKHE_MAX_SLOTS];
    int logN;
};

template <bool SRC_MAC>
__global__ void __launch_bounds__(MKHE_THREADS) k_intt_passA(InvAArgs a, const ModC *mods, const ulonglong2 *twi) {
    MKHE_SMEM(smraw);
    u64 *sm = reinterpret_cast<u64 *>(smraw);
    const int tid = threadIdx.x, tile = blockIdx.x, b = blockIdx.z;
    const int slot = a.slots[blockIdx.y], mi = a.mods[blockIdx.y], oslot = a.out_slots[blockIdx.y];
    const long N = 1L << a.logN;
    const ModC m = mods[mi];
    const long off = (long)slot * N + (long)tile * MKHE_TILE + tid * 8;
    u64 v[8];
    if (SRC_MAC) {
        u64 hi[8], lo[8];
#pragma unroll
        for (int k = 0; k < 8; k++) hi[k] = lo[k] = 0;
        int terms = 0;
        for (int t = 0; t < a.nsets; t++) {
            const u64 *kp = a.key[t].p[b] + off;
            const u64 *hp = a.hst[t].p[b] + off;
            for (int i = 0; i < a.beta; i++) {
                const ulonglong2 *k2 = reinterpret_cast<const ulonglong2 *>(kp + (long)i * a.digit_stride);
                const ulonglong2 *h2 = reinterpret_cast<const ulonglong2 *>(hp + (long)i * a.digit_stride);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    ulonglong2 kk = k2[k], hh = h2[k];
                    mac128(hi[2 * k], lo[2 * k], kk.x, hh.x);
                    mac128(hi[2 * k + 1], lo[2 * k + 1], kk.y, hh.y);
                }
                if ((++terms & 7) == 0) {
#pragma unroll
                    for (int k = 0; k < 8; k++) hi[k] = csub(barrett_lazy(hi[k], m.q, m.mu), m.q);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            u64 h = csub(barrett_lazy(hi[k], m.q, m.mu), m.q);
            v[k] = mont_reduce(h, lo[k], m.q, m.qinv);
        }
    } else {
        const ulonglong2 *p2 = reinterpret_cast<const ulonglong2 *>(a.in.p[b] + off);
#pragma unroll
        for (int k = 0; k < 4; k++) { ulonglong2 x = p2[k]; v[2 * k] = x.x; v[2 * k + 1] = x.y; }
    }
    TwGlobal tw{twi + (long)mi * N, a.logN - 9 - 2, tile};
    contig_inv(v, sm, tw, m);
    u64 *o = a.out.p[b] + (long)oslot * N + (long)tile * MKHE_TILE;
    const int low = tid & 63, hi6 = tid >> 6;
#pragma unroll
    for (int k = 0; k < 8; k++) o[(hi6 << 9) | (k << 6) | low] = v[k];
}

// K2 pass B: remaining S1 inverse stages on columns, N^-1, canonical output (in place capable).
template <int S1>
__global__ void __launch_bounds__(MKHE_THREADS) k_intt_passB(LimbArgs a, const ModC *mods, const ulonglong2 *twi) {
    MKHE_SMEM(smraw);
    u64 *sm = reinterpret_cast<u64 *>(smraw);
    const int tid = threadIdx.x, tile = blockIdx.x, limb = blockIdx.y, poly = blockIdx.z;
    const long N = 1L << a.logN;
    const int mi = a.mod_of_limb[limb];
    const ModC m = mods[mi];
    const u64 *src = a.in.p[poly] + (long)a.slot_of[limb] * N;
    u64 *dst = a.out.p[poly] + (long)a.slot_of[limb] * N;
    TwGlobal tw{twi + (long)mi * N, 0, 0};
    constexpr int EL = 11 - S1;
    const int low = tid & ((1 << EL) - 1), hi = tid >> EL;
    u64 v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = src[col_tile_offset<S1>((hi << (EL + 3)) | (k << EL) | low, tile)];
    cols_inv<S1>(v, sm, tw, m);
#pragma unroll
    for (int k = 0; k < 8; k++) dst[col_tile_offset<S1>((k << 8) | tid, tile)] = v[k];
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
    u64 hi0 = 0, lo0 = 0, hi1 = 0, lo1 = 0;
    for (int t = 0; t < a.nparties; t++) {
        ulonglong2 kk = *reinterpret_cast<const ulonglong2 *>(a.key.p[t] + off);
        ulonglong2 hh = *reinterpret_cast<const ulonglong2 *>(a.hst.p[t] + off);
        mac128(hi0, lo0, kk.x, hh.x);
        mac128(hi1, lo1, kk.y, hh.y);
        if ((t & 7) == 7) {
            hi0 = csub(barrett_lazy(hi0, m.q, m.mu), m.q);
            hi1 = csub(barrett_lazy(hi1, m.q, m.mu), m.q);
        }
    }
    hi0 = csub(barrett_lazy(hi0, m.q, m.mu), m.q);
    hi1 = csub(barrett_lazy(hi1, m.q, m.mu), m.q);
    u64 r0 = mred(mont_reduce(hi0, lo0, m.q, m.qinv), m.r2, m.q, m.qinv);
    u64 r1 = mred(mont_reduce(hi1, lo1, m.q, m.qinv), m.r2, m.q, m.qinv);
    *reinterpret_cast<ulonglong2 *>(a.out + off) = make_ulonglong2(r0, r1);
}

// ------------------------------------------------------------------------------------------------
// K4 / K8  exact RNS basis conversion (modUpExact = reconstructRNS + multSum,
//   mkrlwe/basis_extension.go:337-357,537-646) and the ModDown combine (:203-229).
//   One thread per coefficient.  The fp64 estimate of the overflow count v is reproduced operation by
//   operation (RN convert, RN divide, RN sequential add, truncate).
// ------------------------------------------------------------------------------------------------
#define MKHE_CONV_MAX 16
struct ConvTable {            // device-resident constants of one (source basis -> target basis) pair
    int n1, n2;
    int src_mod[MKHE_CONV_MAX];             // modulus index of source limb i
    int dst_mod[MKHE_CONV_MAX];             // modulus index of target limb j
    u64 qoverqiinvqi[MKHE_CONV_MAX];        // (Q/q_i)^-1 mod q_i, Montgomery
    u64 qoverqimodp[MKHE_CONV_MAX][MKHE_CONV_MAX];   // [j][i] Q/q_i mod p_j, Montgomery
    u64 vtimesqmodp[MKHE_CONV_MAX][MKHE_CONV_MAX + 1];   // [j][v] (-v*Q) mod p_j
    u64 moddown[MKHE_CONV_MAX];             // [j] p_j - MForm(prod src^-1 mod p_j)   (ModDown only)
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
// element-wise kernels
// ------------------------------------------------------------------------------------------------
// K5 tensor products in the NTT domain (mkrlwe/keyswitch_hoisted.go:119-140): all operands canonical;
//   MForm + MulCoeffsMontgomery = plain product mod q.
//   out_0 = A0*B0 ; out_t = [has0] B0*A_t + [has1] A0*B_t
struct TensorArgs {
    const u64 *A0, *B0;           // NTT(op0_0), NTT(op1_0)
    PtrList A, B;                 // per output party: NTT(op0_id) / NTT(op1_id) or nullptr
    PtrList out;                  // [0] = component "0", [1+t]
    int nout;                     // parties in the output
    int nlimbs;
    int mod_of_limb[MKHE_MAX_SLOTS];
    int logN;
};
__global__ void __launch_bounds__(MKHE_THREADS) k_tensor(TensorArgs a, const ModC *mods) {
    const long N = 1L << a.logN;
    const int limb = blockIdx.y;
    const ModC m = mods[a.mod_of_limb[limb]];
    const long off = (long)limb * N + (long)blockIdx.x * MKHE_THREADS + threadIdx.x;
    const u64 a0 = mred(a.A0[off], m.r2, m.q, m.qinv);      // MForm
    const u64 b0 = mred(a.B0[off], m.r2, m.q, m.qinv);
    a.out.p[0][off] = mred(a0, a.B0[off], m.q, m.qinv);
    for (int t = 0; t < a.nout; t++) {
        u64 r = 0;
        if (a.A.p[t]) r = mred(b0, a.A.p[t][off], m.q, m.qinv);
        if (a.B.p[t]) r = csub(r + mred(a0, a.B.p[t][off], m.q, m.qinv), m.q);
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

// register-resident butterfly throughput probe (integer-pipe roofline denominator)
__global__ void __launch_bounds__(MKHE_THREADS) k_bfly_peak(u64 *sink, ModC m, int iters) {
    u64 v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = (u64)(threadIdx.x * 8 + k + blockIdx.x) % m.q;
    u64 w = m.ninv, wsh = m.ninv_sh;
    const u64 q = m.q, twoq = 2 * m.q;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 4; k++) bf_fwd(v[k], v[k + 4], w, wsh, q, twoq);
#pragma unroll
        for (int g = 0; g < 2; g++) { bf_fwd(v[4 * g], v[4 * g + 2], w, wsh, q, twoq); bf_fwd(v[4 * g + 1], v[4 * g + 3], w, wsh, q, twoq); }
#pragma unroll
        for (int g = 0; g < 4; g++) bf_fwd(v[2 * g], v[2 * g + 1], w, wsh, q, twoq);
    }
    u64 s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= v[k];
    if (s == 0x123456789abcdefull) sink[0] = s;
}
