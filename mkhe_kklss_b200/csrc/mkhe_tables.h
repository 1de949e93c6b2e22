// mkhe_tables.h -- host-side generation of per-modulus constants, NTT twiddles (Shoup form) and RNS
// basis-conversion tables.  Follows lattigo v2.3.0 ring.go genNTTParams / primitiveRoot (restated from
// the published algorithm; the module is not vendored in the reference, go.mod:7) and
// mkrlwe/basis_extension.go:34-54,83-153.  A caller that has lattigo can instead pass its own tables
// through mkhe_ctx_set_ntt_tables (include/mkhe.h).
#pragma once
#include <cstdint>
#include <vector>
#include "mkhe_kernels.cuh"

namespace mkhe {

typedef unsigned __int128 u128;

inline u64 h_mulmod(u64 a, u64 b, u64 q) { return (u64)(((u128)a * b) % q); }
inline u64 h_powmod(u64 x, u64 e, u64 q) {
    u64 r = 1 % q;
    x %= q;
    while (e) { if (e & 1) r = h_mulmod(r, x, q); x = h_mulmod(x, x, q); e >>= 1; }
    return r;
}
inline u64 h_invmod(u64 x, u64 q) { return h_powmod(x % q, q - 2, q); }   // q prime
inline u64 h_qinv64(u64 q) { u64 x = 1; for (int i = 0; i < 63; i++) { x *= q; q *= q; } return x; }
inline u64 h_mform(u64 a, u64 q) { return (u64)((((u128)(a % q)) << 64) % q); }
inline u64 h_mred(u64 x, u64 y, u64 q, u64 qinv) {
    u128 m = (u128)x * y;
    u64 mhi = (u64)(m >> 64), mlo = (u64)m;
    u64 hhi = (u64)(((u128)(mlo * qinv) * q) >> 64);
    u64 r = mhi - hhi + q;
    return r >= q ? r - q : r;
}
inline u64 h_shoup(u64 w, u64 q) { return (u64)((((u128)w) << 64) / q); }

inline bool h_is_prime(u64 n) {
    if (n < 2) return false;
    static const u64 sp[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (u64 p : sp) if (n % p == 0) return n == p;
    u64 d = n - 1; int s = 0;
    while (!(d & 1)) { d >>= 1; s++; }
    for (u64 a : sp) {
        u64 x = h_powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int r = 1; r < s; r++) { x = h_mulmod(x, x, n); if (x == n - 1) { comp = false; break; } }
        if (comp) return false;
    }
    return true;
}
inline u64 h_gcd(u64 a, u64 b) { while (b) { u64 t = a % b; a = b; b = t; } return a; }
inline u64 h_rho(u64 n) {
    if (!(n & 1)) return 2;
    for (u64 c = 1;; c++) {
        u64 x = 2, y = 2, d = 1;
        while (d == 1) {
            x = (h_mulmod(x, x, n) + c) % n;
            y = (h_mulmod(y, y, n) + c) % n;
            y = (h_mulmod(y, y, n) + c) % n;
            d = h_gcd(x > y ? x - y : y - x, n);
        }
        if (d != n) return d;
    }
}
inline void h_factor(u64 n, std::vector<u64> &f) {
    if (n == 1) return;
    if (h_is_prime(n)) { for (u64 x : f) if (x == n) return; f.push_back(n); return; }
    u64 d = h_rho(n);
    h_factor(d, f);
    h_factor(n / d, f);
}
// lattigo primitiveRoot: candidates start at 3
inline u64 h_primitive_root(u64 q) {
    std::vector<u64> f;
    h_factor(q - 1, f);
    for (u64 g = 3;; g++) {
        bool ok = true;
        for (u64 p : f) if (h_powmod(g, (q - 1) / p, q) == 1) { ok = false; break; }
        if (ok) return g;
    }
}
inline u64 h_bitrev(u64 x, int bits) { u64 r = 0; for (int i = 0; i < bits; i++) r = (r << 1) | ((x >> i) & 1); return r; }

struct ModTables {
    ModC c;
    std::vector<ulonglong2> twf, twi;     // [N] Shoup pairs, lattigo index order
};

// from plain-form psi powers: psi_plain[idx], psiinv_plain[idx] in lattigo's bit-reversed index order
inline void finish_mod_tables(ModTables &t, int logN, u64 q, const std::vector<u64> &psi, const std::vector<u64> &psiinv, u64 ninv) {
    const size_t N = (size_t)1 << logN;
    t.c.q = q;
    t.c.qinv = h_qinv64(q);
    t.c.mu = (u64)((((u128)1) << 64) / q);
    t.c.r2 = (u64)((((u128)h_mform(1, q)) << 64) % q);
    t.c.ninv = ninv;
    t.c.ninv_sh = h_shoup(ninv, q);
    t.c.w1ninv = h_mulmod(psiinv[1], ninv, q);
    t.c.w1ninv_sh = h_shoup(t.c.w1ninv, q);
    t.c.qd = (double)q;
    t.c.nq = 0 - q;
    t.c.big = q >= ((u64)1 << 57) ? 1u : 0u;
    int bl = 0;
    while (bl < 64 && (q >> bl)) bl++;
    t.c.bshift = (u32)(bl - 2);
    t.c.mu32 = (u32)((((u128)1) << (bl + 30)) / q);
    t.c.pad = 0;
    t.twf.resize(N);
    t.twi.resize(N);
    for (size_t i = 0; i < N; i++) {
        t.twf[i] = make_ulonglong2(psi[i], h_shoup(psi[i], q));
        t.twi[i] = make_ulonglong2(psiinv[i], h_shoup(psiinv[i], q));
    }
}
inline void gen_mod_tables(ModTables &t, int logN, u64 q) {
    const u64 N = (u64)1 << logN;
    u64 g = h_primitive_root(q);
    u64 power = (q - 1) / (2 * N);
    u64 psi1 = h_powmod(g, power, q), psiinv1 = h_powmod(g, (q - 1) - power, q);
    std::vector<u64> psi(N), psiinv(N);
    u64 a = 1, b = 1;
    for (u64 j = 0; j < N; j++) {
        u64 idx = h_bitrev(j, logN);
        psi[idx] = a;
        psiinv[idx] = b;
        a = h_mulmod(a, psi1, q);
        b = h_mulmod(b, psiinv1, q);
    }
    finish_mod_tables(t, logN, q, psi, psiinv, h_invmod(N % q, q));
}
// lattigo tables arrive in Montgomery form
inline void set_mod_tables_from_mont(ModTables &t, int logN, u64 q, const u64 *psiM, const u64 *psiinvM, u64 ninvM) {
    const size_t N = (size_t)1 << logN;
    u64 qinv = h_qinv64(q);
    std::vector<u64> psi(N), psiinv(N);
    for (size_t i = 0; i < N; i++) { psi[i] = h_mred(psiM[i], 1, q, qinv); psiinv[i] = h_mred(psiinvM[i], 1, q, qinv); }
    finish_mod_tables(t, logN, q, psi, psiinv, h_mred(ninvM, 1, q, qinv));
}

// basisextenderparameters (mkrlwe/basis_extension.go:83-153) + genModDownParams (:34-54) for one
// (source basis -> target basis) pair.  moddown[j] = p_j - MForm(prod_i src_i^-1 mod p_j).
inline void gen_conv_table(ConvTable &c, const std::vector<u64> &mod, const std::vector<int> &src, const std::vector<int> &dst) {
    c.n1 = (int)src.size();
    c.n2 = (int)dst.size();
    for (int i = 0; i < c.n1; i++) {
        c.src_mod[i] = src[i];
        u64 qi = mod[src[i]];
        u64 star = 1;
        for (int j = 0; j < c.n1; j++) if (j != i) star = h_mulmod(star, mod[src[j]] % qi, qi);
        c.qoverqiinvqi[i] = h_mform(h_invmod(star, qi), qi);
    }
    for (int j = 0; j < c.n2; j++) {
        c.dst_mod[j] = dst[j];
        u64 pj = mod[dst[j]];
        u64 Qmod = 1;
        for (int i = 0; i < c.n1; i++) {
            u64 s = 1;
            for (int u = 0; u < c.n1; u++) if (u != i) s = h_mulmod(s, mod[src[u]] % pj, pj);
            c.qoverqimodp[j][i] = h_mform(s, pj);
            Qmod = h_mulmod(Qmod, mod[src[i]] % pj, pj);
        }
        u64 v = pj - Qmod;
        c.vtimesqmodp[j][0] = 0;
        for (int i = 1; i <= c.n1; i++) { u64 t = c.vtimesqmodp[j][i - 1] + v; c.vtimesqmodp[j][i] = t >= pj ? t - pj : t; }
        u64 inv = 1;
        for (int i = 0; i < c.n1; i++) inv = h_mulmod(inv, h_invmod(mod[src[i]] % pj, pj), pj);
        c.moddown[j] = pj - h_mform(inv, pj);
        if (c.n1 <= MKHE_LIFT_MAX_SRC) {          // closed form of the key switch's ModDown (k_moddown_Q)
            c.md_sinv[j][0] = inv;
            c.md_sinv[j][1] = h_shoup(inv, pj);
            for (int i = 0; i < c.n1; i++) {
                const u64 ni = pj - h_invmod(mod[src[i]] % pj, pj);
                c.md_nsrcinv[j][i][0] = ni;
                c.md_nsrcinv[j][i][1] = h_shoup(ni, pj);
            }
        }
    }
}

// modUpParams[..][digit][nsrc-2] of NewDecomposer (mkrlwe/basis_extension.go:377-421): basisextenderparameters of the
// digit's limbs Q[src0 .. src0+nsrc) against every modulus of Q u P (indexed by modulus index)
inline void gen_lift_table(LiftTable &t, const std::vector<u64> &mod, int nQP, int src0, int nsrc) {
    memset(&t, 0, sizeof t);
    t.nsrc = nsrc;
    t.src_limb0 = src0;
    for (int i = 0; i < nsrc; i++) {
        u64 qi = mod[src0 + i];
        u64 star = 1;
        for (int j = 0; j < nsrc; j++) if (j != i) star = h_mulmod(star, mod[src0 + j] % qi, qi);
        t.qoverqiinvqi[i] = h_mform(h_invmod(star, qi), qi);
    }
    for (int j = 0; j < nQP; j++) {
        u64 pj = mod[j];
        u64 Qmod = 1;
        for (int i = 0; i < nsrc; i++) {
            u64 s = 1;
            for (int u = 0; u < nsrc; u++) if (u != i) s = h_mulmod(s, mod[src0 + u] % pj, pj);
            t.qoverqimodp[j][i] = h_mform(s, pj);
            Qmod = h_mulmod(Qmod, mod[src0 + i] % pj, pj);
        }
        u64 v = pj - Qmod;
        t.vtimesqmodp[j][0] = 0;
        for (int i = 1; i <= nsrc; i++) { u64 w = t.vtimesqmodp[j][i - 1] + v; t.vtimesqmodp[j][i] = w >= pj ? w - pj : w; }
    }
}

}  // namespace mkhe
