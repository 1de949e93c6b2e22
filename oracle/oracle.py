"""CPU ORACLE -- Python face of oracle/libmkhe_oracle.so.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this module.  The product package (mkhe_kklss_b200/) never does.

PARITY UNPINNED: see the header of oracle/mkhe_oracle.h.

Layout of this file mirrors the reference's packages:
  Ring / BasisExtender      lattigo ring + mkrlwe/basis_extension.go
  MKParams                  mkrlwe/params.go          (CRS, alpha/beta/gamma)
  KeyGenerator              mkrlwe/keygen.go, mkbfv/keygen.go
  Encryptor / Decryptor     mkrlwe/encryptor.go, mkrlwe/decryptor.go
  KeySwitcher               mkrlwe/keyswitch.go, keyswitch_hoisted.go
  CKKSEvaluator             mkckks/evaluator.go       (MulRelinNew, Rescale, HoistedForm, Rotate*)
  BFVEvaluator              mkbfv/evaluator.go        (MulRelinNew)
Ciphertexts are `Ciphertext(value: dict id -> uint64[nlimbs, N], scale)`; component "0" uses the key "0"
exactly like the reference (mkrlwe/elements.go:17-33); party ids are ints here.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from functools import reduce

import math

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmkhe_oracle.so")

u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)
intp = C.POINTER(C.c_int)
ptrp = C.POINTER(C.c_void_p)


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc); building the checker is not using it."""
    src = os.path.join(_HERE, "mkhe_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "mkhe_oracle.h")),
            os.path.getmtime(os.path.join(_HERE, "..", "include", "mkhe_prng.h"))):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()                      # no-op when the library is newer than its sources
        _lib = C.CDLL(_LIB_PATH)
        _lib.ork_ring_new.restype = C.c_void_p
        _lib.ork_be_new.restype = C.c_void_p
        _lib.ork_ks_new.restype = C.c_void_p
        _lib.ork_bfv_new.restype = C.c_void_p
        _lib.ork_primitive_root.restype = C.c_uint64
        _lib.ork_primitive_root.argtypes = [C.c_uint64]
        _lib.ork_galois_element_for_rotation.restype = C.c_uint64
        _lib.ork_modexp.restype = C.c_uint64
        _lib.ork_modexp.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        _lib.ork_prng_next.restype = C.c_uint64
        _lib.ork_set_threads.restype = C.c_int
    return _lib


def set_threads(n: int) -> int:
    return lib().ork_set_threads(C.c_int(n))


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(u64p)


def _vp(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def _u64arr(xs):
    return (C.c_uint64 * len(xs))(*[int(x) for x in xs])


def _intarr(xs):
    return (C.c_int * len(xs))(*[int(x) for x in xs])


def _ptrarr(arrs):
    """array of pointers to numpy buffers (entries may be None)"""
    return (C.c_void_p * len(arrs))(*[None if a is None else a.ctypes.data for a in arrs])


# --------------------------------------------------------------------------------------------
# ring
# --------------------------------------------------------------------------------------------
class Ring:
    """lattigo ring.Ring restated (tables generated as lattigo's genNTTParams does)."""

    def __init__(self, logN: int, moduli):
        self.logN, self.N = logN, 1 << logN
        self.moduli = [int(q) for q in moduli]
        self.nmod = len(self.moduli)
        self.ptr = C.c_void_p(lib().ork_ring_new(C.c_int(logN), _u64arr(self.moduli), C.c_int(self.nmod)))

    def new_poly(self, nlimbs=None):
        return np.zeros((self.nmod if nlimbs is None else nlimbs, self.N), dtype=np.uint64)

    def _call(self, fn, level, *arrs):
        getattr(lib(), fn)(self.ptr, C.c_int(level), *[_p(a) for a in arrs])

    def ntt(self, a, level=None, out=None):
        level = a.shape[0] - 1 if level is None else level
        out = np.zeros_like(a) if out is None else out
        self._call("ork_ntt_lvl", level, a, out)
        return out

    def intt(self, a, level=None, out=None, lazy=False):
        level = a.shape[0] - 1 if level is None else level
        out = np.zeros_like(a) if out is None else out
        self._call("ork_intt_lazy_lvl" if lazy else "ork_intt_lvl", level, a, out)
        return out

    def mform(self, a, level=None, out=None):
        level = a.shape[0] - 1 if level is None else level
        out = np.zeros_like(a) if out is None else out
        self._call("ork_mform_lvl", level, a, out)
        return out

    def invmform(self, a, level=None, out=None):
        level = a.shape[0] - 1 if level is None else level
        out = np.zeros_like(a) if out is None else out
        self._call("ork_invmform_lvl", level, a, out)
        return out

    def mul_mont(self, a, b, level=None, out=None):
        level = a.shape[0] - 1 if level is None else level
        out = np.zeros_like(a) if out is None else out
        self._call("ork_mul_mont_lvl", level, a, b, out)
        return out

    def mul_mont_add(self, a, b, out, level=None):
        level = a.shape[0] - 1 if level is None else level
        self._call("ork_mul_mont_add_lvl", level, a, b, out)
        return out

    def mul_mont_sub(self, a, b, out, level=None):
        level = a.shape[0] - 1 if level is None else level
        self._call("ork_mul_mont_sub_lvl", level, a, b, out)
        return out

    def add(self, a, b, level=None, out=None):
        level = a.shape[0] - 1 if level is None else level
        out = np.zeros_like(a) if out is None else out
        self._call("ork_add_lvl", level, a, b, out)
        return out

    def sub(self, a, b, level=None, out=None):
        level = a.shape[0] - 1 if level is None else level
        out = np.zeros_like(a) if out is None else out
        self._call("ork_sub_lvl", level, a, b, out)
        return out

    def neg(self, a, level=None, out=None):
        level = a.shape[0] - 1 if level is None else level
        out = np.zeros_like(a) if out is None else out
        self._call("ork_neg_lvl", level, a, out)
        return out

    def reduce(self, a, level=None, out=None):
        level = a.shape[0] - 1 if level is None else level
        out = np.zeros_like(a) if out is None else out
        self._call("ork_reduce_lvl", level, a, out)
        return out

    def mul_residues(self, a, residues, level=None, out=None):
        level = a.shape[0] - 1 if level is None else level
        out = np.zeros_like(a) if out is None else out
        lib().ork_mul_residues_lvl(self.ptr, C.c_int(level), _p(a), _u64arr(residues), _p(out))
        return out

    def mul_bigint(self, a, k: int, level=None, out=None):
        """MulScalarBigint: multiply by the integer k (reduced per limb here, in Python ints)."""
        return self.mul_residues(a, [k % q for q in self.moduli], level, out)

    def permute(self, a, galEl, level=None):
        level = a.shape[0] - 1 if level is None else level
        out = np.zeros_like(a)
        lib().ork_permute(self.ptr, C.c_int(level), _p(a), C.c_uint64(galEl), _p(out))
        return out

    def permute_ntt(self, a, galEl, level=None):
        level = a.shape[0] - 1 if level is None else level
        out = np.zeros_like(a)
        lib().ork_permute_ntt(self.ptr, C.c_int(level), _p(a), C.c_uint64(galEl), _p(out))
        return out

    def div_round_by_last_modulus_many(self, level, nb, p0, p1):
        lib().ork_div_round_by_last_modulus_many(self.ptr, C.c_int(level), C.c_int(nb), _p(p0), _p(p1))

    def lift_small(self, small: np.ndarray, level=None):
        level = self.nmod - 1 if level is None else level
        out = np.zeros((level + 1, self.N), dtype=np.uint64)
        lib().ork_lift_small(self.ptr, C.c_int(level), small.ctypes.data_as(i64p), _p(out))
        return out

    def tables(self, i):
        """(NttPsi, NttPsiInv, NttNInv) of modulus i, Montgomery form, as lattigo exposes them."""
        N = self.N

        class _R(C.Structure):
            _fields_ = [("logN", C.c_int), ("N", C.c_int), ("nmod", C.c_int), ("q", u64p), ("qinv", u64p),
                        ("bred", u64p), ("ninv", u64p), ("psi", u64p), ("psiinv", u64p), ("rescale", u64p)]
        r = C.cast(self.ptr, C.POINTER(_R)).contents
        psi = np.ctypeslib.as_array(r.psi, shape=(self.nmod * N,))[i * N:(i + 1) * N].copy()
        psiinv = np.ctypeslib.as_array(r.psiinv, shape=(self.nmod * N,))[i * N:(i + 1) * N].copy()
        return psi, psiinv, int(r.ninv[i])


def galois_element_for_rotation(logN: int, k: int) -> int:
    return int(lib().ork_galois_element_for_rotation(C.c_int(logN), C.c_int(k)))


class PRNG:
    """xoshiro256** seeded with splitmix64 (seed rule: SURVEY 8d)."""

    def __init__(self, seed: int):
        self.state = (C.c_uint64 * 4)()
        lib().ork_prng_seed(self.state, C.c_uint64(seed))

    def uniform(self, ring: Ring, level=None):
        level = ring.nmod - 1 if level is None else level
        out = np.zeros((level + 1, ring.N), dtype=np.uint64)
        lib().ork_sample_uniform(self.state, ring.ptr, C.c_int(level), _p(out))
        return out

    def ternary(self, N, pzero=0.5):
        out = np.zeros(N, dtype=np.int64)
        lib().ork_sample_ternary(self.state, C.c_int(N), C.c_double(pzero), out.ctypes.data_as(i64p))
        return out

    def gaussian(self, N, sigma=3.2, bound=19):
        out = np.zeros(N, dtype=np.int64)
        lib().ork_sample_gaussian(self.state, C.c_int(N), C.c_double(sigma), C.c_int(bound), out.ctypes.data_as(i64p))
        return out


class CtrPRNG:
    """the counter-based samplers of include/mkhe_prng.h behind the PRNG interface: every sampled polynomial takes the next
    stream id(s) (a uniform poly: one per limb), in call order -- the order the device-side generators document
    (include/mkhe.h, mkhe_keygen_*).  Swapping it in for PRNG makes KeyGenerator / Encryptor / MKParams reproduce what the GPU
    makes from the same (seed, first stream)."""

    def __init__(self, seed: int, stream: int = 0):
        self.seed, self.stream = int(seed) & (2**64 - 1), int(stream)

    def uniform(self, ring: Ring, level=None):
        level = ring.nmod - 1 if level is None else level
        out = np.zeros((level + 1, ring.N), dtype=np.uint64)
        lib().ork_ctr_uniform(C.c_uint64(self.seed), C.c_uint64(self.stream), ring.ptr, C.c_int(level), _p(out))
        self.stream += level + 1
        return out

    def ternary(self, N, pzero=0.5):
        out = np.zeros(N, dtype=np.int64)
        lib().ork_ctr_ternary(C.c_uint64(self.seed), C.c_uint64(self.stream), C.c_int(N), C.c_uint64(int(pzero * 2**53)), out.ctypes.data_as(i64p))
        self.stream += 1
        return out

    def gaussian(self, N, sigma=3.2, bound=19):
        assert sigma == 3.2 and bound == 19, "mkhe-ctr-1 fixes sigma = 3.2, bound = 19 (rlwe.DefaultSigma)"
        out = np.zeros(N, dtype=np.int64)
        lib().ork_ctr_gaussian(C.c_uint64(self.seed), C.c_uint64(self.stream), C.c_int(N), out.ctypes.data_as(i64p))
        self.stream += 1
        return out


class BasisExtender:
    """mkrlwe.FastBasisExtender (basis_extension.go:12-357)."""

    def __init__(self, ringQ: Ring, ringP: Ring):
        self.ringQ, self.ringP = ringQ, ringP
        self.ptr = C.c_void_p(lib().ork_be_new(ringQ.ptr, ringP.ptr))

    def modup_q_to_p(self, levelQ, levelP, polQ):
        out = np.zeros((levelP + 1, self.ringQ.N), dtype=np.uint64)
        lib().ork_be_modup_q_to_p(self.ptr, C.c_int(levelQ), C.c_int(levelP), _p(polQ), _p(out))
        return out

    def modup_p_to_q(self, levelP, levelQ, polP):
        out = np.zeros((levelQ + 1, self.ringQ.N), dtype=np.uint64)
        lib().ork_be_modup_p_to_q(self.ptr, C.c_int(levelP), C.c_int(levelQ), _p(polP), _p(out))
        return out

    def moddown_qp_to_q(self, levelQ, levelP, p1Q, p1P):
        out = np.zeros((levelQ + 1, self.ringQ.N), dtype=np.uint64)
        lib().ork_be_moddown_qp_to_q(self.ptr, C.c_int(levelQ), C.c_int(levelP), _p(p1Q), _p(p1P), _p(out))
        return out

    def moddown_qp_to_q_ntt(self, levelQ, levelP, p1Q, p1P):
        """ModDownQPtoQNTT (basis_extension.go:239-290); p1P is left untouched here (the reference overwrites it with its lazy INTT)"""
        out = np.zeros((levelQ + 1, self.ringQ.N), dtype=np.uint64)
        tmp = np.ascontiguousarray(p1P).copy()
        lib().ork_be_moddown_qp_to_q_ntt(self.ptr, C.c_int(levelQ), C.c_int(levelP), _p(np.ascontiguousarray(p1Q)), _p(tmp), _p(out))
        return out

    def moddown_qp_to_p(self, levelQ, levelP, p1Q, p1P):
        out = np.zeros((levelP + 1, self.ringQ.N), dtype=np.uint64)
        lib().ork_be_moddown_qp_to_p(self.ptr, C.c_int(levelQ), C.c_int(levelP), _p(p1Q), _p(p1P), _p(out))
        return out


# --------------------------------------------------------------------------------------------
# containers (mkrlwe/elements.go, keys.go)
# --------------------------------------------------------------------------------------------
@dataclass
class Ciphertext:
    value: dict                      # "0" -> poly, id(int) -> poly ; poly = uint64[nlimbs, N]
    scale: float = 0.0

    def ids(self):
        return sorted(k for k in self.value if k != "0")

    def level(self):
        return self.value["0"].shape[0] - 1

    def copy(self):
        return Ciphertext({k: v.copy() for k, v in self.value.items()}, self.scale)


@dataclass
class SecretKey:
    id: int
    Q: np.ndarray                    # NTT + Montgomery
    P: np.ndarray
    small: np.ndarray = None         # the signed coefficient vector (oracle convenience)


@dataclass
class RelinKey:                      # mkrlwe.RelinearizationKey (keys.go:34-37): Value[0..2] = b, d, v
    id: int
    b: np.ndarray
    d: np.ndarray
    v: np.ndarray


@dataclass
class BFVRelinKey:                   # mkbfv.RelinearizationKey (mkbfv/keys.go:6-9)
    id: int
    b1: np.ndarray
    d1: np.ndarray
    v: np.ndarray
    b2: np.ndarray
    d2: np.ndarray


class MKParams:
    """mkrlwe.Parameters (params.go:8-99): ring Q, ring P, gamma, CRS."""

    CRS_DEFAULT = (0, -1, -2, -3, -4)

    def __init__(self, logN, Q, P, gamma=2, seed=0xB2000000, crs_rots=None, sigma=3.2):
        self.logN, self.N = logN, 1 << logN
        self.Q, self.P = [int(x) for x in Q], [int(x) for x in P]
        self.gamma, self.sigma = gamma, sigma
        self.ringQ, self.ringP = Ring(logN, self.Q), Ring(logN, self.P)
        self.nQ, self.nP = len(self.Q), len(self.P)
        self.D = self.nQ + self.nP
        self.seed = seed
        self.CRS = {}
        idxs = list(self.CRS_DEFAULT)
        if crs_rots is None:
            crs_rots = [1 << i for i in range(logN - 1)]       # params.go:44-46
        for idx in idxs + list(crs_rots):
            self.add_crs(idx)

    def alpha(self):
        return self.nP // self.gamma                              # params.go:63-65

    def beta(self, levelQ):
        a = self.alpha()
        return -(-(levelQ + 1) // a)                              # params.go:67-71

    def max_level(self):
        return self.nQ - 1

    def new_swk(self):
        return np.zeros((self.beta(self.max_level()), self.D, self.N), dtype=np.uint64)

    def add_crs(self, idx, prng=None):
        """uniform QP polys, then MForm (params.go:49-58, 77-99); one PRNG stream per index."""
        if idx in self.CRS and prng is None:
            return
        prng = prng if prng is not None else PRNG(self.seed ^ (0xC125 << 20) ^ (idx & 0xFFFFF))
        swk = self.new_swk()
        for i in range(swk.shape[0]):
            swk[i, :self.nQ] = self.ringQ.mform(prng.uniform(self.ringQ))
            swk[i, self.nQ:] = self.ringP.mform(prng.uniform(self.ringP))
        self.CRS[idx] = swk

    def galois_element_for_rotation(self, k):
        return galois_element_for_rotation(self.logN, k)


# --------------------------------------------------------------------------------------------
# key generation (mkrlwe/keygen.go)
# --------------------------------------------------------------------------------------------
class KeyGenerator:
    def __init__(self, params: MKParams, seed=1, prng=None):
        self.params = params
        self.prng = prng if prng is not None else PRNG(params.seed + 0x1000 + seed)     # prng=CtrPRNG(..): the device's streams

    # QP helpers ---------------------------------------------------------------------------
    def _qp_ntt_mform(self, small):
        p = self.params
        q = p.ringQ.mform(p.ringQ.ntt(p.ringQ.lift_small(small)))
        pp = p.ringP.mform(p.ringP.ntt(p.ringP.lift_small(small)))
        return q, pp

    def _gaussian_error_ntt(self):
        """GenGaussianError keygen.go:123-133 (NTT, not Montgomery) as one [D, N] array"""
        p = self.params
        e = self.prng.gaussian(p.N, p.sigma, int(6 * p.sigma))
        return np.concatenate([p.ringQ.ntt(p.ringQ.lift_small(e)), p.ringP.ntt(p.ringP.lift_small(e))])

    def _qp(self, fn, a, *rest, **kw):
        """apply a Ring method to the Q and P halves of [D, N] arrays"""
        p = self.params
        outs = []
        for ring, sl in ((p.ringQ, slice(0, p.nQ)), (p.ringP, slice(p.nQ, p.D))):
            args = [x[sl] if isinstance(x, np.ndarray) and x.ndim == 2 and x.shape[0] == p.D else x for x in (a,) + rest]
            outs.append(getattr(ring, fn)(*[np.ascontiguousarray(x) if isinstance(x, np.ndarray) else x for x in args], **kw))
        return np.concatenate(outs)

    # keys ---------------------------------------------------------------------------------
    def gen_secret_key(self, id):
        """GenSecretKey keygen.go:58-60 -> ternary P(0)=1/2, NTT + MForm (keygen.go:44-55)"""
        small = self.prng.ternary(self.params.N, 0.5)
        q, pp = self._qp_ntt_mform(small)
        return SecretKey(id, q, pp, small)

    def sk_qp(self, sk):
        return np.concatenate([sk.Q, sk.P])

    def gen_public_key(self, sk):
        """GenPublicKey keygen.go:88-109: pk0 = -a*s + e (NTT, not Montgomery), pk1 = a = CRS[0][0]"""
        p = self.params
        pk0 = self._gaussian_error_ntt()
        pk1 = p.CRS[0][0].copy()
        s = self.sk_qp(sk)
        # MulCoeffsMontgomeryAndSub(sk, pk1, pk0)
        outQ = np.ascontiguousarray(pk0[:p.nQ]); outP = np.ascontiguousarray(pk0[p.nQ:])
        p.ringQ.mul_mont_sub(np.ascontiguousarray(s[:p.nQ]), np.ascontiguousarray(pk1[:p.nQ]), outQ)
        p.ringP.mul_mont_sub(np.ascontiguousarray(s[p.nQ:]), np.ascontiguousarray(pk1[p.nQ:]), outP)
        return (np.concatenate([outQ, outP]), pk1)

    def gen_switching_key(self, sk):
        """GenSwitchingKey keygen.go:269-327: swk[i] = MForm(NTT(e_i)) ; limbs [i*alpha, (i+1)*alpha) += P*s"""
        p = self.params
        Pbig = reduce(lambda a, b: a * b, p.P, 1)
        Ps = p.ringQ.mul_bigint(sk.Q, Pbig)                     # :289
        swk = p.new_swk()
        for i in range(swk.shape[0]):
            e = self.prng.gaussian(p.N, p.sigma, int(6 * p.sigma))
            swk[i, :p.nQ] = p.ringQ.mform(p.ringQ.ntt(p.ringQ.lift_small(e)))
            swk[i, p.nQ:] = p.ringP.mform(p.ringP.ntt(p.ringP.lift_small(e)))
            for j in range(p.alpha()):                          # :307-323 (limbs of digit i; partial last digit)
                idx = i * p.alpha() + j
                if idx >= p.nQ:
                    break
                q = np.uint64(p.Q[idx])
                t = swk[i, idx] + Ps[idx]
                swk[i, idx] = np.where(t >= q, t - q, t)       # CRed :320-322
        return swk

    def _mul_sub_digits(self, a, s, out):
        """out[i] -= a[i] * s  (MulCoeffsMontgomeryAndSubLvl over QP, every digit)"""
        p = self.params
        for i in range(out.shape[0]):
            oq = np.ascontiguousarray(out[i, :p.nQ]); op = np.ascontiguousarray(out[i, p.nQ:])
            p.ringQ.mul_mont_sub(np.ascontiguousarray(a[i, :p.nQ]), s.Q, oq)
            p.ringP.mul_mont_sub(np.ascontiguousarray(a[i, p.nQ:]), s.P, op)
            out[i, :p.nQ] = oq; out[i, p.nQ:] = op

    def _gen_b_digit(self, ai, sk):
        """one digit of b = -s*a + e in MForm (keygen.go:161-167)"""
        p = self.params
        t = np.concatenate([p.ringQ.mul_mont(np.ascontiguousarray(ai[:p.nQ]), sk.Q),
                            p.ringP.mul_mont(np.ascontiguousarray(ai[p.nQ:]), sk.P)])
        t = np.concatenate([p.ringQ.invmform(np.ascontiguousarray(t[:p.nQ])), p.ringP.invmform(np.ascontiguousarray(t[p.nQ:]))])
        e = self._gaussian_error_ntt()
        t = np.concatenate([p.ringQ.sub(np.ascontiguousarray(e[:p.nQ]), np.ascontiguousarray(t[:p.nQ])),
                            p.ringP.sub(np.ascontiguousarray(e[p.nQ:]), np.ascontiguousarray(t[p.nQ:]))])
        return np.concatenate([p.ringQ.mform(np.ascontiguousarray(t[:p.nQ])), p.ringP.mform(np.ascontiguousarray(t[p.nQ:]))])

    def _gen_b(self, a, sk):
        b = self.params.new_swk()
        for i in range(b.shape[0]):
            b[i] = self._gen_b_digit(a[i], sk)
        return b

    def _gen_v(self, sk, r):
        """v = -s*u - r*g + e (keygen.go:178-186)"""
        p = self.params
        u = p.CRS[-1]
        v = self.gen_switching_key(r)
        for i in range(v.shape[0]):
            vq = np.ascontiguousarray(v[i, :p.nQ]); vp = np.ascontiguousarray(v[i, p.nQ:])
            p.ringQ.mul_mont_add(np.ascontiguousarray(u[i, :p.nQ]), sk.Q, vq)
            p.ringP.mul_mont_add(np.ascontiguousarray(u[i, p.nQ:]), sk.P, vp)
            v[i, :p.nQ] = p.ringQ.neg(vq); v[i, p.nQ:] = p.ringP.neg(vp)
        # lattigo Neg stores q - x (x == 0 -> q); keys are only ever multiplied, value mod q is what matters,
        # but keep the limbs canonical so that uploads are canonical:
        for j in range(p.D):
            q = np.uint64((p.Q + p.P)[j])
            v[:, j] = np.where(v[:, j] == q, np.uint64(0), v[:, j])
        return v

    def gen_relin_key(self, sk, r):
        """GenRelinearizationKey keygen.go:137-187"""
        p = self.params
        a = p.CRS[0]
        b = self._gen_b(a, sk)
        d = self.gen_switching_key(sk)
        self._mul_sub_digits(a, r, d)                           # d = -r*a + s*g + e
        v = self._gen_v(sk, r)
        return RelinKey(sk.id, b, d, v)

    def gen_rotation_key(self, rotidx, sk):
        """GenRotationKey keygen.go:190-229: rk = -sigma^-1(s)*a_rot + s*g + e"""
        p = self.params
        assert rotidx in p.CRS, "Cannot GenRotationKey: CRS for given rot idx is not generated"
        while rotidx < 0:
            rotidx += p.N // 2
        galEl = p.galois_element_for_rotation(rotidx)
        galInv = pow(galEl, -1, 2 * p.N)
        skOut = SecretKey(sk.id, p.ringQ.permute_ntt(sk.Q, galInv), p.ringP.permute_ntt(sk.P, galInv))
        rk = self.gen_switching_key(sk)
        self._mul_sub_digits(p.CRS[rotidx], skOut, rk)
        return rk

    def gen_conjugation_key(self, sk):
        """GenConjugationKey keygen.go:240-266: cj = -s*a_cj + sigma_c(s)*g + e"""
        p = self.params
        galEl = 2 * p.N - 1
        skOut = SecretKey(sk.id, p.ringQ.permute_ntt(sk.Q, galEl), p.ringP.permute_ntt(sk.P, galEl))
        cj = self.gen_switching_key(skOut)
        self._mul_sub_digits(p.CRS[-2], sk, cj)
        return cj


class Encryptor:
    """mkrlwe.Encryptor.Encrypt, coefficient-domain ciphertext branch (encryptor.go:55-118)."""

    def __init__(self, params: MKParams, seed=7, prng=None):
        self.params = params
        self.prng = prng if prng is not None else PRNG(params.seed + 0x2000 + seed)

    def encrypt(self, pt: np.ndarray, pk, id, scale=0.0):
        p = self.params
        rq = p.ringQ
        level = pt.shape[0] - 1
        w = rq.mform(rq.ntt(rq.lift_small(self.prng.ternary(p.N, 0.5), level)))
        c0 = rq.intt(rq.mul_mont(w, np.ascontiguousarray(pk[0][:level + 1])))
        c1 = rq.intt(rq.mul_mont(w, np.ascontiguousarray(pk[1][:level + 1])))
        c0 = rq.add(c0, rq.lift_small(self.prng.gaussian(p.N, p.sigma, int(6 * p.sigma)), level))
        c1 = rq.add(c1, rq.lift_small(self.prng.gaussian(p.N, p.sigma, int(6 * p.sigma)), level))
        c0 = rq.add(c0, pt)
        return Ciphertext({"0": c0, id: c1}, scale)


class Decryptor:
    """mkrlwe.Decryptor (decryptor.go:26-66); returns the coefficient-domain plaintext poly."""

    def __init__(self, params: MKParams):
        self.params = params

    def decrypt(self, ct: Ciphertext, sks: dict):
        rq = self.params.ringQ
        level = ct.level()
        acc = ct.value["0"].copy()
        for id in ct.ids():
            sk = sks[id]
            t = rq.ntt(ct.value[id], level)
            t = rq.mul_mont(t, np.ascontiguousarray(sk.Q[:level + 1]), level)
            t = rq.intt(t, level)
            acc = rq.add(acc, t, level)
        return rq.reduce(acc, level)


# --------------------------------------------------------------------------------------------
# KeySwitcher (mkrlwe/keyswitch.go, keyswitch_hoisted.go)
# --------------------------------------------------------------------------------------------
class KeySwitcher:
    def __init__(self, params: MKParams):
        self.params = params
        self.ptr = C.c_void_p(lib().ork_ks_new(params.ringQ.ptr, params.ringP.ptr, C.c_int(params.gamma)))

    def decompose(self, levelQ, a, ad=None):
        ad = self.params.new_swk() if ad is None else ad
        lib().ork_ks_decompose(self.ptr, C.c_int(levelQ), _p(a), _p(ad))
        return ad

    def external_product(self, levelQ, a, bg):
        c = np.zeros((levelQ + 1, self.params.N), dtype=np.uint64)
        lib().ork_ks_external_product(self.ptr, C.c_int(levelQ), _p(a), _p(bg), _p(c))
        return c

    def external_product_hoisted(self, levelQ, ah, bg):
        c = np.zeros((levelQ + 1, self.params.N), dtype=np.uint64)
        lib().ork_ks_external_product_hoisted(self.ptr, C.c_int(levelQ), _p(ah), _p(bg), _p(c))
        return c

    @staticmethod
    def _ct_arrays(ct):
        ids = ct.ids()
        return ids, [ct.value["0"]] + [ct.value[i] for i in ids]

    @staticmethod
    def _by_id(d, maxid):
        return _ptrarr([d.get(i) for i in range(maxid + 1)])

    def mul_and_relin_hoisted(self, op0, op1, h0, h1, rlks: dict, ctOut: Ciphertext):
        """MulAndRelinHoisted keyswitch_hoisted.go:44-179; h0/h1 = dict id->swk or None (nil)"""
        level = ctOut.level()
        if op0.level() < level:
            raise ValueError("Cannot MulAndRelin: op0 and op1 have different levels")
        ids0, a0 = self._ct_arrays(op0)
        ids1, a1 = self._ct_arrays(op1)
        idsO, aO = self._ct_arrays(ctOut)
        maxid = max(list(rlks) + [0])
        lib().ork_ks_mul_and_relin_hoisted(
            self.ptr, C.c_int(level),
            C.c_int(len(ids0)), _intarr(ids0), _ptrarr(a0), None if h0 is None else _ptrarr([h0[i] for i in ids0]),
            C.c_int(len(ids1)), _intarr(ids1), _ptrarr(a1), None if h1 is None else _ptrarr([h1[i] for i in ids1]),
            self._by_id({i: k.b for i, k in rlks.items()}, maxid),
            self._by_id({i: k.d for i, k in rlks.items()}, maxid),
            self._by_id({i: k.v for i, k in rlks.items()}, maxid),
            _p(self.params.CRS[-1]),
            C.c_int(len(idsO)), _intarr(idsO), _ptrarr(aO))

    def mul_and_relin(self, op0, op1, rlks, ctOut):
        self.mul_and_relin_hoisted(op0, op1, None, None, rlks, ctOut)

    def rotate_hoisted(self, ctIn, rotidx, hoisted, rks: dict, ctOut):
        """RotateHoisted keyswitch_hoisted.go:183-247; rks: id -> {rot -> swk}"""
        level = ctOut.level()
        if ctIn.level() < level:
            raise ValueError("Cannot Rotate: ctIn and ctOut have different levels")
        ids, a = self._ct_arrays(ctIn)
        _, o = self._ct_arrays(ctOut)
        r = rotidx
        while r < 0:
            r += self.params.N // 2
        lib().ork_ks_rotate_hoisted(self.ptr, C.c_int(level), C.c_int(rotidx), C.c_int(len(ids)), _intarr(ids),
                                    _ptrarr(a), _ptrarr([hoisted[i] for i in ids]),
                                    _ptrarr([rks[i][r] for i in ids]), _p(self.params.CRS[r]), _ptrarr(o))

    def rotate(self, ctIn, rotidx, rks, ctOut):
        level = ctOut.level()
        ids, a = self._ct_arrays(ctIn)
        _, o = self._ct_arrays(ctOut)
        r = rotidx
        while r < 0:
            r += self.params.N // 2
        lib().ork_ks_rotate(self.ptr, C.c_int(level), C.c_int(rotidx), C.c_int(len(ids)), _intarr(ids),
                            _ptrarr(a), _ptrarr([rks[i][r] for i in ids]), _p(self.params.CRS[r]), _ptrarr(o))

    def conjugate(self, ctIn, cks: dict, ctOut):
        level = ctOut.level()
        ids, a = self._ct_arrays(ctIn)
        _, o = self._ct_arrays(ctOut)
        lib().ork_ks_conjugate(self.ptr, C.c_int(level), C.c_int(len(ids)), _intarr(ids), _ptrarr(a),
                               _ptrarr([cks[i] for i in ids]), _p(self.params.CRS[-2]), _ptrarr(o))


# --------------------------------------------------------------------------------------------
# MK-CKKS evaluator drivers (mkckks/evaluator.go)
# --------------------------------------------------------------------------------------------
class CKKSEvaluator:
    def __init__(self, params: MKParams, scale: float):
        self.params, self.scale = params, float(scale)
        self.ksw = KeySwitcher(params)

    def new_ciphertext(self, ids, level, scale):
        N = self.params.N
        v = {"0": np.zeros((level + 1, N), dtype=np.uint64)}
        for i in ids:
            v[i] = np.zeros((level + 1, N), dtype=np.uint64)
        return Ciphertext(v, scale)

    def _new_binary(self, op0, op1):
        """newCiphertextBinary mkckks/evaluator.go:306-316: union idset, min level"""
        ids = sorted(set(op0.ids()) | set(op1.ids()))
        return self.new_ciphertext(ids, min(op0.level(), op1.level()), max(op0.scale, op1.scale))

    def hoisted_form(self, ct):
        """HoistedForm evaluator.go:543-553"""
        return {i: self.ksw.decompose(ct.level(), ct.value[i]) for i in ct.ids()}

    def rescale(self, ctIn, minScale, ctOut):
        """Rescale evaluator.go:359-398 (returns ValueError where the reference returns error)"""
        q = self.params.Q
        if minScale <= 0:
            raise ValueError("cannot Rescale: minScale is 0")
        if ctIn.scale == 0:
            raise ValueError("cannot Rescale: ciphertext scale is 0")
        if ctIn.level() == 0:
            raise ValueError("cannot Rescale: input Ciphertext already at level 0")
        ctOut.scale = ctIn.scale
        nb = 0
        level = ctIn.level()
        while level - nb >= 0 and ctOut.scale / float(q[level - nb]) >= minScale / 2:
            ctOut.scale /= float(q[level - nb])
            nb += 1
        if nb > 0:
            for k in list(ctOut.value):
                out = ctOut.value[k]
                self.params.ringQ.div_round_by_last_modulus_many(level, nb, ctIn.value[k], out)
                ctOut.value[k] = np.ascontiguousarray(out[:level + 1 - nb])
        return nb

    def mul_relin_hoisted_new(self, op0, op1, h0, h1, rlks):
        """MulRelinHoistedNew / mulRelinHoisted evaluator.go:558-581"""
        ctOut = self._new_binary(op0, op1)
        ctOut.scale = op0.scale * op1.scale
        self.ksw.mul_and_relin_hoisted(op0, op1, h0, h1, rlks, ctOut)
        self.rescale(ctOut, self.scale, ctOut)
        return ctOut

    def mul_relin_new(self, op0, op1, rlks):
        """MulRelinNew evaluator.go:416-443 (hoists into the key set's pool, once if op0 is op1)"""
        if op0 is op1:
            h = self.hoisted_form(op0)
            return self.mul_relin_hoisted_new(op0, op1, h, h, rlks)
        return self.mul_relin_hoisted_new(op0, op1, self.hoisted_form(op0), self.hoisted_form(op1), rlks)

    def _norm_rot(self, rot):
        n2 = self.params.N // 2
        while rot >= n2:
            rot -= n2
        while rot < 0:
            rot += n2
        return rot

    def rotate_hoisted_new(self, ct, rot, hoisted, rks):
        """RotateHoistedNew / rotateHoisted evaluator.go:585-617"""
        out = self.new_ciphertext(ct.ids(), ct.level(), ct.scale)
        rot = self._norm_rot(rot)
        if rot == 0:
            return ct.copy()
        if rot not in self.params.CRS:
            raise RuntimeError("Hoisted rotation only works for precomputed rotation keys")
        self.ksw.rotate_hoisted(ct, rot, hoisted, rks, out)
        return out

    def rotate_new(self, ct, rot, rks):
        """RotateNew / rotate evaluator.go:485-525 (power-of-two chaining when rot is not in the CRS)"""
        out = self.new_ciphertext(ct.ids(), ct.level(), ct.scale)
        rot = self._norm_rot(rot)
        if rot == 0:
            return ct.copy()
        if rot in self.params.CRS:
            self.ksw.rotate(ct, rot, rks, out)
            return out
        tmp = ct.copy()
        k = 1
        while rot > 0:
            if rot % 2:
                self.ksw.rotate(tmp, k, rks, out)
                tmp = out.copy()
            rot //= 2
            k *= 2
        return out

    def conjugate_new(self, ct, cks):
        out = self.new_ciphertext(ct.ids(), ct.level(), ct.scale)
        self.ksw.conjugate(ct, cks, out)
        return out

    # ---- element-wise evaluator ops either side of the key switches ---------------------------
    def get_const_and_scale(self, level, constant):
        """getConstAndScale mkckks/evaluator.go:39-93: a constant with a fractional part is scaled by q_level"""
        scale = 1.0
        if isinstance(constant, complex):
            cReal, cImag = constant.real, constant.imag
        else:
            cReal, cImag = float(constant), 0.0
        if isinstance(constant, (float, complex)):
            for c in (cReal, cImag):
                if c != 0 and c - float(int(c)) != 0:
                    scale = float(self.params.Q[level])
        return cReal, cImag, scale

    @staticmethod
    def scale_up_exact(value, n, q):
        """scaleUpExact mkckks/utils.go:59-86: big.NewFloat(n*value) has 53 bits of precision, so the product and the
        + 0.5 are float64 operations (round to nearest even), Int() truncates, then mod q; negative values give q - res
        (so -0.0... never; a negative multiple of q gives q, not 0)"""
        neg = value < 0
        x = (-n * value) if neg else (n * value)
        x = x + 0.5
        res = int(x) % q
        return q - res if neg else res

    def const_residues(self, level, constant):
        """the two per-limb multipliers of MultByConst (evaluator.go:117-198): coefficients [0, N/2) are multiplied by
        a + b*psi^(N/2)... (NttPsi[i][1]), coefficients [N/2, N) by a - b*(...) -- the reference applies the NTT-domain formula
        to whatever domain the ciphertext is in; reproduced as written.  Returns (first[], second[], scale), plain residues."""
        cReal, cImag, scale = self.get_const_and_scale(level, constant)
        first, second = [], []
        ring = self.params.ringQ
        for i in range(level + 1):
            qi = ring.moduli[i]
            sReal = self.scale_up_exact(cReal, scale, qi) if cReal != 0 else 0
            sc = sReal
            sImag = 0
            if cImag != 0:
                sImag = self.scale_up_exact(cImag, scale, qi)
                psi1 = int(ring.tables(i)[0][1])                       # NttPsi[i][1], Montgomery form
                sImag = (sImag * psi1 * pow(1 << 64, -1, qi)) % qi     # MRed
                sc = sc + sImag
                if sc >= qi:
                    sc -= qi                                           # CRed
            first.append(sc % qi)                                      # MForm reduces
            if cImag != 0:
                sc = sReal + (qi - sImag)
                if sc >= qi:
                    sc -= qi
            second.append(sc % qi)
        return first, second, scale

    def mult_by_const(self, ct0, constant, ctOut):
        """MultByConst evaluator.go:117-198; ctOut may be ct0.  Only the components of ct0 and the limbs <= min level are
        written, like the reference."""
        level = min(ct0.level(), ctOut.level())
        first, second, scale = self.const_residues(level, constant)
        ring = self.params.ringQ
        h = self.params.N // 2
        for u in ct0.value:
            a = ct0.value[u][:level + 1]
            lo = ring.mul_residues(np.ascontiguousarray(a), first, level)
            hi = ring.mul_residues(np.ascontiguousarray(a), second, level)
            lo[:, h:] = hi[:, h:]
            ctOut.value[u][:level + 1] = lo
        ctOut.scale = ct0.scale * scale

    def _evaluate_new(self, c0, c1, sub):
        """AddNew / SubNew = newCiphertextBinary + evaluateInPlace's third branch (ctOut is neither input), evaluator.go:
        200-351: the operand with the smaller scale is multiplied by floor(ratio) when that exceeds 1, components present
        in one operand only are copied, and Sub negates (q - x, unreduced) the ones missing from op0."""
        ctOut = self._new_binary(c0, c1)
        level = min(c0.level(), c1.level(), ctOut.level())
        ring = self.params.ringQ
        s0, s1 = c0.scale, c1.scale
        t0, t1 = c0, c1
        if s1 > s0 and math.floor(s1 / s0) > 1:
            t0 = self.new_ciphertext(c0.ids(), c0.level(), ctOut.scale)
            self.mult_by_const(c0, float(math.floor(s1 / s0)), t0)
        elif s0 > s1 and math.floor(s0 / s1) > 1:
            t1 = self.new_ciphertext(c1.ids(), c1.level(), ctOut.scale)
            self.mult_by_const(c1, float(math.floor(s0 / s1)), t1)
        fn = ring.sub if sub else ring.add
        for k in ctOut.value:
            in0, in1 = k in c0.value, k in c1.value
            if in0 and in1:
                ctOut.value[k] = fn(np.ascontiguousarray(t0.value[k][:level + 1]), np.ascontiguousarray(t1.value[k][:level + 1]), level)
            elif in0:
                ctOut.value[k] = t0.value[k][:level + 1].copy()
            else:
                v = t1.value[k][:level + 1].copy()
                ctOut.value[k] = ring.neg(np.ascontiguousarray(v), level) if sub else v
        return ctOut

    def add_new(self, c0, c1):
        return self._evaluate_new(c0, c1, False)

    def sub_new(self, c0, c1):
        return self._evaluate_new(c0, c1, True)

    def mul_ptxt_new(self, ct, pt_value, pt_scale):
        """MulPtxtNew evaluator.go:465-481: NTT(pt) in Montgomery form, per component NTT, MulCoeffsMontgomery, InvNTT;
        then Rescale to the default scale"""
        level = ct.level()
        ring = self.params.ringQ
        out = self.new_ciphertext(ct.ids(), level, ct.scale * pt_scale)
        ptn = ring.mform(ring.ntt(np.ascontiguousarray(pt_value[:level + 1]), level), level)
        for k in ct.value:
            t = ring.ntt(np.ascontiguousarray(ct.value[k][:level + 1]), level)
            t = ring.mul_mont(t, ptn, level)
            out.value[k] = ring.intt(t, level)
        self.rescale(out, self.scale, out)
        return out

    def drop_level(self, ct, levels):
        level = ct.level()
        for k in ct.value:
            ct.value[k] = np.ascontiguousarray(ct.value[k][:level + 1 - levels])


# --------------------------------------------------------------------------------------------
# CKKS encode / decode for decrypt-precision checks (lattigo ckks.Encoder semantics:
# slot j <-> evaluation at zeta^(5^j)).  Plain numpy FFT of size 2N.
# --------------------------------------------------------------------------------------------
def _rot_group(N):
    idx = np.zeros(N // 2, dtype=np.int64)
    g = 1
    for j in range(N // 2):
        idx[j] = g
        g = (g * 5) % (2 * N)
    return idx


def ckks_encode(params: MKParams, values: np.ndarray, scale: float, level=None):
    N = params.N
    level = params.max_level() if level is None else level
    rg = _rot_group(N)
    V = np.zeros(2 * N, dtype=np.complex128)
    V[rg] = values
    V[2 * N - rg] = np.conj(values)
    c = np.fft.fft(V)[:N].real / N
    ints = [int(round(float(x) * scale)) for x in c]
    out = np.zeros((level + 1, N), dtype=np.uint64)
    for i in range(level + 1):
        q = params.Q[i]
        out[i] = np.array([x % q for x in ints], dtype=np.uint64)
    return out


def crt_centered(moduli, poly):
    """CRT-reconstruct each coefficient to a centred Python int"""
    Qbig = reduce(lambda a, b: a * b, moduli, 1)
    acc = [0] * poly.shape[1]
    for i, q in enumerate(moduli):
        Qi = Qbig // q
        w = Qi * pow(Qi % q, -1, q)
        col = poly[i].tolist()
        acc = [(a + w * c) for a, c in zip(acc, col)]
    half = Qbig // 2
    out = []
    for a in acc:
        a %= Qbig
        out.append(a - Qbig if a > half else a)
    return out


def ckks_decode(params: MKParams, pt: np.ndarray, scale: float):
    N = params.N
    level = pt.shape[0] - 1
    ints = crt_centered(params.Q[:level + 1], pt)
    c = np.array([float(x) / scale for x in ints], dtype=np.float64)
    pad = np.zeros(2 * N, dtype=np.complex128)
    pad[:N] = c
    ev = np.fft.ifft(pad) * (2 * N)          # sum_i c_i e^{+2 pi i ik / 2N}
    return ev[_rot_group(N)]


# --------------------------------------------------------------------------------------------
# MK-BFV (mkbfv/*.go)
# --------------------------------------------------------------------------------------------
class BFVParams(MKParams):
    """mkbfv.Parameters (mkbfv/params.go:27-83): Q, QMul, R = Q u QMul, P, T"""

    def __init__(self, logN, Q, QMul, P, T, gamma=2, seed=0xB2000003, crs_rots=()):
        super().__init__(logN, Q, P, gamma, seed, crs_rots)
        assert len(Q) == len(QMul), "cannot NewParametersFromLiteral: length of Q & QMul is not equal"
        self.QMul = [int(x) for x in QMul]
        self.T = int(T)
        self.ringQMul = Ring(logN, self.QMul)
        self.ringR = Ring(logN, self.Q + self.QMul)


class BFVKeyGenerator(KeyGenerator):
    def gen_bfv_switching_keys(self, sk):
        """GenBFVSwitchingKey mkbfv/keygen.go:91-162: swk_k[i] = MForm(s * G_i + e), G_i = floor(R/m_i * t * [(R/m_i)^-1]_{m_i} * P / QMul)"""
        p = self.params
        Qb = reduce(lambda a, b: a * b, p.Q, 1)
        QMb = reduce(lambda a, b: a * b, p.QMul, 1)
        Rb = Qb * QMb
        Pb = reduce(lambda a, b: a * b, p.P, 1)
        sQ, sP = p.ringQ.invmform(sk.Q), p.ringP.invmform(sk.P)
        outs = []
        for mods in (p.Q, p.QMul):
            swk = p.new_swk()
            for i in range(swk.shape[0]):
                m = mods[i]
                Gi = Rb // m
                Ti = pow(Gi % m, -1, m)
                Gi = (Gi * p.T * Ti * Pb) // QMb
                q = p.ringQ.mul_bigint(sQ, Gi)
                pp = p.ringP.mul_bigint(sP, Gi)
                e = self._gaussian_error_ntt()
                q = p.ringQ.add(q, np.ascontiguousarray(e[:p.nQ]))
                pp = p.ringP.add(pp, np.ascontiguousarray(e[p.nQ:]))
                swk[i, :p.nQ] = p.ringQ.mform(q)
                swk[i, p.nQ:] = p.ringP.mform(pp)
            outs.append(swk)
        return outs

    def gen_bfv_relin_key(self, sk, r):
        """GenRelinearizationKey mkbfv/keygen.go:24-88"""
        p = self.params
        a1, a2 = p.CRS[0], p.CRS[-3]
        b1, b2 = p.new_swk(), p.new_swk()
        for i in range(b1.shape[0]):                            # :51-63: e1, e2 drawn alternately, digit by digit
            b1[i] = self._gen_b_digit(a1[i], sk)
            b2[i] = self._gen_b_digit(a2[i], sk)
        d1, d2 = self.gen_bfv_switching_keys(sk)
        self._mul_sub_digits(a1, r, d1)
        self._mul_sub_digits(a2, r, d2)
        v = self._gen_v(sk, r)
        return BFVRelinKey(sk.id, b1, d1, v, b2, d2)


class BFVEvaluator:
    def __init__(self, params: BFVParams):
        self.params = params
        p = params
        QMb = reduce(lambda a, b: a * b, p.QMul, 1)
        self.ptr = C.c_void_p(lib().ork_bfv_new(p.ringQ.ptr, p.ringQMul.ptr, p.ringR.ptr, p.ringP.ptr, C.c_int(p.gamma),
                                                C.c_uint64(p.T), _u64arr([QMb % q for q in p.Q])))

    def modup_q_to_r(self, polyQ):
        out = np.zeros((2 * self.params.nQ, self.params.N), dtype=np.uint64)
        lib().ork_bfv_modup_q_to_r(self.ptr, _p(polyQ), _p(out))
        return out

    def rescale_q_to_r(self, polyQ):
        out = np.zeros((2 * self.params.nQ, self.params.N), dtype=np.uint64)
        lib().ork_bfv_rescale(self.ptr, _p(polyQ), _p(out))
        return out

    def quantize(self, polyR):
        out = np.zeros((self.params.nQ, self.params.N), dtype=np.uint64)
        lib().ork_bfv_quantize(self.ptr, _p(polyR), _p(out))
        return out

    def decompose_bfv(self, levelQ, aR):
        ad1, ad2 = self.params.new_swk(), self.params.new_swk()
        lib().ork_bfv_decompose(self.ptr, C.c_int(levelQ), _p(aR), _p(ad1), _p(ad2))
        return ad1, ad2

    def mul_relin_new(self, ct0, ct1, rlks: dict):
        """MulRelinNew -> mulRelinHoisted mkbfv/evaluator.go:84-150"""
        p = self.params
        level = p.max_level()
        ct0R = Ciphertext({k: self.modup_q_to_r(v) for k, v in ct0.value.items()})
        ct1R = Ciphertext({k: self.rescale_q_to_r(v) for k, v in ct1.value.items()})
        h0 = {i: self.decompose_bfv(level, ct0R.value[i]) for i in ct0.ids()}
        h1 = {i: self.decompose_bfv(level, ct1R.value[i]) for i in ct1.ids()}
        ids = sorted(set(ct0.ids()) | set(ct1.ids()))
        out = Ciphertext({k: np.zeros((level + 1, p.N), dtype=np.uint64) for k in ["0"] + ids})
        self.mul_and_relin_bfv_hoisted(ct0R, ct1R, h0, h1, rlks, out)
        return out

    def mul_and_relin_bfv_hoisted(self, ct0R, ct1R, h0, h1, rlks, out):
        p = self.params
        level = out.level()
        ids0, a0 = KeySwitcher._ct_arrays(ct0R)
        ids1, a1 = KeySwitcher._ct_arrays(ct1R)
        idsO, aO = KeySwitcher._ct_arrays(out)
        maxid = max(list(rlks) + [0])
        by = lambda f: KeySwitcher._by_id({i: f(k) for i, k in rlks.items()}, maxid)
        lib().ork_bfv_mul_and_relin_hoisted(
            self.ptr, C.c_int(level),
            C.c_int(len(ids0)), _intarr(ids0), _ptrarr(a0),
            None if h0 is None else _ptrarr([h0[i][0] for i in ids0]), None if h0 is None else _ptrarr([h0[i][1] for i in ids0]),
            C.c_int(len(ids1)), _intarr(ids1), _ptrarr(a1),
            None if h1 is None else _ptrarr([h1[i][0] for i in ids1]), None if h1 is None else _ptrarr([h1[i][1] for i in ids1]),
            by(lambda k: k.b1), by(lambda k: k.b2), by(lambda k: k.d1), by(lambda k: k.d2), by(lambda k: k.v),
            _p(p.CRS[-1]), C.c_int(len(idsO)), _intarr(idsO), _ptrarr(aO))


def bfv_encode(params: BFVParams, m: np.ndarray):
    """coefficient-wise plaintext: round(Q*m/T) mod Q (lattigo bfv ScaleUp semantics)"""
    Qb = reduce(lambda a, b: a * b, params.Q, 1)
    T = params.T
    ints = [((Qb * int(x)) + T // 2) // T for x in m]
    out = np.zeros((params.nQ, params.N), dtype=np.uint64)
    for i, q in enumerate(params.Q):
        out[i] = np.array([x % q for x in ints], dtype=np.uint64)
    return out


def bfv_decode(params: BFVParams, pt: np.ndarray):
    """round(T*x/Q) mod T (lattigo bfv ScaleDown semantics)"""
    Qb = reduce(lambda a, b: a * b, params.Q, 1)
    T = params.T
    ints = crt_centered(params.Q, pt)
    return np.array([((T * x + Qb // 2) // Qb) % T for x in ints], dtype=np.int64)
