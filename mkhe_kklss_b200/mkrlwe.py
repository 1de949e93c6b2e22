"""Host-side mirror of the reference's `mkrlwe` package for the hot path, over the C ABI (include/mkhe.h).

Same names and argument meaning as the Go API (mkrlwe/keyswitch.go, keyswitch_hoisted.go, elements.go,
keys.go, idset.go); a Go `panic(msg)` becomes `raise RuntimeError(msg)` with the same message.
Everything heavy lives on the device: `Poly` / `SwitchingKey` are handles, a `Ciphertext` is a
`map id -> Poly` exactly like `Ciphertext.Value map[string]*ring.Poly` (elements.go:17-19).

Keys and ciphertexts made elsewhere arrive as host arrays with the reference's byte layout and are uploaded once; `KeyGenerator`,
`Encryptor` and `Decryptor` below make / use them on the device from the documented counter-based
streams of include/mkhe_prng.h, so a key set never crosses PCIe.
"""
from __future__ import annotations

import numpy as np

from ._lib import Context, MkheError


class IDSet:
    """mkrlwe/idset.go"""

    def __init__(self, ids=()):
        self.Value = set(ids)

    def Add(self, id):
        self.Value.add(id)

    def Has(self, id):
        return id in self.Value

    def Union(self, other):
        return IDSet(self.Value | other.Value)

    def Size(self):
        return len(self.Value)

    def sorted(self):
        return sorted(self.Value)


class Poly:
    """device-resident ring.Poly (coefficient domain unless stated otherwise)"""

    def __init__(self, ctx: Context, nlimbs: int, handle=None):
        self.ctx = ctx
        self.h = ctx.poly_alloc(nlimbs) if handle is None else handle
        self.cap = nlimbs

    @classmethod
    def from_numpy(cls, ctx, arr, cap=None):
        arr = np.ascontiguousarray(arr, dtype=np.uint64)
        p = cls(ctx, cap or arr.shape[0])
        ctx.poly_upload(p.h, arr)
        ctx.poly_set_nlimbs(p.h, arr.shape[0])
        return p

    def numpy(self):
        return self.ctx.poly_download(self.h)

    def nlimbs(self):
        return self.ctx.poly_get_nlimbs(self.h)

    def set_nlimbs(self, n):
        self.ctx.poly_set_nlimbs(self.h, n)

    def free(self):
        if self.h:
            self.ctx.poly_free(self.h)
            self.h = 0

    def __del__(self):
        # what runtime.SetFinalizer does on the Go side: a dropped poly returns its block to the device pool
        # (stream-ordered: no synchronisation).  A context that is already closed has released everything itself.
        try:
            if self.h and getattr(self.ctx, "ptr", None):
                self.free()
        except Exception:
            pass


class SwitchingKey:
    """mkrlwe.SwitchingKey{Value []rlwe.PolyQP} (keys.go:23-25) on the device; also the value type of a
    HoistedCiphertext (elements.go:5-15).  Host layout: uint64[beta_max, nQ+nP, N]."""

    def __init__(self, ctx: Context, arr=None):
        self.ctx = ctx
        self.h = ctx.swk_alloc()
        if arr is not None:
            ctx.swk_upload(self.h, arr)

    def numpy(self):
        return self.ctx.swk_download(self.h)

    def free(self):
        if self.h:
            self.ctx.swk_free(self.h)
            self.h = 0

    def __del__(self):
        try:
            if self.h and getattr(self.ctx, "ptr", None):
                self.free()
        except Exception:
            pass


class Ciphertext:
    """mkrlwe.Ciphertext (elements.go:17-62): Value["0"] plus one poly per party id."""

    def __init__(self, value: dict):
        self.Value = value

    @classmethod
    def new(cls, ctx, idset, level, cap=None):
        """NewCiphertext (elements.go:22-33)"""
        ids = idset.sorted() if isinstance(idset, IDSet) else sorted(idset)
        v = {"0": Poly(ctx, cap or level + 1)}
        for i in ids:
            v[i] = Poly(ctx, cap or level + 1)
        ct = cls(v)
        for p in v.values():
            p.set_nlimbs(level + 1)
        return ct

    @classmethod
    def from_numpy(cls, ctx, value: dict, cap=None):
        return cls({k: Poly.from_numpy(ctx, a, cap) for k, a in value.items()})

    def numpy(self):
        return {k: p.numpy() for k, p in self.Value.items()}

    def IDSet(self):
        return IDSet(k for k in self.Value if k != "0")

    def ids(self):
        return sorted(k for k in self.Value if k != "0")

    def Level(self):
        return self.Value["0"].nlimbs() - 1

    def handles(self, ids=None):
        ids = self.ids() if ids is None else ids
        return [self.Value["0"].h] + [self.Value[i].h for i in ids]

    def free(self):
        for p in self.Value.values():
            p.free()


class RelinearizationKey:
    """mkrlwe.RelinearizationKey{Value [3]*SwitchingKey} = (b, d, v) (keys.go:34-37)"""

    def __init__(self, ctx, id, b, d, v):
        self.ID = id
        self.Value = [SwitchingKey(ctx, b), SwitchingKey(ctx, d), SwitchingKey(ctx, v)]


class RelinearizationKeySet:
    """mkrlwe.RelinearizationKeySet (keys.go:53-57,165-198).  HoistPool lives in the context."""

    def __init__(self):
        self.Value = {}

    def AddRelinearizationKey(self, rlk):
        self.Value[rlk.ID] = rlk

    def GetRelinearizationKey(self, id):
        if id not in self.Value:
            raise RuntimeError("cannot GetRelinearizationKey: there is no relinearization key for given id")
        return self.Value[id]


class RotationKeySet:
    """mkrlwe.RotationKeySet map[id]map[rot] (keys.go:60-62,133-163)"""

    def __init__(self):
        self.Value = {}

    def AddRotationKey(self, id, rotidx, swk: SwitchingKey):
        self.Value.setdefault(id, {})[rotidx] = swk

    def GetRotationKey(self, id, rotidx):
        if id not in self.Value:
            raise RuntimeError("cannot GetRotationKey: there is no rotation key for given id")
        if rotidx not in self.Value[id]:
            raise RuntimeError("cannot GetRotationKey: there is no rotation key for given rotidx")
        return self.Value[id][rotidx]


class ConjugationKeySet:
    def __init__(self):
        self.Value = {}

    def AddConjugationKey(self, id, swk):
        self.Value[id] = swk

    def GetConjugationKey(self, id):
        if id not in self.Value:
            raise RuntimeError("cannot GetConjugationKey: there is no conjugation key for given id")
        return self.Value[id]


class Parameters:
    """mkrlwe.Parameters (params.go:8-75): the device context plus the uploaded CRS switching keys."""

    def __init__(self, logN, Q, P, gamma=2, device=0, QMul=None, T=0, lib=None):
        self.ctx = Context(logN, Q, P, gamma, device, QMul, T, lib)
        self.logN, self.n, self.gamma = logN, 1 << logN, gamma
        self.Q, self.P = list(Q), list(P)
        self.CRS = {}

    def N(self):
        return self.n

    def QCount(self):
        return len(self.Q)

    def PCount(self):
        return len(self.P)

    def MaxLevel(self):
        return len(self.Q) - 1

    def Alpha(self):
        return len(self.P) // self.gamma

    def Beta(self, levelQ):
        return -(-(levelQ + 1) // self.Alpha())

    def Gamma(self):
        return self.gamma

    def SetCRS(self, idx, arr):
        """upload a CRS[idx] made elsewhere (params.go:37-58)"""
        self.CRS[idx] = SwitchingKey(self.ctx, arr)

    CRS_STREAM_STRIDE = 1 << 20          # streams of CRS[idx] start at (idx mod 2^32) * 2^20: disjoint for every index

    def AddCRS(self, idx, seed):
        """params.go:77-99 on the device: beta uniform QP polys in Montgomery form from (seed, streams of idx).  Every party
        that uses the same seed obtains the same CRS -- which is what a common reference string is."""
        swk = SwitchingKey(self.ctx)
        self.ctx.sample_crs(seed, (idx & 0xFFFFFFFF) * self.CRS_STREAM_STRIDE, swk.h)
        self.CRS[idx] = swk
        return swk


class FastBasisExtender:
    """mkrlwe.FastBasisExtender (basis_extension.go:12-357): the conversions the key switch uses are fused into its kernels
    (k_bcast_ntt_pass1 / k_decomp_lift = ModUp of a digit, k_moddown_P / k_moddown_Q = ModDownQPtoQ); the stand-alone NTT-domain
    ModDown is exposed for completeness of the type."""

    def __init__(self, params):
        self.ctx = params.ctx

    def ModDownQPtoQNTT(self, levelQ, levelP, p1QP: Poly, p2Q: Poly):
        """basis_extension.go:239-290; p1QP = the PolyQP as one device poly (Q limbs then P limbs), levelP = the maximum"""
        if levelP != self.ctx.nP - 1:
            raise RuntimeError("ModDownQPtoQNTT: levelP must be the maximum level of P")
        self.ctx.moddown_qp_to_q_ntt(levelQ, p1QP.h, p2Q.h)


class KeySwitcher:
    """mkrlwe.KeySwitcher (keyswitch.go:8-16): method bodies are C-ABI calls, pools live in the context."""

    def __init__(self, params: Parameters):
        self.Parameters = params
        self.ctx = params.ctx

    def Decompose(self, levelQ, a: Poly, ad: SwitchingKey):
        """keyswitch.go:49-73"""
        self.ctx.decompose(levelQ, a.h, ad.h)

    def ExternalProduct(self, levelQ, a: Poly, bg: SwitchingKey, c: Poly):
        """keyswitch.go:79-118"""
        self.ctx.external_product(levelQ, a.h, bg.h, c.h)

    def ExternalProductHoisted(self, levelQ, aHoisted: SwitchingKey, bg: SwitchingKey, c: Poly):
        """keyswitch_hoisted.go:10-40"""
        self.ctx.external_product_hoisted(levelQ, aHoisted.h, bg.h, c.h)

    def MulAndRelinHoisted(self, op0, op1, op0Hoisted, op1Hoisted, rlkSet, ctOut):
        """keyswitch_hoisted.go:44-179; op*Hoisted is a dict id -> SwitchingKey or None (Go nil)"""
        level = ctOut.Level()
        if op0.Level() < level:
            raise RuntimeError("Cannot MulAndRelin: op0 and op1 have different levels")
        ids0, ids1, idsO = op0.ids(), op1.ids(), ctOut.ids()
        if set(idsO) != set(ids0) | set(ids1):
            raise RuntimeError("Cannot MulAndRelin: ctOut idset is not the union of the operands' idsets")
        self.ctx.mul_relin_hoisted(
            level, ids0, op0.handles(ids0), None if op0Hoisted is None else [op0Hoisted[i].h for i in ids0],
            ids1, op1.handles(ids1), None if op1Hoisted is None else [op1Hoisted[i].h for i in ids1],
            [rlkSet.GetRelinearizationKey(i).Value[0].h for i in ids1],
            [rlkSet.GetRelinearizationKey(i).Value[1].h for i in ids0],
            [rlkSet.GetRelinearizationKey(i).Value[2].h for i in ids0],
            self.Parameters.CRS[-1].h, idsO, ctOut.handles(idsO))

    def MulAndRelin(self, op0, op1, rlkSet, ctOut):
        """keyswitch.go:122-230"""
        self.MulAndRelinHoisted(op0, op1, None, None, rlkSet, ctOut)

    def _norm(self, rotidx):
        while rotidx < 0:
            rotidx += self.Parameters.N() // 2
        return rotidx

    def RotateHoisted(self, ctIn, rotidx, ctInHoisted, rkSet, ctOut):
        """keyswitch_hoisted.go:183-247"""
        level = ctOut.Level()
        if ctIn.Level() < level:
            raise RuntimeError("Cannot Rotate: ctIn and ctOut have different levels")
        rotidx = self._norm(rotidx)
        ids = ctIn.ids()
        self.ctx.rotate_hoisted(level, rotidx, ctIn.handles(ids), [ctInHoisted[i].h for i in ids],
                                [rkSet.GetRotationKey(i, rotidx).h for i in ids], self.Parameters.CRS[rotidx].h,
                                ctOut.handles(ids))

    def Rotate(self, ctIn, rotidx, rkSet, ctOut):
        """keyswitch.go:234-298"""
        level = ctOut.Level()
        if ctIn.Level() < level:
            raise RuntimeError("Cannot Rotate: ctIn and ctOut have different levels")
        rotidx = self._norm(rotidx)
        ids = ctIn.ids()
        self.ctx.rotate(level, rotidx, ctIn.handles(ids), [rkSet.GetRotationKey(i, rotidx).h for i in ids],
                        self.Parameters.CRS[rotidx].h, ctOut.handles(ids))

    def Conjugate(self, ctIn, ckSet, ctOut):
        """keyswitch.go:302-332"""
        level = ctOut.Level()
        if ctIn.Level() < level:
            raise RuntimeError("Cannot Conjugate: ctIn and ctOut have different levels")
        ids = ctIn.ids()
        self.ctx.conjugate(level, ctIn.handles(ids), [ckSet.GetConjugationKey(i).h for i in ids],
                           self.Parameters.CRS[-2].h, ctOut.handles(ids))


class Decryptor:
    """mkrlwe.Decryptor (decryptor.go:8-66) on the device: `skSet` maps a party id to its SecretKey.Value.Q uploaded as a Poly
    (NTT domain, Montgomery form).  Returns the coefficient-domain plaintext poly (rlwe.Plaintext.Value); decoding stays on the
    host like in the reference (mkckks/decryptor.go)."""

    def __init__(self, params):
        self.params = params
        self.ctx = params.ctx

    def Decrypt(self, ciphertext, skSet):
        ids = ciphertext.ids()
        for i in ids:
            if i not in skSet:
                raise RuntimeError("Cannot Decrypt: there is a missing secretkey")
        level = ciphertext.Level()
        pt = Poly(self.ctx, level + 1)
        self.ctx.decrypt(level, ciphertext.handles(ids), [skSet[i].h for i in ids], pt.h)
        return pt


class SecretKey:
    """mkrlwe.SecretKey{Value rlwe.PolyQP, ID} (keys.go:10-13) on the device: one poly of nQ + nP limbs (Q limbs then P limbs),
    NTT domain, Montgomery form.  `.h` is what Decryptor / the key generator take."""

    def __init__(self, ctx, id, poly=None):
        self.ID = id
        self.Value = poly if poly is not None else Poly(ctx, ctx.D)
        self.h = self.Value.h


class PublicKey:
    """mkrlwe.PublicKey{Value [2]rlwe.PolyQP, ID} (keys.go:15-19)"""

    def __init__(self, ctx, id):
        self.ID = id
        self.Value = [Poly(ctx, ctx.D), Poly(ctx, ctx.D)]


class KeyGenerator:
    """mkrlwe.KeyGenerator (keygen.go) on the device.  The reference seeds a Blake2b PRNG from crypto/rand per generator; here the
    caller gives (seed, first stream) of the counter-based generator "mkhe-ctr-1" (include/mkhe_prng.h) and every call consumes
    the number of streams include/mkhe.h documents, in the order the Go methods draw their polynomials, so a CPU
    restatement of keygen.go drawing from the same streams makes the same keys bit for bit (tests/parity.py::check_keygen)."""

    def __init__(self, params: Parameters, seed: int, stream: int = 0):
        self.params, self.ctx = params, params.ctx
        self.seed, self.stream = int(seed), int(stream)

    def _take(self, n):
        s = self.stream
        self.stream += n
        return s

    def GenSecretKey(self, id):
        """keygen.go:58-60"""
        return self.GenSecretKeyWithDistrib(0.5, id)

    def GenSecretKeyWithDistrib(self, p, id):
        """keygen.go:68-76"""
        sk = SecretKey(self.ctx, id)
        self.ctx.keygen_secret(self.seed, self._take(1), p, sk.h)
        return sk

    def GenPublicKey(self, sk: SecretKey):
        """keygen.go:88-109"""
        pk = PublicKey(self.ctx, sk.ID)
        self.ctx.keygen_public(self.seed, self._take(1), sk.h, self.params.CRS[0].h, pk.Value[0].h, pk.Value[1].h)
        return pk

    def GenKeyPair(self, id):
        sk = self.GenSecretKey(id)
        return sk, self.GenPublicKey(sk)

    def GenSwitchingKey(self, skIn: SecretKey, swk: SwitchingKey):
        """keygen.go:269-327"""
        self.ctx.keygen_switching_key(self.seed, self._take(self.ctx.beta_max), skIn.h, swk.h)

    def GenRelinearizationKey(self, sk: SecretKey, r: SecretKey):
        """keygen.go:137-187"""
        if self.params.PCount() == 0:
            raise RuntimeError("modulus P is empty")
        rlk = RelinearizationKey(self.ctx, sk.ID, None, None, None)
        self.ctx.keygen_relin(self.seed, self._take(3 * self.ctx.beta_max), sk.h, r.h, self.params.CRS[0].h, self.params.CRS[-1].h,
                              rlk.Value[0].h, rlk.Value[1].h, rlk.Value[2].h)
        return rlk

    def GenRotationKey(self, rotidx, sk: SecretKey):
        """keygen.go:190-229"""
        if rotidx not in self.params.CRS:
            raise RuntimeError("Cannot GenRotationKey: CRS for given rot idx is not generated")
        rk = SwitchingKey(self.ctx)
        self.ctx.keygen_rotation(self.seed, self._take(self.ctx.beta_max), rotidx, sk.h, self.params.CRS[rotidx].h, rk.h)
        return rk

    def GenDefaultRotationKeys(self, sk: SecretKey, rtkSet: RotationKeySet):
        """keygen.go:232-237: the powers of two"""
        for i in range(self.params.logN - 1):
            rtkSet.AddRotationKey(sk.ID, 1 << i, self.GenRotationKey(1 << i, sk))

    def GenConjugationKey(self, sk: SecretKey):
        """keygen.go:240-266"""
        ck = SwitchingKey(self.ctx)
        self.ctx.keygen_conjugation(self.seed, self._take(self.ctx.beta_max), sk.h, self.params.CRS[-2].h, ck.h)
        return ck


class Encryptor:
    """mkrlwe.Encryptor.Encrypt (encryptor.go:55-118), coefficient-domain ciphertexts (what mkckks / mkbfv make), on the device.
    Each call takes three streams (w, e0, e1) of (seed, stream)."""

    def __init__(self, params: Parameters, seed: int, stream: int = 0):
        self.params, self.ctx = params, params.ctx
        self.seed, self.stream = int(seed), int(stream)

    def Encrypt(self, plaintext: Poly, pk: PublicKey, ctOut: Ciphertext):
        """plaintext: coefficient-domain rlwe.Plaintext.Value on the device (or None: an encryption of zero)"""
        levelQ = ctOut.Level() if plaintext is None else min(plaintext.nlimbs() - 1, ctOut.Level())
        if pk.ID not in ctOut.Value:
            ctOut.Value[pk.ID] = Poly(self.ctx, levelQ + 1)
        s = self.stream
        self.stream += 3
        self.ctx.encrypt(self.seed, s, levelQ, None if plaintext is None else plaintext.h, pk.Value[0].h, pk.Value[1].h,
                         ctOut.Value["0"].h, ctOut.Value[pk.ID].h)
        ctOut.Value["0"].set_nlimbs(levelQ + 1)
        ctOut.Value[pk.ID].set_nlimbs(levelQ + 1)
