"""Host-side mirror of the reference's `mkbfv` hot path (mkbfv/evaluator.go, keyswitch*.go,
basis_extension.go) over the C ABI."""
from __future__ import annotations

from . import mkrlwe
from .mkrlwe import Poly, SwitchingKey


class Parameters(mkrlwe.Parameters):
    """mkbfv.Parameters (mkbfv/params.go:27-107): Q, QMul, P, T; R = Q u QMul; gamma = 2"""

    def __init__(self, logN, Q, QMul, P, T, device=0, lib=None):
        if len(Q) != len(QMul):
            raise RuntimeError("cannot NewParametersFromLiteral: length of Q & QMul is not equal")
        super().__init__(logN, Q, P, 2, device, QMul=QMul, T=T, lib=lib)
        self.QMul, self.t = list(QMul), int(T)

    def T(self):
        return self.t


class RelinearizationKey:
    """mkbfv.RelinearizationKey{Value [2]*mkrlwe.RelinearizationKey} (mkbfv/keys.go:6-9); only
    Value[0].Value[2] (= v) of the third slot is used (mkbfv/keyswitch_hoisted.go:191)."""

    def __init__(self, ctx, id, b1=None, d1=None, v=None, b2=None, d2=None):
        self.ID = id
        self.b1, self.d1, self.v = SwitchingKey(ctx, b1), SwitchingKey(ctx, d1), SwitchingKey(ctx, v)
        self.b2, self.d2 = SwitchingKey(ctx, b2), SwitchingKey(ctx, d2)


class RelinearizationKeySet:
    def __init__(self):
        self.Value = {}

    def AddRelinearizationKey(self, rlk):
        self.Value[rlk.ID] = rlk

    def GetRelinearizationKey(self, id):
        if id not in self.Value:
            raise RuntimeError("cannot GetRelinearizationKey: there is no relinearization key for given id")
        return self.Value[id]


class KeyGenerator(mkrlwe.KeyGenerator):
    """mkbfv.KeyGenerator (mkbfv/keygen.go:8-162) on the device"""

    def GenRelinearizationKey(self, sk, r):
        """mkbfv/keygen.go:24-88 with GenBFVSwitchingKey (:91-162)"""
        rlk = RelinearizationKey(self.ctx, sk.ID)
        crs = self.params.CRS
        self.ctx.keygen_bfv_relin(self.seed, self._take(5 * self.ctx.beta_max), sk.h, r.h, crs[0].h, crs[-3].h, crs[-1].h,
                                  rlk.b1.h, rlk.b2.h, rlk.d1.h, rlk.d2.h, rlk.v.h)
        return rlk


class FastBasisExtender:
    """mkbfv.FastBasisExtender (mkbfv/basis_extension.go)"""

    def __init__(self, params: Parameters):
        self.ctx = params.ctx

    def ModUpQtoR(self, polyQ: Poly, polyR: Poly):
        self.ctx.bfv_modup_q_to_r(polyQ.h, polyR.h)

    def Rescale(self, polyQ: Poly, polyR: Poly):
        self.ctx.bfv_rescale_q_to_r(polyQ.h, polyR.h)

    def Quantize(self, polyR: Poly, polyQ: Poly):
        self.ctx.bfv_quantize(polyR.h, polyQ.h)


class KeySwitcher(mkrlwe.KeySwitcher):
    """mkbfv.KeySwitcher (mkbfv/keyswitch.go:6-55)"""

    def DecomposeBFV(self, levelQ, aR: Poly, ad1: SwitchingKey, ad2: SwitchingKey):
        """mkbfv/keyswitch.go:57-81"""
        self.ctx.bfv_decompose(levelQ, aR.h, ad1.h, ad2.h)

    def MulAndRelinBFVHoisted(self, op0, op1, op0Hoisted1, op0Hoisted2, op1Hoisted1, op1Hoisted2, rlkSet, ctOut):
        """mkbfv/keyswitch_hoisted.go:39-207 (op0/op1 in basis R)"""
        level = ctOut.Level()
        ids0, ids1, idsO = op0.ids(), op1.ids(), ctOut.ids()
        g = rlkSet.GetRelinearizationKey
        hs = lambda h, ids: None if h is None else [h[i].h for i in ids]
        self.ctx.bfv_mul_relin_hoisted(
            level, ids0, op0.handles(ids0), hs(op0Hoisted1, ids0), hs(op0Hoisted2, ids0),
            ids1, op1.handles(ids1), hs(op1Hoisted1, ids1), hs(op1Hoisted2, ids1),
            [g(i).b1.h for i in ids1], [g(i).b2.h for i in ids1], [g(i).d1.h for i in ids0], [g(i).d2.h for i in ids0],
            [g(i).v.h for i in ids0], self.Parameters.CRS[-1].h, idsO, ctOut.handles(idsO))


class Evaluator:
    """mkbfv.Evaluator (mkbfv/evaluator.go)"""

    def __init__(self, params: Parameters):
        self.params = params
        self.ctx = params.ctx
        self.ksw = KeySwitcher(params)
        self.conv = FastBasisExtender(params)

    def MulRelinNew(self, op0, op1, rlkSet):
        """mkbfv/evaluator.go:84-150: ModUpQtoR / Rescale of every component, DecomposeBFV, MulAndRelinBFVHoisted
        -- one device call (mkhe_bfv_mul_relin)."""
        idset = op0.IDSet().Union(op1.IDSet())
        ctOut = mkrlwe.Ciphertext.new(self.ctx, idset, self.params.MaxLevel())
        ids0, ids1, idsO = op0.ids(), op1.ids(), ctOut.ids()
        g = rlkSet.GetRelinearizationKey
        self.ctx.bfv_mul_relin(
            ids0, op0.handles(ids0), ids1, op1.handles(ids1),
            [g(i).b1.h for i in ids1], [g(i).b2.h for i in ids1], [g(i).d1.h for i in ids0], [g(i).d2.h for i in ids0],
            [g(i).v.h for i in ids0], self.params.CRS[-1].h, idsO, ctOut.handles(idsO))
        return ctOut

    # -- the rest of the public evaluator (mkbfv/evaluator.go:25-82,152-226): element-wise ops on ring Q and the key switches
    #    of mkrlwe.KeySwitcher, all at the single (maximum) level BFV works at
    def newCiphertextBinary(self, op0, op1):
        return mkrlwe.Ciphertext.new(self.ctx, op0.IDSet().Union(op1.IDSet()), self.params.MaxLevel())

    def _evaluate_new(self, op0, op1, sub):
        """evaluateInPlace (evaluator.go:31-46) + Sub's NegLvl of the components missing from op0 (:67-75)"""
        ctOut = self.newCiphertextBinary(op0, op1)
        level = self.params.MaxLevel()
        for k in ctOut.Value:
            in0, in1 = k in op0.Value, k in op1.Value
            o = ctOut.Value[k].h
            if in0 and in1:
                (self.ctx.poly_sub if sub else self.ctx.poly_add)(level, op0.Value[k].h, op1.Value[k].h, o)
            elif in0:
                self.ctx.poly_copy_lvl(level, o, op0.Value[k].h)
            elif sub:
                self.ctx.poly_neg(level, op1.Value[k].h, o)
            else:
                self.ctx.poly_copy_lvl(level, o, op1.Value[k].h)
        return ctOut

    def AddNew(self, op0, op1):
        return self._evaluate_new(op0, op1, False)

    def SubNew(self, op0, op1):
        return self._evaluate_new(op0, op1, True)

    def _copy(self, ct):
        out = mkrlwe.Ciphertext.new(self.ctx, ct.IDSet(), self.params.MaxLevel())
        for k in ct.Value:
            self.ctx.poly_copy(out.Value[k].h, ct.Value[k].h)
        return out

    def RotateNew(self, ct0, rotidx, rkSet):
        """evaluator.go:152-196 (power-of-two chaining when rotidx has no CRS entry)"""
        n2 = self.params.N() // 2
        rotidx %= n2
        if rotidx == 0:
            return self._copy(ct0)
        ctOut = mkrlwe.Ciphertext.new(self.ctx, ct0.IDSet(), self.params.MaxLevel())
        if rotidx in self.params.CRS:
            self.ksw.Rotate(ct0, rotidx, rkSet, ctOut)
            return ctOut
        ctTmp = self._copy(ct0)
        k = 1
        while rotidx > 0:
            if rotidx % 2 != 0:
                self.ksw.Rotate(ctTmp, k, rkSet, ctOut)
                for key in ctOut.Value:
                    self.ctx.poly_copy(ctTmp.Value[key].h, ctOut.Value[key].h)
            rotidx //= 2
            k *= 2
        ctTmp.free()
        return ctOut

    def ConjugateNew(self, ct0, ckSet):
        """evaluator.go:198-226"""
        ctOut = mkrlwe.Ciphertext.new(self.ctx, ct0.IDSet(), self.params.MaxLevel())
        self.ksw.Conjugate(ct0, ckSet, ctOut)
        return ctOut
