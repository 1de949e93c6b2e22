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

    def __init__(self, ctx, id, b1, d1, v, b2, d2):
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
