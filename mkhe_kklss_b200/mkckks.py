"""Host-side mirror of the reference's `mkckks.Evaluator` hot-path methods (mkckks/evaluator.go) over the
C ABI.  Scale / level bookkeeping stays on the host exactly as in Go; polynomial work is on the device.
"""
from __future__ import annotations

from . import mkrlwe
from .mkrlwe import IDSet, Poly, SwitchingKey


class Parameters(mkrlwe.Parameters):
    """mkckks.Parameters (mkckks/params.go:9-40): mkrlwe parameters + default scale.  The reference hard-wires gamma = 2
    (:20); the keyword exists so the tests can reach other digit widths alpha = #P/gamma."""

    def __init__(self, logN, Q, P, scale, device=0, lib=None, gamma=2):
        super().__init__(logN, Q, P, gamma, device, lib=lib)
        self.scale = float(scale)

    def Scale(self):
        return self.scale


class Ciphertext(mkrlwe.Ciphertext):
    """mkckks.Ciphertext{*mkrlwe.Ciphertext, Scale} (mkckks/elements.go:11-17)"""

    def __init__(self, value, scale=0.0):
        super().__init__(value)
        self.Scale = float(scale)

    @classmethod
    def new(cls, params, idset, level, scale):
        base = mkrlwe.Ciphertext.new(params.ctx, idset, level)
        return cls(base.Value, scale)

    @classmethod
    def from_numpy(cls, ctx, value, scale, cap=None):
        base = mkrlwe.Ciphertext.from_numpy(ctx, value, cap)
        return cls(base.Value, scale)

    def ScalingFactor(self):
        return self.Scale


class Evaluator:
    """mkckks.Evaluator (evaluator.go:13-37)"""

    def __init__(self, params: Parameters):
        self.params = params
        self.ctx = params.ctx
        self.ksw = mkrlwe.KeySwitcher(params)

    def ShallowCopy(self):
        """a second evaluator over the same parameters, keys and ciphertexts whose ops run on their own lane of the device
        (mkhe_ctx_fork): what another NewEvaluator / NewKeySwitcher over one Parameters value is in the reference
        (mkrlwe/keyswitch.go:33-47: private pools, shared tables; cf. FastBasisExtender.ShallowCopy,
        mkrlwe/basis_extension.go:155-175).  Independent ops issued on the two evaluators overlap on the GPU; uses of one
        ciphertext on both are ordered by the library."""
        import copy
        p2 = copy.copy(self.params)
        p2.ctx = self.params.ctx.fork()
        return Evaluator(p2)

    # -- helpers ------------------------------------------------------------------------------
    def newCiphertextBinary(self, op0, op1):
        """evaluator.go:306-316: union idset, min level, max scale"""
        idset = op0.IDSet().Union(op1.IDSet())
        return Ciphertext.new(self.params, idset, min(op0.Level(), op1.Level()), max(op0.Scale, op1.Scale))

    def DropLevel(self, ct0, levels):
        """evaluator.go:106-112: a reslice on the host, a view change on the device"""
        level = ct0.Level()
        for p in ct0.Value.values():
            p.set_nlimbs(level + 1 - levels)

    # -- hot path -----------------------------------------------------------------------------
    def HoistedForm(self, ct):
        """evaluator.go:543-553"""
        out = {}
        for id in ct.ids():
            out[id] = SwitchingKey(self.ctx)
            self.ksw.Decompose(ct.Level(), ct.Value[id], out[id])
        return out

    def _nb_rescales(self, scale, level, minScale):
        """the loop of evaluator.go:375-383"""
        Q = self.params.Q
        nb = 0
        while level - nb >= 0 and scale / float(Q[level - nb]) >= minScale / 2:
            scale /= float(Q[level - nb])
            nb += 1
        return nb, scale

    def Rescale(self, ctIn, minScale, ctOut):
        """evaluator.go:359-398; returns the error string instead of a Go error (None = nil)"""
        if minScale <= 0:
            return "cannot Rescale: minScale is 0"
        if ctIn.Scale == 0:
            return "cannot Rescale: ciphertext scale is 0"
        if ctIn.Level() == 0:
            return "cannot Rescale: input Ciphertext already at level 0"
        level = ctIn.Level()
        nb, ctOut.Scale = self._nb_rescales(ctIn.Scale, level, minScale)
        if nb > 0:
            for k in ctOut.Value:
                self.ctx.rescale(level, nb, ctIn.Value[k].h, ctOut.Value[k].h)
        return None

    def MulRelinHoistedNew(self, op0, op1, op0Hoisted, op1Hoisted, rlkSet):
        """evaluator.go:558-581"""
        ctOut = self.newCiphertextBinary(op0, op1)
        level = min(op0.Level(), op1.Level(), ctOut.Level())
        if ctOut.Level() > level:
            self.DropLevel(ctOut, ctOut.Level() - level)
        ctOut.Scale = op0.ScalingFactor() * op1.ScalingFactor()
        self.ksw.MulAndRelinHoisted(op0, op1, op0Hoisted, op1Hoisted, rlkSet, ctOut)
        self.Rescale(ctOut, self.params.Scale(), ctOut)
        return ctOut

    def MulRelinNew(self, op0, op1, rlkSet):
        """evaluator.go:416-443: hoist (once when op0 is op1), MulAndRelinHoisted, Rescale -- ONE device call
        (mkhe_ckks_mul_relin); this is what the reference benchmark times (mkckks_benchmark_test.go:78-82)."""
        if self.params.Alpha() > 1 and op0 is not op1 and op0.Level() != op1.Level():
            # The reference decomposes every operand at ITS OWN level (evaluator.go:431-437).  With alpha > 1 the last digit
            # used at the output level then comes from a full alpha-limb group of the higher operand, not from the partial
            # group a decomposition at the output level would lift: hoist explicitly, per operand (alpha = 1: identical).
            h0, h1 = self.HoistedForm(op0), self.HoistedForm(op1)
            out = self.MulRelinHoistedNew(op0, op1, h0, h1, rlkSet)
            for h in list(h0.values()) + list(h1.values()):
                h.free()
            return out
        ctOut = self.newCiphertextBinary(op0, op1)
        level = ctOut.Level()
        scale = op0.ScalingFactor() * op1.ScalingFactor()
        nb, newScale = (0, scale)
        if scale != 0 and level != 0 and self.params.Scale() > 0:
            nb, newScale = self._nb_rescales(scale, level, self.params.Scale())
        ids0, ids1, idsO = op0.ids(), op1.ids(), ctOut.ids()
        self.ctx.ckks_mul_relin(
            level, nb, op0 is op1, ids0, op0.handles(ids0), ids1, op1.handles(ids1),
            [rlkSet.GetRelinearizationKey(i).Value[0].h for i in ids1],
            [rlkSet.GetRelinearizationKey(i).Value[1].h for i in ids0],
            [rlkSet.GetRelinearizationKey(i).Value[2].h for i in ids0],
            self.params.CRS[-1].h, idsO, ctOut.handles(idsO))
        ctOut.Scale = newScale
        return ctOut

    def MulRelinLimbSharded(self, op0, op1, rlkSet, ctOut=None):
        """MulRelinNew executed by a team of ranks (mkhe_ckks_mul_relin_limbs): this rank computes the limb slots it owns, the
        exchanges are peer stores fused into the producing kernels; every rank passes the same operands and gets the whole
        result.  The context must have joined a team (Context.team_import / team_join_local).  ctOut: a ciphertext made by
        newCiphertextBinary(op0, op1) beforehand (nothing is allocated during the call then), or None."""
        if ctOut is None:
            ctOut = self.newCiphertextBinary(op0, op1)
        level = min(op0.Level(), op1.Level())
        scale = op0.ScalingFactor() * op1.ScalingFactor()
        nb, newScale = (0, scale)
        if scale != 0 and level != 0 and self.params.Scale() > 0:
            nb, newScale = self._nb_rescales(scale, level, self.params.Scale())
        ids0, ids1, idsO = op0.ids(), op1.ids(), ctOut.ids()
        self.ctx.ckks_mul_relin_limbs(
            level, nb, ids0, op0.handles(ids0), ids1, op1.handles(ids1),
            [rlkSet.GetRelinearizationKey(i).Value[0].h for i in ids1],
            [rlkSet.GetRelinearizationKey(i).Value[1].h for i in ids0],
            [rlkSet.GetRelinearizationKey(i).Value[2].h for i in ids0],
            self.params.CRS[-1].h, idsO, ctOut.handles(idsO))
        ctOut.Scale = newScale
        return ctOut

    def _normalize(self, rotidx):
        n2 = self.params.N() // 2
        while rotidx >= n2:
            rotidx -= n2
        while rotidx < 0:
            rotidx += n2
        return rotidx

    def _copy(self, ct):
        out = Ciphertext.new(self.params, ct.IDSet(), ct.Level(), ct.Scale)
        for k in ct.Value:
            self.ctx.poly_copy(out.Value[k].h, ct.Value[k].h)
        return out

    def RotateHoistedNew(self, ct0, rotidx, ct0Hoisted, rkSet):
        """evaluator.go:585-617"""
        rotidx = self._normalize(rotidx)
        if rotidx == 0:
            return self._copy(ct0)
        if rotidx not in self.params.CRS:
            raise RuntimeError("Hoisted rotation only works for precomputed rotation keys")
        ctOut = Ciphertext.new(self.params, ct0.IDSet(), ct0.Level(), ct0.Scale)
        self.ksw.RotateHoisted(ct0, rotidx, ct0Hoisted, rkSet, ctOut)
        return ctOut

    def RotateNew(self, ct0, rotidx, rkSet):
        """evaluator.go:485-525 (power-of-two chaining when rotidx has no CRS entry)"""
        rotidx = self._normalize(rotidx)
        if rotidx == 0:
            return self._copy(ct0)
        ctOut = Ciphertext.new(self.params, ct0.IDSet(), ct0.Level(), ct0.Scale)
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
        """evaluator.go:527-540"""
        ctOut = Ciphertext.new(self.params, ct0.IDSet(), ct0.Level(), ct0.Scale)
        self.ksw.Conjugate(ct0, ckSet, ctOut)
        return ctOut

    def CopyNew(self, ct):
        """Ciphertext.CopyNew (mkckks/elements.go)"""
        return self._copy(ct)

    def DropLevelNew(self, ct0, levels):
        """evaluator.go:98-102"""
        out = self._copy(ct0)
        self.DropLevel(out, levels)
        return out

    def getConstAndScale(self, level, constant):
        """evaluator.go:39-93: a float / complex constant with a fractional part is scaled by q_level"""
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

    def MultByConst(self, ct0, constant, ctOut):
        """evaluator.go:117-198 (ctOut may be ct0): the per-limb multipliers are derived on the library side exactly like the
        reference; only the components of ct0 are written"""
        level = min(ct0.Level(), ctOut.Level())
        cReal, cImag, scale = self.getConstAndScale(level, constant)
        ks = list(ct0.Value)
        self.ctx.ckks_mult_by_const(level, [ct0.Value[k].h for k in ks], [ctOut.Value[k].h for k in ks], cReal, cImag, scale)
        ctOut.Scale = ct0.Scale * scale

    def MulPtxtNew(self, ct, ptValue, ptScale):
        """evaluator.go:465-481; ptValue = ckks.Plaintext.Value as a device Poly (coefficient domain), ptScale its Scale"""
        ctOut = Ciphertext.new(self.params, ct.IDSet(), ct.Level(), ct.Scale * ptScale)
        ks = list(ct.Value)
        self.ctx.ckks_mul_ptxt(ct.Level(), ptValue.h, [ct.Value[k].h for k in ks], [ctOut.Value[k].h for k in ks])
        self.Rescale(ctOut, self.params.Scale(), ctOut)
        return ctOut

    def AddNew(self, op0, op1):
        """evaluator.go:319-330 = newCiphertextBinary + evaluateInPlace (:200-304)"""
        return self._evaluate_new(op0, op1, False)

    def SubNew(self, op0, op1):
        """evaluator.go:332-357: evaluateInPlace with SubLvl, then NegLvl (q - x, unreduced) of the components missing from op0"""
        return self._evaluate_new(op0, op1, True)

    def _evaluate_new(self, op0, op1, sub):
        """evaluateInPlace's third branch (ctOut is a fresh element): the operand with the smaller scale is first multiplied
        by floor(ratio) when that exceeds 1 (into a pool ciphertext, evaluator.go:274-292), components held by one operand
        only are copied (:297-303)"""
        import math
        ctOut = self.newCiphertextBinary(op0, op1)
        level = min(op0.Level(), op1.Level(), ctOut.Level())
        s0, s1 = op0.ScalingFactor(), op1.ScalingFactor()
        t0, t1, pool = op0, op1, None
        if s1 > s0 and math.floor(s1 / s0) > 1:
            pool = t0 = Ciphertext.new(self.params, op0.IDSet(), op0.Level(), ctOut.Scale)
            self.MultByConst(op0, float(math.floor(s1 / s0)), t0)
            t0.Scale = ctOut.Scale
        elif s0 > s1 and math.floor(s0 / s1) > 1:
            pool = t1 = Ciphertext.new(self.params, op1.IDSet(), op1.Level(), ctOut.Scale)
            self.MultByConst(op1, float(math.floor(s0 / s1)), t1)
            t1.Scale = ctOut.Scale
        for k in ctOut.Value:
            in0, in1 = k in op0.Value, k in op1.Value
            o = ctOut.Value[k].h
            if in0 and in1:
                (self.ctx.poly_sub if sub else self.ctx.poly_add)(level, t0.Value[k].h, t1.Value[k].h, o)
            elif in0:
                self.ctx.poly_copy_lvl(level, o, t0.Value[k].h)
            elif sub:
                self.ctx.poly_neg(level, t1.Value[k].h, o)
            else:
                self.ctx.poly_copy_lvl(level, o, t1.Value[k].h)
        if pool is not None:
            pool.free()
        return ctOut
