"""The op sequence of the reference's encrypted CNN inference (cnn/cnn.go:10-96 driven by cnn/cnn_test.go:121-162), written once
against a Go-named evaluator facade so that the oracle and the device run the very same program (BASELINE config 5).
Ciphertext CONTENTS are synthetic (uniform limbs; the reference's mnist_test.csv is not shipped and throughput / parity do not
depend on the data); ids, levels, scales, rotation indices and the order of operations are the reference's."""
import math

MODEL, DATA = 0, 1                       # modelOwner / dataOwner (cnn/cnn_test.go:35-36)
# genTestParams (cnn/cnn_test.go:185-201): extra CRS / rotation keys, plus every power of two below N/2
EXTRA_ROTS = [14, 15, 384, 512, 640, 768, 896, 8191, 8190, 8188, 8184]


def cnn_rotations(logN):
    return sorted(set(EXTRA_ROTS + [1 << i for i in range(logN - 1)]))


def convolution(E, ctImage, ctImageHoisted, ctKernels, ctKernelsHoisted):
    """Convolution, cnn/cnn.go:10-40"""
    convOut = E.MulRelinHoistedNew(ctImage, ctKernels[0], ctImageHoisted, ctKernelsHoisted[0])
    for j, rot in ((1, 1), (2, 14), (3, 15)):
        temp = E.RotateHoistedNew(ctImage, rot, ctImageHoisted)
        tempHoisted = E.HoistedForm(temp)
        temp = E.MulRelinHoistedNew(temp, ctKernels[j], tempHoisted, ctKernelsHoisted[j])
        convOut = E.AddNew(convOut, temp)
    for rot in (2048, 1024):
        temp = E.RotateNew(convOut, rot)
        convOut = E.AddNew(convOut, temp)
    return convOut


def fc1_layer(E, ctVec, ctVecHoisted, ctMat, ctMatHoisted, ctBias):
    """FC1Layer, cnn/cnn.go:42-71"""
    fc1Out = None
    for i in range(len(ctMat)):
        temp = E.RotateHoistedNew(ctVec, i * 128, ctVecHoisted)
        tempHoisted = E.HoistedForm(temp)
        temp = E.MulRelinHoistedNew(temp, ctMat[i], tempHoisted, ctMatHoisted[i])
        fc1Out = E.CopyNew(temp) if i == 0 else E.AddNew(fc1Out, temp)
    for i in range(int(math.log2(128))):
        temp = E.RotateNew(fc1Out, 1 << i)
        fc1Out = E.AddNew(fc1Out, temp)
    return E.AddNew(fc1Out, ctBias)


def fc2_layer(E, ctVec, ctMat, ctBias, ptMask, ptScale):
    """FC2Layer, cnn/cnn.go:73-96"""
    fc2Out = E.MulPtxtNew(ctVec, ptMask, ptScale)
    for i in range(int(math.log2(16))):
        temp = E.RotateNew(fc2Out, -1 * (1 << i))
        fc2Out = E.AddNew(fc2Out, temp)
    fc2Out = E.MulRelinNew(fc2Out, ctMat)
    for i in range(int(math.log2(64))):
        temp = E.RotateNew(fc2Out, 128 * (1 << i))
        fc2Out = E.AddNew(fc2Out, temp)
    return E.AddNew(fc2Out, ctBias)


def inference(E, ctImage, ctKernels, ctFC1, ctFC2, ctB1, ctB2, ptMask, ptScale):
    """TestCNN's evaluation, cnn/cnn_test.go:121-162"""
    ctImageHoisted = E.HoistedForm(ctImage)
    ctKernelsHoisted = [E.HoistedForm(c) for c in ctKernels]
    ctFC1Hoisted = [E.HoistedForm(c) for c in ctFC1]
    convOut = convolution(E, ctImage, ctImageHoisted, ctKernels, ctKernelsHoisted)
    convOutHoisted = E.HoistedForm(convOut)
    square1Out = E.MulRelinHoistedNew(convOut, convOut, convOutHoisted, convOutHoisted)
    square1OutHoisted = E.HoistedForm(square1Out)
    fc1Out = fc1_layer(E, square1Out, square1OutHoisted, ctFC1, ctFC1Hoisted, ctB1)
    fc1OutHoisted = E.HoistedForm(fc1Out)
    square2Out = E.MulRelinHoistedNew(fc1Out, fc1Out, fc1OutHoisted, fc1OutHoisted)
    return fc2_layer(E, square2Out, ctFC2, ctB2, ptMask, ptScale), {"conv": convOut, "square1": square1Out, "fc1": fc1Out}


class DeviceFacade:
    """mkckks.Evaluator of this package with the key sets bound, Go method names"""

    def __init__(self, ev, rlk, rk):
        self.ev, self.rlk, self.rk = ev, rlk, rk
        self.counts = {}

    def _c(self, n):
        self.counts[n] = self.counts.get(n, 0) + 1

    def HoistedForm(self, ct):
        self._c("HoistedForm"); return self.ev.HoistedForm(ct)

    def MulRelinHoistedNew(self, a, b, ha, hb):
        self._c("MulRelinHoistedNew"); return self.ev.MulRelinHoistedNew(a, b, ha, hb, self.rlk)

    def MulRelinNew(self, a, b):
        self._c("MulRelinNew"); return self.ev.MulRelinNew(a, b, self.rlk)

    def RotateHoistedNew(self, ct, rot, h):
        self._c("RotateHoistedNew"); return self.ev.RotateHoistedNew(ct, rot, h, self.rk)

    def RotateNew(self, ct, rot):
        self._c("RotateNew"); return self.ev.RotateNew(ct, rot, self.rk)

    def AddNew(self, a, b):
        self._c("AddNew"); return self.ev.AddNew(a, b)

    def CopyNew(self, ct):
        return self.ev.CopyNew(ct)

    def MulPtxtNew(self, ct, pt, scale):
        self._c("MulPtxtNew"); return self.ev.MulPtxtNew(ct, pt, scale)


class OracleFacade:
    """the oracle's CKKSEvaluator behind the same names"""

    def __init__(self, ev, rlk, rk):
        self.ev, self.rlk, self.rk = ev, rlk, rk

    def HoistedForm(self, ct):
        return self.ev.hoisted_form(ct)

    def MulRelinHoistedNew(self, a, b, ha, hb):
        return self.ev.mul_relin_hoisted_new(a, b, ha, hb, self.rlk)

    def MulRelinNew(self, a, b):
        return self.ev.mul_relin_new(a, b, self.rlk)

    def RotateHoistedNew(self, ct, rot, h):
        return self.ev.rotate_hoisted_new(ct, rot, h, self.rk)

    def RotateNew(self, ct, rot):
        return self.ev.rotate_new(ct, rot, self.rk)

    def AddNew(self, a, b):
        return self.ev.add_new(a, b)

    def CopyNew(self, ct):
        return ct.copy()

    def MulPtxtNew(self, ct, pt, scale):
        return self.ev.mul_ptxt_new(ct, pt, scale)


def make_inputs(lit, seed=0xB2000005):
    """seeded synthetic inputs of the flow on the oracle side: parameters with the CNN's CRS entries, uniform relinearisation and
    rotation keys of both parties, image / kernel / weight / bias ciphertexts at the top level, the mask plaintext"""
    from oracle import oracle as O
    import parity
    rots = [r for r in cnn_rotations(lit.logN) if r < lit.N // 2]
    op = O.MKParams(lit.logN, lit.Q, lit.P, lit.gamma, seed=seed, crs_rots=rots)
    prng = O.PRNG(seed ^ 0x5EED)
    sw = lambda: parity.uniform_swk(prng, op)
    rlk = {i: O.RelinKey(i, sw(), sw(), sw()) for i in (MODEL, DATA)}
    rk = {i: {r: sw() for r in rots} for i in (MODEL, DATA)}
    L = op.max_level()

    def ct(ids):
        v = {"0": prng.uniform(op.ringQ, L)}
        for i in ids:
            v[i] = prng.uniform(op.ringQ, L)
        return O.Ciphertext(v, lit.scale)

    return {"op": op, "rlk": rlk, "rk": rk, "rots": rots, "image": ct([DATA]), "kernels": [ct([MODEL]) for _ in range(4)],
            "fc1": [ct([MODEL]) for _ in range(8)], "fc2": ct([MODEL]), "b1": ct([MODEL]), "b2": ct([MODEL]),
            "mask": prng.uniform(op.ringQ, L)}


def run_oracle(lit, inp=None):
    from oracle import oracle as O
    inp = inp or make_inputs(lit)
    E = OracleFacade(O.CKKSEvaluator(inp["op"], lit.scale), inp["rlk"], inp["rk"])
    return inference(E, inp["image"], inp["kernels"], inp["fc1"], inp["fc2"], inp["b1"], inp["b2"], inp["mask"], lit.scale)


def run_device(lit, inp, lib=None):
    """the same inputs uploaded through the host mirror; returns (result, intermediates, facade, context)"""
    from mkhe_kklss_b200 import mkckks, mkrlwe
    dp = mkckks.Parameters(lit.logN, lit.Q, lit.P, lit.scale, lib=lib, gamma=lit.gamma)
    for idx, arr in inp["op"].CRS.items():
        dp.SetCRS(idx, arr)
    rl, rk = mkrlwe.RelinearizationKeySet(), mkrlwe.RotationKeySet()
    for i in (MODEL, DATA):
        k = inp["rlk"][i]
        rl.AddRelinearizationKey(mkrlwe.RelinearizationKey(dp.ctx, i, k.b, k.d, k.v))
        for r, a in inp["rk"][i].items():
            rk.AddRotationKey(i, r, mkrlwe.SwitchingKey(dp.ctx, a))
    up = lambda c: mkckks.Ciphertext.from_numpy(dp.ctx, c.value, c.scale)
    E = DeviceFacade(mkckks.Evaluator(dp), rl, rk)
    out, mid = inference(E, up(inp["image"]), [up(c) for c in inp["kernels"]], [up(c) for c in inp["fc1"]], up(inp["fc2"]),
                         up(inp["b1"]), up(inp["b2"]), mkrlwe.Poly.from_numpy(dp.ctx, inp["mask"]), lit.scale)
    return out, mid, E, dp.ctx
