"""The op sequence of the reference's encrypted CNN inference (cnn/cnn.go:10-96 driven by cnn/cnn_test.go:121-162) lives in
mkhe_kklss_b200/cnn.py, written once against a Go-named evaluator facade so that the oracle and the device run the very same
program (BASELINE config 5).  This file adds the oracle's facade and the seeded synthetic inputs.
Ciphertext CONTENTS are synthetic (uniform limbs; the reference's mnist_test.csv is not shipped and throughput / parity do not
depend on the data); ids, levels, scales, rotation indices and the order of operations are the reference's."""
from mkhe_kklss_b200.cnn import (MODEL, DATA, EXTRA_ROTS, cnn_rotations, convolution, fc1_layer, fc2_layer, hoist_model, infer,  # noqa: F401
                                 inference, DeviceFacade)


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
