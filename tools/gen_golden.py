"""Generates tests/golden/oracle_digests.json: SHA-256 digests of the oracle's outputs on seeded inputs.

The reference cannot run in this environment (Go + un-vendored lattigo v2.3.0), so these fixtures do NOT come
from the reference; they freeze the oracle restatement so that any later change to it is caught
(tests/test_oracle.py::test_golden_fixtures) and give the GPU tests a second, committed comparison point.

    python tools/gen_golden.py            # rewrite the fixture
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def _digest(arr):
    return hashlib.sha256(np.ascontiguousarray(arr, dtype=np.uint64).tobytes()).hexdigest()


def golden_cases():
    """(name, callable returning dict component -> array); inputs are derived from fixed seeds"""
    from oracle import oracle as O
    from mkhe_kklss_b200 import params as PR
    import parity

    def ckks_case(lit, k, same):
        p = O.MKParams(lit.logN, lit.Q, lit.P, 2, seed=0xB2000001, crs_rots=[1, 2])
        prng = O.PRNG(0xB2000001 ^ 0x5EED)
        rl = {i: O.RelinKey(i, parity.uniform_swk(prng, p), parity.uniform_swk(prng, p), parity.uniform_swk(prng, p)) for i in range(k)}
        rk = {i: {2: parity.uniform_swk(prng, p)} for i in range(k)}
        ev = O.CKKSEvaluator(p, lit.scale)
        L = p.max_level()
        mk = lambda: O.Ciphertext({**{"0": prng.uniform(p.ringQ, L)}, **{i: prng.uniform(p.ringQ, L) for i in range(k)}}, lit.scale)
        c0 = mk()
        c1 = c0 if same else mk()
        out = ev.mul_relin_new(c0, c1, rl)
        rot = ev.rotate_hoisted_new(c0, 2, ev.hoisted_form(c0), rk)
        res = {f"mul[{kk}]": v for kk, v in out.value.items()}
        res.update({f"rot[{kk}]": v for kk, v in rot.value.items()})
        return res

    def bfv_case(lit, k):
        p = O.BFVParams(lit.logN, lit.Q, lit.QMul, lit.P, lit.T, seed=0xB2000003)
        prng = O.PRNG(0xB2000003 ^ 0xBF5EED)
        u = lambda: parity.uniform_swk(prng, p)
        rl = {i: O.BFVRelinKey(i, u(), u(), u(), u(), u()) for i in range(k)}
        ev = O.BFVEvaluator(p)
        L = p.max_level()
        mk = lambda: O.Ciphertext({**{"0": prng.uniform(p.ringQ, L)}, **{i: prng.uniform(p.ringQ, L) for i in range(k)}})
        c0, c1 = mk(), mk()
        out = ev.mul_relin_new(c0, c1, rl)
        res = {f"mul[{kk}]": v for kk, v in out.value.items()}
        res["modup"] = ev.modup_q_to_r(c0.value["0"])
        res["rescale"] = ev.rescale_q_to_r(c0.value["0"])
        return res

    def cnn_flow_case(lit):
        import cnn_flow as F
        out, mid = F.run_oracle(lit)
        res = {f"fc2Out[{kk}]": v for kk, v in out.value.items()}
        for name, ct in mid.items():
            res.update({f"{name}[{kk}]": v for kk, v in ct.value.items()})
        return res

    def elementwise_case(lit):
        p = O.MKParams(lit.logN, lit.Q, lit.P, 2, seed=0xB2000006, crs_rots=[])
        prng = O.PRNG(0xB2000006 ^ 0x5EED)
        ev = O.CKKSEvaluator(p, lit.scale)
        L = p.max_level()
        mk = lambda ids, lvl, sc: O.Ciphertext({**{"0": prng.uniform(p.ringQ, lvl)}, **{i: prng.uniform(p.ringQ, lvl) for i in ids}}, sc)
        a, b, c = mk([0, 1], L, lit.scale), mk([1, 2], L - 1, lit.scale * 7.3), mk([0], L, lit.scale)
        res = {}
        for nm, ct in (("add", ev.add_new(a, b)), ("sub", ev.sub_new(a, b)), ("sub2", ev.sub_new(c, b))):
            res.update({f"{nm}[{kk}]": v for kk, v in ct.value.items()})
        for n, const in enumerate((3, -2.5, complex(0.5, -1.25))):
            o = ev.new_ciphertext([0, 1], L, 0.0)
            ev.mult_by_const(a, const, o)
            res.update({f"const{n}[{kk}]": v for kk, v in o.value.items()})
        pt = prng.uniform(p.ringQ, L)
        res.update({f"mulptxt[{kk}]": v for kk, v in ev.mul_ptxt_new(a, pt, lit.scale).value.items()})
        return res

    return [
        ("cnn_flow_PN14QP433_logN12", lambda: cnn_flow_case(PR.CNN_PN14QP433.at_logn(12))),
        ("elementwise_PN14QP439_logN12", lambda: elementwise_case(PR.CKKS_PN14QP439.at_logn(12))),
        ("ckks_PN14QP439_logN12_k2", lambda: ckks_case(PR.CKKS_PN14QP439.at_logn(12), 2, False)),
        ("ckks_PN14QP439_logN12_k2_square", lambda: ckks_case(PR.CKKS_PN14QP439.at_logn(12), 2, True)),
        ("ckks_PN15QP880_logN12_k3", lambda: ckks_case(PR.CKKS_PN15QP880.at_logn(12), 3, False)),
        ("cnn_PN14QP433_logN12_k2", lambda: ckks_case(PR.CNN_PN14QP433.at_logn(12), 2, False)),
        ("bfv_PN14QP439_logN12_k2", lambda: bfv_case(PR.BFV_PN14QP439.at_logn(12), 2)),
    ]


def compute_digests():
    out = {}
    for name, fn in golden_cases():
        out[name] = {k: _digest(v) for k, v in sorted(fn().items())}
    return out


if __name__ == "__main__":
    path = os.path.join(ROOT, "tests", "golden", "oracle_digests.json")
    with open(path, "w") as f:
        json.dump(compute_digests(), f, indent=1, sort_keys=True)
    print("wrote", path)
