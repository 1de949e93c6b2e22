"""Replays a golden dump (format "mkhe-dump-1": go/dump_golden_test.go writes it from the REAL Go reference, tools/replay_dump.py
--make writes the same files from the CPU oracle) through libmkhe_b200.so and / or the oracle and compares every output limb of
MulRelinNew and RotateHoistedNew bit for bit.

    python tools/replay_dump.py DIR                  # device (needs a GPU) and oracle against the dumped outputs
    python tools/replay_dump.py DIR --oracle-only    # CPU only
    python tools/replay_dump.py DIR --make [--logn 12 --parties 2]     # write an oracle-made dump (tool-chain self test)

With a dump made by the Go reference this is the out-of-band pin of "bit-exact with the Go evaluator" (DESIGN.md section 5).
The NTT tables in the dump (lattigo's NttPsi / NttPsiInv / NttNInv, Montgomery form) are handed to the library verbatim
(mkhe_ctx_set_ntt_tables) and compared with the tables the library and the oracle generate themselves."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _rd(d, name, shape):
    a = np.fromfile(os.path.join(d, name), dtype="<u8")
    return a.reshape(shape)


def _wr(d, name, arr):
    np.ascontiguousarray(arr, dtype="<u8").tofile(os.path.join(d, name))


def load(d):
    meta = json.load(open(os.path.join(d, "meta.json")))
    assert meta["format"] == "mkhe-dump-1", meta.get("format")
    N, nQ, nP, k, L = 1 << meta["logN"], len(meta["Q"]), len(meta["P"]), len(meta["ids"]), meta["level"]
    D = nQ + nP
    swk = lambda name: _rd(d, name, (nQ, D, N))
    ct = lambda prefix, lv: {**{"0": _rd(d, f"{prefix}_c0.bin", (lv + 1, N))}, **{i: _rd(d, f"{prefix}_p{i}.bin", (lv + 1, N)) for i in range(k)}}
    x = {"meta": meta, "N": N, "k": k, "L": L,
         "psi": [_rd(d, f"psi_{m}.bin", (N,)) for m in range(D)], "psiinv": [_rd(d, f"psiinv_{m}.bin", (N,)) for m in range(D)],
         "ninv": _rd(d, "ninv.bin", (D,)), "u": swk("crs_u.bin"), "a": swk("crs_rot.bin"),
         "rlk": [(swk(f"rlk_{i}_b.bin"), swk(f"rlk_{i}_d.bin"), swk(f"rlk_{i}_v.bin")) for i in range(k)],
         "rk": [swk(f"rk_{i}.bin") for i in range(k)], "ct0": ct("ct0", L), "ct1": ct("ct1", L),
         "mul": ct("mul", meta["mul_level"]), "rot": ct("rot", L)}
    return x


def make(d, logn, parties, seed=0xB2000007):
    """an oracle-made dump in the same format (self test of the tool chain; NOT a pin of the reference)"""
    from oracle import oracle as O
    from mkhe_kklss_b200 import params as PR
    import parity
    lit = PR.CKKS_PN15QP880.at_logn(logn)
    os.makedirs(d, exist_ok=True)
    rot = 2
    p = O.MKParams(lit.logN, lit.Q, lit.P, 2, seed=seed, crs_rots=[rot])
    prng = O.PRNG(seed ^ 0x5EED)
    sw = lambda: parity.uniform_swk(prng, p)
    ids = list(range(parties))
    rlk = {i: O.RelinKey(i, sw(), sw(), sw()) for i in ids}
    rk = {i: {rot: sw()} for i in ids}
    L = p.max_level()
    mk = lambda: O.Ciphertext({**{"0": prng.uniform(p.ringQ, L)}, **{i: prng.uniform(p.ringQ, L) for i in ids}}, lit.scale)
    c0, c1 = mk(), mk()
    ev = O.CKKSEvaluator(p, lit.scale)
    mul = ev.mul_relin_new(c0, c1, rlk)
    rt = ev.rotate_hoisted_new(c0, rot, ev.hoisted_form(c0), rk)
    for m, ring_i in [(m, (p.ringQ, m)) for m in range(len(lit.Q))] + [(len(lit.Q) + j, (p.ringP, j)) for j in range(len(lit.P))]:
        psi, psiinv, _ = ring_i[0].tables(ring_i[1])
        _wr(d, f"psi_{m}.bin", psi)
        _wr(d, f"psiinv_{m}.bin", psiinv)
    _wr(d, "ninv.bin", np.array([p.ringQ.tables(i)[2] for i in range(len(lit.Q))] + [p.ringP.tables(j)[2] for j in range(len(lit.P))], dtype=np.uint64))
    _wr(d, "crs_u.bin", p.CRS[-1])
    _wr(d, "crs_rot.bin", p.CRS[rot])
    for i in ids:
        _wr(d, f"rlk_{i}_b.bin", rlk[i].b); _wr(d, f"rlk_{i}_d.bin", rlk[i].d); _wr(d, f"rlk_{i}_v.bin", rlk[i].v)
        _wr(d, f"rk_{i}.bin", rk[i][rot])
    for prefix, c in (("ct0", c0), ("ct1", c1), ("mul", mul), ("rot", rt)):
        _wr(d, f"{prefix}_c0.bin", c.value["0"])        # component "0"; party t of the sorted id list -> _p<t>
        for i in ids:
            _wr(d, f"{prefix}_p{i}.bin", c.value[i])
    meta = {"format": "mkhe-dump-1", "params": lit.name, "logN": lit.logN, "Q": [int(q) for q in lit.Q], "P": [int(q) for q in lit.P],
            "gamma": 2, "scale": lit.scale, "ids": [f"user{i}" for i in ids], "level": L, "rot": rot, "mul_level": mul.level(),
            "mul_scale": mul.scale, "source": "oracle (tool-chain self test, not the Go reference)"}
    json.dump(meta, open(os.path.join(d, "meta.json"), "w"), indent=1)


# ---- mkbfv dumps (meta["scheme"] == "bfv": go/dump_golden_bfv_test.go) -------------------------------------------------------
def load_bfv(d):
    meta = json.load(open(os.path.join(d, "meta.json")))
    N, nQ, nP, k, L = 1 << meta["logN"], len(meta["Q"]), len(meta["P"]), len(meta["ids"]), meta["level"]
    D, M = nQ + nP, 2 * nQ + nP
    swk = lambda name: _rd(d, name, (nQ, D, N))
    ct = lambda prefix: {**{"0": _rd(d, f"{prefix}_c0.bin", (L + 1, N))}, **{i: _rd(d, f"{prefix}_p{i}.bin", (L + 1, N)) for i in range(k)}}
    return {"meta": meta, "N": N, "k": k, "L": L,
            "psi": [_rd(d, f"psi_{m}.bin", (N,)) for m in range(M)], "psiinv": [_rd(d, f"psiinv_{m}.bin", (N,)) for m in range(M)],
            "ninv": _rd(d, "ninv.bin", (M,)), "u": swk("crs_u.bin"),
            "rlk": [{n: swk(f"rlk_{i}_{n}.bin") for n in ("b1", "d1", "v", "b2", "d2")} for i in range(k)],
            "ct0": ct("ct0"), "ct1": ct("ct1"), "mul": ct("mul")}


def make_bfv(d, logn, parties, seed=0xB2000017):
    """an oracle-made mkbfv dump in the same format (self test of the tool chain; NOT a pin of the reference)"""
    from oracle import oracle as O
    from mkhe_kklss_b200 import params as PR
    import parity
    lit = PR.BFV_PN14QP439.at_logn(logn)
    os.makedirs(d, exist_ok=True)
    p = O.BFVParams(lit.logN, lit.Q, lit.QMul, lit.P, lit.T, seed=seed)
    prng = O.PRNG(seed ^ 0x5EED)
    sw = lambda: parity.uniform_swk(prng, p)
    ids = list(range(parties))
    rlk = {i: O.BFVRelinKey(i, sw(), sw(), sw(), sw(), sw()) for i in ids}
    L = p.max_level()
    mk = lambda: O.Ciphertext({**{"0": prng.uniform(p.ringQ, L)}, **{i: prng.uniform(p.ringQ, L) for i in ids}})
    c0, c1 = mk(), mk()
    mul = O.BFVEvaluator(p).mul_relin_new(c0, c1, rlk)
    rings = [(p.ringQ, i) for i in range(len(lit.Q))] + [(p.ringP, j) for j in range(len(lit.P))] + [(p.ringQMul, i) for i in range(len(lit.QMul))]
    ninv = []
    for m, (ring, i) in enumerate(rings):
        psi, psiinv, ni = ring.tables(i)
        _wr(d, f"psi_{m}.bin", psi)
        _wr(d, f"psiinv_{m}.bin", psiinv)
        ninv.append(ni)
    _wr(d, "ninv.bin", np.array(ninv, dtype=np.uint64))
    _wr(d, "crs_u.bin", p.CRS[-1])
    for i in ids:
        for n in ("b1", "d1", "v", "b2", "d2"):
            _wr(d, f"rlk_{i}_{n}.bin", getattr(rlk[i], n))
    for prefix, c in (("ct0", c0), ("ct1", c1), ("mul", mul)):
        _wr(d, f"{prefix}_c0.bin", c.value["0"])
        for i in ids:
            _wr(d, f"{prefix}_p{i}.bin", c.value[i])
    meta = {"format": "mkhe-dump-1", "scheme": "bfv", "params": lit.name, "logN": lit.logN, "Q": [int(q) for q in lit.Q],
            "P": [int(q) for q in lit.P], "QMul": [int(q) for q in lit.QMul], "T": int(lit.T), "gamma": 2,
            "ids": [f"user{i}" for i in ids], "level": L, "source": "oracle (tool-chain self test, not the Go reference)"}
    json.dump(meta, open(os.path.join(d, "meta.json"), "w"), indent=1)


def replay_bfv_oracle(x, report):
    from oracle import oracle as O
    meta, k = x["meta"], x["k"]
    p = O.BFVParams(meta["logN"], meta["Q"], meta["QMul"], meta["P"], meta["T"], seed=1)
    nQ, nP = len(meta["Q"]), len(meta["P"])
    same = True
    for m in range(2 * nQ + nP):
        ring, i = (p.ringQ, m) if m < nQ else ((p.ringP, m - nQ) if m < nQ + nP else (p.ringQMul, m - nQ - nP))
        psi, psiinv, ninv = ring.tables(i)
        same = same and np.array_equal(psi, x["psi"][m]) and np.array_equal(psiinv, x["psiinv"][m]) and int(ninv) == int(x["ninv"][m])
    report.append(("oracle NTT tables (Q, P, QMul) == dumped lattigo tables", same, []))
    p.CRS[-1] = x["u"]
    rlk = {i: O.BFVRelinKey(i, x["rlk"][i]["b1"], x["rlk"][i]["d1"], x["rlk"][i]["v"], x["rlk"][i]["b2"], x["rlk"][i]["d2"]) for i in range(k)}
    c0 = O.Ciphertext({kk: v.copy() for kk, v in x["ct0"].items()})
    c1 = O.Ciphertext({kk: v.copy() for kk, v in x["ct1"].items()})
    _cmp(O.BFVEvaluator(p).mul_relin_new(c0, c1, rlk).value, x["mul"], "oracle mkbfv MulRelinNew", report)


def replay_bfv_device(x, report, lib=None, use_dumped_tables=True):
    from mkhe_kklss_b200 import mkbfv, mkrlwe
    meta, k = x["meta"], x["k"]
    dp = mkbfv.Parameters(meta["logN"], meta["Q"], meta["QMul"], meta["P"], meta["T"], lib=lib)
    if use_dumped_tables:
        for m in range(2 * len(meta["Q"]) + len(meta["P"])):
            dp.ctx.set_ntt_tables(m, x["psi"][m], x["psiinv"][m], int(x["ninv"][m]))
    dp.SetCRS(-1, x["u"])
    rl = mkbfv.RelinearizationKeySet()
    for i in range(k):
        r = x["rlk"][i]
        rl.AddRelinearizationKey(mkbfv.RelinearizationKey(dp.ctx, i, r["b1"], r["d1"], r["v"], r["b2"], r["d2"]))
    c0, c1 = mkrlwe.Ciphertext.from_numpy(dp.ctx, x["ct0"]), mkrlwe.Ciphertext.from_numpy(dp.ctx, x["ct1"])
    tag = "device" + (" (dumped tables)" if use_dumped_tables else " (own tables)")
    _cmp(mkbfv.Evaluator(dp).MulRelinNew(c0, c1, rl).numpy(), x["mul"], f"{tag} mkbfv MulRelinNew", report)
    dp.ctx.close()


def _cmp(got, want, what, report):
    ok = set(got) == set(want) and all(np.array_equal(got[k], want[k]) for k in want)
    bad = [str(k) for k in want if k not in got or not np.array_equal(got[k], want[k])]
    report.append((what, ok, bad))
    return ok


def replay_oracle(x, report):
    from oracle import oracle as O
    meta, k, rot = x["meta"], x["k"], x["meta"]["rot"]
    p = O.MKParams(meta["logN"], meta["Q"], meta["P"], meta["gamma"], seed=1, crs_rots=[rot])
    # the oracle generates its own tables (lattigo's primitiveRoot rule): they must be the dumped ones
    nQ = len(meta["Q"])
    same = True
    for m in range(nQ + len(meta["P"])):
        ring, i = (p.ringQ, m) if m < nQ else (p.ringP, m - nQ)
        psi, psiinv, ninv = ring.tables(i)
        same = same and np.array_equal(psi, x["psi"][m]) and np.array_equal(psiinv, x["psiinv"][m]) and int(ninv) == int(x["ninv"][m])
    report.append(("oracle NTT tables == dumped lattigo tables", same, []))
    p.CRS[-1], p.CRS[rot] = x["u"], x["a"]
    rlk = {i: O.RelinKey(i, *x["rlk"][i]) for i in range(k)}
    rk = {i: {rot: x["rk"][i]} for i in range(k)}
    ev = O.CKKSEvaluator(p, meta["scale"])
    c0 = O.Ciphertext({kk: v.copy() for kk, v in x["ct0"].items()}, meta["scale"])
    c1 = O.Ciphertext({kk: v.copy() for kk, v in x["ct1"].items()}, meta["scale"])
    mul = ev.mul_relin_new(c0, c1, rlk)
    _cmp(mul.value, x["mul"], "oracle MulRelinNew", report)
    report.append(("oracle MulRelinNew scale", mul.scale == meta["mul_scale"], []))
    rt = ev.rotate_hoisted_new(c0, rot, ev.hoisted_form(c0), rk)
    _cmp(rt.value, x["rot"], "oracle RotateHoistedNew", report)


def replay_device(x, report, lib=None, use_dumped_tables=True):
    from mkhe_kklss_b200 import mkckks, mkrlwe
    meta, k, rot = x["meta"], x["k"], x["meta"]["rot"]
    dp = mkckks.Parameters(meta["logN"], meta["Q"], meta["P"], meta["scale"], lib=lib, gamma=meta["gamma"])
    if use_dumped_tables:
        for m in range(len(meta["Q"]) + len(meta["P"])):
            dp.ctx.set_ntt_tables(m, x["psi"][m], x["psiinv"][m], int(x["ninv"][m]))
    dp.SetCRS(-1, x["u"])
    dp.SetCRS(rot, x["a"])
    rl, rks = mkrlwe.RelinearizationKeySet(), mkrlwe.RotationKeySet()
    for i in range(k):
        b, d, v = x["rlk"][i]
        rl.AddRelinearizationKey(mkrlwe.RelinearizationKey(dp.ctx, i, b, d, v))
        rks.AddRotationKey(i, rot, mkrlwe.SwitchingKey(dp.ctx, x["rk"][i]))
    ev = mkckks.Evaluator(dp)
    c0 = mkckks.Ciphertext.from_numpy(dp.ctx, x["ct0"], meta["scale"])
    c1 = mkckks.Ciphertext.from_numpy(dp.ctx, x["ct1"], meta["scale"])
    tag = "device" + (" (dumped tables)" if use_dumped_tables else " (own tables)")
    mul = ev.MulRelinNew(c0, c1, rl)
    _cmp(mul.numpy(), x["mul"], f"{tag} MulRelinNew", report)
    report.append((f"{tag} MulRelinNew scale", mul.Scale == meta["mul_scale"], []))
    rt = ev.RotateHoistedNew(c0, rot, ev.HoistedForm(c0), rks)
    _cmp(rt.numpy(), x["rot"], f"{tag} RotateHoistedNew", report)
    dp.ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dir")
    ap.add_argument("--make", action="store_true")
    ap.add_argument("--logn", type=int, default=12)
    ap.add_argument("--parties", type=int, default=2)
    ap.add_argument("--oracle-only", action="store_true")
    ap.add_argument("--device-only", action="store_true")
    ap.add_argument("--bfv", action="store_true", help="with --make: write an mkbfv dump")
    ap.add_argument("--lib", default=None, help="development: another build of the library")
    args = ap.parse_args()
    if args.make:
        (make_bfv if args.bfv else make)(args.dir, args.logn, args.parties)
        print("wrote", args.dir)
        return 0
    lib = None
    if args.lib:
        from mkhe_kklss_b200 import _lib
        lib = _lib.Library(os.path.abspath(args.lib))
    bfv = json.load(open(os.path.join(args.dir, "meta.json"))).get("scheme") == "bfv"
    x = load_bfv(args.dir) if bfv else load(args.dir)
    report = []
    if not args.device_only:
        (replay_bfv_oracle if bfv else replay_oracle)(x, report)
    if not args.oracle_only:
        (replay_bfv_device if bfv else replay_device)(x, report, lib=lib, use_dumped_tables=True)
        (replay_bfv_device if bfv else replay_device)(x, report, lib=lib, use_dumped_tables=False)
    print("dump source:", x["meta"].get("source"))
    for what, ok, bad in report:
        print(("PASS " if ok else "FAIL ") + what + ("" if ok else f"  mismatching components: {bad}"))
    return 0 if all(ok for _, ok, _ in report) else 1


if __name__ == "__main__":
    sys.exit(main())
