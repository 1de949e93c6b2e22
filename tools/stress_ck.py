#!/usr/bin/env python
"""GPU race hunt with per-launch checksums: two independent contexts run identical MulRelinNew calls concurrently; the SECOND
context (MKHE_DEBUG_CK=2) checksums every scratch buffer after every kernel launch.  All ops are identical, so the checksum rows of
every op must be identical: the first row that differs names the kernel after which a buffer first deviates."""
import ctypes as C
import os, sys
NCTX = int(os.environ.get("MKHE_STRESS_CTXS", "2"))      # 1: a single context, single stream (the plain back-to-back case)
os.environ["MKHE_DEBUG_CK"] = str(NCTX)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity
from mkhe_kklss_b200 import params as PR, mkckks

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = 8
ws = [parity.CKKSWorld(PR.CKKS_PN15QP880, 2, rots=(2,)) for _ in range(NCTX)]
st = []
for w in ws:
    ids, level = w.ids, w.op.max_level()
    o0, d0 = w.random_ct(ids, level)
    o1, d1 = w.random_ct(ids, level)
    want = w.oev.mul_relin_new(o0, o1, w.o_rlk)
    outs = [mkckks.Ciphertext.new(w.dp, ids, level, w.lit.scale) for _ in range(B)]
    g = w.d_rlk.GetRelinearizationKey
    kb, kd, kv = ([g(i).Value[j].h for i in ids] for j in range(3))
    nb, _ = w.dev._nb_rescales(w.lit.scale * w.lit.scale, level, w.lit.scale)
    st.append((w, ids, level, d0, d1, want, outs, kb, kd, kv, nb))


def fetch(w):
    buf = (C.c_uint64 * (4096 * 16))()
    n = C.c_int(0)
    names = C.create_string_buffer(1 << 20)
    w.ctx.check(w.ctx.dll.mkhe_debug_ck_fetch(w.ctx.ptr, buf, 4096, C.byref(n), names, C.c_size_t(1 << 20)))
    rows = np.array(buf[:16 * n.value], dtype=np.uint64).reshape(-1, 16)
    return rows, names.value.decode().split("\n")[:n.value]


def run(op_fn):
    for i in range(B):
        for s in st:
            op_fn(s, i)
    for s in st:
        s[0].ctx.sync()


def mul(s, i):
    w, ids, level, d0, d1, want, outs, kb, kd, kv, nb = s
    w.ctx.ckks_mul_relin(level, nb, False, ids, d0.handles(ids), ids, d1.handles(ids), kb, kd, kv, w.dp.CRS[-1].h, ids, outs[i].handles(ids))


def snap(w, row):
    nbytes = C.c_size_t(0)
    buf = np.empty(8 << 20, dtype=np.uint64)
    w.ctx.check(w.ctx.dll.mkhe_debug_snap_read(w.ctx.ptr, C.c_int(row), buf.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_size_t(buf.nbytes), C.byref(nbytes)))
    return buf[:nbytes.value // 8].copy()


run(mul)                       # warm-up: scratch pools get allocated
fetch(ws[-1])
if os.environ.get("MKHE_DEBUG_SNAP"):
    N = 1 << 15
    for r in range(rounds):
        run(mul)
        rows, names = fetch(ws[-1])
        per = len(names) // B
        w, ids, level, d0, d1, want, outs, kb, kd, kv, nb = st[-1]
        ok = [all(np.array_equal(w.ctx.poly_download(outs[i].Value[key].h, level + 1 - nb), want.value[key]) for key in ["0"] + ids) for i in range(B)]
        if all(ok) or not any(ok):
            continue
        gi = ok.index(True)
        for i in range(B):
            if ok[i]:
                continue
            for j in range(per):
                if i * per + j >= 64 or gi * per + j >= 64:
                    break
                got, good = snap(w, i * per + j), snap(w, gi * per + j)
                d = np.flatnonzero(got != good)
                if len(d):
                    unit = 17 * N if os.environ["MKHE_DEBUG_SNAP"] == "accqp" else 14 * N
                    prod, rem = d // unit, d % unit
                    slot, pos = rem // N, rem % N
                    print(f"round {r} op {i} (good op {gi}): FIRST differing snapshot = #{j} after {names[i * per + j]}: {len(d)} words; products {np.unique(prod).tolist()} "
                          f"slots {np.unique(slot).tolist()} tiles {np.unique(pos // 2048).tolist()} pos {pos.min()}..{pos.max()}", flush=True)
                    break
            else:
                print(f"round {r} op {i}: output WRONG but no snapshot differs", flush=True)
    print("done")
    sys.exit(0)
ref = None
for r in range(rounds):
    run(mul)
    rows, names = fetch(ws[-1])
    per = len(rows) // B
    w, ids, level, d0, d1, want, outs, kb, kd, kv, nb = st[-1]
    for i in range(B):
        blk = rows[i * per:(i + 1) * per]
        ok_out = all(np.array_equal(w.ctx.poly_download(outs[i].Value[key].h, level + 1 - nb), want.value[key]) for key in ["0"] + ids)
        if ref is None and ok_out:
            ref = blk.copy()
            ref_i = i
            print(f"reference op: {per} launches; scratch order: {names[0]}")
            continue
        if ref is not None and not np.array_equal(blk, ref) and os.environ.get("MKHE_DEBUG_SNAP") and not ok_out:
            badrows = sorted(set(np.argwhere(blk != ref)[:, 0].tolist()))
            def snap(row):
                nbytes = C.c_size_t(0)
                buf = np.empty(64 << 20, dtype=np.uint64)
                w.ctx.check(w.ctx.dll.mkhe_debug_snap_read(w.ctx.ptr, C.c_int(row), buf.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_size_t(buf.nbytes), C.byref(nbytes)))
                return buf[:nbytes.value // 8].copy()
            l0 = badrows[0]
            got, good = snap(i * per + l0), snap(ref_i * per + l0)
            N = 1 << 15
            d = np.flatnonzero(got != good)
            prod, rem = d // (17 * N), d % (17 * N)
            slot, pos = rem // N, rem % N
            print(f"round {r} op {i}: snapshot after launch {l0}: {len(d)} differing words; products {np.unique(prod).tolist()} slots {np.unique(slot).tolist()} "
                  f"pos {pos.min()}..{pos.max()} distinct 512-blocks {np.unique(pos // 512).tolist()[:20]} first words {d[:6].tolist()}", flush=True)
            for dd in d[:4]:
                print(f"      word {dd}: got {got[dd]:#x} good {good[dd]:#x}")
        if ref is not None and not np.array_equal(blk, ref):
            bad = np.argwhere(blk != ref)
            l0 = bad[:, 0].min()
            cols = sorted(set(bad[bad[:, 0] == l0][:, 1].tolist()))
            sc = names[i * per + l0].split()
            print(f"round {r} op {i}: output {'OK' if ok_out else 'WRONG'}; first deviation after launch {l0} = {sc[0]}, buffers {[sc[1 + c] for c in cols if 1 + c < len(sc)]}; "
                  f"rows deviating {sorted(set(bad[:, 0].tolist()))[:12]}", flush=True)
        elif not ok_out:
            print(f"round {r} op {i}: output WRONG but all scratch checksums equal the reference", flush=True)
print("done")
