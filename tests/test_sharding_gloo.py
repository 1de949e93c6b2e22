"""CPU, world_size = 2, gloo: the party-sharded MulRelin algebra.  Each rank owns half of the parties'
relinearization keys, forms its partial x / y with the oracle's primitives, the partials are summed with an all-reduce
and reduced mod q (exactly what mkhe_ckks_mul_relin_sharded does with NCCL on the device), each rank finishes the
components of its own parties, and the assembled ciphertext must equal the single-process oracle bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mkhe_kklss_b200 import params as PR
from mkhe_kklss_b200 import sharding


def test_partition_is_balanced_and_complete():
    for k in (1, 2, 3, 4, 8, 13, 32):
        for world in (1, 2, 4, 8):
            ids = list(range(10, 10 + k))
            parts = [sharding.owned_parties(ids, world, r) for r in range(world)]
            assert sorted(sum(parts, [])) == ids
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1
            for i in ids:
                assert i in parts[sharding.owner_of(ids, world, i)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _allreduce_u64(arr):
    t = torch.from_numpy(arr.view(np.int64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy().view(np.uint64)


def _worker(rank, world, port, k, logN, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import parity
    from oracle import oracle as O
    O.set_threads(1)
    lit = PR.CKKS_PN14QP439.at_logn(logN)
    p = O.MKParams(lit.logN, lit.Q, lit.P, 2, seed=0xB2000004, crs_rots=[])
    prng = O.PRNG(0xB2000004 ^ 0x5EED)
    ids = list(range(k))
    rl = {i: O.RelinKey(i, parity.uniform_swk(prng, p), parity.uniform_swk(prng, p), parity.uniform_swk(prng, p)) for i in ids}
    L = p.max_level()
    mk = lambda: O.Ciphertext({**{"0": prng.uniform(p.ringQ, L)}, **{i: prng.uniform(p.ringQ, L) for i in ids}}, lit.scale)
    c0, c1 = mk(), mk()
    ks = O.KeySwitcher(p)
    own = sharding.owned_parties(ids, world, rank)
    h0 = {i: ks.decompose(L, c0.value[i]) for i in own}
    h1 = {i: ks.decompose(L, c1.value[i]) for i in own}
    beta, nQ, D = p.beta(L), p.nQ, p.D
    mods = p.Q + p.P

    def partial(keys, hs):
        acc = p.new_swk()
        for i in own:
            for dg in range(beta):
                for ring, sl in ((p.ringQ, slice(0, nQ)), (p.ringP, slice(nQ, D))):
                    a = np.ascontiguousarray(acc[dg, sl])
                    ring.mul_mont_add(np.ascontiguousarray(keys[i][dg, sl]), np.ascontiguousarray(hs[i][dg, sl]), a)
                    acc[dg, sl] = a
        for dg in range(beta):
            acc[dg, :nQ] = p.ringQ.mform(np.ascontiguousarray(acc[dg, :nQ]))
            acc[dg, nQ:] = p.ringP.mform(np.ascontiguousarray(acc[dg, nQ:]))
        return acc

    def exchange(part):
        s = _allreduce_u64(part)
        for j, q in enumerate(mods):
            s[:, j] = sharding.reduce_partials([s[:, j]], q)
        return s

    x = exchange(partial({i: rl[i].d for i in own}, h0))
    y = exchange(partial({i: rl[i].b for i in own}, h1))
    rq = p.ringQ
    A0, B0 = rq.ntt(c0.value["0"]), rq.ntt(c1.value["0"])
    out = {}
    c0_part = rq.intt(rq.mul_mont(rq.mform(A0), B0)) if rank == 0 else rq.new_poly()
    for i in own:
        t = rq.mul_mont(rq.mform(B0), rq.ntt(c0.value[i]))
        rq.mul_mont_add(rq.mform(A0), rq.ntt(c1.value[i]), t)
        ci = rq.intt(t)
        ci = rq.add(ci, ks.external_product_hoisted(L, h1[i], x))
        pi = ks.external_product_hoisted(L, h0[i], y)
        hp = ks.decompose(L, pi)
        c0_part = rq.add(c0_part, ks.external_product_hoisted(L, hp, rl[i].v))
        out[i] = rq.add(ci, ks.external_product_hoisted(L, hp, p.CRS[-1]))
    s = _allreduce_u64(c0_part)
    for j, q in enumerate(p.Q):
        s[j] = sharding.reduce_partials([s[j]], q)
    out["0"] = s
    gathered = [None] * world
    dist.all_gather_object(gathered, {k_: v for k_, v in out.items() if k_ != "0" or rank == 0})
    if rank == 0:
        full = {}
        for g in gathered:
            full.update(g)
        ref = O.CKKSEvaluator(p, lit.scale)
        want = ref._new_binary(c0, c1)
        ref.ksw.mul_and_relin_hoisted(c0, c1, ref.hoisted_form(c0), ref.hoisted_form(c1), rl, want)
        ok = set(full) == set(want.value) and all(np.array_equal(full[kk], want.value[kk]) for kk in want.value)
        out_q.put(bool(ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("k", [2, 3])
def test_party_sharded_mul_relin_matches_oracle(k):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, 10, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=300)
        assert pr.exitcode == 0
    assert q.get(timeout=5) is True
