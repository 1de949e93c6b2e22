"""-m gpu, needs >= 2 GPUs (skipped on a single-GPU box): party-sharded MulRelinNew over NCCL must equal the oracle /
single-GPU result bit for bit (SURVEY test plan T5)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ndev():
    from mkhe_kklss_b200 import _lib
    try:
        return _lib.default_library().device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, k, logN, out_q, p2p=False):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [os.path.dirname(here), here]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import parity
    from oracle import oracle as O
    from mkhe_kklss_b200 import mkckks, mkrlwe, params as PR, sharding
    lit = PR.CKKS_PN15QP880.at_logn(logN)
    p = O.MKParams(lit.logN, lit.Q, lit.P, 2, seed=0xB2000004, crs_rots=[])
    prng = O.PRNG(0xB2000004 ^ 0x5EED)
    ids = list(range(k))
    rl = {i: O.RelinKey(i, parity.uniform_swk(prng, p), parity.uniform_swk(prng, p), parity.uniform_swk(prng, p)) for i in ids}
    L = p.max_level()
    mk = lambda: {**{"0": prng.uniform(p.ringQ, L)}, **{i: prng.uniform(p.ringQ, L) for i in ids}}
    v0, v1 = mk(), mk()
    dp = mkckks.Parameters(lit.logN, lit.Q, lit.P, lit.scale, device=rank)
    dp.SetCRS(-1, p.CRS[-1])
    uid = [dp.ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    dp.ctx.comm_init(world, rank, uid[0])
    if p2p:                                         # fused exchange over peer memory instead of the big all-reduce
        hs = [None] * world
        dist.all_gather_object(hs, dp.ctx.p2p_export())
        dp.ctx.p2p_import(world, rank, hs)
    own = sharding.owned_parties(ids, world, rank)
    rlk = mkrlwe.RelinearizationKeySet()
    for i in own:                                   # a rank only ever holds its own parties' keys
        rlk.AddRelinearizationKey(mkrlwe.RelinearizationKey(dp.ctx, i, rl[i].b, rl[i].d, rl[i].v))
    ev = sharding.ShardedEvaluator(mkckks.Evaluator(dp), world, rank)
    c0 = mkckks.Ciphertext.from_numpy(dp.ctx, v0, lit.scale)
    c1 = mkckks.Ciphertext.from_numpy(dp.ctx, v1, lit.scale)
    res = ev.MulRelinNew(c0, c1, rlk, all_ids=ids)
    if p2p:                                         # back to back: the exchange buffers are reused without host synchronisation
        for _ in range(3):
            res = ev.MulRelinNew(c0, c1, rlk, all_ids=ids)
    mine = {kk: res.Value[kk].numpy() for kk in res.valid if kk != "0" or rank == 0}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        full = {}
        for g in gathered:
            full.update(g)
        oev = O.CKKSEvaluator(p, lit.scale)
        want = oev.mul_relin_new(O.Ciphertext(v0, lit.scale), O.Ciphertext(v1, lit.scale), rl)
        ok = set(full) == set(want.value) and all(np.array_equal(full[kk], want.value[kk]) for kk in want.value)
        out_q.put(bool(ok))
    dist.destroy_process_group()
    dp.ctx.close()


@pytest.mark.skipif(_ndev() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("k,world,p2p", [(4, 2, False), (3, 2, False), (4, 2, True), (3, 2, True)])
def test_party_sharded_mul_relin_nccl(k, world, p2p):
    import torch.multiprocessing as mp
    if _ndev() < world:
        pytest.skip("not enough GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, 12, q, p2p)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=600)
        assert pr.exitcode == 0
    assert q.get(timeout=5) is True
