"""Host-side party sharding for multi-GPU MulRelin: one process per GPU, rank g owns the
relinearization keys of the parties in `owned_parties(ids, world, g)`; ciphertexts are replicated (they are small).

The device-side exchange (ncclAllReduce of the partial x, y and of the c_0 contributions, then a reduction mod q) is inside
`mkhe_ckks_mul_relin_sharded`; this module only decides ownership and assembles the complete ciphertext afterwards.
"""
from __future__ import annotations


def owned_parties(ids, world: int, rank: int):
    """contiguous, balanced partition of the sorted party ids; every rank gets floor or ceil of k/world parties"""
    ids = sorted(ids)
    k = len(ids)
    base, extra = divmod(k, world)
    start = rank * base + min(rank, extra)
    return ids[start:start + base + (1 if rank < extra else 0)]


def owner_of(ids, world: int, id) -> int:
    for r in range(world):
        if id in owned_parties(ids, world, r):
            return r
    raise KeyError(id)


def reduce_partials(partials, q):
    """what the device does after the all-reduce: sum of <= 8 canonical residues, one reduction mod q
    (exact because the reference accumulates with modular adds)"""
    import numpy as np
    acc = np.zeros_like(partials[0])
    for p in partials:
        acc = acc + p                      # < 8 * 2^60 < 2^64: no wrap
    return acc % np.uint64(q)


class ShardedEvaluator:
    """mkckks.Evaluator.MulRelinNew over `world` ranks (each rank constructs one with its own Parameters/context)."""

    def __init__(self, evaluator, world: int, rank: int):
        self.ev, self.world, self.rank = evaluator, world, rank

    def MulRelinNew(self, op0, op1, rlkSet, all_ids=None):
        from .mkckks import Ciphertext
        ev = self.ev
        ctOut = ev.newCiphertextBinary(op0, op1)
        level = ctOut.Level()
        scale = op0.ScalingFactor() * op1.ScalingFactor()
        nb, newScale = ev._nb_rescales(scale, level, ev.params.Scale())
        ids0, ids1, idsO = op0.ids(), op1.ids(), ctOut.ids()
        own = owned_parties(all_ids if all_ids is not None else idsO, self.world, self.rank)
        key = lambda i, j: rlkSet.Value[i].Value[j].h if i in rlkSet.Value else 0
        ev.ctx.ckks_mul_relin_sharded(level, nb, ids0, op0.handles(ids0), ids1, op1.handles(ids1), own,
                                      [key(i, 0) for i in ids1], [key(i, 1) for i in ids0], [key(i, 2) for i in ids0],
                                      ev.params.CRS[-1].h, idsO, ctOut.handles(idsO))
        ctOut.Scale = newScale
        ctOut.valid = ["0"] + [i for i in idsO if i in own]
        return ctOut
