"""Development aid, CPU only: the CUDA kernel SOURCES compiled for the coroutine emulator (tools/emu) must agree
with the oracle bit for bit.  This exercises index arithmetic / synchronisation structure of the very code that
runs on the B200; it is not a product path (the package never loads this library)."""
import pytest

import parity
from mkhe_kklss_b200 import params as PR
from mkhe_kklss_b200._lib import Library


@pytest.fixture(scope="module")
def emu():
    import __graft_entry__ as g
    return Library(g.build_emulator())


@pytest.mark.parametrize("logN", [12, 13])
def test_ckks_kernels_under_emulation(emu, logN):
    w = parity.CKKSWorld(PR.CKKS_PN14QP439.at_logn(logN), 2, lib=emu)
    parity.run_ckks_suite(w, quick=(logN != 12))
    w.close()


def test_lanes_under_emulation(emu):
    """forked contexts: shared handles, private pools, the cross-lane bookkeeping (host logic; the emulator is synchronous)"""
    w = parity.CKKSWorld(PR.CKKS_PN14QP439.at_logn(12), 2, lib=emu)
    parity.check_lanes(w, rounds=1)
    parity.check_mul_relin_new(w, w.ids, w.ids)          # the root is intact after its fork is gone
    w.close()


def test_cnn_parameter_set_under_emulation(emu):
    """48-bit primes with a 58-bit q0: digits far above the target modulus"""
    w = parity.CKKSWorld(PR.CNN_PN14QP433.at_logn(12), 2, lib=emu)
    parity.run_ckks_suite(w, quick=True)
    w.close()


def test_bfv_kernels_under_emulation(emu):
    w = parity.BFVWorld(PR.BFV_PN14QP439.at_logn(12), 2, lib=emu)
    parity.check_bfv_conv(w)
    parity.check_bfv_mul_relin(w, [0, 1], [0, 1])
    parity.check_bfv_mul_relin(w, [0], [0, 1])
    parity.check_bfv_linear_ops(w)
    w.close()


@pytest.mark.parametrize("lit", [PR.PN16QP1761_Q7, PR.PN16QP1761_Q7_ALPHA4], ids=lambda l: l.name)
def test_wide_digits_under_emulation(emu, lit):
    """four special primes: gamma = 2 -> alpha = 2 (the set the reference keeps commented out, mkrlwe_test.go:22-35),
    gamma = 1 -> alpha = 4.  Multi-limb digits take the exact lift of DecomposeAndSplit's general branch, a single-limb
    last digit the broadcast, and the lower levels end in partial digits."""
    w = parity.CKKSWorld(lit.at_logn(12), 2, lib=emu)
    parity.run_ckks_suite(w, quick=True)
    parity.check_decompose(w, level=5)
    parity.check_decompose(w, level=4)
    parity.check_mul_relin_new(w, w.ids, w.ids, level=5)
    w.close()


def test_mixed_levels_and_abi_checks_under_emulation(emu):
    w = parity.CKKSWorld(PR.PN16QP1761_Q7.at_logn(12), 2, lib=emu)
    parity.check_mixed_levels(w)
    parity.check_abi_negative(w)
    w.close()
    w = parity.CKKSWorld(PR.CKKS_PN14QP439.at_logn(12), 2, lib=emu, rots=(1,))
    parity.check_fork_after_queued_work(w, rounds=2)
    w.close()


def test_limb_sharded_single_rank_under_emulation(emu):
    """the limb-sharded op with a team of ONE rank (the emulator runs one rank at a time): team memory indirection of the P parts,
    the broadcast of p_id, the gather + Rescale path -- against the oracle"""
    parity.check_limb_sharded(PR.CKKS_PN14QP439.at_logn(12), 2, 1, lib=emu)
    parity.check_limb_sharded(PR.CKKS_PN15QP880.at_logn(12), 3, 1, lib=emu, level=5, ids0=[0, 1], ids1=[1, 2])
    parity.check_team_allgather(PR.CKKS_PN14QP439.at_logn(12), 1, lib=emu)


def test_pn16qp1761_full_limb_count_under_emulation(emu):
    """34 + 4 limbs, 17 two-limb digits"""
    w = parity.CKKSWorld(PR.PN16QP1761.at_logn(12), 2, lib=emu, rots=(1,))
    parity.check_decompose(w)
    parity.check_mul_relin_new(w, w.ids, w.ids)
    parity.check_rotate(w, w.ids, 1)
    w.close()


def test_elementwise_semantics_under_emulation(emu):
    """AddNew / SubNew / MultByConst / MulPtxtNew on real encryptions: decrypted results against the plain computation, with
    the reference's precision thresholds (pins what the element-wise ops MEAN, not only that oracle and device agree)"""
    w = parity.CKKSWorld(PR.CKKS_PN14QP439.at_logn(12), 2, lib=emu, real_keys=True)
    parity.check_ckks_semantics(w)
    w.close()


def test_cnn_flow_under_emulation(emu):
    """the reference's CNN op sequence at logN = 12 (rotations beyond N/2 wrap, the ones without a key chain powers of two)"""
    parity.check_cnn_flow(PR.CNN_PN14QP433.at_logn(12), lib=emu)


def test_keygen_and_encrypt_under_emulation(emu):
    """device-side KeyGenerator / Encryptor against the oracle's on the same counter-based streams, then the
    encrypt -> MulRelin -> Rotate -> Conjugate -> Decrypt flow on device-made material only"""
    parity.check_keygen(PR.CKKS_PN14QP439.at_logn(12), lib=emu)
    parity.check_keygen(PR.PN16QP1761_Q7.at_logn(12), lib=emu, rots=(1,), semantics=False)     # alpha = 2: multi-limb gadget digits
    parity.check_bfv_keygen(PR.BFV_PN14QP439.at_logn(12), lib=emu)
