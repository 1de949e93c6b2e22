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
    w.close()
