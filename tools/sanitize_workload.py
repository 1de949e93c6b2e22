"""development: the logN = 12 parity suite driven under compute-sanitizer (tools/sanitize.sh).  Every op is compared with the
oracle as usual; the sanitizer watches the kernels (memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory
hazards, including the TMA-fed rings of k_ntt_pass2 and k_mac_intt)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from mkhe_kklss_b200 import params as PR  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "ckks"
if which == "ckks":
    w = parity.CKKSWorld(PR.CKKS_PN14QP439.at_logn(12), 2)
    parity.check_ntt(w)
    parity.check_decompose(w)
    parity.check_external_product(w)
    parity.check_mul_relin_new(w, w.ids, w.ids)
    parity.check_mul_relin_hoisted(w, w.ids[:1], w.ids)
    parity.check_rotate(w, w.ids, 2)
    parity.check_conjugate(w, w.ids)
    parity.check_rescale(w)
    w.close()
elif which == "ckks15":
    w = parity.CKKSWorld(PR.CKKS_PN15QP880.at_logn(12), 3, rots=(1,))
    parity.check_mul_relin_new(w, w.ids, w.ids)
    parity.check_rotate(w, w.ids, 1)
    w.close()
elif which == "bfv":
    w = parity.BFVWorld(PR.BFV_PN14QP439.at_logn(12), 2)
    parity.check_bfv_conv(w)
    parity.check_bfv_mul_relin(w, w.ids, w.ids)
    w.close()
elif which == "wide":
    w = parity.CKKSWorld(PR.PN16QP1761_Q7.at_logn(12), 2)
    parity.check_decompose(w)
    parity.check_mul_relin_new(w, w.ids, w.ids)
    w.close()
elif which == "keygen":
    # round 2: samplers, key generation, encryption, the stand-alone NTT-domain ModDown
    parity.check_keygen(PR.CKKS_PN14QP439.at_logn(12), semantics=False)
    parity.check_bfv_keygen(PR.BFV_PN14QP439.at_logn(12))
    w = parity.CKKSWorld(PR.CKKS_PN14QP439.at_logn(12), 2)
    parity.check_moddown_ntt(w)
    w.close()
print("sanitize workload", which, "OK")
