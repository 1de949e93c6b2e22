"""The golden-dump tool chain (go/dump_golden_test.go writes the format from the real Go reference; tools/replay_dump.py replays
it): here the dump is made by the oracle, which tests the format, the loaders and the replay paths -- not the reference."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import replay_dump as R  # noqa: E402


@pytest.fixture(scope="module")
def dump_dir(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("mkhe_dump"))
    R.make(d, 12, 2)
    return d


def test_replay_through_the_oracle(dump_dir):
    report = []
    R.replay_oracle(R.load(dump_dir), report)
    assert report and all(ok for _, ok, _ in report), report


def test_go_dumper_and_replayer_agree_on_file_names():
    """every file the replayer loads is written by the Go dumper under the same name"""
    go = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "go", "dump_golden_test.go")).read()
    for name in ('"psi_"', '"psiinv_"', '"ninv.bin"', '"crs_u.bin"', '"crs_rot.bin"', '"rlk_"', '"_b.bin"', '"_d.bin"', '"_v.bin"', '"rk_"',
                 '"_c0.bin"', '"_p"', '"ct0"', '"ct1"', '"mul"', '"rot"', '"mkhe-dump-1"', '"mul_level"', '"mul_scale"'):
        assert name in go, name


@pytest.mark.gpu
def test_replay_through_the_device(dump_dir):
    x = R.load(dump_dir)
    report = []
    R.replay_device(x, report, use_dumped_tables=True)      # lattigo's tables handed over verbatim (mkhe_ctx_set_ntt_tables)
    R.replay_device(x, report, use_dumped_tables=False)     # the library's own tables
    assert report and all(ok for _, ok, _ in report), report


@pytest.fixture(scope="module")
def bfv_dump_dir(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("mkhe_dump_bfv"))
    R.make_bfv(d, 12, 2)
    return d


def test_bfv_replay_through_the_oracle(bfv_dump_dir):
    report = []
    R.replay_bfv_oracle(R.load_bfv(bfv_dump_dir), report)
    assert report and all(ok for _, ok, _ in report), report


def test_bfv_go_dumper_and_replayer_agree_on_file_names():
    go = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "go", "dump_golden_bfv_test.go")).read()
    for name in ('"psi_"', '"psiinv_"', '"ninv.bin"', '"crs_u.bin"', '"rlk_"', '"_b1.bin"', '"_d1.bin"', '"_v.bin"', '"_b2.bin"', '"_d2.bin"',
                 '"_c0.bin"', '"_p"', '"ct0"', '"ct1"', '"mul"', '"mkhe-dump-1"', '"scheme": "bfv"', '"QMul"', '"T"'):
        assert name in go, name


def test_go_shim_binds_only_declared_entry_points():
    """every C.mkhe_* call of go/mkgpu/mkgpu.go names an entry point include/mkhe.h declares"""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "mkhe.h")).read()
    declared = set(re.findall(r"\b(mkhe_[a-z0-9_]+)\s*\(", hdr))
    used = set(re.findall(r"C\.(mkhe_[a-z0-9_]+)\(", open(os.path.join(root, "go", "mkgpu", "mkgpu.go")).read()))
    assert used and used <= declared, sorted(used - declared)


@pytest.mark.gpu
def test_bfv_replay_through_the_device(bfv_dump_dir):
    x = R.load_bfv(bfv_dump_dir)
    report = []
    R.replay_bfv_device(x, report, use_dumped_tables=True)
    R.replay_bfv_device(x, report, use_dumped_tables=False)
    assert report and all(ok for _, ok, _ in report), report
