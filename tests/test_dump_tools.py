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
