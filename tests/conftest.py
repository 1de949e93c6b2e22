import os
import sys

# several ranks of a limb-sharded team live in ONE process on ONE device in these tests: their in-kernel barriers spin until the
# other ranks' kernels have run, so every stream needs a hardware queue of its own (the default of 8 connections would make
# streams share queues and a spinning barrier could sit in front of the very kernels it waits for)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
