import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


# several ranks as contexts of this process (tests/test_gpu_sharded_local.py): one hardware queue per stream, so that a
# kernel waiting for its peers can never sit in front of a peer's kernel (read at CUDA initialisation)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


# E_t = T[dim_t] commitments regrouped from the dim_t bucket sums (MsmJob::group_*, msm.cu): on by default from 2^18
# points per rank; the tests lower the threshold so that proofs with >= 2^10 lookups per rank take that path (smaller
# ones keep the plain MSM) — read once per process by the library
os.environ.setdefault("B200_MSM_GROUP_MIN_POINTS", "1024")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
