import hashlib
import json
import os
import sys

import numpy as np
import pytest

# Every kernel of the library is loaded when CUDA initialises: with lazy loading the first launch of a kernel may wait
# for running kernels of other streams, and test_gpu_parity.py keeps a minutes-long single-image launch running under
# the other tests (named_configs_job).
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # one hardware work queue per stream of the lanes / single-image pipelines

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def sha(b) -> str:
    return hashlib.sha256(bytes(b)).hexdigest()


@pytest.fixture(scope="session")
def manifest():
    return json.load(open(os.path.join(GOLDEN, "manifest.json")))


@pytest.fixture(scope="session")
def oracle():
    from cpu_codecs import Oracle, build_oracle
    build_oracle()
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from cpu_codecs import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref/libnblic_ref.so not built (reference tree absent)")
    return Ref()


@pytest.fixture(scope="session")
def kodak(oracle, manifest):
    """dict name -> uint8 raster, recovered by decoding the committed reference e1n0 streams with the
    oracle and checked against the manifest's pixel hashes (so the oracle decode is pinned too)."""
    out = {}
    for name, ent in sorted(manifest["kodak"].items()):
        data = open(os.path.join(GOLDEN, "kodak_e1n0", name + ".nblic"), "rb").read()
        assert sha(data) == ent["streams"]["e1n0"]["sha256"]
        img, near, effort = oracle.n_decode(data)
        assert (near, effort) == (0, 1)
        assert sha(img.tobytes()) == ent["pixels_sha256"], name
        out[name] = img
    return out
