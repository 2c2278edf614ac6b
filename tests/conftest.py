import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


_HAVE_GPU = None


def have_gpu():
    """True when libflacb200 can create an engine on device 0 (asked once per session)."""
    global _HAVE_GPU
    if _HAVE_GPU is None:
        try:
            from pyflac_b200 import _native as nat
            nat.Engine(0).close()
            _HAVE_GPU = True
        except Exception:
            _HAVE_GPU = False
    return _HAVE_GPU


def pytest_collection_modifyitems(config, items):
    # GPU-marked tests skip (not error) on a host without a CUDA device
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if gpu_items and not have_gpu():
        skip = pytest.mark.skip(reason="no CUDA device (libflacb200 has no CPU fallback)")
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def checkers():
    """Build the test-only checker libraries once per session."""
    import _checkers
    _checkers.build_checkers()
    return _checkers


def golden_cases():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "manifest.json")) as f:
        return json.load(f)["cases"]


def load_golden(case):
    import numpy as np
    g = os.path.join(ROOT, "tests", "golden")
    x = np.load(os.path.join(g, case["name"] + ".pcm.npy"))
    with open(os.path.join(g, case["name"] + ".flac"), "rb") as f:
        flac = f.read()
    return x, flac


def fixture_cases():
    """pyFLAC's own tests/data/*.flac (imported by tests/golden/make_fixtures.py)."""
    import json
    with open(os.path.join(ROOT, "tests", "golden", "fixtures", "fixtures.json")) as f:
        return json.load(f)["cases"]


def fixture_path(name):
    return os.path.join(ROOT, "tests", "golden", "fixtures", name)


def pcm_md5(x, bps):
    """MD5 the way STREAMINFO defines it: interleaved little-endian samples of (bps+7)/8 bytes."""
    import hashlib
    import numpy as np
    w = (bps + 7) // 8
    raw = np.ascontiguousarray(np.asarray(x).astype("<i4")).view(np.uint8).reshape(-1, 4)
    return hashlib.md5(np.ascontiguousarray(raw[:, :w]).tobytes()).hexdigest()


def foreign_cases():
    """decode fixtures the presets never produce (tests/golden/make_foreign.py): tuned libFLAC output and hand-written streams"""
    import json
    with open(os.path.join(ROOT, "tests", "golden", "foreign", "foreign.json")) as f:
        return json.load(f)["cases"]


def foreign_path(name):
    return os.path.join(ROOT, "tests", "golden", "foreign", name)
