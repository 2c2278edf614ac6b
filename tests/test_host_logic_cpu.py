"""CPU test: the host logic of the drop-in encoder / decoder handles against the reference's libFLAC 1.4.3, session by session
(tools/host_logic_check.py: the search for the stream marker with ID3v2 tags and junk in front of it, every metadata block type
through the metadata callback under the respond / ignore filters, blocks larger than one read, malformed and truncated metadata,
getters / decode positions / refused calls around init, flush, reset and finish; encoder streams without a sample with every failing
callback).  The shipped library has no CPU path -- a handle cannot even be initialised without a CUDA device -- so the comparison
runs on a SCRATCH build under a temporary directory in which those two init checks are compiled out (tools/host_logic_check.sh says
how); everything in these sessions stops short of the first audio frame, i.e. never needs a kernel.  The same sessions continue into
the frames on the GPU in tests/test_gpu_zz_host_logic.py."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_handle_api_host_logic_matches_libflac(checkers, tmp_path):
    if not checkers.ref_available():
        pytest.skip("oracle/_ref not present")
    if not os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")):
        pytest.skip("nvcc not present")
    from pyflac_b200 import _native
    _native.lib()                                                  # the product's objects are current (the scratch build links its kernels from them)
    r = subprocess.run(["bash", os.path.join(ROOT, "tools", "host_logic_check.sh"), str(tmp_path / "scratch")],
                       capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    last = r.stdout.strip().splitlines()[-1]
    assert last.endswith(" 0 differ from libFLAC") and int(last.split()[0]) >= 3400, tail
