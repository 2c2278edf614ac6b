"""The reference's OWN package and tests on top of libflacb200.so.

oracle/build_refpkg.py (run by __graft_entry__.build() where /root/reference exists) copies the reference's pyflac/ sources,
tests/ and examples/passthrough.py into the git-ignored oracle/_ref/refpkg/, changes the one link line of
pyflac/builder/build_args.py:49-51 to `-lflacb200` and builds its two cffi modules with the reference's own commands.
Here the UNMODIFIED reference tests (tests/test_encoder.py, tests/test_decoder.py: 39 tests) and examples/passthrough.py
run in a subprocess against the CUDA library; `soundfile` (libsndfile, absent from this image) is stood in for by
tests/soundfile_shim (this repo's WAV reader / writer)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "oracle", "_ref", "refpkg")


def _env():
    path = os.pathsep.join([os.path.join(ROOT, "tests", "soundfile_shim"), PKG, ROOT, os.environ.get("PYTHONPATH", "")])
    return dict(os.environ, PYTHONPATH=path)


def _need_pkg():
    if not any(f.startswith("_encoder") and f.endswith(".so") for f in (os.listdir(os.path.join(PKG, "pyflac")) if os.path.isdir(os.path.join(PKG, "pyflac")) else [])):
        pytest.skip("oracle/_ref/refpkg not built (needs /root/reference at build time)")


def test_reference_package_binds_this_library():
    """`import pyflac` resolves every FLAC__* symbol out of libflacb200.so (no libFLAC is loaded)."""
    _need_pkg()
    code = ("import pyflac, sys\n"
            "from pyflac._encoder import lib as el\n"
            "from pyflac._decoder import lib as dl\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libflacb200.so' in maps and 'libFLAC' not in maps, maps\n"
            "assert pyflac.__file__.startswith(sys.argv[1]), pyflac.__file__\n"
            "print('ok', pyflac.__version__)\n")
    r = subprocess.run([sys.executable, "-c", code, PKG], capture_output=True, text=True, env=_env(), cwd=PKG)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr


def test_reference_test_suite_unchanged():
    """/root/reference/tests/test_encoder.py + test_decoder.py, byte-for-byte copies, all green on the CUDA path."""
    _need_pkg()
    r = subprocess.run([sys.executable, "-m", "pytest", "tests", "-q", "-x", "-p", "no:cacheprovider"], capture_output=True, text=True,
                       env=_env(), cwd=PKG, timeout=900)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail
    n = int(r.stdout.strip().splitlines()[-1].split(" passed")[0].split()[-1])
    assert n == 39, tail


@pytest.mark.parametrize("wav", ["mono.wav", "stereo.wav", "32bit.wav"])
def test_reference_passthrough_example(wav):
    """examples/passthrough.py: StreamEncoder -> StreamDecoder, `np.array_equal` on every block (passthrough.py:76)."""
    _need_pkg()
    r = subprocess.run([sys.executable, os.path.join("examples", "passthrough.py"), os.path.join("tests", "data", wav)],
                       capture_output=True, text=True, env=_env(), cwd=PKG, timeout=600)
    assert r.returncode == 0 and "Verified OK" in r.stdout, r.stdout + r.stderr[-2000:]
