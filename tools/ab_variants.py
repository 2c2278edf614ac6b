"""Build compile-time variants of one kernel file and time a command with each (one gpurun call for a whole A/B sweep).

    python tools/ab_variants.py build dec_kernels.cu  base:  ring8:-DFB_DEC_RING=8  ctas12:-DFB_DEC_MIN_CTAS=12
        -> _variants/libflacb200_<tag>.so (the other objects are reused from pyflac_b200/csrc/_obj; run build.py first)
    python tools/ab_variants.py run -- python tools/prof_decode.py 4096 131072        (on the GPU box)
        -> runs the command once per variant with the variant copied over pyflac_b200/libflacb200.so, then restores it

Every variant goes through the same tests before its timing counts: put the pytest command in front of the timing one.
_variants/ is scratch (git-ignored); it travels to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyflac_b200 import build as B  # noqa: E402

VAR = os.path.join(ROOT, "_variants")


def build(src, specs):
    os.makedirs(VAR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cflags = [f for f in B.NVCC_FLAGS if f != "-shared"]
    cflags = cflags[:cflags.index("-cudart")]
    objdir = os.path.join(B.CSRC, "_obj")
    others = [os.path.join(objdir, s + ".o") for s in B.SOURCES if s != src]
    for o in others:
        if not os.path.exists(o):
            raise SystemExit("missing " + o + ": run python pyflac_b200/build.py first")
    for spec in specs:
        tag, _, defs = spec.partition(":")
        obj = os.path.join(VAR, f"{src}.{tag}.o")
        subprocess.run([nvcc] + cflags + [d for d in defs.split(",") if d] + ["-Xptxas", "-v", "-c", "-o", obj, os.path.join(B.CSRC, src)], check=True)
        lib = os.path.join(VAR, f"libflacb200_{tag}.so")
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", lib] + others + [obj, "-lpthread"], check=True)
        print("built", lib)


def run(cmd):
    keep = B.LIB + ".keep"
    shutil.copyfile(B.LIB, keep)
    try:
        for f in sorted(os.listdir(VAR)):
            if f.startswith("libflacb200_") and f.endswith(".so"):
                shutil.copyfile(os.path.join(VAR, f), B.LIB)
                print("==", f[len("libflacb200_"):-3], flush=True)
                subprocess.run(cmd)
    finally:
        shutil.move(keep, B.LIB)


if __name__ == "__main__":
    if len(sys.argv) >= 4 and sys.argv[1] == "build":
        build(sys.argv[2], sys.argv[3:])
    elif len(sys.argv) >= 4 and sys.argv[1] == "run" and sys.argv[2] == "--":
        run(sys.argv[3:])
    else:
        raise SystemExit(__doc__)
