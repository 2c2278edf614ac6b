#!/usr/bin/env python
"""TEST INFRASTRUCTURE: put the reference's own Python package on top of libflacb200.so.

Copies /root/reference/pyflac (Python sources, cffi builders, libFLAC headers -- NOT its libFLAC binaries), its tests/ and
examples/passthrough.py into the git-ignored oracle/_ref/refpkg/, repoints the one link line of
pyflac/builder/build_args.py:49-51 at pyflac_b200/libflacb200.so exactly as INTEGRATION.md describes, and runs the
reference's own two build commands (scripts/install.sh:3-4).  The result travels to the GPU box with the gpurun snapshot;
tests/test_gpu_reference_suite.py then runs the reference's unmodified tests against the CUDA library.
Nothing under pyflac_b200/ uses any of this.  Needs /root/reference (this container only).
"""
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("REF", "/root/reference")
DEST = os.path.join(HERE, "_ref", "refpkg")


def main():
    if not os.path.isdir(os.path.join(REF, "pyflac")):
        print("oracle/_ref/refpkg:", "using prebuilt files" if os.path.isdir(DEST) else "reference not available")
        return 0
    lib = os.path.join(ROOT, "pyflac_b200", "libflacb200.so")
    if not os.path.exists(lib):
        print("oracle/_ref/refpkg: build pyflac_b200/libflacb200.so first")
        return 1
    shutil.rmtree(DEST, ignore_errors=True)
    os.makedirs(DEST)
    shutil.copytree(os.path.join(REF, "pyflac"), os.path.join(DEST, "pyflac"),
                    ignore=shutil.ignore_patterns("libraries", "__pycache__", "*.so", "*.o", "*.c"))
    shutil.copytree(os.path.join(REF, "tests"), os.path.join(DEST, "tests"), ignore=shutil.ignore_patterns("__pycache__"))
    os.makedirs(os.path.join(DEST, "examples"))
    shutil.copy(os.path.join(REF, "examples", "passthrough.py"), os.path.join(DEST, "examples", "passthrough.py"))
    subprocess.run(["chmod", "-R", "u+w", DEST], check=True)
    # the one link line: libraries / library_dirs / rpath of the linux-x86_64 branch
    ba = os.path.join(DEST, "pyflac", "builder", "build_args.py")
    src = open(ba).read()
    src, n1 = re.subn(r"build_kwargs\['libraries'\] = \['FLAC-12\.1\.0'\]", "build_kwargs['libraries'] = ['flacb200']", src)
    libdir = os.path.join(ROOT, "pyflac_b200")
    src, n2 = re.subn(r"(elif system == 'Linux':.*?)build_kwargs\['library_dirs'\] = \[[^\n]*\]",
                      lambda m: m.group(1) + "build_kwargs['library_dirs'] = [%r]" % libdir, src, count=1, flags=re.S)
    # relocatable: the extension finds the library relative to itself (oracle/_ref/refpkg/pyflac -> pyflac_b200)
    src, n3 = re.subn(r"build_kwargs\['extra_link_args'\] = \['-Wl,-rpath,\$ORIGIN/libraries/' \+ architecture\]",
                      "build_kwargs['extra_link_args'] = ['-Wl,-rpath,$ORIGIN/../../../../pyflac_b200']", src)
    assert (n1, n2, n3) == (1, 1, 1), (n1, n2, n3)
    open(ba, "w").write(src)
    for b in ("encoder.py", "decoder.py"):
        r = subprocess.run([sys.executable, os.path.join("pyflac", "builder", b)], cwd=DEST, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout[-2000:] + r.stderr[-4000:])
            return 1
    for junk in ("_encoder.c", "_decoder.c", "_encoder.o", "_decoder.o"):
        p = os.path.join(DEST, "pyflac", junk)
        if os.path.exists(p):
            os.remove(p)
    print("oracle/_ref/refpkg built: reference pyflac", "linked against", lib)
    return 0


if __name__ == "__main__":
    sys.exit(main())
