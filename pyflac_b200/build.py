"""Build libflacb200.so (sm_100a) in-tree with nvcc.  No JIT cache: the .so travels with the repo snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libflacb200.so")
SOURCES = ["engine.cu", "enc_analyze.cu", "enc_pack.cu", "enc_fused.cu", "flac_api_enc.cu", "dec_kernels.cu", "dec_engine.cu", "flac_api_dec.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",            # decision arithmetic must not be contracted; explicit fma() still emits DFMA
    "-Xcompiler", "-fPIC,-O2,-fno-fast-math,-ffp-contract=off",
    "-shared", "-cudart", "static",
]


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", f)
                                                                 for f in os.listdir(os.path.join(HERE, "..", "include"))]
    return any(os.path.getmtime(d) > t for d in deps)


def _deps_mtime():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    inc = os.path.join(HERE, "..", "include")
    deps += [os.path.join(inc, f) for f in os.listdir(inc)]
    return max(os.path.getmtime(d) for d in deps)


def build(force=False, verbose=False):
    """One nvcc -c per source (in parallel, objects cached under csrc/_obj), then one link.  Safe when several processes get here
    at once (the ranks of a torchrun launch on a box whose copy of the tree looks stale): one builds under a file lock, the
    others wait and find the library fresh; the library itself appears atomically (link to a temporary name, then rename)."""
    if not force and not needs_build():
        return LIB
    import fcntl
    try:
        lock = open(os.path.join(HERE, ".build.lock"), "w")
    except OSError:
        lock = None                                     # read-only tree: nothing to serialise against, the build below fails loudly
    try:
        if lock is not None:
            fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not needs_build():             # another process built it while this one waited
            return LIB
        return _build_locked(force, verbose)
    finally:
        if lock is not None:
            fcntl.flock(lock, fcntl.LOCK_UN)
            lock.close()


def _build_locked(force, verbose):
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(CSRC, "_obj")
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f not in ("-shared",)]
    cflags = cflags[:cflags.index("-cudart")]          # link-only flags stay out of the compile step
    hdr_t = _deps_mtime()

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj, ""
        cmd = [nvcc] + cflags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed on " + src)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(compile_one, sources()))
    if verbose:
        for _, log in res:
            print(log)
    tmp = LIB + ".tmp%d" % os.getpid()
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", tmp] + [o for o, _ in res] + ["-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed linking libflacb200.so")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
