"""In-tree build of the CUDA library (sm_100a only).

`python -m lanczos_hls_b200.build` or `lanczos_hls_b200.build.build()` compiles
csrc/*.cu + csrc/*.cpp into lanczos_hls_b200/liblanczos_b200.so with nvcc.  nvcc
cross-compiles without a GPU; the built .so travels to the GPU box with the tree.

Every source becomes one object under lanczos_hls_b200/_obj/ (compiled in parallel, re-used when neither
the source, the headers nor the flags changed), then one link step.  `extra` flags (e.g. -DLZB_...) build a
development variant into another directory without touching the product library.
"""
import concurrent.futures
import glob
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblanczos_b200.so")
STAMP = os.path.join(HERE, ".build_stamp")
OBJ = os.path.join(HERE, "_obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build liblanczos_b200.so")


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def _headers():
    h = sorted(glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")))
    h.append(os.path.join(HERE, "..", "include", "lanczos_b200.h"))
    return h


def _hash(files, flags):
    h = hashlib.sha256()
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(flags).encode())
    return h.hexdigest()


def _compile_one(src, flags, objdir, verbose):
    deps = [src] + _headers()
    with open(src) as fh:
        if '#include "lanczos_' in fh.read():      # a translation unit that includes another .cu (lanczos_dyn2.cu)
            deps += [f for f in _sources() if f != src]
    key = _hash(deps, flags)
    obj = os.path.join(objdir, os.path.basename(src) + ".o")
    stamp = obj + ".stamp"
    if os.path.exists(obj) and os.path.exists(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == key:
                return obj, ""
    cmd = [_nvcc()] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s%s" % (os.path.basename(src), res.stdout, res.stderr))
    with open(stamp, "w") as fh:
        fh.write(key)
    return obj, res.stdout + res.stderr


def build(force=False, verbose=False, extra=(), out=None, only=None):
    """Compile what changed and link. Returns the path of the shared library.

    extra: additional nvcc flags (development variants); out: output library path (default: the product library);
    only: basenames of the sources to compile with `extra` (the others use the product flags and objects)."""
    extra = list(extra)
    lib = out or LIB
    variant = bool(extra) or out is not None
    flags = NVCC_FLAGS + extra
    objdir = OBJ if not variant else os.path.join(os.path.dirname(os.path.abspath(lib)), "_obj")
    os.makedirs(objdir, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    digest = _hash(_sources() + _headers(), flags)
    stamp = STAMP if not variant else lib + ".stamp"
    if not force and os.path.exists(lib) and os.path.exists(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == digest:
                return lib
    if force:
        for f in glob.glob(os.path.join(objdir, "*.stamp")):
            os.remove(f)
    jobs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        for src in _sources():
            use_extra = variant and (only is None or os.path.basename(src) in only)
            jobs.append(ex.submit(_compile_one, src, flags if use_extra else NVCC_FLAGS, objdir if use_extra else OBJ, verbose))
        objs = []
        for j in jobs:
            obj, log = j.result()
            objs.append(obj)
            if verbose and log:
                sys.stderr.write(log)
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", lib] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking " + os.path.basename(lib))
    with open(stamp, "w") as fh:
        fh.write(digest)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
