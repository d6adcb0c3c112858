"""In-tree build of librodc.so (hand-written sm_100a CUDA behind the C ABI of include/rodc.h).

``python -m rapidobjectdetectionusingcascadedcnns_b200.build`` cross-compiles without a GPU.
The .so lives next to this file so it travels to the GPU box with the repo snapshot.

Staleness is decided by CONTENT, not by mtime: every object file and the library carry a stamp with the
SHA-256 of the sources, headers and flags they were built from, so an edited ``csrc`` is never silently
ignored in favour of an old ``.so`` and a snapshot copy (which does not preserve mtimes) never triggers a
needless rebuild on the GPU box.

``build_library(probes=True)`` builds ``librodc_probes.so`` with ``-DRODC_PROBES`` (measurement aids of
``profiles/``: operand-bandwidth probe and drain-interval override of fc1).  The shipped library does not
contain those paths and nothing in the package loads the probe build.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "build")
LIB = os.path.join(HERE, "librodc.so")
LIB_PROBES = os.path.join(HERE, "librodc_probes.so")
SOURCES = ["plan.cu", "pyramid.cu", "features.cu", "fc.cu", "fc1_tc.cu", "stage.cu", "compact.cu", "group.cu", "prof.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _headers():
    hs = sorted(os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".cuh", ".h")))
    hs.append(os.path.join(os.path.dirname(HERE), "include", "rodc.h"))
    return hs


def _digest(paths, extra: str = "") -> str:
    h = hashlib.sha256(extra.encode())
    for p in paths:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stamp_ok(target: str, digest: str) -> bool:
    try:
        return os.path.exists(target) and open(target + ".stamp").read().strip() == digest
    except OSError:
        return False


def source_digest(probes: bool = False) -> str:
    """SHA-256 over every source, header and flag the library is built from."""
    flags = NVCC_FLAGS + (["-DRODC_PROBES"] if probes else [])
    return _digest([os.path.join(CSRC, s) for s in SOURCES] + _headers(), " ".join(flags))


def is_current(probes: bool = False) -> bool:
    return _stamp_ok(LIB_PROBES if probes else LIB, source_digest(probes))


def build_library(force: bool = False, verbose: bool = False, probes: bool = False) -> str:
    """Compile csrc/*.cu and link the library (serialised across processes with a file lock, so the
    ranks of a torchrun job cannot race when the library is missing or stale)."""
    import fcntl
    os.makedirs(OBJ, exist_ok=True)
    with open(os.path.join(OBJ, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose, probes)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool, probes: bool) -> str:
    lib = LIB_PROBES if probes else LIB
    total = source_digest(probes)
    if not force and _stamp_ok(lib, total):
        return lib
    headers = _headers()
    nvcc = _nvcc()
    flags = NVCC_FLAGS + (["-DRODC_PROBES"] if probes else [])
    suffix = ".probes.o" if probes else ".o"

    def compile_one(src):
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", suffix))
        d = _digest([s] + headers, " ".join(flags))
        if force or not _stamp_ok(o, d):
            cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
            if verbose:
                sys.stderr.write(r.stderr)
            with open(o + ".stamp", "w") as f:
                f.write(d)
        return o

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(lib + ".stamp", "w") as f:
        f.write(total)
    return lib


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, probes="--probes" in sys.argv))
