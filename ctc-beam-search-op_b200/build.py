"""Builds lib/libctcx.so (the C-ABI library, include/ctcx.h) with nvcc for sm_100a, in-tree. The
kernel families live in separate translation units (csrc/k_*.cu) that compile in parallel."""
import hashlib
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB_PATH = os.path.join(LIB_DIR, "libctcx.so")
HASH_PATH = os.path.join(LIB_DIR, "libctcx.srchash")
PUBLIC_HEADER = os.path.join(os.path.dirname(PKG_DIR), "include", "ctcx.h")
UNITS = ["ctcx_api.cu", "k_narrow.cu", "k_wide.cu", "k_generic.cu", "k_post.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    # score arithmetic must not be contracted or re-associated (bit-exact parity with the reference)
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "--expt-extended-lambda", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _sources():
    """Every file the library is built from: all of csrc/ plus the public header."""
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)
                   if f.endswith((".cu", ".cuh", ".h")))
    return files + [PUBLIC_HEADER]


def _unit_hash(unit):
    """Content hash of one translation unit's inputs (coarse: every header counts for every unit),
    keyed by file NAME, not by absolute path: the tree is copied to other boxes."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in _sources():
        if f.endswith(".cu") and os.path.basename(f) != unit:
            continue
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode() + b"\0" + fh.read())
    return h.hexdigest()


def source_hash():
    """Hash of everything the library is built from (sources, headers, flags)."""
    h = hashlib.sha256()
    for u in UNITS:
        h.update(_unit_hash(u).encode())
    return h.hexdigest()


def is_stale():
    """True if lib/libctcx.so is missing or was built from different sources. Content based: file
    times do not survive copying the tree to another box."""
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as fh:
        return fh.read().strip() != source_hash()


def _compile(unit, verbose):
    obj = os.path.join(OBJ_DIR, unit[:-3] + ".o")
    stamp = obj + ".hash"
    want = _unit_hash(unit)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == want:
        return obj
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-c", os.path.join(CSRC, unit), "-o", obj]
    subprocess.check_call(cmd)
    with open(stamp, "w") as fh:
        fh.write(want + "\n")
    return obj


def build(force=False, verbose=False):
    """Compile the library if missing or built from other sources. Returns the .so path."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        objs = list(ex.map(lambda u: _compile(u, verbose), UNITS))
    subprocess.check_call([nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-Xcompiler", "-fPIC", "-o", LIB_PATH] + objs)
    with open(HASH_PATH, "w") as fh:
        fh.write(source_hash() + "\n")
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
