"""Builds lib/libctcx.so (the C-ABI library, include/ctcx.h) with nvcc for sm_100a, in-tree."""
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libctcx.so")
SOURCES = [os.path.join(CSRC, "ctcx_api.cu")]
HEADERS = [os.path.join(CSRC, "ctcx_kernels.cuh"), os.path.join(CSRC, "ctcx_beam_v2.cuh"),
           os.path.join(CSRC, "ctcx_beam_v3.cuh"),
           os.path.join(CSRC, "ctcx_beam_wide.cuh"),
           os.path.join(CSRC, "ctcx_math.cuh"),
           os.path.join(os.path.dirname(PKG_DIR), "include", "ctcx.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    # score arithmetic must not be contracted or re-associated (bit-exact parity with the reference)
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "--expt-extended-lambda", "-Xcompiler", "-fPIC", "-shared",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


HASH_PATH = os.path.join(LIB_DIR, "libctcx.srchash")


def source_hash():
    """Hash of everything the library is built from (sources, headers, flags)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in SOURCES + HEADERS:
        with open(f, "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()


def is_stale():
    """True if lib/libctcx.so is missing or was built from different sources. Content based: file
    times do not survive copying the tree to another box."""
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as fh:
        return fh.read().strip() != source_hash()


def build(force=False, verbose=False):
    """Compile the library if missing or built from other sources. Returns the .so path."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    subprocess.check_call(cmd)
    with open(HASH_PATH, "w") as fh:
        fh.write(source_hash() + "\n")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
