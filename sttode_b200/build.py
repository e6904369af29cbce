"""Build libsttode_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m sttode_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libsttode_sm100.so")
SOURCES = ["encoder.cu", "encoder_tc.cu", "backward.cu", "gemm_tc.cu", "decoder.cu", "decoder_tc.cu", "pmath.cu", "tc_selftest.cu", "comm.cu", "capi.cu"]      # + new files: keep in sync with csrc/
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "kernels.cuh"), os.path.join(CSRC, "tc.cuh"),
           os.path.join(os.path.dirname(HERE), "include", "sttode.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--compiler-options", "-fPIC", "-shared"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    return "nvcc"


HASH_PATH = LIB_PATH + ".srchash"


def source_hash():
    """sha256 over every source, header and compiler flag that goes into the library."""
    import hashlib
    h = hashlib.sha256()
    for path in [os.path.join(CSRC, s) for s in SOURCES] + HEADERS:
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(os.environ.get("STTODE_TC_TIMELINE", "").encode())
    return h.hexdigest()


def is_stale():
    """True when the .so is missing or was built from other sources (recorded source hash differs; mtimes are not trusted —
    the .so is git-ignored and ships prebuilt with the snapshot)."""
    if not os.path.isfile(LIB_PATH) or not os.path.isfile(HASH_PATH):
        return True
    with open(HASH_PATH) as f:
        return f.read().strip() != source_hash()


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    extra = ["-DSTTODE_TC_TIMELINE=" + os.environ["STTODE_TC_TIMELINE"]] if os.environ.get("STTODE_TC_TIMELINE") else []
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    with open(HASH_PATH, "w") as f:
        f.write(source_hash() + "\n")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
