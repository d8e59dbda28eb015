"""Build the sm_100a shared library (C ABI of include/text2pos_b200.h) in-tree with nvcc.

    python -m text2pos_cvpr2022_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libtext2pos_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + [
        os.path.join(os.path.dirname(HERE), "include", "text2pos_b200.h")
    ]


STAMP_PATH = os.path.join(LIB_DIR, "BUILD_STAMP")


def source_hash() -> str:
    """sha256 over every source the library is built from (+ the compile flags): what ``BUILD_STAMP`` records next to the
    ``.so`` and what ``_lib.load()`` checks, so a source edit can never ship with a stale binary (mtimes do not survive the
    snapshot to the GPU box)."""
    import hashlib

    h = hashlib.sha256()
    h.update(" ".join(ARCH_FLAGS + ["-O3", "-std=c++17"]).encode())
    for p in sorted(_deps()):
        if os.path.exists(p):
            h.update(os.path.basename(p).encode())
            h.update(open(p, "rb").read())
    return h.hexdigest()


def built_hash() -> str:
    try:
        return open(STAMP_PATH).read().strip()
    except OSError:
        return ""


def is_stale() -> bool:
    return not os.path.exists(LIB_PATH) or built_hash() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    common = [NVCC, *ARCH_FLAGS, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
    if verbose:
        common += ["-Xptxas", "-v"]
    procs = []
    objs = []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        procs.append((src, subprocess.Popen(common + ["-c", src, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out = p.communicate()[0].decode()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed on {src}\n")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    tmp = LIB_PATH + ".tmp"
    subprocess.check_call([NVCC, *ARCH_FLAGS, "-shared", "-o", tmp, *objs])
    os.replace(tmp, LIB_PATH)
    with open(STAMP_PATH, "w") as f:
        f.write(source_hash() + "\n")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
