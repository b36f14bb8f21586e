"""Build libswga_b200.so in-tree with nvcc for sm_100a (B200).  `python -m swga_b200.build`."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libswga_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--use_fast_math",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale():
    if not os.path.isfile(LIB):
        return True
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [
        os.path.join(os.path.dirname(HERE), "include", "swga_b200.h")]
    return any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        nvcc = "nvcc"
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + sources()
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libswga_b200.so")
    with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:
        f.write(res.stdout)
    return LIB


def build_set_finder(ref=None):
    """Compile the reference's set_finder (cliquer + swga front end) UNCHANGED from its own sources,
    where they lie, into swga_b200/bin/ -- the host remainder of the path.  `ref` (default: $SWGA_REFERENCE,
    else /root/reference) is a checkout of eclarke/swga.  Returns the path, or None when neither a prebuilt
    binary nor the reference sources are available (sets.set_finder() then says what to do)."""
    ref = ref or os.environ.get("SWGA_REFERENCE") or "/root/reference"
    out = os.path.join(HERE, "bin", "set_finder")
    if os.access(out, os.X_OK):
        return out
    src = os.path.join(ref, "ext", "cliquer")
    files = [os.path.join(src, f) for f in ("set_finder.c", "cliquer.c", "graph.c", "reorder.c")]
    if not all(os.path.isfile(f) for f in files):
        return None
    os.makedirs(os.path.dirname(out), exist_ok=True)
    # flags of ext/cliquer/Makefile:10,26
    subprocess.check_call(["gcc", "-O3", "-fomit-frame-pointer", "-funroll-loops", "-w", "-DENABLE_LONG_OPTIONS"]
                          + files + ["-o", out])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_set_finder())
