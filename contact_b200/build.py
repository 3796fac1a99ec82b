"""Build the CUDA library in-tree: contact_b200/lib/libcontact_addon_b200.so (sm_100a only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
SO = os.path.join(LIBDIR, "libcontact_addon_b200.so")


def _sources():
    out = [os.path.join(HERE, "..", "include", "contact_addon_b200.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h", ".inc")):
            out.append(os.path.join(CSRC, f))
    return out


STAMP = os.path.join(LIBDIR, "build_stamp.txt")


def _digest():
    """Content hash of all sources: the staleness test (mtimes miss an edit made while a 4-minute nvcc run is under way)."""
    import hashlib
    h = hashlib.sha256()
    for f in _sources():
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(SO):
        return True
    if os.path.exists(STAMP):
        return open(STAMP).read().strip() != _digest()
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    os.makedirs(LIBDIR, exist_ok=True)
    digest = _digest()                      # of the sources as they are when the compiler starts
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # (no --split-compile: it halves the build time but the split ptxas units cost 20 % on k_snorm_batch -- 47.2 k against 58.4 k
    #  solves/s on the same box, gpurun_out/ab_split.log -- through spills in the warp-resident product)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
           "-Xcompiler", "-fPIC", "-shared", "-o", SO, os.path.join(CSRC, "contact_addon_b200.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libcontact_addon_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    with open(STAMP, "w") as fh:
        fh.write(digest + "\n")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
