"""Build ``librrin_b200.so`` in-tree with nvcc for sm_100a (no torch extension machinery:
the library has a plain C ABI and links only the static CUDA runtime)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "librrin_b200.so")
SOURCES = ["conv3x3.cu", "conv3x3_f16.cu", "glue.cu", "engine.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build librrin_b200.so")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "rrin_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, diag: bool = False, defines=(), lib_out: str = LIB) -> str:
    """``diag=True`` (``python -m rrin_b200.build --diag``) compiles the timing knobs and per-role cycle counters
    (RRIN_CONV_DBG / RRIN_CONV_PROF) into the conv kernels; the default library has neither."""
    if not force and not diag and not defines and lib_out == LIB and not _stale():
        return LIB
    nvcc = _nvcc()
    obj_dir = os.path.join(ROOT, "build" if lib_out == LIB else "build_" + os.path.basename(lib_out).replace(".", "_"))
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *(["-DRRIN_DIAG"] if diag else []), *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        objs.append(obj)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_out + ".tmp", *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(lib_out + ".tmp", lib_out)
    return lib_out


if __name__ == "__main__":
    # A/B builds: python -m rrin_b200.build --force -DNAME ... --out rrin_b200/librrin_b200_alt.so  (then RRIN_LIB=<that path>)
    defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    outp = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else LIB
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, diag="--diag" in sys.argv, defines=defs, lib_out=os.path.abspath(outp)))
