"""Build libgsf.so (sm_100a only) in-tree with nvcc; no torch, no JIT cache.

    python -m gps_optimize_slam_b200.build            # rebuild if sources are newer
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.environ.get("GSF_LIB_OUT") or os.path.join(PKG_DIR, "libgsf.so")   # GSF_LIB_OUT: A/B builds (with GSF_NVCC_EXTRA)
SOURCES = ["gsf_fused.cu", "gsf_fast.cu", "gsf_long.cu", "gsf_ate.cu", "gsf_kernels.cu", "gsf_ransac.cu", "gsf_grid.cu", "gsf_ekf_api.cu", "gsf_synth.cu", "gsf_gpsfilter.cu", "gsf_text.cu", "gsf_f32.cu", "gsf_assoc_long.cu", "gsf_capi.cu"]
EXTRA = os.environ.get("GSF_NVCC_EXTRA", "").split()
PER_FILE = {}                     # per-source extra flags (tuning hook)
NVCC_FLAGS = EXTRA + ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--fmad=true"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libgsf.so cannot be built")


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(PKG_DIR, "..", "include", "gsf.h"))
    return any(os.path.getmtime(d) > built for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB_PATH
    nvcc = _nvcc()
    objs = []
    build_dir = os.path.join(PKG_DIR, "build", os.path.basename(LIB_PATH).replace(".so", ""))
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *PER_FILE.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    link = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
