"""Builds ``csrc/libcavoid_sm100a.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m gym_collision_avoidance_b200.build [--force] [--verbose]

Flags: ``-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false`` -- FMA contraction is disabled on
purpose: the ORCA solve must reproduce the reference's x86-64 float32 arithmetic bit for bit and the float64
state update must round like NumPy, which an a*b+c -> fma rewrite would break.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
ROOT = os.path.dirname(PKG)
OUT = os.path.join(CSRC, "libcavoid_sm100a.so")
SOURCES = ["cav_step.cu", "cav_tpa.cu", "cav_tpa_n.cu", "cav_policy.cu", "cav_ga3c.cu", "cav_laser.cu", "cav_reset.cu", "cav_scenario.cu", "cav_abi.cu"]
HEADERS = ["cav_common.cuh", "cav_orca.cuh", "cav_tpa_common.cuh", "cav_sortnet.cuh", "cav_scenario_gen.cuh", "cav_hd.cuh", "cav_keys.cuh", "cav_tti.cuh", "cav_math.cuh", "cav_move.cuh", "cav_orca_core.cuh", "cav_kdtree.cuh", "cav_kernels.h", os.path.join(ROOT, "include", "cavoid.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h)
                                                        for h in HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    extra = ["-Xptxas", "-v"] if verbose else []
    extra += os.environ.get("CAV_NVCC_EXTRA", "").split()  # tuning experiments only (e.g. -DCAV_TPA_N_MINB_F32=6)

    def cc(src):
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout))
        if verbose:
            print(r.stdout)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(cc, SOURCES))
    cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
