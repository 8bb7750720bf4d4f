"""Build recipe for libgckpp_b200.so (sm_100a only) -- explicit nvcc, in-tree output.

Steps: (1) kppgen emits gen/*.h, gen/*.cuh from the committed mechanism IR;
(2) each .cu is compiled to an object with `-gencode arch=compute_100a,code=sm_100a -lineinfo`;
(3) objects are linked into geos_chem_b200/libgckpp_b200.so (plain C ABI, no torch).
Objects are rebuilt only when a source/dependency is newer.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
GEN = os.path.join(CSRC, "gen")
# GCKPP_BUILD_TAG=<tag> builds a variant (other -D flags) into its own object directory and lib<...>_<tag>.so
TAG = os.environ.get("GCKPP_BUILD_TAG", "")
OBJ = os.path.join(CSRC, "build" + ("_" + TAG if TAG else ""))
LIB = os.path.join(HERE, "libgckpp_b200%s.so" % ("_" + TAG if TAG else ""))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]

# translation unit -> extra flags
UNITS = {
    # Update_RCONST and the small kernels: no FMA contraction, so the rate laws round operation by operation like the
    # reference's expressions (the uptake laws contain cancellations that amplify a fused rounding); 0.1 % of a step
    "kernels_misc.cu": ["-fmad=false"],
    # arithmetic-reference kernel: no FMA contraction so sums round like the reference's
    "ros_generic.cu": ["-fmad=false"],
    # production kernel: shared-memory-resident, FMA allowed
    "ros_smem.cu": (["-DSMEM_PROFILE"] if os.environ.get("GCKPP_SMEM_PROFILE") else []) + os.environ.get("GCKPP_SMEM_DEFS", "").split(),
    # production kernel: one warp per cell
    "ros_warp.cu": (["-DWARP_PROFILE"] if os.environ.get("GCKPP_WARP_PROFILE") else []) + os.environ.get("GCKPP_WARP_DEFS", "").split(),
    # Do_FullChem's pieces around the integration, reference operation order
    "post.cu": ["-fmad=false"],
    # one cell per thread, generated straight-line code (small mechanisms)
    "ros_unrolled.cu": os.environ.get("GCKPP_UNR_DEFS", "").split(),
    # one cell per lane, streamed workspace
    "ros_lane.cu": (["-DLANE_PROFILE"] if os.environ.get("GCKPP_LANE_PROFILE") else []) + os.environ.get("GCKPP_LANE_DEFS", "").split(),
    "gckpp_gpu.cu": os.environ.get("GCKPP_SMEM_DEFS", "").split() + os.environ.get("GCKPP_WARP_DEFS", "").split(),
}


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list if os.path.exists(s))


def generate():
    sys.path.insert(0, os.path.dirname(HERE))
    from geos_chem_b200.kppgen import emit_cuda
    emit_cuda.main(["--out", GEN])


def build(verbose=False):
    generate()
    os.makedirs(OBJ, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps += [os.path.join(GEN, f) for f in os.listdir(GEN)]
    deps.append(os.path.join(HERE, "..", "include", "gckpp_gpu.h"))
    objs = []
    logs = []
    procs = []
    for unit, extra in UNITS.items():
        src = os.path.join(CSRC, unit)
        obj = os.path.join(OBJ, unit.replace(".cu", ".o"))
        objs.append(obj)
        if _newer([src] + deps, obj):
            cmd = [NVCC] + ARCH + COMMON + extra + ["-c", src, "-o", obj]
            procs.append((unit, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for unit, cmd, p in procs:
        out, _ = p.communicate()
        logs.append("$ " + " ".join(cmd) + "\n" + out)
        if p.returncode != 0:
            sys.stderr.write(logs[-1])
            raise RuntimeError("nvcc failed for " + unit)
    if procs or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        subprocess.check_call(cmd)
    if logs:
        with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
            f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
