"""The C-ABI library loads and exports every symbol include/gckpp_gpu.h declares (no compute calls:
this runs without a GPU)."""
import ctypes
import os
import re

import pytest

from geos_chem_b200 import kpp

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "gckpp_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gckpp_gpu_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "libgckpp_b200.so does not export %s" % s
    assert sorted(kpp.EXPORTS) == syms


def test_dims_and_names_without_gpu(lib):
    d = kpp.mech_dims("fullchem")
    assert (d["nvar"], d["nfix"], d["nspec"], d["nreact"], d["lu_nonzero"]) == (353, 3, 356, 1058, 5683)
    assert (d["nphot"], d["next"]) == (177, 113)
    assert kpp.mech_dims("Hg")["nvar"] == 32 and kpp.mech_dims("carbon")["nvar"] == 12
    names = kpp.spc_names("fullchem")
    assert names[0] == "CH2I2" and names[-3:] == ["H2", "N2", "O2"]


def test_no_cpu_fallback(lib):
    """without a CUDA device the product path must fail loudly, not compute on the CPU"""
    import torch
    if torch.cuda.is_available():
        return
    try:
        kpp.KppSolver("fullchem")
    except kpp.KppError as e:
        assert "gckpp_gpu_init failed" in str(e)
    else:
        raise AssertionError("KppSolver constructed without a GPU")


def test_product_does_not_touch_oracle():
    """only tests/, bench.py and __graft_entry__.smoke() may use oracle/"""
    pkg = os.path.join(ROOT, "geos_chem_b200")
    for dp, _, fs in os.walk(pkg):
        if "build" in dp.split(os.sep):
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "pyoracle" not in txt and "kpp_oracle" not in txt and "libkpp_oracle" not in txt, os.path.join(dp, f)


def test_smem_kernel_plan_fits_b200(lib):
    """host-only query of the shared-memory kernel's static plan: it must fit one B200 SM (227 KB opt-in)"""
    from geos_chem_b200 import kpp
    for mech in ("fullchem", "Hg"):
        p = kpp.plan_info(mech)
        assert 0 < p["smem_bytes"] + 16 <= 227 * 1024, p      # + the 16 bytes of static shared memory
        assert p["cells_per_block"] >= 1
    p = kpp.plan_info("fullchem")
    # head/tail split of the elimination DAG: 21 LU rounds and 19 sweep rounds instead of 72 + 68 levels
    assert (p["n_lu"], p["n_fwd"], p["n_bwd"]) == (21, 10, 9)


def test_production_kernel_sass_has_no_divergence_checks_or_spills(lib):
    """The block kernel's control flow is warp-uniform for the compiler (warp index and bundle length come through
    broadcasts, csrc/ros_smem.cu): its SASS must contain no BRA.DIV / WARPSYNC and no local-memory traffic.  Those
    scaffolds cost 8 % of the kernel's throughput when they crept in (profiles/r02au_*, r02av_*)."""
    import shutil
    import subprocess
    obj = os.path.join(os.path.dirname(kpp.LIB_PATH), "csrc", "build", "ros_smem.o")
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(obj) or not os.path.exists(cuobjdump):
        pytest.skip("no object file / cuobjdump (the library was built elsewhere)")
    sass = subprocess.run([cuobjdump, "-sass", obj], capture_output=True, text=True, check=True).stdout
    blocks = sass.split("Function : ")
    prod = [b for b in blocks if "ros_smem_kernel" in b.split("\n", 1)[0] and "fullchem_dims" in b.split("\n", 1)[0]]
    assert len(prod) == 2, [b.split("\n", 1)[0] for b in blocks[1:]]          # the plain and the auto-reduce instance
    for b in prod:
        name = b.split("\n", 1)[0]
        assert "BRA.DIV" not in b and "WARPSYNC" not in b, name
        assert " LDL" not in b and " STL" not in b, name
        assert "DFMA" in b and "LDGSTS" in b, name
