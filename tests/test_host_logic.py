"""Host-side logic: mechanism IR, sample I/O, grid generator, column sharding (incl. a 2-rank gloo run)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from geos_chem_b200 import grid, sample
from geos_chem_b200.kppgen import ir

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
REF_SAMPLE = "/root/reference/KPP/standalone/Beijing_L1_20190701_0040.txt"


def test_ir_counts():
    m = ir.load("fullchem")
    assert (m.nvar, m.nfix, m.nreact, m.lu_nonzero) == (353, 3, 1058, 5683)
    sched = m.lu_schedule()
    assert sum(len(st) for st in sched) == 2773                    # L entries
    assert sum(len(u) for st in sched for _, _, u in st) == 20790  # LU multiply-adds
    assert sum(1 for e in m.JVS if not e) == 1313                  # structural zeros of Jac_SP
    assert sum(len(e) for e in m.P_VAR) == 3862
    arity = [len(e[0].factors) - 1 for e in m.A]
    assert (arity.count(1), arity.count(2), arity.count(3)) == (275, 779, 4)
    # Q1: two reactions carry a literal instead of RCT(r)
    lits = [r for r, e in enumerate(m.A) if e[0].factors[0][0] == "N"]
    assert lits == [125, 734]
    assert m.rconst[125] is None and m.rconst[734] is None
    hg = ir.load("Hg")
    assert (hg.nvar, hg.nreact, hg.lu_nonzero) == (32, 94, 161)
    assert ir.load("carbon").has_jac is False


def test_sample_roundtrip(fx):
    s = {k: fx[k] for k in ("level", "cosSZA", "Hstart", "Hexit", "fileTotSteps", "OperatorTimestep", "pressure_hPa",
                            "temperature_K", "numden", "h2o_vmr", "cloud_fraction", "longitude", "latitude",
                            "location", "timestamp", "ICNTRL", "RCNTRL", "names")}
    s["C"], s["ATOL"], s["R"], s["A"] = list(fx["C"]), list(fx["ATOL"]), list(fx["R"]), list(fx["A"])
    txt = sample.format_sample(s)
    back = sample.parse_sample(txt, spc_names=fx["names"], nreact=1058)
    assert back["C"] == s["C"] and back["R"] == s["R"] and back["A"] == s["A"] and back["ATOL"] == s["ATOL"]
    assert back["ICNTRL"] == s["ICNTRL"] and back["RCNTRL"] == s["RCNTRL"]
    assert back["fileTotSteps"] == 12 and back["Hexit"] == 497.8023
    bad = txt.replace("CH2IBr,", "CH2IBx,")
    with pytest.raises(sample.SampleError):
        sample.parse_sample(bad, spc_names=fx["names"])
    out = sample.format_output(fx["names"][:2], [5579.513118142906, 0.0], [1.0e-300, -2.5])
    assert "CH2I2,  5.5795131181429060E+003,  1.0000000000000000E-300" in out


@pytest.mark.skipif(not os.path.exists(REF_SAMPLE), reason="reference tree not present (GPU box)")
def test_reader_on_the_reference_file(fx):
    s = sample.read_sample(REF_SAMPLE, spc_names=fx["names"], nreact=1058)
    assert np.array_equal(np.array(s["C"]), fx["C"]) and np.array_equal(np.array(s["R"]), fx["R"])
    assert s["ICNTRL"] == fx["ICNTRL"] and s["fileTotSteps"] == 12


def test_grid_is_shard_invariant():
    full = grid.make_grid("4x5", limit=None)
    n = full["conc"].shape[1]
    assert n == 238464
    parts = [grid.column_shard(grid.GRIDS["4x5"], r, 2) for r in range(2)]
    assert sorted(np.concatenate(parts).tolist()) == list(range(n))
    cells = parts[1][::997]
    sub = grid.make_cells(cells, grid.GRIDS["4x5"])
    assert np.array_equal(sub["conc"], full["conc"][:, cells])
    assert np.array_equal(sub["photol"], full["photol"][:, cells])
    assert np.array_equal(sub["hstart"], full["hstart"][cells])
    assert 185.0 <= full["temp"].min() and full["temp"].max() <= 310.0
    assert abs((full["cossza"] <= 0).mean() - 0.5) < 0.01
    # every L of a column lands on one rank
    NX, NY, NZ = grid.GRIDS["4x5"]
    assert len(set((parts[0] % (NX * NY)) % 2)) == 1


def test_two_rank_gloo_sharding(tmp_path):
    """N>1 host path: column sharding + the diagnostics all-reduce, on CPU with gloo, world_size 2"""
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, numpy as np, torch, torch.distributed as dist\n"
        "sys.path.insert(0, %r)\n"
        "from geos_chem_b200 import grid, dist_diag\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "cells = grid.column_shard((8, 4, 3), r, w)\n"
        "ist = np.zeros((8, len(cells)), np.int32); ist[2] = 10 + r; ist[3] = 9; ierr = np.ones(len(cells), np.int32)\n"
        "if r == 1: ierr[0] = -7\n"
        "d = dist_diag.reduce_diagnostics(ist, ierr)\n"
        "if r == 0:\n"
        "    assert d['cells'] == 96 and d['failed'] == 1 and d['max_nstp'] == 11, d\n"
        "    assert d['sum_nstp'] == 48 * 10 + 48 * 11 and d['hist_nstp'][10] == 48 and d['hist_nstp'][11] == 48, d\n"
        "    print('OK')\n"
        "dist.destroy_process_group()\n" % ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
