"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (geos_chem_b200) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MECH_ID = {"fullchem": 0, "Hg": 1, "carbon": 2}
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


LIBS = {"strict": "libkpp_oracle.so", "fma": "libkpp_oracle_fma.so", "native": "libkpp_oracle_native.so"}


def build(variant="strict"):
    """strict: -O2 -march=x86-64-v3 (travels to the GPU box); native: -O3 -march=native, built on the machine that
    runs it (the CPU-baseline arm of bench.py); both without FMA contraction"""
    target = LIBS[variant]
    subprocess.check_call(["make", "-C", HERE, "-j8", target], stdout=subprocess.DEVNULL)
    return os.path.join(HERE, target)


def native_is_stale():
    """the native build must come from this machine's compiler run (never shipped): rebuild when its CPU differs"""
    tag = os.path.join(HERE, "build", "native", "cpu.txt")
    try:
        cpu = [l for l in open("/proc/cpuinfo") if l.startswith("model name")][0].strip()
    except Exception:
        cpu = "unknown"
    lib = os.path.join(HERE, LIBS["native"])
    if os.path.exists(lib) and os.path.exists(tag) and open(tag).read() == cpu:
        return False, tag, cpu
    return True, tag, cpu


class Oracle:
    def __init__(self, variant="strict", autobuild=True):
        self.variant = variant
        if variant == "native":
            stale, tag, cpu = native_is_stale()
            if stale:
                try:
                    subprocess.call(["rm", "-rf", os.path.join(HERE, "build", "native"), os.path.join(HERE, LIBS["native"])])
                    build("native")
                    with open(tag, "w") as f:
                        f.write(cpu)
                except Exception:
                    self.variant = variant = "strict"       # no compiler: fall back to the shipped strict build
        path = os.path.join(HERE, LIBS[variant])
        if not os.path.exists(path):
            if not autobuild:
                raise FileNotFoundError(path)
            build(variant)
        self.lib = L = C.CDLL(path)
        L.kpp_oracle_dims.argtypes = [C.c_int, _ip]
        L.kpp_oracle_integrate_cell.argtypes = [C.c_int, C.c_double, C.c_double, _dp, _dp, _dp, _dp, _ip, _dp, _ip, _dp]
        L.kpp_oracle_integrate.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp, _dp, _dp, _ip, _dp,
                                           C.c_void_p, _dp, _ip, _dp, _ip, C.c_int]
        L.kpp_oracle_update_rconst.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_void_p, C.c_void_p, _dp, C.c_int]
        L.kpp_oracle_fun.argtypes = [C.c_int, _dp, _dp, _dp, _dp]
        L.kpp_oracle_jac.argtypes = [C.c_int, _dp, _dp, _dp]
        L.kpp_oracle_decomp.argtypes = [C.c_int, _dp]
        L.kpp_oracle_solve.argtypes = [C.c_int, _dp, _dp]
        L.kpp_oracle_set_keep_active.argtypes = [C.c_int, C.c_int, _ip]

    def set_keep_active(self, mech, idx0):
        """keepSpcActive of the auto-reduce solver (global, like the Fortran module variable); [] switches it off"""
        a = np.ascontiguousarray(idx0, np.int32)
        if a.size == 0:
            a = np.zeros(1, np.int32)
            n = 0
        else:
            n = a.size
        assert self.lib.kpp_oracle_set_keep_active(MECH_ID[mech], n, a) == 0

    def dims(self, mech="fullchem"):
        d = np.zeros(7, np.int32)
        assert self.lib.kpp_oracle_dims(MECH_ID[mech], d) == 0
        return dict(zip(("nvar", "nfix", "nspec", "nreact", "lu_nonzero", "nphot", "next"), (int(x) for x in d)))

    def integrate_cell(self, mech, tin, tout, conc, rconst, atol, rtol, icntrl, rcntrl):
        c = np.ascontiguousarray(conc, np.float64).copy()
        ist = np.zeros(20, np.int32)
        rst = np.zeros(20, np.float64)
        ierr = self.lib.kpp_oracle_integrate_cell(
            MECH_ID[mech], tin, tout, c, np.ascontiguousarray(rconst, np.float64),
            np.ascontiguousarray(atol, np.float64), np.ascontiguousarray(rtol, np.float64),
            np.ascontiguousarray(icntrl, np.int32), np.ascontiguousarray(rcntrl, np.float64), ist, rst)
        return c, ist, rst, ierr

    def integrate(self, mech, tin, tout, conc, rconst, atol, rtol, icntrl, rcntrl, hstart=None, nthreads=0):
        """conc [nspec, ncell], rconst [nreact, ncell] (cell-fastest) -> conc_out, istatus[8,ncell], rstatus[4,ncell], ierr[ncell]"""
        conc = np.ascontiguousarray(conc, np.float64)
        rconst = np.ascontiguousarray(rconst, np.float64)
        ncell = conc.shape[1]
        out = np.empty_like(conc)
        ist = np.zeros((8, ncell), np.int32)
        rst = np.zeros((4, ncell), np.float64)
        ierr = np.zeros(ncell, np.int32)
        hs = None
        if hstart is not None:
            hstart = np.ascontiguousarray(hstart, np.float64)
            hs = hstart.ctypes.data_as(C.c_void_p)
        rc = self.lib.kpp_oracle_integrate(
            MECH_ID[mech], ncell, tin, tout, conc, rconst, np.ascontiguousarray(atol, np.float64),
            np.ascontiguousarray(rtol, np.float64), np.ascontiguousarray(icntrl, np.int32),
            np.ascontiguousarray(rcntrl, np.float64), hs, out, ist, rst, ierr, nthreads)
        assert rc == 0
        return out, ist, rst, ierr

    def update_rconst(self, mech, temp, numden, h2o, photol=None, khet=None, nthreads=0):
        d = self.dims(mech)
        temp = np.ascontiguousarray(temp, np.float64)
        ncell = temp.shape[0]
        out = np.empty((d["nreact"], ncell), np.float64)
        ph = kh = None
        if photol is not None:
            photol = np.ascontiguousarray(photol, np.float64); assert photol.shape == (d["nphot"], ncell)
            ph = photol.ctypes.data_as(C.c_void_p)
        if khet is not None:
            khet = np.ascontiguousarray(khet, np.float64); assert khet.shape == (d["next"], ncell)
            kh = khet.ctypes.data_as(C.c_void_p)
        rc = self.lib.kpp_oracle_update_rconst(MECH_ID[mech], ncell, temp, np.ascontiguousarray(numden, np.float64),
                                               np.ascontiguousarray(h2o, np.float64), ph, kh, out, nthreads)
        assert rc == 0
        return out

    def fun(self, mech, conc, rconst):
        d = self.dims(mech)
        vdot = np.zeros(d["nvar"]); a = np.zeros(d["nreact"])
        self.lib.kpp_oracle_fun(MECH_ID[mech], np.ascontiguousarray(conc, np.float64), np.ascontiguousarray(rconst, np.float64), vdot, a)
        return vdot, a

    def jac(self, mech, conc, rconst):
        d = self.dims(mech)
        jvs = np.zeros(d["lu_nonzero"])
        self.lib.kpp_oracle_jac(MECH_ID[mech], np.ascontiguousarray(conc, np.float64), np.ascontiguousarray(rconst, np.float64), jvs)
        return jvs

    def decomp(self, mech, jvs):
        j = np.ascontiguousarray(jvs, np.float64).copy()
        ier = self.lib.kpp_oracle_decomp(MECH_ID[mech], j)
        return j, ier

    def solve(self, mech, jvs, x):
        xx = np.ascontiguousarray(x, np.float64).copy()
        self.lib.kpp_oracle_solve(MECH_ID[mech], np.ascontiguousarray(jvs, np.float64), xx)
        return xx
