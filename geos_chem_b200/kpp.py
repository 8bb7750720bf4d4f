"""Host-side mirror of the reference's KPP interface over the C ABI (include/gckpp_gpu.h).

Names and argument meaning follow the reference so callers (and the parity tests) read like
the Fortran they replace:

  Update_RCONST   KPP/<mech>/gckpp_Rates.F90:408
  Integrate       KPP/<mech>/gckpp_Integrator.F90:80      (TIN, TOUT, ICNTRL_U, RCNTRL_U -> ISTATUS, RSTATUS, IERR)
  Fun             KPP/<mech>/gckpp_Function.F90:51         (V, F, RCT -> Vdot, Aout)
  Jac_SP / KppDecomp / KppSolve   gckpp_Jacobian.F90:48, gckpp_LinearAlgebra.F90:46, :644

The module variables the Fortran passes implicitly (C, RCONST, TEMP, NUMDEN, H2O, PHOTOL, ATOL,
RTOL) are explicit arrays over cells here, CELL-FASTEST ([k, ncell], C-contiguous).
numpy arrays go through the host entry points (copies inside the call); torch CUDA tensors
go through the *_device entry points (no copies).  There is no CPU implementation: if the
CUDA library is missing or no GPU is present, construction fails.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GCKPP_B200_LIB") or os.path.join(_HERE, "libgckpp_b200.so")
MECH_ID = {"fullchem": 0, "Hg": 1, "carbon": 2}

# ISTATUS / RSTATUS slots (gckpp_Integrator.F90:57-63)
Nfun, Njac, Nstp, Nacc, Nrej, Ndec, Nsol, Nsng = range(8)
Ntexit, Nhexit, Nhnew, NARthr = range(4)

_lib = None


class KppError(RuntimeError):
    pass


def load_library(path=None):
    """dlopen libgckpp_b200.so and declare the C ABI. Raises if the library was not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise KppError("%s not found: build it with `python -m geos_chem_b200.build` "
                       "(__graft_entry__.build()); there is no CPU fallback" % path)
    L = C.CDLL(path)
    vp, ip, dp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)
    L.gckpp_gpu_dims.argtypes = [C.c_int, ip]
    L.gckpp_gpu_spc_name.argtypes = [C.c_int, C.c_int]
    L.gckpp_gpu_spc_name.restype = C.c_char_p
    L.gckpp_gpu_init.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.gckpp_gpu_finalize.argtypes = [vp]
    L.gckpp_gpu_set_option.argtypes = [vp, C.c_char_p, C.c_int]
    sig = [vp, C.c_int, C.c_double, C.c_double] + [vp] * 17
    L.gckpp_gpu_integrate.argtypes = sig
    L.gckpp_gpu_integrate_device.argtypes = sig
    L.gckpp_gpu_update_rconst.argtypes = [vp, C.c_int] + [vp] * 6
    L.gckpp_gpu_update_rconst_device.argtypes = [vp, C.c_int] + [vp] * 6
    L.gckpp_gpu_fun.argtypes = [vp, C.c_int] + [vp] * 4
    L.gckpp_gpu_jac.argtypes = [vp, C.c_int] + [vp] * 3
    L.gckpp_gpu_decomp.argtypes = [vp, C.c_int] + [vp] * 2
    L.gckpp_gpu_solve.argtypes = [vp, C.c_int] + [vp] * 2
    L.gckpp_gpu_last_stats.argtypes = [vp, dp]
    L.gckpp_gpu_set_stream.argtypes = [vp, vp]
    L.gckpp_gpu_fp64_peak.argtypes = [C.c_int, dp, dp]
    L.gckpp_gpu_last_error.restype = C.c_char_p
    L.gckpp_gpu_plan_info.argtypes = [C.c_int, ip]
    L.gckpp_gpu_set_keep_active.argtypes = [vp, C.c_int, ip]
    L.gckpp_gpu_warp_plan.argtypes = [C.c_int, ip, C.c_int, vp, C.c_int64]
    L.gckpp_gpu_set_sr_mw.argtypes = [vp, C.c_int, vp]
    L.gckpp_gpu_set_het.argtypes = [vp, vp, vp]
    L.gckpp_gpu_set_species_data.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    for sfx in ("", "_device"):
        getattr(L, "gckpp_gpu_zero_species" + sfx).argtypes = [vp, C.c_int, vp, C.c_int, vp]
        getattr(L, "gckpp_gpu_post_integrate" + sfx).argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp, vp]
        getattr(L, "gckpp_gpu_prod_loss" + sfx).argtypes = [vp, C.c_int, vp, C.c_double, C.c_int, vp, vp]
        getattr(L, "gckpp_gpu_oh_reactivity" + sfx).argtypes = [vp, C.c_int, vp, vp, vp]
    _lib = L
    return L


EXPORTS = ["gckpp_gpu_dims", "gckpp_gpu_spc_name", "gckpp_gpu_init", "gckpp_gpu_finalize", "gckpp_gpu_set_option",
           "gckpp_gpu_integrate", "gckpp_gpu_integrate_device", "gckpp_gpu_update_rconst",
           "gckpp_gpu_update_rconst_device", "gckpp_gpu_fun", "gckpp_gpu_jac", "gckpp_gpu_decomp",
           "gckpp_gpu_solve", "gckpp_gpu_last_stats", "gckpp_gpu_last_error", "gckpp_gpu_set_stream",
           "gckpp_gpu_fp64_peak", "gckpp_gpu_plan_info", "gckpp_gpu_set_keep_active", "gckpp_gpu_warp_plan",
           "gckpp_gpu_zero_species", "gckpp_gpu_zero_species_device", "gckpp_gpu_post_integrate",
           "gckpp_gpu_post_integrate_device", "gckpp_gpu_prod_loss", "gckpp_gpu_prod_loss_device",
           "gckpp_gpu_oh_reactivity", "gckpp_gpu_oh_reactivity_device", "gckpp_gpu_set_sr_mw", "gckpp_gpu_set_het",
           "gckpp_gpu_set_species_data"]


def plan_info(mech):
    """static plan of the shared-memory kernel (host-only query)"""
    L = load_library()
    d = (C.c_int32 * 8)()
    if L.gckpp_gpu_plan_info(MECH_ID[mech], d) != 0:
        raise KppError(L.gckpp_gpu_last_error().decode())
    return dict(zip(("smem_bytes", "stream_rows", "resident_rows", "rounds", "n_lu", "n_fwd", "n_bwd", "cells_per_block"),
                    (int(x) for x in d)))


def warp_plan(mech):
    """per-warp table streams of the warp-group kernel as the host plan lays them out (host-only query)"""
    L = load_library()
    info = (C.c_int32 * 256)()
    if L.gckpp_gpu_warp_plan(MECH_ID[mech], info, 256, None, 0) != 0:
        raise KppError(L.gckpp_gpu_last_error().decode())
    wg, cpb, smem, rows, nseg, nph, rs, per = (int(x) for x in info[:8])
    stream = np.zeros((rows, 32, 4), np.uint32)
    if L.gckpp_gpu_warp_plan(MECH_ID[mech], info, 256, stream.ctypes.data_as(C.c_void_p), stream.size) != 0:
        raise KppError(L.gckpp_gpu_last_error().decode())
    nlev = [int(x) for x in info[8:8 + nph]]
    warps = []
    for w in range(wg):
        o = [int(x) for x in info[8 + nph + w * per: 8 + nph + (w + 1) * per]]
        warps.append(dict(off=o[0], rows=o[1], seg_off=o[2:2 + nseg], nb=o[2 + nseg:2 + nseg + nph]))
    return dict(wg=wg, cells_per_block=cpb, smem_bytes=smem, ring_slots=rs, stream=stream, warps=warps, nlev=nlev)


def mech_dims(mech):
    L = load_library()
    d = (C.c_int32 * 7)()
    if L.gckpp_gpu_dims(MECH_ID[mech], d) != 0:
        raise KppError(L.gckpp_gpu_last_error().decode())
    return dict(zip(("nvar", "nfix", "nspec", "nreact", "lu_nonzero", "nphot", "next"), (int(x) for x in d)))


def spc_names(mech):
    L = load_library()
    n = mech_dims(mech)["nspec"]
    return [L.gckpp_gpu_spc_name(MECH_ID[mech], i).decode() for i in range(n)]


def _is_torch(x):
    return x is not None and type(x).__module__.startswith("torch")


def _np(x, dtype, shape=None, name="array"):
    if x is None:
        return None
    a = np.ascontiguousarray(x, dtype=dtype)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError("%s: expected shape %s, got %s" % (name, tuple(shape), tuple(a.shape)))
    return a


def _ptr(a):
    if a is None:
        return None
    if _is_torch(a):
        return C.c_void_p(a.data_ptr())
    return a.ctypes.data_as(C.c_void_p)


class KppSolver:
    """One solver instance = one mechanism on one GPU (gckpp_gpu_init ... gckpp_gpu_finalize)."""

    def __init__(self, mech="fullchem", device=0, max_cells=1 << 20, retry=False):
        self.L = load_library()
        self.mech = mech
        self.dims = mech_dims(mech)
        self.device = device
        h = C.c_void_p()
        rc = self.L.gckpp_gpu_init(MECH_ID[mech], device, max_cells, C.byref(h))
        if rc != 0:
            raise KppError("gckpp_gpu_init failed (%d): %s" % (rc, self.L.gckpp_gpu_last_error().decode()))
        self.h = h
        self._user_stream = False
        if retry:
            self.set_option("retry", 1)

    def close(self):
        if getattr(self, "h", None):
            self.L.gckpp_gpu_finalize(self.h)
            self.h = None

    __del__ = close

    def set_option(self, key, value):
        rc = self.L.gckpp_gpu_set_option(self.h, key.encode(), int(value))
        if rc != 0:
            raise KppError(self.L.gckpp_gpu_last_error().decode())

    def set_keep_active(self, idx0):
        """keepSpcActive of the auto-reduce solver: 0-based variable-species indices that are never removed
        (fullchem_AutoReduceFuncs.F90:40-140); an empty list switches keepActive off"""
        a = np.ascontiguousarray(idx0, np.int32)
        rc = self.L.gckpp_gpu_set_keep_active(self.h, int(a.size), a.ctypes.data_as(C.POINTER(C.c_int32)))
        if rc != 0:
            raise KppError(self.L.gckpp_gpu_last_error().decode())

    def set_stream(self, cuda_stream_ptr):
        """run on a caller-owned CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); None = the default:
        torch tensors run on torch's current stream, host arrays on the handle's own stream"""
        self._user_stream = bool(cuda_stream_ptr)
        self.L.gckpp_gpu_set_stream(self.h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None)

    def _torch_stream(self, *tensors):
        """device entry points: every array must be a CUDA tensor of this device, and the work is ordered after the
        kernels that produced them -- it runs on torch's current stream unless set_stream() chose one"""
        import torch
        for t in tensors:
            if t is None:
                continue
            if not _is_torch(t) or not t.is_cuda or t.device.index != self.device:
                raise ValueError("device entry point: every array must be a CUDA tensor on cuda:%d" % self.device)
            if not t.is_contiguous():
                raise ValueError("device entry point: tensors must be contiguous")
        if not self._user_stream:
            self.L.gckpp_gpu_set_stream(self.h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))

    def _check(self, rc, what):
        if rc < 0:
            raise KppError("%s failed (%d): %s" % (what, rc, self.L.gckpp_gpu_last_error().decode()))
        return rc

    # ------------------------------------------------------------------ Update_RCONST
    def Update_RCONST(self, TEMP, NUMDEN, H2O, PHOTOL=None, khet=None, out=None):
        """RCONST[NREACT, ncell] from per-cell TEMP, NUMDEN, H2O [ncell], PHOTOL [NPHOT, ncell] and the
        externally supplied constants khet [NEXT, ncell] (K_MT, K_CLD, State_Het laws)."""
        d = self.dims
        if _is_torch(TEMP):
            import torch
            ncell = TEMP.shape[0]
            if out is None:
                out = torch.empty((d["nreact"], ncell), dtype=torch.float64, device=TEMP.device)
            self._torch_stream(TEMP, NUMDEN, H2O, PHOTOL, khet, out)
            rc = self.L.gckpp_gpu_update_rconst_device(self.h, ncell, _ptr(TEMP), _ptr(NUMDEN), _ptr(H2O),
                                                       _ptr(PHOTOL), _ptr(khet), _ptr(out))
            self._check(rc, "Update_RCONST")
            return out
        TEMP = _np(TEMP, np.float64)
        ncell = TEMP.shape[0]
        NUMDEN = _np(NUMDEN, np.float64, (ncell,), "NUMDEN")
        H2O = _np(H2O, np.float64, (ncell,), "H2O")
        PHOTOL = _np(PHOTOL, np.float64, (d["nphot"], ncell), "PHOTOL")
        khet = _np(khet, np.float64, (d["next"], ncell), "khet")
        out = np.empty((d["nreact"], ncell), np.float64)
        rc = self.L.gckpp_gpu_update_rconst(self.h, ncell, _ptr(TEMP), _ptr(NUMDEN), _ptr(H2O), _ptr(PHOTOL),
                                            _ptr(khet), _ptr(out))
        self._check(rc, "Update_RCONST")
        return out

    # ------------------------------------------------------------------ Integrate
    def Integrate(self, TIN, TOUT, C_in, RCONST=None, ATOL=None, RTOL=None, ICNTRL_U=None, RCNTRL_U=None,
                  hstart=None, active=None, TEMP=None, NUMDEN=None, H2O=None, PHOTOL=None, khet=None,
                  C_out=None, ISTATUS=None, RSTATUS=None, IERR=None):
        """Batched Integrate.  C_in [NSPEC, ncell]; RCONST [NREACT, ncell] or None (then TEMP, NUMDEN,
        H2O [, PHOTOL, khet] are required and Update_RCONST runs on the device first).
        Returns (C_out, ISTATUS[8, ncell], RSTATUS[4, ncell], IERR[ncell], n_failed_twice)."""
        d = self.dims
        ic = _np(ICNTRL_U if ICNTRL_U is not None else np.zeros(20), np.int32, (20,), "ICNTRL_U")
        rcn = _np(RCNTRL_U if RCNTRL_U is not None else np.zeros(20), np.float64, (20,), "RCNTRL_U")
        at = _np(ATOL, np.float64, (d["nvar"],), "ATOL")
        rt = _np(RTOL, np.float64, (d["nvar"],), "RTOL")
        if _is_torch(C_in):
            import torch
            ncell = C_in.shape[1]
            dev = C_in.device
            assert C_in.dtype == torch.float64 and C_in.is_contiguous()
            if C_out is None:
                C_out = torch.empty_like(C_in)
            if ISTATUS is None:
                ISTATUS = torch.empty((8, ncell), dtype=torch.int32, device=dev)
            if RSTATUS is None:
                RSTATUS = torch.empty((4, ncell), dtype=torch.float64, device=dev)
            if IERR is None:
                IERR = torch.empty((ncell,), dtype=torch.int32, device=dev)
            self._torch_stream(C_in, RCONST, TEMP, NUMDEN, H2O, PHOTOL, khet, hstart, active, C_out, ISTATUS, RSTATUS, IERR)
            fn = self.L.gckpp_gpu_integrate_device
        else:
            C_in = _np(C_in, np.float64)
            if C_in.ndim != 2 or C_in.shape[0] != d["nspec"]:
                raise ValueError("C_in: expected [%d, ncell], got %s" % (d["nspec"], C_in.shape))
            ncell = C_in.shape[1]
            RCONST = _np(RCONST, np.float64, (d["nreact"], ncell), "RCONST")
            TEMP = _np(TEMP, np.float64, (ncell,), "TEMP")
            NUMDEN = _np(NUMDEN, np.float64, (ncell,), "NUMDEN")
            H2O = _np(H2O, np.float64, (ncell,), "H2O")
            PHOTOL = _np(PHOTOL, np.float64, (d["nphot"], ncell), "PHOTOL")
            khet = _np(khet, np.float64, (d["next"], ncell), "khet")
            hstart = _np(hstart, np.float64, (ncell,), "hstart")
            active = _np(active, np.uint8, (ncell,), "active")
            C_out = np.empty_like(C_in)
            ISTATUS = np.zeros((8, ncell), np.int32)
            RSTATUS = np.zeros((4, ncell), np.float64)
            IERR = np.zeros((ncell,), np.int32)
            fn = self.L.gckpp_gpu_integrate
        rc = fn(self.h, ncell, float(TIN), float(TOUT), _ptr(C_in), _ptr(RCONST), _ptr(TEMP), _ptr(NUMDEN),
                _ptr(H2O), _ptr(PHOTOL), _ptr(khet), _ptr(at), _ptr(rt), _ptr(ic), _ptr(rcn), _ptr(hstart),
                _ptr(active), _ptr(C_out), _ptr(ISTATUS), _ptr(RSTATUS), _ptr(IERR))
        if rc < 0 and rc >= -5:
            # option errors are per-call in the reference too: every cell reports the same IERR
            return C_out, ISTATUS, RSTATUS, IERR, rc
        self._check(rc, "Integrate")
        return C_out, ISTATUS, RSTATUS, IERR, rc

    # ------------------------------------------------------------------ diagnostics / pieces (host arrays)
    def Fun(self, C_in, RCONST):
        d = self.dims
        C_in = _np(C_in, np.float64)
        ncell = C_in.shape[1]
        RCONST = _np(RCONST, np.float64, (d["nreact"], ncell), "RCONST")
        vdot = np.empty((d["nvar"], ncell), np.float64)
        aout = np.empty((d["nreact"], ncell), np.float64)
        self._check(self.L.gckpp_gpu_fun(self.h, ncell, _ptr(C_in), _ptr(RCONST), _ptr(vdot), _ptr(aout)), "Fun")
        return vdot, aout

    def Jac_SP(self, C_in, RCONST):
        d = self.dims
        C_in = _np(C_in, np.float64)
        ncell = C_in.shape[1]
        RCONST = _np(RCONST, np.float64, (d["nreact"], ncell), "RCONST")
        jvs = np.empty((d["lu_nonzero"], ncell), np.float64)
        self._check(self.L.gckpp_gpu_jac(self.h, ncell, _ptr(C_in), _ptr(RCONST), _ptr(jvs)), "Jac_SP")
        return jvs

    def KppDecomp(self, JVS):
        d = self.dims
        j = _np(JVS, np.float64).copy()
        ncell = j.shape[1]
        ier = np.zeros(ncell, np.int32)
        self._check(self.L.gckpp_gpu_decomp(self.h, ncell, _ptr(j), _ptr(ier)), "KppDecomp")
        return j, ier

    def KppSolve(self, JVS, X):
        j = _np(JVS, np.float64)
        x = _np(X, np.float64).copy()
        self._check(self.L.gckpp_gpu_solve(self.h, j.shape[1], _ptr(j), _ptr(x)), "KppSolve")
        return x

    # ------------------------------------------------------------------ heterogeneous laws on the device
    NHET = 104
    # rows 0-47: first part; rows 48-103: second part (read only after set_species_data); order = include/gckpp_gpu.h
    HET_FIELDS = ("SUNCOS", "stratBox", "SSA_is_Alk", "SSA_is_Acid", "SSC_is_Alk", "SSC_is_Acid", "f_Alk_SSA", "f_Alk_SSC",
                  "f_Acid_SSA", "f_Acid_SSC", "ClearFr", "aClArea", "aClRadi", "Cl_conc_SSA", "Cl_conc_SSC", "gamma_HO2",
                  "H_PLUS", "NO3_molal", "SO4_molal", "HSO4_molal") + tuple("xArea%d" % k for k in range(1, 15)) + \
        tuple("xRadi%d" % k for k in range(1, 15)) + \
        ("natSurface", "TurnOffHetRates", "CldFr", "aIce", "aLiq", "rIce", "rLiq", "pHCloud", "pHSSA1", "pHSSA2",
         "Cl_conc_Cld", "Br_conc_Cld", "Br_conc_SSA", "Br_conc_SSC", "Br_over_Cl_Cld", "Br_over_Cl_SSA", "Br_over_Cl_SSC",
         "frac_Br_CldA", "frac_Br_CldC", "frac_Br_CldG", "frac_Cl_CldA", "frac_Cl_CldC", "frac_Cl_CldG", "frac_SALACL",
         "frac_HSO3_aq", "HSO3m", "HCl_theta", "HBr_theta", "HNO3_theta", "H_conc_LCl", "H_conc_SSA", "H_conc_SSC",
         "HSO3_aq", "SO3_aq", "TSO3_aq", "aWater1", "aWater2") + tuple("KHETI_SLA%d" % k for k in range(1, 12)) + \
        ("AClVol", "xVol_ORC", "xVol_SSC", "xH2O_SUL", "xH2O_ORC", "xH2O_SSC", "OMOC_POA", "OMOC_OPOA")

    def set_species_data(self, sr_mw, mw, henry_k0, henry_cr):
        """set_sr_mw plus MW, HENRY_K0, HENRY_CR (1:NSPEC) of gckpp_Global: switches the second part of the device-side
        heterogeneous laws on (cloud / halogen uptake, fullchem_RateLawFuncs.F90:803-3238)"""
        n = self.dims["nspec"]
        a = [_np(x, np.float64, (n,), nm) for x, nm in ((sr_mw, "sr_mw"), (mw, "mw"), (henry_k0, "henry_k0"), (henry_cr, "henry_cr"))]
        self._check(self.L.gckpp_gpu_set_species_data(self.h, n, *[_ptr(x) for x in a]), "set_species_data")

    def set_sr_mw(self, sr_mw):
        """SR_MW(1:NSPEC) of gckpp_Global (SQRT of the molecular weights); None switches the device-side laws off"""
        if sr_mw is None:
            self._check(self.L.gckpp_gpu_set_sr_mw(self.h, 0, None), "set_sr_mw")
            return
        a = _np(sr_mw, np.float64, (self.dims["nspec"],), "sr_mw")
        self._check(self.L.gckpp_gpu_set_sr_mw(self.h, a.size, _ptr(a)), "set_sr_mw")

    def set_het(self, het, conc=None):
        """HetState fields [NHET, ncell] (order: HET_FIELDS) for the next calls; numpy for the host entry points, a CUDA
        tensor for the device ones; None = off.  The arrays must stay alive until those calls have been made."""
        self._het_keep = (het, conc)
        if het is not None and not _is_torch(het):
            het = _np(het, np.float64)
            conc = None if conc is None else _np(conc, np.float64)
            self._het_keep = (het, conc)
        if het is not None and het.shape[0] != self.NHET:
            raise ValueError("het: expected [%d, ncell]" % self.NHET)
        self._check(self.L.gckpp_gpu_set_het(self.h, _ptr(het), _ptr(conc)), "set_het")

    # ------------------------------------------------------------------ Do_FullChem's pieces around the integration
    # numpy arrays are staged through the host entry points (a new array is returned), torch CUDA tensors are
    # modified in place on the device
    def _ids(self, ids0):
        a = np.ascontiguousarray(ids0, np.int32).reshape(-1)
        return a, (a.ctypes.data_as(C.c_void_p) if a.size else None)

    def zero_species(self, C_io, ids0):
        """C(PL_Kpp_Id(F)) = 0 for the prod/loss family species (GeosCore/fullchem_mod.F90:941-946)"""
        ids, pi = self._ids(ids0)
        if _is_torch(C_io):
            self._torch_stream(C_io)
            self._check(self.L.gckpp_gpu_zero_species_device(self.h, C_io.shape[1], _ptr(C_io), ids.size, pi), "zero_species")
            return C_io
        c = _np(C_io, np.float64).copy()
        self._check(self.L.gckpp_gpu_zero_species(self.h, c.shape[1], _ptr(c), ids.size, pi), "zero_species")
        return c

    def post_integrate(self, C_io, scale_ids0=(), scale_div=(), spc_mask=None, negatives=None):
        """fullchem_ConvertEquivToAlk (C(id) = C(id) / div), the KppNegatives count and C = MAX(C, 0) over the species
        with spc_mask != 0 (GeosCore/fullchem_mod.F90:1284-1287, 1326-1348).  Returns (C, negatives)."""
        ids, pi = self._ids(scale_ids0)
        div = np.ascontiguousarray(scale_div, np.float64).reshape(-1)
        if div.size != ids.size:
            raise ValueError("scale_ids0 and scale_div differ in length")
        mask = None if spc_mask is None else np.ascontiguousarray(spc_mask, np.uint8)
        if mask is not None and mask.shape != (self.dims["nspec"],):
            raise ValueError("spc_mask: expected [%d]" % self.dims["nspec"])
        pd = div.ctypes.data_as(C.c_void_p) if div.size else None
        pm = mask.ctypes.data_as(C.c_void_p) if mask is not None else None
        if _is_torch(C_io):
            self._torch_stream(C_io, negatives)
            self._check(self.L.gckpp_gpu_post_integrate_device(self.h, C_io.shape[1], _ptr(C_io), ids.size, pi, pd, pm,
                                                               _ptr(negatives)), "post_integrate")
            return C_io, negatives
        c = _np(C_io, np.float64).copy()
        neg = None if negatives is None else np.ascontiguousarray(negatives, np.float32).copy()
        self._check(self.L.gckpp_gpu_post_integrate(self.h, c.shape[1], _ptr(c), ids.size, pi, pd, pm, _ptr(neg)), "post_integrate")
        return c, neg

    def prod_loss(self, C_in, dt, ids0, out=None):
        """State_Diag%Prod / %Loss(slot) = C(KppId(slot)) / DT (GeosCore/fullchem_mod.F90:1463-1492) -> [nslots, ncell]"""
        ids, pi = self._ids(ids0)
        if _is_torch(C_in):
            import torch
            if out is None:
                out = torch.empty((ids.size, C_in.shape[1]), dtype=torch.float64, device=C_in.device)
            self._torch_stream(C_in, out)
            self._check(self.L.gckpp_gpu_prod_loss_device(self.h, C_in.shape[1], _ptr(C_in), float(dt), ids.size, pi, _ptr(out)), "prod_loss")
            return out
        c = _np(C_in, np.float64)
        out = np.empty((ids.size, c.shape[1]), np.float64)
        self._check(self.L.gckpp_gpu_prod_loss(self.h, c.shape[1], _ptr(c), float(dt), ids.size, pi, _ptr(out)), "prod_loss")
        return out

    def Get_OHreactivity(self, C_in, RCONST, out=None):
        """Get_OHreactivity(CC, RR, OHreact) of KPP/fullchem/gckpp_Util.F90:983-1040 over cells -> [ncell] (1/s)"""
        if _is_torch(C_in):
            import torch
            if out is None:
                out = torch.empty((C_in.shape[1],), dtype=torch.float64, device=C_in.device)
            self._torch_stream(C_in, RCONST, out)
            self._check(self.L.gckpp_gpu_oh_reactivity_device(self.h, C_in.shape[1], _ptr(C_in), _ptr(RCONST), _ptr(out)), "Get_OHreactivity")
            return out
        c = _np(C_in, np.float64)
        rc = _np(RCONST, np.float64, (self.dims["nreact"], c.shape[1]), "RCONST")
        out = np.empty((c.shape[1],), np.float64)
        self._check(self.L.gckpp_gpu_oh_reactivity(self.h, c.shape[1], _ptr(c), _ptr(rc), _ptr(out)), "Get_OHreactivity")
        return out

    def last_stats(self):
        s = (C.c_double * 16)()
        self.L.gckpp_gpu_last_stats(self.h, s)
        keys = ("integrate_ms", "rconst_ms", "copy_ms", "cells", "retried", "failed_twice", "launches", "sum_nstp",
                "sum_nacc", "device_ms", "failed_integrations", "waves", "rconst_launches", "kernel")
        return dict(zip(keys, (float(x) for x in s)))


def fp64_peak(device=0):
    """measured FP64 FMA peak of the device [TFLOP/s] (register-resident DFMA chains on every SM)"""
    L = load_library()
    tf, ms = C.c_double(), C.c_double()
    rc = L.gckpp_gpu_fp64_peak(device, C.byref(tf), C.byref(ms))
    if rc != 0:
        raise KppError(L.gckpp_gpu_last_error().decode())
    return tf.value
