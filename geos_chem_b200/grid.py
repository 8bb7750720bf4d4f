"""Synthetic global grids for the chemistry hot path.

No met fields or restart files exist offline, so every cell is derived from the reference's
KPP-standalone sample (tests/golden/beijing_l1_20190701_0040.json) and a deterministic,
shard-invariant counter hash (value depends only on seed, cell index, species/stream), so a
rank that owns any subset of cells generates exactly the values a single process would.

Per cell (cell = I + NX*(J + NY*L), I fastest -- the order of Conc(I,J,L) and of the collapsed
OpenMP loop, fullchem_mod.F90:528-546):
  T      level/latitude profile in [185, 310] K + noise       P  72-level profile 1013 -> 0.01 hPa
  NUMDEN P/(kB T)                                             H2O vmr 3e-2 (surface) .. 3e-6 (aloft)
  cosSZA cos(longitude hour angle)*cos(lat): half the columns are dark
  PHOTOL fixture J-values * max(cosSZA,0)/0.6833
  khet   fixture K_MT/K_CLD/het constants * lognormal(0, 0.5)
  C0     fixture mixing ratios * 10^u, u ~ U(-0.5, 0.5) per variable species per cell;
         fixed species (H2, N2, O2) follow NUMDEN exactly
  hstart "warm": U(100, 700) s (KPPHvalue carried over from a previous step);  "cold": 0
"""
import json
import os

import numpy as np

_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
GOLDEN = os.path.join(_ROOT, "tests", "golden", "beijing_l1_20190701_0040.json")
SEED = 20190701
KB = 1.38064852e-23  # J/K (Headers/physconstants.F90:67)

GRIDS = {
    "4x5": (72, 46, 72),        # 238,464 cells
    "2x2.5": (144, 91, 72),     # 943,488 cells
    "c180": (6 * 180, 180, 72),  # 13,996,800 cells (cubed-sphere equivalent)
}


def load_fixture(path=GOLDEN):
    with open(path) as f:
        g = json.load(f)
    out = dict(g)
    out["C"] = np.array([float(x) for x in g["C"]])
    out["ATOL"] = np.array([float(x) for x in g["ATOL"]])
    out["R"] = np.array([float(x) for x in g["R"]])
    out["A"] = np.array([float(x) for x in g["A"]])
    return out


def _hash_u01(seed, cell, k, stream):
    """counter hash -> uniform in (0,1): splitmix64 finaliser over (seed, cell, k, stream)"""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
             + cell.astype(np.uint64) * np.uint64(0xBF58476D1CE4E5B9)
             + np.uint64(k) * np.uint64(0x94D049BB133111EB)
             + np.uint64(stream) * np.uint64(0xD6E8FEB86659FD93))
        z ^= z >> np.uint64(30); z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27); z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
        z ^= z >> np.uint64(33); z *= np.uint64(0xFF51AFD7ED558CCD)
        z ^= z >> np.uint64(33)
    return ((z >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def _normal(seed, cell, k, stream):
    u1 = _hash_u01(seed, cell, k, stream)
    u2 = _hash_u01(seed, cell, k, stream + 1)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def rate_layout(mech="fullchem"):
    """(gas, phot[(r,k)], ext, null, nphot) index lists of Update_RCONST for a mechanism"""
    from .kppgen import ir, rates
    return rates.rate_layout(ir.load(mech))


def fixture_inputs(fx, mech="fullchem"):
    """PHOTOL[nphot] and khet[next] implied by the fixture's R vector."""
    gas, phot, ext, null, nphot = rate_layout(mech)
    photol = np.zeros(nphot)
    for r, k in phot:
        photol[k] = fx["R"][r]
    khet = np.array([fx["R"][r] for r in ext])
    return photol, khet


def make_cells(cells, shape, hstart="warm", seed=SEED, fx=None):
    """Inputs for the given linear cell indices of an NX x NY x NZ grid. Returns a dict of
    cell-fastest float64 arrays: conc [356, n], temp/numden/h2o/press/cossza/hstart [n],
    photol [177, n], khet [113, n], plus atol/rtol [353]."""
    fx = fx or load_fixture()
    NX, NY, NZ = shape
    cells = np.asarray(cells, dtype=np.int64)
    n = cells.shape[0]
    I = cells % NX
    J = (cells // NX) % NY
    Lv = cells // (NX * NY)
    lat = (-90.0 + 180.0 * (J + 0.5) / NY) * np.pi / 180.0
    lon = 2.0 * np.pi * (I + 0.5) / NX
    # pressure: 72-level hybrid-like profile, dense near the surface, 1013 -> 0.01 hPa
    eta = 1.0 - (Lv + 0.5) / NZ
    press = 0.01 + (1013.25 - 0.01) * eta ** 3.2
    zstar = 7.0 * np.log(1013.25 / press)  # log-pressure height, km
    tsfc = 302.0 - 42.0 * np.sin(lat) ** 2
    temp = np.where(zstar < 12.0, tsfc - 6.5 * zstar * (tsfc - 217.0) / 78.0,
            np.where(zstar < 20.0, 217.0,
             np.where(zstar < 47.0, 217.0 + 2.0 * (zstar - 20.0),
              np.where(zstar < 51.0, 271.0, 271.0 - 2.8 * (zstar - 51.0)))))
    temp = temp + 3.0 * (2.0 * _hash_u01(seed, cells, 0, 11) - 1.0)
    temp = np.clip(temp, 185.0, 310.0)
    numden = press * 100.0 / (KB * temp) * 1.0e-6
    vmr = 3.0e-2 * (press / 1013.25) ** 3 * np.exp(0.5 * _normal(seed, cells, 0, 21))
    vmr = np.clip(vmr, 3.0e-6, 3.0e-2)
    h2o = vmr * numden
    cossza = np.cos(lon) * np.cos(lat)
    # rate inputs
    photol0, khet0 = fixture_inputs(fx)
    sun = np.maximum(cossza, 0.0) / fx["cosSZA"]
    photol = photol0[:, None] * sun[None, :]
    khet = np.empty((khet0.shape[0], n))
    for k in range(khet0.shape[0]):
        khet[k] = khet0[k] * np.exp(0.5 * _normal(seed, cells, k, 31))
    # concentrations: fixture mixing ratios, perturbed by up to half a decade per species
    nspec = fx["C"].shape[0]
    nvar = nspec - 3
    scale = numden / fx["numden"]
    conc = np.empty((nspec, n))
    for s in range(nvar):
        u = _hash_u01(seed, cells, s, 41) - 0.5
        conc[s] = fx["C"][s] * scale * 10.0 ** u
    for s in range(nvar, nspec):
        conc[s] = fx["C"][s] * scale
    if hstart == "warm":
        hs = 100.0 + 600.0 * _hash_u01(seed, cells, 0, 51)
    elif hstart == "cold":
        hs = np.zeros(n)
    else:
        raise ValueError("hstart must be 'warm' or 'cold'")
    atol = fx["ATOL"][:nvar].copy()
    rtol = np.full(nvar, 0.5e-2)
    icntrl = np.zeros(20, np.int32)
    icntrl[0], icntrl[2], icntrl[6], icntrl[14] = 1, 4, 1, -1   # fullchem_AutoReduceFuncs.F90:241-260
    rcntrl = np.zeros(20)
    return dict(conc=conc, temp=temp, numden=numden, h2o=h2o, press=press, cossza=cossza, photol=photol,
                khet=khet, hstart=hs, atol=atol, rtol=rtol, icntrl=icntrl, rcntrl=rcntrl, cells=cells)


def column_shard(shape, rank, world):
    """Linear cell indices owned by `rank`: (I,J) columns dealt round-robin over ranks (every L of a
    column stays on one GPU; interleaving balances day and night columns), in ascending cell order."""
    NX, NY, NZ = shape
    cols = np.arange(NX * NY, dtype=np.int64)
    mine = cols[cols % world == rank]
    cells = (mine[None, :] + (NX * NY) * np.arange(NZ, dtype=np.int64)[:, None]).reshape(-1)
    return cells


def make_grid(name="4x5", hstart="warm", rank=0, world=1, seed=SEED, limit=None):
    shape = GRIDS[name] if isinstance(name, str) else tuple(name)
    cells = column_shard(shape, rank, world)
    if limit is not None:
        cells = cells[:limit]
    g = make_cells(cells, shape, hstart=hstart, seed=seed)
    g["shape"] = shape
    return g


def make_small_mech(mech, cells, seed=SEED):
    """Config 5: inputs of the carbon and Hg mechanisms for the given cells.  The reference ships no sample for
    them, so concentrations and rate constants are log-uniform in ranges that keep the systems stiff but
    integrable (documented here, shard-invariant like make_cells):
      Hg      C0 in 1e2..1e8 molec/cm3, RCONST in 1e-16..1e-11, one 3600 s step, RTOL 1e-2 (mercury_mod.F90:951-1167)
      carbon  C0 in 1e4..1e12 molec/cm3, RCONST in 1e-16..1e-9, one 3600 s forward-Euler step"""
    from .kppgen import ir
    m = ir.load(mech)
    cells = np.asarray(cells, dtype=np.int64)
    n = cells.shape[0]
    lo_c, hi_c, lo_r, hi_r = (2.0, 8.0, -16.0, -11.0) if mech == "Hg" else (4.0, 12.0, -16.0, -9.0)
    conc = np.empty((m.nspec, n))
    for s in range(m.nspec):
        conc[s] = 10.0 ** (lo_c + (hi_c - lo_c) * _hash_u01(seed, cells, s, 61))
    rconst = np.empty((m.nreact, n))
    for r in range(m.nreact):
        rconst[r] = 10.0 ** (lo_r + (hi_r - lo_r) * _hash_u01(seed, cells, r, 71))
    icntrl = np.zeros(20, np.int32)
    icntrl[0], icntrl[2], icntrl[6], icntrl[14] = 1, 4, 1, -1
    return dict(conc=conc, rconst=rconst, atol=np.full(m.nvar, 1e-2), rtol=np.full(m.nvar, 1e-2), icntrl=icntrl,
                rcntrl=np.zeros(20), cells=cells, dt=3600.0)


def replicate_fixture(n, fx=None):
    """Config 1: the fixture cell replicated n times (zero divergence)."""
    fx = fx or load_fixture()
    nvar = fx["C"].shape[0] - 3
    conc = np.repeat(fx["C"][:, None], n, axis=1)
    rconst = np.repeat(fx["R"][:, None], n, axis=1)
    icntrl = np.array(fx["ICNTRL"], np.int32)
    icntrl[11] = 0  # the fixture run had auto-reduce off (ICNTRL(12)=0); ICNTRL(14) is only read when it is on
    rcntrl = np.array(fx["RCNTRL"], np.float64)
    rcntrl[2] = fx["Hstart"]
    return dict(conc=conc, rconst=rconst, atol=fx["ATOL"][:nvar].copy(), rtol=np.full(nvar, 0.5e-2),
                icntrl=icntrl, rcntrl=rcntrl, hstart=np.full(n, fx["Hstart"]), dt=fx["OperatorTimestep"])


# species the auto-reduce solver never removes (fullchem_AutoReduce_KeepHalogensActive,
# KPP/fullchem/fullchem_AutoReduceFuncs.F90:69-108; Shen et al. 2020 GMD Table 1)
KEEP_ACTIVE_HALOGENS = ["AERI", "Br", "Br2", "BrCl", "BrNO2", "BrNO3", "BrO", "BrSALA", "BrSALC", "HBr", "HOBr", "Cl",
                        "Cl2", "Cl2O2", "ClNO2", "ClNO3", "ClO", "ClOO", "OClO", "HCl", "HOCl", "I", "I2", "IO", "I2O2",
                        "HI", "ISALA", "ISALC", "I2O4", "I2O3", "INO", "IONO", "IONO2", "ICl", "IBr", "HOI", "SALACl",
                        "SALCCl", "SALAAL", "SALCAL"]


def keep_active_indices(names):
    """0-based indices of KEEP_ACTIVE_HALOGENS in SPC_NAMES (Fortran identifiers are case-insensitive)"""
    up = {n.upper(): i for i, n in enumerate(names)}
    return [up[k.upper()] for k in KEEP_ACTIVE_HALOGENS]
