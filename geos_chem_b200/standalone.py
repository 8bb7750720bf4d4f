"""GPU equivalent of the KPP standalone box model (KPP/standalone/kpp_standalone.F90:97-169, `fullmech`).

  python -m geos_chem_b200.standalone SAMPLE.txt [OUTPUT.txt] [--rtol 0.5e-2] [--kernel 0|1] [--replicate N]

Reads a sample written by KppSa_Write_Samples (kppsa_interface_mod.F90:684-843), integrates it for the
operator timestep through the C ABI (libgckpp_b200.so) exactly as `fullmech` does -- RCNTRL(3) = Hstart,
ATOL < 0 -> 1e-2, RTOL = the command-line value, C = Cinit, RCONST = R -- prints the same four lines and
the same two consistency warnings, and writes the box model's output file (kpp_standalone.F90:171-231).
"""
import argparse
import sys

import numpy as np

from . import kpp, sample


def fullmech(s, rtol_value=0.5e-2, kernel=None, replicate=1, device=0):
    names = kpp.spc_names("fullchem")
    d = kpp.mech_dims("fullchem")
    cinit = np.asarray(s["C"], np.float64)
    atol = np.asarray(s["ATOL"], np.float64)[:d["nvar"]].copy()
    atol[atol < 0.0] = 1.0e-2                                  # kpp_standalone.F90:118-120
    rtol = np.full(d["nvar"], rtol_value)
    rcntrl = np.asarray(s["RCNTRL"], np.float64).copy()
    rcntrl[2] = s["Hstart"]                                    # :114
    icntrl = np.asarray(s["ICNTRL"], np.int32)
    conc = np.repeat(cinit[:, None], replicate, axis=1)
    rconst = np.repeat(np.asarray(s["R"], np.float64)[:, None], replicate, axis=1)
    solver = kpp.KppSolver("fullchem", device=device, max_cells=max(replicate, 1))
    if kernel is not None:
        solver.set_option("kernel", kernel)
    c, ist, rst, ierr, _ = solver.Integrate(0.0, s["OperatorTimestep"], conc, rconst, atol, rtol, icntrl, rcntrl)
    solver.close()
    return names, cinit, c[:, 0], ist[:, 0], rst[:, 0], int(ierr[0])


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("sample")
    ap.add_argument("output", nargs="?")
    ap.add_argument("--rtol", type=float, default=0.5e-2)
    ap.add_argument("--kernel", type=int, default=None)
    ap.add_argument("--replicate", type=int, default=1)
    a = ap.parse_args(argv)
    s = sample.read_sample(a.sample, spc_names=kpp.spc_names("fullchem"), nreact=kpp.mech_dims("fullchem")["nreact"])
    names, cinit, cfinal, ist, rst, ierr = fullmech(s, a.rtol, a.kernel, a.replicate)
    print(" Number of internal timesteps (from 3D run): %5d" % s["fileTotSteps"])
    print(" Number of internal timesteps ( standalone): %5d" % ist[kpp.Nstp])
    print(" Hexit (from 3D run): %10.2f" % s["Hexit"])
    print(" Hexit ( standalone): %10.2f" % rst[kpp.Nhexit])
    ok = True
    if s["fileTotSteps"] != ist[kpp.Nstp]:
        print("Warning: Number of internal steps do not match 3D grid cell"); ok = False
    if abs(s["Hexit"] - rst[kpp.Nhexit]) / s["Hexit"] > 0.001:
        print("Warning: final timestep does not match 3D grid cell within 0.1%"); ok = False
    if ierr != 1:
        print("Integrate returned IERR = %d" % ierr); ok = False
    if a.output:
        with open(a.output, "w") as f:
            f.write(sample.format_output(names, cinit, cfinal))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
