#!/usr/bin/env python3
"""Turn the reference's only known-answer vector for the KPP hot path into a committed fixture.

Source : /root/reference/KPP/standalone/Beijing_L1_20190701_0040.txt  (written by a 3-D
         GEOS-Chem run through GeosCore/kppsa_interface_mod.F90:569-843)
Output : tests/golden/beijing_l1_20190701_0040.json
Holds  : header metadata, ICNTRL/RCNTRL, the 3-D model's own answer (12 internal steps,
         Hexit 497.8023 s), C0/ATOL for the 356 species, R1..R1058, A1..A1058.
Numbers are kept as the file's decimal strings so nothing is lost in conversion.
Run in the build container only (the GPU box has no /root/reference).
"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from geos_chem_b200.sample import read_sample

REF = os.environ.get("GEOSCHEM_REF", "/root/reference")
src = os.path.join(REF, "KPP/standalone/Beijing_L1_20190701_0040.txt")
s = read_sample(src)
keep = {k: s[k] for k in ("level", "cosSZA", "Hstart", "Hexit", "fileTotSteps", "OperatorTimestep",
                          "pressure_hPa", "temperature_K", "numden", "h2o_vmr", "cloud_fraction",
                          "longitude", "latitude", "location", "timestamp", "ICNTRL", "RCNTRL", "names")}
keep["C"] = s["C_str"]; keep["ATOL"] = ["%.2E" % a for a in s["ATOL"]]
keep["R"] = s["R_str"]; keep["A"] = s["A_str"]
keep["source"] = "geoschem/geos-chem KPP/standalone/Beijing_L1_20190701_0040.txt via tools/make_golden.py"
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "beijing_l1_20190701_0040.json")
with open(dst, "w") as f:
    json.dump(keep, f, indent=0, separators=(",", ":"))
    f.write("\n")
print(len(keep["C"]), len(keep["R"]), len(keep["A"]), keep["fileTotSteps"], keep["Hexit"], os.path.getsize(dst))
