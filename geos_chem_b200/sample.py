"""KPP-standalone sample files: reader and writers.

Input format  = what `KppSa_Write_Samples` writes (GeosCore/kppsa_interface_mod.F90:684-843)
                and `read_input` parses (KPP/standalone/kpp_standalone_init.F90:8-207).
Output format = `write_output` of the box model (KPP/standalone/kpp_standalone.F90:171-231).

The reader applies the same rules as the Fortran one: line 1 is the number of header
lines; header fields are found by substring match and read after the first ':'; after
'ICNTRL integrator options used:' come two lines of (10i6), after 'RCNTRL ...' four lines
of (5F13.6); then NSPEC lines 'Name, value[, ATOL]' whose names must equal SPC_NAMES,
NREACT lines 'Rn, value' and NREACT lines 'An, value' (ignored by the Fortran reader, kept here).
"""
import re

HEADER_FIELDS = {
    "GEOS-Chem Vertical Level:": ("level", int),
    "Cosine of solar zenith angle:": ("cosSZA", float),
    "Init KPP Timestep (seconds):": ("Hstart", float),
    "Exit KPP Timestep (seconds):": ("Hexit", float),
    "Number of internal timesteps:": ("fileTotSteps", int),
    "Chemistry operator timestep (seconds):": ("OperatorTimestep", float),
    # not parsed by the Fortran reader ("for reference" fields), but needed to rebuild met inputs
    "Pressure (hPa):": ("pressure_hPa", float),
    "Temperature (K):": ("temperature_K", float),
    "Dry air density (molec/cm3):": ("numden", float),
    "Water vapor mixing ratio (vol H2O/vol dry air):": ("h2o_vmr", float),
    "Cloud fraction:": ("cloud_fraction", float),
    "Longitude (degrees):": ("longitude", float),
    "Latitude (degrees):": ("latitude", float),
    "Location:": ("location", str),
    "Timestamp:": ("timestamp", str),
}


class SampleError(ValueError):
    pass


def parse_sample(text, spc_names=None, nreact=None):
    """Parse a KPP-standalone sample. Returns a dict with header fields, ICNTRL(20), RCNTRL(20),
    names, C (list of float), ATOL (list; -1.0 where absent, like the Fortran default),
    R, A and the verbatim value strings (C_str, R_str, A_str) so a fixture can be re-emitted exactly."""
    lines = text.split("\n")
    try:
        nheader = int(lines[0].split()[0])
    except (IndexError, ValueError):
        raise SampleError("first line must hold the number of header lines")
    out = {"nheader": nheader, "ICNTRL": [0] * 20, "RCNTRL": [0.0] * 20}
    # defaults the Fortran reader sets before parsing (kpp_standalone_init.F90:60-63)
    out["ICNTRL"][0], out["ICNTRL"][2], out["ICNTRL"][6], out["ICNTRL"][14] = 1, 4, 1, -1
    parse_i = parse_r = False
    i1 = r1 = 0
    for ln in lines[1:1 + nheader]:
        for key, (name, typ) in HEADER_FIELDS.items():
            if key in ln:
                val = ln[ln.index(":") + 1:].strip()
                if key == "Timestamp:":
                    val = ln[ln.index("Timestamp:") + len("Timestamp:"):].strip()
                out[name] = typ(val) if typ is not str else val
        if "ICNTRL integrator options used:" in ln:
            parse_i = True
            continue
        if parse_i:
            vals = [int(ln[6 * k:6 * k + 6]) for k in range(10)]  # (10i6)
            out["ICNTRL"][i1:i1 + 10] = vals
            i1 += 10
            if i1 >= 20:
                parse_i = False
                continue
        if "RCNTRL integrator options used:" in ln:
            parse_r = True
            continue
        if parse_r:
            vals = [float(ln[13 * k:13 * k + 13]) for k in range(5)]  # (5F13.6)
            out["RCNTRL"][r1:r1 + 5] = vals
            r1 += 5
            if r1 >= 20:
                parse_r = False
                continue
    body = [ln for ln in lines[1 + nheader:] if ln.strip()]
    names, cs, cstr, atol = [], [], [], []
    k = 0
    while k < len(body) and not re.match(r"^R\d+,", body[k]):
        parts = [p.strip() for p in body[k].split(",")]
        names.append(parts[0])
        cstr.append(parts[1])
        cs.append(float(parts[1]))
        atol.append(float(parts[2]) if len(parts) > 2 and parts[2] else -1.0)
        k += 1
    if spc_names is not None:
        if len(names) != len(spc_names):
            raise SampleError("expected %d species, found %d" % (len(spc_names), len(names)))
        for i, (a, b) in enumerate(zip(names, spc_names)):
            if a != b:  # the Fortran reader stops here too
                raise SampleError("species name mismatch at %d: expected %s, found %s" % (i + 1, b, a))
    rs, rstr, as_, astr = [], [], [], []
    for ln in body[k:]:
        parts = [p.strip() for p in ln.split(",")]
        tag = parts[0]
        if tag[0] == "R":
            assert int(tag[1:]) == len(rs) + 1
            rstr.append(parts[1]); rs.append(float(parts[1]))
        elif tag[0] == "A":
            assert int(tag[1:]) == len(as_) + 1
            astr.append(parts[1]); as_.append(float(parts[1]))
        else:
            raise SampleError("unexpected line %r" % ln)
    if nreact is not None and len(rs) != nreact:
        raise SampleError("expected %d rate constants, found %d" % (nreact, len(rs)))
    out.update(names=names, C=cs, C_str=cstr, ATOL=atol, R=rs, R_str=rstr, A=as_, A_str=astr)
    return out


def read_sample(path, spc_names=None, nreact=None):
    with open(path) as f:
        return parse_sample(f.read(), spc_names, nreact)


def _es25(x):
    """Fortran ES25.16E3: 25 wide, 16 decimals, 3-digit exponent"""
    s = "%.16E" % x
    mant, exp = s.split("E")
    sign, digits = exp[0], exp[1:].lstrip("0") or "0"
    return ("%sE%s%03d" % (mant, sign, int(digits))).rjust(25)


def format_sample(s):
    """Write the sample format of KppSa_Write_Samples (kppsa_interface_mod.F90:684-843):
    header block, then 'Name, value, ATOL' / 'Rn, value' / 'An, value' lines."""
    hdr = []
    hdr.append("=" * 76)
    hdr.append("")
    hdr.append("                  KPP Standalone Atmospheric Chemical State")
    hdr.append("Meteorological and general grid cell metadata    ")
    hdr.append("")
    def put(label, val):
        hdr.append(label.ljust(49) + val)
    put("Location:", str(s.get("location", "")))
    put("Timestamp:", str(s.get("timestamp", "")))
    put("Longitude (degrees):", "%11.4f" % s.get("longitude", 0.0))
    put("Latitude (degrees):", "%11.4f" % s.get("latitude", 0.0))
    put("GEOS-Chem Vertical Level:", "%6d" % s.get("level", 1))
    put("Pressure (hPa):", "%11.4f" % s.get("pressure_hPa", 0.0))
    put("Temperature (K):", "%11.2f" % s.get("temperature_K", 0.0))
    put("Dry air density (molec/cm3):", "%11.4E" % s.get("numden", 0.0))
    put("Water vapor mixing ratio (vol H2O/vol dry air):", "%11.4E" % s.get("h2o_vmr", 0.0))
    put("Cloud fraction:", "%11.4E" % s.get("cloud_fraction", 0.0))
    put("Cosine of solar zenith angle:", "%11.4E" % s.get("cosSZA", 0.0))
    hdr.append("")
    hdr.append("KPP Integrator-specific parameters               ")
    hdr.append("")
    put("Init KPP Timestep (seconds):", "%11.4f" % s["Hstart"])
    put("Exit KPP Timestep (seconds):", "%11.4f" % s["Hexit"])
    put("Chemistry operator timestep (seconds):", "%11.4f" % s["OperatorTimestep"])
    put("Number of internal timesteps:", "%6d" % s["fileTotSteps"])
    hdr.append("ICNTRL integrator options used:")
    for r in range(2):
        hdr.append("".join("%6d" % v for v in s["ICNTRL"][10 * r:10 * r + 10]))
    hdr.append("RCNTRL integrator options used:")
    for r in range(4):
        hdr.append("".join("%13.6f" % v for v in s["RCNTRL"][5 * r:5 * r + 5]))
    hdr.append("")
    hdr.append("CSV data of full chemical state, including species concentrations,")
    hdr.append("rate constants (R) and instantaneous reaction rates (A).")
    hdr.append("All concentration units are in molec/cm3 and rates in molec/cm3/s.")
    hdr.append("")
    hdr.append("=" * 76)
    hdr.append("Name,   Value,   Absolute Tolerance")
    out = [str(len(hdr))] + hdr
    cstr = s.get("C_str")
    for i, nm in enumerate(s["names"]):
        v = cstr[i] if cstr else _es25(s["C"][i]).strip()
        out.append("%s,  %s,  %.2E" % (nm, v, s["ATOL"][i]))
    rstr, astr = s.get("R_str"), s.get("A_str")
    for i, v in enumerate(s["R"]):
        out.append("R%d,  %s" % (i + 1, rstr[i] if rstr else _es25(v).strip()))
    for i, v in enumerate(s.get("A", [])):
        out.append("A%d,  %s" % (i + 1, astr[i] if astr else _es25(v).strip()))
    return "\n".join(out) + "\n"


def format_output(names, cinit, cfinal):
    """The CSV tail of the box model's output file (kpp_standalone.F90:222-229):
    'Name,<ES25.16E3>,<ES25.16E3>' per species."""
    lines = ["Species Name,Initial Concentration (molec/cm3),Final Concentration (molec/cm3)"]
    for nm, a, b in zip(names, cinit, cfinal):
        lines.append("%s,%s,%s" % (nm, _es25(a), _es25(b)))
    return "\n".join(lines) + "\n"
