"""Seeded synthetic inputs and the operator case table shared by the CPU and GPU parity suites.

A *case* is a list of arguments in the exact order of include/fcb200_api.inc, with
  - input fields as float32 numpy arrays (shape (ny, nx), nx fastest -- the reference's i = y*nx + x),
  - output fields pre-filled with a sentinel so that "left untouched" is observable,
  - the in/out ValuesDefined flag as a 1-element int32 array.
`run(api, case)` copies the arguments, calls `<prefix><op>` and returns (ret, outputs, flag).

Synthetic fields follow SURVEY.md 8(d): smooth sums of sinusoids plus a little white noise, in
physically plausible ranges (so that the saturation table stays in range unless a test wants
otherwise), with optional undefined-value patterns.
"""
from __future__ import annotations

import copy

import numpy as np

ALL, NONE, SOME = 0, 1, 2
SENTINEL = np.float32(-777.25)
UNDEF = np.float32(1.0e35)


# ----------------------------------------------------------------------------------------- fields
def smooth(rng, nx, ny, lo, hi, noise=0.01):
    """smooth field in [lo, hi]: a few sinusoids + white noise, float32 (ny, nx)"""
    y, x = np.mgrid[0:ny, 0:nx]
    f = np.zeros((ny, nx))
    for _ in range(3):
        kx, ky = rng.uniform(0.5, 4.0, 2)
        ph = rng.uniform(0, 2 * np.pi, 2)
        f += rng.uniform(0.3, 1.0) * np.sin(2 * np.pi * kx * x / max(nx, 2) + ph[0]) * np.cos(2 * np.pi * ky * y / max(ny, 2) + ph[1])
    f += noise * rng.standard_normal((ny, nx)) * 3
    f = (f - f.min()) / max(f.max() - f.min(), 1e-12)
    return (lo + (hi - lo) * f).astype(np.float32)


def uniform(rng, nx, ny, lo, hi):
    return rng.uniform(lo, hi, (ny, nx)).astype(np.float32)


FIELD_RANGES = {
    "tk": (215.0, 310.0),      # temperature, K
    "tc": (-25.0, 5.0),        # air temperature, deg C (icing)
    "th": (260.0, 330.0),      # potential temperature, K
    "q": (1e-6, 2e-2),         # specific humidity, kg/kg
    "rh": (1.0, 100.0),        # relative humidity, %
    "rh01": (0.4, 1.0),        # relative humidity, fraction (icing)
    "p": (150.0, 1040.0),      # pressure field, hPa
    "ps": (950.0, 1050.0),     # surface pressure, hPa
    "pmsl": (960.0, 1030.0),
    "wind": (-30.0, 30.0),
    "z": (4800.0, 5900.0),     # geopotential height / Montgomery-like
    "sst": (-1.0, 8.0),
    "sal": (30.0, 35.0),
    "aice": (0.0, 0.6),
    "wave": (0.0, 8.0),
    "pw": (3.0, 12.0),
    "depth": (20.0, 3000.0),
    "any": (-50.0, 50.0),
    "precip": (0.0, 5.0),      # mm
    "snow": (0.0, 0.6),        # mm water
    "tc30": (-40.0, 35.0),     # temperature, deg C (cvtemp / cvhum Celsius modes)
    "pfl": (5.0, 1050.0),      # pressure for the flight-level table, hPa (beyond both ends of the table)
    "pos": (0.05, 900.0),      # strictly positive (log, pow)
    "small": (-8.0, 8.0),      # exponent range (exp, 10^x)
    "snoww": (-0.5, 20.0),     # snow water, kg/m2 (some non-positive)
    "t2m": (255.0, 278.0),     # 2 m temperature / dew point around freezing, K
}


def field(rng, kind, nx, ny):
    if kind == "xm" or kind == "ym":
        return (smooth(rng, nx, ny, 0.97, 1.03, noise=0.0) / np.float32(2 * 2500.0)).astype(np.float32)
    if kind == "fc":
        return smooth(rng, nx, ny, 1.1e-4, 1.4e-4, noise=0.0)
    if kind == "fc0":  # Coriolis field crossing zero (momentum coordinates clamp)
        return smooth(rng, nx, ny, -3e-5, 1.4e-4, noise=0.0)
    lo, hi = FIELD_RANGES[kind]
    if kind in ("aice", "wave", "pw", "depth", "sal", "precip", "snow"):
        return uniform(rng, nx, ny, lo, hi)
    return smooth(rng, nx, ny, lo, hi)


def apply_mask(rng, a, mask, undef):
    """in-place undefined pattern"""
    ny, nx = a.shape
    if mask == "none":
        return a
    if mask == "bernoulli":
        a[rng.random(a.shape) < 0.3] = undef
    elif mask == "sparse":
        a[rng.random(a.shape) < 0.02] = undef
    elif mask == "nan":
        a[rng.random(a.shape) < 0.05] = np.nan
        a[rng.random(a.shape) < 0.05] = undef
    elif mask == "blobs":
        y, x = np.mgrid[0:ny, 0:nx]
        for _ in range(4):
            cx, cy = rng.uniform(0, nx), rng.uniform(0, ny)
            r = rng.uniform(0.05, 0.25) * max(nx, ny)
            a[(x - cx) ** 2 + (y - cy) ** 2 < r * r] = undef
    elif mask == "edge":  # only row 0, column 0 and the last column: the wrap-around flag quirk
        a[0, :: max(1, nx // 7)] = undef
        a[:: max(1, ny // 5), 0] = undef
        a[1:: max(1, ny // 3), nx - 1] = undef
    elif mask == "corner":
        a[0, 0] = undef
    elif mask == "all":
        a[:] = undef
    else:
        raise ValueError(mask)
    return a


# ----------------------------------------------------------------------------------------- case table
# Each spec: list of argument descriptors, in API order.
#   "nx", "ny", "undef", "flag", "out"
#   ("in", kind)            input field of a FIELD_RANGES kind (masked by the case's mask)
#   ("in!", kind)           input field that is never masked (map ratios, Coriolis)
#   ("f", name, default)    float scalar, overridable by params[name]
#   ("i", name, default)    int scalar
#   ("s", name, default)    string
#   ("members", kind)       ensemble member list + count (two C arguments)
#   "flags_in"              per-member int32 flags
#   ("limits", default)     float list + count (two C arguments)
#   "inout"                 shapiro2_filter's `field`
SPECS = {
    "pleveltemp": ["nx", "ny", ("in", "tk"), ("f", "p", 500.0), ("s", "unit", "kelvin"), ("i", "compute", 3), "out", "flag", "undef"],
    "plevelhum": ["nx", "ny", ("in", "tk"), ("in", "q"), ("f", "p", 850.0), ("s", "unit", "celsius"), ("i", "compute", 1), "out", "flag", "undef"],
    "hleveltemp": ["nx", "ny", ("in", "tk"), ("in", "ps"), ("f", "alevel", 50.0), ("f", "blevel", 0.7), ("s", "unit", "kelvin"), ("i", "compute", 3),
                   "out", "flag", "undef"],
    "hlevelthe": ["nx", "ny", ("in", "tk"), ("in", "q"), ("in", "ps"), ("f", "alevel", 50.0), ("f", "blevel", 0.7), ("i", "compute", 1), "out", "flag",
                  "undef"],
    "hlevelhum": ["nx", "ny", ("in", "tk"), ("in", "q"), ("in", "ps"), ("f", "alevel", 50.0), ("f", "blevel", 0.7), ("s", "unit", "celsius"),
                  ("i", "compute", 1), "out", "flag", "undef"],
    "hlevelducting": ["nx", "ny", ("in", "tk"), ("in", "q"), ("in", "ps"), ("f", "alevel", 50.0), ("f", "blevel", 0.7), ("i", "compute", 1), "out",
                      "flag", "undef"],
    "hlevelpressure": ["nx", "ny", ("in", "ps"), ("f", "alevel", 50.0), ("f", "blevel", 0.7), "out", "flag", "undef"],
    "aleveltemp": ["nx", "ny", ("in", "tk"), ("in", "p"), ("s", "unit", "kelvin"), ("i", "compute", 3), "out", "flag", "undef"],
    "alevelthe": ["nx", "ny", ("in", "tk"), ("in", "q"), ("in", "p"), ("i", "compute", 1), "out", "flag", "undef"],
    "alevelhum": ["nx", "ny", ("in", "tk"), ("in", "q"), ("in", "p"), ("s", "unit", "celsius"), ("i", "compute", 1), "out", "flag", "undef"],
    "alevelducting": ["nx", "ny", ("in", "tk"), ("in", "q"), ("in", "p"), ("i", "compute", 1), "out", "flag", "undef"],
    "ilevelgwind": ["nx", "ny", ("in", "z"), ("in!", "xm"), ("in!", "ym"), ("in!", "fc"), "out", "out", "flag", "undef"],
    "relvort": ["nx", "ny", ("in", "wind"), ("in", "wind"), ("in!", "xm"), ("in!", "ym"), "out", "flag", "undef"],
    "absvort": ["nx", "ny", ("in", "wind"), ("in", "wind"), ("in!", "xm"), ("in!", "ym"), ("in!", "fc"), "out", "flag", "undef"],
    "divergence": ["nx", "ny", ("in", "wind"), ("in", "wind"), ("in!", "xm"), ("in!", "ym"), "out", "flag", "undef"],
    "advection": ["nx", "ny", ("in", "tk"), ("in", "wind"), ("in", "wind"), ("in!", "xm"), ("in!", "ym"), ("f", "hours", 1.0), "out", "flag", "undef"],
    "gradient": ["nx", "ny", ("in", "tk"), ("in!", "xm"), ("in!", "ym"), ("i", "compute", 1), "out", "flag", "undef"],
    "shapiro2_filter": ["nx", "ny", ("inout", "tk"), "out", "flag", "undef"],
    "windCooling": ["nx", "ny", ("in", "tk"), ("in", "wind"), ("in", "wind"), ("i", "compute", 1), "out", "flag", "undef"],
    "thermalFrontParameter": ["nx", "ny", ("in", "tk"), ("in!", "xm"), ("in!", "ym"), "out", "flag", "undef"],
    "momentumXcoordinate": ["nx", "ny", ("in", "wind"), ("in!", "xm"), ("in!", "fc0"), ("f", "fcoriolisMin", 5e-5), "out", "flag", "undef"],
    "momentumYcoordinate": ["nx", "ny", ("in", "wind"), ("in!", "ym"), ("in!", "fc0"), ("f", "fcoriolisMin", 5e-5), "out", "flag", "undef"],
    "jacobian": ["nx", "ny", ("in", "tk"), ("in", "z"), ("in!", "xm"), ("in!", "ym"), "out", "flag", "undef"],
    "vesselIcingOverland": ["nx", "ny", ("in", "tc"), ("in", "sst"), ("in", "wind"), ("in", "wind"), ("in", "sal"), ("in", "aice"), "out", "flag",
                            "undef"],
    "vesselIcingMertins": ["nx", "ny", ("in", "tc"), ("in", "sst"), ("in", "wind"), ("in", "wind"), ("in", "sal"), ("in", "aice"), "out", "flag",
                           "undef"],
    "vesselIcingModStall": ["nx", "ny", ("in", "sal"), ("in", "wave"), ("in", "wind"), ("in", "wind"), ("in", "tc"), ("in", "rh01"), ("in", "sst"),
                            ("in", "pmsl"), ("in", "pw"), ("in", "aice"), ("in", "depth"), ("f", "vs", 5.0), ("f", "alpha", 2.6), ("f", "zmin", 4.0),
                            ("f", "zmax", 4.0), "out", "flag", "undef"],
    "vesselIcingMincog": ["nx", "ny", ("in", "sal"), ("in", "wave"), ("in", "wind"), ("in", "wind"), ("in", "tc"), ("in", "rh01"), ("in", "sst"),
                          ("in", "pmsl"), ("in", "pw"), ("in", "aice"), ("in", "depth"), ("f", "vs", 5.0), ("f", "alpha", 2.6), ("f", "zmin", 4.0),
                          ("f", "zmax", 4.0), ("i", "alt", 1), "out", "flag", "undef"],
    "fieldOPERfield": [("i", "compute", 1), "nx", "ny", ("in", "any"), ("in", "any"), "out", "flag", "undef"],
    "meanValue": ["nx", "ny", ("members", "tk"), "flags_in", "out", "flag", "undef"],
    "stddevValue": ["nx", "ny", ("members", "tk"), "flags_in", "out", "flag", "undef"],
    "extremeValue": [("i", "compute", 1), "nx", "ny", ("members", "tk"), "out", "flag", "undef"],
    "probability": [("i", "compute", 1), "nx", "ny", ("members", "tk"), "flags_in", ("limits", (262.0, 280.0)), "out", "flag", "undef"],
    # the rest of the reference's Python subset
    "kIndex": ["nx", "ny", ("in", "tk"), ("in", "tk"), ("in", "rh"), ("in", "tk"), ("in", "rh"), ("f", "p500", 500.0), ("f", "p700", 700.0),
               ("f", "p850", 850.0), ("i", "compute", 1), "out", "flag", "undef"],
    "ductingIndex": ["nx", "ny", ("in", "tk"), ("in", "rh"), ("f", "p850", 850.0), ("i", "compute", 1), "out", "flag", "undef"],
    "showalterIndex": ["nx", "ny", ("in", "tk"), ("in", "tk"), ("in", "rh"), ("f", "p500", 500.0), ("f", "p850", 850.0), ("i", "compute", 1), "out",
                       "flag", "undef"],
    "boydenIndex": ["nx", "ny", ("in", "tk"), ("in", "z"), ("in", "z"), ("f", "p700", 700.0), ("f", "p1000", 1000.0), ("i", "compute", 1), "out",
                    "flag", "undef"],
    "sweatIndex": ["nx", "ny", ("in", "tk"), ("in", "tk"), ("in", "tk"), ("in", "tk"), ("in", "wind"), ("in", "wind"), ("in", "wind"), ("in", "wind"),
                   "out", "flag", "undef"],
    "seaSoundSpeed": ["nx", "ny", ("in", "sst"), ("in", "sal"), ("f", "z", 50.0), ("i", "compute", 1), "out", "flag", "undef"],
    "cvtemp": ["nx", "ny", ("in", "tk"), ("i", "compute", 1), "out", "flag", "undef"],
    "cvhum": ["nx", "ny", ("in", "tk"), ("in", "rh"), ("s", "unit", "kelvin"), ("i", "compute", 1), "out", "flag", "undef"],
    "abshum": ["nx", "ny", ("in", "tk"), ("in", "rh"), "out", "flag", "undef"],
    "underCooledRain": ["nx", "ny", ("in", "precip"), ("in", "snow"), ("in", "tk"), ("f", "precipMin", 0.5), ("f", "snowRateMax", 0.1),
                        ("f", "tcMax", 0.0), "out", "flag", "undef"],
    # the rest of SURVEY.md 8f rank 1
    "plevelthe": ["nx", "ny", ("in", "tk"), ("in", "rh"), ("f", "p", 850.0), ("i", "compute", 1), "out", "flag", "undef"],
    "pleveldz2tmean": ["nx", "ny", ("in", "z"), ("in", "z"), ("f", "p1", 1000.0), ("f", "p2", 500.0), ("i", "compute", 1), "out", "flag", "undef"],
    "plevelducting": ["nx", "ny", ("in", "tk"), ("in", "q"), ("f", "p", 850.0), ("i", "compute", 1), "out", "flag", "undef"],
    "vectorabs": ["nx", "ny", ("in", "wind"), ("in", "wind"), "out", "flag", "undef"],
    "pressure2FlightLevel": ["nx", "ny", ("in", "pfl"), "out", "flag", "undef"],
    "values2classes": ["nx", "ny", ("in", "tk"), "out", ("limits", (230.0, 250.0, 262.5, 270.0, 280.0, 300.0)), "flag", "undef"],
    "minvalueFields": ["nx", "ny", ("in", "any"), ("in", "any"), "out", "flag", "undef"],
    "minvalueFieldConst": ["nx", "ny", ("in", "any"), ("f", "value", 3.5), "out", "flag", "undef"],
    "maxvalueFields": ["nx", "ny", ("in", "any"), ("in", "any"), "out", "flag", "undef"],
    "maxvalueFieldConst": ["nx", "ny", ("in", "any"), ("f", "value", 3.5), "out", "flag", "undef"],
    "absvalueField": ["nx", "ny", ("in", "any"), "out", "flag", "undef"],
    "log10Field": ["nx", "ny", ("in", "pos"), "out", "flag", "undef"],
    "pow10Field": ["nx", "ny", ("in", "small"), "out", "flag", "undef"],
    "logField": ["nx", "ny", ("in", "pos"), "out", "flag", "undef"],
    "expField": ["nx", "ny", ("in", "small"), "out", "flag", "undef"],
    "powerField": ["nx", "ny", ("in", "pos"), ("f", "value", 1.7), "out", "flag", "undef"],
    "replaceUndefined": ["nx", "ny", ("in", "any"), ("f", "value", -1.0), "out", "flag", "undef"],
    "replaceDefined": ["nx", "ny", ("in", "any"), ("f", "value", -1.0), "out", "flag", "undef"],
    "fieldOPERconstant": [("i", "compute", 1), "nx", "ny", ("in", "any"), ("f", "value", 2.5), "out", "flag", "undef"],
    "constantOPERfield": [("i", "compute", 1), "nx", "ny", ("f", "value", 2.5), ("in", "any"), "out", "flag", "undef"],
    "sumFields": ["nx", "ny", ("members", "tk"), "out", "flag", "undef"],
    "snow_in_cm": ["nx", "ny", ("in", "snoww"), ("in", "t2m"), ("in", "t2m"), "out", "flag", "undef"],
    # geostrophic stencil siblings
    "plevelgwind_xcomp": ["nx", "ny", ("in", "z"), ("in!", "xm"), ("in!", "ym"), ("in!", "fc"), "out", "flag", "undef"],
    "plevelgwind_ycomp": ["nx", "ny", ("in", "z"), ("in!", "xm"), ("in!", "ym"), ("in!", "fc"), "out", "flag", "undef"],
    "plevelgvort": ["nx", "ny", ("in", "z"), ("in!", "xm"), ("in!", "ym"), ("in!", "fc"), "out", "flag", "undef"],
    # neighbourhood functions; the field is never masked (they require ALL_DEFINED and sort raw values)
    "neighbourProbFunctions": ["nx", "ny", ("in!", "tk"), ("limits", (270.0, 2.0)), ("i", "compute", 5), "out", "flag", "undef"],
    "neighbourFunctions": ["nx", "ny", ("in!", "tk"), ("limits", (2.0, 2.0)), ("i", "compute", 1), "out", "flag", "undef"],
    "plevelqvector": ["nx", "ny", ("in", "z"), ("in", "tk"), ("in!", "xm"), ("in!", "ym"), ("in!", "fc"), ("f", "p", 700.0), ("i", "compute", 1), "out",
                      "flag", "undef"],
}

# operators whose device result may differ from the CPU by transcendental ulps (powf/expf/exp/pow/tanh);
# everything else must be bit-identical.  Values are relative tolerances on defined points.
TRANSCENDENTAL = {
    "hleveltemp": 1e-5, "hlevelthe": 1e-5, "hlevelhum": 1e-5, "hlevelducting": 1e-5,
    "aleveltemp": 1e-5, "alevelthe": 1e-5, "alevelhum": 1e-5, "alevelducting": 1e-5,
    "windCooling": 1e-5, "vesselIcingModStall": 1e-5, "vesselIcingMincog": 1e-4,
    "abshum": 1e-6,  # double exp(): CUDA's is within 1 ulp of a double, the float result differs about once in 2^29
    # device libm: logf 1 ulp, log10f / expf 2 ulp, powf 4 ulp, double pow / exp 1 ulp of a double
    "log10Field": 1e-6, "logField": 1e-6, "expField": 1e-6, "powerField": 2e-6, "pow10Field": 1e-6, "snow_in_cm": 1e-6,
}


class Case:
    def __init__(self, name, args, out_idx, flag_idx, undef, params):
        self.name, self.args, self.out_idx, self.flag_idx, self.undef, self.params = name, args, out_idx, flag_idx, undef, params

    def __repr__(self):
        return "Case(%s, %s)" % (self.name, self.params)


KAPPA = 287.0 / 1004.0


def theta_input(name, params):
    """True if the operator/mode reads its first field as POTENTIAL temperature: the synthetic
    temperature is then divided by the Exner factor so that the physical temperature stays in the
    range of the saturation table (-100..100 C) -- otherwise the lookup is out of range (-> undef)
    or, just below -100 C, linearly extrapolated through zero where any rounding difference of
    powf is amplified without bound."""
    c = params.get("compute", None)
    if name in ("pleveltemp", "hleveltemp", "aleveltemp"):
        return c in (1, 2, 5)
    if name in ("plevelhum", "hlevelhum", "alevelhum", "hlevelducting", "alevelducting", "hlevelthe", "alevelthe"):
        return c is not None and c % 2 == 0
    return False


def humidity_is_rh(name, params):
    """True if the operator/mode reads its second field as relative humidity (%) instead of q (kg/kg)."""
    c = params.get("compute", None)
    if name == "plevelhum":
        return c in (3, 4, 5, 6, 9, 10)
    if name in ("hlevelhum", "alevelhum"):
        return c in (3, 4, 7, 8, 11, 12)
    if name in ("hlevelducting", "alevelducting"):
        return c in (3, 4)
    return False


def abs_floor(name, params):
    """Magnitude below which a value is compared absolutely (rtol * floor) instead of relatively:
    temperatures in Celsius and temperature differences are judged on the Kelvin scale."""
    c = params.get("compute", None)
    unit = params.get("unit", "")
    if name in ("pleveltemp", "hleveltemp", "aleveltemp") and c in (1, 2):
        return 273.15
    if name in ("plevelhum", "hlevelhum", "alevelhum") and c is not None and c >= 5:
        return 273.15
    if name == "windCooling":
        return 13.12  # the polynomial's constant term: the result is a difference of O(10) terms clamped at 0
    if name in ("vesselIcingModStall", "vesselIcingMincog"):
        return 1.0  # cm/h; icing rates of interest are O(0.1 .. 10)
    return 0.0


def build(name, nx, ny, seed=0, undef=UNDEF, flag_in=SOME, mask="none", nmembers=5, member_flags=None, alias=False, **params):
    """Build one case.  `params` override scalar defaults (compute=..., unit=..., p=...)."""
    rng = np.random.default_rng(seed)
    undef = np.float32(undef)
    spec = SPECS[name]
    full = {d[1]: params.get(d[1], d[2]) for d in spec if isinstance(d, tuple) and d[0] in ("f", "i", "s")}
    args, out_idx, flag_idx = [], [], None
    in_fields = []  # (arg index, kind, maskable)
    for d in spec:
        if d == "nx":
            args.append(nx)
        elif d == "ny":
            args.append(ny)
        elif d == "undef":
            args.append(float(undef))
        elif d == "flag":
            flag_idx = len(args)
            args.append(np.array([flag_in], dtype=np.int32))
        elif d == "out":
            out_idx.append(len(args))
            args.append(np.full((ny, nx), SENTINEL, dtype=np.float32))
        elif d == "flags_in":
            mf = member_flags if member_flags is not None else [flag_in] * nmembers
            args.append(np.array(mf, dtype=np.int32))
        elif d[0] in ("in", "in!", "inout"):
            kind = params.get("kinds", {}).get(len(in_fields), d[1])  # kinds={position: kind} overrides the spec's field kind
            in_fields.append((len(args), kind, d[0] != "in!"))
            args.append(field(rng, kind, nx, ny))
        elif d[0] in ("f", "i", "s"):
            args.append(full[d[1]])
        elif d[0] == "members":
            base = field(rng, d[1], nx, ny)
            members = []
            for _ in range(nmembers):
                m = (base + rng.standard_normal((ny, nx)).astype(np.float32) * np.float32(3.0)).astype(np.float32)
                apply_mask(rng, m, mask, undef)
                members.append(m)
            args.append(members)
            args.append(nmembers)
        elif d[0] == "limits":
            lim = list(params.get("limits", d[1]))
            args.append(np.array(lim if lim else [0.0], dtype=np.float32))
            args.append(len(lim))
        else:
            raise ValueError(d)
    if humidity_is_rh(name, full):
        args[in_fields[1][0]] = field(rng, "rh", nx, ny)
    if theta_input(name, full) and in_fields:
        # first field is theta: derive it from the synthetic temperature and the level's pressure
        kinds = [k for _, k, _ in in_fields]
        t_idx = in_fields[0][0]
        if "p" in kinds:
            p = args[in_fields[kinds.index("p")][0]].astype(np.float64)
        elif "ps" in kinds:
            p = full.get("alevel", 0.0) + full.get("blevel", 1.0) * args[in_fields[kinds.index("ps")][0]].astype(np.float64)
        else:
            p = np.float64(full.get("p", 1000.0))
        with np.errstate(all="ignore"):
            exner = np.where(p > 0, np.abs(p) / 1000.0, 1.0) ** KAPPA
        args[t_idx] = (args[t_idx].astype(np.float64) / exner).astype(np.float32)
    for idx, _, maskable in in_fields:
        if maskable:
            apply_mask(rng, args[idx], mask, undef)
    case = Case(name, args, out_idx, flag_idx, undef, dict(params, nx=nx, ny=ny, seed=seed, flag_in=flag_in, mask=mask))
    case.floor = abs_floor(name, full)
    if alias:  # output aliases the first input field (allowed for elementwise operators and shapiro2_filter)
        case.alias = (in_fields[0][0], out_idx[0])
    return case


def run(api, case, to_device=None):
    """Call `<prefix><op>` on a private copy of the case.  `to_device(arr) -> tensor` moves field
    arguments to the GPU first (device-pointer path); outputs come back as numpy arrays."""
    args = copy.deepcopy(case.args)
    alias = getattr(case, "alias", None)
    if alias:
        args[alias[1]] = args[alias[0]]
    moved = {}
    if to_device is not None:
        for k, a in enumerate(args):
            if isinstance(a, np.ndarray) and a.dtype == np.float32 and a.ndim == 2:
                if alias and k == alias[1]:
                    continue
                moved[k] = to_device(a)
            elif isinstance(a, list):
                moved[k] = [to_device(m) for m in a]
        if alias:
            moved[alias[1]] = moved[alias[0]]
    call_args = [moved.get(k, a) for k, a in enumerate(args)]
    ret = api.call(case.name, *call_args)
    outs = []
    for k in case.out_idx:
        o = call_args[k]
        outs.append(o.cpu().numpy() if hasattr(o, "cpu") else o)
    return ret, outs, int(args[case.flag_idx][0])


def undefined_mask(a, undef):
    return (a == undef) | np.isnan(a)


def compare(case, got, want, rtol=0.0):
    """Returns a list of human-readable mismatches (empty = parity).  `got`/`want` = (ret, outs, flag).
    The undefined mask and the flag must be bit-exact; values are compared bit-for-bit when rtol == 0,
    otherwise by relative error on the defined points."""
    problems = []
    if got[0] != want[0]:
        problems.append("return value %r != %r" % (got[0], want[0]))
    if got[2] != want[2]:
        problems.append("fDefined %r != %r" % (got[2], want[2]))
    for k, (g, w) in enumerate(zip(got[1], want[1])):
        mg, mw = undefined_mask(g, case.undef), undefined_mask(w, case.undef)
        if not np.array_equal(mg, mw):
            problems.append("output %d: undefined mask differs at %d points" % (k, int((mg != mw).sum())))
            continue
        if rtol == 0.0:
            same = (g.view(np.uint32) == w.view(np.uint32)) | (np.isnan(g) & np.isnan(w)) | (g == w)
            if not same.all():
                bad = np.argwhere(~same)
                y, x = bad[0]
                problems.append("output %d: %d points differ bitwise, first at (y=%d,x=%d): %r vs %r" % (k, len(bad), y, x, g[y, x], w[y, x]))
        else:
            d = ~mw
            with np.errstate(all="ignore"):
                err = np.abs(g[d].astype(np.float64) - w[d].astype(np.float64))
                scale = np.maximum(np.abs(w[d].astype(np.float64)), max(getattr(case, "floor", 0.0), 1e-30))
                rel = np.where(err == 0, 0.0, err / scale)
                # both infinite with the same sign counts as equal
                rel = np.where(np.isinf(g[d]) & (g[d] == w[d]), 0.0, rel)
                # anything else that is not a number (inf against -inf, inf against a finite value ...) is a mismatch, not a value to skip
                rel = np.where(np.isnan(rel), np.inf, rel)
            if rel.size and np.max(rel) > rtol:
                worst = int(np.argmax(rel))
                problems.append("output %d: max relative error %.3g > %.1g (got %r, want %r)" % (k, float(np.max(rel)), rtol, g[d][worst], w[d][worst]))
    return problems
