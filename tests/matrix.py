"""The parity matrix: which (operator, parameters) combinations are exercised, on the CPU
(oracle vs compiled reference) and on the GPU (product vs oracle / reference)."""
import itertools

from cases import ALL, NONE, SOME

# parameter variants per operator (each dict overrides the defaults of cases.SPECS)
VARIANTS = {
    "pleveltemp": [dict(compute=c, unit=u) for c, u in [(1, "celsius"), (1, "kelvin"), (2, ""), (3, "kelvin"), (4, ""), (5, "")]]
    + [dict(compute=3, p=-1.0), dict(compute=6), dict(compute=0, unit=""), dict(compute=4, p=1013.25)],
    "plevelhum": [dict(compute=c, unit=u) for c in range(1, 13) for u in ("celsius", "kelvin")] + [dict(compute=0), dict(compute=13), dict(compute=3, p=0.0)],
    "hleveltemp": [dict(compute=c, unit=u) for c, u in [(1, "celsius"), (2, "kelvin"), (3, "x"), (4, ""), (5, ""), (7, "")]]
    + [dict(compute=3, alevel=0.0, blevel=0.0), dict(compute=3, alevel=-1.0), dict(compute=3, blevel=1.5), dict(compute=3, alevel=0.0, blevel=1.0)],
    "hlevelthe": [dict(compute=1), dict(compute=2), dict(compute=3), dict(compute=1, blevel=-0.1)],
    "hlevelhum": [dict(compute=c, unit=u) for c in range(1, 13) for u in ("celsius", "kelvin")] + [dict(compute=0), dict(compute=13), dict(compute=1, alevel=-5.0)],
    "hlevelducting": [dict(compute=c) for c in (1, 2, 3, 4, 5)] + [dict(compute=1, blevel=2.0)],
    "hlevelpressure": [dict(), dict(alevel=0.0, blevel=0.0), dict(alevel=0.0, blevel=1.0)],
    "aleveltemp": [dict(compute=c, unit=u) for c, u in [(1, "celsius"), (1, "kelvin"), (2, "celsius"), (3, "kelvin"), (4, ""), (5, "")]] + [dict(compute=0), dict(compute=6)],
    "alevelthe": [dict(compute=1), dict(compute=2), dict(compute=3)],
    "alevelhum": [dict(compute=c, unit=u) for c in range(1, 13) for u in ("celsius", "kelvin")] + [dict(compute=0), dict(compute=13)],
    "alevelducting": [dict(compute=c) for c in (1, 2, 3, 4, 5)],
    "ilevelgwind": [dict()],
    "relvort": [dict()],
    "absvort": [dict()],
    "divergence": [dict()],
    "advection": [dict(hours=1.0), dict(hours=1.0 / 3600.0), dict(hours=24.0)],
    "gradient": [dict(compute=c) for c in (1, 2, 3, 4, 5)],
    "shapiro2_filter": [dict()],
    "windCooling": [dict(compute=1), dict(compute=2), dict(compute=3)],
    "thermalFrontParameter": [dict()],
    "momentumXcoordinate": [dict(), dict(fcoriolisMin=-8e-5)],
    "momentumYcoordinate": [dict(), dict(fcoriolisMin=0.0)],
    "jacobian": [dict()],
    "vesselIcingOverland": [dict()],
    "vesselIcingMertins": [dict()],
    "vesselIcingModStall": [dict(), dict(zmin=2.0, zmax=4.0), dict(zmin=4.0, zmax=2.0), dict(zmin=1.0, zmax=2.5), dict(vs=-1.0), dict(alpha=0.3)],
    "vesselIcingMincog": [dict(), dict(alt=2), dict(zmin=3.0, zmax=4.0), dict(zmin=1.0, zmax=2.5), dict(alpha=-0.1), dict(alpha=0.3, vs=8.0)],
    "fieldOPERfield": [dict(compute=c) for c in (1, 2, 3, 4, 5)],
    "meanValue": [dict()],
    "stddevValue": [dict()],
    "extremeValue": [dict(compute=c) for c in (1, 2, 3, 4, 5)],
    "probability": [dict(compute=c) for c in (1, 2, 3, 4, 5, 6)] + [dict(compute=3, limits=(270.0,)), dict(compute=1, limits=()), dict(compute=7)],
    "kIndex": [dict(compute=1), dict(compute=2), dict(compute=3), dict(compute=1, p500=700.0), dict(compute=1, p500=0.0), dict(compute=2, p850=925.0)],
    "ductingIndex": [dict(compute=1), dict(compute=2), dict(compute=0), dict(compute=1, p850=-1.0)],
    "showalterIndex": [dict(compute=1), dict(compute=2), dict(compute=3), dict(compute=1, p500=900.0), dict(compute=2, p500=400.0, p850=925.0)],
    "boydenIndex": [dict(compute=1), dict(compute=2), dict(compute=3), dict(compute=1, p700=1000.0)],
    "sweatIndex": [dict()],
    "seaSoundSpeed": [dict(compute=1), dict(compute=2, kinds={0: "tk"}), dict(compute=3), dict(compute=1, z=-1200.0)],
    "cvtemp": [dict(compute=1), dict(compute=2, kinds={0: "tc30"}), dict(compute=3), dict(compute=3, kinds={0: "tc30"}), dict(compute=4),
               dict(compute=4, kinds={0: "tc30"}), dict(compute=5)],
    "cvhum": [dict(compute=1, unit="kelvin"), dict(compute=1, unit="celsius"), dict(compute=2, unit=""), dict(compute=3, unit="", kinds={0: "tc30"}),
              dict(compute=4, unit="", kinds={1: "tk"}), dict(compute=4, unit="1", kinds={1: "tk"}), dict(compute=5, unit="1", kinds={0: "tc30", 1: "tc30"}),
              dict(compute=6, unit="")],
    "abshum": [dict()],
    "underCooledRain": [dict(), dict(precipMin=0.0, snowRateMax=0.5, tcMax=20.0)],
    "plevelthe": [dict(compute=1), dict(compute=2), dict(compute=3), dict(compute=1, p=0.0)],
    "pleveldz2tmean": [dict(compute=1), dict(compute=2), dict(compute=3), dict(compute=4), dict(compute=1, p1=500.0, p2=1000.0), dict(compute=1, p1=500.0, p2=500.0)],
    "plevelducting": [dict(compute=1), dict(compute=2), dict(compute=3, kinds={1: "rh"}), dict(compute=4, kinds={1: "rh"}), dict(compute=5), dict(compute=1, p=-3.0)],
    "vectorabs": [dict()],
    "pressure2FlightLevel": [dict()],
    "values2classes": [dict(), dict(limits=(260.0, 280.0)), dict(limits=(260.0,)), dict(limits=(250.0, 255.0, 260.0, 265.0, 270.0, 275.0, 280.0, 285.0, 290.0, 295.0))],
    "minvalueFields": [dict()],
    "minvalueFieldConst": [dict(), dict(value=1.0e35)],
    "maxvalueFields": [dict()],
    "maxvalueFieldConst": [dict(), dict(value=1.0e35)],
    "absvalueField": [dict()],
    "log10Field": [dict(), dict(kinds={0: "any"})],
    "pow10Field": [dict()],
    "logField": [dict(), dict(kinds={0: "any"})],
    "expField": [dict()],
    "powerField": [dict(), dict(value=-0.5), dict(value=2.0), dict(value=1.0e35)],
    "replaceUndefined": [dict(), dict(value=1.0e35)],
    "replaceDefined": [dict(), dict(value=1.0e35)],
    "fieldOPERconstant": [dict(compute=c) for c in (1, 2, 3, 4, 5)] + [dict(compute=4, value=0.0), dict(compute=2, value=1.0e35)],
    "constantOPERfield": [dict(compute=c) for c in (1, 2, 3, 4, 5)] + [dict(compute=4, value=0.0), dict(compute=2, value=1.0e35)],
    "sumFields": [dict()],
    "snow_in_cm": [dict()],
    "plevelgwind_xcomp": [dict()],
    "plevelgwind_ycomp": [dict()],
    "plevelgvort": [dict()],
    "plevelqvector": [dict(compute=c) for c in (1, 2, 3, 4, 5)] + [dict(compute=1, p=0.0)],
    "neighbourProbFunctions": [dict(compute=5), dict(compute=6), dict(compute=5, limits=(270.0, 0.0)), dict(compute=6, limits=(262.0, 1.0)),
                               dict(compute=5, limits=(270.0, 3.0)), dict(compute=1, limits=(270.0, 1.0)), dict(compute=5, limits=(270.0,)),
                               dict(compute=5, limits=(270.0, 50.0)), dict(compute=6, limits=(270.0, 9.0))],
    "neighbourFunctions": [dict(compute=1), dict(compute=2), dict(compute=3), dict(compute=1, limits=(1.0,)), dict(compute=2, limits=(3.0, 1.0)),
                           dict(compute=4, limits=(90.0, 2.0, 1.0)), dict(compute=4, limits=(50.0, 1.0)), dict(compute=4, limits=(0.0, 1.0, 2.0)),
                           dict(compute=5, limits=(270.0, 2.0, 1.0)), dict(compute=6, limits=(270.0, 1.0, 3.0)), dict(compute=4, limits=(100.0, 1.0)),
                           dict(compute=1, limits=(1.0, 4.0)), dict(compute=1, limits=(0.0,)), dict(compute=4, limits=(50.0,)), dict(compute=7, limits=(1.0, 1.0)),
                           dict(compute=2, limits=(40.0, 1.0))],
}

STENCILS = {"plevelqvector", "plevelgwind_xcomp", "plevelgwind_ycomp", "plevelgvort", "ilevelgwind", "relvort", "absvort", "divergence", "advection", "gradient", "shapiro2_filter", "thermalFrontParameter", "jacobian"}
ENSEMBLE = {"meanValue", "stddevValue", "extremeValue", "probability"}
SLOW = {"vesselIcingModStall", "vesselIcingMincog"}

# (mask, flag_in) combinations: the flag may lie about the data in both directions, like real callers
MASKS = [("none", ALL), ("none", SOME), ("bernoulli", SOME), ("nan", SOME), ("nan", ALL), ("edge", SOME), ("corner", SOME), ("all", SOME), ("all", NONE),
         ("blobs", SOME)]

GRIDS_SMALL = [(37, 23), (3, 3), (4, 5), (64, 8)]


def small_matrix():
    """(name, params, nx, ny, mask, flag) -- a few thousand tiny cases, seconds on one core"""
    out = []
    for name, variants in VARIANTS.items():
        grids = GRIDS_SMALL if name not in SLOW else [(9, 7), (3, 3)]
        for v in variants:
            for gi, (nx, ny) in enumerate(grids):
                masks = MASKS if gi == 0 else [("none", ALL), ("bernoulli", SOME)]
                for mask, flag in masks:
                    out.append((name, v, nx, ny, mask, flag))
    return out


def case_id(c):
    name, v, nx, ny, mask, flag = c
    return "%s-%s-%dx%d-%s-f%d" % (name, ",".join("%s=%s" % kv for kv in sorted(v.items())), nx, ny, mask, flag)
