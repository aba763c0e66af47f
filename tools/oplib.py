"""Operator work-lists for bench.py and tools/opbench.py: synthetic inputs of BASELINE.json's five configurations
 and, for every operator, the argument list of its batched C-ABI entry point plus the argument list of the
reference's one-field call on the FIRST field of the same inputs (for the CPU timing beside it).

An operator's arguments are written once, as a list of items:
    "nx", "ny", "nf"                       grid and batch size (the single-field call has no "nf")
    ("F", kind)                            per-field input, [nf, ny, nx] on the device (masked when the row is masked)
    ("G", kind)                            grid-constant input, [ny, nx] (map ratios, Coriolis, surface pressure: never masked)
    ("S", value)                           scalar (float, int or str), same in both calls
    ("SV", numpy array)                    per-field scalar array (host); the single-field call takes element 0
    "OUT"                                  output, [nf, ny, nx]
    "FLAGS"                                in/out ValuesDefined per field
    ("MEMBERS", kind, M)                   ensemble: M member arrays [nt, ny, nx] + the count M (two C arguments)
    "MFLAGS"                               per (time, member) input flags
    ("LIMITS", values)                     float list + count (two C arguments)
Algorithmic bytes per point follow SURVEY.md 8(a), batched figures (grid-constant arrays amortised over the batch).
"""
from __future__ import annotations

import numpy as np

MEPS = (949, 1069)
ECMWF = (3600, 1801)
UNDEF = 1.0e35
ALL_DEFINED, NONE_DEFINED, SOME_DEFINED = 0, 1, 2

# value ranges of the synthetic fields
RANGES = {"t": (215.0, 305.0), "q": (1e-6, 2e-2), "p": (300.0, 1040.0), "w": (-30.0, 30.0), "any": (-50.0, 50.0), "tc": (-25.0, 5.0), "sst": (-1.0, 8.0),
          "sal": (30.0, 35.0), "aice": (0.0, 0.6), "wave": (0.0, 8.0), "rh01": (0.4, 1.0), "pmsl": (960.0, 1030.0), "pw": (3.0, 12.0),
          "depth": (20.0, 3000.0), "rh": (1.0, 100.0), "z": (4800.0, 5900.0), "ps": (950.0, 1040.0), "tens": (250.0, 300.0),
          "xm": (1.9e-4, 2.1e-4), "ym": (1.9e-4, 2.1e-4), "fc": (1.1e-4, 1.4e-4), "xme": (4.5e-5, 9.0e-5), "yme": (4.497e-5, 4.497e-5)}


class Inputs:
    """Seeded synthetic fields on one device: a smooth part (sinusoids) plus white noise, scaled into the kind's range."""

    def __init__(self, torch, device, seed=1234):
        self.torch, self.device = torch, device
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(seed)
        self._xy = {}

    def _smooth(self, grid):
        if grid not in self._xy:
            t = self.torch
            nx, ny = grid
            x = t.arange(nx, device=self.device, dtype=t.float32)[None, :]
            y = t.arange(ny, device=self.device, dtype=t.float32)[:, None]
            self._xy[grid] = 0.5 + 0.35 * t.sin(x * (6.28 * 3 / nx) + 0.3) * t.cos(y * (6.28 * 2 / ny) + 0.1)
        return self._xy[grid]

    def field(self, kind, grid, nf=None, mask=0.0):
        t = self.torch
        lo, hi = RANGES[kind]
        nx, ny = grid
        shape = (ny, nx) if nf is None else (nf, ny, nx)
        a = t.rand(shape, device=self.device, generator=self.gen, dtype=t.float32)
        if kind not in ("aice", "wave", "pw", "depth", "sal"):  # those are white noise over their range
            a.mul_(0.15).add_(self._smooth(grid))
        a.mul_(hi - lo).add_(lo)
        if mask > 0:
            a[t.rand(shape, device=self.device, generator=self.gen) < mask] = UNDEF
        return a


class Row:
    def __init__(self, name, cfg, call, grid, nf, bpp, items, masked=None, cpu_rows=None, note=None, common_mask=False):
        self.name, self.cfg, self.call, self.grid, self.nf, self.bpp, self.items = name, cfg, call, grid, nf, bpp, items
        self.masked = masked          # None = follow the run's mask; a number pins it (cfg4's 5 %)
        self.common_mask = common_mask  # the SAME points are undefined in every per-field input (a land/sea or below-ground mask)
        self.cpu_rows = cpu_rows      # rows of the grid in the CPU sample (None = the whole field)
        self.note = note


def _stencil(nin, extra=(), maps=("xm", "ym"), nout=1, first="tens"):
    kinds = [first] + ["w"] * (nin - 1)
    return ["nx", "ny", "nf"] + [("F", k) for k in kinds] + [("G", m) for m in maps] + [("S", e) for e in extra] + ["OUT"] * nout + ["FLAGS", "UNDEF"]


def _ew(kinds, scalars=(), lead=()):
    return [("S", s) for s in lead] + ["nx", "ny", "nf"] + [("F", k) for k in kinds] + [("S", s) for s in scalars] + ["OUT", "FLAGS", "UNDEF"]


def _ens(M, lead=(), limits=None, mflags=True):
    items = [("S", s) for s in lead] + ["nx", "ny", "nf", ("MEMBERS", "tens", M)]
    if mflags:
        items.append("MFLAGS")
    if limits is not None:
        items.append(("LIMITS", limits))
    return items + ["OUT", "FLAGS", "UNDEF"]


def _hybrid_levels(nf):
    eta = (np.arange(nf) % 65 + 0.5) / 65.0
    return (200.0 * (1 - eta) * eta * 2.0 + 10.0 * (1 - eta)).astype(np.float32), (eta ** 1.5).astype(np.float32)


ICING6 = ["tc", "sst", "w", "w", "sal", "aice"]
ICING11 = ["sal", "wave", "w", "w", "tc", "rh01", "sst", "pmsl", "pw", "aice", "depth"]
EMAPS = ("xme", "yme")  # ECMWF lat-lon map ratios


def rows(levels_cfg3=137, nt_cfg4=8):
    """the per-operator battery: BASELINE.json configs 1, 3, 4, 5 and the single operators of config 2"""
    a65, b65 = _hybrid_levels(65)
    r = [
        # cfg1: the MEPS 500 hPa case -- as a rotating batch of 64 fields so that the number is an HBM number (one 4 MB field is
        # L2 resident and launch-latency bound: that is bench.py's cfg1_latency record)
        Row("pleveltemp_c3", "cfg1", "pleveltemp_batched", MEPS, 128, 8, ["nx", "ny", "nf", ("F", "t"), ("SV", np.full(128, 500.0, np.float32)), ("S", "kelvin"), ("S", 3), "OUT", "FLAGS", "UNDEF"]),
        Row("relvort", "cfg1", "relvort_batched", MEPS, 64, 12, _stencil(2, first="w")),
        Row("divergence", "cfg1", "divergence_batched", MEPS, 64, 12, _stencil(2, first="w")),
        # cfg2: the single operators behind the headline chain
        Row("aleveltemp_c3", "cfg2", "aleveltemp_batched", MEPS, 65, 12, _ew(["t", "p"], ("kelvin", 3))),
        Row("alevelhum_c1", "cfg2", "alevelhum_batched", MEPS, 65, 16, _ew(["t", "q", "p"], ("celsius", 1))),
        Row("alevelhum_c5", "cfg2", "alevelhum_batched", MEPS, 65, 16, _ew(["t", "q", "p"], ("celsius", 5))),
        Row("alevelthe_c1", "cfg2", "alevelthe_batched", MEPS, 65, 16, _ew(["t", "q", "p"], (1,))),
        Row("hlevelhum_c1", "cfg2", "hlevelhum_batched", MEPS, 65, 12, ["nx", "ny", "nf", ("F", "t"), ("F", "q"), ("G", "ps"), ("SV", a65), ("SV", b65), ("S", "celsius"), ("S", 1), "OUT", "FLAGS", "UNDEF"]),
        Row("hleveltemp_c3", "cfg2", "hleveltemp_batched", MEPS, 65, 8, ["nx", "ny", "nf", ("F", "t"), ("G", "ps"), ("SV", a65), ("SV", b65), ("S", "kelvin"), ("S", 3), "OUT", "FLAGS", "UNDEF"]),
        # cfg3: ECMWF 0.1 degree, all levels in one batch
        Row("advection", "cfg3", "advection_batched", ECMWF, levels_cfg3, 16, _stencil(3, extra=(1.0,), maps=EMAPS)),
        Row("thermalFrontParameter", "cfg3", "thermalFrontParameter_batched", ECMWF, levels_cfg3, 8, _stencil(1, maps=EMAPS)),
        Row("shapiro2_filter", "cfg3", "shapiro2_filter_batched", ECMWF, levels_cfg3, 8, ["nx", "ny", "nf", ("F", "tens"), "OUT", "FLAGS", "UNDEF"]),
        Row("gradient_c3", "cfg3", "gradient_batched", ECMWF, min(levels_cfg3, 32), 8, _stencil(1, extra=(3,), maps=EMAPS)),
        # cfg4: 30-member ensemble; bytes per OUTPUT point = 4 (M + 1); member flags ALL_DEFINED and SOME_DEFINED with 5 % undefined
    ]
    for tag, mk in (("", 0.0), ("_masked5", 0.05)):
        r += [
            Row("meanValue" + tag, "cfg4", "meanValue_batched", MEPS, nt_cfg4, 124, _ens(30), masked=mk, cpu_rows=267),
            Row("stddevValue" + tag, "cfg4", "stddevValue_batched", MEPS, nt_cfg4, 124, _ens(30), masked=mk, cpu_rows=267),
            Row("extremeValue_max" + tag, "cfg4", "extremeValue_batched", MEPS, nt_cfg4, 124, _ens(30, lead=(1,), mflags=False), masked=mk, cpu_rows=267),
            Row("extremeValue_min" + tag, "cfg4", "extremeValue_batched", MEPS, nt_cfg4, 124, _ens(30, lead=(2,), mflags=False), masked=mk, cpu_rows=267),
            Row("probability_above" + tag, "cfg4", "probability_batched", MEPS, nt_cfg4, 124, _ens(30, lead=(1,), limits=[275.0]), masked=mk, cpu_rows=267),
        ]
    # cfg5: 30 % of every maskable input undefined, flags SOME_DEFINED
    m = 0.3
    r += [
        Row("vesselIcingOverland_masked30", "cfg5", "vesselIcingOverland_batched", MEPS, 40, 28, _ew(ICING6), masked=m),
        Row("vesselIcingMertins_masked30", "cfg5", "vesselIcingMertins_batched", MEPS, 40, 28, _ew(ICING6), masked=m),
        Row("vesselIcingModStall_masked30", "cfg5", "vesselIcingModStall_batched", MEPS, 4, 48, _ew(ICING11, (5.0, 2.6, 4.0, 4.0)), masked=m, cpu_rows=48,
            note="compute-bound (double RK4 x 50, two fixed-point loops): judged against an instruction bound, not HBM"),
        Row("vesselIcingMincog_masked30", "cfg5", "vesselIcingMincog_batched", MEPS, 4, 48, _ew(ICING11, (5.0, 2.6, 4.0, 4.0, 1)), masked=m, cpu_rows=48,
            note="compute-bound (float RK4 x 50 + 17 bisections per height): judged against an instruction bound, not HBM"),
    ]
    for c, nm in ((1, "add"), (2, "sub"), (3, "mul"), (4, "div")):
        r.append(Row("fieldOPERfield_%s_masked30" % nm, "cfg5", "fieldOPERfield_batched", MEPS, 96, 12, _ew(["any", "any"], (), lead=(c,)), masked=m))
    r += [
        Row("relvort_masked30", "cfg5", "relvort_batched", MEPS, 64, 12, _stencil(2, first="w"), masked=m),
        Row("divergence_masked30", "cfg5", "divergence_batched", MEPS, 64, 12, _stencil(2, first="w"), masked=m),
        Row("advection_masked30", "cfg5", "advection_batched", MEPS, 48, 16, _stencil(3, extra=(1.0,)), masked=m),
        Row("gradient_c3_masked30", "cfg5", "gradient_batched", MEPS, 96, 8, _stencil(1, extra=(3,)), masked=m),
        Row("jacobian_masked30", "cfg5", "jacobian_batched", MEPS, 64, 12, _stencil(2), masked=m),
        Row("thermalFrontParameter_masked30", "cfg5", "thermalFrontParameter_batched", MEPS, 96, 8, _stencil(1), masked=m),
        Row("shapiro2_filter_masked30", "cfg5", "shapiro2_filter_batched", MEPS, 96, 8, ["nx", "ny", "nf", ("F", "tens"), "OUT", "FLAGS", "UNDEF"], masked=m),
        Row("alevelhum_c5_masked30", "cfg5", "alevelhum_batched", MEPS, 65, 16, _ew(["t", "q", "p"], ("celsius", 5)), masked=m,
            note="t, q and p undefined INDEPENDENTLY: 15 % of the points have a defined t, q over an undefined p, which the reference lets flow into "
                 "RH and Td (FC.cc:1429) -- they take the IEEE redo"),
        Row("alevelhum_c5_masked30_common", "cfg5", "alevelhum_batched", MEPS, 65, 16, _ew(["t", "q", "p"], ("celsius", 5)), masked=m, common_mask=True,
            note="the same 30 % of the points undefined in t, q and p (one mask for the level, as below-ground points are)"),
    ]
    return r


def extra_rows():
    """the rest of the operator surface (tools/opbench.py --all): SURVEY.md 8a rows not named in a BASELINE config and 8f's next rows"""
    a96, b96 = _hybrid_levels(96)
    geo = ["nx", "ny", "nf", ("F", "z"), ("G", "xm"), ("G", "ym"), ("G", "fc"), "OUT", "FLAGS", "UNDEF"]
    return [
        Row("absvort", "8a", "absvort_batched", MEPS, 64, 12, _stencil(2, maps=("xm", "ym", "fc"), first="w")),
        Row("jacobian", "8a", "jacobian_batched", MEPS, 64, 12, _stencil(2)),
        Row("ilevelgwind", "8a", "ilevelgwind_batched", MEPS, 64, 12, ["nx", "ny", "nf", ("F", "z"), ("G", "xm"), ("G", "ym"), ("G", "fc"), "OUT", "OUT", "FLAGS", "UNDEF"]),
        Row("advection_meps", "8a", "advection_batched", MEPS, 48, 16, _stencil(3, extra=(1.0,))),
        Row("thermalFrontParameter_meps", "8a", "thermalFrontParameter_batched", MEPS, 96, 8, _stencil(1)),
        Row("shapiro2_filter_meps", "8a", "shapiro2_filter_batched", MEPS, 96, 8, ["nx", "ny", "nf", ("F", "tens"), "OUT", "FLAGS", "UNDEF"]),
        Row("pleveltemp_c4", "8a", "pleveltemp_batched", MEPS, 128, 8, ["nx", "ny", "nf", ("F", "t"), ("SV", np.full(128, 500.0, np.float32)), ("S", "kelvin"), ("S", 4), "OUT", "FLAGS", "UNDEF"]),
        Row("plevelhum_c1", "8a", "plevelhum_batched", MEPS, 96, 12, ["nx", "ny", "nf", ("F", "t"), ("F", "q"), ("SV", np.full(96, 850.0, np.float32)), ("S", "celsius"), ("S", 1), "OUT", "FLAGS", "UNDEF"]),
        Row("plevelhum_c7", "8a", "plevelhum_batched", MEPS, 96, 12, ["nx", "ny", "nf", ("F", "t"), ("F", "q"), ("SV", np.full(96, 850.0, np.float32)), ("S", "celsius"), ("S", 7), "OUT", "FLAGS", "UNDEF"]),
        Row("hlevelhum_c5", "8a", "hlevelhum_batched", MEPS, 96, 12, ["nx", "ny", "nf", ("F", "t"), ("F", "q"), ("G", "ps"), ("SV", a96), ("SV", b96), ("S", "celsius"), ("S", 5), "OUT", "FLAGS", "UNDEF"]),
        Row("alevelhum_c7", "8a", "alevelhum_batched", MEPS, 65, 16, _ew(["t", "rh", "p"], ("celsius", 7))),
        Row("alevelducting_c1", "8a", "alevelducting_batched", MEPS, 65, 16, _ew(["t", "q", "p"], (1,))),
        Row("windCooling", "8a", "windCooling_batched", MEPS, 65, 16, _ew(["t", "w", "w"], (1,))),
        Row("fieldOPERfield_add", "8a", "fieldOPERfield_batched", MEPS, 96, 12, _ew(["any", "any"], (), lead=(1,))),
        Row("fieldOPERfield_div", "8a", "fieldOPERfield_batched", MEPS, 96, 12, _ew(["any", "any"], (), lead=(4,))),
        Row("vesselIcingOverland", "8a", "vesselIcingOverland_batched", MEPS, 40, 28, _ew(ICING6)),
        Row("vesselIcingMertins", "8a", "vesselIcingMertins_batched", MEPS, 40, 28, _ew(ICING6)),
        Row("vesselIcingModStall", "8a", "vesselIcingModStall_batched", MEPS, 4, 48, _ew(ICING11, (5.0, 2.6, 4.0, 4.0)), cpu_rows=48),
        Row("vesselIcingMincog", "8a", "vesselIcingMincog_batched", MEPS, 4, 48, _ew(ICING11, (5.0, 2.6, 4.0, 4.0, 1)), cpu_rows=48),
        Row("kIndex", "8f", "kIndex_batched", MEPS, 48, 24, _ew(["t", "t", "rh", "t", "rh"], (500.0, 700.0, 850.0, 1))),
        Row("ductingIndex", "8f", "ductingIndex_batched", MEPS, 96, 12, _ew(["t", "rh"], (850.0, 1))),
        Row("showalterIndex", "8f", "showalterIndex_batched", MEPS, 64, 20, _ew(["t", "t", "rh"], (500.0, 850.0, 1))),
        Row("boydenIndex", "8f", "boydenIndex_batched", MEPS, 64, 16, _ew(["t", "z", "z"], (700.0, 1000.0, 1))),
        Row("seaSoundSpeed", "8f", "seaSoundSpeed_batched", MEPS, 96, 12, _ew(["sst", "sal"], (50.0, 1))),
        Row("cvtemp_c1", "8f", "cvtemp_batched", MEPS, 128, 8, _ew(["t"], (1,))),
        Row("cvhum_c1", "8f", "cvhum_batched", MEPS, 96, 12, _ew(["t", "rh"], ("kelvin", 1))),
        Row("abshum", "8f", "abshum_batched", MEPS, 96, 12, _ew(["t", "rh"])),
        Row("plevelthe_c1", "8f", "plevelthe_batched", MEPS, 96, 12, _ew(["t", "rh"], (850.0, 1))),
        Row("vectorabs", "8f", "vectorabs_batched", MEPS, 96, 12, _ew(["w", "w"])),
        Row("pressure2FlightLevel", "8f", "pressure2FlightLevel_batched", MEPS, 128, 8, _ew(["p"])),
        Row("plevelgwind_ycomp", "8f", "plevelgwind_ycomp_batched", MEPS, 64, 8, geo),
        Row("plevelgvort", "8f", "plevelgvort_batched", MEPS, 64, 8, geo),
        Row("plevelqvector_c1", "8f", "plevelqvector_batched", MEPS, 48, 12,
            ["nx", "ny", "nf", ("F", "z"), ("F", "t"), ("G", "xm"), ("G", "ym"), ("G", "fc"), ("S", 700.0), ("S", 1), "OUT", "FLAGS", "UNDEF"]),
    ]


class Built:
    """device-resident arguments of one row (one or two input sets) and their single-field host twins"""

    def __init__(self, row, inputs, mask, sets=1):
        self.row = row
        torch = inputs.torch
        nx, ny = row.grid
        nf = row.nf
        self.mask = row.masked if row.masked is not None else mask
        flag = SOME_DEFINED if self.mask > 0 else ALL_DEFINED
        self.sets, self.flag_arrays, self.keep = [], [], []
        grid_const = {}
        for _ in range(sets):
            args, flag_arrays = [], []
            shared = None
            if row.common_mask and self.mask > 0:
                shared = torch.rand((nf, ny, nx), device=inputs.device, generator=inputs.gen) < self.mask
            for it in row.items:
                if it == "nx":
                    args.append(nx)
                elif it == "ny":
                    args.append(ny)
                elif it == "nf":
                    args.append(nf)
                elif it == "UNDEF":
                    args.append(UNDEF)
                elif it == "OUT":
                    args.append(torch.empty((nf, ny, nx), dtype=torch.float32, device=inputs.device))
                elif it == "FLAGS":
                    f = np.full(nf, flag, np.int32)
                    flag_arrays.append((f, f.copy()))
                    args.append(f)
                elif it == "MFLAGS":
                    f = np.full(nf * self._M(), flag, np.int32)
                    args.append(f)
                elif it[0] == "F":
                    if shared is not None:
                        a = inputs.field(it[1], row.grid, nf, 0.0)
                        a[shared] = UNDEF
                        args.append(a)
                    else:
                        args.append(inputs.field(it[1], row.grid, nf, self.mask))
                elif it[0] == "G":
                    if it[1] not in grid_const:
                        grid_const[it[1]] = inputs.field(it[1], row.grid)
                    args.append(grid_const[it[1]])
                elif it[0] in ("S", "SV"):
                    args.append(it[1])
                elif it[0] == "MEMBERS":
                    base = inputs.field(it[1], row.grid, nf, 0.0)
                    members = []
                    for _j in range(it[2]):
                        mbr = base + torch.randn(base.shape, device=inputs.device, generator=inputs.gen) * 3.0
                        if self.mask > 0:
                            mbr[torch.rand(base.shape, device=inputs.device, generator=inputs.gen) < self.mask] = UNDEF
                        members.append(mbr)
                    del base
                    args += [members, it[2]]
                elif it[0] == "LIMITS":
                    args += [np.array(it[1], np.float32), len(it[1])]
                else:
                    raise ValueError(it)
            self.sets.append(args)
            self.flag_arrays.append(flag_arrays)

    def _M(self):
        for it in self.row.items:
            if isinstance(it, tuple) and it[0] == "MEMBERS":
                return it[2]
        return 1

    def reset_flags(self):
        for fa in self.flag_arrays:
            for live, saved in fa:
                live[...] = saved

    @property
    def points(self):
        return self.row.grid[0] * self.row.grid[1] * self.row.nf

    def single_host_args(self):
        """the reference's one-field call on field 0 of set 0 (host numpy arrays); rows of the grid cut to row.cpu_rows"""
        nx, ny = self.row.grid
        rows_ = self.row.cpu_rows or ny
        rows_ = min(rows_, ny)
        out, pos = [], 0
        args = self.sets[0]
        for it in self.row.items:
            a = args[pos]
            if it == "nf":
                pos += 1
                continue
            if it == "ny":
                out.append(rows_)
            elif it in ("nx", "UNDEF"):
                out.append(a)
            elif it == "OUT":
                out.append(np.empty((rows_, nx), np.float32))
            elif it == "FLAGS":
                out.append(np.array([a[0]], np.int32))
            elif it == "MFLAGS":
                out.append(np.ascontiguousarray(a[: self._M()]))
            elif it[0] == "F":
                out.append(np.ascontiguousarray(a[0, :rows_].cpu().numpy()))
            elif it[0] == "G":
                out.append(np.ascontiguousarray(a[:rows_].cpu().numpy()))
            elif it[0] == "S":
                out.append(a)
            elif it[0] == "SV":
                out.append(float(a[0]))
            elif it[0] == "MEMBERS":
                out.append([np.ascontiguousarray(mm[0, :rows_].cpu().numpy()) for mm in a])
                out.append(args[pos + 1])
                pos += 1
            elif it[0] == "LIMITS":
                out.append(a)
                out.append(args[pos + 1])
                pos += 1
            pos += 1
        return out, rows_ * nx


def time_row(gpu, torch, stream, built, min_seconds=0.3, sampler=None):
    """device-resident time of one batched call: groups of deferred launches between two CUDA events on the launching stream
    (a drain between groups recycles the library's scratch arena), repeated until `min_seconds` of wall time have passed and
    the clock sampler has seen two samples under this load.  Returns (ms per launch, launches, wall t0, wall t1)."""
    import time
    row = built.row
    rc = gpu.call(row.call, *built.sets[0])  # warm-up, immediate mode: the call is accepted, the scratch arena has grown
    assert rc == 1, (row.name, rc, gpu.last_error())
    built.reset_flags()
    torch.cuda.synchronize()
    bytes_per_launch = row.bpp * built.points
    reps = max(1, min(8, int(6e9 // bytes_per_launch)))
    total_ms, launches, k = 0.0, 0, 0
    wall0 = time.time()
    while True:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gpu.begin_deferred()
        e0.record(stream)
        for _ in range(reps):
            gpu.call(row.call, *built.sets[k % len(built.sets)])
            k += 1
        e1.record(stream)
        gpu.end_deferred()
        torch.cuda.synchronize()
        built.reset_flags()
        total_ms += e0.elapsed_time(e1)
        launches += reps
        el = time.time() - wall0
        if (el >= min_seconds and (sampler is None or sampler.samples_since(wall0) >= 2)) or el > 4 * min_seconds + 1.0:
            break
    return total_ms / launches, launches, wall0, time.time()
