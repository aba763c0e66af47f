"""ctypes binding of the C-ABI declared in include/fcb200.h.

The prototypes are not restated here: they are parsed from include/fcb200_api.inc and
include/fcb200_batched.inc, so the binding cannot drift from the header.  The same parser binds
the test oracles (oracle/libfcoracle.so `fco_`, oracle/_ref/libfcref.so `fcref_`), which share the
signatures of fcb200_api.inc -- that binding is made by tests/, never by this package.

Pointer arguments accept
  * numpy float32 / int32 arrays (host memory),
  * torch tensors (host or CUDA; `.data_ptr()` is passed through),
  * raw integer addresses, or None (NULL).
"""
from __future__ import annotations

import ctypes
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
INCLUDE = os.path.join(ROOT, "include")
LIB_PATH = os.path.join(HERE, "lib", "libfcb200.so")

ALL_DEFINED, NONE_DEFINED, SOME_DEFINED = 0, 1, 2

_CT = {
    "int": ctypes.c_int,
    "float": ctypes.c_float,
    "const char*": ctypes.c_char_p,
    "const float*": ctypes.c_void_p,
    "float*": ctypes.c_void_p,
    "int*": ctypes.c_void_p,
    "const int*": ctypes.c_void_p,
    "const float* const*": ctypes.c_void_p,
}


def parse_inc(path: str, macro: str) -> dict:
    """{name: [(ctype_string, arg_name), ...]} for every `macro(name, (args))` entry of an .inc file."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(macro + r"\(\s*(\w+)\s*,\s*\((.*?)\)\s*\)", text, flags=re.S):
        name, args = m.group(1), m.group(2)
        parsed = []
        for a in args.split(","):
            a = " ".join(a.split())
            mm = re.match(r"^(.*?)(\w+)$", a)
            typ = mm.group(1).strip().replace(" *", "*")
            parsed.append((typ, mm.group(2)))
        out[name] = parsed
    return out


def _addr(x):
    """address of a host/device buffer, plus an object to keep alive"""
    if x is None:
        return None, None
    if isinstance(x, int):
        return x, None
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("array arguments must be C-contiguous")
        return x.ctypes.data, x
    if hasattr(x, "data_ptr"):  # torch tensor
        if not x.is_contiguous():
            raise ValueError("tensor arguments must be contiguous")
        return x.data_ptr(), x
    raise TypeError("unsupported buffer argument: %r" % type(x))


class Api:
    """One shared library exposing `<prefix><op>` for the operators of the given .inc files."""

    def __init__(self, lib_path: str, prefix: str, batched: bool):
        if not os.path.exists(lib_path):
            raise FileNotFoundError(
                "%s is missing: build it first (python -c 'import __graft_entry__ as g; g.build()'); "
                "there is no fallback implementation" % lib_path)
        self.lib = ctypes.CDLL(lib_path)
        self.prefix = prefix
        self.sigs = parse_inc(os.path.join(INCLUDE, "fcb200_api.inc"), "FC_FN")
        if batched:
            self.sigs.update(parse_inc(os.path.join(INCLUDE, "fcb200_batched.inc"), "FCB_FN"))
        self.fns = {}
        for name, args in self.sigs.items():
            fn = getattr(self.lib, prefix + name)
            fn.restype = ctypes.c_int
            fn.argtypes = [_CT[t] for t, _ in args]
            self.fns[name] = fn

    def has(self, name: str) -> bool:
        return name in self.fns

    def call(self, name: str, *args) -> int:
        sig = self.sigs[name]
        if len(args) != len(sig):
            raise TypeError("%s expects %d arguments (%s), got %d" % (name, len(sig), ", ".join(n for _, n in sig), len(args)))
        conv, keep = [], []
        for (typ, argname), v in zip(sig, args):
            if typ == "int":
                conv.append(int(v))
            elif typ == "float":
                conv.append(float(v))
            elif typ == "const char*":
                conv.append(v.encode() if isinstance(v, str) else v)
            elif typ == "const float* const*":
                addrs = []
                for item in v:
                    a, k = _addr(item)
                    addrs.append(a)
                    keep.append(k)
                table = (ctypes.c_void_p * max(1, len(addrs)))(*addrs)
                keep.append(table)
                conv.append(ctypes.cast(table, ctypes.c_void_p))
            else:
                if typ in ("int*", "const int*") and isinstance(v, np.ndarray) and v.dtype != np.int32:
                    raise TypeError("%s.%s must be an int32 array" % (name, argname))
                if typ in ("const float*", "float*") and isinstance(v, np.ndarray) and v.dtype != np.float32:
                    raise TypeError("%s.%s must be a float32 array" % (name, argname))
                a, k = _addr(v)
                keep.append(k)
                conv.append(a)
        return self.fns[name](*conv)


class Fcb200(Api):
    """The product library.  Raises if libfcb200.so is missing -- there is no CPU fallback."""

    def __init__(self, lib_path: str = LIB_PATH):
        super().__init__(lib_path, "fcb200_", batched=True)
        L = self.lib
        L.fcb200_last_error.restype = ctypes.c_char_p
        L.fcb200_version.restype = ctypes.c_char_p
        L.fcb200_launch_count.restype = ctypes.c_ulonglong
        L.fcb200_set_stream.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.fcb200_set_device.argtypes = [ctypes.c_int]
        L.fcb200_graph_end.argtypes = [ctypes.POINTER(ctypes.c_void_p)]
        for f in (L.fcb200_graph_launch, L.fcb200_graph_kernels, L.fcb200_graph_destroy):
            f.argtypes = [ctypes.c_void_p]
        L.fcb200_slab_unique_id.argtypes = [ctypes.c_char_p]
        L.fcb200_slab_init.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p]
        L.fcb200_slab_partition.argtypes = [ctypes.c_int] * 4 + [ctypes.POINTER(ctypes.c_int)] * 4
        L.fcb200_slab_exchange.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.fcb200_slab_combine_flags.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.fcb200_slab_bytes_sent.restype = ctypes.c_ulonglong

    def last_error(self) -> str:
        return self.lib.fcb200_last_error().decode()

    def version(self) -> str:
        return self.lib.fcb200_version().decode()

    def device_count(self) -> int:
        return self.lib.fcb200_device_count()

    def set_device(self, device: int) -> None:
        self._check(self.lib.fcb200_set_device(device))

    def set_stream(self, cuda_stream, use_it: bool = True) -> None:
        self._check(self.lib.fcb200_set_stream(ctypes.c_void_p(cuda_stream or 0), 1 if use_it else 0))

    def begin_deferred(self) -> None:
        self._check(self.lib.fcb200_begin_deferred())

    def end_deferred(self) -> None:
        self._check(self.lib.fcb200_end_deferred())

    def synchronize(self) -> None:
        self._check(self.lib.fcb200_synchronize())

    def launch_count(self) -> int:
        return int(self.lib.fcb200_launch_count())

    # ---- graphs (include/fcb200.h): a chain of calls on device-resident fields, captured once, replayed with one launch
    def graph_begin(self) -> None:
        self._check(self.lib.fcb200_graph_begin())

    def graph_end(self) -> int:
        """Close the capture; returns the graph handle.  Raises (with the failing call's error) if a captured call failed."""
        h = ctypes.c_void_p()
        self._check(self.lib.fcb200_graph_end(ctypes.byref(h)))
        return h.value

    def graph_launch(self, graph: int) -> None:
        self._check(self.lib.fcb200_graph_launch(graph))

    def graph_kernels(self, graph: int) -> int:
        return self._check(self.lib.fcb200_graph_kernels(graph))

    def graph_destroy(self, graph: int) -> None:
        self._check(self.lib.fcb200_graph_destroy(graph))

    # ---- row slabs (include/fcb200.h): one large grid over several GPUs, NCCL halo exchange
    def slab_unique_id(self) -> bytes:
        buf = ctypes.create_string_buffer(128)
        self._check(self.lib.fcb200_slab_unique_id(buf))
        return buf.raw

    def slab_init(self, rank: int, nranks: int, unique_id: bytes) -> None:
        self._check(self.lib.fcb200_slab_init(rank, nranks, ctypes.create_string_buffer(unique_id, 128)))

    def slab_finalize(self) -> None:
        self.lib.fcb200_slab_finalize()

    def slab_partition(self, ny: int, halo: int, rank: int, nranks: int):
        """(r0, r1, lo, hi): owned rows and extended-slab rows of `rank`, or None if ny is too small"""
        v = [ctypes.c_int() for _ in range(4)]
        if self.lib.fcb200_slab_partition(ny, halo, rank, nranks, *[ctypes.byref(x) for x in v]) != 1:
            return None
        return tuple(x.value for x in v)

    def slab_exchange(self, ext, nx: int, ext_rows: int, nfields: int, halo: int) -> None:
        addr, _keep = _addr(ext)
        self._check(self.lib.fcb200_slab_exchange(addr, nx, ext_rows, nfields, halo))

    def slab_combine_flags(self, flags: np.ndarray) -> None:
        assert flags.dtype == np.int32 and flags.flags["C_CONTIGUOUS"]
        self._check(self.lib.fcb200_slab_combine_flags(flags.ctypes.data, flags.size))

    def slab_reduce_flags(self) -> None:
        """deferred mode: enqueue the cross-rank combination of the flags of the calls queued so far (final at end_deferred)"""
        self._check(self.lib.fcb200_slab_reduce_flags())

    def slab_bytes_sent(self) -> int:
        return int(self.lib.fcb200_slab_bytes_sent())

    def _check(self, r: int) -> int:
        if r < 0:
            raise RuntimeError(self.last_error() or "fcb200 runtime error")
        return r

    def call(self, name: str, *args) -> int:  # noqa: D102 - same contract as Api.call, loud on runtime errors
        return self._check(super().call(name, *args))


_singleton = None


def load() -> Fcb200:
    global _singleton
    if _singleton is None:
        # FCB200_LIB: a development switch (tools/shape_variants.sh builds libraries that differ in one kernel shape)
        _singleton = Fcb200(os.environ.get("FCB200_LIB") or LIB_PATH)
    return _singleton
