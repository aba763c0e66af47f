"""CPU checks of the boundary: the C-ABI library loads and exports every symbol include/*.h declares
(no compute call is made: there is no GPU here), and fails loudly instead of falling back."""
import ctypes
import importlib
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = importlib.import_module("mi-fieldcalc_b200")
capi = pkg.capi


def test_library_exports_every_declared_symbol():
    assert os.path.exists(capi.LIB_PATH), "libfcb200.so not built: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(capi.LIB_PATH)
    sigs = capi.parse_inc(os.path.join(capi.INCLUDE, "fcb200_api.inc"), "FC_FN")
    sigs.update(capi.parse_inc(os.path.join(capi.INCLUDE, "fcb200_batched.inc"), "FCB_FN"))
    assert len(sigs) >= 64
    runtime = ["fcb200_slab_reduce_flags", "fcb200_graph_begin", "fcb200_graph_end", "fcb200_graph_launch", "fcb200_graph_kernels", "fcb200_graph_destroy", "fcb200_version", "fcb200_last_error", "fcb200_device_count", "fcb200_set_device", "fcb200_set_stream", "fcb200_begin_deferred",
               "fcb200_end_deferred", "fcb200_in_deferred", "fcb200_synchronize", "fcb200_launch_count", "fcb200_slab_unique_id", "fcb200_slab_init",
               "fcb200_slab_finalize", "fcb200_slab_rank", "fcb200_slab_nranks", "fcb200_slab_partition", "fcb200_slab_exchange",
               "fcb200_slab_combine_flags", "fcb200_slab_bytes_sent"]
    # every prototype of fcb200.h outside the two .inc lists is in `runtime` (the header is the contract)
    import re
    header = open(os.path.join(capi.INCLUDE, "fcb200.h")).read()
    declared = set(re.findall(r"\b(fcb200_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", header, flags=re.S)))
    assert declared == set(runtime), sorted(declared ^ set(runtime))
    missing = [n for n in ["fcb200_" + k for k in sigs] + runtime if not hasattr(lib, n)]
    assert not missing, missing


def test_every_single_field_entry_has_a_batched_twin():
    single = capi.parse_inc(os.path.join(capi.INCLUDE, "fcb200_api.inc"), "FC_FN")
    batched = capi.parse_inc(os.path.join(capi.INCLUDE, "fcb200_batched.inc"), "FCB_FN")
    assert sorted(k + "_batched" for k in single) == sorted(k for k in batched if k not in ("alevel_chain_batched", "hlevel_chain_batched"))


def test_oracle_and_reference_share_the_signature_list(oracle):
    single = capi.parse_inc(os.path.join(capi.INCLUDE, "fcb200_api.inc"), "FC_FN")
    assert set(oracle.fns) == set(single)


def test_no_silent_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    api = pkg.load()
    assert api.device_count() < 0
    a = np.ones((4, 4), np.float32)
    flag = np.array([0], np.int32)
    with pytest.raises(RuntimeError):
        api.call("fieldOPERfield", 1, 4, 4, a, a, a.copy(), flag, 1e35)
    assert "fcb200" in api.last_error()


def test_slab_partition_is_the_python_partition():
    """fcb200_slab_partition (C++, what a slab caller uses) == distributed.partition_rows / slab_bounds (what the gloo tests prove)"""
    D = importlib.import_module("mi-fieldcalc_b200.distributed")
    api = pkg.load()
    for ny in (3, 9, 41, 1069, 1801):
        for world in (1, 2, 3, 8):
            for halo in (1, 2):
                for rank in range(world):
                    got = api.slab_partition(ny, halo, rank, world)
                    r0, r1 = D.partition_rows(ny, world)[rank]
                    lo, hi = D.slab_bounds(r0, r1, ny, halo)
                    if D.min_rows_ok(ny, world, halo):
                        assert got == (r0, r1, lo, hi), (ny, world, halo, rank, got)
                    elif got is not None:
                        assert got == (r0, r1, lo, hi)
