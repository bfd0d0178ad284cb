"""The C-ABI library loads and exports every symbol include/mmdgpu.h declares; host-only entry points behave."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, has_gpu, synth_case
from simple_mmd_renderer_b200 import capi, lib
from simple_mmd_renderer_b200.poser import MmdGpuError, bezier_table, plan_arrays

HEADER = os.path.join(ROOT, "include", "mmdgpu.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"MMDGPU_API\s+[\w\s\*]+?\b(mmdgpu_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 48
    h = lib.load()
    for n in names:
        assert hasattr(h, n), f"{n} declared in mmdgpu.h but not exported by libmmdgpu.so"
        assert n in lib.SIGNATURES, f"{n} has no ctypes signature in lib.py"
    assert set(lib.SIGNATURES) == set(names)


def test_header_compiles_as_c_and_cxx(tmp_path):
    for comp, std, ext in (("gcc", "-std=c99", "c"), ("g++", "-std=c++11", "cc")):
        src = tmp_path / f"t.{ext}"
        src.write_text('#include "mmdgpu.h"\nint main(void){ return mmdgpu_version() > 0 ? 0 : 1; }\n')
        r = subprocess.run([comp, std, "-Wall", "-Werror", "-pedantic", f"-I{ROOT}/include", "-fsyntax-only", str(src)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_library_is_sm100a_only():
    r = subprocess.run(["cuobjdump", "-lelf", lib.SO_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\w+", r.stdout))
    assert archs == {"sm_100a"}, archs


def test_version_and_status_strings():
    h = lib.load()
    assert h.mmdgpu_version() >= 1
    for s in range(0, -7, -1):
        assert h.mmdgpu_status_string(s)


def test_no_device_means_error_not_fallback():
    """Without a GPU the library refuses to create a context: there is no CPU path."""
    h = lib.load()
    ctx = C.c_void_p()
    st = h.mmdgpu_context_create(0, None, C.byref(ctx))
    if has_gpu():
        assert st == capi.OK
        h.mmdgpu_context_destroy(ctx)
    else:
        assert st == capi.ERR_CUDA and not ctx.value
        assert h.mmdgpu_last_error(None)


def test_null_arguments_are_rejected():
    h = lib.load()
    assert h.mmdgpu_reset_posing(None) == capi.ERR_INVALID_ARG
    assert h.mmdgpu_deform(None) == capi.ERR_INVALID_ARG
    assert h.mmdgpu_update(None, None, None) == capi.ERR_INVALID_ARG
    assert h.mmdgpu_context_synchronize(None) == capi.ERR_INVALID_ARG
    assert h.mmdgpu_model_vertex_count(None) == 0
    assert h.mmdgpu_plan_create(None, None, None, None, 0) == capi.ERR_INVALID_ARG


def test_bad_indices_are_rejected_where_libmmd_reads_out_of_bounds():
    cfg, model, _ = synth_case("tiny")
    bad = dict(model)
    ids = model["bone_id"].copy()
    ids[5, 0] = model["n_bones"] + 3
    bad["bone_id"] = ids
    with pytest.raises(MmdGpuError) as e:
        plan_arrays(bad)
    assert e.value.status == capi.ERR_BAD_INDEX
    bad = dict(model)
    ent = model["vertex_morph_entries"].copy()
    ent["vertex"][0] = model["n_vertices"]
    bad["vertex_morph_entries"] = ent
    with pytest.raises(MmdGpuError) as e:
        plan_arrays(bad)
    assert e.value.status == capi.ERR_BAD_INDEX


def test_group_morph_cycle_is_rejected():
    cfg, model, _ = synth_case("tiny_full")
    bad = dict(model)
    mt = model["morph_type"]
    groups = np.flatnonzero(mt == capi.MORPH_GROUP)
    assert groups.size >= 2
    ge = model["group_morph_entries"].copy()
    g0 = int(groups[0])
    ge["morph"][int(model["morph_entry_begin"][g0])] = g0      # a group that contains itself
    bad["group_morph_entries"] = ge
    with pytest.raises(MmdGpuError) as e:
        plan_arrays(bad)
    assert e.value.status == capi.ERR_BAD_INDEX


def test_bezier_table_linear_and_endpoints():
    assert bezier_table([20, 20, 107, 107]) is None          # c0.x == c0.y and c1.x == c1.y -> linear
    t = bezier_table([10, 90, 100, 30])
    assert t is not None and t.shape == (32,)
    assert 0 < t[0] < 1e-5 and abs(t[31] - 1.0) < 1e-5      # tab[0] is ~2.7e-7, not 0
    assert (np.diff(t) > -1e-6).all()


def test_cxx_shim_and_example_compile_and_link(tmp_path):
    """include/mmdgpu.hpp (the libmmd-named C++ mirror) and examples/headless_update.cc build against the library."""
    exe = tmp_path / "headless_update"
    r = subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-Werror", f"-I{ROOT}/include",
                        f"{ROOT}/examples/headless_update.cc", lib.SO_PATH, "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
