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


def test_deep_group_morph_chain_is_refused_not_overflowed():
    """A crafted chain of nested group morphs (each holding the next) must come back as a status, not as a host stack
    overflow: the expansion is iterative and capped at 64 levels (libmmd itself recurses once per level)."""
    cfg, model, _ = synth_case("tiny")

    def chain(n):
        m = dict(model)
        base = int(model["n_morphs"])
        m["n_morphs"] = base + n
        m["morph_type"] = np.concatenate([model["morph_type"], np.full(n, capi.MORPH_GROUP, np.uint8)])
        ge = np.zeros(n, capi.GROUP_MORPH_ENTRY)
        ge["morph"] = np.concatenate([base + 1 + np.arange(n - 1), [0]]).astype(np.uint32)   # i -> i + 1 ... -> morph 0
        ge["rate"] = 1.0
        m["group_morph_entries"], m["n_group_morph_entries"] = ge, n
        m["morph_entry_begin"] = np.concatenate([model["morph_entry_begin"], np.arange(n, dtype=np.uint32)])
        m["morph_entry_count"] = np.concatenate([model["morph_entry_count"], np.ones(n, np.uint32)])
        return m
    plan = plan_arrays(chain(40))                         # legal depth
    assert plan[capi.PLAN_APP_SLOT_MORPH].size > 40
    with pytest.raises(MmdGpuError) as e:
        plan_arrays(chain(300_000))
    assert e.value.status == capi.ERR_UNSUPPORTED


def test_ik_solves_that_reach_each_other_are_refused():
    """An IK bone that is (transitively) a link or target of its own solve makes libmmd recurse forever
    (poser_impl.inl:203-206); nesting deeper than three levels is refused as unsupported."""
    from simple_mmd_renderer_b200 import synth
    model, _ = synth.make_ik_nested()
    plan_arrays(model)                                    # three levels: accepted
    ikb = np.flatnonzero(model["bone_flags"] & capi.BONE_HAS_IK)
    bad = dict(model)
    tgt = model["ik_target"].copy()
    tgt[ikb[0]] = ikb[0]                                  # its own target
    bad["ik_target"] = tgt
    with pytest.raises(MmdGpuError) as e:
        plan_arrays(bad)
    assert e.value.status == capi.ERR_BAD_INDEX
    # a fourth level: the innermost solve's target becomes an IK bone as well
    depth = {}

    def levels(b):
        lb, lc = int(model["ik_link_begin"][b]), int(model["ik_link_count"][b])
        kids = [int(x) for x in model["ik_link_bone"][lb:lb + lc]] + [int(model["ik_target"][b])]
        return 1 + max([levels(k) for k in kids if model["bone_flags"][k] & capi.BONE_HAS_IK] or [0])
    outer = [int(b) for b in ikb if levels(int(b)) == 3][0]
    inner = int(model["ik_target"][int(model["ik_target"][outer])])       # level 3 solve sits on this bone
    assert model["bone_flags"][inner] & capi.BONE_HAS_IK
    deep = dict(model)
    fl = model["bone_flags"].copy()
    leaf_target = int(model["ik_target"][inner])
    fl[leaf_target] |= capi.BONE_HAS_IK
    deep["bone_flags"] = fl
    for k, v in (("ik_target", 0), ("ik_iterations", 2), ("ik_angle_limit", 1.0), ("ik_link_begin", 0), ("ik_link_count", 0)):
        a = model[k].copy(); a[leaf_target] = v; deep[k] = a
    with pytest.raises(MmdGpuError) as e:
        plan_arrays(deep)
    assert e.value.status == capi.ERR_UNSUPPORTED


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
