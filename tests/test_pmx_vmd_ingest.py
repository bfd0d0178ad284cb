"""PMX / VMD byte streams: the host-side parsers must produce exactly the plan and the flattened
motion that the flat-array entry points produce for the same data, including the name join."""
import numpy as np
import pytest

import pmxio
from conftest import synth_case
from simple_mmd_renderer_b200 import capi
from simple_mmd_renderer_b200.poser import HostPlan, MmdGpuError

SKIP_PLAN = set()


def _version(model: dict) -> float:
    return 2.1 if (model["skin_type"] == capi.SKIN_QDEF).any() else 2.0


def _representable(model: dict) -> dict:
    """The PMX container cannot hold an append index >= the bone count in a narrow field (the stress config has
    one, which libmmd drops): mirror what the file will say."""
    m = dict(model)
    ap = model["bone_append_parent"].copy()
    ap[(ap < 0) | (ap >= model["n_bones"])] = -1
    m["bone_append_parent"] = ap
    return m


@pytest.mark.parametrize("name", ["tiny", "tiny_full", "small"])
@pytest.mark.parametrize("utf8", [False, True])
def test_pmx_bytes_give_the_same_plan_as_arrays(name, utf8):
    cfg, model, _ = synth_case(name)
    model = _representable(model)
    want = HostPlan(arrays=model).arrays()
    got = HostPlan(pmx_bytes=pmxio.write_pmx(model, utf8=utf8, version=_version(model), extra_uv=2 if utf8 else 0)).arrays()
    for k in want:
        a, b = want[k], got[k]
        assert a.shape == b.shape, k
        np.testing.assert_array_equal(a.view(np.uint8) if a.dtype.kind == "f" else a,
                                      b.view(np.uint8) if b.dtype.kind == "f" else b, err_msg=f"plan array {k}")


def _representable_motion(motion: dict) -> dict:
    """A VMD file cannot hold a registered-but-empty track (no record = no track): drop those."""
    keep = np.flatnonzero(motion["bone_track_key_count"] > 0)
    m = dict(motion)
    m["n_bone_tracks"] = int(keep.size)
    for k in ("bone_track_bone", "bone_track_key_begin", "bone_track_key_count"):
        m[k] = motion[k][keep]
    return m


@pytest.mark.parametrize("name", ["tiny", "tiny_full"])
@pytest.mark.parametrize("utf8", [False, True])
def test_vmd_bytes_give_the_same_motion_as_arrays(name, utf8):
    cfg, model, motion = synth_case(name)
    model = _representable(model)
    motion = _representable_motion(motion)
    plan = HostPlan(pmx_bytes=pmxio.write_pmx(model, utf8=utf8, version=_version(model)))
    want = HostPlan(arrays=model).anim_from_arrays(motion, model["n_bones"], model["n_morphs"])
    got = plan.anim_from_vmd(pmxio.write_vmd(motion))
    for k in want:
        a, b = want[k], got[k]
        assert a.shape == b.shape, k
        np.testing.assert_array_equal(a.view(np.uint8) if a.dtype.kind == "f" else a,
                                      b.view(np.uint8) if b.dtype.kind == "f" else b, err_msg=f"anim array {k}")


def test_motion_storage_is_sorted_and_last_duplicate_wins():
    """std::map<frame, key> semantics (motion_impl.inl:221-227): unsorted input, duplicate frames."""
    cfg, model, motion = synth_case("tiny_full")      # stress config: shuffled keys + one duplicated frame per track
    a = HostPlan(arrays=model).anim_from_arrays(motion, model["n_bones"], model["n_morphs"])
    beg, cnt, kf = a[capi.ANIM_BONE_KEY_BEGIN], a[capi.ANIM_BONE_KEY_COUNT], a[capi.ANIM_KEY_FRAME]
    keys = motion["bone_keys"]
    kr = a[capi.ANIM_KEY_R].reshape(-1, 4)
    for t in range(int(motion["n_bone_tracks"])):
        b = int(motion["bone_track_bone"][t])
        seg = keys[int(motion["bone_track_key_begin"][t]):][:int(motion["bone_track_key_count"][t])]
        last = {}
        for k in seg:
            last[int(k["frame"])] = k
        frames = sorted(last)
        assert a[capi.ANIM_BONE_TRACKED][b] == 1
        np.testing.assert_array_equal(kf[beg[b]:beg[b] + cnt[b]], frames)
        for i, f in enumerate(frames):
            np.testing.assert_array_equal(kr[beg[b] + i].view(np.uint32), last[f]["rotation"].view(np.uint32))
    untracked = set(range(model["n_bones"])) - set(int(x) for x in motion["bone_track_bone"])
    assert untracked and all(a[capi.ANIM_BONE_TRACKED][b] == 0 for b in untracked)


def test_bezier_tables_are_deduplicated_and_linear_is_flagged():
    cfg, model, motion = synth_case("tiny")
    a = HostPlan(arrays=model).anim_from_arrays(motion, model["n_bones"], model["n_morphs"])
    curves = a[capi.ANIM_KEY_CURVE]
    n_tables = a[capi.ANIM_TABLES].size // 32
    interp = motion["bone_keys"]["interp"].reshape(-1, 4)
    distinct_nonlinear = {tuple(q) for q in interp.tolist() if not (q[0] == q[1] and q[2] == q[3])}
    assert n_tables == len(distinct_nonlinear)
    assert (curves[curves != 0xFFFFFFFF] < n_tables).all()
    lin = np.asarray([q[0] == q[1] and q[2] == q[3] for q in interp.tolist()])
    assert lin.any()


def test_malformed_streams_are_rejected():
    cfg, model, motion = synth_case("tiny")
    good = pmxio.write_pmx(model)
    for bad in (b"", b"PMD " + good[4:], good[: len(good) // 2], good[:40]):
        with pytest.raises(MmdGpuError) as e:
            HostPlan(pmx_bytes=bad)
        assert e.value.status == capi.ERR_PARSE
    plan = HostPlan(pmx_bytes=good)
    vmd = pmxio.write_vmd(motion)
    for bad in (b"Vocaloid Motion Data file".ljust(50, b"\0"), vmd[:100], vmd[: len(vmd) - 40]):
        with pytest.raises(MmdGpuError) as e:
            plan.anim_from_vmd(bad)
        assert e.value.status == capi.ERR_PARSE
    # a plan made from flat arrays has no names: a VMD cannot be joined to it
    with pytest.raises(MmdGpuError):
        HostPlan(arrays=model).anim_from_vmd(vmd)


def test_pmx21_qdef_is_accepted_and_pmx20_qdef_rejected():
    cfg, model, _ = synth_case("tiny_full")       # has QDEF-tagged vertices
    assert (model["skin_type"] == capi.SKIN_QDEF).any()
    HostPlan(pmx_bytes=pmxio.write_pmx(_representable(model), version=2.1))
    with pytest.raises(MmdGpuError) as e:
        HostPlan(pmx_bytes=pmxio.write_pmx(_representable(model), version=2.0))
    assert e.value.status == capi.ERR_PARSE


def test_parsers_survive_corrupted_streams():
    """Robustness: truncated and bit-flipped PMX / VMD streams must come back as a status code (or parse, when the
    damage hit a field that does not matter), never as a crash or a hang.  libmmd's readers throw or read out of
    bounds on such input."""
    cfg, model, motion = synth_case("tiny_full")
    pmx = pmxio.write_pmx(_representable(model), version=2.1)
    vmd = pmxio.write_vmd(_representable_motion(motion))
    plan = HostPlan(pmx_bytes=pmx)
    rng = np.random.default_rng(123)
    outcomes = {"ok": 0, "err": 0}
    for trial in range(150):
        buf = bytearray(pmx)
        mode = trial % 3
        if mode == 0:
            buf = buf[: int(rng.integers(0, len(buf)))]
        elif mode == 1:
            for _ in range(int(rng.integers(1, 6))):
                buf[int(rng.integers(0, len(buf)))] = int(rng.integers(0, 256))
        else:   # corrupt a count / index field region near the start of a section
            at = int(rng.integers(0, min(len(buf), 4096)))
            buf[at:at + 4] = bytes(int(x) for x in rng.integers(0, 256, 4))
        try:
            HostPlan(pmx_bytes=bytes(buf)).close()
            outcomes["ok"] += 1
        except MmdGpuError as e:
            assert e.status in (capi.ERR_PARSE, capi.ERR_BAD_INDEX, capi.ERR_INVALID_ARG, capi.ERR_UNSUPPORTED, capi.ERR_OOM)
            outcomes["err"] += 1
    assert outcomes["err"] > 20
    for trial in range(100):
        buf = bytearray(vmd)
        if trial % 2 == 0:
            buf = buf[: int(rng.integers(0, len(buf)))]
        else:
            for _ in range(int(rng.integers(1, 6))):
                buf[int(rng.integers(0, len(buf)))] = int(rng.integers(0, 256))
        try:
            plan.anim_from_vmd(bytes(buf))
        except MmdGpuError as e:
            assert e.status in (capi.ERR_PARSE, capi.ERR_BAD_INDEX, capi.ERR_INVALID_ARG)


def test_host_code_under_address_and_ub_sanitizers(tmp_path):
    """tools/fuzz_host.cc: PMX / VMD parsing, plan building and motion flattening compiled with
    -fsanitize=address,undefined and driven with valid, truncated and bit-flipped streams."""
    import subprocess
    from conftest import ROOT
    cfg, model, motion = synth_case("tiny_full")
    (tmp_path / "m.pmx").write_bytes(pmxio.write_pmx(_representable(model), version=2.1))
    (tmp_path / "m.vmd").write_bytes(pmxio.write_vmd(_representable_motion(motion)))
    exe = tmp_path / "fuzz_host"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                        f"-I{ROOT}/include", f"{ROOT}/tools/fuzz_host.cc", f"{ROOT}/simple_mmd_renderer_b200/csrc/host_plan.cpp",
                        f"{ROOT}/simple_mmd_renderer_b200/csrc/pmx_vmd.cpp", "-o", str(exe)], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer runtime not available: " + r.stderr[-200:])
    r = subprocess.run([str(exe), str(tmp_path / "m.pmx"), str(tmp_path / "m.vmd"), "300"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "no sanitizer report" in r.stdout


def test_material_morphs_survive_the_pmx_container():
    """Material morph records (index + 113-byte pmx_material_morph, interprete/pmx_types.inl:61-72): same grouped plan
    from bytes as from arrays; an index of 255 in a 1-byte field means every material (pmx_reader_impl.inl:327-334)."""
    from simple_mmd_renderer_b200 import synth
    cfg, model, _ = synth_case("tiny_full")
    model = synth.add_material_morphs(_representable(model))
    want = HostPlan(arrays=model, extensions=True).arrays()
    got = HostPlan(pmx_bytes=pmxio.write_pmx(model, version=_version(model)), extensions=True).arrays()
    row = want[capi.PLAN_MATERIAL_MORPH_ROW]
    assert row.size == 6 and row[-1] * 120 == want[capi.PLAN_MATERIAL_MORPH].size and row[-1] > 12
    for k in want:
        np.testing.assert_array_equal(want[k].view(np.uint8) if want[k].dtype.kind == "f" else want[k],
                                      got[k].view(np.uint8) if got[k].dtype.kind == "f" else got[k], err_msg=f"plan array {k}")
    # libmmd-exact mode carries no material plan at all
    assert HostPlan(arrays=model).arrays()[capi.PLAN_MATERIAL_MORPH].size == 0
