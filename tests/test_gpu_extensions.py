"""Extension mode (mmdgpu_options.extensions = 1): spherical SDEF, dual-quaternion QDEF, applied UV morphs.

PARITY UNPINNED — libmmd implements none of these (it lerps SDEF like BDEF2, cannot represent QDEF, ignores UV
morphs), so these tests pin the CUDA path to (1) the libmmd oracle for everything that is NOT an extension, (2) an
independent fp64 restatement of the documented formulas (tests/ext_reference.py), tolerance stated below, and
(3) self-consistency properties."""
import numpy as np
import pytest

import ext_reference as ref
from conftest import assert_bitwise, synth_case
from simple_mmd_renderer_b200 import capi
from simple_mmd_renderer_b200.poser import Frames, MmdGpuError, Model, Motion

pytestmark = pytest.mark.gpu
TOL = dict(rtol=2e-5, atol=2e-5)     # fp32 device arithmetic vs the fp64 restatement


def _norm_types(model):
    import oracle
    t, ids, w = oracle.Restatement(model, None).skinning()     # Model::Normalize output (SDEF stays 3, QDEF -> 2)
    return t, ids, w


def test_extension_formulas_against_fp64_restatement(ctx):
    import oracle
    cfg, model, motion = synth_case("tiny_full")
    assert (model["skin_type"] == capi.SKIN_SDEF).any() and (model["skin_type"] == capi.SKIN_QDEF).any()
    assert (model["morph_type"] == capi.MORPH_UV).any()
    orc = oracle.Restatement(model, motion)
    m = Model(ctx, model, extensions=True)
    a = Motion(m, motion)
    frames = [7, 42, 88]
    fr = Frames(m, 1, len(frames))
    fr.update(a, frames)
    t_norm, ids, _ = _norm_types(model)
    is_sdef = t_norm == capi.SKIN_SDEF
    is_qdef = model["skin_type"] == capi.SKIN_QDEF
    plain = ~(is_sdef | is_qdef)
    assert is_sdef.any() and is_qdef.any()
    for k, f in enumerate(frames):
        want = orc.run_frame(f)
        pos, nrm = fr.download(k, capi.STREAM_POSITION), fr.download(k, capi.STREAM_NORMAL)
        skin = fr.bone_matrices(k)
        assert_bitwise(skin, want["skin"], f"frame {f} skinning matrices (unchanged by extensions)")
        # everything that is not an extension stays bit-identical to libmmd
        assert_bitwise(pos[plain], want["pos"][plain], f"frame {f} BDEF positions")
        assert_bitwise(nrm[plain], want["nrm"][plain], f"frame {f} BDEF normals")
        T = ref.bone_transforms(skin)
        dv, duv = ref.morph_images(model, fr.morph_rates(k))
        P = model["position"].astype(np.float64) + dv
        N = model["normal"].astype(np.float64)
        for i in np.flatnonzero(is_sdef)[:200]:
            p, n = ref.sdef(P[i], N[i], int(ids[i, 0]), int(ids[i, 1]), float(model["weight"][i, 0]),
                            model["sdef_c"][i].astype(np.float64), model["sdef_r0"][i].astype(np.float64),
                            model["sdef_r1"][i].astype(np.float64), T)
            np.testing.assert_allclose(pos[i], p, err_msg=f"SDEF position vertex {i}", **TOL)
            np.testing.assert_allclose(nrm[i], n, err_msg=f"SDEF normal vertex {i}", **TOL)
        for i in np.flatnonzero(is_qdef)[:200]:
            p, n = ref.qdef(P[i], N[i], [int(x) for x in model["bone_id"][i]], model["weight"][i].astype(np.float64), T)
            np.testing.assert_allclose(pos[i], p, err_msg=f"QDEF position vertex {i}", **TOL)
            np.testing.assert_allclose(nrm[i], n, err_msg=f"QDEF normal vertex {i}", **TOL)
        uv = fr.download(k, capi.STREAM_UV)
        np.testing.assert_allclose(uv, model["uv"].astype(np.float64) + duv, rtol=1e-6, atol=1e-6, err_msg=f"frame {f} UV")
    # interleaved layout carries the morphed UV in the record
    fi = Frames(m, 1, 1, capi.LAYOUT_INTERLEAVED_SOKOL32)
    fi.update(a, [42])
    rec = fi.download(0, capi.STREAM_INTERLEAVED)
    assert_bitwise(rec[:, 6:8], fr.download(1, capi.STREAM_UV), "interleaved uv")
    assert_bitwise(rec[:, 3:6], fr.download(1, capi.STREAM_NORMAL), "interleaved normal")


def _rig(n_vertices=64):
    """Hand-made rig: root <- parent <- child, vertices tagged SDEF (parent/child) and QDEF."""
    rng = np.random.default_rng(11)
    nv = n_vertices
    pos = rng.uniform(-1, 1, (nv, 3)).astype(np.float32)
    nrm = rng.normal(size=(nv, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True).astype(np.float32)
    st = np.where(np.arange(nv) % 2 == 0, capi.SKIN_SDEF, capi.SKIN_QDEF).astype(np.uint8)
    bid = np.zeros((nv, 4), np.int32)
    bid[:, 0], bid[:, 1], bid[:, 2], bid[:, 3] = 2, 1, 0, 1
    w = np.zeros((nv, 4), np.float32)
    w[st == capi.SKIN_SDEF, 0] = rng.uniform(0.2, 0.8, int((st == capi.SKIN_SDEF).sum())).astype(np.float32)
    w4 = rng.random((nv, 4)).astype(np.float32) + 0.1
    w4 /= w4.sum(1, keepdims=True)
    w[st == capi.SKIN_QDEF] = w4[st == capi.SKIN_QDEF]
    C = (pos + rng.uniform(-0.1, 0.1, (nv, 3))).astype(np.float32)
    R0 = (C + rng.uniform(-0.3, 0.3, (nv, 3))).astype(np.float32)
    R1 = (C + rng.uniform(-0.3, 0.3, (nv, 3))).astype(np.float32)
    uvm = np.zeros(5, capi.UV_MORPH_ENTRY)
    uvm["vertex"] = [0, 3, 3, 10, 63]
    uvm["offset"] = rng.uniform(-0.2, 0.2, (5, 4)).astype(np.float32)
    return dict(
        n_vertices=nv, position=pos, normal=nrm, uv=rng.random((nv, 2)).astype(np.float32), skin_type=st, bone_id=bid,
        weight=w, sdef_c=C, sdef_r0=R0, sdef_r1=R1,
        n_bones=3, bone_position=np.asarray([[0, 0, 0], [0, 1, 0], [0, 2, 0]], np.float32),
        bone_parent=np.asarray([-1, 0, 1], np.int32), bone_transform_level=np.zeros(3, np.int32),
        bone_flags=np.zeros(3, np.uint16), bone_append_parent=np.full(3, -1, np.int32),
        bone_append_ratio=np.zeros(3, np.float32), ik_target=np.full(3, -1, np.int32), ik_iterations=np.zeros(3, np.int32),
        ik_angle_limit=np.zeros(3, np.float32), ik_link_begin=np.zeros(3, np.uint32), ik_link_count=np.zeros(3, np.uint32),
        n_ik_links=0, ik_link_bone=np.zeros(0, np.int32), ik_link_has_limit=np.zeros(0, np.uint8),
        ik_link_lo=np.zeros((0, 3), np.float32), ik_link_hi=np.zeros((0, 3), np.float32),
        n_morphs=1, morph_type=np.asarray([capi.MORPH_UV], np.uint8), morph_entry_begin=np.zeros(1, np.uint32),
        morph_entry_count=np.asarray([5], np.uint32),
        n_vertex_morph_entries=0, vertex_morph_entries=np.zeros(0, capi.VERTEX_MORPH_ENTRY),
        n_uv_morph_entries=5, uv_morph_entries=uvm,
        n_bone_morph_entries=0, bone_morph_entries=np.zeros(0, capi.BONE_MORPH_ENTRY),
        n_group_morph_entries=0, group_morph_entries=np.zeros(0, capi.GROUP_MORPH_ENTRY))


def _posed(ctx, model, poses, ext, morph_w=0.0):
    m = Model(ctx, model, extensions=ext)
    fr = Frames(m, 1, 1)
    fr.reset_posing()
    for b, (t, q) in poses.items():
        fr.set_bone_pose(0, b, t, q)
    fr.set_morph_pose(0, 0, morph_w)
    fr.pre_physics_posing()
    fr.post_physics_posing()
    fr.deform()
    out = dict(pos=fr.download(0, capi.STREAM_POSITION), nrm=fr.download(0, capi.STREAM_NORMAL))
    if ext:
        out["uv"] = fr.download(0, capi.STREAM_UV)
    return out


def test_sdef_and_dqs_reduce_to_linear_blend_for_pure_translations(ctx):
    """With identity rotations every bone shares q = 1: SDEF == BDEF2 lerp and DQS == BDEF4 blend (== libmmd compat)."""
    model = _rig()
    ident = [0, 0, 0, 1]
    poses = {0: ([0.3, -0.2, 0.1], ident), 1: ([0.0, 0.5, -0.4], ident), 2: ([-0.7, 0.1, 0.2], ident)}
    a = _posed(ctx, model, poses, ext=True)
    b = _posed(ctx, model, poses, ext=False)
    np.testing.assert_allclose(a["pos"], b["pos"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(a["nrm"], b["nrm"], rtol=1e-5, atol=1e-5)


def test_dqs_with_one_dominant_bone_is_a_rigid_transform(ctx):
    model = _rig()
    q = model["skin_type"] == capi.SKIN_QDEF
    model["weight"][q] = [1, 0, 0, 0]
    h = np.float32(np.sqrt(0.5))
    poses = {2: ([0.1, 0.2, 0.3], [h, 0, 0, h]), 1: ([0, 0, 0], [0, h, 0, h])}
    a = _posed(ctx, model, poses, ext=True)
    b = _posed(ctx, model, poses, ext=False)      # BDEF4 with weights (1,0,0,0) == the single bone's matrix
    np.testing.assert_allclose(a["pos"][q], b["pos"][q], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(a["nrm"][q], b["nrm"][q], rtol=1e-5, atol=2e-5)
    # rigid: lengths of normals are preserved by a quaternion rotation
    np.testing.assert_allclose(np.linalg.norm(a["nrm"][q], axis=1), 1.0, atol=1e-5)


def test_uv_morph_is_linear_in_the_weight(ctx):
    model = _rig()
    u0 = _posed(ctx, model, {}, ext=True, morph_w=0.0)["uv"]
    u1 = _posed(ctx, model, {}, ext=True, morph_w=0.35)["uv"]
    u2 = _posed(ctx, model, {}, ext=True, morph_w=0.70)["uv"]
    assert_bitwise(u0, model["uv"], "UV at weight 0")
    np.testing.assert_allclose(u2 - u0, 2.0 * (u1 - u0), rtol=1e-5, atol=1e-6)
    touched = np.zeros(model["n_vertices"], bool)
    touched[model["uv_morph_entries"]["vertex"]] = True
    assert_bitwise(u1[~touched], model["uv"][~touched], "untouched UVs")
    assert (np.abs(u1[touched] - model["uv"][touched]).max(axis=1) > 0).all()


def test_uv_stream_needs_extensions(ctx):
    cfg, model, motion = synth_case("tiny")
    fr = Frames(Model(ctx, model), 1, 1)
    with pytest.raises(MmdGpuError) as e:
        fr.download(0, capi.STREAM_UV)
    assert e.value.status == capi.ERR_INVALID_ARG


# ---------------------------------------------------------------- material morph accumulation (extension)
def _material_case():
    from simple_mmd_renderer_b200 import synth
    cfg, model, motion = synth_case("tiny_full")
    return synth.add_material_morphs(model), motion


def _material_rates(model, seed=5):
    rng = np.random.default_rng(seed)
    w = rng.uniform(0.0, 1.0, int(model["n_morphs"])).astype(np.float32)
    w[rng.random(w.size) < 0.2] = 0.0
    return w


def test_material_images_against_fp64_restatement(ctx):
    model, _ = _material_case()
    m = Model(ctx, model, extensions=True)
    assert m.n_materials == 5
    fr = Frames(m, 3, 1)
    np.testing.assert_array_equal(fr.material_images(1)[:, 0], 1.0)     # a fresh Poser: ResetPosing state
    np.testing.assert_array_equal(fr.material_images(1)[:, 1], 0.0)
    weights = [_material_rates(model, s) for s in (5, 6, 7)]
    fr.reset_posing()
    for slot, w in enumerate(weights):
        for i, x in enumerate(w):
            fr.set_morph_pose(slot, i, float(x))
    fr.pre_physics_posing()
    fr.post_physics_posing()
    for slot, w in enumerate(weights):
        want = ref.material_images(model, w)
        got = fr.material_images(slot)
        assert (np.abs(want[:, 0] - 1.0).max() > 1e-3) and (np.abs(want[:, 1]).max() > 1e-3)
        np.testing.assert_allclose(got, want, **TOL)
    fr.reset_posing()
    np.testing.assert_array_equal(fr.material_images(0)[:, 0], 1.0)
    np.testing.assert_array_equal(fr.material_images(0)[:, 1], 0.0)


def test_material_images_through_the_fused_update_and_a_clip(ctx):
    from simple_mmd_renderer_b200 import synth
    model, _ = _material_case()
    cfg = synth.TINY_FULL
    motion = synth.make_motion(cfg, model)          # keys every morph, the material ones included
    m = Model(ctx, model, extensions=True)
    mo = Motion(m, motion)
    fr = Frames(m, 1, 4)
    frames = [3, 17, 18, 40]
    fr.update([mo], frames)
    for slot in range(4):
        want = ref.material_images(model, fr.morph_rates(slot))
        np.testing.assert_allclose(fr.material_images(slot), want, **TOL)
    # the deformation itself does not depend on material morphs: same positions as the libmmd-exact model
    plain = Frames(Model(ctx, model), 1, 4)
    plain.update([Motion(plain.model, motion)], frames)
    for slot in range(4):
        t = _norm_types(model)[0]
        keep = (t != capi.SKIN_SDEF) & (model["skin_type"] != capi.SKIN_QDEF)
        assert_bitwise(fr.download(slot, capi.STREAM_POSITION)[keep], plain.download(slot, capi.STREAM_POSITION)[keep],
                       "positions of non-extension vertices")


def test_material_image_properties(ctx):
    """Self-consistency: a multiply entry at rate 1 yields its value; additive images are linear in the rate; an entry
    for "every material" reaches all of them; a skipped morph (rate < 1e-7) leaves the images untouched."""
    from simple_mmd_renderer_b200 import synth
    cfg, base, _ = synth_case("tiny")
    model = synth.add_material_morphs(base, n_materials=3, n_morphs=2, entries_per_morph=1, in_group=False)
    ent = model["material_morph_entries"].copy()
    ent["material"] = [1, -1]
    ent["method"] = [capi.MATERIAL_MUL, capi.MATERIAL_ADD]
    model["material_morph_entries"] = ent
    first = int(model["n_morphs"]) - 2
    m = Model(ctx, model, extensions=True)

    def images(w_mul, w_add):
        fr = Frames(m, 1, 1)
        fr.reset_posing()
        fr.set_morph_pose(0, first, w_mul)
        fr.set_morph_pose(0, first + 1, w_add)
        fr.pre_physics_posing()
        return fr.material_images(0)
    a = images(1.0, 0.0)
    np.testing.assert_allclose(a[1, 0], ent["value"][0], rtol=1e-6)
    np.testing.assert_array_equal(a[[0, 2], 0], 1.0)
    np.testing.assert_array_equal(a[:, 1], 0.0)
    b1, b2 = images(0.0, 0.3), images(0.0, 0.6)
    np.testing.assert_allclose(b2[:, 1], 2.0 * b1[:, 1], rtol=1e-6)
    for t in range(3):
        np.testing.assert_allclose(b1[t, 1], ent["value"][1] * np.float32(0.3), rtol=1e-6)
    c = images(5e-8, -1.0)                                 # both skipped: below 1e-7 / negative (poser_impl.inl:329)
    np.testing.assert_array_equal(c[:, 0], 1.0)
    np.testing.assert_array_equal(c[:, 1], 0.0)


def test_material_images_are_untouched_in_libmmd_exact_mode(ctx):
    """libmmd allocates the images as 1 / 0 and never fills them (poser_impl.inl:31-36, 355-358)."""
    model, _ = _material_case()
    fr = Frames(Model(ctx, model), 1, 1)
    for i, x in enumerate(_material_rates(model)):
        fr.set_morph_pose(0, i, float(x))
    fr.pre_physics_posing()
    img = fr.material_images(0)
    assert img.shape == (5, 2, capi.MATERIAL_FIELDS)
    np.testing.assert_array_equal(img[:, 0], 1.0)
    np.testing.assert_array_equal(img[:, 1], 0.0)
