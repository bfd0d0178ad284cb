"""Index tier (bit-exact): Model::Normalize rewrite, evaluation order, IK classification, morph application
slots, per-vertex morph CSR and the wave schedule of the bone program — all host-side, no GPU."""
import numpy as np
import pytest

import oracle
from conftest import synth_case
from golden_util import load_golden
from simple_mmd_renderer_b200 import capi, synth
from simple_mmd_renderer_b200.poser import plan_arrays

CASES = ["tiny", "tiny_full", "small", "C2"]


@pytest.mark.parametrize("name", CASES)
def test_normalize_rewrite_matches_oracle(name):
    cfg, model, motion = synth_case(name)
    plan = plan_arrays(model)
    t, ids, w = oracle.Restatement(model, None).skinning()
    np.testing.assert_array_equal(plan[capi.PLAN_SKIN_TYPE], t)
    pid = plan[capi.PLAN_BONE_ID].reshape(-1, 4).astype(np.int32)
    pw = plan[capi.PLAN_WEIGHT].reshape(-1, 4)
    # the plan additionally folds Deform's Lerp shortcuts (w < 1e-7 -> bone 1, w > 0.99999988 -> bone 0) into the
    # stream; everything else must be identical to Model::Normalize's output
    two = (t == capi.SKIN_BDEF2) | (t == capi.SKIN_SDEF)
    lo = two & (w[:, 0] < np.float32(1e-7))
    hi = two & (w[:, 0] > np.float32(1.0 - 1e-7))
    plain = ~(lo | hi)
    n_ids = np.where(t == capi.SKIN_BDEF1, 1, np.where(t == capi.SKIN_BDEF4, 4, 2))
    for k in range(4):
        use = plain & (n_ids > k)
        np.testing.assert_array_equal(pid[use, k], ids[use, k])
    np.testing.assert_array_equal(pid[lo, 0], ids[lo, 1])
    np.testing.assert_array_equal(pid[hi, 0], ids[hi, 0])
    four = t == capi.SKIN_BDEF4
    np.testing.assert_array_equal(pw[four].view(np.uint32), w[four].view(np.uint32))
    np.testing.assert_array_equal(pw[two & plain, 0].view(np.uint32), w[two & plain, 0].view(np.uint32))


@pytest.mark.parametrize("name", ["tiny", "tiny_full", "small", "ik_zoo"])
def test_normalize_and_ik_class_match_libmmd_golden(name):
    cfg, model, motion = synth_case(name)
    g = load_golden(name)
    plan = plan_arrays(model)
    np.testing.assert_array_equal(plan[capi.PLAN_SKIN_TYPE], g["norm_type"])
    np.testing.assert_array_equal(plan[capi.PLAN_IK_FIX_TYPE], g["ik_fix"])
    np.testing.assert_array_equal(plan[capi.PLAN_IK_EULER_ORDER], g["ik_order"])


@pytest.mark.parametrize("name", CASES)
def test_evaluation_order_is_libmmd_sort(name):
    """poser_impl.inl:100-109, 500-510: two lists split by the post-physics flag, sorted by (level, index)."""
    cfg, model, _ = synth_case(name)
    plan = plan_arrays(model)
    flags, level = model["bone_flags"], model["bone_transform_level"]
    nb = model["n_bones"]
    pre = sorted((b for b in range(nb) if not flags[b] & capi.BONE_POST_PHYSICS), key=lambda b: (int(level[b]), b))
    post = sorted((b for b in range(nb) if flags[b] & capi.BONE_POST_PHYSICS), key=lambda b: (int(level[b]), b))
    np.testing.assert_array_equal(plan[capi.PLAN_ORDER_PRE], pre)
    np.testing.assert_array_equal(plan[capi.PLAN_ORDER_POST], post)
    # program order = EVAL(b) [IK(b)] over pre, SKIN over pre, then the same over post
    kinds, bones = plan[capi.PLAN_OP_KIND], plan[capi.PLAN_OP_BONE]
    want = []
    for lst in (pre, post):
        for b in lst:
            want.append((0, b))
            if flags[b] & capi.BONE_HAS_IK:
                want.append((1, b))
        want += [(2, b) for b in lst]
    assert list(zip(kinds.tolist(), bones.tolist())) == want


def _app_slots(model):
    """DFS expansion of UpdateMorphTransform's recursion (poser_impl.inl:328-360)."""
    mt, mb, mc = model["morph_type"], model["morph_entry_begin"], model["morph_entry_count"]
    ge = model["group_morph_entries"]
    slots = []

    def visit(m, parent, mult):
        me = len(slots)
        slots.append((m, parent, mult))
        if mt[m] == capi.MORPH_GROUP:
            for j in range(int(mc[m])):
                e = ge[int(mb[m]) + j]
                visit(int(e["morph"]), me, np.float32(e["rate"]))
    for m in range(int(model["n_morphs"])):
        visit(m, -1, np.float32(1.0))
    return slots


@pytest.mark.parametrize("name", CASES)
def test_morph_csr_is_the_transpose_in_application_order(name):
    cfg, model, _ = synth_case(name)
    plan = plan_arrays(model)
    slots = _app_slots(model)
    np.testing.assert_array_equal(plan[capi.PLAN_APP_SLOT_MORPH], [s[0] for s in slots])
    np.testing.assert_array_equal(plan[capi.PLAN_APP_SLOT_PARENT], [s[1] for s in slots])
    np.testing.assert_array_equal(plan[capi.PLAN_APP_SLOT_MULT].view(np.uint32),
                                  np.asarray([s[2] for s in slots], np.float32).view(np.uint32))
    # brute-force transpose: (vertex, slot, entry order) sorted
    mt, mb, mc = model["morph_type"], model["morph_entry_begin"], model["morph_entry_count"]
    ve = model["vertex_morph_entries"]
    vs, ss, offs = [], [], []
    for s, (m, _, _) in enumerate(slots):
        if mt[m] != capi.MORPH_VERTEX:
            continue
        seg = ve[int(mb[m]):int(mb[m]) + int(mc[m])]
        vs.append(seg["vertex"].astype(np.int64))
        ss.append(np.full(seg.size, s, np.int64))
        offs.append(seg["offset"])
    nv = int(model["n_vertices"])
    if vs:
        v = np.concatenate(vs); s = np.concatenate(ss); o = np.concatenate(offs)
        order = np.argsort(v, kind="stable")       # stable: keeps (slot, entry order) inside a vertex
        row = np.zeros(nv + 1, np.uint32)
        np.add.at(row, v + 1, 1)
        row = np.cumsum(row).astype(np.uint32)
        np.testing.assert_array_equal(plan[capi.PLAN_CSR_ROW_PTR], row)
        np.testing.assert_array_equal(plan[capi.PLAN_CSR_SLOT], s[order].astype(np.uint32))
        np.testing.assert_array_equal(plan[capi.PLAN_CSR_OFFSET].reshape(-1, 3).view(np.uint32), o[order].view(np.uint32))
    else:
        assert plan[capi.PLAN_CSR_SLOT].size == 0


def _op_sets(model, kind, b):
    """External read set and write set of one program op, restated from poser_impl.inl:142-311."""
    nb = int(model["n_bones"])
    flags, parent = model["bone_flags"], model["bone_parent"]
    app = model["bone_append_parent"]
    is_link = np.zeros(nb, bool)
    for ikb in range(nb):
        if flags[ikb] & capi.BONE_HAS_IK:
            lb, lc = int(model["ik_link_begin"][ikb]), int(model["ik_link_count"][ikb])
            is_link[model["ik_link_bone"][lb:lb + lc]] = True

    def eval_sets(x):
        R, W = [], []
        if flags[x] & (capi.BONE_APPEND_ROTATE | capi.BONE_APPEND_TRANSLATE) and 0 <= app[x] < nb:
            R.append(("tot", int(app[x])))
        if is_link[x]:
            R.append(("ik", x)); W.append(("pre", x))
        if 0 <= parent[x] < nb:
            R.append(("local", int(parent[x])))
        W += [("tot", x), ("local", x)]
        return R, W
    if kind == 0:
        return eval_sets(b)
    if kind == 2:
        return [("local", b)], [("skin", b)]
    R, W = [], []
    wrote = set()

    def solve(ikb):
        """The IK block of UpdateBoneTransform(ikb); links and target are re-evaluated with UpdateBoneTransform itself
        (poser_impl.inl:203-206), so one that has IK runs its own solve right there."""
        lb, lc = int(model["ik_link_begin"][ikb]), int(model["ik_link_count"][ikb])
        links = [int(x) for x in model["ik_link_bone"][lb:lb + lc]]
        if ("local", ikb) not in wrote:
            R.append(("local", ikb))
        for l in links:
            W.append(("ik", l)); wrote.add(("ik", l))
        for x in list(reversed(links)) + [int(model["ik_target"][ikb])]:
            r2, w2 = eval_sets(x)
            R.extend(v for v in r2 if v not in wrote)
            W.extend(w2)
            wrote.update(w2)
            if flags[x] & capi.BONE_HAS_IK:
                solve(x)
        for l in links:
            if 0 <= parent[l] < nb and ("local", int(parent[l])) not in wrote:
                R.append(("local", int(parent[l])))
    solve(b)
    return R, W


@pytest.mark.parametrize("name", ["tiny", "tiny_full", "small", "C2", "ik_zoo", "ik_nested"])
def test_wave_schedule_preserves_sequential_semantics(name):
    """Every op must observe, in the wave program, exactly the writers it observes in libmmd's sequential program,
    and ops that share a wave must not touch each other's state."""
    cfg, model, _ = synth_case(name)
    plan = plan_arrays(model)
    kinds, bones, wave = plan[capi.PLAN_OP_KIND], plan[capi.PLAN_OP_BONE], plan[capi.PLAN_OP_WAVE]
    wb, wops = plan[capi.PLAN_WAVE_BEGIN], plan[capi.PLAN_WAVE_OPS]
    n = kinds.size
    sets = [_op_sets(model, int(kinds[i]), int(bones[i])) for i in range(n)]

    def observed(order_groups):
        last = {}
        seen = [None] * n
        for grp in order_groups:
            # all ops of a group read the state left by earlier groups
            for i in grp:
                seen[i] = tuple(sorted((v, last.get(v, -1)) for v in set(sets[i][0])))
            for i in grp:
                for v in sets[i][1]:
                    last[v] = i
        return seen, last
    seq_seen, seq_last = observed([[i] for i in range(n)])
    groups = [[int(x) for x in wops[wb[w]:wb[w + 1]]] for w in range(wb.size - 1)]
    assert sorted(sum(groups, [])) == list(range(n))
    for w, grp in enumerate(groups):
        assert all(wave[i] == w for i in grp)
        touched = {}
        for i in grp:
            for v in set(sets[i][0]) | set(sets[i][1]):
                touched.setdefault(v, []).append(i)
        for v, ops in touched.items():
            writers = [i for i in ops if v in sets[i][1]]
            assert not (writers and len(set(ops)) > 1), f"wave {w}: ops {ops} conflict on {v}"
    wav_seen, wav_last = observed(groups)
    assert wav_seen == seq_seen
    assert wav_last == seq_last
    split = int(plan[capi.PLAN_WAVE_PHASE_SPLIT][0])
    post = set(int(b) for b in plan[capi.PLAN_ORDER_POST])
    for i in range(n):
        b = int(bones[i])
        assert (wave[i] >= split) == (b in post), "pre / post physics ops must not share a launch segment"


def test_waves_are_shallow_on_the_headline_model():
    """C3 (1 k bones): the schedule must expose parallelism — far fewer waves than bones."""
    cfg = synth.CONFIGS["C3"]
    small = synth.SynthConfig("c3_bones", cfg.config_id, 2000, cfg.n_bones, 4, 30)
    plan = plan_arrays(synth.make_model(small))
    n_waves = plan[capi.PLAN_WAVE_BEGIN].size - 1
    assert n_waves < 64


@pytest.mark.parametrize("name", ["tiny", "tiny_full", "small", "C2"])
def test_tile_layout_is_a_faithful_permutation(name):
    """Device vertex layout: every TILE-vertex tile is stored in a tile-local order; the permutation, the
    tile-local bone lists and the sliced-ELL morph table must reproduce the PMX-order arrays exactly."""
    cfg, model, _ = synth_case(name)
    plan = plan_arrays(model)
    TILE, V = 512, 4            # kTileVerts, kVertsPerThread (host_plan.hpp)
    WARPS = TILE // V // 32     # warps per CTA; group g = step j * WARPS + warp w
    GROUPS = TILE // 32
    nv = int(model["n_vertices"])
    orig = plan[capi.PLAN_TILE_ORIG].astype(np.int64)
    nvp = orig.size
    assert nvp % TILE == 0 and nvp >= nv and nvp - nv < TILE
    n_tiles = nvp // TILE
    tiles = orig.reshape(n_tiles, TILE)
    assert (np.sort(tiles, axis=1) == np.arange(TILE)).all(), "tile_orig must be a permutation of each tile"
    src = (np.arange(nvp) // TILE) * TILE + orig           # PMX vertex of every storage position
    real = src < nv
    # device type per PMX vertex, derived independently from the Normalize output + Deform's Lerp shortcuts
    t = plan[capi.PLAN_SKIN_TYPE].astype(np.int64)
    pw = plan[capi.PLAN_WEIGHT].reshape(-1, 4)
    pid = plan[capi.PLAN_BONE_ID].reshape(-1, 4).astype(np.int64)
    two = (t == capi.SKIN_BDEF2) | (t == capi.SKIN_SDEF)
    dev = np.where(t == capi.SKIN_BDEF1, 0, np.where(t == capi.SKIN_BDEF4, 2, 1))
    dev = np.where(two & ((pw[:, 0] < np.float32(1e-7)) | (pw[:, 0] > np.float32(1.0 - 1e-7))), 0, dev)
    st_type = plan[capi.PLAN_TILE_TYPE].astype(np.int64)
    np.testing.assert_array_equal(st_type[real], dev[src[real]])
    assert (st_type[~real] == 0).all()
    # tile-local bone indices map back to the global ids
    tb_begin = plan[capi.PLAN_TILE_BONE_BEGIN].astype(np.int64)
    tb = plan[capi.PLAN_TILE_BONES].astype(np.int64)
    local = plan[capi.PLAN_TILE_LOCAL_ID].reshape(-1, 4).astype(np.int64)
    n_ids = np.where(st_type == 0, 1, np.where(st_type == 2, 4, 2))
    for ti in range(n_tiles):
        bones = tb[tb_begin[ti]:tb_begin[ti + 1]]
        assert (np.diff(bones) > 0).all()
        sl = slice(ti * TILE, (ti + 1) * TILE)
        for k in range(4):
            use = real[sl] & (n_ids[sl] > k)
            np.testing.assert_array_equal(bones[local[sl][use, k]], pid[src[sl][use], k])
        used = set()
        for k in range(4):
            used |= set(pid[src[sl][real[sl] & (n_ids[sl] > k)], k].tolist())
        assert used == set(bones.tolist()) or (not used and bones.tolist() == [0])
    # sliced ELL -> per-vertex entry lists == CSR rows, application order kept; padding points at the zero slot
    row = plan[capi.PLAN_CSR_ROW_PTR].astype(np.int64)
    cslot = plan[capi.PLAN_CSR_SLOT]
    coff = plan[capi.PLAN_CSR_OFFSET].reshape(-1, 3)
    base = plan[capi.PLAN_ELL_BASE].astype(np.int64)
    rounds = plan[capi.PLAN_ELL_ROUNDS].astype(np.int64)
    eslot = plan[capi.PLAN_ELL_SLOT]
    eoff = plan[capi.PLAN_ELL_OFFSET].reshape(-1, 3)
    pad = plan[capi.PLAN_APP_SLOT_MORPH].size
    total_real = 0
    for ti in range(n_tiles):
        for j in range(V):
            for w in range(WARPS):
                g = ti * GROUPS + j * WARPS + w
                cnts = []
                for l in range(32):
                    pos = ti * TILE + (w * 32 + l) * V + j
                    v = src[pos]
                    cnt = int(row[v + 1] - row[v]) if v < nv else 0
                    cnts.append(cnt)
                    at = base[g] + np.arange(rounds[g]) * 32 + l
                    np.testing.assert_array_equal(eslot[at[:cnt]], cslot[row[v]:row[v] + cnt] if v < nv else [])
                    if cnt:
                        np.testing.assert_array_equal(eoff[at[:cnt]].view(np.uint32), coff[row[v]:row[v] + cnt].view(np.uint32))
                    assert (eslot[at[cnt:]] == pad).all() and (eoff[at[cnt:]] == 0).all()
                    total_real += cnt
                assert rounds[g] == max(cnts)
    assert total_real == cslot.size
    # the point of the sort: a warp step (32 consecutive ranks) is one skinning type except at the boundaries
    mixed = 0
    for ti in range(n_tiles):
        for j in range(V):
            for w in range(WARPS):
                pos = ti * TILE + (w * 32 + np.arange(32)) * V + j
                mixed += len(set(st_type[pos].tolist())) > 1
    assert mixed <= 3 * n_tiles


@pytest.mark.parametrize("name", ["small", "C2"])
def test_tile_order_pairs_lanes_on_their_leading_bones(name):
    """The tile order puts two vertices of one skinning type that agree on their leading bone ids into an aligned lane
    pair (2k, 2k + 1), and fills 32-lane groups with pairs of one agreement level, so that the skinning kernel's LDS.128
    of bone k reads ONE palette cell per lane pair (2 shared-memory wavefronts instead of 4).  Checked here: the share of
    (group, type, bone position) load sets in which every lane pair agrees, and that it is what the order was built for
    (the same model ordered without pairing has almost none)."""
    import os, subprocess, sys, json
    cfg, model, _ = synth_case(name)

    def fast_share(plan):
        TILE, V = 512, 4
        WARPS = TILE // V // 32
        st = plan[capi.PLAN_TILE_TYPE].astype(np.int64)
        n_tiles = st.size // TILE
        j, w, l = np.meshgrid(np.arange(V), np.arange(WARPS), np.arange(32), indexing="ij")
        pos = ((w * 32 + l) * V + j).reshape(V * WARPS, 32)
        ty = st.reshape(n_tiles, TILE)[:, pos]
        lid = plan[capi.PLAN_TILE_LOCAL_ID].reshape(-1, 4).astype(np.int64).reshape(n_tiles, TILE, 4)[:, pos]
        fast = total = 0
        for t, keep in {0: 1, 1: 2, 2: 4}.items():
            m = ty == t
            present = m.any(axis=2)
            same_type = m[:, :, 0::2] == m[:, :, 1::2]
            for k in range(keep):
                agree = (lid[:, :, 0::2, k] == lid[:, :, 1::2, k]) | ~m[:, :, 0::2]
                fast += int(((agree & same_type).all(axis=2) & present).sum())
                total += int(present.sum())
        return fast / max(1, total)

    share = fast_share(plan_arrays(model))
    assert share > 0.45, share
    # the same plan without pairing, in a fresh process (the knob is read once per process)
    code = ("import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r);"
            "from conftest import synth_case; from simple_mmd_renderer_b200.poser import plan_arrays;"
            "from simple_mmd_renderer_b200 import capi; import numpy as np;"
            "cfg, model, _ = synth_case(%r); p = plan_arrays(model);"
            "print(json.dumps({'orig': p[capi.PLAN_TILE_ORIG].tolist()}))") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                                 os.path.dirname(os.path.abspath(__file__)), name)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True,
                         env=dict(os.environ, MMDGPU_TILE_PAIRING="0")).stdout
    orig_unpaired = np.asarray(json.loads(out.strip().splitlines()[-1])["orig"])
    assert not np.array_equal(orig_unpaired, plan_arrays(model)[capi.PLAN_TILE_ORIG]), "the pairing knob must change the order"


def test_material_morph_plan_is_grouped_by_material_in_application_order():
    """Extension plan: entries under each material appear in application-slot (DFS) order, an "every material" entry
    is repeated under each material, and values / methods are carried over unchanged."""
    from simple_mmd_renderer_b200 import synth
    cfg, model, _ = synth_case("tiny_full")
    model = synth.add_material_morphs(model)
    plan = plan_arrays(model, extensions=True)
    row = plan[capi.PLAN_MATERIAL_MORPH_ROW]
    rec = plan[capi.PLAN_MATERIAL_MORPH].view(np.dtype([("node", "<i4"), ("method", "<u4"), ("value", "<f4", (28,))]))
    node_morph = plan[capi.PLAN_APP_SLOT_MORPH]
    nmat = int(model["n_materials"])
    assert row.size == nmat + 1 and row[0] == 0 and row[-1] == rec.size
    ent, mb, mc = model["material_morph_entries"], model["morph_entry_begin"], model["morph_entry_count"]
    for mat in range(nmat):
        want = []
        for node, m in enumerate(node_morph):
            if model["morph_type"][m] != capi.MORPH_MATERIAL:
                continue
            for e in ent[int(mb[m]):int(mb[m]) + int(mc[m])]:
                if e["material"] < 0 or e["material"] >= nmat or e["material"] == mat:
                    want.append((node, int(e["method"]), e["value"]))
        got = rec[row[mat]:row[mat + 1]]
        assert len(got) == len(want), mat
        for g, (node, method, value) in zip(got, want):
            assert g["node"] == node and g["method"] == method
            np.testing.assert_array_equal(g["value"].view(np.uint32), value.view(np.uint32))
    # a material morph inside a group shows up once per visit (direct + through the group)
    n_material_morphs = int((model["morph_type"] == capi.MORPH_MATERIAL).sum())
    assert sum(1 for m in node_morph if model["morph_type"][m] == capi.MORPH_MATERIAL) == n_material_morphs + 2


BONE_STATIC = np.dtype([("local_offset", "<f4", (3,)), ("parent", "<i4"), ("position", "<f4", (3,)), ("append_parent", "<i4"),
                        ("append_ratio", "<f4"), ("flags", "<u4"), ("link_slot", "<i4"), ("morph_slot", "<i4")])
IK_DESC = np.dtype([("bone", "<i4"), ("target", "<i4"), ("iterations", "<i4"), ("angle_limit", "<f4"), ("link_begin", "<i4"),
                    ("link_count", "<i4"), ("pad", "<i4", (2,))])
IK_LINK = np.dtype([("bone", "<i4"), ("limited", "u1"), ("fix", "u1"), ("order", "u1"), ("pad", "u1"), ("lo", "<f4", (3,)),
                    ("hi", "<f4", (3,))])
IK_IMAGE = np.dtype([("bones_begin", "<i4"), ("n_bones", "<i4"), ("lslots_begin", "<i4"), ("n_lslots", "<i4"),
                     ("mslots_begin", "<i4"), ("n_mslots", "<i4"), ("region_f4", "<i4"), ("pad", "<i4")])
HAS_PARENT, APPEND_ROT, APPEND_TRANS, IS_LINK = 1, 2, 4, 8


@pytest.mark.parametrize("name", ["tiny_full", "small", "C2", "ik_zoo"])
def test_ik_images_are_faithful_renumberings(name):
    """Chain-local images of the CCD IK solves (device design): every image holds the solve's links, target and IK bone
    plus the parents / append parents evaluating them reads; the translated static records, IK descriptor and links
    name the same bones as the originals once mapped back through the image's bone list; read-only bones carry no
    references; the evaluated set is exactly links + target."""
    cfg, model, _ = synth_case(name)
    plan = plan_arrays(model)
    static = plan[capi.PLAN_BONE_STATIC].view(BONE_STATIC)
    desc = plan[capi.PLAN_IK_DESC].view(IK_DESC)
    links = plan[capi.PLAN_IK_LINK].view(IK_LINK)
    img = plan[capi.PLAN_IK_IMAGE].view(IK_IMAGE)
    ibones, iwritten = plan[capi.PLAN_IK_IMAGE_BONES], plan[capi.PLAN_IK_IMAGE_WRITTEN]
    istatic = plan[capi.PLAN_IK_IMAGE_STATIC].view(BONE_STATIC)
    ils, ims = plan[capi.PLAN_IK_IMAGE_LINK_SLOTS], plan[capi.PLAN_IK_IMAGE_MORPH_SLOTS]
    idesc = plan[capi.PLAN_IK_IMAGE_DESC].view(IK_DESC)
    ilinks = plan[capi.PLAN_IK_IMAGE_LINKS].view(IK_LINK)
    assert img.size == desc.size == idesc.size > 0
    for k in range(desc.size):
        I, d, di = img[k], desc[k], idesc[k]
        gb = ibones[I["bones_begin"]:I["bones_begin"] + I["n_bones"]]
        wr = iwritten[I["bones_begin"]:I["bones_begin"] + I["n_bones"]]
        st = istatic[I["bones_begin"]:I["bones_begin"] + I["n_bones"]]
        ls = ils[I["lslots_begin"]:I["lslots_begin"] + I["n_lslots"]]
        ms = ims[I["mslots_begin"]:I["mslots_begin"] + I["n_mslots"]]
        assert len(set(gb.tolist())) == gb.size, "image bones are distinct"
        assert I["region_f4"] % 2 == 1 and I["region_f4"] >= 7 * I["n_bones"] + 2 * I["n_lslots"] + 2 * I["n_mslots"]
        chain = [int(links[d["link_begin"] + j]["bone"]) for j in range(d["link_count"])]
        assert gb[di["bone"]] == d["bone"] and gb[di["target"]] == d["target"]
        assert di["iterations"] == d["iterations"] and di["angle_limit"] == d["angle_limit"] and di["link_count"] == d["link_count"]
        for j in range(d["link_count"]):
            a, b = links[d["link_begin"] + j], ilinks[di["link_begin"] + j]
            assert gb[b["bone"]] == a["bone"]
            for f in ("limited", "fix", "order"):
                assert a[f] == b[f]
            np.testing.assert_array_equal(a["lo"], b["lo"]); np.testing.assert_array_equal(a["hi"], b["hi"])
        assert sorted(gb[wr != 0].tolist()) == sorted(set(chain + [int(d["target"])])), "evaluated = links + target"
        for i in range(gb.size):
            o, t = static[gb[i]], st[i]
            np.testing.assert_array_equal(o["local_offset"], t["local_offset"]); np.testing.assert_array_equal(o["position"], t["position"])
            if wr[i]:
                assert t["flags"] == o["flags"] and t["append_ratio"] == o["append_ratio"]
                if o["flags"] & HAS_PARENT:
                    assert gb[t["parent"]] == o["parent"]
                if o["flags"] & (APPEND_ROT | APPEND_TRANS):
                    assert gb[t["append_parent"]] == o["append_parent"]
                assert (t["link_slot"] < 0) == (o["link_slot"] < 0) and (t["morph_slot"] < 0) == (o["morph_slot"] < 0)
                if o["link_slot"] >= 0:
                    assert ls[t["link_slot"]] == o["link_slot"]
                if o["morph_slot"] >= 0:
                    assert ms[t["morph_slot"]] == o["morph_slot"]
            else:
                assert t["parent"] == -1 and t["append_parent"] == -1 and t["link_slot"] == -1 and t["morph_slot"] == -1
                assert not (t["flags"] & (HAS_PARENT | APPEND_ROT | APPEND_TRANS | IS_LINK))
