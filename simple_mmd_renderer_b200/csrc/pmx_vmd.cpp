// pmx_vmd.cpp — PMX 2.0 / 2.1 and VMD byte streams -> the flat descriptors of include/mmdgpu.h.
//
// Layouts follow what the reference reads: L/reader/pmx_reader_impl.inl:16-449 with the packed records of
// L/reader/interprete/pmx_types.inl:17-95, and L/reader/vmd_reader_impl.inl:9-79 with
// L/reader/interprete/vmd_types.inl:17-37 (L/ = 3rd_party/libmmd/include/mmd/).  Only what the deformation
// path consumes is kept (vertices, bones, morphs; bone and morph key frames); everything else is skipped.
//
// Deliberate deviations from libmmd's readers:
//   * names are joined byte-exactly after decoding to UTF-16 here (CP932 table for VMD, UTF-16LE / UTF-8 for
//     PMX) instead of through iconv / mbstowcs, which break the join on glibc;
//   * PMX 2.1 and skinning type 4 (QDEF) are accepted; soft-body / flip / impulse records are skipped;
//   * every read is bounds-checked and reports MMDGPU_ERR_PARSE instead of throwing.
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "host_plan.hpp"

namespace mmdgpu {

namespace {

const uint16_t kSjis[60 * 188] = {
#include "sjis_table.inc"
};

struct Reader {
    const uint8_t* p;
    size_t n, at = 0;
    bool ok = true;
    Reader(const void* b, size_t len) : p(static_cast<const uint8_t*>(b)), n(len) {}
    bool need(size_t k) {
        if (!ok || k > n - at) { ok = false; return false; }
        return true;
    }
    template <class T> T get() {
        T v{};
        if (need(sizeof(T))) { std::memcpy(&v, p + at, sizeof(T)); at += sizeof(T); }
        return v;
    }
    void skip(size_t k) { if (need(k)) at += k; }
    void floats(float* dst, int k) { for (int i = 0; i < k; ++i) dst[i] = get<float>(); }
    // FileReader::ReadIndex (L/util/dwarf_impl.inl:84-103) for bone / morph / material / rigid-body indices.
    // libmmd widens 1- and 2-byte indices unsigned, so the PMX "-1" arrives as 255 / 65535 there and is
    // caught (for parents) by the `< bone_num` test; here every width maps its all-ones pattern to -1.
    int32_t index(int size) {
        switch (size) {
        case 1: { uint8_t v = get<uint8_t>(); return v == 0xFFu ? -1 : int32_t(v); }
        case 2: { uint16_t v = get<uint16_t>(); return v == 0xFFFFu ? -1 : int32_t(v); }
        case 4: return get<int32_t>();
        default: ok = false; return -1;
        }
    }
    // Vertex indices are unsigned at every width (PMX specification; identical in libmmd for 1 and 2 bytes).
    uint32_t vindex(int size) {
        switch (size) {
        case 1: return get<uint8_t>();
        case 2: return get<uint16_t>();
        case 4: return uint32_t(get<int32_t>());
        default: ok = false; return 0;
        }
    }
    std::string text() {  // int32 byte length + raw bytes
        const int32_t len = get<int32_t>();
        if (len < 0 || !need(size_t(len))) { ok = false; return std::string(); }
        std::string s(reinterpret_cast<const char*>(p + at), size_t(len));
        at += size_t(len);
        return s;
    }
};

mmdgpu_status fail(std::string& err, const std::string& msg) {
    err = msg;
    return MMDGPU_ERR_PARSE;
}

// ---- name decoding to UTF-16 code units
std::u16string from_utf16le(const std::string& s) {
    std::u16string r;
    for (size_t i = 0; i + 1 < s.size(); i += 2) r.push_back(char16_t(uint8_t(s[i]) | (uint16_t(uint8_t(s[i + 1])) << 8)));
    return r;
}
std::u16string from_utf8(const std::string& s) {
    std::u16string r;
    size_t i = 0;
    while (i < s.size()) {
        const uint8_t c = uint8_t(s[i]);
        uint32_t cp;
        int extra;
        if (c < 0x80) { cp = c; extra = 0; }
        else if ((c & 0xE0) == 0xC0) { cp = c & 0x1F; extra = 1; }
        else if ((c & 0xF0) == 0xE0) { cp = c & 0x0F; extra = 2; }
        else if ((c & 0xF8) == 0xF0) { cp = c & 0x07; extra = 3; }
        else { cp = 0xFFFD; extra = 0; }
        ++i;
        for (int k = 0; k < extra && i < s.size(); ++k, ++i) cp = (cp << 6) | (uint8_t(s[i]) & 0x3F);
        if (cp >= 0x10000) {
            cp -= 0x10000;
            r.push_back(char16_t(0xD800 + (cp >> 10)));
            r.push_back(char16_t(0xDC00 + (cp & 0x3FF)));
        } else r.push_back(char16_t(cp));
    }
    return r;
}
// CP932 bytes, terminated by the first NUL (mmd_string<15>, L/util/dwarf.inl)
std::u16string from_cp932(const char* b, size_t n) {
    std::u16string r;
    size_t i = 0;
    while (i < n && b[i] != 0) {
        const uint8_t c = uint8_t(b[i]);
        const bool lead = (c >= 0x81 && c <= 0x9F) || (c >= 0xE0 && c <= 0xFC);
        if (lead && i + 1 < n) {
            const uint8_t t = uint8_t(b[i + 1]);
            const int li = (c <= 0x9F) ? (c - 0x81) : (31 + c - 0xE0);
            int ti = -1;
            if (t >= 0x40 && t <= 0x7E) ti = t - 0x40;
            else if (t >= 0x80 && t <= 0xFC) ti = 63 + t - 0x80;
            const uint16_t u = (ti >= 0) ? kSjis[li * 188 + ti] : 0;
            r.push_back(u ? char16_t(u) : char16_t(0xFFFD));
            i += 2;
        } else if (c >= 0xA1 && c <= 0xDF) {
            r.push_back(char16_t(0xFF61 + (c - 0xA1)));  // half-width katakana
            ++i;
        } else {
            r.push_back(char16_t(c));  // ASCII; a truncated lead byte at the end of the field stays as is
            ++i;
        }
    }
    return r;
}

}  // namespace

void ParsedModel::finish() {
    desc = mmdgpu_model_desc{};
    desc.n_vertices = uint32_t(skin_type.size());
    desc.position = position.data(); desc.normal = normal.data(); desc.uv = uv.data();
    desc.skin_type = skin_type.data(); desc.bone_id = bone_id.data(); desc.weight = weight.data();
    desc.sdef_c = sdef_c.data(); desc.sdef_r0 = sdef_r0.data(); desc.sdef_r1 = sdef_r1.data();
    desc.n_bones = uint32_t(bone_parent.size());
    desc.bone_position = bone_position.data(); desc.bone_parent = bone_parent.data();
    desc.bone_transform_level = bone_transform_level.data(); desc.bone_flags = bone_flags.data();
    desc.bone_append_parent = bone_append_parent.data(); desc.bone_append_ratio = bone_append_ratio.data();
    desc.ik_target = ik_target.data(); desc.ik_iterations = ik_iterations.data();
    desc.ik_angle_limit = ik_angle_limit.data(); desc.ik_link_begin = ik_link_begin.data();
    desc.ik_link_count = ik_link_count.data();
    desc.n_ik_links = uint32_t(ik_link_bone.size());
    desc.ik_link_bone = ik_link_bone.data(); desc.ik_link_has_limit = ik_link_has_limit.data();
    desc.ik_link_lo = ik_link_lo.data(); desc.ik_link_hi = ik_link_hi.data();
    desc.n_morphs = uint32_t(morph_type.size());
    desc.morph_type = morph_type.data(); desc.morph_entry_begin = morph_entry_begin.data();
    desc.morph_entry_count = morph_entry_count.data();
    desc.n_vertex_morph_entries = uint32_t(vme.size()); desc.vertex_morph_entries = vme.data();
    desc.n_uv_morph_entries = uint32_t(uvme.size()); desc.uv_morph_entries = uvme.data();
    desc.n_bone_morph_entries = uint32_t(bme.size()); desc.bone_morph_entries = bme.data();
    desc.n_group_morph_entries = uint32_t(gme.size()); desc.group_morph_entries = gme.data();
    desc.n_materials = n_materials;
    desc.n_material_morph_entries = uint32_t(mme.size()); desc.material_morph_entries = mme.data();
}

mmdgpu_status parse_pmx(const void* bytes, size_t n, ParsedModel& o, std::string& err) {
    o = ParsedModel();
    Reader r(bytes, n);
    // header: "PMX " + float version + u8 count of flag bytes (8) — pmx_reader_impl.inl:21-29
    char magic[4];
    for (char& c : magic) c = char(r.get<uint8_t>());
    const float version = r.get<float>();
    const uint8_t n_flags = r.get<uint8_t>();
    if (!r.ok || std::memcmp(magic, "PMX ", 4) != 0) return fail(err, "not a PMX file");
    if (!(version == 2.0f || version == 2.1f)) return fail(err, "unsupported PMX version");
    if (n_flags < 8) return fail(err, "PMX header has fewer than 8 flag bytes");
    o.utf8 = r.get<uint8_t>() > 0;
    const int extra_uv = r.get<uint8_t>();
    const int vsz = r.get<uint8_t>(), tsz = r.get<uint8_t>(), msz = r.get<uint8_t>(), bsz = r.get<uint8_t>(),
              mosz = r.get<uint8_t>(), rsz = r.get<uint8_t>();
    (void)rsz;
    r.skip(size_t(n_flags) - 8);
    if (extra_uv > 4) return fail(err, "more than 4 extra UV channels");
    for (int i = 0; i < 4; ++i) r.text();  // name, name_en, description, description_en
    if (!r.ok) return fail(err, "truncated PMX header");

    // ---- vertices, pmx_reader_impl.inl:49-104
    const int32_t nv = r.get<int32_t>();
    if (nv < 0) return fail(err, "negative vertex count");
    if (size_t(nv) > (n - r.at) / 38 + 1) return fail(err, "vertex count exceeds the file size");
    o.position.resize(size_t(nv) * 3); o.normal.resize(size_t(nv) * 3); o.uv.resize(size_t(nv) * 2);
    o.skin_type.resize(size_t(nv)); o.bone_id.assign(size_t(nv) * 4, 0); o.weight.assign(size_t(nv) * 4, 0.0f);
    o.sdef_c.assign(size_t(nv) * 3, 0.0f); o.sdef_r0.assign(size_t(nv) * 3, 0.0f); o.sdef_r1.assign(size_t(nv) * 3, 0.0f);
    for (int32_t i = 0; i < nv && r.ok; ++i) {
        r.floats(&o.position[size_t(i) * 3], 3);
        r.floats(&o.normal[size_t(i) * 3], 3);
        r.floats(&o.uv[size_t(i) * 2], 2);
        r.skip(size_t(extra_uv) * 16);
        const uint8_t t = r.get<uint8_t>();
        int32_t* id = &o.bone_id[size_t(i) * 4];
        float* w = &o.weight[size_t(i) * 4];
        switch (t) {
        case MMDGPU_SKIN_BDEF1: id[0] = r.index(bsz); w[0] = 1.0f; break;
        case MMDGPU_SKIN_BDEF2: id[0] = r.index(bsz); id[1] = r.index(bsz); w[0] = r.get<float>(); break;
        case MMDGPU_SKIN_BDEF4:
        case MMDGPU_SKIN_QDEF:
            for (int k = 0; k < 4; ++k) id[k] = r.index(bsz);
            r.floats(w, 4);
            break;
        case MMDGPU_SKIN_SDEF:
            id[0] = r.index(bsz); id[1] = r.index(bsz); w[0] = r.get<float>();
            r.floats(&o.sdef_c[size_t(i) * 3], 3); r.floats(&o.sdef_r0[size_t(i) * 3], 3); r.floats(&o.sdef_r1[size_t(i) * 3], 3);
            break;
        default: return fail(err, "invalid skinning type at vertex " + std::to_string(i));
        }
        if (t == MMDGPU_SKIN_QDEF && version < 2.1f) return fail(err, "QDEF skinning in a PMX 2.0 file");
        // An unused BDEF4 / QDEF lane is stored as index -1 with weight 0 by some exporters; libmmd would read
        // bone_images_[SIZE_MAX].  A zero-weight lane contributes +0 for any finite matrix: point it at bone 0.
        if (t == MMDGPU_SKIN_BDEF4 || t == MMDGPU_SKIN_QDEF)
            for (int k = 0; k < 4; ++k)
                if (id[k] < 0 && w[k] == 0.0f) id[k] = 0;
        o.skin_type[size_t(i)] = t;
        r.skip(4);  // edge scale
    }
    if (!r.ok) return fail(err, "truncated vertex section");

    // ---- faces, textures, materials: skipped (pmx_reader_impl.inl:106-196)
    const int32_t n_face_idx = r.get<int32_t>();
    if (n_face_idx < 0) return fail(err, "negative face index count");
    r.skip(size_t(n_face_idx) * size_t(vsz));
    const int32_t n_tex = r.get<int32_t>();
    for (int32_t i = 0; i < n_tex && r.ok; ++i) r.text();
    const int32_t n_mat = r.get<int32_t>();
    for (int32_t i = 0; i < n_mat && r.ok; ++i) {
        r.text(); r.text();
        r.skip(65);                    // pmx_material_basic
        r.skip(size_t(tsz) * 2);       // texture, sphere texture
        r.skip(1);                     // sphere mode
        const uint8_t shared_toon = r.get<uint8_t>();
        r.skip(shared_toon ? 1 : size_t(tsz));
        r.text();                      // memo
        r.skip(4);                     // face index count
    }
    if (!r.ok || n_tex < 0 || n_mat < 0) return fail(err, "truncated face / texture / material section");
    o.n_materials = uint32_t(n_mat);

    // ---- bones, pmx_reader_impl.inl:191-265
    const int32_t nb = r.get<int32_t>();
    if (nb < 0 || size_t(nb) > (n - r.at) / 20 + 1) return fail(err, "bad bone count");
    for (int32_t b = 0; b < nb && r.ok; ++b) {
        o.bone_names.push_back(r.text());
        r.text();
        float pos[3];
        r.floats(pos, 3);
        o.bone_position.insert(o.bone_position.end(), pos, pos + 3);
        const int32_t parent = r.index(bsz);
        o.bone_parent.push_back((parent >= 0 && parent < nb) ? parent : -1);  // pmx_reader_impl.inl:198-203
        o.bone_transform_level.push_back(r.get<int32_t>());
        const uint16_t flag = r.get<uint16_t>();
        o.bone_flags.push_back(flag);
        if (flag & 0x0001) r.index(bsz); else r.skip(12);  // child: bone index or offset
        int32_t ap = -1;
        float ar = 0.0f;
        if (flag & (MMDGPU_BONE_APPEND_ROTATE | MMDGPU_BONE_APPEND_TRANSLATE)) { ap = r.index(bsz); ar = r.get<float>(); }
        o.bone_append_parent.push_back(ap);
        o.bone_append_ratio.push_back(ar);
        if (flag & 0x0400) r.skip(12);   // fixed rotation axis
        if (flag & 0x0800) r.skip(24);   // local axes
        if (flag & 0x2000) r.skip(4);    // external parent key
        int32_t target = -1, iters = 0;
        float angle = 0.0f;
        uint32_t lbegin = uint32_t(o.ik_link_bone.size()), lcount = 0;
        if (flag & MMDGPU_BONE_HAS_IK) {
            target = r.index(bsz);
            iters = r.get<int32_t>();
            angle = r.get<float>();
            const int32_t nl = r.get<int32_t>();
            if (nl < 0 || size_t(nl) > n) return fail(err, "bad IK link count at bone " + std::to_string(b));
            for (int32_t j = 0; j < nl && r.ok; ++j) {
                o.ik_link_bone.push_back(r.index(bsz));
                const uint8_t has = r.get<uint8_t>() != 0;
                o.ik_link_has_limit.push_back(has);
                float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
                if (has) { r.floats(lo, 3); r.floats(hi, 3); }
                o.ik_link_lo.insert(o.ik_link_lo.end(), lo, lo + 3);
                o.ik_link_hi.insert(o.ik_link_hi.end(), hi, hi + 3);
            }
            lcount = uint32_t(nl);
        }
        o.ik_target.push_back(target); o.ik_iterations.push_back(iters); o.ik_angle_limit.push_back(angle);
        o.ik_link_begin.push_back(lbegin); o.ik_link_count.push_back(lcount);
    }
    if (!r.ok) return fail(err, "truncated bone section");

    // ---- morphs, pmx_reader_impl.inl:267-360
    const int32_t nm = r.get<int32_t>();
    if (nm < 0 || size_t(nm) > (n - r.at) / 14 + 1) return fail(err, "bad morph count");
    for (int32_t m = 0; m < nm && r.ok; ++m) {
        o.morph_names.push_back(r.text());
        r.text();
        r.skip(1);  // panel / category
        const uint8_t type = r.get<uint8_t>();
        const int32_t cnt = r.get<int32_t>();
        if (cnt < 0 || size_t(cnt) > n) return fail(err, "bad morph entry count at morph " + std::to_string(m));
        uint32_t begin = 0, kept = uint32_t(cnt);
        switch (type) {
        case MMDGPU_MORPH_GROUP:
            begin = uint32_t(o.gme.size());
            for (int32_t j = 0; j < cnt && r.ok; ++j) {
                mmdgpu_group_morph_entry e;
                e.morph = uint32_t(r.index(mosz));
                e.rate = r.get<float>();
                o.gme.push_back(e);
            }
            break;
        case MMDGPU_MORPH_VERTEX:
            begin = uint32_t(o.vme.size());
            for (int32_t j = 0; j < cnt && r.ok; ++j) {
                mmdgpu_vertex_morph_entry e;
                e.vertex = r.vindex(vsz);
                r.floats(e.offset, 3);
                o.vme.push_back(e);
            }
            break;
        case MMDGPU_MORPH_BONE:
            begin = uint32_t(o.bme.size());
            for (int32_t j = 0; j < cnt && r.ok; ++j) {
                mmdgpu_bone_morph_entry e;
                e.bone = uint32_t(r.index(bsz));
                r.floats(e.translation, 3);
                r.floats(e.rotation, 4);
                o.bme.push_back(e);
            }
            break;
        case MMDGPU_MORPH_UV: case MMDGPU_MORPH_EXT_UV1: case MMDGPU_MORPH_EXT_UV2: case MMDGPU_MORPH_EXT_UV3:
        case MMDGPU_MORPH_EXT_UV4:
            begin = uint32_t(o.uvme.size());
            for (int32_t j = 0; j < cnt && r.ok; ++j) {
                mmdgpu_uv_morph_entry e;
                e.vertex = r.vindex(vsz);
                r.floats(e.offset, 4);
                o.uvme.push_back(e);
            }
            break;
        case MMDGPU_MORPH_MATERIAL:  // index + pmx_material_morph (113 B, interprete/pmx_types.inl:61-72)
            begin = uint32_t(o.mme.size());
            for (int32_t j = 0; j < cnt && r.ok; ++j) {
                mmdgpu_material_morph_entry e;
                const int64_t mi = r.index(msz);
                e.material = (mi < 0 || mi >= n_mat) ? -1 : int32_t(mi);
                e.method = r.get<uint8_t>();
                r.floats(e.value, MMDGPU_MATERIAL_FIELDS);
                if (e.method > MMDGPU_MATERIAL_ADD) return fail(err, "unknown material morph method at morph " + std::to_string(m));
                o.mme.push_back(e);
            }
            break;
        case 9:   // PMX 2.1 flip morph: morph index + rate
            r.skip(size_t(cnt) * (size_t(mosz) + 4));
            kept = 0;
            break;
        case 10:  // PMX 2.1 impulse morph: rigid body index + local flag + velocity + torque
            r.skip(size_t(cnt) * (size_t(rsz) + 25));
            kept = 0;
            break;
        default: return fail(err, "unknown morph type at morph " + std::to_string(m));
        }
        // types the deformation path ignores keep their number so that morph indices stay those of the file
        o.morph_type.push_back(type <= MMDGPU_MORPH_MATERIAL ? type : uint8_t(MMDGPU_MORPH_MATERIAL));
        o.morph_entry_begin.push_back(begin);
        o.morph_entry_count.push_back(kept);
    }
    if (!r.ok) return fail(err, "truncated morph section");
    // display frames, rigid bodies, joints (and PMX 2.1 soft bodies) are not part of the deformation path.
    o.finish();
    return MMDGPU_OK;
}

void ParsedMotion::finish() {
    desc = mmdgpu_anim_desc{};
    desc.n_bone_tracks = uint32_t(bone_track_bone.size());
    desc.bone_track_bone = bone_track_bone.data();
    desc.bone_track_key_begin = bt_begin.data(); desc.bone_track_key_count = bt_count.data();
    desc.n_bone_keys = uint32_t(bone_keys.size()); desc.bone_keys = bone_keys.data();
    desc.n_morph_tracks = uint32_t(morph_track_morph.size());
    desc.morph_track_morph = morph_track_morph.data();
    desc.morph_track_key_begin = mt_begin.data(); desc.morph_track_key_count = mt_count.data();
    desc.n_morph_keys = uint32_t(morph_keys.size()); desc.morph_keys = morph_keys.data();
}

mmdgpu_status parse_vmd(const void* bytes, size_t n, const Plan& plan, ParsedMotion& o, std::string& err) {
    o = ParsedMotion();
    Reader r(bytes, n);
    char magic[30];
    for (char& c : magic) c = char(r.get<uint8_t>());
    r.skip(20);  // model name
    if (!r.ok || std::strncmp(magic, "Vocaloid Motion Data 0002", 30) != 0) return fail(err, "not a VMD file");
    if (plan.bone_names.size() != plan.nb || plan.morph_names.size() != plan.nm)
        return fail(err, "the model carries no names (create it from PMX bytes to join a VMD by name)");

    // model names -> indices.  MotionPlayer's name join (L/motion/poser_impl.inl:522-537) looks every model
    // bone / morph up in the motion, so model items that share a name all follow the same track.
    const bool utf8 = plan.names_utf8;
    auto decode = [&](const std::string& s) { return utf8 ? from_utf8(s) : from_utf16le(s); };
    std::map<std::u16string, std::vector<int32_t>> bone_of, morph_of;
    for (uint32_t b = 0; b < plan.nb; ++b) bone_of[decode(plan.bone_names[b])].push_back(int32_t(b));
    for (uint32_t m = 0; m < plan.nm; ++m) morph_of[decode(plan.morph_names[m])].push_back(int32_t(m));

    // ---- bone records: 15-byte name, u32 frame, T, R, 4 x 16 interpolation bytes (vmd_types.inl:22-31)
    const uint32_t n_bone = r.get<uint32_t>();
    if (!r.ok || size_t(n_bone) > (n - r.at) / 111) return fail(err, "bone record count exceeds the file size");
    std::map<int32_t, std::vector<mmdgpu_bone_key>> by_bone;
    for (uint32_t i = 0; i < n_bone; ++i) {
        char name[15];
        std::memcpy(name, r.p + r.at, 15);
        r.skip(15);
        mmdgpu_bone_key k;
        k.frame = r.get<uint32_t>();
        r.floats(k.translation, 3);
        r.floats(k.rotation, 4);
        for (int c = 0; c < 4; ++c) {
            const int8_t* blk = reinterpret_cast<const int8_t*>(r.p + r.at);
            // control points are bytes [0], [4], [8], [12] of the channel's block (vmd_reader_impl.inl:32-61)
            k.interp[c][0] = blk[0]; k.interp[c][1] = blk[4]; k.interp[c][2] = blk[8]; k.interp[c][3] = blk[12];
            r.skip(16);
        }
        auto it = bone_of.find(from_cp932(name, 15));
        if (it != bone_of.end())
            for (int32_t b : it->second) by_bone[b].push_back(k);
    }
    const uint32_t n_morph = r.get<uint32_t>();
    if (!r.ok || size_t(n_morph) > (n - r.at) / 23) return fail(err, "morph record count exceeds the file size");
    std::map<int32_t, std::vector<mmdgpu_morph_key>> by_morph;
    for (uint32_t i = 0; i < n_morph; ++i) {
        char name[15];
        std::memcpy(name, r.p + r.at, 15);
        r.skip(15);
        mmdgpu_morph_key k;
        k.frame = r.get<uint32_t>();
        k.weight = r.get<float>();
        auto it = morph_of.find(from_cp932(name, 15));
        if (it != morph_of.end())
            for (int32_t m : it->second) by_morph[m].push_back(k);
    }
    if (!r.ok) return fail(err, "truncated VMD");
    for (auto& kv : by_bone) {
        o.bone_track_bone.push_back(kv.first);
        o.bt_begin.push_back(uint32_t(o.bone_keys.size()));
        o.bt_count.push_back(uint32_t(kv.second.size()));
        o.bone_keys.insert(o.bone_keys.end(), kv.second.begin(), kv.second.end());
    }
    for (auto& kv : by_morph) {
        o.morph_track_morph.push_back(kv.first);
        o.mt_begin.push_back(uint32_t(o.morph_keys.size()));
        o.mt_count.push_back(uint32_t(kv.second.size()));
        o.morph_keys.insert(o.morph_keys.end(), kv.second.begin(), kv.second.end());
    }
    o.finish();
    return MMDGPU_OK;
}

}  // namespace mmdgpu
