// mmdgpu_libmmd_adapter.hpp — OPTIONAL: flatten an already loaded mmd::Model / mmd::Motion into the descriptor
// structs of mmdgpu.h.  Not part of the core library: it compiles only where libmmd's headers are on
// the include path, and the including translation unit must have included them first, the way main.cpp does:
//
//     #include <math.h>
//     #include <stdlib.h>          // float abs() overloads before libmmd
//     #include <mmd/mmd.hxx>
//     #include <mmdgpu_libmmd_adapter.hpp>
//
//     mmdgpu::FlatModel fm;  mmdgpu::Flatten(*g_state.model, fm);             // after PmxReader::ReadModel
//     mmdgpu::Model gpu_model(ctx, fm.desc);
//     mmdgpu::FlatMotion fa; mmdgpu::Flatten(*g_state.motion, *g_state.model, fa);   // after VmdReader::ReadMotion
//     mmdgpu::Motion gpu_motion(gpu_model, fa.desc);
//
// Only libmmd's public accessors are used (L/model/model.inl, L/motion/motion.inl).  The model is taken as libmmd
// holds it, i.e. after Model::Normalize (pmx_reader_impl.inl:441); the rewrite is idempotent, so the plan built from
// this descriptor equals the plan built from the PMX bytes.  Tracks are joined to bones / morphs by exact name, as
// MotionPlayer does (poser_impl.inl:522-537).
#ifndef MMDGPU_LIBMMD_ADAPTER_HPP_INCLUDED
#define MMDGPU_LIBMMD_ADAPTER_HPP_INCLUDED

#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "mmdgpu.h"

namespace mmdgpu {

struct FlatModel {
    std::vector<float> position, normal, uv, weight, sdef_c, sdef_r0, sdef_r1, bone_position, bone_append_ratio, ik_angle_limit,
        ik_link_lo, ik_link_hi;
    std::vector<uint8_t> skin_type, ik_link_has_limit, morph_type;
    std::vector<int32_t> bone_id, bone_parent, bone_transform_level, bone_append_parent, ik_target, ik_iterations, ik_link_bone;
    std::vector<uint16_t> bone_flags;
    std::vector<uint32_t> ik_link_begin, ik_link_count, morph_entry_begin, morph_entry_count;
    std::vector<mmdgpu_vertex_morph_entry> vertex_morph_entries;
    std::vector<mmdgpu_uv_morph_entry> uv_morph_entries;
    std::vector<mmdgpu_bone_morph_entry> bone_morph_entries;
    std::vector<mmdgpu_group_morph_entry> group_morph_entries;
    std::vector<mmdgpu_material_morph_entry> material_morph_entries;
    mmdgpu_model_desc desc;
};

struct FlatMotion {
    std::vector<int32_t> bone_track_bone, morph_track_morph;
    std::vector<uint32_t> bone_track_key_begin, bone_track_key_count, morph_track_key_begin, morph_track_key_count;
    std::vector<mmdgpu_bone_key> bone_keys;
    std::vector<mmdgpu_morph_key> morph_keys;
    mmdgpu_anim_desc desc;
};

namespace adapter_detail {
inline int32_t index_or_none(size_t i) { return (i == mmd::nil || i > size_t(0x7FFFFFFF)) ? -1 : int32_t(i); }
template <class V3> inline void push3(std::vector<float>& dst, const V3& v) {
    dst.push_back(v.p.x); dst.push_back(v.p.y); dst.push_back(v.p.z);
}
// Bezier::GetC returns a reference to a temporary (L/util/math_impl.inl:1385-1392), so the control points cannot be
// read through the public interface.  The members are reached with the explicit-instantiation idiom instead (an
// explicit instantiation may name private members); they hold 3 x the control point (SetC, math_impl.inl:1393-1397).
template <class Tag> struct member_of { typedef typename Tag::type type; };
template <class Tag, typename Tag::type Member> struct expose {
    friend typename Tag::type reveal(Tag) { return Member; }
};
struct bezier_c0 { typedef mmd::Vector2D<float> mmd::Bezier<float, 32>::*type; friend type reveal(bezier_c0); };
struct bezier_c1 { typedef mmd::Vector2D<float> mmd::Bezier<float, 32>::*type; friend type reveal(bezier_c1); };
template struct expose<bezier_c0, &mmd::Bezier<float, 32>::c_0>;
template struct expose<bezier_c1, &mmd::Bezier<float, 32>::c_1>;

// One VMD control byte from the float libmmd stored: 3 * (byte * (1 / 127.f)) (vmd_reader_impl.inl:30-61, SetC).
inline int8_t control_byte(float c3) {
    const long b = std::lround(double(c3) / 3.0 * 127.0);
    if (b < -128 || b > 127 || (float(int(b)) * (1.0f / 127.0f)) * 3.0f != c3)
        throw std::runtime_error("mmdgpu adapter: Bezier control point is not a VMD byte / 127");
    return int8_t(b);
}
inline void control_quad(const mmd::Bezier<float, 32>& ip, int8_t out[4]) {
    const mmd::Vector2D<float>& c0 = ip.*reveal(bezier_c0());
    const mmd::Vector2D<float>& c1 = ip.*reveal(bezier_c1());
    if (c0.p.x == c0.p.y && c1.p.x == c1.p.y) {  // linear (math_impl.inl:1399), including default-constructed interpolators
        out[0] = 20; out[1] = 20; out[2] = 107; out[3] = 107;
        return;
    }
    out[0] = control_byte(c0.p.x); out[1] = control_byte(c0.p.y); out[2] = control_byte(c1.p.x); out[3] = control_byte(c1.p.y);
}
}  // namespace adapter_detail

inline void Flatten(mmd::Model& model, FlatModel& o) {
    using namespace adapter_detail;
    o = FlatModel();
    const size_t nv = model.GetVertexNum(), nb = model.GetBoneNum(), nm = model.GetMorphNum();
    bool any_sdef = false;
    for (size_t i = 0; i < nv; ++i) {
        mmd::Model::Vertex<mmd::ref> v = model.GetVertex(i);
        push3(o.position, v.GetCoordinate());
        push3(o.normal, v.GetNormal());
        o.uv.push_back(v.GetUVCoordinate().p.x);
        o.uv.push_back(v.GetUVCoordinate().p.y);
        const mmd::Model::SkinningOperator& op = v.GetSkinningOperator();
        int32_t id[4] = {0, 0, 0, 0};
        float w[4] = {0.f, 0.f, 0.f, 0.f}, c[3] = {0, 0, 0}, r0[3] = {0, 0, 0}, r1[3] = {0, 0, 0};
        uint8_t t = MMDGPU_SKIN_BDEF1;
        switch (op.GetSkinningType()) {
        case mmd::Model::SkinningOperator::SKINNING_BDEF1:
            id[0] = int32_t(op.GetBDEF1().GetBoneID());
            w[0] = 1.0f;
            break;
        case mmd::Model::SkinningOperator::SKINNING_BDEF2:
            t = MMDGPU_SKIN_BDEF2;
            id[0] = int32_t(op.GetBDEF2().GetBoneID(0));
            id[1] = int32_t(op.GetBDEF2().GetBoneID(1));
            w[0] = op.GetBDEF2().GetBoneWeight();
            break;
        case mmd::Model::SkinningOperator::SKINNING_BDEF4:
            t = MMDGPU_SKIN_BDEF4;
            for (size_t k = 0; k < 4; ++k) {
                id[k] = int32_t(op.GetBDEF4().GetBoneID(k));
                w[k] = op.GetBDEF4().GetBoneWeight(k);
            }
            break;
        default:  // SDEF, and anything Deform's default: label treats as one (poser_impl.inl:417-426)
            t = MMDGPU_SKIN_SDEF;
            any_sdef = true;
            id[0] = int32_t(op.GetSDEF().GetBoneID(0));
            id[1] = int32_t(op.GetSDEF().GetBoneID(1));
            w[0] = op.GetSDEF().GetBoneWeight();
            c[0] = op.GetSDEF().GetC().p.x; c[1] = op.GetSDEF().GetC().p.y; c[2] = op.GetSDEF().GetC().p.z;
            r0[0] = op.GetSDEF().GetR0().p.x; r0[1] = op.GetSDEF().GetR0().p.y; r0[2] = op.GetSDEF().GetR0().p.z;
            r1[0] = op.GetSDEF().GetR1().p.x; r1[1] = op.GetSDEF().GetR1().p.y; r1[2] = op.GetSDEF().GetR1().p.z;
            break;
        }
        o.skin_type.push_back(t);
        for (int k = 0; k < 4; ++k) { o.bone_id.push_back(id[k]); o.weight.push_back(w[k]); }
        for (int k = 0; k < 3; ++k) { o.sdef_c.push_back(c[k]); o.sdef_r0.push_back(r0[k]); o.sdef_r1.push_back(r1[k]); }
    }
    for (size_t b = 0; b < nb; ++b) {
        const mmd::Model::Bone& bone = model.GetBone(b);
        push3(o.bone_position, bone.GetPosition());
        o.bone_parent.push_back(index_or_none(bone.GetParentIndex()));
        o.bone_transform_level.push_back(int32_t(bone.GetTransformLevel()));
        o.bone_flags.push_back(uint16_t((bone.IsHasIK() ? MMDGPU_BONE_HAS_IK : 0) | (bone.IsAppendRotate() ? MMDGPU_BONE_APPEND_ROTATE : 0) |
                                        (bone.IsAppendTranslate() ? MMDGPU_BONE_APPEND_TRANSLATE : 0) |
                                        (bone.IsPostPhysics() ? MMDGPU_BONE_POST_PHYSICS : 0)));
        const bool appends = bone.IsAppendRotate() || bone.IsAppendTranslate();
        o.bone_append_parent.push_back(appends ? index_or_none(bone.GetAppendIndex()) : -1);
        o.bone_append_ratio.push_back(appends ? bone.GetAppendRatio() : 0.0f);
        o.ik_link_begin.push_back(uint32_t(o.ik_link_bone.size()));
        if (bone.IsHasIK()) {
            o.ik_target.push_back(index_or_none(bone.GetIKTargetIndex()));
            o.ik_iterations.push_back(int32_t(bone.GetCCDIterateLimit()));
            o.ik_angle_limit.push_back(bone.GetCCDAngleLimit());
            for (size_t l = 0; l < bone.GetIKLinkNum(); ++l) {
                const mmd::Model::Bone::IKLink& link = bone.GetIKLink(l);
                o.ik_link_bone.push_back(index_or_none(link.GetLinkIndex()));
                o.ik_link_has_limit.push_back(link.IsHasLimit() ? 1 : 0);
                if (link.IsHasLimit()) {
                    push3(o.ik_link_lo, link.GetLoLimit());
                    push3(o.ik_link_hi, link.GetHiLimit());
                } else {
                    for (int k = 0; k < 3; ++k) { o.ik_link_lo.push_back(0.f); o.ik_link_hi.push_back(0.f); }
                }
            }
            o.ik_link_count.push_back(uint32_t(bone.GetIKLinkNum()));
        } else {
            o.ik_target.push_back(-1);
            o.ik_iterations.push_back(0);
            o.ik_angle_limit.push_back(0.0f);
            o.ik_link_count.push_back(0);
        }
    }
    size_t n_materials = 0;
    for (size_t p = 0; p < model.GetPartNum(); ++p) ++n_materials;  // one material per part (model.inl:169-170)
    for (size_t m = 0; m < nm; ++m) {
        const mmd::Model::Morph& morph = model.GetMorph(m);
        const int type = int(morph.GetType());
        const size_t n = morph.GetMorphDataNum();
        uint32_t begin = 0, count = uint32_t(n);
        if (type == MMDGPU_MORPH_GROUP) {
            begin = uint32_t(o.group_morph_entries.size());
            for (size_t j = 0; j < n; ++j) {
                mmdgpu_group_morph_entry e;
                e.morph = uint32_t(morph.GetMorphData(j).GetGroupMorph().GetMorphIndex());
                e.rate = morph.GetMorphData(j).GetGroupMorph().GetMorphRate();
                o.group_morph_entries.push_back(e);
            }
        } else if (type == MMDGPU_MORPH_VERTEX) {
            begin = uint32_t(o.vertex_morph_entries.size());
            for (size_t j = 0; j < n; ++j) {
                const mmd::Model::Morph::MorphData::VertexMorph& s = morph.GetMorphData(j).GetVertexMorph();
                mmdgpu_vertex_morph_entry e;
                e.vertex = uint32_t(s.GetVertexIndex());
                e.offset[0] = s.GetOffset().p.x; e.offset[1] = s.GetOffset().p.y; e.offset[2] = s.GetOffset().p.z;
                o.vertex_morph_entries.push_back(e);
            }
        } else if (type == MMDGPU_MORPH_BONE) {
            begin = uint32_t(o.bone_morph_entries.size());
            for (size_t j = 0; j < n; ++j) {
                const mmd::Model::Morph::MorphData::BoneMorph& s = morph.GetMorphData(j).GetBoneMorph();
                mmdgpu_bone_morph_entry e;
                e.bone = uint32_t(s.GetBoneIndex());
                e.translation[0] = s.GetTranslation().p.x; e.translation[1] = s.GetTranslation().p.y; e.translation[2] = s.GetTranslation().p.z;
                for (int k = 0; k < 4; ++k) e.rotation[k] = s.GetRotation().v[k];
                o.bone_morph_entries.push_back(e);
            }
        } else if (type >= MMDGPU_MORPH_UV && type <= MMDGPU_MORPH_EXT_UV4) {
            begin = uint32_t(o.uv_morph_entries.size());
            for (size_t j = 0; j < n; ++j) {
                const mmd::Model::Morph::MorphData::UVMorph& s = morph.GetMorphData(j).GetUVMorph();
                mmdgpu_uv_morph_entry e;
                e.vertex = uint32_t(s.GetVertexIndex());
                for (int k = 0; k < 4; ++k) e.offset[k] = s.GetOffset().v[k];
                o.uv_morph_entries.push_back(e);
            }
        } else if (type == MMDGPU_MORPH_MATERIAL) {
            begin = uint32_t(o.material_morph_entries.size());
            for (size_t j = 0; j < n; ++j) {
                const mmd::Model::Morph::MorphData::MaterialMorph& s = morph.GetMorphData(j).GetMaterialMorph();
                mmdgpu_material_morph_entry e;
                e.material = (s.IsGlobal() || s.GetMaterialIndex() >= n_materials) ? -1 : int32_t(s.GetMaterialIndex());
                e.method = uint32_t(s.GetMethod());
                float* v = e.value;
                for (int k = 0; k < 4; ++k) v[k] = s.GetDiffuse().v[k];
                for (int k = 0; k < 3; ++k) v[4 + k] = s.GetSpecular().v[k];
                v[7] = s.GetShininess();
                for (int k = 0; k < 3; ++k) v[8 + k] = s.GetAmbient().v[k];
                for (int k = 0; k < 4; ++k) v[11 + k] = s.GetEdgeColor().v[k];
                v[15] = s.GetEdgeSize();
                for (int k = 0; k < 4; ++k) v[16 + k] = s.GetTexture().v[k];
                for (int k = 0; k < 4; ++k) v[20 + k] = s.GetSubTexture().v[k];
                for (int k = 0; k < 4; ++k) v[24 + k] = s.GetToonTexture().v[k];
                o.material_morph_entries.push_back(e);
            }
        } else {
            count = 0;
        }
        o.morph_type.push_back(uint8_t(type <= MMDGPU_MORPH_MATERIAL ? type : MMDGPU_MORPH_MATERIAL));
        o.morph_entry_begin.push_back(begin);
        o.morph_entry_count.push_back(count);
    }

    mmdgpu_model_desc& d = o.desc;
    d = mmdgpu_model_desc();
    d.n_vertices = uint32_t(nv);
    d.position = o.position.data(); d.normal = o.normal.data(); d.uv = o.uv.data();
    d.skin_type = o.skin_type.data(); d.bone_id = o.bone_id.data(); d.weight = o.weight.data();
    if (any_sdef) { d.sdef_c = o.sdef_c.data(); d.sdef_r0 = o.sdef_r0.data(); d.sdef_r1 = o.sdef_r1.data(); }
    d.n_bones = uint32_t(nb);
    d.bone_position = o.bone_position.data(); d.bone_parent = o.bone_parent.data();
    d.bone_transform_level = o.bone_transform_level.data(); d.bone_flags = o.bone_flags.data();
    d.bone_append_parent = o.bone_append_parent.data(); d.bone_append_ratio = o.bone_append_ratio.data();
    d.ik_target = o.ik_target.data(); d.ik_iterations = o.ik_iterations.data(); d.ik_angle_limit = o.ik_angle_limit.data();
    d.ik_link_begin = o.ik_link_begin.data(); d.ik_link_count = o.ik_link_count.data();
    d.n_ik_links = uint32_t(o.ik_link_bone.size());
    d.ik_link_bone = o.ik_link_bone.data(); d.ik_link_has_limit = o.ik_link_has_limit.data();
    d.ik_link_lo = o.ik_link_lo.data(); d.ik_link_hi = o.ik_link_hi.data();
    d.n_morphs = uint32_t(nm);
    d.morph_type = o.morph_type.data(); d.morph_entry_begin = o.morph_entry_begin.data(); d.morph_entry_count = o.morph_entry_count.data();
    d.n_vertex_morph_entries = uint32_t(o.vertex_morph_entries.size()); d.vertex_morph_entries = o.vertex_morph_entries.data();
    d.n_uv_morph_entries = uint32_t(o.uv_morph_entries.size()); d.uv_morph_entries = o.uv_morph_entries.data();
    d.n_bone_morph_entries = uint32_t(o.bone_morph_entries.size()); d.bone_morph_entries = o.bone_morph_entries.data();
    d.n_group_morph_entries = uint32_t(o.group_morph_entries.size()); d.group_morph_entries = o.group_morph_entries.data();
    d.n_materials = uint32_t(n_materials);
    d.n_material_morph_entries = uint32_t(o.material_morph_entries.size()); d.material_morph_entries = o.material_morph_entries.data();
}

// Tracks of `motion` that name a bone / morph of `model` (the rest are ignored, as MotionPlayer ignores them).
// A name that several bones share animates all of them, as in poser_impl.inl:522-537.
// `track_prefix`: under glibc libmmd's Shift-JIS conversion leaves a byte-order mark in front of every VMD name
// (L/util/dwarf_impl.inl:205-232), so that no track ever matches a bone; pass L"\xfeff" there to join anyway
// (VmdTrackPrefix() detects it from the motion's own name).
inline std::wstring VmdTrackPrefix(const mmd::Motion& motion) {
    const std::wstring& n = motion.GetName();
    return (!n.empty() && n[0] == wchar_t(0xFEFF)) ? std::wstring(1, wchar_t(0xFEFF)) : std::wstring();
}
inline void Flatten(const mmd::Motion& motion, mmd::Model& model, FlatMotion& o, const std::wstring& track_prefix = std::wstring()) {
    using namespace adapter_detail;
    o = FlatMotion();
    for (size_t b = 0; b < model.GetBoneNum(); ++b) {
        const std::wstring name = track_prefix + model.GetBone(b).GetName();
        if (!motion.IsBoneRegistered(name)) continue;
        o.bone_track_bone.push_back(int32_t(b));
        o.bone_track_key_begin.push_back(uint32_t(o.bone_keys.size()));
        uint32_t n = 0;
        for (size_t f = motion.QueryBoneKeyframeForward(name, 0); f != mmd::nil; f = motion.QueryBoneKeyframeForward(name, f + 1)) {
            const mmd::Motion::BoneKeyframe& kf = motion.GetBoneKeyframe(name, f);
            mmdgpu_bone_key k;
            k.frame = uint32_t(f);
            k.translation[0] = kf.GetTranslation().p.x; k.translation[1] = kf.GetTranslation().p.y; k.translation[2] = kf.GetTranslation().p.z;
            for (int c = 0; c < 4; ++c) k.rotation[c] = kf.GetRotation().v[c];
            control_quad(kf.GetXInterpolator(), k.interp[0]);
            control_quad(kf.GetYInterpolator(), k.interp[1]);
            control_quad(kf.GetZInterpolator(), k.interp[2]);
            control_quad(kf.GetRInterpolator(), k.interp[3]);
            o.bone_keys.push_back(k);
            ++n;
        }
        o.bone_track_key_count.push_back(n);
    }
    for (size_t m = 0; m < model.GetMorphNum(); ++m) {
        const std::wstring name = track_prefix + model.GetMorph(m).GetName();
        if (!motion.IsMorphRegistered(name)) continue;
        o.morph_track_morph.push_back(int32_t(m));
        o.morph_track_key_begin.push_back(uint32_t(o.morph_keys.size()));
        uint32_t n = 0;
        for (size_t f = motion.QueryMorphKeyframeForward(name, 0); f != mmd::nil; f = motion.QueryMorphKeyframeForward(name, f + 1)) {
            mmdgpu_morph_key k;
            k.frame = uint32_t(f);
            k.weight = motion.GetMorphKeyframe(name, f).GetWeight();
            o.morph_keys.push_back(k);
            ++n;
        }
        o.morph_track_key_count.push_back(n);
    }
    mmdgpu_anim_desc& d = o.desc;
    d = mmdgpu_anim_desc();
    d.n_bone_tracks = uint32_t(o.bone_track_bone.size());
    d.bone_track_bone = o.bone_track_bone.data();
    d.bone_track_key_begin = o.bone_track_key_begin.data(); d.bone_track_key_count = o.bone_track_key_count.data();
    d.n_bone_keys = uint32_t(o.bone_keys.size()); d.bone_keys = o.bone_keys.data();
    d.n_morph_tracks = uint32_t(o.morph_track_morph.size());
    d.morph_track_morph = o.morph_track_morph.data();
    d.morph_track_key_begin = o.morph_track_key_begin.data(); d.morph_track_key_count = o.morph_track_key_count.data();
    d.n_morph_keys = uint32_t(o.morph_keys.size()); d.morph_keys = o.morph_keys.data();
}

}  // namespace mmdgpu

#endif  // MMDGPU_LIBMMD_ADAPTER_HPP_INCLUDED
