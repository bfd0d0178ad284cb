// TEST INFRASTRUCTURE — not product code.  Nothing under simple_mmd_renderer_b200/ links, imports or
// executes this file; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs do.
//
// ref_harness.cc — the reference itself (libmmd, header-only) behind a flat C interface.
//
// Compiled by oracle/Makefile from the headers where they lie under
// /root/reference/3rd_party/libmmd/include into oracle/_ref/libmmd_ref.so (git-ignored).  No libmmd
// source is copied into this repository.  The harness builds mmd::Model / mmd::Motion
// programmatically from the same flat descriptors the CUDA library takes (include/mmdgpu.h), because
// libmmd's own file readers cannot join names on glibc, and drives the exact call
// sequence of main.cpp:1788-1821 with physics off:
//     ResetPosing(); SeekFrame(f); PrePhysicsPosing(); PostPhysicsPosing(); Deform();
//
// <math.h> and <stdlib.h> come first so that the unqualified abs(float) inside namespace mmd resolves
// to the float overload exactly as it does in main.cpp's translation unit.
#include <math.h>
#include <stdlib.h>

#include <mmd/mmd.hxx>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cwchar>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../include/mmdgpu.h"

using namespace mmd;

namespace {

Vector3f V3(const float* p) {
    Vector3f v;
    v.p.x = p[0];
    v.p.y = p[1];
    v.p.z = p[2];
    return v;
}

std::wstring Name(wchar_t prefix, unsigned i) {
    wchar_t buf[32];
    swprintf(buf, 32, L"%lc%u", prefix, i);
    return buf;
}

// BoneImage is private to Poser; PhysicsReactor is its friend and hands out references
// (L/motion/physics.inl:32-40).  This subclass exists only to reach that accessor.
class Spy : public PhysicsReactor {
public:
    void AddPoser(Poser&) override {}
    void RemovePoser(Poser&) override {}
    void Reset() override {}
    void React(float) override {}
    void SetGravityStrength(float) override {}
    void SetGravityDirection(const Vector3f&) override {}
    float GetGravityStrength() const override { return 0; }
    Vector3f GetGravityDirection() const override { return Vector3f(); }
    void SetFloor(bool) override {}
    bool IsHasFloor() const override { return false; }

    static void SkinningMatrix(Poser& p, size_t b, float* out16) {
        BoneImageReference im = GetPoserBoneImage(p, b);
        for (int k = 0; k < 16; ++k) out16[k] = im.skinning_matrix_.v[k];
    }
    static void LocalMatrix(Poser& p, size_t b, float* out16) {
        BoneImageReference im = GetPoserBoneImage(p, b);
        for (int k = 0; k < 16; ++k) out16[k] = im.local_matrix_.v[k];
    }
    // what a host physics reactor does between Pre and Post (PoserMotionState::Synchronize / Fix,
    // L/../mmd-bullet/mmd-bullet_impl.inl:34-56): overwrite a bone's skinning and local matrices in place
    static void Override(Poser& p, size_t b, const float* skin16, const float* local16) {
        BoneImageReference im = GetPoserBoneImage(p, b);
        for (int k = 0; k < 16; ++k) im.skinning_matrix_.v[k] = skin16[k];
        if (local16)
            for (int k = 0; k < 16; ++k) im.local_matrix_.v[k] = local16[k];
    }
    static void Pose(Poser& p, size_t b, float* out7) {
        BoneImageReference im = GetPoserBoneImage(p, b);
        for (int k = 0; k < 3; ++k) out7[k] = im.translation_.v[k];
        for (int k = 0; k < 4; ++k) out7[3 + k] = im.rotation_.v[k];
    }
    static void IkClass(Poser& p, size_t b, std::vector<uint8_t>& fix, std::vector<uint8_t>& order) {
        BoneImageReference im = GetPoserBoneImage(p, b);
        for (size_t j = 0; j < im.ik_links_.size(); ++j) {
            // FIX_NONE, FIX_X, FIX_Y, FIX_Z, FIX_ALL  ->  0..4 ; ORDER_ZXY, ORDER_XYZ, ORDER_YZX -> 1, 2, 0
            fix.push_back((uint8_t)im.ik_fix_types_[j]);
            static const uint8_t map[3] = {1, 2, 0};
            order.push_back(map[(int)im.ik_transform_orders_[j]]);
        }
    }
};

void BuildModel(Model& model, const mmdgpu_model_desc& d) {
    model.SetExtraUVNumber(0);
    for (uint32_t b = 0; b < d.n_bones; ++b) {
        Model::Bone& bone = model.NewBone();
        std::wstring nm = Name(L'b', b);
        bone.SetName(nm);
        bone.SetNameEn(nm);
        bone.SetPosition(V3(d.bone_position + 3 * b));
        int32_t parent = d.bone_parent[b];
        bone.SetParentIndex(parent < 0 ? nil : (size_t)parent);
        bone.SetTransformLevel((size_t)d.bone_transform_level[b]);
        uint16_t fl = d.bone_flags[b];
        bone.SetChildUseID(false);
        bone.SetRotatable(true);
        bone.SetMovable(true);
        bone.SetVisible(true);
        bone.SetControllable(true);
        bone.SetHasIK((fl & MMDGPU_BONE_HAS_IK) != 0);
        bone.SetAppendRotate((fl & MMDGPU_BONE_APPEND_ROTATE) != 0);
        bone.SetAppendTranslate((fl & MMDGPU_BONE_APPEND_TRANSLATE) != 0);
        bone.SetRotAxisFixed(false);
        bone.SetUseLocalAxis(false);
        bone.SetPostPhysics((fl & MMDGPU_BONE_POST_PHYSICS) != 0);
        bone.SetReceiveTransform(false);
        float up[3] = {0, 1, 0};
        bone.SetChildOffset(V3(up));
        int32_t ap = d.bone_append_parent ? d.bone_append_parent[b] : -1;
        bone.SetAppendIndex(ap < 0 ? nil : (size_t)ap);
        bone.SetAppendRatio(d.bone_append_ratio ? d.bone_append_ratio[b] : 0.0f);
        bone.SetExportKey(0);
        bone.SetIKTargetIndex(nil);
        bone.SetCCDIterateLimit(0);
        bone.SetCCDAngleLimit(0);
        if (fl & MMDGPU_BONE_HAS_IK) {
            int32_t tgt = d.ik_target[b];
            bone.SetIKTargetIndex(tgt < 0 ? nil : (size_t)tgt);
            bone.SetCCDIterateLimit((size_t)d.ik_iterations[b]);
            bone.SetCCDAngleLimit(d.ik_angle_limit[b]);
            for (uint32_t j = 0; j < d.ik_link_count[b]; ++j) {
                uint32_t l = d.ik_link_begin[b] + j;
                Model::Bone::IKLink& link = bone.NewIKLink();
                link.SetLinkIndex((size_t)d.ik_link_bone[l]);
                link.SetHasLimit(d.ik_link_has_limit[l] != 0);
                link.SetLoLimit(V3(d.ik_link_lo + 3 * l));
                link.SetHiLimit(V3(d.ik_link_hi + 3 * l));
            }
        }
    }
    const float zero3[3] = {0, 0, 0};
    for (uint32_t i = 0; i < d.n_vertices; ++i) {
        Model::Vertex<ref> v = model.NewVertex();
        v.SetCoordinate(V3(d.position + 3 * i));
        v.SetNormal(V3(d.normal + 3 * i));
        Vector2f uv;
        uv.v[0] = d.uv ? d.uv[2 * i] : 0.0f;
        uv.v[1] = d.uv ? d.uv[2 * i + 1] : 0.0f;
        v.SetUVCoordinate(uv);
        v.SetEdgeScale(1.0f);
        Model::SkinningOperator& op = v.GetSkinningOperator();
        const int32_t* id = d.bone_id + 4 * i;
        const float* w = d.weight + 4 * i;
        switch (d.skin_type[i]) {
        case MMDGPU_SKIN_BDEF1:
            op.SetSkinningType(Model::SkinningOperator::SKINNING_BDEF1);
            op.GetBDEF1().SetBoneID((size_t)id[0]);
            break;
        case MMDGPU_SKIN_BDEF2:
            op.SetSkinningType(Model::SkinningOperator::SKINNING_BDEF2);
            op.GetBDEF2().SetBoneID(0, (size_t)id[0]);
            op.GetBDEF2().SetBoneID(1, (size_t)id[1]);
            op.GetBDEF2().SetBoneWeight(w[0]);
            break;
        case MMDGPU_SKIN_SDEF:
            op.SetSkinningType(Model::SkinningOperator::SKINNING_SDEF);
            op.GetSDEF().SetBoneID(0, (size_t)id[0]);
            op.GetSDEF().SetBoneID(1, (size_t)id[1]);
            op.GetSDEF().SetBoneWeight(w[0]);
            op.GetSDEF().SetC(V3(d.sdef_c ? d.sdef_c + 3 * i : zero3));
            op.GetSDEF().SetR0(V3(d.sdef_r0 ? d.sdef_r0 + 3 * i : zero3));
            op.GetSDEF().SetR1(V3(d.sdef_r1 ? d.sdef_r1 + 3 * i : zero3));
            break;
        case MMDGPU_SKIN_BDEF4:
        case MMDGPU_SKIN_QDEF:  // libmmd has no QDEF (L/model/model.inl:23-28): the oracle declares BDEF4
        default:
            op.SetSkinningType(Model::SkinningOperator::SKINNING_BDEF4);
            for (int j = 0; j < 4; ++j) {
                op.GetBDEF4().SetBoneID(j, (size_t)id[j]);
                op.GetBDEF4().SetBoneWeight(j, w[j]);
            }
            break;
        }
    }
    for (uint32_t m = 0; m < d.n_morphs; ++m) {
        Model::Morph& mo = model.NewMorph();
        std::wstring nm = Name(L'm', m);
        mo.SetName(nm);
        mo.SetNameEn(nm);
        mo.SetCategory(Model::Morph::MORPH_CAT_OTHER);
        mo.SetType((Model::Morph::MorphType)d.morph_type[m]);
        uint32_t b = d.morph_entry_begin[m], n = d.morph_entry_count[m];
        for (uint32_t j = 0; j < n; ++j) {
            Model::Morph::MorphData& md = mo.NewMorphData();
            switch (d.morph_type[m]) {
            case MMDGPU_MORPH_GROUP:
                md.GetGroupMorph().SetMorphIndex(d.group_morph_entries[b + j].morph);
                md.GetGroupMorph().SetMorphRate(d.group_morph_entries[b + j].rate);
                break;
            case MMDGPU_MORPH_VERTEX:
                md.GetVertexMorph().SetVertexIndex(d.vertex_morph_entries[b + j].vertex);
                md.GetVertexMorph().SetOffset(V3(d.vertex_morph_entries[b + j].offset));
                break;
            case MMDGPU_MORPH_BONE: {
                md.GetBoneMorph().SetBoneIndex(d.bone_morph_entries[b + j].bone);
                md.GetBoneMorph().SetTranslation(V3(d.bone_morph_entries[b + j].translation));
                Vector4f r;
                for (int k = 0; k < 4; ++k) r.v[k] = d.bone_morph_entries[b + j].rotation[k];
                md.GetBoneMorph().SetRotation(r);
                break;
            }
            case MMDGPU_MORPH_UV:
            case MMDGPU_MORPH_EXT_UV1:
            case MMDGPU_MORPH_EXT_UV2:
            case MMDGPU_MORPH_EXT_UV3:
            case MMDGPU_MORPH_EXT_UV4: {
                md.GetUVMorph().SetVertexIndex(d.uv_morph_entries[b + j].vertex);
                Vector4f o;
                for (int k = 0; k < 4; ++k) o.v[k] = d.uv_morph_entries[b + j].offset[k];
                md.GetUVMorph().SetOffset(o);
                break;
            }
            default:
                break;
            }
        }
    }
    // PmxReader::ReadModel ends with this (L/reader/pmx_reader_impl.inl:441).
    model.Normalize();
}

void BuildMotion(Motion& motion, const mmdgpu_anim_desc& a) {
    const float r = 1.0f / 127.0f;  // L/reader/vmd_reader_impl.inl:30
    for (uint32_t t = 0; t < a.n_bone_tracks; ++t) {
        std::wstring nm = Name(L'b', (unsigned)a.bone_track_bone[t]);
        motion.RegisterBone(nm);
        for (uint32_t j = 0; j < a.bone_track_key_count[t]; ++j) {
            const mmdgpu_bone_key& k = a.bone_keys[a.bone_track_key_begin[t] + j];
            Motion::BoneKeyframe& kf = motion.GetBoneKeyframe(nm, k.frame);
            kf.SetTranslation(V3(k.translation));
            Vector4f q;
            for (int c = 0; c < 4; ++c) q.v[c] = k.rotation[c];
            kf.SetRotation(q);
            interpolator* ip[4] = {&kf.GetXInterpolator(), &kf.GetYInterpolator(), &kf.GetZInterpolator(),
                                   &kf.GetRInterpolator()};
            for (int c = 0; c < 4; ++c) {
                Vector2f c0, c1;
                c0.p.x = k.interp[c][0] * r;
                c0.p.y = k.interp[c][1] * r;
                c1.p.x = k.interp[c][2] * r;
                c1.p.y = k.interp[c][3] * r;
                ip[c]->SetC(c0, c1);
            }
        }
    }
    for (uint32_t t = 0; t < a.n_morph_tracks; ++t) {
        std::wstring nm = Name(L'm', (unsigned)a.morph_track_morph[t]);
        motion.RegisterMorph(nm);
        for (uint32_t j = 0; j < a.morph_track_key_count[t]; ++j) {
            const mmdgpu_morph_key& k = a.morph_keys[a.morph_track_key_begin[t] + j];
            motion.GetMorphKeyframe(nm, k.frame).SetWeight(k.weight);
        }
    }
}

void OneFrame(Poser& poser, MotionPlayer& player, size_t frame) {
    // main.cpp:1788-1821, physics off
    poser.ResetPosing();
    player.SeekFrame(frame);
    poser.PrePhysicsPosing();
    poser.PostPhysicsPosing();
    poser.Deform();
}

}  // namespace

struct ref_session {
    Model model;
    Motion motion;
    std::unique_ptr<Poser> poser;
    std::unique_ptr<MotionPlayer> player;
    bool has_motion = false;
};

extern "C" {

__attribute__((visibility("default"))) ref_session* ref_create(const mmdgpu_model_desc* m,
                                                                const mmdgpu_anim_desc* a_or_null) {
    ref_session* s = new ref_session();
    BuildModel(s->model, *m);
    if (a_or_null) {
        BuildMotion(s->motion, *a_or_null);
        s->has_motion = true;
    }
    s->poser.reset(new Poser(s->model));
    s->player.reset(new MotionPlayer(s->motion, *s->poser));
    return s;
}

__attribute__((visibility("default"))) void ref_destroy(ref_session* s) { delete s; }

// Skinning operators after Model::Normalize, in the descriptor's numbering (SDEF stays 3).
__attribute__((visibility("default"))) void ref_get_skinning(ref_session* s, uint8_t* type, int32_t* id4,
                                                              float* w4) {
    size_t n = s->model.GetVertexNum();
    for (size_t i = 0; i < n; ++i) {
        Model::Vertex<ref> vx = s->model.GetVertex(i);
        const Model::SkinningOperator& op = vx.GetSkinningOperator();
        int32_t* id = id4 + 4 * i;
        float* w = w4 + 4 * i;
        for (int k = 0; k < 4; ++k) {
            id[k] = -1;
            w[k] = 0;
        }
        switch (op.GetSkinningType()) {
        case Model::SkinningOperator::SKINNING_BDEF1:
            type[i] = MMDGPU_SKIN_BDEF1;
            id[0] = (int32_t)op.GetBDEF1().GetBoneID();
            break;
        case Model::SkinningOperator::SKINNING_BDEF2:
            type[i] = MMDGPU_SKIN_BDEF2;
            id[0] = (int32_t)op.GetBDEF2().GetBoneID(0);
            id[1] = (int32_t)op.GetBDEF2().GetBoneID(1);
            w[0] = op.GetBDEF2().GetBoneWeight();
            break;
        case Model::SkinningOperator::SKINNING_SDEF:
            type[i] = MMDGPU_SKIN_SDEF;
            id[0] = (int32_t)op.GetSDEF().GetBoneID(0);
            id[1] = (int32_t)op.GetSDEF().GetBoneID(1);
            w[0] = op.GetSDEF().GetBoneWeight();
            break;
        default:
            type[i] = MMDGPU_SKIN_BDEF4;
            for (int k = 0; k < 4; ++k) {
                id[k] = (int32_t)op.GetBDEF4().GetBoneID(k);
                w[k] = op.GetBDEF4().GetBoneWeight(k);
            }
            break;
        }
    }
}

// Per-link IK classification Poser::Poser computed (L/motion/poser_impl.inl:78-91), concatenated over
// IK bones in bone order.  fix: 0 NONE 1 X 2 Y 3 Z 4 ALL; order: 0 YZX 1 ZXY 2 XYZ.  Returns the count.
__attribute__((visibility("default"))) uint32_t ref_get_ik_class(ref_session* s, uint8_t* fix, uint8_t* order,
                                                                  uint32_t cap) {
    std::vector<uint8_t> f, o;
    for (size_t b = 0; b < s->model.GetBoneNum(); ++b)
        if (s->model.GetBone(b).IsHasIK()) Spy::IkClass(*s->poser, b, f, o);
    for (size_t i = 0; i < f.size() && i < cap; ++i) {
        fix[i] = f[i];
        order[i] = o[i];
    }
    return (uint32_t)f.size();
}

// One frame of main.cpp:1788-1821.  Any output pointer may be NULL.
//   pos, nrm : nv x 3     skin, local : nb x 16     poses : nb x 7 (T xyz, R xyzw)     rates : nm
__attribute__((visibility("default"))) int ref_run_frame(ref_session* s, uint32_t frame, float* pos, float* nrm,
                                                          float* skin, float* local, float* poses, float* rates) {
    OneFrame(*s->poser, *s->player, frame);
    size_t nv = s->model.GetVertexNum(), nb = s->model.GetBoneNum(), nm = s->model.GetMorphNum();
    if (pos) memcpy(pos, s->poser->pose_image.coordinates.data(), nv * 12);
    if (nrm) memcpy(nrm, s->poser->pose_image.normals.data(), nv * 12);
    for (size_t b = 0; b < nb; ++b) {
        if (skin) Spy::SkinningMatrix(*s->poser, b, skin + 16 * b);
        if (local) Spy::LocalMatrix(*s->poser, b, local + 16 * b);
        if (poses) Spy::Pose(*s->poser, b, poses + 7 * b);
    }
    if (rates) {
        for (size_t m = 0; m < nm; ++m) {
            std::wstring nmw = s->model.GetMorph(m).GetName();
            rates[m] = s->motion.IsMorphRegistered(nmw) ? s->motion.GetMorphPose(nmw, (size_t)frame).GetWeight() : 0.0f;
        }
    }
    return 0;
}

// main.cpp:1788-1821 with a host reactor in the middle: ResetPosing, SeekFrame, PrePhysicsPosing, then the listed
// bones' skinning (and optionally local) matrices are overwritten as PhysicsReactor::React would, PostPhysicsPosing, Deform.
__attribute__((visibility("default"))) int ref_run_frame_override(ref_session* s, uint32_t frame, uint32_t n, const int32_t* bones,
                                                                   const float* skin16, const float* local16_or_null,
                                                                   float* pos, float* nrm, float* skin, float* local) {
    Poser& p = *s->poser;
    p.ResetPosing();
    s->player->SeekFrame(frame);
    p.PrePhysicsPosing();
    for (uint32_t i = 0; i < n; ++i)
        Spy::Override(p, (size_t)bones[i], skin16 + 16 * i, local16_or_null ? local16_or_null + 16 * i : nullptr);
    p.PostPhysicsPosing();
    p.Deform();
    size_t nv = s->model.GetVertexNum(), nb = s->model.GetBoneNum();
    if (pos) memcpy(pos, p.pose_image.coordinates.data(), nv * 12);
    if (nrm) memcpy(nrm, p.pose_image.normals.data(), nv * 12);
    for (size_t b = 0; b < nb; ++b) {
        if (skin) Spy::SkinningMatrix(p, b, skin + 16 * b);
        if (local) Spy::LocalMatrix(p, b, local + 16 * b);
    }
    return 0;
}

// Sub-frame path: ResetPosing, MotionPlayer::SeekTime(seconds) (L/motion/poser_impl.inl:548-555), Pre, Post, Deform.
__attribute__((visibility("default"))) int ref_run_time(ref_session* s, double seconds, float* pos, float* nrm,
                                                         float* skin, float* poses, float* rates) {
    Poser& p = *s->poser;
    p.ResetPosing();
    s->player->SeekTime(seconds);
    p.PrePhysicsPosing();
    p.PostPhysicsPosing();
    p.Deform();
    size_t nv = s->model.GetVertexNum(), nb = s->model.GetBoneNum(), nm = s->model.GetMorphNum();
    if (pos) memcpy(pos, p.pose_image.coordinates.data(), nv * 12);
    if (nrm) memcpy(nrm, p.pose_image.normals.data(), nv * 12);
    for (size_t b = 0; b < nb; ++b) {
        if (skin) Spy::SkinningMatrix(p, b, skin + 16 * b);
        if (poses) Spy::Pose(p, b, poses + 7 * b);
    }
    if (rates)
        for (size_t m = 0; m < nm; ++m) {
            std::wstring nmw = s->model.GetMorph(m).GetName();
            rates[m] = s->motion.IsMorphRegistered(nmw) ? s->motion.GetMorphPose(nmw, seconds).GetWeight() : 0.0f;
        }
    return 0;
}

// Manual posing path: ResetPosing, SetBonePose/SetMorphPose for the listed items, Pre, Post, Deform.
__attribute__((visibility("default"))) int ref_run_manual(ref_session* s, uint32_t n_bone_poses,
                                                           const int32_t* bone, const float* pose7,
                                                           uint32_t n_morph_poses, const int32_t* morph,
                                                           const float* weight, float* pos, float* nrm, float* skin) {
    Poser& p = *s->poser;
    p.ResetPosing();
    for (uint32_t i = 0; i < n_bone_poses; ++i) {
        Vector4f r;
        for (int k = 0; k < 4; ++k) r.v[k] = pose7[7 * i + 3 + k];
        p.SetBonePose((size_t)bone[i], Motion::BonePose(V3(pose7 + 7 * i), r));
    }
    for (uint32_t i = 0; i < n_morph_poses; ++i) p.SetMorphPose((size_t)morph[i], Motion::MorphPose(weight[i]));
    p.PrePhysicsPosing();
    p.PostPhysicsPosing();
    p.Deform();
    size_t nv = s->model.GetVertexNum(), nb = s->model.GetBoneNum();
    if (pos) memcpy(pos, p.pose_image.coordinates.data(), nv * 12);
    if (nrm) memcpy(nrm, p.pose_image.normals.data(), nv * 12);
    if (skin)
        for (size_t b = 0; b < nb; ++b) Spy::SkinningMatrix(p, b, skin + 16 * b);
    return 0;
}

// main.cpp:838-859 repack of the current pose_image into 32-byte sokol vertices.
__attribute__((visibility("default"))) void ref_repack_sokol32(ref_session* s, float* out8) {
    size_t nv = s->model.GetVertexNum();
    const float mmd_to_meter = 0.1f;
    for (size_t i = 0; i < nv; ++i) {
        Model::Vertex<ref> vertex = s->model.GetVertex(i);
        Vector2f uv = vertex.GetUVCoordinate();
        const Vector3f& p = s->poser->pose_image.coordinates[i];
        const Vector3f& n = s->poser->pose_image.normals[i];
        float* o = out8 + 8 * i;
        o[0] = p.p.x * mmd_to_meter;
        o[1] = p.p.y * mmd_to_meter;
        o[2] = p.p.z * mmd_to_meter;
        o[3] = n.p.x;
        o[4] = n.p.y;
        o[5] = n.p.z;
        o[6] = uv.v[0];
        o[7] = uv.v[1];
    }
}

// CPU baseline: time the frame loop over `frames` with n_threads host threads, each owning a private
// Poser + MotionPlayer over the shared read-only Model / Motion (frames split in contiguous blocks).
// Returns wall seconds; *checksum receives a value that depends on every thread's last frame.
__attribute__((visibility("default"))) double ref_time_frames(ref_session* s, const uint32_t* frames,
                                                               uint32_t n_frames, uint32_t n_threads,
                                                               double* checksum) {
    if (n_threads < 1) n_threads = 1;
    std::vector<std::unique_ptr<Poser>> posers(n_threads);
    std::vector<std::unique_ptr<MotionPlayer>> players(n_threads);
    for (uint32_t t = 0; t < n_threads; ++t) {
        posers[t].reset(new Poser(s->model));
        players[t].reset(new MotionPlayer(s->motion, *posers[t]));
    }
    std::vector<double> sums(n_threads, 0.0);
    auto work = [&](uint32_t t) {
        uint32_t lo = (uint32_t)((uint64_t)n_frames * t / n_threads);
        uint32_t hi = (uint32_t)((uint64_t)n_frames * (t + 1) / n_threads);
        double acc = 0;
        size_t nv = s->model.GetVertexNum();
        for (uint32_t i = lo; i < hi; ++i) {
            OneFrame(*posers[t], *players[t], frames[i]);
            if (nv) acc += posers[t]->pose_image.coordinates[i % nv].p.x + posers[t]->pose_image.normals[(i * 7) % nv].p.y;
        }
        sums[t] = acc;
    };
    auto t0 = std::chrono::steady_clock::now();
    if (n_threads == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (uint32_t t = 0; t < n_threads; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    double c = 0;
    for (double x : sums) c += x;
    if (checksum) *checksum = c;
    return sec;
}

// Function-level known-answer interface: libmmd's own math on one row of `in` per case (layout: oracle/mmd_oracle.c,
// port_math_kat).  These are the functions SURVEY section 4 item 3 lists (L/util/math_impl.inl:1047-1224, 1265-1340,
// 1372-1428); tests/golden/make_math_kat.py stores their outputs as fixtures.
__attribute__((visibility("default"))) int ref_math_kat(int op, const float* in, uint32_t n, float* out) {
    static const int kin[11] = {5, 9, 9, 5, 4, 4, 8, 4, 4, 32, 3};
    static const int kout[11] = {1, 4, 4, 3, 4, 4, 4, 9, 4, 16, 3};
    if (op < 0 || op > 10) return -1;
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = in + (size_t)i * kin[op];
        float* o = out + (size_t)i * kout[op];
        auto quat = [](const float* p) { Quaternionf q; q.i = p[0]; q.j = p[1]; q.k = p[2]; q.e = p[3]; return q; };
        auto putq = [](const Quaternionf& q, float* p) { p[0] = q.i; p[1] = q.j; p[2] = q.k; p[3] = q.e; };
        auto vec3 = [](const float* p) { Vector3f v; v.v[0] = p[0]; v.v[1] = p[1]; v.v[2] = p[2]; return v; };
        switch (op) {
        case 0: {   // control bytes exactly as VmdReader feeds them (vmd_reader_impl.inl:32-41)
            const float r = 1.0f / 127.0f;
            Vector2f c0, c1;
            c0.p.x = (signed char)a[0] * r; c0.p.y = (signed char)a[1] * r;
            c1.p.x = (signed char)a[2] * r; c1.p.y = (signed char)a[3] * r;
            Bezier<float> bz;
            bz.SetC(c0, c1);
            o[0] = bz[a[4]];
            break;
        }
        case 1: {
            Vector4f x, y;
            for (int c = 0; c < 4; ++c) { x.v[c] = a[c]; y.v[c] = a[4 + c]; }
            Vector4f v = NLerp(x, y)[a[8]];
            for (int c = 0; c < 4; ++c) o[c] = v.v[c];
            break;
        }
        case 2: putq(SLerp(quat(a), quat(a + 4))[a[8]], o); break;
        case 3: {
            const int order = (int)a[4];
            Vector3f e = order == 1 ? QuaternionToZXY(quat(a)) : order == 2 ? QuaternionToXYZ(quat(a)) : QuaternionToYZX(quat(a));
            o[0] = e.v[0]; o[1] = e.v[1]; o[2] = e.v[2];
            break;
        }
        case 4: {
            const int order = (int)a[3];
            putq(order == 1 ? ZXYToQuaternion(vec3(a)) : order == 2 ? XYZToQuaternion(vec3(a)) : YZXToQuaternion(vec3(a)), o);
            break;
        }
        case 5: putq(AxisToQuaternion(vec3(a), a[3]), o); break;
        case 6: putq(quat(a) * quat(a + 4), o); break;
        case 7: {
            Matrix4f m = quat(a).ToRotateMatrix();
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) o[3 * r + c] = m.r.v[r].v[c];
            break;
        }
        case 8: putq(quat(a).Inverse(), o); break;
        case 9: {
            Matrix4f x, y;
            for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { x.r.v[r].v[c] = a[4 * r + c]; y.r.v[r].v[c] = a[16 + 4 * r + c]; }
            Matrix4f m = x * y;
            for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) o[4 * r + c] = m.r.v[r].v[c];
            break;
        }
        case 10: { Vector3f v = vec3(a).Normalize(); o[0] = v.v[0]; o[1] = v.v[1]; o[2] = v.v[2]; break; }
        }
    }
    return 0;
}

}  // extern "C"
