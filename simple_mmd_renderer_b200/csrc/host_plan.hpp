// host_plan.hpp — GPU-free flattening of a PMX-shaped model / VMD-shaped motion into the arrays the
// sm_100a kernels consume.  Everything libmmd precomputes in Model::Normalize
// (L/model/model_impl.inl:406-452) and Poser::Poser (L/motion/poser_impl.inl:16-128) happens here, plus
// what the device design adds: the wave schedule of the bone program, the morph application-slot tree
// and the per-vertex morph CSR.  (L/ = 3rd_party/libmmd/include/mmd/ of the reference.)
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/mmdgpu.h"

namespace mmdgpu {

// ---- per-bone static record read by the hierarchy kernel (48 bytes, three float4 loads) -----------
struct BoneStatic {
    float local_offset[3];  // position - parent position, or position (poser_impl.inl:40-46)
    int32_t parent;         // -1 = none
    float position[3];      // global_offset_matrix_ row 3 = -position (poser_impl.inl:36)
    int32_t append_parent;  // valid iff flags & (kAppendRot | kAppendTrans)
    float append_ratio;
    uint32_t flags;
    int32_t link_slot;      // index into the compact ikR / preIK arrays, -1 if the bone is no IK link
    int32_t morph_slot;     // index into the compact morphR / morphT arrays, -1 if no bone morph touches it
};
enum : uint32_t {
    kHasParent = 1u, kAppendRot = 2u, kAppendTrans = 4u, kIsLink = 8u, kHasIk = 16u, kPostPhysics = 32u
    // bits 31:16 of an IK bone's flags: index of its IkDesc (nested solves look it up from the bone)
};
constexpr int32_t kMaxIkDepth = 3;   // solve inside solve inside solve; deeper nesting is refused at load

struct IkDesc {
    int32_t bone, target, iterations;  // iterations already min(limit, 256) (poser_impl.inl:96)
    float angle_limit;
    int32_t link_begin, link_count;
    int32_t pad[2];
};
struct IkLink {
    int32_t bone;
    uint8_t limited, fix, order, pad;  // fix: 0 NONE 1 X 2 Y 3 Z 4 ALL ; order: 0 YZX 1 ZXY 2 XYZ
    float lo[3], hi[3];                // min / max after the swap of poser_impl.inl:74-77
};

// Chain-local image of one CCD IK solve (device design, no libmmd counterpart): the few bones a solve touches -
// its links, its target, the IK bone, their parents and append parents - renumbered 0..n_bones-1 together with
// translated copies of their static records, so that a solve can run on a < 2 KB private copy of the state.
struct IkImage {
    int32_t bones_begin, n_bones;    // into ik_img_bones (global bone ids) / ik_img_static (translated records)
    int32_t lslots_begin, n_lslots;  // into ik_img_lslots (global link slots)
    int32_t mslots_begin, n_mslots;  // into ik_img_mslots (global bone-morph slots)
    int32_t region_f4;               // float4 per solve: 7 n_bones + 2 n_lslots + 2 n_mslots, made odd (bank spread)
    int32_t pad;
};

enum : uint8_t { kOpEval = 0, kOpIk = 1, kOpSkin = 2 };
struct Op {
    uint8_t kind;
    int32_t arg;  // bone for EVAL / SKIN, IkDesc index for IK
};

// One entry of a bone morph in application order (per affected bone).
struct BoneMorphEntry {
    int32_t node;  // application slot whose rate scales it
    float translation[3];
    float rotation[4];
};

// One entry of a material morph in application order (per affected material; an "every material" entry is
// repeated under each material).
struct MaterialMorphEntry {
    int32_t node;     // application slot whose rate scales it
    uint32_t method;  // MMDGPU_MATERIAL_MUL / _ADD
    float value[MMDGPU_MATERIAL_FIELDS];
};

// Device-side skinning types (after Model::Normalize, the Lerp shortcuts and the compat mapping).
enum : uint8_t { kDevBdef1 = 0, kDevBdef2 = 1, kDevBdef4 = 2, kDevSdef = 3, kDevQdef = 4 };

struct Plan {
    uint32_t nv = 0, nb = 0, nm = 0;
    bool extensions = false;

    // vertices
    std::vector<uint8_t> norm_type;  // descriptor numbering after Model::Normalize (SDEF stays 3, QDEF -> 2)
    std::vector<uint8_t> dev_type;   // kDev*
    std::vector<uint16_t> bone_id;   // 4 nv, unused lanes 0
    std::vector<float> weight;       // 4 nv, unused lanes 0
    std::vector<float> position, normal, uv;
    std::vector<float> sdef_c, sdef_r0, sdef_r1;  // only when extensions && any SDEF

    // bones
    std::vector<BoneStatic> bones;
    std::vector<int32_t> order_pre, order_post;
    std::vector<IkDesc> iks;
    bool ik_nested = false;          // some solve's link or target is itself an IK bone (poser_impl.inl:203-206)
    std::vector<IkLink> links;       // in descriptor order (PLAN_IK_* arrays index this)
    std::vector<int32_t> link_bones; // compact list: link_slot -> bone
    std::vector<int32_t> reset_bones;// bones whose tot / local state is read before it is written

    // program
    std::vector<Op> ops;             // program order (libmmd's sequential order)
    std::vector<int32_t> op_wave;    // wave per op
    std::vector<int32_t> wave_begin; // n_waves + 1 offsets into wave_ops
    std::vector<int32_t> wave_ops;   // op ids grouped by wave (program order inside a wave)
    int32_t phase_split = 0;         // first wave of the post-physics segment

    // morph application slots (DFS expansion of the morph table, poser_impl.inl:328-360)
    std::vector<int32_t> node_morph, node_parent, node_depth;
    std::vector<float> node_mult;
    std::vector<int32_t> nodes_by_depth, depth_begin;

    // per-vertex CSR of vertex-morph entries in application order
    std::vector<uint32_t> csr_row;   // nv + 1
    std::vector<uint32_t> csr_node;
    std::vector<float> csr_offset;   // 3 per entry
    // per-vertex CSR of UV-morph entries (extensions only)
    std::vector<uint32_t> uv_row, uv_node;
    std::vector<float> uv_offset;    // 4 per entry (only .xy applied to the base UV)
    // ---- device vertex layout: kTileVerts-vertex tiles, each stored in a tile-local order that makes the 32 lanes
    //      of a warp step share a skinning type and a morph entry count (see build_tiles in host_plan.cpp)
    uint32_t nv_pad = 0, n_tiles = 0;
    std::vector<uint16_t> tile_orig;        // nv_pad: storage position -> PMX index within the tile
    std::vector<uint8_t> st_type;           // nv_pad: device skinning type per storage position
    std::vector<uint16_t> st_local_id;      // 4 nv_pad: tile-local bone indices per storage position
    std::vector<float> st_weight;           // 4 nv_pad
    std::vector<uint32_t> tile_bone_begin;  // n_tiles + 1
    std::vector<uint16_t> tile_bones;       // distinct global bone ids of every tile, ascending
    uint32_t max_tile_bones = 0;
    // sliced ELL of the vertex-morph entries: group = (tile, step j, warp w) = 32 lanes; rounds = max entry
    // count of the group; entry (round k, lane l) at ell_base + k * 32 + l; padding uses node = pad_node
    std::vector<uint32_t> ell_base;         // n_tiles * 32
    std::vector<uint32_t> ell_rounds;       // n_tiles * 32
    std::vector<uint32_t> ell_node;
    std::vector<float> ell_offset;          // 3 per entry
    uint32_t pad_node = 0;                  // application slot whose rate is always 0
    // extensions only: sliced ELL of the UV-morph entries (offset.xyzw), spherical-deform parameters
    std::vector<uint32_t> uv_ell_base, uv_ell_rounds, uv_ell_node;
    std::vector<float> uv_ell_offset;       // 4 per entry
    std::vector<float> st_sdef;             // 12 per storage position: C.xyz_, cr0.xyz_, cr1.xyz_

    // bone morphs grouped by affected bone
    std::vector<int32_t> morph_bones;        // morph_slot -> bone
    std::vector<int32_t> bone_morph_row;     // morph_bones.size() + 1
    std::vector<BoneMorphEntry> bone_morph_entries;

    // chain-local images of the CCD IK solves, one per entry of `iks` (see IkImage)
    std::vector<IkImage> ik_img;
    std::vector<int32_t> ik_img_bones;       // global bone id of every image bone
    std::vector<uint8_t> ik_img_written;     // 1: the solve evaluates (writes) this image bone: links, target
    std::vector<BoneStatic> ik_img_static;   // static records, references translated to image indices
    std::vector<int32_t> ik_img_lslots, ik_img_mslots;  // global link / bone-morph slot of every image slot
    std::vector<IkDesc> ik_img_desc;         // bone / target / link_begin translated
    std::vector<IkLink> ik_img_links;        // bone translated
    uint32_t ik_img_max_region = 0;
    bool ik_img_ok = false;

    // extensions only: material morphs grouped by affected material, application order inside a material
    uint32_t n_materials = 0;
    std::vector<int32_t> material_morph_row;  // n_materials + 1
    std::vector<MaterialMorphEntry> material_morph_entries;

    // names (only for models parsed from PMX bytes)
    std::vector<std::string> bone_names, morph_names;
    bool names_utf8 = false;  // PMX text encoding flag: UTF-8, else UTF-16LE

    // scratch for mmdgpu_plan_get
    std::vector<uint8_t> op_kind_u8, ik_fix_u8, ik_order_u8;
    std::vector<int32_t> op_arg_i32, phase_split_i32;
};

constexpr uint32_t kTileVerts = 512;       // vertices per tile = one CTA iteration
constexpr uint32_t kSlotGroup = 4;         // slots (frames / instances) a skinning CTA evaluates together
constexpr uint32_t kVertsPerThread = 4;    // consecutive storage positions one thread owns ("steps" of a warp)
constexpr uint32_t kSkinThreads = kTileVerts / kVertsPerThread;
constexpr uint32_t kSkinWarps = kSkinThreads / 32;
constexpr uint32_t kTileGroups = kTileVerts / 32;  // (step j) x (warp w) groups of 32 lanes
// storage position of sorted rank r inside a tile: group g = r / 32 is step j = g / kSkinWarps of warp
// w = g % kSkinWarps; lane l = r % 32 owns positions (w*32 + l) * kVertsPerThread + j
inline uint32_t tile_position_of_rank(uint32_t r) {
    const uint32_t g = r >> 5, l = r & 31u;
    return ((g % kSkinWarps) * 32u + l) * kVertsPerThread + g / kSkinWarps;
}

// Returns MMDGPU_OK or an error code with a message in `err`.
mmdgpu_status build_plan(const mmdgpu_model_desc& d, const mmdgpu_options* opt, Plan& out, std::string& err);

// ---- animation -----------------------------------------------------------------------------------
struct HostAnim {
    uint32_t nb = 0, nm = 0, length = 0;
    // per model bone: key range (count 0 + tracked 1 = registered-but-empty track)
    std::vector<uint32_t> bone_key_begin, bone_key_count;
    std::vector<uint8_t> bone_tracked;
    std::vector<uint32_t> key_frame;
    std::vector<float> key_T;        // 4 per key (w unused)
    std::vector<float> key_R;        // 4 per key
    std::vector<uint32_t> key_curve; // 4 per key: Bezier table index, 0xFFFFFFFF = linear
    std::vector<float> tables;       // 32 floats per deduplicated table
    std::vector<uint32_t> morph_key_begin, morph_key_count;
    std::vector<uint8_t> morph_tracked;
    std::vector<uint32_t> mkey_frame;
    std::vector<float> mkey_weight;
};
mmdgpu_status build_anim(const mmdgpu_anim_desc& d, uint32_t nb, uint32_t nm, HostAnim& out, std::string& err);

// Bezier::presample of L/util/math_impl.inl:1398-1428 for one VMD control quadruple; returns true if linear.
bool bezier_table(const int8_t ctrl[4], float table[32]);

// ---- PMX / VMD byte streams ----------------------------------------------------------------------
struct ParsedModel {  // owns the arrays a mmdgpu_model_desc points into
    std::vector<float> position, normal, uv, weight, sdef_c, sdef_r0, sdef_r1, bone_position, bone_append_ratio,
        ik_angle_limit, ik_link_lo, ik_link_hi;
    std::vector<uint8_t> skin_type, ik_link_has_limit, morph_type;
    std::vector<int32_t> bone_id, bone_parent, bone_transform_level, bone_append_parent, ik_target, ik_iterations,
        ik_link_bone;
    std::vector<uint16_t> bone_flags;
    std::vector<uint32_t> ik_link_begin, ik_link_count, morph_entry_begin, morph_entry_count;
    std::vector<mmdgpu_vertex_morph_entry> vme;
    std::vector<mmdgpu_uv_morph_entry> uvme;
    std::vector<mmdgpu_bone_morph_entry> bme;
    std::vector<mmdgpu_group_morph_entry> gme;
    std::vector<mmdgpu_material_morph_entry> mme;
    uint32_t n_materials = 0;
    std::vector<std::string> bone_names, morph_names;  // raw bytes as stored (UTF-16LE or UTF-8)
    bool utf8 = false;
    mmdgpu_model_desc desc{};
    void finish();  // point desc at the vectors
};
mmdgpu_status parse_pmx(const void* bytes, size_t n, ParsedModel& out, std::string& err);

struct ParsedMotion {
    std::vector<int32_t> bone_track_bone, morph_track_morph;
    std::vector<uint32_t> bt_begin, bt_count, mt_begin, mt_count;
    std::vector<mmdgpu_bone_key> bone_keys;
    std::vector<mmdgpu_morph_key> morph_keys;
    mmdgpu_anim_desc desc{};
    void finish();
};
// Joins VMD track names (Shift-JIS, 15 bytes, NUL-trimmed) to the plan's names.  PMX names are UTF-16LE or
// UTF-8; the join transcodes the ASCII / half-width subset exactly and otherwise compares through a
// Shift-JIS -> UTF-16 table lookup supplied by the caller-independent decoder in pmx_vmd.cpp.
mmdgpu_status parse_vmd(const void* bytes, size_t n, const Plan& plan, ParsedMotion& out, std::string& err);

}  // namespace mmdgpu
