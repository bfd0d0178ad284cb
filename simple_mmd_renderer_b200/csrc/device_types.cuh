// device_types.cuh — plain structs passed by value to the sm_100a kernels.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "host_plan.hpp"

namespace mmdgpu {

static_assert(kVertsPerThread * kSkinThreads == kTileVerts, "one CTA iteration = one tile");
static_assert(kVertsPerThread == 2 || kVertsPerThread == 4, "vector loads are written for 2 or 4 positions per thread");

// Static model image in HBM.  Vertex streams are structure-of-arrays in TILE STORAGE ORDER (host_plan.hpp):
// a thread owns 4 consecutive storage positions, so every global access is a 16-byte vector, and the 32 lanes
// of a warp step share a skinning type and a morph entry count.
struct DevModel {
    uint32_t nv, nv_pad, nb, nm;
    uint32_t n_nodes, n_nodes_pad;  // morph application slots; padded to a multiple of 4, >= n_nodes + 1 (the always-zero slot)
    uint32_t n_tiles, max_tile_bones;
    uint32_t global_palette;        // 1: tiles touch too many bones to stage; bone ids are global and the palette is read from HBM / L2
    // vertex streams (nv_pad entries each, storage order)
    const float *px, *py, *pz, *nx, *ny, *nz;
    const uint2* ids;        // 4 x u16 tile-local bone indices; bits 15:13 of id0 carry the device skinning type
    const float4* weights;
    const float2* uv;
    const uint16_t* orig;    // per storage position: PMX index within the tile
    const uint2* ell_hdr;    // per 32-lane group (tile, step, warp): (first entry, rounds)
    const float4* ell_ent;   // (offset.xyz, byte offset of the slot's float4 of rates); entry (round k, lane l) at base + 32 k + l
    const uint32_t* tile_bone_begin;  // n_tiles + 1
    const uint16_t* tile_bones;       // distinct bones of each tile
    // extension streams (NULL in libmmd-exact mode)
    uint32_t extensions;
    const float4* sdef;        // 3 per storage position: C, cr0, cr1 (spherical deform)
    const uint2* uv_ell_hdr;   // UV-morph sliced ELL, same group structure as ell_hdr
    const float4* uv_ell_ent;  // (du, dv, byte offset of the slot's float4 of rates, unused)
    // bones
    const BoneStatic* bones;
    const IkDesc* iks;
    const IkLink* links;
    uint32_t ik_nested;      // some solve's link / target has IK itself: solves recurse (kernels.cu)
    const int32_t* reset_bones;
    uint32_t n_reset, n_link_slots, n_morph_slots;
    // program: ops grouped by wave; op word = kind << 28 | arg
    const uint32_t* wave_begin;
    const uint32_t* wave_ops;
    uint32_t n_waves, phase_split, n_ops;
    // morph application slots
    const int32_t* node_morph;
    const int32_t* node_parent;
    const float* node_mult;
    const int32_t* nodes_by_depth;
    const int32_t* depth_begin;
    uint32_t n_depths;
    // bone morphs grouped by bone
    const int32_t* bone_morph_row;
    const BoneMorphEntry* bone_morph_entries;
    // chain-local IK images (one per IkDesc), used by the flat IK-wave kernel
    const IkImage* ik_img;
    const int32_t* ik_img_bones;       // global bone id of every image bone
    const uint8_t* ik_img_written;     // 1: the solve writes this image bone (links, target)
    const BoneStatic* ik_img_static;   // static records with parent / append parent / slots translated to image indices
    const int32_t* ik_img_lslots;      // global link slot of every image link slot
    const int32_t* ik_img_mslots;      // global bone-morph slot of every image morph slot
    const IkDesc* ik_img_desc;         // IkDesc with bone / target / link_begin translated
    const IkLink* ik_img_links;        // IkLink with bone translated
    uint32_t ik_img_max_region;        // largest region_f4 over the model's solves (0: no images)
    // extensions: material morphs grouped by material
    uint32_t n_materials;
    const int32_t* material_morph_row;
    const MaterialMorphEntry* material_morph_entries;
};

// Flattened VMD clip bound to one model.
struct DevAnim {
    const uint32_t* bone_key_begin;
    const uint32_t* bone_key_count;
    const uint8_t* bone_tracked;
    const uint32_t* key_frame;
    const float4* key_T;
    const float4* key_R;
    const uint4* key_curve;   // Bezier table index per channel X, Y, Z, R; 0xFFFFFFFF = linear
    const float* tables;      // 32 floats per table
    const uint32_t* morph_key_begin;
    const uint32_t* morph_key_count;
    const uint8_t* morph_tracked;
    const uint32_t* mkey_frame;
    const float* mkey_weight;
};

// Per-slot dynamic state and outputs.  Slot = instance * n_frames + k.
struct DevFrames {
    uint32_t n_slots, n_instances, n_frames;
    float4* poseR;      // [slot][nb]   BoneImage::rotation_
    float4* poseT;      // [slot][nb]   BoneImage::translation_ (w unused)
    float* rate;        // [slot][nm]   Poser::morph_rates_
    float* node_rate;   // [slot / 4][n_nodes_pad][slot % 4]  rate of every application slot, 0 = skipped
    float4* totR;       // [slot][nb]
    float4* totT;       // [slot][nb]
    float* local;       // [slot][nb][12]  rows 0..3 x cols 0..2 of BoneImage::local_matrix_
    float4* ikR;        // [slot][n_link_slots]
    float4* preIK;      // [slot][n_link_slots]
    float4* morphR;     // [slot][n_morph_slots]
    float4* morphT;     // [slot][n_morph_slots]
    float4* palette;    // [slot][nb][3]   column c of skinning_matrix_: (M0c, M1c, M2c, M3c)
    float4* pal_ext;    // [slot][nb][2]   extensions: rotation quaternion and dual part of the skinning transform
    float2* out_uv;     // extensions, SOA layout: [slot][uv_stride] morphed UV
    float* material_images;  // extensions: [slot][n_materials][2][28] multiplicative then additive image
    // Vertex outputs: the library's own buffers ([slot][nv_pad] records) or caller-owned device memory bound with
    // mmdgpu_frames_bind_output (e.g. a mapped GL vertex buffer, [slot][>= nv] records).  Only vertices < nv are stored.
    float* out_pos;     // SOA: [slot][pos_stride floats], 3 per vertex
    float* out_nrm;     // SOA: [slot][nrm_stride floats]
    float4* out_inter;  // INTERLEAVED: [slot][inter_stride float4], 2 per vertex
    size_t pos_stride, nrm_stride, inter_stride, uv_stride;  // slot strides in elements of the respective pointer
    uint32_t* frame_id; // [slot]
    double* time_s;     // [slot] seconds, for MotionPlayer::SeekTime
};

}  // namespace mmdgpu
