/*
 * mmdgpu.h — C-ABI of the B200-native MMD deformation library (libmmdgpu.so).
 *
 * This is the drop-in boundary for the per-frame deformation path that
 * simple_mmd_renderer drives through libmmd (main.cpp:1786-1825):
 *
 *     ResetPosing -> SeekFrame -> PrePhysicsPosing -> [React] ->
 *     PostPhysicsPosing -> Deform -> UpdateDeformedVertices
 *
 * libmmd is header-only C++ inlined into main.cpp, so there is no FFI in the
 * reference; every entry point below cites the libmmd member (file:line under
 * 3rd_party/libmmd/include/mmd/, abbreviated L/) or the main.cpp lines it
 * replaces.  Signatures are plain C: pointers, sizes, integer status codes.
 * Nothing here throws and nothing here falls back to the CPU: a call that
 * needs the device and cannot reach it returns MMDGPU_ERR_CUDA.
 *
 * Conventions (identical to libmmd, L/util/math.inl:10-19):
 *   - vectors are rows, y = x * M; 4x4 matrices are 16 floats row-major with
 *     the translation in row 3 (elements 12..14);
 *   - quaternions are stored (x, y, z, w) = libmmd's (i, j, k, e);
 *   - indices are int32 with -1 = "none" (libmmd: size_t nil);
 *   - all host arrays are little-endian, borrowed for the duration of the call.
 *
 * Threading: one context per host thread / CUDA stream.  Calls on one context
 * are issued in stream order.  Different contexts are independent.
 */
#ifndef MMDGPU_H_INCLUDED
#define MMDGPU_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define MMDGPU_API __declspec(dllexport)
#else
#define MMDGPU_API __attribute__((visibility("default")))
#endif

#define MMDGPU_VERSION_MAJOR 0
#define MMDGPU_VERSION_MINOR 1

/* ------------------------------------------------------------------ status */

typedef int mmdgpu_status;
enum {
    MMDGPU_OK              = 0,
    MMDGPU_ERR_INVALID_ARG = -1, /* NULL where data is required, sizes inconsistent            */
    MMDGPU_ERR_BAD_INDEX   = -2, /* bone / vertex / morph index out of range, group-morph cycle */
    MMDGPU_ERR_UNSUPPORTED = -3, /* valid input that this build cannot represent               */
    MMDGPU_ERR_CUDA        = -4, /* CUDA runtime error (no device, launch failure, ...)        */
    MMDGPU_ERR_OOM         = -5, /* host or device allocation failed                           */
    MMDGPU_ERR_PARSE       = -6  /* PMX / VMD byte stream malformed                             */
};

/* ----------------------------------------------------------------- handles */

typedef struct mmdgpu_context*   mmdgpu_context_t;   /* one per GPU / stream                        */
typedef struct mmdgpu_model*     mmdgpu_model_t;     /* immutable after create; shared by instances  */
typedef struct mmdgpu_animation* mmdgpu_animation_t; /* flattened VMD tracks bound to one model      */
typedef struct mmdgpu_frames*    mmdgpu_frames_t;    /* device state + output for instances x frames */
typedef struct mmdgpu_plan*      mmdgpu_plan_t;      /* host-only flattened model (no GPU needed)    */

/* ------------------------------------------------------- PMX-shaped model */

/* PMX skinning numbering, L/reader/pmx_reader_impl.inl:67-99 (+4 = PMX 2.1 QDEF). */
enum {
    MMDGPU_SKIN_BDEF1 = 0,
    MMDGPU_SKIN_BDEF2 = 1,
    MMDGPU_SKIN_BDEF4 = 2,
    MMDGPU_SKIN_SDEF  = 3,
    MMDGPU_SKIN_QDEF  = 4
};

/* PMX bone flag bits, L/reader/interprete/pmx_types.inl:46-59.  Only these four are consumed. */
enum {
    MMDGPU_BONE_HAS_IK           = 0x0020,
    MMDGPU_BONE_APPEND_ROTATE    = 0x0100,
    MMDGPU_BONE_APPEND_TRANSLATE = 0x0200,
    MMDGPU_BONE_POST_PHYSICS     = 0x1000
};

/* PMX morph numbering, L/model/model.inl:485-495. */
enum {
    MMDGPU_MORPH_GROUP    = 0,
    MMDGPU_MORPH_VERTEX   = 1,
    MMDGPU_MORPH_BONE     = 2,
    MMDGPU_MORPH_UV       = 3,
    MMDGPU_MORPH_EXT_UV1  = 4,
    MMDGPU_MORPH_EXT_UV2  = 5,
    MMDGPU_MORPH_EXT_UV3  = 6,
    MMDGPU_MORPH_EXT_UV4  = 7,
    MMDGPU_MORPH_MATERIAL = 8
};

typedef struct mmdgpu_vertex_morph_entry { uint32_t vertex; float offset[3]; } mmdgpu_vertex_morph_entry;
typedef struct mmdgpu_uv_morph_entry     { uint32_t vertex; float offset[4]; } mmdgpu_uv_morph_entry;
typedef struct mmdgpu_bone_morph_entry   { uint32_t bone; float translation[3]; float rotation[4]; } mmdgpu_bone_morph_entry;
typedef struct mmdgpu_group_morph_entry  { uint32_t morph; float rate; } mmdgpu_group_morph_entry;
/* Material morph entry, the PMX record of L/reader/interprete/pmx_types.inl:61-72 in its field order:
 * value[0..3] diffuse rgba, [4..6] specular, [7] shininess, [8..10] ambient, [11..14] edge colour,
 * [15] edge size, [16..19] texture, [20..23] sphere (sub) texture, [24..27] toon texture.
 * material < 0 or >= n_materials: every material (pmx_reader_impl.inl:327-334). */
#define MMDGPU_MATERIAL_FIELDS 28
enum { MMDGPU_MATERIAL_MUL = 0, MMDGPU_MATERIAL_ADD = 1 };  /* model.inl:396-399 */
typedef struct mmdgpu_material_morph_entry { int32_t material; uint32_t method; float value[MMDGPU_MATERIAL_FIELDS]; } mmdgpu_material_morph_entry;

/*
 * Flat image of mmd::Model (L/model/model.inl:719-734) restricted to what
 * Poser consumes.  Arrays may be NULL only when their count is zero (or where
 * marked optional).
 */
typedef struct mmdgpu_model_desc {
    /* vertices — Model::VertexInfo, L/model/model.inl:719-726 */
    uint32_t       n_vertices;
    const float*   position;      /* 3n  */
    const float*   normal;        /* 3n  */
    const float*   uv;            /* 2n, optional (NULL = zeros) */
    const uint8_t* skin_type;     /* n, MMDGPU_SKIN_*            */
    const int32_t* bone_id;       /* 4n, unused lanes ignored    */
    const float*   weight;        /* 4n; BDEF2/SDEF use weight[0] (weight of bone_id[0]) */
    const float*   sdef_c;        /* 3n, optional */
    const float*   sdef_r0;       /* 3n, optional */
    const float*   sdef_r1;       /* 3n, optional */

    /* bones — Model::Bone, L/model/model.inl:183-332 */
    uint32_t        n_bones;
    const float*    bone_position;        /* 3nb */
    const int32_t*  bone_parent;          /* nb, -1 or >= nb: no parent (poser_impl.inl:39-46) */
    const int32_t*  bone_transform_level; /* nb */
    const uint16_t* bone_flags;           /* nb, MMDGPU_BONE_* */
    const int32_t*  bone_append_parent;   /* nb, optional */
    const float*    bone_append_ratio;    /* nb, optional */
    /* IK, valid where HAS_IK — Bone::IKInfo */
    const int32_t*  ik_target;            /* nb, optional */
    const int32_t*  ik_iterations;        /* nb */
    const float*    ik_angle_limit;       /* nb */
    const uint32_t* ik_link_begin;        /* nb */
    const uint32_t* ik_link_count;        /* nb */
    uint32_t        n_ik_links;
    const int32_t*  ik_link_bone;         /* n_ik_links, tip-most first as stored in PMX */
    const uint8_t*  ik_link_has_limit;    /* n_ik_links */
    const float*    ik_link_lo;           /* 3*n_ik_links */
    const float*    ik_link_hi;           /* 3*n_ik_links */

    /* morphs — Model::Morph, L/model/model.inl:334-517 */
    uint32_t        n_morphs;
    const uint8_t*  morph_type;           /* nm, MMDGPU_MORPH_* */
    const uint32_t* morph_entry_begin;    /* nm, index into the pool of the morph's type */
    const uint32_t* morph_entry_count;    /* nm */
    uint32_t n_vertex_morph_entries; const mmdgpu_vertex_morph_entry* vertex_morph_entries;
    uint32_t n_uv_morph_entries;     const mmdgpu_uv_morph_entry*     uv_morph_entries;
    uint32_t n_bone_morph_entries;   const mmdgpu_bone_morph_entry*   bone_morph_entries;
    uint32_t n_group_morph_entries;  const mmdgpu_group_morph_entry*  group_morph_entries;
    /* materials: only their count and the material-morph pool (extensions; libmmd never fills
     * Poser::material_mul_images_ / material_add_images_, L/motion/poser.inl:160-161) */
    uint32_t n_materials;
    uint32_t n_material_morph_entries; const mmdgpu_material_morph_entry* material_morph_entries;
} mmdgpu_model_desc;

/* ------------------------------------------------------ VMD-shaped motion */

/*
 * One bone keyframe: the VMD 111-byte record minus the name
 * (L/reader/interprete/vmd_types.inl:22-31).  interp[c] = (x0, y0, x1, y1) of
 * channel c in {X, Y, Z, R} = bytes [0],[4],[8],[12] of the channel's
 * 16-byte block (L/reader/vmd_reader_impl.inl:32-61); signed as in the reader.
 */
typedef struct mmdgpu_bone_key {
    uint32_t frame;
    float    translation[3];
    float    rotation[4];
    int8_t   interp[4][4];
} mmdgpu_bone_key;

typedef struct mmdgpu_morph_key { uint32_t frame; float weight; } mmdgpu_morph_key;

/*
 * Flat image of mmd::Motion (L/motion/motion.inl:128-129) after the name join
 * of MotionPlayer::MotionPlayer (L/motion/poser_impl.inl:522-537): one track
 * per animated model bone / morph.  Keys of a track need not be sorted; equal
 * frames keep the last one (std::map::operator[] semantics,
 * L/motion/motion_impl.inl:221-227).  A bone or morph may appear in at most
 * one track.
 */
typedef struct mmdgpu_anim_desc {
    uint32_t               n_bone_tracks;
    const int32_t*         bone_track_bone;       /* model bone index per track */
    const uint32_t*        bone_track_key_begin;
    const uint32_t*        bone_track_key_count;
    uint32_t               n_bone_keys;
    const mmdgpu_bone_key* bone_keys;
    uint32_t                n_morph_tracks;
    const int32_t*          morph_track_morph;
    const uint32_t*         morph_track_key_begin;
    const uint32_t*         morph_track_key_count;
    uint32_t                n_morph_keys;
    const mmdgpu_morph_key* morph_keys;
} mmdgpu_anim_desc;

/* ------------------------------------------------------------------ output */

typedef enum mmdgpu_layout {
    /* pose_image.coordinates / .normals (L/motion/poser.inl:17-20): two planes of n x float3 */
    MMDGPU_LAYOUT_SOA_POS_NRM = 0,
    /* main.cpp:50-54 Vertex{pos[3]*0.1f, normal[3], uv[2]} = 32 B, as main.cpp:838-859 packs it */
    MMDGPU_LAYOUT_INTERLEAVED_SOKOL32 = 1
} mmdgpu_layout;

typedef enum mmdgpu_stream_id {
    MMDGPU_STREAM_POSITION    = 0, /* SOA: n x float3                    */
    MMDGPU_STREAM_NORMAL      = 1, /* SOA: n x float3                    */
    MMDGPU_STREAM_INTERLEAVED = 2, /* INTERLEAVED: n x 32 B              */
    MMDGPU_STREAM_SKIN_MATRIX = 3, /* device palette, nb x 12 floats (3 columns x float4) */
    MMDGPU_STREAM_UV          = 4  /* SOA + extensions only: n x float2 morphed UV           */
} mmdgpu_stream_id;

/* Behaviour switches.  Zero-initialised = libmmd-exact. */
typedef struct mmdgpu_options {
    /* 0: libmmd-exact (SDEF -> BDEF2 lerp, QDEF -> BDEF4 lerp, UV morphs ignored;
     *    L/motion/poser_impl.inl:417-426, :355-358).
     * 1: extensions — what the PMX format means but libmmd does not implement: spherical SDEF (rotation
     *    slerp(q0, q1, w1) about C, centres blended through R0 / R1), dual-quaternion QDEF, applied UV morphs
     *    (type 3; output stream MMDGPU_STREAM_UV, or the uv fields of the interleaved record).  PARITY UNPINNED:
     *    there is no libmmd behaviour to compare with; tests check them against an fp64 restatement of the same
     *    formulas and through self-consistency properties. */
    uint32_t extensions;
    uint32_t reserved[7];
} mmdgpu_options;

/* ---------------------------------------------------------------- contexts */

MMDGPU_API int         mmdgpu_version(void);
MMDGPU_API const char* mmdgpu_status_string(mmdgpu_status s);

/* cuda_stream_or_null: a cudaStream_t the caller owns (e.g. torch's current stream), or NULL to
 * let the context create its own non-blocking stream. */
MMDGPU_API mmdgpu_status mmdgpu_context_create(int device, void* cuda_stream_or_null, mmdgpu_context_t* out);
MMDGPU_API void          mmdgpu_context_destroy(mmdgpu_context_t ctx);
MMDGPU_API const char*   mmdgpu_last_error(mmdgpu_context_t ctx);
MMDGPU_API mmdgpu_status mmdgpu_context_synchronize(mmdgpu_context_t ctx);
MMDGPU_API void*         mmdgpu_context_stream(mmdgpu_context_t ctx);
/* Number of kernels this context has launched since creation (bench "gpu_launches"). */
MMDGPU_API uint64_t      mmdgpu_context_launch_count(mmdgpu_context_t ctx);

/* Device timing of this context's own kernels: when enabled, every launch is bracketed by CUDA events on the
 * context's stream.  profile_read synchronises the stream, returns the summed device milliseconds and the
 * launch count per kernel since the last read, and resets both. */
typedef enum mmdgpu_kernel_id {
    MMDGPU_KERNEL_POSE_SAMPLE = 0, /* K1: VMD key-frame sampling                      */
    MMDGPU_KERNEL_HIERARCHY   = 1, /* K2: morph rates, bone program, CCD IK, palette  */
    MMDGPU_KERNEL_SKIN        = 2, /* K3: morph gather + skinning                     */
    MMDGPU_KERNEL_COUNT       = 3
} mmdgpu_kernel_id;
MMDGPU_API mmdgpu_status mmdgpu_context_set_profiling(mmdgpu_context_t ctx, int enabled);
MMDGPU_API mmdgpu_status mmdgpu_context_profile_read(mmdgpu_context_t ctx, double ms_total[MMDGPU_KERNEL_COUNT],
                                                     uint64_t launches[MMDGPU_KERNEL_COUNT]);
/* Make the context's compute stream wait for every asynchronous download issued so far, so that an event
 * recorded on the compute stream afterwards covers them (end-to-end timing). */
MMDGPU_API mmdgpu_status mmdgpu_context_join_downloads(mmdgpu_context_t ctx);

/* ------------------------------------------------------------------ models */

/* Replaces: PmxReader::ReadModel's result + Model::Normalize (L/model/model_impl.inl:406-452)
 * + Poser::Poser's static precomputation (L/motion/poser_impl.inl:16-128).  Unlike libmmd,
 * which never bounds-checks, out-of-range bone / vertex / morph ids are rejected with BAD_INDEX. */
MMDGPU_API mmdgpu_status mmdgpu_model_create_from_arrays(mmdgpu_context_t ctx, const mmdgpu_model_desc* desc,
                                                         const mmdgpu_options* opt_or_null, mmdgpu_model_t* out);
/* PMX 2.0 / 2.1 byte stream, layout of L/reader/pmx_reader_impl.inl:16-449. */
MMDGPU_API mmdgpu_status mmdgpu_model_create_from_pmx(mmdgpu_context_t ctx, const void* bytes, size_t n,
                                                      const mmdgpu_options* opt_or_null, mmdgpu_model_t* out);
MMDGPU_API void          mmdgpu_model_destroy(mmdgpu_model_t model);
MMDGPU_API uint32_t      mmdgpu_model_vertex_count(mmdgpu_model_t model);
MMDGPU_API uint32_t      mmdgpu_model_bone_count(mmdgpu_model_t model);
MMDGPU_API uint32_t      mmdgpu_model_morph_count(mmdgpu_model_t model);
MMDGPU_API uint32_t      mmdgpu_model_material_count(mmdgpu_model_t model);
/* The host plan the model was built from (owned by the model). */
MMDGPU_API mmdgpu_plan_t mmdgpu_model_plan(mmdgpu_model_t model);
/* Bone / morph lookup by raw name bytes as stored in the PMX (only for models created from PMX). */
MMDGPU_API int32_t       mmdgpu_model_find_bone(mmdgpu_model_t model, const void* name_bytes, size_t n);
MMDGPU_API int32_t       mmdgpu_model_find_morph(mmdgpu_model_t model, const void* name_bytes, size_t n);

/* -------------------------------------------------------------- animations */

/* Replaces: mmd::Motion storage + MotionPlayer's name join (L/motion/poser_impl.inl:522-537). */
MMDGPU_API mmdgpu_status mmdgpu_animation_create_from_arrays(mmdgpu_context_t ctx, mmdgpu_model_t model,
                                                             const mmdgpu_anim_desc* desc, mmdgpu_animation_t* out);
/* VMD byte stream (50-B header, 111-B bone records, 23-B morph records; L/reader/vmd_reader_impl.inl:9-79).
 * Track names are joined to the model's bone / morph names byte-wise (Shift-JIS, NUL-trimmed). */
MMDGPU_API mmdgpu_status mmdgpu_animation_create_from_vmd(mmdgpu_context_t ctx, mmdgpu_model_t model,
                                                          const void* bytes, size_t n, mmdgpu_animation_t* out);
MMDGPU_API void          mmdgpu_animation_destroy(mmdgpu_animation_t anim);
/* Motion::GetLength (L/motion/motion_impl.inl:242-245): largest key frame. */
MMDGPU_API uint32_t      mmdgpu_animation_length(mmdgpu_animation_t anim);

/* ------------------------------------------------------------------ frames */

/* A frames object holds n_instances x n_frames independent "slots"; slot = instance * n_frames + k.
 * Each slot is one mmd::Poser worth of state: bone poses, morph rates, bone matrices and one
 * deformed vertex buffer (Poser::pose_image, L/motion/poser.inl:17-20). */
MMDGPU_API mmdgpu_status mmdgpu_frames_create(mmdgpu_context_t ctx, mmdgpu_model_t model, uint32_t n_instances,
                                              uint32_t n_frames, mmdgpu_layout layout, mmdgpu_frames_t* out);
MMDGPU_API void          mmdgpu_frames_destroy(mmdgpu_frames_t frames);
MMDGPU_API uint32_t      mmdgpu_frames_slot_count(mmdgpu_frames_t frames);

/* Poser::ResetPosing (L/motion/poser_impl.inl:130-140): zero morph rates, identity bone poses.
 * The redundant Pre+PostPhysicsPosing evaluation libmmd runs inside ResetPosing is not executed:
 * its results are overwritten by the next PrePhysicsPosing. */
MMDGPU_API mmdgpu_status mmdgpu_reset_posing(mmdgpu_frames_t frames);
/* MotionPlayer::SeekFrame(size_t) (L/motion/poser_impl.inl:539-546) for every slot: per_instance has
 * n_instances handles (instance i's clip), frame_per_slot has n_slots frame ids (host memory). */
MMDGPU_API mmdgpu_status mmdgpu_seek_frame(mmdgpu_frames_t frames, const mmdgpu_animation_t* per_instance,
                                           const uint32_t* frame_per_slot);
/* Same, frame ids generated on the device: slot (i, k) seeks first_frame[i] + k * frame_stride.
 * first_frame_per_instance is host memory (n_instances). */
MMDGPU_API mmdgpu_status mmdgpu_seek_frame_range(mmdgpu_frames_t frames, const mmdgpu_animation_t* per_instance,
                                                 const uint32_t* first_frame_per_instance, uint32_t frame_stride);
/* MotionPlayer::SeekTime(double) (L/motion/poser_impl.inl:548-555; sampling of L/motion/motion_impl.inl:321-380,
 * 426-470): frame = seconds * 30 as a double, no snapping to key frames.  time_per_slot: n_slots seconds (host). */
MMDGPU_API mmdgpu_status mmdgpu_seek_time(mmdgpu_frames_t frames, const mmdgpu_animation_t* per_instance,
                                          const double* time_per_slot);
/* mmdgpu_reset_posing followed by mmdgpu_seek_frame / _seek_time as ONE sampling launch: bones and morphs the clip does not
 * animate get identity / zero, the others their sampled key frames (what main.cpp:1788-1796 leaves in the Poser). */
MMDGPU_API mmdgpu_status mmdgpu_reset_and_seek_frame(mmdgpu_frames_t frames, const mmdgpu_animation_t* per_instance,
                                                     const uint32_t* frame_per_slot);
MMDGPU_API mmdgpu_status mmdgpu_reset_and_seek_time(mmdgpu_frames_t frames, const mmdgpu_animation_t* per_instance,
                                                    const double* time_per_slot);
/* ResetPosing + SeekFrame (or SeekTime) + PrePhysicsPosing + PostPhysicsPosing in one call - main.cpp:1788-1810 without the
 * physics step: one sampling launch and ONE hierarchy pass over the whole bone program (pre- and post-physics bones).
 * Deform is the caller's next call. */
MMDGPU_API mmdgpu_status mmdgpu_pose_frame(mmdgpu_frames_t frames, const mmdgpu_animation_t* per_instance,
                                           const uint32_t* frame_per_slot);
MMDGPU_API mmdgpu_status mmdgpu_pose_time(mmdgpu_frames_t frames, const mmdgpu_animation_t* per_instance,
                                          const double* time_per_slot);
/* Poser::SetBonePose / SetMorphPose by index (L/motion/poser_impl.inl:466-480). */
MMDGPU_API mmdgpu_status mmdgpu_set_bone_pose(mmdgpu_frames_t frames, uint32_t slot, uint32_t bone,
                                              const float translation[3], const float rotation[4]);
MMDGPU_API mmdgpu_status mmdgpu_set_morph_pose(mmdgpu_frames_t frames, uint32_t slot, uint32_t morph, float weight);
/* Poser::PrePhysicsPosing / PostPhysicsPosing / Deform (L/motion/poser_impl.inl:362-394, 396-461). */
MMDGPU_API mmdgpu_status mmdgpu_pre_physics_posing(mmdgpu_frames_t frames);
MMDGPU_API mmdgpu_status mmdgpu_post_physics_posing(mmdgpu_frames_t frames);
MMDGPU_API mmdgpu_status mmdgpu_deform(mmdgpu_frames_t frames);
/* Physics hand-back (PoserMotionState::Synchronize / Fix, mmd-bullet_impl.inl:34-56): overwrite one
 * bone's skinning matrix (and, if local_or_null != NULL, its local matrix) between pre and post. */
MMDGPU_API mmdgpu_status mmdgpu_set_skinning_matrix_override(mmdgpu_frames_t frames, uint32_t slot, uint32_t bone,
                                                             const float skinning[16], const float* local_or_null);
/* Fused ResetPosing + SeekFrame + PrePhysicsPosing + PostPhysicsPosing + Deform for every slot
 * (main.cpp:1788-1821 with physics off). */
MMDGPU_API mmdgpu_status mmdgpu_update(mmdgpu_frames_t frames, const mmdgpu_animation_t* per_instance,
                                       const uint32_t* frame_per_slot);
MMDGPU_API mmdgpu_status mmdgpu_update_range(mmdgpu_frames_t frames, const mmdgpu_animation_t* per_instance,
                                             const uint32_t* first_frame_per_instance, uint32_t frame_stride);

/* Device pointer of slot 0 of one output stream and the byte stride between slots.  The vertex streams
 * (POSITION / NORMAL / INTERLEAVED / UV) keep their address for the life of the frames object (unless rebound with
 * mmdgpu_frames_bind_output); SKIN_MATRIX belongs
 * to the per-update state, of which fused updates rotate several copies: ask again after every mmdgpu_update. */
MMDGPU_API mmdgpu_status mmdgpu_frames_device_ptr(mmdgpu_frames_t frames, mmdgpu_stream_id id, void** dptr,
                                                  size_t* slot_stride_bytes);
/* Copy one slot's output stream to host memory (synchronous); feeds sg_update_buffer (main.cpp:862)
 * or Poser::pose_image unchanged.  bytes must equal the stream's size for one slot. */
MMDGPU_API mmdgpu_status mmdgpu_frames_download(mmdgpu_frames_t frames, uint32_t slot, mmdgpu_stream_id id,
                                                void* host_dst, size_t bytes);
/* Asynchronous variant on the context's download stream for a slot range into pinned host memory.  The copy
 * starts when the work queued on the compute stream so far has finished and overlaps whatever is queued afterwards;
 * the next skinning launch on THIS frames object waits (on the device) until the copy has read the vertex buffers,
 * so updating the same object right away is safe - sampling and hierarchy of that update still overlap the copy.
 * Completion on the host: mmdgpu_frames_wait_downloads (this object's copies only) or mmdgpu_context_synchronize. */
MMDGPU_API mmdgpu_status mmdgpu_frames_download_async(mmdgpu_frames_t frames, uint32_t first_slot, uint32_t n_slots,
                                                      mmdgpu_stream_id id, void* pinned_host_dst, size_t bytes);
/* Positions then normals of ONE slot into one pinned block of 2 x n_vertices x 12 bytes (Poser::pose_image.coordinates and
 * .normals, L/motion/poser.inl:17-20, behind each other).  A one-slot frames object - an interactive Poser - keeps the
 * two planes adjacent on the device, so this is a single transfer.  Same ordering rules as mmdgpu_frames_download_async. */
MMDGPU_API mmdgpu_status mmdgpu_frames_download_pair_async(mmdgpu_frames_t frames, uint32_t slot, void* pinned_host_dst,
                                                           size_t bytes);
/* Block the calling thread until every mmdgpu_frames_download_async issued for this frames object has landed in host
 * memory.  Does not wait for compute or for copies of other frames objects: a bake alternating two frames objects
 * hands window k to its sink while window k+1 is being evaluated (simple_mmd_renderer_b200/shard.py BakeDriver). */
MMDGPU_API mmdgpu_status mmdgpu_frames_wait_downloads(mmdgpu_frames_t frames);
/* Non-blocking form: 1 if this object's asynchronous downloads have all landed (or none was issued), 0 if one is
 * still in flight, -1 on a CUDA error. */
MMDGPU_API int           mmdgpu_frames_downloads_done(mmdgpu_frames_t frames);
/* Zero-copy hand-off (replaces the CPU repack + sg_update_buffer of main.cpp:820-863): let the skinning kernel write
 * one vertex output stream straight into caller-owned DEVICE memory - e.g. the pointer cudaGraphicsResourceGetMappedPointer
 * returns for sokol's GL vertex buffer (sg_gl_query_buffer_info, 3rd_party/sokol/sokol_gfx.h:5213) - or
 * into PAGE-LOCKED HOST memory (mmdgpu_host_alloc / cudaHostAlloc / cudaHostRegister): the kernel's stores then cross
 * PCIe while it computes and Poser::pose_image (L/motion/poser.inl:17-20) needs no copy of its own, only
 * mmdgpu_frames_wait_skinning before the host reads it (what mmdgpu::Poser of mmdgpu.hpp does).
 * id: POSITION / NORMAL (SoA frames), INTERLEAVED (sokol32 frames: main.cpp:50-54 records, positions x0.1, static UV),
 * UV (extensions).  Slot s is written at device_ptr + s * slot_stride_bytes; exactly n_vertices records per slot are
 * stored (no padding).  device_ptr and slot_stride_bytes must be 16-byte aligned (32 for INTERLEAVED).
 * device_ptr = NULL restores the library-owned buffer.  The binding applies to launches issued after the call; the
 * caller orders its own use of the buffer against the context's stream (mmdgpu_context_stream / _synchronize).
 * mmdgpu_frames_device_ptr / _download[_async] follow the binding. */
MMDGPU_API mmdgpu_status mmdgpu_frames_bind_output(mmdgpu_frames_t frames, mmdgpu_stream_id id, void* device_ptr,
                                                   size_t slot_stride_bytes);
/* Block the calling thread until the most recent skinning launch of this frames object (mmdgpu_deform or a fused update)
 * has finished: its bound or library-owned outputs are then complete and, when bound to page-locked host memory, visible
 * to the host.  Does not wait for other frames objects or for downloads. */
MMDGPU_API mmdgpu_status mmdgpu_frames_wait_skinning(mmdgpu_frames_t frames);
/* BoneImage::skinning_matrix_ / local_matrix_ of one slot as nb x 16 floats (parity on bone globals). */
MMDGPU_API mmdgpu_status mmdgpu_bone_matrices_download(mmdgpu_frames_t frames, uint32_t slot, float* host_dst);
MMDGPU_API mmdgpu_status mmdgpu_bone_local_matrices_download(mmdgpu_frames_t frames, uint32_t slot, float* host_dst);
/* Sampled BoneImage::rotation_/translation_ (nb x 7 floats: T xyz, R xyzw) and morph_rates_ (nm). */
MMDGPU_API mmdgpu_status mmdgpu_bone_poses_download(mmdgpu_frames_t frames, uint32_t slot, float* host_dst);
MMDGPU_API mmdgpu_status mmdgpu_morph_rates_download(mmdgpu_frames_t frames, uint32_t slot, float* host_dst);
/* Poser::material_mul_images_ / material_add_images_ (L/motion/poser.inl:107-161): n_materials x 2 x
 * MMDGPU_MATERIAL_FIELDS floats, per material the multiplicative image then the additive one, fields in the
 * order of mmdgpu_material_morph_entry.value.  libmmd allocates these images as all 1 / all 0 and never fills
 * them (poser_impl.inl:355-358), and that is what libmmd-exact mode returns.  With extensions = 1 (parity
 * unpinned) material morphs are accumulated in application order during pre_physics_posing / update:
 *   MUL entry: mul = mul * (1 + (value - 1) * rate)      ADD entry: add = add + value * rate. */
MMDGPU_API mmdgpu_status mmdgpu_material_images_download(mmdgpu_frames_t frames, uint32_t slot, float* host_dst);

/* Slots one skinning work item walks for its tile (the run over which a tile's static streams stay in registers). */
MMDGPU_API uint32_t      mmdgpu_frames_slot_run(mmdgpu_frames_t frames);

/* Peer buffers: the receive side of the optional bake gather (BASELINE configs[4]).  The root rank creates a
 * device buffer and gets a 64-byte handle (cudaIpcMemHandle_t) to hand to the other processes of the node by any
 * means; each of them opens it and binds a window of it with mmdgpu_frames_bind_output, so that its skinning kernel
 * stores the baked vertices straight into the root's memory over NVLink - compute and gather in one kernel.  Ordering
 * between producers and the consumer is the caller's (one small collective per window; simple_mmd_renderer_b200/shard.py).
 * release: opened = 1 for a pointer from _open, 0 for one from _create. */
MMDGPU_API mmdgpu_status mmdgpu_peer_buffer_create(mmdgpu_context_t ctx, size_t bytes, void** dptr, unsigned char handle[64]);
MMDGPU_API mmdgpu_status mmdgpu_peer_buffer_open(mmdgpu_context_t ctx, const unsigned char handle[64], void** dptr);
MMDGPU_API mmdgpu_status mmdgpu_peer_buffer_release(mmdgpu_context_t ctx, void* dptr, int opened);

/* Test export (function-level known answers): run ONE device math function of the path on n rows
 * of inputs.  op / row layouts: 0 Bezier {4 control bytes as floats, x} -> lambda (host-built table + device lookup,
 * L/util/math_impl.inl:1372-1428); 1 NLerp {a4, b4, l} -> 4 (:1265-1277); 2 SLerp {a4, b4, l} -> 4 (:1312-1340);
 * 3 quaternion -> Euler {q4, order 0 YZX / 1 ZXY / 2 XYZ} -> 3 (:1059-1137); 4 Euler -> quaternion {e3, order} -> 4
 * (:1156-1224); 5 AxisToQuaternion {axis3, angle} -> 4 (:1047-1058); 6 quaternion product {a4, b4} -> 4 (:510-517);
 * 7 ToRotateMatrix {q4} -> rows 0..2 (9) (:540-563); 8 Inverse {q4} -> 4 (:474-477); 9 affine 4x4 product {a16, b16}
 * -> 16 (:984-1003); 10 Vector3D::Normalize {v3} -> 3 (:393-400). */
MMDGPU_API mmdgpu_status mmdgpu_test_math(mmdgpu_context_t ctx, int op, const float* in, uint32_t n, float* out);

/* Pinned host memory helpers for the download path. */
MMDGPU_API mmdgpu_status mmdgpu_host_alloc(size_t bytes, void** out);
MMDGPU_API void          mmdgpu_host_free(void* p);

/* ------------------------------------------------- host plan (no GPU used) */

/*
 * The flattened, device-ready image of a model, built on the host only.  It is what
 * mmdgpu_model_create_from_arrays uploads; it is exposed so that the index tier (Normalize rewrite,
 * evaluation order, wave schedule, IK classification, morph application slots, per-vertex CSR) can be
 * checked bit-exactly on a machine without a GPU.
 */
MMDGPU_API mmdgpu_status mmdgpu_plan_create(const mmdgpu_model_desc* desc, const mmdgpu_options* opt_or_null,
                                            mmdgpu_plan_t* out, char* err_buf, size_t err_buf_len);
MMDGPU_API void          mmdgpu_plan_destroy(mmdgpu_plan_t plan);

typedef enum mmdgpu_plan_array {
    MMDGPU_PLAN_SKIN_TYPE = 0,      /* u8  [nv]   type after Model::Normalize + Lerp shortcuts        */
    MMDGPU_PLAN_BONE_ID = 1,        /* u16 [4nv]  ids after the rewrite                               */
    MMDGPU_PLAN_WEIGHT = 2,         /* f32 [4nv]                                                      */
    MMDGPU_PLAN_ORDER_PRE = 3,      /* i32 [..]   pre_physics_bones_ after std::sort                  */
    MMDGPU_PLAN_ORDER_POST = 4,     /* i32 [..]   post_physics_bones_ after std::sort                 */
    MMDGPU_PLAN_OP_KIND = 5,        /* u8  [n_ops] 0 EVAL, 1 IK, 2 SKIN, in program order             */
    MMDGPU_PLAN_OP_BONE = 6,        /* i32 [n_ops]                                                    */
    MMDGPU_PLAN_OP_WAVE = 7,        /* i32 [n_ops] wave each op was scheduled into                    */
    MMDGPU_PLAN_WAVE_BEGIN = 8,     /* i32 [n_waves+1] offsets into the wave-ordered op list           */
    MMDGPU_PLAN_WAVE_OPS = 9,       /* i32 [n_ops] op indices (program order ids) grouped by wave      */
    MMDGPU_PLAN_IK_FIX_TYPE = 10,   /* u8  [n_ik_links] 0 NONE 1 X 2 Y 3 Z 4 ALL                       */
    MMDGPU_PLAN_IK_EULER_ORDER = 11,/* u8  [n_ik_links] 0 YZX 1 ZXY 2 XYZ                              */
    MMDGPU_PLAN_APP_SLOT_MORPH = 12,/* i32 [n_app_slots] morph index of every DFS application slot     */
    MMDGPU_PLAN_APP_SLOT_PARENT = 13,/* i32 [n_app_slots] parent slot (group) or -1                    */
    MMDGPU_PLAN_APP_SLOT_MULT = 14, /* f32 [n_app_slots] group rate multiplier                         */
    MMDGPU_PLAN_CSR_ROW_PTR = 15,   /* u32 [nv+1]                                                     */
    MMDGPU_PLAN_CSR_SLOT = 16,      /* u32 [n_csr] application slot of every CSR entry                 */
    MMDGPU_PLAN_CSR_OFFSET = 17,    /* f32 [3 n_csr]                                                  */
    MMDGPU_PLAN_BEZIER_UNUSED = 18,
    MMDGPU_PLAN_WAVE_PHASE_SPLIT = 19,/* i32 [1] index of the first wave of the post-physics segment   */
    /* device vertex layout: 512-vertex tiles stored in a tile-local order (type, morph entry count, index) */
    MMDGPU_PLAN_TILE_ORIG = 20,     /* u16 [nv_pad] storage position -> PMX vertex index within its tile     */
    MMDGPU_PLAN_TILE_TYPE = 21,     /* u8  [nv_pad] device skinning type per storage position (0 B1 1 B2 2 B4 3 SDEF 4 QDEF) */
    MMDGPU_PLAN_TILE_LOCAL_ID = 22, /* u16 [4 nv_pad] tile-local bone index per storage position           */
    MMDGPU_PLAN_TILE_BONE_BEGIN = 23,/* u32 [n_tiles+1] offsets into TILE_BONES                             */
    MMDGPU_PLAN_TILE_BONES = 24,    /* u16 [..] distinct bone ids used by each tile, ascending              */
    MMDGPU_PLAN_ELL_BASE = 25,      /* u32 [16 n_tiles] first entry of every 32-lane group (16 per tile)  */
    MMDGPU_PLAN_ELL_ROUNDS = 26,    /* u32 [16 n_tiles] entries per lane (padded) of every group            */
    MMDGPU_PLAN_ELL_SLOT = 27,      /* u32 [n_ell] application slot; padding = number of application slots  */
    MMDGPU_PLAN_ELL_OFFSET = 28,    /* f32 [3 n_ell]                                                        */
    /* model data as the plan holds it (lets a PMX byte stream be checked against flat arrays on the host) */
    MMDGPU_PLAN_POSITION = 29,      /* f32 [3 nv] */
    MMDGPU_PLAN_NORMAL = 30,        /* f32 [3 nv] */
    MMDGPU_PLAN_UV = 31,            /* f32 [2 nv] */
    MMDGPU_PLAN_BONE_STATIC = 32,   /* u8  [48 nb] per-bone record read by the hierarchy kernel (host_plan.hpp) */
    MMDGPU_PLAN_IK_DESC = 33,       /* u8  [32 n_ik] */
    MMDGPU_PLAN_IK_LINK = 34,       /* u8  [32 n_ik_links] */
    MMDGPU_PLAN_BONE_MORPH = 35,    /* u8  [32 n] bone-morph entries grouped by bone, application order        */
    /* extensions only (empty otherwise): material morphs grouped by material, application order inside a material;
     * an entry for "every material" is repeated under each */
    MMDGPU_PLAN_MATERIAL_MORPH_ROW = 36, /* i32 [n_materials + 1]                                              */
    MMDGPU_PLAN_MATERIAL_MORPH = 37,     /* u8  [120 n] {i32 application slot; u32 method; f32 value[28]}      */
    /* chain-local images of the CCD IK solves (device design, one image per IK bone in PLAN_IK_DESC order): the bones a
     * solve touches renumbered 0..n-1, with translated copies of their static records, IK descriptor and links */
    MMDGPU_PLAN_IK_IMAGE = 38,             /* u8  [32 n] {i32 bones_begin, n_bones, lslots_begin, n_lslots, mslots_begin, n_mslots, region_f4, pad} */
    MMDGPU_PLAN_IK_IMAGE_BONES = 39,       /* i32 global bone id of every image bone                           */
    MMDGPU_PLAN_IK_IMAGE_WRITTEN = 40,     /* u8  1: the solve evaluates this image bone (links, target)        */
    MMDGPU_PLAN_IK_IMAGE_STATIC = 41,      /* u8  [48 n] PLAN_BONE_STATIC records with image-local references   */
    MMDGPU_PLAN_IK_IMAGE_LINK_SLOTS = 42,  /* i32 global link slot of every image link slot                    */
    MMDGPU_PLAN_IK_IMAGE_MORPH_SLOTS = 43, /* i32 global bone-morph slot of every image morph slot             */
    MMDGPU_PLAN_IK_IMAGE_DESC = 44,        /* u8  [32 n] PLAN_IK_DESC records with image-local bone / target    */
    MMDGPU_PLAN_IK_IMAGE_LINKS = 45        /* u8  [32 n] PLAN_IK_LINK records with image-local bone             */
} mmdgpu_plan_array;

/* Host-only: parse a PMX 2.0 / 2.1 byte stream (layout of L/reader/pmx_reader_impl.inl:16-449) and build the
 * plan; bone / morph names are kept for VMD joins. */
MMDGPU_API mmdgpu_status mmdgpu_plan_create_from_pmx(const void* bytes, size_t n, const mmdgpu_options* opt_or_null,
                                                     mmdgpu_plan_t* out, char* err_buf, size_t err_buf_len);

/* Returns a pointer into plan-owned memory and the element count; the pointer is valid until the plan
 * (or the model that owns it) is destroyed. */
MMDGPU_API mmdgpu_status mmdgpu_plan_get(mmdgpu_plan_t plan, mmdgpu_plan_array which, const void** data,
                                         size_t* count);

/* Host-only flattened motion: what mmdgpu_animation_create_* uploads (mmd::Motion's std::map storage,
 * L/motion/motion.inl:128-129, as sorted, de-duplicated key arrays per model bone / morph plus the
 * de-duplicated Bezier tables). */
typedef struct mmdgpu_anim_plan* mmdgpu_anim_plan_t;
MMDGPU_API mmdgpu_status mmdgpu_anim_plan_create(const mmdgpu_anim_desc* desc, uint32_t n_bones, uint32_t n_morphs,
                                                 mmdgpu_anim_plan_t* out, char* err_buf, size_t err_buf_len);
/* VMD byte stream joined by name against a plan that was created from PMX bytes. */
MMDGPU_API mmdgpu_status mmdgpu_anim_plan_create_from_vmd(mmdgpu_plan_t model_plan, const void* bytes, size_t n,
                                                          mmdgpu_anim_plan_t* out, char* err_buf, size_t err_buf_len);
MMDGPU_API void          mmdgpu_anim_plan_destroy(mmdgpu_anim_plan_t plan);
typedef enum mmdgpu_anim_array {
    MMDGPU_ANIM_BONE_KEY_BEGIN = 0,  /* u32 [nb]  first key of every model bone                   */
    MMDGPU_ANIM_BONE_KEY_COUNT = 1,  /* u32 [nb]                                                  */
    MMDGPU_ANIM_BONE_TRACKED = 2,    /* u8  [nb]  1 if the motion registers the bone              */
    MMDGPU_ANIM_KEY_FRAME = 3,       /* u32 [nk]  ascending inside a track, duplicates removed    */
    MMDGPU_ANIM_KEY_T = 4,           /* f32 [4 nk] translation (w unused)                         */
    MMDGPU_ANIM_KEY_R = 5,           /* f32 [4 nk] rotation xyzw                                  */
    MMDGPU_ANIM_KEY_CURVE = 6,       /* u32 [4 nk] Bezier table index of X, Y, Z, R; 0xFFFFFFFF = linear */
    MMDGPU_ANIM_TABLES = 7,          /* f32 [32 nt]                                               */
    MMDGPU_ANIM_MORPH_KEY_BEGIN = 8, /* u32 [nm] */
    MMDGPU_ANIM_MORPH_KEY_COUNT = 9, /* u32 [nm] */
    MMDGPU_ANIM_MORPH_TRACKED = 10,  /* u8  [nm] */
    MMDGPU_ANIM_MKEY_FRAME = 11,     /* u32 [nmk] */
    MMDGPU_ANIM_MKEY_WEIGHT = 12     /* f32 [nmk] */
} mmdgpu_anim_array;
MMDGPU_API mmdgpu_status mmdgpu_anim_plan_get(mmdgpu_anim_plan_t plan, mmdgpu_anim_array which, const void** data,
                                              size_t* count);

/* Presampled Bezier table of Bezier::presample (L/util/math_impl.inl:1398-1428) for one VMD control
 * quadruple (x0, y0, x1, y1).  Returns 1 and leaves table untouched if the curve is linear. */
MMDGPU_API int mmdgpu_bezier_table(const int8_t ctrl[4], float table[32]);

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* MMDGPU_H_INCLUDED */
