// mmdgpu_api.cu — the C-ABI of include/mmdgpu.h: handles, uploads, launches, downloads.
//
// There is no CPU path in this file: every entry point that produces poses, matrices or vertices launches
// the sm_100a kernels of kernels.cu; when CUDA is unavailable the call returns MMDGPU_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "../../include/mmdgpu.h"
#include "device_types.cuh"
#include "host_plan.hpp"
#include "kernels.cuh"

using namespace mmdgpu;

// ------------------------------------------------------------------------------------------- handles
constexpr int kPreStreams = 8;                   // hierarchy chains of fused updates in flight
constexpr int kStateCopies = kPreStreams + 1;    // copies of the per-update state

struct mmdgpu_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t dl_stream = nullptr;
    cudaEvent_t dl_event = nullptr;
    // fused updates run key-frame sampling and the bone hierarchy of update n+1 on this stream while the skinning
    // kernel of update n still runs on `stream`
    // (kPreStreams of them, taken in turn, so that the latency-bound hierarchy chains of consecutive updates overlap each
    // other: a batch of CCD IK solves is one dependent chain of ~0.2 ms that occupies a few hundred threads)
    cudaStream_t pre_stream[kPreStreams] = {};
    // same, highest priority: for models with CCD IK, whose hierarchy kernel is a long latency-bound chain that
    // should claim SM resources as soon as CTAs of the running skinning kernel retire
    cudaStream_t pre_stream_hi[kPreStreams] = {};
    int n_pre_ik = kPreStreams;                   // streams a model with CCD IK uses (MMDGPU_PRE_STREAMS=1..8, experiment knob)
    int n_pre_plain = 2;                          // ... and a model without: its hierarchy is short, deeper buys nothing
    cudaError_t sync_pre() {
        for (int i = 0; i < kPreStreams; ++i)
            for (cudaStream_t st : {pre_stream[i], pre_stream_hi[i]})
                if (st) { cudaError_t e = cudaStreamSynchronize(st); if (e != cudaSuccess) return e; }
        return cudaSuccess;
    }
    std::string err;
    uint64_t launches = 0;
    int max_smem_optin = 0;
    int sm_count = 0;
    // optional per-kernel device timing
    bool profiling = false;
    struct Span { int id; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
};

struct mmdgpu_plan {
    Plan plan;
};
struct mmdgpu_anim_plan {
    HostAnim anim;
};

namespace {

// Owns device allocations of one handle.
struct DevArena {
    std::vector<void*> ptrs;
    ~DevArena() { release(); }
    void release() {
        for (void* p : ptrs) cudaFree(p);
        ptrs.clear();
    }
    cudaError_t alloc(void** out, size_t bytes) {
        *out = nullptr;
        cudaError_t e = cudaMalloc(out, bytes ? bytes : 16);
        if (e == cudaSuccess) ptrs.push_back(*out);
        return e;
    }
};

thread_local std::string g_err;  // errors of calls that have no context to attach them to

}  // namespace

struct mmdgpu_model {
    mmdgpu_context_t ctx = nullptr;
    mmdgpu_plan plan;
    DevModel dev{};
    DevArena mem;
    std::vector<std::pair<uint32_t, uint32_t>> ik_waves;  // (wave, ops in it) of every wave that holds a CCD IK solve
    bool ik_images_ok = false;                            // every solve has a chain-local image that fits
};

struct mmdgpu_animation {
    mmdgpu_context_t ctx = nullptr;
    mmdgpu_model_t model = nullptr;
    // Identity of this clip for the frames objects' "what is uploaded" caches.  Never the handle address: a clip created
    // after another was destroyed usually gets the same heap block back, and its DevAnim points at different arrays.
    uint64_t uid = 0;
    HostAnim host;
    DevAnim dev{};
    DevArena mem;
};


struct mmdgpu_frames {
    mmdgpu_context_t ctx = nullptr;
    mmdgpu_model_t model = nullptr;
    mmdgpu_layout layout = MMDGPU_LAYOUT_SOA_POS_NRM;
    DevFrames dev{};
    DevArena mem;
    DevAnim* d_anims = nullptr;                 // [n_instances], of the selected copy
    std::vector<DevAnim> h_anims;
    bool range_mode = false;
    uint32_t frame_stride = 1;
    uint32_t slots_per_cta = 1;
    // Everything one fused update writes before its skinning kernel runs - sampled poses and rates, the hierarchy state
    // that crosses launches, and what the hierarchy hands to the skinning kernel (palette, extension palette,
    // application-slot rates) - exists kStateCopies = kPreStreams + 1 times, used round-robin: updates n+1 .. n+kPreStreams
    // sample and run their hierarchies (on the pre streams, in turn) while update n's skinning kernel still reads its copy.
    struct StateSet {
        float4 *poseR = nullptr, *poseT = nullptr, *totR = nullptr, *totT = nullptr, *ikR = nullptr, *preIK = nullptr,
               *morphR = nullptr, *morphT = nullptr, *palette = nullptr, *pal_ext = nullptr;
        float *rate = nullptr, *local = nullptr, *node_rate = nullptr, *material_images = nullptr;
        uint32_t* frame_id = nullptr;
        double* time_s = nullptr;
        DevAnim* d_anims = nullptr;
        std::vector<uint64_t> bound;            // uid of the clip whose DevAnim each entry of d_anims currently holds
    } set[kStateCopies];
    int n_copies = kStateCopies;                  // copies this object allocates and rotates: pre streams of its model + 1
    int cur = 0;                                  // copy the step-wise entry points and the downloads use
    int update_turn = 0;                          // which of the pre streams the last fused update used
    cudaEvent_t ev_pre[kStateCopies] = {};    // hierarchy of the update that wrote copy i has finished
    cudaEvent_t ev_skin[kStateCopies] = {};   // skinning that read copy i has finished
    bool skin_recorded[kStateCopies] = {};
    cudaEvent_t ev_main = nullptr;                // main-stream work the next fused update must follow
    bool main_dirty = true;
    // The vertex output streams are single-buffered: a skinning kernel must not overwrite them while an asynchronous
    // download of the previous update is still reading.  ev_dl follows the last copy issued on the download stream.
    bool host_bound = false;                      // some vertex output stream is bound to page-locked host memory
    cudaEvent_t ev_deform = nullptr;              // follows the last step-wise mmdgpu_deform
    cudaEvent_t last_skin = nullptr;              // event behind the most recent skinning launch (fused or step-wise)
    cudaEvent_t ev_dl = nullptr;
    bool dl_pending = false;                      // a skinning launch has yet to wait for ev_dl
    bool dl_recorded = false;                     // ev_dl has been recorded at least once
    // library-owned output buffers (what mmdgpu_frames_bind_output(.., NULL, ..) restores)
    float *own_pos = nullptr, *own_nrm = nullptr;
    float4* own_inter = nullptr;
    float2* own_uv = nullptr;
    void select(int i) {
        const StateSet& x = set[i];
        dev.poseR = x.poseR; dev.poseT = x.poseT; dev.rate = x.rate; dev.totR = x.totR; dev.totT = x.totT; dev.local = x.local;
        dev.ikR = x.ikR; dev.preIK = x.preIK; dev.morphR = x.morphR; dev.morphT = x.morphT;
        dev.palette = x.palette; dev.pal_ext = x.pal_ext; dev.node_rate = x.node_rate; dev.material_images = x.material_images;
        dev.frame_id = x.frame_id; dev.time_s = x.time_s;
        d_anims = x.d_anims;
        cur = i;
    }
    ~mmdgpu_frames() {
        // also the error path of mmdgpu_frames_create: queued launches may still reference the arena freed after this body
        if (ctx && ctx->stream) cudaStreamSynchronize(ctx->stream);
        for (int i = 0; i < kStateCopies; ++i) {
            if (ev_pre[i]) cudaEventDestroy(ev_pre[i]);
            if (ev_skin[i]) cudaEventDestroy(ev_skin[i]);
        }
        if (ev_main) cudaEventDestroy(ev_main);
        if (ev_dl) cudaEventDestroy(ev_dl);
        if (ev_deform) cudaEventDestroy(ev_deform);
    }
};

namespace {

mmdgpu_status set_err(mmdgpu_context_t ctx, mmdgpu_status s, const std::string& msg) {
    if (ctx) ctx->err = msg; else g_err = msg;
    return s;
}
mmdgpu_status cuda_fail(mmdgpu_context_t ctx, cudaError_t e, const char* what) {
    const mmdgpu_status s = (e == cudaErrorMemoryAllocation) ? MMDGPU_ERR_OOM : MMDGPU_ERR_CUDA;
    return set_err(ctx, s, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CU(ctx, call)                                             \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
    } while (0)

template <class T>
cudaError_t upload(mmdgpu_context_t ctx, DevArena& mem, const T* src, size_t n, const T** out) {
    void* d = nullptr;
    cudaError_t e = mem.alloc(&d, n * sizeof(T));
    if (e != cudaSuccess) return e;
    if (n) e = cudaMemcpyAsync(d, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
    *out = static_cast<const T*>(d);
    return e;
}
template <class T>
cudaError_t upload(mmdgpu_context_t ctx, DevArena& mem, const std::vector<T>& v, const T** out) {
    return upload(ctx, mem, v.data(), v.size(), out);
}
template <class T>
cudaError_t dalloc(DevArena& mem, T** out, size_t n, bool zero, cudaStream_t st) {
    void* d = nullptr;
    cudaError_t e = mem.alloc(&d, n * sizeof(T));
    if (e != cudaSuccess) return e;
    if (zero && n) e = cudaMemsetAsync(d, 0, n * sizeof(T), st);
    *out = static_cast<T*>(d);
    return e;
}

uint32_t round_up(uint32_t x, uint32_t m) { return (x + m - 1) / m * m; }
constexpr uint32_t kMaxStagedTileBones = 256;

// How many consecutive slots one CTA walks for its tile (a multiple of the slot group).  Long runs amortise the
// per-item prologue (static streams, first palette); enough items must remain to fill the machine.  Measured on
// B200 (profiles/r01_experiments.md): 64-slot runs are best whenever they still leave >= ~1.5 waves of work items.
// Slot strides of the library-owned output planes: exactly nv records, rounded up to 16 bytes (bulk copies).
// MMDGPU_PADDED_SLOTS=1 (A/B knob): slots a whole number of 512-vertex tiles apart, as in round 1.
inline bool padded_slots() {
    static const bool on = [] { const char* e = std::getenv("MMDGPU_PADDED_SLOTS"); return e && e[0] == '1'; }();
    return on;
}
inline size_t tile_pad(uint32_t nv) { return (size_t(nv) + kTileVerts - 1) / kTileVerts * kTileVerts; }
inline size_t own_plane_stride(uint32_t nv) { return padded_slots() ? tile_pad(nv) * 3 : (size_t(nv) * 3 + 3) / 4 * 4; }   // floats
inline size_t own_inter_stride(uint32_t nv) { return padded_slots() ? tile_pad(nv) * 2 : size_t(nv) * 2; }                // float4
inline size_t own_uv_stride(uint32_t nv) { return padded_slots() ? tile_pad(nv) : (size_t(nv) + 1) / 2 * 2; }             // float2

uint32_t choose_slots_per_cta(uint32_t tiles, uint32_t n_slots, int sm_count) {
    if (tiles == 0 || n_slots == 0) return kSlotGroup;
    if (const char* env = std::getenv("MMDGPU_SLOTS_PER_CTA")) {  // tuning knob for experiments
        const long v = std::strtol(env, nullptr, 10);
        if (v > 0) return uint32_t(std::min<long>(v, n_slots));
    }
    const uint64_t min_items = uint64_t(sm_count > 0 ? sm_count : 148) * 3 * 3 / 2;  // 1.5 waves of 3 CTAs per SM
    uint32_t chunk = std::min<uint32_t>(64, (n_slots + kSlotGroup - 1) / kSlotGroup * kSlotGroup);
    while (chunk > kSlotGroup && uint64_t(tiles) * ((n_slots + chunk - 1) / chunk) < min_items) chunk = std::max(kSlotGroup, chunk / 2);
    return (chunk + kSlotGroup - 1) / kSlotGroup * kSlotGroup;
}

mmdgpu_status upload_model(mmdgpu_model* m) {
    mmdgpu_context_t ctx = m->ctx;
    const Plan& p = m->plan.plan;
    DevModel& D = m->dev;
    D = DevModel{};
    const uint32_t nv = p.nv, nb = p.nb;
    const uint32_t nvp = p.nv_pad;
    D.nv = nv; D.nv_pad = nvp; D.nb = nb; D.nm = p.nm;
    D.n_nodes = uint32_t(p.node_morph.size());
    D.n_nodes_pad = round_up(D.n_nodes + 1, 4);  // + the always-zero slot the ELL padding points at
    D.n_tiles = p.n_tiles;
    D.max_tile_bones = std::max<uint32_t>(1, p.max_tile_bones);
    D.extensions = p.extensions ? 1u : 0u;

    // A tile that touches more bones than fit the staged palettes (4 slots x 2 buffers x 48 B per bone next to the
    // staging tiles) switches the whole model to global bone ids read straight from the slot's palette.
    // MMDGPU_FORCE_FALLBACKS=1 (test knob, tools/gpu_fuzz.py): take the large-model code paths on small models
    const bool force = std::getenv("MMDGPU_FORCE_FALLBACKS") != nullptr && !p.extensions;
    D.global_palette = (!p.extensions && (p.max_tile_bones > kMaxStagedTileBones || force)) ? 1u : 0u;
    if (p.extensions && p.max_tile_bones > kMaxStagedTileBones)
        return set_err(ctx, MMDGPU_ERR_UNSUPPORTED, "extensions with more than " + std::to_string(kMaxStagedTileBones) +
                                                        " distinct bones in one 512-vertex tile");
    if (D.global_palette) D.max_tile_bones = 0;

    // ---- vertex streams, structure of arrays, in tile storage order (host_plan.hpp)
    std::vector<float> plane[6];
    for (auto& v : plane) v.assign(nvp, 0.0f);
    std::vector<uint2> ids(nvp, make_uint2(0, 0));
    std::vector<float4> wts(nvp, make_float4(0.f, 0.f, 0.f, 0.f));
    std::vector<float2> uv(nvp, make_float2(0.f, 0.f));

    for (uint32_t pos = 0; pos < nvp; ++pos) {
        const uint32_t src = (pos / kTileVerts) * kTileVerts + p.tile_orig[pos];
        if (src < nv) {
            for (int k = 0; k < 3; ++k) {
                plane[k][pos] = p.position[size_t(src) * 3 + k];
                plane[3 + k][pos] = p.normal[size_t(src) * 3 + k];
            }
            uv[pos] = make_float2(p.uv[size_t(src) * 2], p.uv[size_t(src) * 2 + 1]);
        }
        uint16_t gid[4] = {0, 0, 0, 0};
        if (D.global_palette) {
            const uint32_t tb0 = p.tile_bone_begin[pos / kTileVerts];
            for (int k = 0; k < 4; ++k) gid[k] = p.tile_bones[tb0 + p.st_local_id[size_t(pos) * 4 + k]];
        }
        const uint16_t* id = D.global_palette ? gid : &p.st_local_id[size_t(pos) * 4];
        ids[pos].x = uint32_t(id[0]) | (uint32_t(p.st_type[pos]) << 13) | (uint32_t(id[1]) << 16);
        ids[pos].y = uint32_t(id[2]) | (uint32_t(id[3]) << 16);
        const float* w = &p.st_weight[size_t(pos) * 4];
        wts[pos] = make_float4(w[0], w[1], w[2], w[3]);
    }
    CU(ctx, upload(ctx, m->mem, plane[0], &D.px)); CU(ctx, upload(ctx, m->mem, plane[1], &D.py));
    CU(ctx, upload(ctx, m->mem, plane[2], &D.pz)); CU(ctx, upload(ctx, m->mem, plane[3], &D.nx));
    CU(ctx, upload(ctx, m->mem, plane[4], &D.ny)); CU(ctx, upload(ctx, m->mem, plane[5], &D.nz));
    CU(ctx, upload(ctx, m->mem, ids, &D.ids));
    CU(ctx, upload(ctx, m->mem, wts, &D.weights));
    CU(ctx, upload(ctx, m->mem, uv, &D.uv));
    CU(ctx, upload(ctx, m->mem, p.tile_orig, &D.orig));
    // ---- sliced-ELL morph entries
    std::vector<uint2> hdr(p.ell_base.size());
    for (size_t g = 0; g < hdr.size(); ++g) hdr[g] = make_uint2(p.ell_base[g], p.ell_rounds[g]);
    std::vector<float4> ent(p.ell_node.size());
    for (size_t e = 0; e < ent.size(); ++e) {
        float slot_bits;
        const uint32_t node = p.ell_node[e] * 16u;  // byte offset of the node's float4 (one rate per slot of a group)
        std::memcpy(&slot_bits, &node, 4);
        ent[e] = make_float4(p.ell_offset[3 * e], p.ell_offset[3 * e + 1], p.ell_offset[3 * e + 2], slot_bits);
    }
    CU(ctx, upload(ctx, m->mem, hdr, &D.ell_hdr));
    CU(ctx, upload(ctx, m->mem, ent, &D.ell_ent));
    CU(ctx, upload(ctx, m->mem, p.tile_bone_begin, &D.tile_bone_begin));
    CU(ctx, upload(ctx, m->mem, p.tile_bones, &D.tile_bones));
    if (p.extensions) {
        // spherical-deform parameters (3 float4 per storage position) and the UV-morph table
        std::vector<float4> sd(size_t(nvp) * 3, make_float4(0.f, 0.f, 0.f, 0.f));
        if (!p.st_sdef.empty())
            for (size_t i = 0; i < sd.size(); ++i)
                sd[i] = make_float4(p.st_sdef[4 * i], p.st_sdef[4 * i + 1], p.st_sdef[4 * i + 2], 0.f);
        CU(ctx, upload(ctx, m->mem, sd, &D.sdef));
        std::vector<uint2> uhdr(p.uv_ell_base.size());
        for (size_t g = 0; g < uhdr.size(); ++g) uhdr[g] = make_uint2(p.uv_ell_base[g], p.uv_ell_rounds[g]);
        std::vector<float4> uent(p.uv_ell_node.size());
        for (size_t e = 0; e < uent.size(); ++e) {
            float slot_bits;
            const uint32_t node = p.uv_ell_node[e] * 16u;
            std::memcpy(&slot_bits, &node, 4);
            uent[e] = make_float4(p.uv_ell_offset[4 * e], p.uv_ell_offset[4 * e + 1], slot_bits, 0.f);
        }
        CU(ctx, upload(ctx, m->mem, uhdr, &D.uv_ell_hdr));
        CU(ctx, upload(ctx, m->mem, uent, &D.uv_ell_ent));
    }
    // ---- bones and the program
    static_assert(sizeof(BoneStatic) == 48, "BoneStatic is read as three float4");
    CU(ctx, upload(ctx, m->mem, p.bones, &D.bones));
    CU(ctx, upload(ctx, m->mem, p.iks, &D.iks));
    CU(ctx, upload(ctx, m->mem, p.links, &D.links));
    D.ik_nested = p.ik_nested ? 1u : 0u;
    CU(ctx, upload(ctx, m->mem, p.reset_bones, &D.reset_bones));
    D.n_reset = uint32_t(p.reset_bones.size());
    D.n_link_slots = uint32_t(p.link_bones.size());
    D.n_morph_slots = uint32_t(p.morph_bones.size());
    std::vector<uint32_t> wb(p.wave_begin.begin(), p.wave_begin.end());
    std::vector<uint32_t> words(p.wave_ops.size());
    for (size_t i = 0; i < words.size(); ++i) {
        const Op& o = p.ops[size_t(p.wave_ops[i])];
        words[i] = (uint32_t(o.kind) << 28) | (uint32_t(o.arg) & 0x0FFFFFFFu);
    }
    CU(ctx, upload(ctx, m->mem, wb, &D.wave_begin));
    CU(ctx, upload(ctx, m->mem, words, &D.wave_ops));
    D.n_waves = uint32_t(p.wave_begin.size() - 1);
    D.phase_split = uint32_t(p.phase_split);
    D.n_ops = uint32_t(p.wave_ops.size());
    CU(ctx, upload(ctx, m->mem, p.node_morph, &D.node_morph));
    CU(ctx, upload(ctx, m->mem, p.node_parent, &D.node_parent));
    CU(ctx, upload(ctx, m->mem, p.node_mult, &D.node_mult));
    CU(ctx, upload(ctx, m->mem, p.nodes_by_depth, &D.nodes_by_depth));
    CU(ctx, upload(ctx, m->mem, p.depth_begin, &D.depth_begin));
    D.n_depths = p.depth_begin.empty() ? 0u : uint32_t(p.depth_begin.size() - 1);
    CU(ctx, upload(ctx, m->mem, p.bone_morph_row, &D.bone_morph_row));
    CU(ctx, upload(ctx, m->mem, p.bone_morph_entries, &D.bone_morph_entries));
    D.n_materials = p.n_materials;
    if (!p.material_morph_entries.empty()) {
        CU(ctx, upload(ctx, m->mem, p.material_morph_row, &D.material_morph_row));
        CU(ctx, upload(ctx, m->mem, p.material_morph_entries, &D.material_morph_entries));
    }
    // ---- chain-local images of the CCD IK solves (built by build_plan; kernels.cu, hierarchy_flat_kernel)
    {
        m->ik_waves.clear();
        for (uint32_t w = 0; w + 1 < p.wave_begin.size(); ++w) {
            bool has_ik = false;
            for (int32_t i = p.wave_begin[w]; i < p.wave_begin[w + 1]; ++i) has_ik |= p.ops[size_t(p.wave_ops[size_t(i)])].kind == kOpIk;
            if (has_ik) m->ik_waves.push_back({w, uint32_t(p.wave_begin[w + 1] - p.wave_begin[w])});
        }
        // 64 solves per CTA must fit the shared memory a CTA may have next to the skinning kernel's
        m->ik_images_ok = p.ik_img_ok && size_t(p.ik_img_max_region) * 64 * sizeof(float4) <= 96 * 1024;
        if (m->ik_images_ok) {
            CU(ctx, upload(ctx, m->mem, p.ik_img, &D.ik_img));
            CU(ctx, upload(ctx, m->mem, p.ik_img_bones, &D.ik_img_bones));
            CU(ctx, upload(ctx, m->mem, p.ik_img_written, &D.ik_img_written));
            CU(ctx, upload(ctx, m->mem, p.ik_img_static, &D.ik_img_static));
            CU(ctx, upload(ctx, m->mem, p.ik_img_lslots, &D.ik_img_lslots));
            CU(ctx, upload(ctx, m->mem, p.ik_img_mslots, &D.ik_img_mslots));
            CU(ctx, upload(ctx, m->mem, p.ik_img_desc, &D.ik_img_desc));
            CU(ctx, upload(ctx, m->mem, p.ik_img_links, &D.ik_img_links));
            D.ik_img_max_region = p.ik_img_max_region;
        }
    }
    CU(ctx, cudaStreamSynchronize(ctx->stream));  // the staging vectors above die with this frame

    const size_t smem = std::max(skin_smem_bytes(D, MMDGPU_LAYOUT_INTERLEAVED_SOKOL32), skin_smem_bytes(D, MMDGPU_LAYOUT_SOA_POS_NRM));
    if (smem + 1024 > size_t(ctx->max_smem_optin))
        return set_err(ctx, MMDGPU_ERR_UNSUPPORTED,
                       "tile palettes + morph slot rates (" + std::to_string(smem) + " B) exceed shared memory per CTA");
    CU(ctx, prepare_skin_kernels(D));
    return MMDGPU_OK;
}

mmdgpu_status upload_anim(mmdgpu_animation* a) {
    mmdgpu_context_t ctx = a->ctx;
    const HostAnim& h = a->host;
    DevAnim& D = a->dev;
    D = DevAnim{};
    CU(ctx, upload(ctx, a->mem, h.bone_key_begin, &D.bone_key_begin));
    CU(ctx, upload(ctx, a->mem, h.bone_key_count, &D.bone_key_count));
    CU(ctx, upload(ctx, a->mem, h.bone_tracked, &D.bone_tracked));
    CU(ctx, upload(ctx, a->mem, h.key_frame, &D.key_frame));
    CU(ctx, upload(ctx, a->mem, reinterpret_cast<const float4*>(h.key_T.data()), h.key_T.size() / 4, &D.key_T));
    CU(ctx, upload(ctx, a->mem, reinterpret_cast<const float4*>(h.key_R.data()), h.key_R.size() / 4, &D.key_R));
    CU(ctx, upload(ctx, a->mem, reinterpret_cast<const uint4*>(h.key_curve.data()), h.key_curve.size() / 4, &D.key_curve));
    CU(ctx, upload(ctx, a->mem, h.tables, &D.tables));
    CU(ctx, upload(ctx, a->mem, h.morph_key_begin, &D.morph_key_begin));
    CU(ctx, upload(ctx, a->mem, h.morph_key_count, &D.morph_key_count));
    CU(ctx, upload(ctx, a->mem, h.morph_tracked, &D.morph_tracked));
    CU(ctx, upload(ctx, a->mem, h.mkey_frame, &D.mkey_frame));
    CU(ctx, upload(ctx, a->mem, h.mkey_weight, &D.mkey_weight));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return MMDGPU_OK;
}

std::atomic<uint64_t> g_next_anim_uid{1};

// Argument check shared by every entry point that takes clips; nothing is modified on failure.
mmdgpu_status check_anims(const mmdgpu_frames* f, const mmdgpu_animation_t* per_instance) {
    mmdgpu_context_t ctx = f->ctx;
    if (!per_instance) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "per_instance animation array is NULL");
    for (uint32_t i = 0; i < f->dev.n_instances; ++i) {
        if (!per_instance[i]) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "animation handle is NULL");
        if (per_instance[i]->model != f->model)
            return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "animation was created for a different model");
    }
    return MMDGPU_OK;
}

mmdgpu_status bind_anims(mmdgpu_frames* f, const mmdgpu_animation_t* per_instance, cudaStream_t st) {
    mmdgpu_context_t ctx = f->ctx;
    const uint32_t ni = f->dev.n_instances;
    if (mmdgpu_status s = check_anims(f, per_instance)) return s;
    std::vector<uint64_t>& bound = f->set[f->cur].bound;
    bool same = bound.size() == ni;
    for (uint32_t i = 0; same && i < ni; ++i) same = bound[i] == per_instance[i]->uid;
    if (same) return MMDGPU_OK;
    // the previous upload may still be in flight from pageable memory semantics' point of view: cudaMemcpyAsync
    // from pageable memory returns after staging, so h_anims may be rewritten immediately.
    f->h_anims.resize(ni);
    for (uint32_t i = 0; i < ni; ++i) f->h_anims[i] = per_instance[i]->dev;
    CU(ctx, cudaMemcpyAsync(f->d_anims, f->h_anims.data(), sizeof(DevAnim) * ni, cudaMemcpyHostToDevice, st));
    bound.resize(ni);
    for (uint32_t i = 0; i < ni; ++i) bound[i] = per_instance[i]->uid;
    return MMDGPU_OK;
}

// Brackets one launch with events when profiling is on.
struct Timed {
    mmdgpu_context_t ctx;
    cudaStream_t st;
    cudaEvent_t b = nullptr;
    Timed(mmdgpu_context_t c, int id, cudaStream_t stream = nullptr) : ctx(c), st(stream ? stream : c->stream) {
        if (!c->profiling) return;
        cudaEvent_t ev[2] = {nullptr, nullptr};
        for (auto& e : ev) {
            if (!c->event_pool.empty()) { e = c->event_pool.back(); c->event_pool.pop_back(); }
            else if (cudaEventCreate(&e) != cudaSuccess) return;
        }
        cudaEventRecord(ev[0], st);
        b = ev[1];
        c->spans.push_back({id, ev[0], ev[1]});
    }
    ~Timed() {
        if (b) cudaEventRecord(b, st);
        ctx->launches++;
    }
};

mmdgpu_status enter(mmdgpu_context_t ctx) {
    if (!ctx) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "context is NULL");
    CU(ctx, cudaSetDevice(ctx->device));
    return MMDGPU_OK;
}

mmdgpu_status do_seek(mmdgpu_frames* f, const mmdgpu_animation_t* per_instance, const uint32_t* frames, bool range,
                      uint32_t stride, bool write_untracked, cudaStream_t st) {
    mmdgpu_context_t ctx = f->ctx;
    if (!frames) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "frame id array is NULL");
    if (mmdgpu_status s = bind_anims(f, per_instance, st)) return s;
    const uint32_t n = range ? f->dev.n_instances : f->dev.n_slots;
    // one frame id travels as a kernel argument: no host-to-device copy on the interactive / single-clip bake path
    // ... and up to kInlineFrameIds of them too (a small crowd): the copy from pageable memory costs more than the launch
    static const bool allow_inline = [] { const char* e = std::getenv("MMDGPU_INLINE_IDS"); return !(e && e[0] == '0'); }();   // A/B knob
    const bool by_value = n == 1, inline_ids = allow_inline && !by_value && n <= kInlineFrameIds;
    if (!by_value && !inline_ids) CU(ctx, cudaMemcpyAsync(f->dev.frame_id, frames, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    f->range_mode = range;
    f->frame_stride = stride;
    SampleSpec sp;
    sp.anims = f->d_anims;
    sp.write_untracked = write_untracked;
    sp.range_mode = range;
    sp.frame_stride = stride;
    sp.frame_by_value = by_value ? frames : nullptr;
    if (inline_ids) { sp.frame_ids_inline = frames; sp.n_inline = n; }
    {
        Timed t(ctx, MMDGPU_KERNEL_POSE_SAMPLE, st);
        CU(ctx, launch_pose_sample(st, f->model->dev, f->dev, sp));
    }
    return MMDGPU_OK;
}

// Batches of this many slots or more run their CCD IK waves in the flat one-thread-per-(op, slot) kernel, on
// chain-local images, between segments of the CTA-per-slot kernel (kernels.cu); below it one launch wins.
// B200, C2 (two chains per slot), G vertex-frames/s with / without: 256 slots 63.9 / 61.1, 512 slots 90.2 / 71.4,
// 1024 slots 102 / 72, 2048 slots 110.9 / 77.4 (profiles/r01_experiments.md).
// Round 2, with up to eight updates' hierarchies in flight (the CTA kernel holds 34 KB of shared memory per slot for the
// whole solve, the flat kernel almost nothing): 128 slots 59.7 / 51.9, 256 slots 79.3 / 64.2 (profiles/r02_experiments.md).
constexpr uint32_t kIkSplitMinSlots = 128;
constexpr size_t kIkSplitMaxWaves = 4;

static bool split_ik_waves(const mmdgpu_frames* f) {
    const mmdgpu_model* m = f->model;
    if (!m->ik_images_ok || m->ik_waves.empty() || m->ik_waves.size() > kIkSplitMaxWaves || !hierarchy_uses_cta_kernel(m->dev))
        return false;
    if (const char* env = std::getenv("MMDGPU_IK_SPLIT")) return env[0] == '1';  // experiment / test knob
    return f->dev.n_slots >= kIkSplitMinSlots;
}

mmdgpu_status do_hierarchy(mmdgpu_frames* f, uint32_t lo, uint32_t hi, bool prologue, cudaStream_t st) {
    mmdgpu_context_t ctx = f->ctx;
    const DevModel& M = f->model->dev;
    if (!split_ik_waves(f)) {
        Timed t(ctx, MMDGPU_KERNEL_HIERARCHY, st);
        CU(ctx, launch_hierarchy(st, M, f->dev, lo, hi, prologue));
        return MMDGPU_OK;
    }
    uint32_t cur = lo;
    bool pro = prologue;
    for (const auto& wn : f->model->ik_waves) {
        const uint32_t w = wn.first;
        if (w < lo || w >= hi) continue;
        if (w > cur || pro) {  // the segment before the IK wave (possibly empty: prologue only)
            Timed t(ctx, MMDGPU_KERNEL_HIERARCHY, st);
            CU(ctx, launch_hierarchy(st, M, f->dev, cur, w, pro));
            pro = false;
        }
        {
            Timed t(ctx, MMDGPU_KERNEL_HIERARCHY, st);
            CU(ctx, launch_hierarchy_wave_flat(st, M, f->dev, w, wn.second));
        }
        cur = w + 1;
    }
    if (cur < hi || pro) {
        Timed t(ctx, MMDGPU_KERNEL_HIERARCHY, st);
        CU(ctx, launch_hierarchy(st, M, f->dev, cur, hi, pro));
    }
    return MMDGPU_OK;
}

mmdgpu_status do_skin(mmdgpu_frames* f) {
    mmdgpu_context_t ctx = f->ctx;
    if (f->model->dev.nv_pad == 0) return MMDGPU_OK;
    if (f->dl_pending) {  // the copy of the previous update's vertices must have read the buffers this launch overwrites
        CU(ctx, cudaStreamWaitEvent(ctx->stream, f->ev_dl, 0));
        f->dl_pending = false;
    }
    {
        Timed t(ctx, MMDGPU_KERNEL_SKIN);
        CU(ctx, launch_skin(ctx->stream, f->model->dev, f->dev, int(f->layout), f->slots_per_cta));
    }
    return MMDGPU_OK;
}

struct StreamView {
    const char* base;
    size_t slot_stride, slot_bytes;
};
mmdgpu_status stream_view(mmdgpu_frames* f, mmdgpu_stream_id id, StreamView& v) {
    const DevModel& M = f->model->dev;
    switch (id) {
    case MMDGPU_STREAM_POSITION:
    case MMDGPU_STREAM_NORMAL:
        if (f->layout != MMDGPU_LAYOUT_SOA_POS_NRM)
            return set_err(f->ctx, MMDGPU_ERR_INVALID_ARG, "frames were created with the interleaved layout");
        v.base = reinterpret_cast<const char*>(id == MMDGPU_STREAM_POSITION ? f->dev.out_pos : f->dev.out_nrm);
        v.slot_stride = (id == MMDGPU_STREAM_POSITION ? f->dev.pos_stride : f->dev.nrm_stride) * 4;
        v.slot_bytes = size_t(M.nv) * 12;
        return MMDGPU_OK;
    case MMDGPU_STREAM_INTERLEAVED:
        if (f->layout != MMDGPU_LAYOUT_INTERLEAVED_SOKOL32)
            return set_err(f->ctx, MMDGPU_ERR_INVALID_ARG, "frames were created with the SoA layout");
        v.base = reinterpret_cast<const char*>(f->dev.out_inter);
        v.slot_stride = f->dev.inter_stride * 16;
        v.slot_bytes = size_t(M.nv) * 32;
        return MMDGPU_OK;
    case MMDGPU_STREAM_SKIN_MATRIX:
        v.base = reinterpret_cast<const char*>(f->dev.palette);
        v.slot_stride = v.slot_bytes = size_t(M.nb) * 48;
        return MMDGPU_OK;
    case MMDGPU_STREAM_UV:
        if (!f->dev.out_uv)
            return set_err(f->ctx, MMDGPU_ERR_INVALID_ARG, "the UV stream exists only for SoA frames of a model created with extensions");
        v.base = reinterpret_cast<const char*>(f->dev.out_uv);
        v.slot_stride = f->dev.uv_stride * 8;
        v.slot_bytes = size_t(M.nv) * 8;
        return MMDGPU_OK;
    }
    return set_err(f->ctx, MMDGPU_ERR_INVALID_ARG, "unknown stream id");
}

}  // namespace

// ------------------------------------------------------------------------------------------- basics
extern "C" {

MMDGPU_API int mmdgpu_version(void) { return MMDGPU_VERSION_MAJOR * 100 + MMDGPU_VERSION_MINOR; }

MMDGPU_API const char* mmdgpu_status_string(mmdgpu_status s) {
    switch (s) {
    case MMDGPU_OK: return "ok";
    case MMDGPU_ERR_INVALID_ARG: return "invalid argument";
    case MMDGPU_ERR_BAD_INDEX: return "index out of range";
    case MMDGPU_ERR_UNSUPPORTED: return "unsupported";
    case MMDGPU_ERR_CUDA: return "CUDA error";
    case MMDGPU_ERR_OOM: return "out of memory";
    case MMDGPU_ERR_PARSE: return "malformed PMX/VMD data";
    }
    return "unknown status";
}

MMDGPU_API mmdgpu_status mmdgpu_context_create(int device, void* cuda_stream_or_null, mmdgpu_context_t* out) {
    if (!out) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceCount");
    if (count == 0) return set_err(nullptr, MMDGPU_ERR_CUDA, "no CUDA device");
    if (device < 0 || device >= count) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "device ordinal out of range");
    // every failure below leaves through mmdgpu_context_destroy, which releases whatever had been created by then
    std::unique_ptr<mmdgpu_context, void (*)(mmdgpu_context*)> c(new (std::nothrow) mmdgpu_context(),
                                                                 [](mmdgpu_context* p) { mmdgpu_context_destroy(p); });
    if (!c) return set_err(nullptr, MMDGPU_ERR_OOM, "host allocation failed");
    c->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    cudaDeviceProp prop{};
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major < 10)
        return set_err(nullptr, MMDGPU_ERR_UNSUPPORTED, "this library is built for sm_100a only; device is sm_" +
                                                            std::to_string(prop.major) + std::to_string(prop.minor));
    c->max_smem_optin = int(prop.sharedMemPerBlockOptin);
    c->sm_count = prop.multiProcessorCount;
    if (cuda_stream_or_null) c->stream = static_cast<cudaStream_t>(cuda_stream_or_null);
    else {
        if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess)
            return cuda_fail(nullptr, e, "cudaStreamCreate");
        c->own_stream = true;
    }
    if ((e = cudaStreamCreateWithFlags(&c->dl_stream, cudaStreamNonBlocking)) != cudaSuccess)
        return cuda_fail(nullptr, e, "cudaStreamCreate");
    if ((e = cudaEventCreateWithFlags(&c->dl_event, cudaEventDisableTiming)) != cudaSuccess)
        return cuda_fail(nullptr, e, "cudaEventCreate");
    {
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        if (const char* env = std::getenv("MMDGPU_PRE_STREAMS")) c->n_pre_ik = std::min(kPreStreams, std::max(1, std::atoi(env)));
        for (int i = 0; i < kPreStreams; ++i) {
            if ((e = cudaStreamCreateWithFlags(&c->pre_stream[i], cudaStreamNonBlocking)) != cudaSuccess)
                return cuda_fail(nullptr, e, "cudaStreamCreate");
            if ((e = cudaStreamCreateWithPriority(&c->pre_stream_hi[i], cudaStreamNonBlocking, greatest)) != cudaSuccess)
                return cuda_fail(nullptr, e, "cudaStreamCreate");
        }
    }
    *out = c.release();
    return MMDGPU_OK;
}

MMDGPU_API void mmdgpu_context_destroy(mmdgpu_context_t ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->sync_pre();
    cudaStreamSynchronize(ctx->stream);
    if (ctx->dl_stream) { cudaStreamSynchronize(ctx->dl_stream); cudaStreamDestroy(ctx->dl_stream); }
    for (int i = 0; i < kPreStreams; ++i) {
        if (ctx->pre_stream[i]) cudaStreamDestroy(ctx->pre_stream[i]);
        if (ctx->pre_stream_hi[i]) cudaStreamDestroy(ctx->pre_stream_hi[i]);
    }
    if (ctx->dl_event) cudaEventDestroy(ctx->dl_event);
    for (auto& sp : ctx->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

MMDGPU_API const char* mmdgpu_last_error(mmdgpu_context_t ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

MMDGPU_API mmdgpu_status mmdgpu_context_synchronize(mmdgpu_context_t ctx) {
    if (mmdgpu_status s = enter(ctx)) return s;
    CU(ctx, ctx->sync_pre());
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->dl_stream));
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_context_set_profiling(mmdgpu_context_t ctx, int enabled) {
    if (mmdgpu_status s = enter(ctx)) return s;
    ctx->profiling = enabled != 0;
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_context_profile_read(mmdgpu_context_t ctx, double ms_total[MMDGPU_KERNEL_COUNT],
                                                     uint64_t launches[MMDGPU_KERNEL_COUNT]) {
    if (mmdgpu_status s = enter(ctx)) return s;
    if (!ms_total || !launches) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "NULL argument");
    CU(ctx, ctx->sync_pre());
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < MMDGPU_KERNEL_COUNT; ++i) { ms_total[i] = 0.0; launches[i] = 0; }
    for (const auto& sp : ctx->spans) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess && sp.id >= 0 && sp.id < MMDGPU_KERNEL_COUNT) {
            ms_total[sp.id] += double(ms);
            launches[sp.id]++;
        }
        ctx->event_pool.push_back(sp.a);
        ctx->event_pool.push_back(sp.b);
    }
    ctx->spans.clear();
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_context_join_downloads(mmdgpu_context_t ctx) {
    if (mmdgpu_status s = enter(ctx)) return s;
    CU(ctx, cudaEventRecord(ctx->dl_event, ctx->dl_stream));
    CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->dl_event, 0));
    return MMDGPU_OK;
}

MMDGPU_API void* mmdgpu_context_stream(mmdgpu_context_t ctx) { return ctx ? ctx->stream : nullptr; }
MMDGPU_API uint64_t mmdgpu_context_launch_count(mmdgpu_context_t ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------------------------------- plan
MMDGPU_API mmdgpu_status mmdgpu_plan_create(const mmdgpu_model_desc* desc, const mmdgpu_options* opt, mmdgpu_plan_t* out,
                                            char* err_buf, size_t err_buf_len) {
    auto report = [&](const std::string& m) {
        g_err = m;
        if (err_buf && err_buf_len) std::snprintf(err_buf, err_buf_len, "%s", m.c_str());
    };
    if (!desc || !out) { report("desc or out is NULL"); return MMDGPU_ERR_INVALID_ARG; }
    *out = nullptr;
    std::unique_ptr<mmdgpu_plan> p(new (std::nothrow) mmdgpu_plan());
    if (!p) { report("host allocation failed"); return MMDGPU_ERR_OOM; }
    std::string err;
    mmdgpu_status s;
    try {
        s = build_plan(*desc, opt, p->plan, err);
    } catch (const std::bad_alloc&) {
        s = MMDGPU_ERR_OOM; err = "host allocation failed";
    }
    if (s != MMDGPU_OK) { report(err); return s; }
    *out = p.release();
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_plan_create_from_pmx(const void* bytes, size_t n, const mmdgpu_options* opt, mmdgpu_plan_t* out,
                                                     char* err_buf, size_t err_buf_len) {
    auto report = [&](const std::string& m) {
        g_err = m;
        if (err_buf && err_buf_len) std::snprintf(err_buf, err_buf_len, "%s", m.c_str());
    };
    if (!bytes || !out) { report("bytes or out is NULL"); return MMDGPU_ERR_INVALID_ARG; }
    *out = nullptr;
    std::string err;
    mmdgpu_status s;
    try {
        ParsedModel pm;
        s = parse_pmx(bytes, n, pm, err);
        if (s != MMDGPU_OK) { report(err); return s; }
        s = mmdgpu_plan_create(&pm.desc, opt, out, err_buf, err_buf_len);
        if (s != MMDGPU_OK) return s;
        (*out)->plan.bone_names = std::move(pm.bone_names);
        (*out)->plan.morph_names = std::move(pm.morph_names);
        (*out)->plan.names_utf8 = pm.utf8;
    } catch (const std::bad_alloc&) {
        report("host allocation failed");
        return MMDGPU_ERR_OOM;
    }
    return MMDGPU_OK;
}

MMDGPU_API void mmdgpu_plan_destroy(mmdgpu_plan_t plan) { delete plan; }

MMDGPU_API mmdgpu_status mmdgpu_anim_plan_create(const mmdgpu_anim_desc* desc, uint32_t n_bones, uint32_t n_morphs,
                                                 mmdgpu_anim_plan_t* out, char* err_buf, size_t err_buf_len) {
    auto report = [&](const std::string& m) {
        g_err = m;
        if (err_buf && err_buf_len) std::snprintf(err_buf, err_buf_len, "%s", m.c_str());
    };
    if (!desc || !out) { report("desc or out is NULL"); return MMDGPU_ERR_INVALID_ARG; }
    *out = nullptr;
    std::unique_ptr<mmdgpu_anim_plan> a(new (std::nothrow) mmdgpu_anim_plan());
    if (!a) { report("host allocation failed"); return MMDGPU_ERR_OOM; }
    std::string err;
    mmdgpu_status s;
    try {
        s = build_anim(*desc, n_bones, n_morphs, a->anim, err);
    } catch (const std::bad_alloc&) {
        s = MMDGPU_ERR_OOM; err = "host allocation failed";
    }
    if (s != MMDGPU_OK) { report(err); return s; }
    *out = a.release();
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_anim_plan_create_from_vmd(mmdgpu_plan_t model_plan, const void* bytes, size_t n,
                                                          mmdgpu_anim_plan_t* out, char* err_buf, size_t err_buf_len) {
    auto report = [&](const std::string& m) {
        g_err = m;
        if (err_buf && err_buf_len) std::snprintf(err_buf, err_buf_len, "%s", m.c_str());
    };
    if (!model_plan || !bytes || !out) { report("NULL argument"); return MMDGPU_ERR_INVALID_ARG; }
    *out = nullptr;
    std::string err;
    try {
        ParsedMotion pm;
        mmdgpu_status s = parse_vmd(bytes, n, model_plan->plan, pm, err);
        if (s != MMDGPU_OK) { report(err); return s; }
        return mmdgpu_anim_plan_create(&pm.desc, model_plan->plan.nb, model_plan->plan.nm, out, err_buf, err_buf_len);
    } catch (const std::bad_alloc&) {
        report("host allocation failed");
        return MMDGPU_ERR_OOM;
    }
}

MMDGPU_API void mmdgpu_anim_plan_destroy(mmdgpu_anim_plan_t plan) { delete plan; }

MMDGPU_API mmdgpu_status mmdgpu_anim_plan_get(mmdgpu_anim_plan_t plan, mmdgpu_anim_array which, const void** data, size_t* count) {
    if (!plan || !data || !count) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "NULL argument");
    const HostAnim& a = plan->anim;
#define RET(vec)                \
    do {                        \
        *data = (vec).data();   \
        *count = (vec).size();  \
        return MMDGPU_OK;       \
    } while (0)
    switch (which) {
    case MMDGPU_ANIM_BONE_KEY_BEGIN: RET(a.bone_key_begin);
    case MMDGPU_ANIM_BONE_KEY_COUNT: RET(a.bone_key_count);
    case MMDGPU_ANIM_BONE_TRACKED: RET(a.bone_tracked);
    case MMDGPU_ANIM_KEY_FRAME: RET(a.key_frame);
    case MMDGPU_ANIM_KEY_T: RET(a.key_T);
    case MMDGPU_ANIM_KEY_R: RET(a.key_R);
    case MMDGPU_ANIM_KEY_CURVE: RET(a.key_curve);
    case MMDGPU_ANIM_TABLES: RET(a.tables);
    case MMDGPU_ANIM_MORPH_KEY_BEGIN: RET(a.morph_key_begin);
    case MMDGPU_ANIM_MORPH_KEY_COUNT: RET(a.morph_key_count);
    case MMDGPU_ANIM_MORPH_TRACKED: RET(a.morph_tracked);
    case MMDGPU_ANIM_MKEY_FRAME: RET(a.mkey_frame);
    case MMDGPU_ANIM_MKEY_WEIGHT: RET(a.mkey_weight);
    }
#undef RET
    return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "unknown animation array");
}

MMDGPU_API mmdgpu_status mmdgpu_plan_get(mmdgpu_plan_t plan, mmdgpu_plan_array which, const void** data, size_t* count) {
    if (!plan || !data || !count) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "NULL argument");
    const Plan& p = plan->plan;
#define RET(vec)                \
    do {                        \
        *data = (vec).data();   \
        *count = (vec).size();  \
        return MMDGPU_OK;       \
    } while (0)
    switch (which) {
    case MMDGPU_PLAN_SKIN_TYPE: RET(p.norm_type);
    case MMDGPU_PLAN_BONE_ID: RET(p.bone_id);
    case MMDGPU_PLAN_WEIGHT: RET(p.weight);
    case MMDGPU_PLAN_ORDER_PRE: RET(p.order_pre);
    case MMDGPU_PLAN_ORDER_POST: RET(p.order_post);
    case MMDGPU_PLAN_OP_KIND: RET(p.op_kind_u8);
    case MMDGPU_PLAN_OP_BONE: RET(p.op_arg_i32);
    case MMDGPU_PLAN_OP_WAVE: RET(p.op_wave);
    case MMDGPU_PLAN_WAVE_BEGIN: RET(p.wave_begin);
    case MMDGPU_PLAN_WAVE_OPS: RET(p.wave_ops);
    case MMDGPU_PLAN_IK_FIX_TYPE: RET(p.ik_fix_u8);
    case MMDGPU_PLAN_IK_EULER_ORDER: RET(p.ik_order_u8);
    case MMDGPU_PLAN_APP_SLOT_MORPH: RET(p.node_morph);
    case MMDGPU_PLAN_APP_SLOT_PARENT: RET(p.node_parent);
    case MMDGPU_PLAN_APP_SLOT_MULT: RET(p.node_mult);
    case MMDGPU_PLAN_CSR_ROW_PTR: RET(p.csr_row);
    case MMDGPU_PLAN_CSR_SLOT: RET(p.csr_node);
    case MMDGPU_PLAN_CSR_OFFSET: RET(p.csr_offset);
    case MMDGPU_PLAN_WAVE_PHASE_SPLIT: RET(p.phase_split_i32);
    case MMDGPU_PLAN_TILE_ORIG: RET(p.tile_orig);
    case MMDGPU_PLAN_TILE_TYPE: RET(p.st_type);
    case MMDGPU_PLAN_TILE_LOCAL_ID: RET(p.st_local_id);
    case MMDGPU_PLAN_TILE_BONE_BEGIN: RET(p.tile_bone_begin);
    case MMDGPU_PLAN_TILE_BONES: RET(p.tile_bones);
    case MMDGPU_PLAN_ELL_BASE: RET(p.ell_base);
    case MMDGPU_PLAN_ELL_ROUNDS: RET(p.ell_rounds);
    case MMDGPU_PLAN_ELL_SLOT: RET(p.ell_node);
    case MMDGPU_PLAN_ELL_OFFSET: RET(p.ell_offset);
    case MMDGPU_PLAN_POSITION: RET(p.position);
    case MMDGPU_PLAN_NORMAL: RET(p.normal);
    case MMDGPU_PLAN_UV: RET(p.uv);
    case MMDGPU_PLAN_BONE_STATIC:
        *data = p.bones.data(); *count = p.bones.size() * sizeof(BoneStatic); return MMDGPU_OK;
    case MMDGPU_PLAN_IK_DESC:
        *data = p.iks.data(); *count = p.iks.size() * sizeof(IkDesc); return MMDGPU_OK;
    case MMDGPU_PLAN_IK_LINK:
        *data = p.links.data(); *count = p.links.size() * sizeof(IkLink); return MMDGPU_OK;
    case MMDGPU_PLAN_IK_IMAGE:
        *data = p.ik_img.data(); *count = p.ik_img.size() * sizeof(IkImage); return MMDGPU_OK;
    case MMDGPU_PLAN_IK_IMAGE_BONES: RET(p.ik_img_bones);
    case MMDGPU_PLAN_IK_IMAGE_WRITTEN: RET(p.ik_img_written);
    case MMDGPU_PLAN_IK_IMAGE_STATIC:
        *data = p.ik_img_static.data(); *count = p.ik_img_static.size() * sizeof(BoneStatic); return MMDGPU_OK;
    case MMDGPU_PLAN_IK_IMAGE_LINK_SLOTS: RET(p.ik_img_lslots);
    case MMDGPU_PLAN_IK_IMAGE_MORPH_SLOTS: RET(p.ik_img_mslots);
    case MMDGPU_PLAN_IK_IMAGE_DESC:
        *data = p.ik_img_desc.data(); *count = p.ik_img_desc.size() * sizeof(IkDesc); return MMDGPU_OK;
    case MMDGPU_PLAN_IK_IMAGE_LINKS:
        *data = p.ik_img_links.data(); *count = p.ik_img_links.size() * sizeof(IkLink); return MMDGPU_OK;
    case MMDGPU_PLAN_MATERIAL_MORPH_ROW: RET(p.material_morph_row);
    case MMDGPU_PLAN_MATERIAL_MORPH:
        *data = p.material_morph_entries.data(); *count = p.material_morph_entries.size() * sizeof(MaterialMorphEntry); return MMDGPU_OK;
    case MMDGPU_PLAN_BONE_MORPH:
        *data = p.bone_morph_entries.data(); *count = p.bone_morph_entries.size() * sizeof(BoneMorphEntry); return MMDGPU_OK;
    default: break;
    }
#undef RET
    return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "unknown plan array");
}

MMDGPU_API int mmdgpu_bezier_table(const int8_t ctrl[4], float table[32]) { return bezier_table(ctrl, table) ? 1 : 0; }

// ------------------------------------------------------------------------------------------- model
MMDGPU_API mmdgpu_status mmdgpu_model_create_from_arrays(mmdgpu_context_t ctx, const mmdgpu_model_desc* desc,
                                                         const mmdgpu_options* opt, mmdgpu_model_t* out) {
    if (mmdgpu_status s = enter(ctx)) return s;
    if (!desc || !out) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "desc or out is NULL");
    *out = nullptr;
    std::unique_ptr<mmdgpu_model> m(new (std::nothrow) mmdgpu_model());
    if (!m) return set_err(ctx, MMDGPU_ERR_OOM, "host allocation failed");
    m->ctx = ctx;
    std::string err;
    mmdgpu_status s;
    try {
        s = build_plan(*desc, opt, m->plan.plan, err);
        if (s != MMDGPU_OK) return set_err(ctx, s, err);
        s = upload_model(m.get());
    } catch (const std::bad_alloc&) {
        return set_err(ctx, MMDGPU_ERR_OOM, "host allocation failed");
    }
    if (s != MMDGPU_OK) return s;
    *out = m.release();
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_model_create_from_pmx(mmdgpu_context_t ctx, const void* bytes, size_t n,
                                                      const mmdgpu_options* opt, mmdgpu_model_t* out) {
    if (mmdgpu_status s = enter(ctx)) return s;
    if (!bytes || !out) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "bytes or out is NULL");
    *out = nullptr;
    ParsedModel pm;
    std::string err;
    mmdgpu_status s;
    try {
        s = parse_pmx(bytes, n, pm, err);
    } catch (const std::bad_alloc&) {
        return set_err(ctx, MMDGPU_ERR_OOM, "host allocation failed");
    }
    if (s != MMDGPU_OK) return set_err(ctx, s, err);
    s = mmdgpu_model_create_from_arrays(ctx, &pm.desc, opt, out);
    if (s != MMDGPU_OK) return s;
    (*out)->plan.plan.bone_names = std::move(pm.bone_names);
    (*out)->plan.plan.morph_names = std::move(pm.morph_names);
    (*out)->plan.plan.names_utf8 = pm.utf8;
    return MMDGPU_OK;
}

MMDGPU_API void mmdgpu_model_destroy(mmdgpu_model_t model) {
    if (!model) return;
    cudaSetDevice(model->ctx->device);
    model->ctx->sync_pre();
    cudaStreamSynchronize(model->ctx->stream);
    delete model;
}
MMDGPU_API uint32_t mmdgpu_model_vertex_count(mmdgpu_model_t m) { return m ? m->plan.plan.nv : 0; }
MMDGPU_API uint32_t mmdgpu_model_bone_count(mmdgpu_model_t m) { return m ? m->plan.plan.nb : 0; }
MMDGPU_API uint32_t mmdgpu_model_morph_count(mmdgpu_model_t m) { return m ? m->plan.plan.nm : 0; }
MMDGPU_API uint32_t mmdgpu_model_material_count(mmdgpu_model_t m) { return m ? m->plan.plan.n_materials : 0; }
MMDGPU_API mmdgpu_plan_t mmdgpu_model_plan(mmdgpu_model_t m) { return m ? &m->plan : nullptr; }

static int32_t find_name(const std::vector<std::string>& names, const void* bytes, size_t n) {
    if (!bytes) return -1;
    const std::string key(static_cast<const char*>(bytes), n);
    for (size_t i = 0; i < names.size(); ++i)
        if (names[i] == key) return int32_t(i);
    return -1;
}
MMDGPU_API int32_t mmdgpu_model_find_bone(mmdgpu_model_t m, const void* name_bytes, size_t n) {
    return m ? find_name(m->plan.plan.bone_names, name_bytes, n) : -1;
}
MMDGPU_API int32_t mmdgpu_model_find_morph(mmdgpu_model_t m, const void* name_bytes, size_t n) {
    return m ? find_name(m->plan.plan.morph_names, name_bytes, n) : -1;
}

// ------------------------------------------------------------------------------------------- animation
MMDGPU_API mmdgpu_status mmdgpu_animation_create_from_arrays(mmdgpu_context_t ctx, mmdgpu_model_t model,
                                                             const mmdgpu_anim_desc* desc, mmdgpu_animation_t* out) {
    if (mmdgpu_status s = enter(ctx)) return s;
    if (!model || !desc || !out) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "model, desc or out is NULL");
    if (model->ctx != ctx) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "model belongs to a different context");
    *out = nullptr;
    std::unique_ptr<mmdgpu_animation> a(new (std::nothrow) mmdgpu_animation());
    if (!a) return set_err(ctx, MMDGPU_ERR_OOM, "host allocation failed");
    a->ctx = ctx;
    a->model = model;
    a->uid = g_next_anim_uid.fetch_add(1, std::memory_order_relaxed);
    std::string err;
    mmdgpu_status s;
    try {
        s = build_anim(*desc, model->plan.plan.nb, model->plan.plan.nm, a->host, err);
        if (s != MMDGPU_OK) return set_err(ctx, s, err);
        s = upload_anim(a.get());
    } catch (const std::bad_alloc&) {
        return set_err(ctx, MMDGPU_ERR_OOM, "host allocation failed");
    }
    if (s != MMDGPU_OK) return s;
    *out = a.release();
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_animation_create_from_vmd(mmdgpu_context_t ctx, mmdgpu_model_t model, const void* bytes,
                                                          size_t n, mmdgpu_animation_t* out) {
    if (mmdgpu_status s = enter(ctx)) return s;
    if (!model || !bytes || !out) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "model, bytes or out is NULL");
    *out = nullptr;
    ParsedMotion pm;
    std::string err;
    mmdgpu_status s;
    try {
        s = parse_vmd(bytes, n, model->plan.plan, pm, err);
    } catch (const std::bad_alloc&) {
        return set_err(ctx, MMDGPU_ERR_OOM, "host allocation failed");
    }
    if (s != MMDGPU_OK) return set_err(ctx, s, err);
    return mmdgpu_animation_create_from_arrays(ctx, model, &pm.desc, out);
}

MMDGPU_API void mmdgpu_animation_destroy(mmdgpu_animation_t a) {
    if (!a) return;
    cudaSetDevice(a->ctx->device);
    a->ctx->sync_pre();
    cudaStreamSynchronize(a->ctx->stream);
    delete a;
}
MMDGPU_API uint32_t mmdgpu_animation_length(mmdgpu_animation_t a) { return a ? a->host.length : 0; }

// ------------------------------------------------------------------------------------------- frames
MMDGPU_API mmdgpu_status mmdgpu_frames_create(mmdgpu_context_t ctx, mmdgpu_model_t model, uint32_t n_instances,
                                              uint32_t n_frames, mmdgpu_layout layout, mmdgpu_frames_t* out) {
    if (mmdgpu_status s = enter(ctx)) return s;
    if (!model || !out) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "model or out is NULL");
    if (model->ctx != ctx) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "model belongs to a different context");
    if (n_instances == 0 || n_frames == 0) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "n_instances and n_frames must be > 0");
    if (uint64_t(n_instances) * n_frames > 65535u)
        return set_err(ctx, MMDGPU_ERR_UNSUPPORTED, "more than 65535 slots in one frames object (grid.y limit)");
    if (layout != MMDGPU_LAYOUT_SOA_POS_NRM && layout != MMDGPU_LAYOUT_INTERLEAVED_SOKOL32)
        return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "unknown layout");
    *out = nullptr;
    std::unique_ptr<mmdgpu_frames> f(new (std::nothrow) mmdgpu_frames());
    if (!f) return set_err(ctx, MMDGPU_ERR_OOM, "host allocation failed");
    f->ctx = ctx;
    f->model = model;
    f->layout = layout;
    const DevModel& M = model->dev;
    DevFrames& F = f->dev;
    const size_t ns = size_t(n_instances) * n_frames;
    F.n_slots = uint32_t(ns); F.n_instances = n_instances; F.n_frames = n_frames;
    cudaStream_t st = ctx->stream;
    f->n_copies = (model->plan.plan.iks.empty() ? ctx->n_pre_plain : ctx->n_pre_ik) + 1;
    for (int i = 0; i < f->n_copies; ++i) {
        mmdgpu_frames::StateSet& x = f->set[i];
        CU(ctx, dalloc(f->mem, &x.poseR, ns * M.nb, false, st));
        CU(ctx, dalloc(f->mem, &x.poseT, ns * M.nb, false, st));
        CU(ctx, dalloc(f->mem, &x.rate, ns * M.nm, true, st));
        // application-slot rates, [slot / 4][node][slot % 4] (kernels.cu)
        CU(ctx, dalloc(f->mem, &x.node_rate, (ns + kSlotGroup - 1) / kSlotGroup * kSlotGroup * M.n_nodes_pad, true, st));
        CU(ctx, dalloc(f->mem, &x.totR, ns * M.nb, true, st));
        CU(ctx, dalloc(f->mem, &x.totT, ns * M.nb, true, st));
        CU(ctx, dalloc(f->mem, &x.local, ns * M.nb * 12, true, st));
        CU(ctx, dalloc(f->mem, &x.ikR, ns * M.n_link_slots, true, st));
        CU(ctx, dalloc(f->mem, &x.preIK, ns * M.n_link_slots, true, st));
        CU(ctx, dalloc(f->mem, &x.morphR, ns * M.n_morph_slots, true, st));
        CU(ctx, dalloc(f->mem, &x.morphT, ns * M.n_morph_slots, true, st));
        CU(ctx, dalloc(f->mem, &x.palette, ns * M.nb * 3, true, st));
        if (M.extensions) CU(ctx, dalloc(f->mem, &x.pal_ext, ns * M.nb * 2, true, st));
        if (M.material_morph_entries)
            CU(ctx, dalloc(f->mem, &x.material_images, ns * M.n_materials * 2 * MMDGPU_MATERIAL_FIELDS, true, st));
        CU(ctx, dalloc(f->mem, &x.frame_id, ns, true, st));
        CU(ctx, dalloc(f->mem, &x.time_s, ns, true, st));
        CU(ctx, dalloc(f->mem, &x.d_anims, size_t(n_instances), true, st));
        CU(ctx, cudaEventCreateWithFlags(&f->ev_pre[i], cudaEventDisableTiming));
        CU(ctx, cudaEventCreateWithFlags(&f->ev_skin[i], cudaEventDisableTiming));
    }
    CU(ctx, cudaEventCreateWithFlags(&f->ev_main, cudaEventDisableTiming));
    f->select(0);
    CU(ctx, cudaEventCreateWithFlags(&f->ev_dl, cudaEventDisableTiming));
    CU(ctx, cudaEventCreateWithFlags(&f->ev_deform, cudaEventDisableTiming));
    if (layout == MMDGPU_LAYOUT_SOA_POS_NRM) {
        // one allocation, position planes then normal planes: a one-slot object (an interactive Poser) can hand both to the
        // host with a single copy (mmdgpu_frames_download_pair_async)
        CU(ctx, dalloc(f->mem, &F.out_pos, 2 * ns * M.nv_pad * 3, false, st));
        // (the skinning kernel stores exactly nv vertices per slot, so slots - and the normal planes behind the position
        // planes - follow each other at the next 16-byte boundary: for nv % 4 == 0 a stream of all slots is one contiguous
        // block and leaves in a plain 1-D copy instead of a pitched one)
        F.out_nrm = F.out_pos + ns * own_plane_stride(M.nv);
        if (M.extensions) CU(ctx, dalloc(f->mem, &F.out_uv, ns * M.nv_pad, false, st));
    } else {
        CU(ctx, dalloc(f->mem, &F.out_inter, ns * M.nv_pad * 2, false, st));
    }
    F.pos_stride = F.nrm_stride = own_plane_stride(M.nv);  // floats
    F.inter_stride = own_inter_stride(M.nv);               // float4
    F.uv_stride = own_uv_stride(M.nv);                     // float2
    f->own_pos = F.out_pos; f->own_nrm = F.out_nrm; f->own_inter = F.out_inter; f->own_uv = F.out_uv;
    f->slots_per_cta = choose_slots_per_cta(M.n_tiles, F.n_slots, ctx->sm_count);
    // Poser::Poser ends with ResetPosing() (poser_impl.inl:125-127): a fresh object holds identity poses (both copies).
    for (int i = f->n_copies - 1; i >= 0; --i) {
        f->select(i);
        Timed t(ctx, MMDGPU_KERNEL_POSE_SAMPLE);
        SampleSpec reset;
        reset.write_untracked = true;
        CU(ctx, launch_pose_sample(st, M, F, reset));
    }
    *out = f.release();
    return MMDGPU_OK;
}

MMDGPU_API void mmdgpu_frames_destroy(mmdgpu_frames_t f) {
    if (!f) return;
    cudaSetDevice(f->ctx->device);
    f->ctx->sync_pre();
    cudaStreamSynchronize(f->ctx->stream);
    cudaStreamSynchronize(f->ctx->dl_stream);
    delete f;
}
MMDGPU_API uint32_t mmdgpu_frames_slot_count(mmdgpu_frames_t f) { return f ? f->dev.n_slots : 0; }

MMDGPU_API mmdgpu_status mmdgpu_reset_posing(mmdgpu_frames_t f) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    f->main_dirty = true;
    {
        Timed t(f->ctx, MMDGPU_KERNEL_POSE_SAMPLE);
        SampleSpec reset;
        reset.write_untracked = true;
        CU(f->ctx, launch_pose_sample(f->ctx->stream, f->model->dev, f->dev, reset));
    }
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_seek_frame(mmdgpu_frames_t f, const mmdgpu_animation_t* per_instance,
                                           const uint32_t* frame_per_slot) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    f->main_dirty = true;
    return do_seek(f, per_instance, frame_per_slot, false, 1, false, f->ctx->stream);
}

MMDGPU_API mmdgpu_status mmdgpu_seek_frame_range(mmdgpu_frames_t f, const mmdgpu_animation_t* per_instance,
                                                 const uint32_t* first_frame_per_instance, uint32_t frame_stride) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    f->main_dirty = true;
    return do_seek(f, per_instance, first_frame_per_instance, true, frame_stride, false, f->ctx->stream);
}

static mmdgpu_status seek_time_common(mmdgpu_frames_t f, const mmdgpu_animation_t* per_instance, const double* time_per_slot,
                                      bool write_untracked) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    mmdgpu_context_t ctx = f->ctx;
    if (mmdgpu_status s = enter(ctx)) return s;
    if (!time_per_slot) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "time array is NULL");
    f->main_dirty = true;
    if (mmdgpu_status s = bind_anims(f, per_instance, ctx->stream)) return s;
    const bool by_value = f->dev.n_slots == 1;
    if (!by_value)
        CU(ctx, cudaMemcpyAsync(f->dev.time_s, time_per_slot, sizeof(double) * f->dev.n_slots, cudaMemcpyHostToDevice, ctx->stream));
    {
        Timed t(ctx, MMDGPU_KERNEL_POSE_SAMPLE);
        SampleSpec sp;
        sp.anims = f->d_anims;
        sp.write_untracked = write_untracked;
        sp.time_mode = true;
        sp.time_by_value = by_value ? time_per_slot : nullptr;
        CU(ctx, launch_pose_sample(ctx->stream, f->model->dev, f->dev, sp));
    }
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_seek_time(mmdgpu_frames_t f, const mmdgpu_animation_t* per_instance, const double* time_per_slot) {
    return seek_time_common(f, per_instance, time_per_slot, false);
}

MMDGPU_API mmdgpu_status mmdgpu_reset_and_seek_time(mmdgpu_frames_t f, const mmdgpu_animation_t* per_instance, const double* time_per_slot) {
    return seek_time_common(f, per_instance, time_per_slot, true);
}

// ResetPosing + SeekFrame / SeekTime + PrePhysicsPosing + PostPhysicsPosing on the context's stream: one sampling launch
// (identity / zero for what the clip does not animate) and the whole bone program in one hierarchy pass.
MMDGPU_API mmdgpu_status mmdgpu_pose_frame(mmdgpu_frames_t f, const mmdgpu_animation_t* per_instance, const uint32_t* frame_per_slot) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    f->main_dirty = true;
    if (mmdgpu_status s = do_seek(f, per_instance, frame_per_slot, false, 1, true, f->ctx->stream)) return s;
    return do_hierarchy(f, 0, f->model->dev.n_waves, true, f->ctx->stream);
}
MMDGPU_API mmdgpu_status mmdgpu_pose_time(mmdgpu_frames_t f, const mmdgpu_animation_t* per_instance, const double* time_per_slot) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = seek_time_common(f, per_instance, time_per_slot, true)) return s;
    return do_hierarchy(f, 0, f->model->dev.n_waves, true, f->ctx->stream);
}

MMDGPU_API mmdgpu_status mmdgpu_reset_and_seek_frame(mmdgpu_frames_t f, const mmdgpu_animation_t* per_instance,
                                                     const uint32_t* frame_per_slot) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    f->main_dirty = true;
    return do_seek(f, per_instance, frame_per_slot, false, 1, true, f->ctx->stream);
}

MMDGPU_API mmdgpu_status mmdgpu_set_bone_pose(mmdgpu_frames_t f, uint32_t slot, uint32_t bone, const float T[3],
                                              const float R[4]) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    if (!T || !R) return set_err(f->ctx, MMDGPU_ERR_INVALID_ARG, "pose is NULL");
    if (slot >= f->dev.n_slots || bone >= f->model->dev.nb) return set_err(f->ctx, MMDGPU_ERR_BAD_INDEX, "slot or bone out of range");
    f->main_dirty = true;
    const float t4[4] = {T[0], T[1], T[2], 0.0f};
    const size_t at = size_t(slot) * f->model->dev.nb + bone;
    CU(f->ctx, cudaMemcpyAsync(f->dev.poseT + at, t4, 16, cudaMemcpyHostToDevice, f->ctx->stream));
    CU(f->ctx, cudaMemcpyAsync(f->dev.poseR + at, R, 16, cudaMemcpyHostToDevice, f->ctx->stream));
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_set_morph_pose(mmdgpu_frames_t f, uint32_t slot, uint32_t morph, float weight) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    if (slot >= f->dev.n_slots || morph >= f->model->dev.nm) return set_err(f->ctx, MMDGPU_ERR_BAD_INDEX, "slot or morph out of range");
    f->main_dirty = true;
    CU(f->ctx, cudaMemcpyAsync(f->dev.rate + size_t(slot) * f->model->dev.nm + morph, &weight, 4, cudaMemcpyHostToDevice,
                               f->ctx->stream));
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_pre_physics_posing(mmdgpu_frames_t f) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    f->main_dirty = true;
    return do_hierarchy(f, 0, f->model->dev.phase_split, true, f->ctx->stream);
}

MMDGPU_API mmdgpu_status mmdgpu_post_physics_posing(mmdgpu_frames_t f) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    const DevModel& M = f->model->dev;
    if (M.phase_split >= M.n_waves) return MMDGPU_OK;  // no post-physics bones: nothing to evaluate
    f->main_dirty = true;
    return do_hierarchy(f, M.phase_split, M.n_waves, false, f->ctx->stream);
}

MMDGPU_API mmdgpu_status mmdgpu_deform(mmdgpu_frames_t f) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    // this launch reads state copy `cur` from the main stream: the fused updates' pre streams, which rewrite the copies
    // round-robin, must follow it (ev_main)
    f->main_dirty = true;
    if (mmdgpu_status s = do_skin(f)) return s;
    CU(f->ctx, cudaEventRecord(f->ev_deform, f->ctx->stream));
    f->last_skin = f->ev_deform;
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_set_skinning_matrix_override(mmdgpu_frames_t f, uint32_t slot, uint32_t bone,
                                                             const float skinning[16], const float* local_or_null) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    if (!skinning) return set_err(f->ctx, MMDGPU_ERR_INVALID_ARG, "matrix is NULL");
    const uint32_t nb = f->model->dev.nb;
    if (slot >= f->dev.n_slots || bone >= nb) return set_err(f->ctx, MMDGPU_ERR_BAD_INDEX, "slot or bone out of range");
    f->main_dirty = true;
    float cols[12];
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 4; ++r) cols[4 * c + r] = skinning[4 * r + c];
    CU(f->ctx, cudaMemcpyAsync(f->dev.palette + (size_t(slot) * nb + bone) * 3, cols, 48, cudaMemcpyHostToDevice, f->ctx->stream));
    if (local_or_null) {
        float rows[12];
        for (int r = 0; r < 4; ++r)
            for (int c = 0; c < 3; ++c) rows[3 * r + c] = local_or_null[4 * r + c];
        CU(f->ctx, cudaMemcpyAsync(f->dev.local + (size_t(slot) * nb + bone) * 12, rows, 48, cudaMemcpyHostToDevice, f->ctx->stream));
    }
    return MMDGPU_OK;
}

static mmdgpu_status update_common(mmdgpu_frames_t f, const mmdgpu_animation_t* per_instance, const uint32_t* frames,
                                   bool range, uint32_t stride) {
    // ResetPosing + SeekFrame collapse into one sampling launch that writes identity / zero for items the
    // clip does not animate (main.cpp:1788-1796).  Sampling and the hierarchy run on one of two alternating pre streams
    // into the next of the kStateCopies copies of the per-update state (mmdgpu_frames::StateSet), so they overlap the
    // previous update's hierarchy (other pre stream, other copy) and the skinning kernels still reading older copies.
    // Ordering: a copy is rewritten only after the skinning kernel that read it has finished (ev_skin), which in turn
    // ran after the hierarchy that wrote it (ev_pre); work issued on the main stream in between (step-wise calls,
    // pose uploads) is followed through ev_main by BOTH pre streams.
    mmdgpu_context_t ctx = f->ctx;
    // arguments are checked before anything rotates: a rejected call leaves `cur` on the copy of the last good update
    if (!frames) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "frame id array is NULL");
    if (mmdgpu_status s = check_anims(f, per_instance)) return s;
    const bool has_ik = !f->model->plan.plan.iks.empty();
    const int n_pre = f->n_copies - 1;
    const int next = (f->cur + 1) % f->n_copies;
    f->update_turn = (f->update_turn + 1) % n_pre;
    cudaStream_t pre = has_ik ? ctx->pre_stream_hi[f->update_turn] : ctx->pre_stream[f->update_turn];
    if (f->skin_recorded[next]) CU(ctx, cudaStreamWaitEvent(pre, f->ev_skin[next], 0));
    if (f->main_dirty) {  // step-wise calls / uploads issued on the main stream since the last fused update
        CU(ctx, cudaEventRecord(f->ev_main, ctx->stream));
        // any copy may have been touched from the main stream: the update after this one waits as well
        for (int i = 0; i < kPreStreams; ++i) CU(ctx, cudaStreamWaitEvent(has_ik ? ctx->pre_stream_hi[i] : ctx->pre_stream[i], f->ev_main, 0));
        f->main_dirty = false;
    }
    f->select(next);
    if (mmdgpu_status s = do_seek(f, per_instance, frames, range, stride, true, pre)) return s;
    const DevModel& M = f->model->dev;
    if (mmdgpu_status s = do_hierarchy(f, 0, M.n_waves, true, pre)) return s;
    CU(ctx, cudaEventRecord(f->ev_pre[next], pre));
    CU(ctx, cudaStreamWaitEvent(ctx->stream, f->ev_pre[next], 0));
    if (mmdgpu_status s = do_skin(f)) return s;
    CU(ctx, cudaEventRecord(f->ev_skin[next], ctx->stream));
    f->skin_recorded[next] = true;
    f->last_skin = f->ev_skin[next];
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_update(mmdgpu_frames_t f, const mmdgpu_animation_t* per_instance, const uint32_t* frame_per_slot) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    return update_common(f, per_instance, frame_per_slot, false, 1);
}

MMDGPU_API mmdgpu_status mmdgpu_update_range(mmdgpu_frames_t f, const mmdgpu_animation_t* per_instance,
                                             const uint32_t* first_frame_per_instance, uint32_t frame_stride) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    return update_common(f, per_instance, first_frame_per_instance, true, frame_stride);
}

MMDGPU_API mmdgpu_status mmdgpu_frames_device_ptr(mmdgpu_frames_t f, mmdgpu_stream_id id, void** dptr, size_t* slot_stride_bytes) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (!dptr) return set_err(f->ctx, MMDGPU_ERR_INVALID_ARG, "dptr is NULL");
    StreamView v{};
    if (mmdgpu_status s = stream_view(f, id, v)) return s;
    *dptr = const_cast<char*>(v.base);
    if (slot_stride_bytes) *slot_stride_bytes = v.slot_stride;
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_frames_download(mmdgpu_frames_t f, uint32_t slot, mmdgpu_stream_id id, void* host_dst, size_t bytes) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    if (!host_dst) return set_err(f->ctx, MMDGPU_ERR_INVALID_ARG, "host_dst is NULL");
    if (slot >= f->dev.n_slots) return set_err(f->ctx, MMDGPU_ERR_BAD_INDEX, "slot out of range");
    StreamView v{};
    if (mmdgpu_status s = stream_view(f, id, v)) return s;
    if (bytes != v.slot_bytes) return set_err(f->ctx, MMDGPU_ERR_INVALID_ARG, "bytes must equal the size of one slot of the stream");
    if (bytes) CU(f->ctx, cudaMemcpyAsync(host_dst, v.base + size_t(slot) * v.slot_stride, bytes, (f->host_bound ? cudaMemcpyDefault : cudaMemcpyDeviceToHost), f->ctx->stream));
    CU(f->ctx, cudaStreamSynchronize(f->ctx->stream));
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_frames_download_async(mmdgpu_frames_t f, uint32_t first_slot, uint32_t n_slots,
                                                      mmdgpu_stream_id id, void* pinned_host_dst, size_t bytes) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    mmdgpu_context_t ctx = f->ctx;
    if (mmdgpu_status s = enter(ctx)) return s;
    if (!pinned_host_dst) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "host_dst is NULL");
    if (uint64_t(first_slot) + n_slots > f->dev.n_slots) return set_err(ctx, MMDGPU_ERR_BAD_INDEX, "slot range out of range");
    StreamView v{};
    if (mmdgpu_status s = stream_view(f, id, v)) return s;
    if (bytes != v.slot_bytes * n_slots) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "bytes must equal n_slots x the size of one slot");
    if (bytes == 0) return MMDGPU_OK;
    CU(ctx, cudaEventRecord(ctx->dl_event, ctx->stream));
    CU(ctx, cudaStreamWaitEvent(ctx->dl_stream, ctx->dl_event, 0));
    const char* src = v.base + size_t(first_slot) * v.slot_stride;
    if (v.slot_bytes == v.slot_stride || n_slots == 1)   // contiguous: one plain copy
        CU(ctx, cudaMemcpyAsync(pinned_host_dst, src, bytes, (f->host_bound ? cudaMemcpyDefault : cudaMemcpyDeviceToHost), ctx->dl_stream));
    else
        CU(ctx, cudaMemcpy2DAsync(pinned_host_dst, v.slot_bytes, src, v.slot_stride, v.slot_bytes, n_slots,
                                  (f->host_bound ? cudaMemcpyDefault : cudaMemcpyDeviceToHost), ctx->dl_stream));
    CU(ctx, cudaEventRecord(f->ev_dl, ctx->dl_stream));
    f->dl_pending = f->dl_recorded = true;
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_frames_download_pair_async(mmdgpu_frames_t f, uint32_t slot, void* pinned_host_dst, size_t bytes) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    mmdgpu_context_t ctx = f->ctx;
    if (mmdgpu_status s = enter(ctx)) return s;
    if (!pinned_host_dst) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "host_dst is NULL");
    if (f->layout != MMDGPU_LAYOUT_SOA_POS_NRM) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "frames were created with the interleaved layout");
    if (slot >= f->dev.n_slots) return set_err(ctx, MMDGPU_ERR_BAD_INDEX, "slot out of range");
    const DevModel& M = f->model->dev;
    const size_t plane = size_t(M.nv) * 12;
    if (bytes != 2 * plane) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "bytes must equal positions + normals of one slot");
    if (bytes == 0) return MMDGPU_OK;
    const char* pos = reinterpret_cast<const char*>(f->dev.out_pos) + size_t(slot) * f->dev.pos_stride * 4;
    const char* nrm = reinterpret_cast<const char*>(f->dev.out_nrm) + size_t(slot) * f->dev.nrm_stride * 4;
    CU(ctx, cudaEventRecord(ctx->dl_event, ctx->stream));
    CU(ctx, cudaStreamWaitEvent(ctx->dl_stream, ctx->dl_event, 0));
    char* dst = static_cast<char*>(pinned_host_dst);
    if (nrm == pos + plane)   // one-slot object with unpadded planes next to each other: one transfer
        CU(ctx, cudaMemcpyAsync(dst, pos, bytes, (f->host_bound ? cudaMemcpyDefault : cudaMemcpyDeviceToHost), ctx->dl_stream));
    else {
        CU(ctx, cudaMemcpyAsync(dst, pos, plane, (f->host_bound ? cudaMemcpyDefault : cudaMemcpyDeviceToHost), ctx->dl_stream));
        CU(ctx, cudaMemcpyAsync(dst + plane, nrm, plane, (f->host_bound ? cudaMemcpyDefault : cudaMemcpyDeviceToHost), ctx->dl_stream));
    }
    CU(ctx, cudaEventRecord(f->ev_dl, ctx->dl_stream));
    f->dl_pending = f->dl_recorded = true;
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_frames_wait_downloads(mmdgpu_frames_t f) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    if (f->dl_recorded) CU(f->ctx, cudaEventSynchronize(f->ev_dl));
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_frames_wait_skinning(mmdgpu_frames_t f) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    if (f->last_skin) CU(f->ctx, cudaEventSynchronize(f->last_skin));
    return MMDGPU_OK;
}

MMDGPU_API int mmdgpu_frames_downloads_done(mmdgpu_frames_t f) {
    if (!f || !f->dl_recorded) return 1;
    if (enter(f->ctx) != MMDGPU_OK) return -1;
    const cudaError_t e = cudaEventQuery(f->ev_dl);
    if (e == cudaSuccess) return 1;
    if (e == cudaErrorNotReady) { cudaGetLastError(); return 0; }
    return -1;
}

MMDGPU_API mmdgpu_status mmdgpu_frames_bind_output(mmdgpu_frames_t f, mmdgpu_stream_id id, void* device_ptr, size_t slot_stride_bytes) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    mmdgpu_context_t ctx = f->ctx;
    if (mmdgpu_status s = enter(ctx)) return s;
    const DevModel& M = f->model->dev;
    const bool soa = f->layout == MMDGPU_LAYOUT_SOA_POS_NRM;
    size_t rec = 0, align = 16;
    switch (id) {
    case MMDGPU_STREAM_POSITION: case MMDGPU_STREAM_NORMAL:
        if (!soa) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "frames were created with the interleaved layout");
        rec = 12; break;
    case MMDGPU_STREAM_INTERLEAVED:
        if (soa) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "frames were created with the SoA layout");
        rec = 32; align = 32; break;  // records leave the SM as one 256-bit store each
    case MMDGPU_STREAM_UV:
        if (!f->own_uv) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "the UV stream exists only for SoA frames of a model created with extensions");
        rec = 8; break;
    default: return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "only the vertex output streams can be bound");
    }
    // queued kernels keep the pointers they were launched with; later launches see the new binding
    f->main_dirty = true;
    DevFrames& F = f->dev;
    if (!device_ptr) {  // back to the library-owned buffer
        if (id == MMDGPU_STREAM_POSITION) { F.out_pos = f->own_pos; F.pos_stride = own_plane_stride(M.nv); }
        else if (id == MMDGPU_STREAM_NORMAL) { F.out_nrm = f->own_nrm; F.nrm_stride = own_plane_stride(M.nv); }
        else if (id == MMDGPU_STREAM_INTERLEAVED) { F.out_inter = f->own_inter; F.inter_stride = own_inter_stride(M.nv); }
        else { F.out_uv = f->own_uv; F.uv_stride = own_uv_stride(M.nv); }
        return MMDGPU_OK;
    }
    if (reinterpret_cast<uintptr_t>(device_ptr) % align || slot_stride_bytes % align)
        return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "bound output pointer and slot stride must be " + std::to_string(align) + "-byte aligned");
    if (slot_stride_bytes < size_t(M.nv) * rec && F.n_slots > 1)
        return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "slot stride is smaller than one slot of the stream");
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, device_ptr) != cudaSuccess) attr.type = cudaMemoryTypeUnregistered;
    if (attr.type == cudaMemoryTypeHost && attr.devicePointer) {
        // page-locked host memory (mmdgpu_host_alloc, cudaHostAlloc, cudaHostRegister): the skinning kernel's stores travel
        // over PCIe while it computes; no separate device-to-host copy
        device_ptr = attr.devicePointer;
        f->host_bound = true;   // (stays set: it only selects cudaMemcpyDefault for the downloads of this object)
    } else if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged) {
        cudaGetLastError();
        return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "bound output pointer is neither device memory nor page-locked host memory");
    }
    if (id == MMDGPU_STREAM_POSITION) { F.out_pos = static_cast<float*>(device_ptr); F.pos_stride = slot_stride_bytes / 4; }
    else if (id == MMDGPU_STREAM_NORMAL) { F.out_nrm = static_cast<float*>(device_ptr); F.nrm_stride = slot_stride_bytes / 4; }
    else if (id == MMDGPU_STREAM_INTERLEAVED) { F.out_inter = static_cast<float4*>(device_ptr); F.inter_stride = slot_stride_bytes / 16; }
    else { F.out_uv = static_cast<float2*>(device_ptr); F.uv_stride = slot_stride_bytes / 8; }
    return MMDGPU_OK;
}

static mmdgpu_status download_small(mmdgpu_frames_t f, const void* src, size_t bytes, std::vector<float>& tmp) {
    tmp.resize(bytes / 4);
    CU(f->ctx, cudaMemcpyAsync(tmp.data(), src, bytes, cudaMemcpyDeviceToHost, f->ctx->stream));
    CU(f->ctx, cudaStreamSynchronize(f->ctx->stream));
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_bone_matrices_download(mmdgpu_frames_t f, uint32_t slot, float* host_dst) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    if (!host_dst) return set_err(f->ctx, MMDGPU_ERR_INVALID_ARG, "host_dst is NULL");
    if (slot >= f->dev.n_slots) return set_err(f->ctx, MMDGPU_ERR_BAD_INDEX, "slot out of range");
    const uint32_t nb = f->model->dev.nb;
    std::vector<float> tmp;
    if (mmdgpu_status s = download_small(f, f->dev.palette + size_t(slot) * nb * 3, size_t(nb) * 48, tmp)) return s;
    for (uint32_t b = 0; b < nb; ++b) {
        float* o = host_dst + size_t(b) * 16;
        for (int r = 0; r < 4; ++r) {
            for (int c = 0; c < 3; ++c) o[4 * r + c] = tmp[size_t(b) * 12 + 4 * c + r];
            o[4 * r + 3] = (r == 3) ? 1.0f : 0.0f;
        }
    }
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_bone_local_matrices_download(mmdgpu_frames_t f, uint32_t slot, float* host_dst) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    if (!host_dst) return set_err(f->ctx, MMDGPU_ERR_INVALID_ARG, "host_dst is NULL");
    if (slot >= f->dev.n_slots) return set_err(f->ctx, MMDGPU_ERR_BAD_INDEX, "slot out of range");
    const uint32_t nb = f->model->dev.nb;
    std::vector<float> tmp;
    if (mmdgpu_status s = download_small(f, f->dev.local + size_t(slot) * nb * 12, size_t(nb) * 48, tmp)) return s;
    for (uint32_t b = 0; b < nb; ++b) {
        float* o = host_dst + size_t(b) * 16;
        for (int r = 0; r < 4; ++r) {
            for (int c = 0; c < 3; ++c) o[4 * r + c] = tmp[size_t(b) * 12 + 3 * r + c];
            o[4 * r + 3] = (r == 3) ? 1.0f : 0.0f;
        }
    }
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_bone_poses_download(mmdgpu_frames_t f, uint32_t slot, float* host_dst) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    if (!host_dst) return set_err(f->ctx, MMDGPU_ERR_INVALID_ARG, "host_dst is NULL");
    if (slot >= f->dev.n_slots) return set_err(f->ctx, MMDGPU_ERR_BAD_INDEX, "slot out of range");
    const uint32_t nb = f->model->dev.nb;
    std::vector<float> t, r;
    if (mmdgpu_status s = download_small(f, f->dev.poseT + size_t(slot) * nb, size_t(nb) * 16, t)) return s;
    if (mmdgpu_status s = download_small(f, f->dev.poseR + size_t(slot) * nb, size_t(nb) * 16, r)) return s;
    for (uint32_t b = 0; b < nb; ++b) {
        float* o = host_dst + size_t(b) * 7;
        o[0] = t[4 * b]; o[1] = t[4 * b + 1]; o[2] = t[4 * b + 2];
        o[3] = r[4 * b]; o[4] = r[4 * b + 1]; o[5] = r[4 * b + 2]; o[6] = r[4 * b + 3];
    }
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_morph_rates_download(mmdgpu_frames_t f, uint32_t slot, float* host_dst) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    if (slot >= f->dev.n_slots) return set_err(f->ctx, MMDGPU_ERR_BAD_INDEX, "slot out of range");
    const uint32_t nm = f->model->dev.nm;
    if (nm == 0) return MMDGPU_OK;
    if (!host_dst) return set_err(f->ctx, MMDGPU_ERR_INVALID_ARG, "host_dst is NULL");
    CU(f->ctx, cudaMemcpyAsync(host_dst, f->dev.rate + size_t(slot) * nm, size_t(nm) * 4, cudaMemcpyDeviceToHost, f->ctx->stream));
    CU(f->ctx, cudaStreamSynchronize(f->ctx->stream));
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_material_images_download(mmdgpu_frames_t f, uint32_t slot, float* host_dst) {
    if (!f) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "frames is NULL");
    if (mmdgpu_status s = enter(f->ctx)) return s;
    if (slot >= f->dev.n_slots) return set_err(f->ctx, MMDGPU_ERR_BAD_INDEX, "slot out of range");
    const uint32_t nmat = f->model->dev.n_materials;
    if (nmat == 0) return MMDGPU_OK;
    if (!host_dst) return set_err(f->ctx, MMDGPU_ERR_INVALID_ARG, "host_dst is NULL");
    const size_t per = size_t(nmat) * 2 * MMDGPU_MATERIAL_FIELDS;
    if (!f->dev.material_images) {
        // libmmd-exact mode, or a model without material morphs: the images libmmd allocates and never touches
        for (uint32_t m = 0; m < nmat; ++m)
            for (uint32_t k = 0; k < MMDGPU_MATERIAL_FIELDS; ++k) {
                host_dst[(size_t(m) * 2) * MMDGPU_MATERIAL_FIELDS + k] = 1.0f;
                host_dst[(size_t(m) * 2 + 1) * MMDGPU_MATERIAL_FIELDS + k] = 0.0f;
            }
        return MMDGPU_OK;
    }
    CU(f->ctx, cudaMemcpyAsync(host_dst, f->dev.material_images + size_t(slot) * per, per * 4, cudaMemcpyDeviceToHost, f->ctx->stream));
    CU(f->ctx, cudaStreamSynchronize(f->ctx->stream));
    return MMDGPU_OK;
}

MMDGPU_API uint32_t mmdgpu_frames_slot_run(mmdgpu_frames_t f) { return f ? f->slots_per_cta : 0; }

// Test export: runs ONE device math function (csrc/mmd_math.cuh) on n rows of inputs.
MMDGPU_API mmdgpu_status mmdgpu_test_math(mmdgpu_context_t ctx, int op, const float* in, uint32_t n, float* out) {
    if (mmdgpu_status s = enter(ctx)) return s;
    static const int kin[11] = {5, 9, 9, 5, 4, 4, 8, 4, 4, 32, 3}, kout[11] = {1, 4, 4, 3, 4, 4, 4, 9, 4, 16, 3};
    if (op < 0 || op > 10 || !in || !out) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "unknown op or NULL argument");
    DevArena mem;
    std::vector<float> tables;
    std::vector<uint32_t> curve;
    if (op == 0) {   // Bezier: the host builds the presampled tables exactly as for a clip, the device looks them up
        curve.resize(n);
        for (uint32_t i = 0; i < n; ++i) {
            const int8_t c[4] = {int8_t(in[5 * i]), int8_t(in[5 * i + 1]), int8_t(in[5 * i + 2]), int8_t(in[5 * i + 3])};
            float t[32];
            if (bezier_table(c, t)) curve[i] = 0xFFFFFFFFu;
            else { curve[i] = uint32_t(tables.size() / 32); tables.insert(tables.end(), t, t + 32); }
        }
    }
    const float *d_in = nullptr, *d_tab = nullptr;
    const uint32_t* d_curve = nullptr;
    float* d_out = nullptr;
    CU(ctx, upload(ctx, mem, in, size_t(n) * kin[op], &d_in));
    CU(ctx, upload(ctx, mem, tables, &d_tab));
    CU(ctx, upload(ctx, mem, curve, &d_curve));
    CU(ctx, dalloc(mem, &d_out, size_t(n) * kout[op], true, ctx->stream));
    CU(ctx, launch_math_kat(ctx->stream, op, d_in, n, d_out, d_tab, d_curve));
    CU(ctx, cudaMemcpyAsync(out, d_out, size_t(n) * kout[op] * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return MMDGPU_OK;
}

// ---- peer buffers: the receive side of a bake gather that the skinning kernels of other ranks write into directly
MMDGPU_API mmdgpu_status mmdgpu_peer_buffer_create(mmdgpu_context_t ctx, size_t bytes, void** dptr, unsigned char handle[64]) {
    if (mmdgpu_status s = enter(ctx)) return s;
    if (!dptr || !handle) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "NULL argument");
    *dptr = nullptr;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the C-ABI");
    void* p = nullptr;
    CU(ctx, cudaMalloc(&p, bytes ? bytes : 16));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(ctx, e, "cudaIpcGetMemHandle"); }
    std::memcpy(handle, &h, 64);
    *dptr = p;
    return MMDGPU_OK;
}
MMDGPU_API mmdgpu_status mmdgpu_peer_buffer_open(mmdgpu_context_t ctx, const unsigned char handle[64], void** dptr) {
    if (mmdgpu_status s = enter(ctx)) return s;
    if (!dptr || !handle) return set_err(ctx, MMDGPU_ERR_INVALID_ARG, "NULL argument");
    *dptr = nullptr;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    // mapped into this context's device; peer access to the exporting device is enabled as part of the call
    CU(ctx, cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
    return MMDGPU_OK;
}
MMDGPU_API mmdgpu_status mmdgpu_peer_buffer_release(mmdgpu_context_t ctx, void* dptr, int opened) {
    if (mmdgpu_status s = enter(ctx)) return s;
    if (!dptr) return MMDGPU_OK;
    CU(ctx, ctx->sync_pre());
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (opened) CU(ctx, cudaIpcCloseMemHandle(dptr));
    else CU(ctx, cudaFree(dptr));
    return MMDGPU_OK;
}

MMDGPU_API mmdgpu_status mmdgpu_host_alloc(size_t bytes, void** out) {
    if (!out) return set_err(nullptr, MMDGPU_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 16, cudaHostAllocDefault);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaHostAlloc");
    return MMDGPU_OK;
}
MMDGPU_API void mmdgpu_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"
