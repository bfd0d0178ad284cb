// mmdgpu.hpp — header-only C++ mirror of the libmmd classes simple_mmd_renderer's frame loop drives
// (main.cpp:1786-1825, :820-863), implemented on the mmdgpu_* C-ABI (mmdgpu.h).
//
//   libmmd (L/ = 3rd_party/libmmd/include/mmd/)                     here
//   --------------------------------------------------------------  ---------------------------------------------
//   mmd::Model + PmxReader::ReadModel   L/reader/pmx_reader_impl.inl  mmdgpu::Model        (PMX bytes or flat arrays)
//   mmd::Motion + VmdReader::ReadMotion L/reader/vmd_reader_impl.inl  mmdgpu::Motion       (VMD bytes or flat arrays)
//   mmd::Poser                          L/motion/poser.inl:15-45      mmdgpu::Poser        same method names, pose_image
//   mmd::MotionPlayer                   L/motion/poser.inl:184-198    mmdgpu::MotionPlayer SeekFrame / SeekTime
//
// Error behaviour: libmmd's engine methods are void and never throw; its readers throw mmd::exception.  Here the
// constructors that parse or allocate throw mmdgpu::Error; the per-frame methods throw only on a CUDA failure
// (there is no CPU fallback to continue on).
#ifndef MMDGPU_HPP_INCLUDED
#define MMDGPU_HPP_INCLUDED

#include <cstddef>
#include <cstdlib>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "mmdgpu.h"

namespace mmdgpu {

class Error : public std::runtime_error {
public:
    Error(mmdgpu_status s, const std::string& what) : std::runtime_error(what), status(s) {}
    mmdgpu_status status;
};

// Element type of Poser::pose_image.  In a translation unit that has libmmd's headers (main.cpp keeps them for mmd::Model)
// it IS mmd::Vector3f, so that `const mmd::Vector3f& pos = poser->pose_image.coordinates[i];` (main.cpp:843-844)
// compiles unchanged; elsewhere a 12-byte struct with the member spelling main.cpp reads (pos.p.x, main.cpp:848-854).
#ifdef __MMD_H_7F46DEA0A2C1F5902D557E3545B096B5_INCLUDED__
using Vector3f = mmd::Vector3f;
#else
struct Vector3f {
    struct { float x, y, z; } p;
};
#endif
static_assert(sizeof(Vector3f) == 12, "Vector3f must be three packed floats");

class Context {
public:
    explicit Context(int device = 0, void* cuda_stream = nullptr) {
        mmdgpu_status s = mmdgpu_context_create(device, cuda_stream, &h_);
        if (s != MMDGPU_OK) throw Error(s, std::string("mmdgpu_context_create: ") + mmdgpu_last_error(nullptr));
    }
    ~Context() { mmdgpu_context_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    mmdgpu_context_t handle() const { return h_; }
    void check(mmdgpu_status s, const char* what) const {
        if (s != MMDGPU_OK) throw Error(s, std::string(what) + ": " + mmdgpu_last_error(h_));
    }
    void Synchronize() const { check(mmdgpu_context_synchronize(h_), "mmdgpu_context_synchronize"); }

private:
    mmdgpu_context_t h_ = nullptr;
};

inline std::vector<unsigned char> ReadFile(const std::string& path) {
    std::FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw Error(MMDGPU_ERR_INVALID_ARG, "cannot open " + path);
    std::vector<unsigned char> buf;
    unsigned char tmp[1 << 16];
    size_t n;
    while ((n = std::fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    std::fclose(f);
    return buf;
}

class Model {
public:
    Model(Context& ctx, const void* pmx_bytes, size_t n, const mmdgpu_options* opt = nullptr) : ctx_(ctx) {
        ctx_.check(mmdgpu_model_create_from_pmx(ctx.handle(), pmx_bytes, n, opt, &h_), "mmdgpu_model_create_from_pmx");
    }
    Model(Context& ctx, const mmdgpu_model_desc& desc, const mmdgpu_options* opt = nullptr) : ctx_(ctx) {
        ctx_.check(mmdgpu_model_create_from_arrays(ctx.handle(), &desc, opt, &h_), "mmdgpu_model_create_from_arrays");
    }
    ~Model() { mmdgpu_model_destroy(h_); }
    Model(const Model&) = delete;
    Model& operator=(const Model&) = delete;
    size_t GetVertexNum() const { return mmdgpu_model_vertex_count(h_); }
    size_t GetBoneNum() const { return mmdgpu_model_bone_count(h_); }
    size_t GetMorphNum() const { return mmdgpu_model_morph_count(h_); }
    size_t GetMaterialNum() const { return mmdgpu_model_material_count(h_); }
    mmdgpu_model_t handle() const { return h_; }
    Context& context() const { return ctx_; }

private:
    Context& ctx_;
    mmdgpu_model_t h_ = nullptr;
};

class Motion {
public:
    Motion(Model& model, const void* vmd_bytes, size_t n) : model_(model) {
        model.context().check(mmdgpu_animation_create_from_vmd(model.context().handle(), model.handle(), vmd_bytes, n, &h_),
                              "mmdgpu_animation_create_from_vmd");
    }
    Motion(Model& model, const mmdgpu_anim_desc& desc) : model_(model) {
        model.context().check(mmdgpu_animation_create_from_arrays(model.context().handle(), model.handle(), &desc, &h_),
                              "mmdgpu_animation_create_from_arrays");
    }
    ~Motion() { mmdgpu_animation_destroy(h_); }
    Motion(const Motion&) = delete;
    Motion& operator=(const Motion&) = delete;
    size_t GetLength() const { return mmdgpu_animation_length(h_); }  // Motion::GetLength, motion_impl.inl:242-245
    mmdgpu_animation_t handle() const { return h_; }
    Model& model() const { return model_; }

private:
    Model& model_;
    mmdgpu_animation_t h_ = nullptr;
};

class Poser {
public:
    // Poser::pose_image (L/motion/poser.inl:17-20).  The two vectors live in one page-locked host block that the
    // skinning kernel writes DIRECTLY (mmdgpu_frames_bind_output on host memory): Deform() is the transfer, the first
    // access afterwards only waits for that kernel (mmdgpu_frames_wait_skinning).
    class LazyVectors {
    public:
        size_t size() const { return owner_->model_.GetVertexNum(); }
        const Vector3f& operator[](size_t i) const { return data()[i]; }
        const Vector3f* data() const {
            owner_->Fetch();
            return host_;
        }
        const Vector3f* begin() const { return data(); }
        const Vector3f* end() const { return data() + size(); }

    private:
        friend class Poser;
        explicit LazyVectors(Poser* owner) : owner_(owner) {}
        Poser* owner_;
        Vector3f* host_ = nullptr;
    };
    struct PoseImage {
        LazyVectors coordinates, normals;
    } pose_image;

    // Poser::Poser (poser_impl.inl:16-128) ends with ResetPosing(); Deform();
    explicit Poser(Model& model, mmdgpu_layout layout = MMDGPU_LAYOUT_SOA_POS_NRM)
        : pose_image{LazyVectors(this), LazyVectors(this)},
          model_(model),
          layout_(layout) {
        check(mmdgpu_frames_create(model.context().handle(), model.handle(), 1, 1, layout, &frames_), "mmdgpu_frames_create");
        // one page-locked block: coordinates, then normals at the next 16-byte boundary (bound outputs are 16-byte aligned)
        const size_t bytes = model.GetVertexNum() * sizeof(Vector3f), nrm_at = (bytes + 15) / 16 * 16;
        void* a = nullptr;
        if (mmdgpu_host_alloc(nrm_at + bytes, &a) != MMDGPU_OK) {
            mmdgpu_frames_destroy(frames_);
            throw Error(MMDGPU_ERR_OOM, "page-locked pose_image allocation failed");
        }
        pose_image.coordinates.host_ = static_cast<Vector3f*>(a);
        pose_image.normals.host_ = reinterpret_cast<Vector3f*>(static_cast<char*>(a) + nrm_at);
        if (layout == MMDGPU_LAYOUT_SOA_POS_NRM && bytes) {
            // MMDGPU_POSE_IMAGE_COPY=1 keeps the deformed buffer on the device and copies it on first access instead
            const char* e = std::getenv("MMDGPU_POSE_IMAGE_COPY");
            if (!(e && e[0] == '1') &&
                mmdgpu_frames_bind_output(frames_, MMDGPU_STREAM_POSITION, pose_image.coordinates.host_, nrm_at) == MMDGPU_OK) {
                if (mmdgpu_frames_bind_output(frames_, MMDGPU_STREAM_NORMAL, pose_image.normals.host_, nrm_at) == MMDGPU_OK) host_bound_ = true;
                else mmdgpu_frames_bind_output(frames_, MMDGPU_STREAM_POSITION, nullptr, 0);
            }
        }
        ResetPosing();
        Deform();
    }
    ~Poser() {
        mmdgpu_frames_destroy(frames_);
        mmdgpu_host_free(pose_image.coordinates.host_);   // the block holds both mirrors
    }
    Poser(const Poser&) = delete;
    Poser& operator=(const Poser&) = delete;

    // poser_impl.inl:130-140 — zero rates, identity poses, then a full Pre + PostPhysicsPosing.
    //
    // main.cpp:1788-1810 calls ResetPosing(); SeekFrame(f); PrePhysicsPosing(); PostPhysicsPosing() every frame.  The
    // evaluation ResetPosing ends with is dead there: PrePhysicsPosing clears every bone's scratch state and Pre + Post
    // recompute all bones from the poses alone (poser_impl.inl:362-394).  The calls are therefore recorded, and when the
    // sequence Reset [, Seek], Pre, Post completes it is issued as ONE sampling launch (identity / zero for what the clip
    // does not animate, the sampled key frames for the rest) plus the two hierarchy passes.  Any other continuation
    // (Deform right after ResetPosing, SetBonePose, a download, Pre without Post ...) replays the recorded calls exactly
    // as libmmd would have run them, so every observable state is libmmd's.
    void ResetPosing() {
        Flush();
        pending_ = kReset;
    }
    void SetBonePose(size_t index, const float translation[3], const float rotation_xyzw[4]) {
        Flush();
        check(mmdgpu_set_bone_pose(frames_, 0, uint32_t(index), translation, rotation_xyzw), "mmdgpu_set_bone_pose");
    }
    void SetMorphPose(size_t index, float weight) {
        Flush();
        check(mmdgpu_set_morph_pose(frames_, 0, uint32_t(index), weight), "mmdgpu_set_morph_pose");
    }
    void PrePhysicsPosing() {
        if (pending_ == kReset || pending_ == (kReset | kSeek)) { pending_ |= kPre; return; }
        Flush();
        check(mmdgpu_pre_physics_posing(frames_), "mmdgpu_pre_physics_posing");
    }
    void PostPhysicsPosing() {
        if (pending_ & kPre) {   // Reset [, Seek], Pre, Post: nothing of ResetPosing's own evaluation survives
            const unsigned had = pending_;
            pending_ = kNone;
            if (had & kSeek) {   // one call: one sampling launch + one hierarchy pass
                if (seek_by_time_) check(mmdgpu_pose_time(frames_, &seek_anim_, &seek_seconds_), "mmdgpu_pose_time");
                else check(mmdgpu_pose_frame(frames_, &seek_anim_, &seek_frame_), "mmdgpu_pose_frame");
                return;
            }
            check(mmdgpu_reset_posing(frames_), "mmdgpu_reset_posing");
            check(mmdgpu_pre_physics_posing(frames_), "mmdgpu_pre_physics_posing");
            check(mmdgpu_post_physics_posing(frames_), "mmdgpu_post_physics_posing");
            return;
        }
        Flush();
        check(mmdgpu_post_physics_posing(frames_), "mmdgpu_post_physics_posing");
    }
    void Deform() {
        Flush();
        check(mmdgpu_deform(frames_), "mmdgpu_deform");
        image_valid_ = false;
    }
    // ResetPosing + SeekFrame + Pre + Post + Deform in one call (the fused path; physics off).
    void Update(const Motion& motion, size_t frame) {
        pending_ = kNone;        // everything recorded so far is overwritten by the fused update
        mmdgpu_animation_t a = motion.handle();
        const uint32_t f = uint32_t(frame);
        check(mmdgpu_update(frames_, &a, &f), "mmdgpu_update");
        image_valid_ = false;
    }
    // Physics hand-back between Pre and Post (PoserMotionState::Synchronize / Fix, mmd-bullet_impl.inl:34-56).
    void OverrideSkinningMatrix(size_t bone, const float skinning[16], const float* local_or_null = nullptr) {
        Flush();   // with physics on, the recorded calls run exactly as libmmd's (the host Bullet step dominates the frame anyway)
        check(mmdgpu_set_skinning_matrix_override(frames_, 0, uint32_t(bone), skinning, local_or_null),
              "mmdgpu_set_skinning_matrix_override");
    }
    // The 32-byte Vertex{pos*0.1f, normal, uv} records main.cpp:838-859 builds, ready for sg_update_buffer
    // (requires layout = MMDGPU_LAYOUT_INTERLEAVED_SOKOL32).
    void DownloadInterleaved(void* dst_vertices) {
        Flush();
        check(mmdgpu_frames_download(frames_, 0, MMDGPU_STREAM_INTERLEAVED, dst_vertices, model_.GetVertexNum() * 32),
              "mmdgpu_frames_download");
    }
    // Poser::material_mul_images_ / material_add_images_ (poser.inl:160-161): n_materials x 2 x 28 floats; all 1 / all 0
    // unless the model was created with extensions (libmmd never fills them).
    void DownloadMaterialImages(float* dst) { Flush(); check(mmdgpu_material_images_download(frames_, 0, dst), "mmdgpu_material_images_download"); }
    void DownloadSkinningMatrices(float* dst_nb_x_16) { Flush(); check(mmdgpu_bone_matrices_download(frames_, 0, dst_nb_x_16), "mmdgpu_bone_matrices_download"); }
    const Model& GetModel() const { return model_; }
    Model& GetModel() { return model_; }
    // The frames object behind this Poser, with every recorded call issued (for mmdgpu_frames_bind_output and friends).
    mmdgpu_frames_t frames() { Flush(); return frames_; }

private:
    friend class MotionPlayer;
    enum : unsigned { kNone = 0, kReset = 1, kSeek = 2, kPre = 4 };
    void check(mmdgpu_status s, const char* what) const { model_.context().check(s, what); }
    // MotionPlayer::SeekFrame / SeekTime land here
    void Seek(mmdgpu_animation_t anim, bool by_time, uint32_t frame, double seconds) {
        if (pending_ != kReset) Flush();
        seek_anim_ = anim; seek_by_time_ = by_time; seek_frame_ = frame; seek_seconds_ = seconds;
        if (pending_ == kReset) { pending_ |= kSeek; return; }
        IssueSeek(false);
    }
    void IssueSeek(bool with_reset) {
        if (seek_by_time_)
            check(with_reset ? mmdgpu_reset_and_seek_time(frames_, &seek_anim_, &seek_seconds_) : mmdgpu_seek_time(frames_, &seek_anim_, &seek_seconds_),
                  "mmdgpu_seek_time");
        else
            check(with_reset ? mmdgpu_reset_and_seek_frame(frames_, &seek_anim_, &seek_frame_) : mmdgpu_seek_frame(frames_, &seek_anim_, &seek_frame_),
                  "mmdgpu_seek_frame");
    }
    // Issue the recorded calls one by one, as libmmd would have executed them.
    void Flush() {
        const unsigned had = pending_;
        pending_ = kNone;
        if (had & kReset) {
            check(mmdgpu_reset_posing(frames_), "mmdgpu_reset_posing");
            check(mmdgpu_pre_physics_posing(frames_), "mmdgpu_pre_physics_posing");
            check(mmdgpu_post_physics_posing(frames_), "mmdgpu_post_physics_posing");
        }
        if (had & kSeek) IssueSeek(false);
        if (had & kPre) check(mmdgpu_pre_physics_posing(frames_), "mmdgpu_pre_physics_posing");
    }
    // both planes of pose_image in one round trip (SoA layout; the interleaved layout has DownloadInterleaved)
    void Fetch() {
        Flush();
        if (image_valid_) return;
        if (host_bound_) {
            // still our binding?  (a caller may have re-bound the planes through frames(); then pose_image is a copy again)
            void *p = nullptr, *n = nullptr;
            if (mmdgpu_frames_device_ptr(frames_, MMDGPU_STREAM_POSITION, &p, nullptr) != MMDGPU_OK || p != pose_image.coordinates.host_ ||
                mmdgpu_frames_device_ptr(frames_, MMDGPU_STREAM_NORMAL, &n, nullptr) != MMDGPU_OK || n != pose_image.normals.host_)
                host_bound_ = false;
        }
        if (host_bound_) {
            check(mmdgpu_frames_wait_skinning(frames_), "mmdgpu_frames_wait_skinning");   // the kernel wrote pose_image itself
        } else if (layout_ == MMDGPU_LAYOUT_SOA_POS_NRM) {
            const size_t bytes = model_.GetVertexNum() * sizeof(Vector3f);
            if (bytes % 16 == 0) {   // the planes are adjacent on the device and on the host: one transfer
                check(mmdgpu_frames_download_pair_async(frames_, 0, pose_image.coordinates.host_, 2 * bytes), "mmdgpu_frames_download_pair_async");
            } else {
                check(mmdgpu_frames_download_async(frames_, 0, 1, MMDGPU_STREAM_POSITION, pose_image.coordinates.host_, bytes), "mmdgpu_frames_download_async");
                check(mmdgpu_frames_download_async(frames_, 0, 1, MMDGPU_STREAM_NORMAL, pose_image.normals.host_, bytes), "mmdgpu_frames_download_async");
            }
            check(mmdgpu_frames_wait_downloads(frames_), "mmdgpu_frames_wait_downloads");   // these copies only, not the whole context
        } else {
            throw Error(MMDGPU_ERR_INVALID_ARG, "pose_image needs the SoA layout; interleaved Posers use DownloadInterleaved or a bound output");
        }
        image_valid_ = true;
    }
    Model& model_;
    mmdgpu_layout layout_;
    mmdgpu_frames_t frames_ = nullptr;
    bool image_valid_ = false;
    bool host_bound_ = false;                     // the skinning kernel writes pose_image's host block directly
    unsigned pending_ = kNone;
    mmdgpu_animation_t seek_anim_ = nullptr;
    bool seek_by_time_ = false;
    uint32_t seek_frame_ = 0;
    double seek_seconds_ = 0.0;
};

class MotionPlayer {
public:
    MotionPlayer(const Motion& motion, Poser& poser) : motion_(motion), poser_(poser) {
        if (&motion.model() != &poser.GetModel()) throw Error(MMDGPU_ERR_INVALID_ARG, "motion and poser use different models");
    }
    // MotionPlayer::SeekFrame(size_t), poser_impl.inl:539-546
    void SeekFrame(size_t frame) { poser_.Seek(motion_.handle(), false, uint32_t(frame), 0.0); }
    // MotionPlayer::SeekTime(double), poser_impl.inl:548-555 (main.cpp itself only uses SeekFrame)
    void SeekTime(double time_seconds) { poser_.Seek(motion_.handle(), true, 0u, time_seconds); }

private:
    const Motion& motion_;
    Poser& poser_;
};

}  // namespace mmdgpu

#endif  // MMDGPU_HPP_INCLUDED
