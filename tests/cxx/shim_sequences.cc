// Call sequences on mmdgpu::Poser (include/mmdgpu.hpp) against the same calls issued eagerly through the C-ABI.
// The shim records ResetPosing / SeekFrame / PrePhysicsPosing and issues them fused once PostPhysicsPosing completes
// main.cpp's sequence; any other continuation must replay them exactly.  Every sequence below ends in the same observable
// state both ways or the program exits non-zero.   usage: shim_sequences model.pmx motion.vmd
#include <cstdio>
#include <cstring>
#include <functional>
#include <vector>

#include "mmdgpu.hpp"

struct Eager {   // libmmd's semantics, one C-ABI call per libmmd call (ResetPosing = reset + Pre + Post)
    mmdgpu::Context& ctx;
    mmdgpu_frames_t f = nullptr;
    mmdgpu_animation_t anim;
    Eager(mmdgpu::Context& c, mmdgpu::Model& m, mmdgpu_animation_t a) : ctx(c), anim(a) {
        ctx.check(mmdgpu_frames_create(c.handle(), m.handle(), 1, 1, MMDGPU_LAYOUT_SOA_POS_NRM, &f), "frames_create");
        Reset();
        Deform();
    }
    ~Eager() { mmdgpu_frames_destroy(f); }
    void Reset() { ctx.check(mmdgpu_reset_posing(f), "reset"); Pre(); Post(); }
    void Seek(uint32_t fr) { ctx.check(mmdgpu_seek_frame(f, &anim, &fr), "seek"); }
    void SeekTime(double t) { ctx.check(mmdgpu_seek_time(f, &anim, &t), "seek_time"); }
    void Pre() { ctx.check(mmdgpu_pre_physics_posing(f), "pre"); }
    void Post() { ctx.check(mmdgpu_post_physics_posing(f), "post"); }
    void Deform() { ctx.check(mmdgpu_deform(f), "deform"); }
    void Bone(uint32_t b, const float* T, const float* R) { ctx.check(mmdgpu_set_bone_pose(f, 0, b, T, R), "bone"); }
};

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s model.pmx motion.vmd\n", argv[0]); return 2; }
    try {
        const auto pmx = mmdgpu::ReadFile(argv[1]), vmd = mmdgpu::ReadFile(argv[2]);
        mmdgpu::Context ctx(0);
        mmdgpu::Model model(ctx, pmx.data(), pmx.size());
        mmdgpu::Motion motion(model, vmd.data(), vmd.size());
        mmdgpu::Poser poser(model);
        mmdgpu::MotionPlayer player(motion, poser);
        Eager eager(ctx, model, motion.handle());
        const size_t nv = model.GetVertexNum(), nb = model.GetBoneNum();
        std::vector<float> pa(nv * 3), na(nv * 3), ma(nb * 16), mb(nb * 16);
        int failures = 0;
        auto compare = [&](const char* what) {
            ctx.check(mmdgpu_frames_download(eager.f, 0, MMDGPU_STREAM_POSITION, pa.data(), nv * 12), "download");
            ctx.check(mmdgpu_frames_download(eager.f, 0, MMDGPU_STREAM_NORMAL, na.data(), nv * 12), "download");
            ctx.check(mmdgpu_bone_matrices_download(eager.f, 0, ma.data()), "matrices");
            poser.DownloadSkinningMatrices(mb.data());
            const bool ok = std::memcmp(pa.data(), poser.pose_image.coordinates.data(), nv * 12) == 0 &&
                            std::memcmp(na.data(), poser.pose_image.normals.data(), nv * 12) == 0 &&
                            std::memcmp(ma.data(), mb.data(), nb * 64) == 0;
            std::printf("%-58s %s\n", what, ok ? "same" : "DIFFERENT");
            failures += !ok;
        };
        const float T[3] = {0.3f, -0.2f, 0.1f}, R[4] = {0.1f, 0.2f, -0.1f, 0.9695360f};
        // 1. main.cpp's frame: the fused path of the shim
        for (uint32_t fr : {7u, 33u, 88u}) {
            poser.ResetPosing(); player.SeekFrame(fr); poser.PrePhysicsPosing(); poser.PostPhysicsPosing(); poser.Deform();
            eager.Reset(); eager.Seek(fr); eager.Pre(); eager.Post(); eager.Deform();
            compare("Reset, Seek, Pre, Post, Deform");
        }
        // 2. ResetPosing observed directly
        poser.ResetPosing(); poser.Deform();
        eager.Reset(); eager.Deform();
        compare("Reset, Deform");
        // 3. sought poses never evaluated: Deform sees ResetPosing's own evaluation
        poser.ResetPosing(); player.SeekFrame(40); poser.Deform();
        eager.Reset(); eager.Seek(40); eager.Deform();
        compare("Reset, Seek, Deform");
        // 4. Pre without Post: post-physics bones keep the reset-pose evaluation
        poser.ResetPosing(); player.SeekFrame(41); poser.PrePhysicsPosing(); poser.Deform();
        eager.Reset(); eager.Seek(41); eager.Pre(); eager.Deform();
        compare("Reset, Seek, Pre, Deform");
        // 5. manual posing between Reset and Pre
        poser.ResetPosing(); poser.SetBonePose(3, T, R); poser.PrePhysicsPosing(); poser.PostPhysicsPosing(); poser.Deform();
        eager.Reset(); eager.Bone(3, T, R); eager.Pre(); eager.Post(); eager.Deform();
        compare("Reset, SetBonePose, Pre, Post, Deform");
        // 6. manual posing after the seek
        poser.ResetPosing(); player.SeekFrame(12); poser.SetBonePose(5, T, R); poser.PrePhysicsPosing(); poser.PostPhysicsPosing(); poser.Deform();
        eager.Reset(); eager.Seek(12); eager.Bone(5, T, R); eager.Pre(); eager.Post(); eager.Deform();
        compare("Reset, Seek, SetBonePose, Pre, Post, Deform");
        // 7. no reset at all: poses of untracked bones persist from the previous frame
        player.SeekFrame(60); poser.PrePhysicsPosing(); poser.PostPhysicsPosing(); poser.Deform();
        eager.Seek(60); eager.Pre(); eager.Post(); eager.Deform();
        compare("Seek, Pre, Post, Deform (no reset)");
        // 8. sub-frame time
        poser.ResetPosing(); player.SeekTime(1.2345); poser.PrePhysicsPosing(); poser.PostPhysicsPosing(); poser.Deform();
        eager.Reset(); eager.SeekTime(1.2345); eager.Pre(); eager.Post(); eager.Deform();
        compare("Reset, SeekTime, Pre, Post, Deform");
        // 9. two resets and two seeks in a row
        poser.ResetPosing(); poser.ResetPosing(); player.SeekFrame(5); player.SeekFrame(70); poser.PrePhysicsPosing(); poser.PostPhysicsPosing(); poser.Deform();
        eager.Reset(); eager.Reset(); eager.Seek(5); eager.Seek(70); eager.Pre(); eager.Post(); eager.Deform();
        compare("Reset, Reset, Seek, Seek, Pre, Post, Deform");
        // 10. physics hand-back between Pre and Post
        {
            std::vector<float> m16(16, 0.f);
            m16[0] = m16[5] = m16[10] = m16[15] = 1.f; m16[12] = 0.5f; m16[13] = -1.f;
            poser.ResetPosing(); player.SeekFrame(22); poser.PrePhysicsPosing(); poser.OverrideSkinningMatrix(2, m16.data(), m16.data()); poser.PostPhysicsPosing(); poser.Deform();
            eager.Reset(); eager.Seek(22); eager.Pre();
            ctx.check(mmdgpu_set_skinning_matrix_override(eager.f, 0, 2, m16.data(), m16.data()), "override");
            eager.Post(); eager.Deform();
            compare("Reset, Seek, Pre, Override, Post, Deform");
        }
        return failures ? 1 : 0;
    } catch (const mmdgpu::Error& e) {
        std::fprintf(stderr, "mmdgpu error %d: %s\n", e.status, e.what());
        return 3;
    }
}
