// headless_update — the frame loop of main.cpp:1786-1825 (physics off) on the mmdgpu C++ shim, without a window.
//
//   headless_update model.pmx motion.vmd first_frame n_frames [out.bin]
//
// Drives the libmmd-named call sequence once per frame
//     ResetPosing(); SeekFrame(f); PrePhysicsPosing(); PostPhysicsPosing(); Deform();
// reads pose_image like UpdateDeformedVertices (main.cpp:820-863) does, and writes the last frame's
// coordinates + normals (n x 3 floats each) to out.bin so that a test can compare them with libmmd's.
//
// build: g++ -std=c++14 -O2 -Iinclude examples/headless_update.cc simple_mmd_renderer_b200/libmmdgpu.so -o headless_update
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "mmdgpu.hpp"

int main(int argc, char** argv) {
    if (argc < 5) {
        std::fprintf(stderr, "usage: %s model.pmx motion.vmd first_frame n_frames [out.bin]\n", argv[0]);
        return 2;
    }
    try {
        const std::vector<unsigned char> pmx = mmdgpu::ReadFile(argv[1]), vmd = mmdgpu::ReadFile(argv[2]);
        const size_t first = std::strtoul(argv[3], nullptr, 10), count = std::strtoul(argv[4], nullptr, 10);
        mmdgpu::Context ctx(0);
        mmdgpu::Model model(ctx, pmx.data(), pmx.size());
        mmdgpu::Motion motion(model, vmd.data(), vmd.size());
        mmdgpu::Poser poser(model);
        mmdgpu::MotionPlayer player(motion, poser);
        std::printf("model: %zu vertices, %zu bones, %zu morphs; motion length %zu frames\n", model.GetVertexNum(),
                    model.GetBoneNum(), model.GetMorphNum(), motion.GetLength());
        double checksum = 0.0;
        const auto t0 = std::chrono::steady_clock::now();
        for (size_t f = first; f < first + count; ++f) {
            poser.ResetPosing();
            player.SeekFrame(f);
            poser.PrePhysicsPosing();
            poser.PostPhysicsPosing();
            poser.Deform();
            const size_t n = poser.pose_image.coordinates.size();
            if (n) checksum += poser.pose_image.coordinates[f % n].p.x + poser.pose_image.normals[(7 * f) % n].p.y;
        }
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("%zu frames in %.3f ms (%.1f us / frame incl. download), checksum %.6f\n", count, sec * 1e3,
                    count ? sec * 1e6 / double(count) : 0.0, checksum);
        if (argc > 5) {
            std::FILE* o = std::fopen(argv[5], "wb");
            if (!o) return 3;
            const size_t n = poser.pose_image.coordinates.size();
            std::fwrite(poser.pose_image.coordinates.data(), sizeof(mmdgpu::Vector3f), n, o);
            std::fwrite(poser.pose_image.normals.data(), sizeof(mmdgpu::Vector3f), n, o);
            std::fclose(o);
        }
    } catch (const mmdgpu::Error& e) {
        std::fprintf(stderr, "mmdgpu error %d: %s\n", e.status, e.what());
        return 1;
    }
    return 0;
}
