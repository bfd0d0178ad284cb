// Host-side memory-safety check (compute-sanitizer is closed on the GPU pool; this covers everything that runs
// before the first kernel): PMX / VMD parsing, plan building and motion flattening under AddressSanitizer + UBSan,
// on a valid stream and on truncated / bit-flipped mutations of it.
//
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -I include \
//       tools/fuzz_host.cc simple_mmd_renderer_b200/csrc/host_plan.cpp simple_mmd_renderer_b200/csrc/pmx_vmd.cpp -o /tmp/fuzz_host
//   /tmp/fuzz_host model.pmx motion.vmd 2000
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "../simple_mmd_renderer_b200/csrc/host_plan.hpp"

static std::vector<unsigned char> slurp(const char* path) {
    std::vector<unsigned char> b;
    if (FILE* f = std::fopen(path, "rb")) {
        unsigned char tmp[65536];
        size_t n;
        while ((n = std::fread(tmp, 1, sizeof tmp, f)) > 0) b.insert(b.end(), tmp, tmp + n);
        std::fclose(f);
    }
    return b;
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const std::vector<unsigned char> pmx = slurp(argv[1]), vmd = slurp(argv[2]);
    const int trials = std::atoi(argv[3]);
    if (pmx.empty() || vmd.empty()) return 3;
    std::mt19937 rng(12345);
    int ok = 0, rejected = 0, anim_ok = 0;
    mmdgpu::Plan good;
    {
        mmdgpu::ParsedModel pm;
        std::string err;
        if (mmdgpu::parse_pmx(pmx.data(), pmx.size(), pm, err) != MMDGPU_OK || mmdgpu::build_plan(pm.desc, nullptr, good, err) != MMDGPU_OK) {
            std::fprintf(stderr, "the unmodified stream must parse: %s\n", err.c_str());
            return 4;
        }
        good.bone_names = pm.bone_names; good.morph_names = pm.morph_names; good.names_utf8 = pm.utf8;
    }
    for (int t = 0; t < trials; ++t) {
        std::vector<unsigned char> buf = (t & 1) ? vmd : pmx;
        const int mode = (t >> 1) % 3;
        if (t >= 2) {
            if (mode == 0) buf.resize(rng() % buf.size());
            else if (mode == 1) for (int k = 0, n = 1 + rng() % 8; k < n; ++k) buf[rng() % buf.size()] = (unsigned char)rng();
            else { size_t at = rng() % std::min<size_t>(buf.size() - 4, 8192); for (int k = 0; k < 4; ++k) buf[at + k] = (unsigned char)rng(); }
        }
        std::string err;
        if (t & 1) {
            mmdgpu::ParsedMotion mo;
            if (mmdgpu::parse_vmd(buf.data(), buf.size(), good, mo, err) == MMDGPU_OK) {
                mmdgpu::HostAnim a;
                if (mmdgpu::build_anim(mo.desc, good.nb, good.nm, a, err) == MMDGPU_OK) ++anim_ok; else ++rejected;
            } else ++rejected;
        } else {
            mmdgpu::ParsedModel pm;
            mmdgpu::Plan p;
            if (mmdgpu::parse_pmx(buf.data(), buf.size(), pm, err) == MMDGPU_OK && mmdgpu::build_plan(pm.desc, nullptr, p, err) == MMDGPU_OK) ++ok;
            else ++rejected;
        }
    }
    std::printf("fuzz_host: %d trials, %d models accepted, %d motions accepted, %d rejected, no sanitizer report\n", trials, ok, anim_ok, rejected);
    return 0;
}
