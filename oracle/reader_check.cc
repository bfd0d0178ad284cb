// TEST INFRASTRUCTURE.  Pins the product's PMX / VMD byte-stream parsers against libmmd's own readers.
//   libmmd PmxReader / VmdReader (L/reader/pmx_reader_impl.inl, vmd_reader_impl.inl)  ->  mmd::Model / mmd::Motion
//   -> include/mmdgpu_libmmd_adapter.hpp  ->  flat descriptors, dumped to <out> as (name, bytes) records.
// tests/test_reader_pin.py builds the host plan from that dump and from the same files' bytes and requires the
// two to be byte-identical.  Built only where the reference headers are mounted (oracle/Makefile ->
// oracle/_ref/reader_check); no reference source is copied.
#include <math.h>
#include <stdlib.h>

#include <mmd/mmd.hxx>

#include <clocale>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../include/mmdgpu_libmmd_adapter.hpp"

static FILE* g_out;
template <class T> static void put(const char* name, const std::vector<T>& v) {
    const uint32_t nl = uint32_t(std::strlen(name));
    const uint64_t nb = uint64_t(v.size()) * sizeof(T);
    std::fwrite(&nl, 4, 1, g_out);
    std::fwrite(name, 1, nl, g_out);
    std::fwrite(&nb, 8, 1, g_out);
    if (nb) std::fwrite(v.data(), 1, size_t(nb), g_out);
}
static void put_u32(const char* name, uint32_t x) { put(name, std::vector<uint32_t>(1, x)); }

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: reader_check model.pmx motion.vmd out.bin\n"); return 2; }
    std::setlocale(LC_ALL, "");
    try {
        mmd::Model model;
        {
            mmd::FileReader file{std::string(argv[1])};
            mmd::PmxReader reader(file);
            reader.ReadModel(model);
        }
        mmd::Motion motion;
        {
            mmd::FileReader file{std::string(argv[2])};
            mmd::VmdReader reader(file);
            reader.ReadMotion(motion);
        }
        mmdgpu::FlatModel m;
        mmdgpu::Flatten(model, m);
        mmdgpu::FlatMotion a;
        mmdgpu::Flatten(motion, model, a, mmdgpu::VmdTrackPrefix(motion));
        g_out = std::fopen(argv[3], "wb");
        if (!g_out) return 2;
        put_u32("n_vertices", m.desc.n_vertices); put_u32("n_bones", m.desc.n_bones); put_u32("n_ik_links", m.desc.n_ik_links);
        put_u32("n_morphs", m.desc.n_morphs); put_u32("n_materials", m.desc.n_materials);
        put_u32("has_sdef", m.desc.sdef_c ? 1u : 0u);
        put("position", m.position); put("normal", m.normal); put("uv", m.uv); put("skin_type", m.skin_type);
        put("bone_id", m.bone_id); put("weight", m.weight); put("sdef_c", m.sdef_c); put("sdef_r0", m.sdef_r0); put("sdef_r1", m.sdef_r1);
        put("bone_position", m.bone_position); put("bone_parent", m.bone_parent); put("bone_transform_level", m.bone_transform_level);
        put("bone_flags", m.bone_flags); put("bone_append_parent", m.bone_append_parent); put("bone_append_ratio", m.bone_append_ratio);
        put("ik_target", m.ik_target); put("ik_iterations", m.ik_iterations); put("ik_angle_limit", m.ik_angle_limit);
        put("ik_link_begin", m.ik_link_begin); put("ik_link_count", m.ik_link_count); put("ik_link_bone", m.ik_link_bone);
        put("ik_link_has_limit", m.ik_link_has_limit); put("ik_link_lo", m.ik_link_lo); put("ik_link_hi", m.ik_link_hi);
        put("morph_type", m.morph_type); put("morph_entry_begin", m.morph_entry_begin); put("morph_entry_count", m.morph_entry_count);
        put("vertex_morph_entries", m.vertex_morph_entries); put("uv_morph_entries", m.uv_morph_entries);
        put("bone_morph_entries", m.bone_morph_entries); put("group_morph_entries", m.group_morph_entries);
        put("material_morph_entries", m.material_morph_entries);
        put("bone_track_bone", a.bone_track_bone); put("bone_track_key_begin", a.bone_track_key_begin);
        put("bone_track_key_count", a.bone_track_key_count); put("bone_keys", a.bone_keys);
        put("morph_track_morph", a.morph_track_morph); put("morph_track_key_begin", a.morph_track_key_begin);
        put("morph_track_key_count", a.morph_track_key_count); put("morph_keys", a.morph_keys);
        std::fclose(g_out);
        std::printf("libmmd readers: %zu vertices, %zu bones, %zu morphs; %u bone tracks (%u keys), %u morph tracks (%u keys) joined\n",
                    model.GetVertexNum(), model.GetBoneNum(), model.GetMorphNum(), a.desc.n_bone_tracks, a.desc.n_bone_keys,
                    a.desc.n_morph_tracks, a.desc.n_morph_keys);
    } catch (const std::exception& e) {
        std::printf("EXCEPTION %s\n", e.what());
        return 1;
    } catch (...) {
        std::printf("EXCEPTION (non-standard)\n");
        return 1;
    }
    return 0;
}
