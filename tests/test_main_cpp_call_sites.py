"""The drop-in claim of INTEGRATION.md section 2, compiled instead of asserted: the reference's own frame-loop text
(main.cpp:1786-1825) and its vertex upload (main.cpp:820-863), extracted from /root/reference AT TEST TIME (never
committed), must build unchanged against include/mmdgpu.hpp once g_state.poser / g_state.motion_player are the mmdgpu
types.  main.cpp keeps libmmd's mmd::Model for everything else, so libmmd's real headers are on the include path; the
sokol calls and the application state are stubs.  No GPU is used: the program is compiled and linked, and its main()
does not create a context."""
import os
import subprocess

import pytest

from conftest import ROOT
from simple_mmd_renderer_b200 import lib

MAIN_CPP = "/root/reference/main.cpp"
LIBMMD_INC = "/root/reference/3rd_party/libmmd/include"

PRELUDE = r'''
#include <math.h>
#include <stdlib.h>
#include <cstdint>
#include <cstddef>
#include <iostream>
#include <memory>
#include <vector>
#include <mmd/mmd.hxx>          // the reference keeps mmd::Model / mmd::Motion (triangles, materials, UVs, sequencer)
#include <mmdgpu.hpp>           // ... and swaps Poser / MotionPlayer for the GPU ones

// ---- stubs for what is not on the deformation path
struct sg_buffer { uint32_t id; };
struct sg_range { const void* ptr; size_t size; };
static size_t g_uploaded_bytes = 0;
static void sg_update_buffer(sg_buffer, const sg_range& r) { g_uploaded_bytes = r.size; }
struct PhysicsReactorStub { void React(float) {} };
'''

STATE = r'''
static struct AppState {
    std::shared_ptr<mmd::Model> model;                       // main.cpp keeps libmmd's model
    std::unique_ptr<mmdgpu::Poser> poser;                    // was std::unique_ptr<mmd::Poser>        (main.cpp:663)
    std::unique_ptr<mmdgpu::MotionPlayer> motion_player;     // was std::unique_ptr<mmd::MotionPlayer> (main.cpp:676,717)
    std::unique_ptr<PhysicsReactorStub> physics_reactor;
    bool model_loaded = false, motion_loaded = false, physics_enabled = false;
    sg_buffer vertex_buffer{0};
    float time = 0.0f;
} g_state;
'''


def _lines(text, lo, hi):
    """1-based inclusive line range."""
    return "\n".join(text.splitlines()[lo - 1:hi]) + "\n"


@pytest.mark.skipif(not os.path.exists(MAIN_CPP), reason="needs the reference sources (/root/reference is not on the GPU box)")
def test_reference_frame_loop_and_vertex_upload_compile_unchanged(tmp_path):
    src = open(MAIN_CPP, encoding="utf-8", errors="replace").read()
    vertex_struct = _lines(src, 49, 54)
    upload = _lines(src, 820, 863)
    loop = _lines(src, 1784, 1825)
    assert "struct Vertex" in vertex_struct and "float uv[2];" in vertex_struct
    assert "void UpdateDeformedVertices() {" in upload and "sg_update_buffer" in upload and upload.count("{") == upload.count("}")
    assert "g_state.poser->ResetPosing();" in loop and "g_state.motion_player->SeekFrame(frame);" in loop
    assert "g_state.poser->Deform();" in loop and "UpdateDeformedVertices();" in loop
    assert loop.count("{") == loop.count("}"), "the extracted block must be balanced"
    tu = (PRELUDE + vertex_struct + STATE + upload +
          "static void frame_callback_body() {\n" + loop + "}\n"
          "int main(int argc, char**) { if (argc > 1000) frame_callback_body(); return int(g_uploaded_bytes); }\n")
    path = tmp_path / "call_sites.cc"
    path.write_text(tu)
    exe = tmp_path / "call_sites"
    r = subprocess.run(["g++", "-std=c++17", "-O0", "-w", f"-I{LIBMMD_INC}", f"-I{ROOT}/include", str(path), lib.SO_PATH,
                        f"-Wl,-rpath,{os.path.dirname(lib.SO_PATH)}", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
