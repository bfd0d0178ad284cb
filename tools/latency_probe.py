"""GPU box: per-frame latency of the interactive (one Poser, one frame per call) path on a C1-sized model, through
the C++ shim (examples/headless_update.cc), next to libmmd's single-thread time for the same frames."""
import os, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import pmxio
import oracle
from simple_mmd_renderer_b200 import lib, synth

cfg = synth.C1
model = synth.make_model(cfg)
motion = synth.make_motion(cfg, model)
d = tempfile.mkdtemp()
open(os.path.join(d, "m.pmx"), "wb").write(pmxio.write_pmx(model))
open(os.path.join(d, "m.vmd"), "wb").write(pmxio.write_vmd(motion))
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(d, "headless_update")
subprocess.run(["g++", "-std=c++14", "-O2", f"-I{root}/include", f"{root}/examples/headless_update.cc", lib.SO_PATH, "-o", exe], check=True)
for mode, pdl, label in (("0", "1", "default: pose_image written by the skinning kernel (page-locked host block bound as its output), programmatic dependent launch"),
                         ("0", "0", "MMDGPU_PDL=0: ordinary launches"),
                         ("1", "1", "MMDGPU_POSE_IMAGE_COPY=1: pose_image copied from the device on first access"),
                         ("1", "0", "MMDGPU_POSE_IMAGE_COPY=1 MMDGPU_PDL=0 (the path of the previous measurement)")):
    for _ in range(3):
        r = subprocess.run([exe, os.path.join(d, "m.pmx"), os.path.join(d, "m.vmd"), "0", "300"], capture_output=True, text=True,
                           env=dict(os.environ, MMDGPU_POSE_IMAGE_COPY=mode, MMDGPU_PDL=pdl))
    print(label + ":\n  " + r.stdout.strip().replace("\n", "\n  "))
ses = oracle.Reference(model, motion) if oracle.have_reference() else oracle.Restatement(model, motion)
sec, _ = ses.time_frames(np.arange(300, dtype=np.uint32), 1)
print(f"libmmd CPU, 1 thread: 300 frames in {sec*1e3:.1f} ms ({sec/300*1e6:.0f} us / frame)")
