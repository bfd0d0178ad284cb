"""A short slice of tools/gpu_fuzz.py inside the GPU test-suite: a few seeds of every phase (random rigs, CCD IK chain
sets, skeleton topologies, morph graphs, key-frame structures, crowds, extension mode), each compared bit-for-bit with
the CPU oracle.  The long runs are logged in profiles/r01_experiments.md."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import gpu_fuzz  # noqa: E402

pytestmark = pytest.mark.gpu
PHASES = {"rig": gpu_fuzz.rig_case, "ik": gpu_fuzz.ik_case, "topo": gpu_fuzz.topo_case, "morph": gpu_fuzz.morph_case,
          "motion": gpu_fuzz.motion_case, "crowd": gpu_fuzz.crowd_case, "ext": gpu_fuzz.ext_case}


@pytest.mark.parametrize("phase", sorted(PHASES))
def test_fuzz_slice(ctx, phase):
    gpu_fuzz.ctx = ctx
    for seed in range(40000, 40006):
        ok, what = PHASES[phase](seed)
        assert ok, f"{phase} seed {seed}: {what}"
