mkdir -p gpurun_out
(MMDGPU_FORCE_FALLBACKS=1 timeout 400 python tools/gpu_fuzz.py 4000 60 all 2>&1 | tail -8) | tee gpurun_out/r3c_fuzz_fallbacks.txt
(MMDGPU_SKIN_SCALAR=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -2) | tee gpurun_out/r3c_pytest_scalar.txt
(MMDGPU_SOKOL_STAGED=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "sokol or interleaved or layout" 2>&1 | tail -2) | tee gpurun_out/r3c_pytest_staged1.txt
(MMDGPU_SOKOL_STAGED=0 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "sokol or interleaved or layout" 2>&1 | tail -2) | tee gpurun_out/r3c_pytest_staged0.txt
(MMDGPU_PDL=0 MMDGPU_TILE_PAIRING=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2) | tee gpurun_out/r3c_pytest_nopdl_nopair.txt
(timeout 400 python tools/gpu_fuzz.py 5000 150 rig 2>&1 | tail -2) | tee gpurun_out/r3c_fuzz_rig.txt
