mkdir -p gpurun_out/r2w
timeout 400 python -m pytest tests/test_sky_gpu.py -q -x -k "rmul or matmul" > gpurun_out/r2w/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2w/pytest.log
