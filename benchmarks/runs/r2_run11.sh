mkdir -p gpurun_out/r2n
timeout 300 python -m pytest tests/test_sky_gpu.py -q -x -k "matmul" > gpurun_out/r2n/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2n/pytest.log
timeout 200 python benchmarks/configs.py --configs 4f32 --reps 9 > gpurun_out/r2n/f32.jsonl 2> gpurun_out/r2n/f32.err
timeout 200 python benchmarks/configs.py --configs 4f32 --reps 9 --opt sky.mm_f32_kernel=1 > gpurun_out/r2n/f32_old.jsonl 2>> gpurun_out/r2n/f32.err
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_structured_add_gpu.py tests/test_sky_gpu.py tests/test_bbb_gpu.py -q -x -k "structured or float32_kernels or trsm_many or qr_panel or staged or prepared or cached" > gpurun_out/r2n/memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2n/memcheck.log
