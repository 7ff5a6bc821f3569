mkdir -p gpurun_out/r2aa
timeout 300 python -m pytest tests/test_sky_gpu.py -q -x -k "trsm or sweep" > gpurun_out/r2aa/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2aa/pytest.log
for i in 1 2 3; do
timeout 200 python benchmarks/configs.py --configs 5 --reps 12 --no-oracle >> gpurun_out/r2aa/cfg5_s32.jsonl 2>> gpurun_out/r2aa/err.log
timeout 200 python benchmarks/configs.py --configs 5 --reps 12 --no-oracle --opt trsm.flow_ctr_stride=1 >> gpurun_out/r2aa/cfg5_s1.jsonl 2>> gpurun_out/r2aa/err.log
done
