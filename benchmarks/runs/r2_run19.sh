mkdir -p gpurun_out/r2x
timeout 300 python -m pytest tests/test_sky_gpu.py -q -x -k "trsm" > gpurun_out/r2x/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2x/pytest.log
timeout 200 python benchmarks/configs.py --configs 5 --reps 9 --no-oracle > gpurun_out/r2x/cfg5_bk64.jsonl 2> gpurun_out/r2x/err.log
timeout 200 python benchmarks/configs.py --configs 5 --reps 9 --opt trsm.flow_bk=32 > gpurun_out/r2x/cfg5_bk32.jsonl 2>> gpurun_out/r2x/err.log
timeout 200 python benchmarks/configs.py --configs 5 --reps 9 --no-oracle --opt trsm.flow_bk=32 --opt trsm.flow_ctas_per_sm=4 > gpurun_out/r2x/cfg5_bk32_c4.jsonl 2>> gpurun_out/r2x/err.log
