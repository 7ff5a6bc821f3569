mkdir -p gpurun_out/r2h
timeout 600 python -m pytest tests/test_sky_gpu.py tests/test_bbb_gpu.py -q -k "trsm or matmul" > gpurun_out/r2h/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2h/pytest.log
timeout 300 python benchmarks/configs.py --configs 5 --reps 9 > gpurun_out/r2h/cfg5.jsonl 2> gpurun_out/r2h/cfgs.err
timeout 300 python benchmarks/configs.py --configs bbbmm --reps 9 > gpurun_out/r2h/bbbmm.jsonl 2>> gpurun_out/r2h/cfgs.err
timeout 300 python benchmarks/configs.py --configs 5 --reps 9 --no-oracle --opt trsm.cond_check=0 > gpurun_out/r2h/cfg5_nocond.jsonl 2>> gpurun_out/r2h/cfgs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h/cfg5_launches.csv python benchmarks/configs.py --configs 5 --reps 1 --no-oracle > gpurun_out/r2h/ncu_cfg5.log 2>&1
