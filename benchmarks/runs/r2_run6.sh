mkdir -p gpurun_out/r2f
timeout 600 python -m pytest tests/test_sky_gpu.py tests/test_bbb_gpu.py -q -k "trsm or matmul or triangular" > gpurun_out/r2f/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2f/pytest.log
timeout 300 python benchmarks/configs.py --configs 4,bbbmm,bbbtri,5 --reps 6 > gpurun_out/r2f/cfgs.jsonl 2> gpurun_out/r2f/cfgs.err
timeout 300 python benchmarks/configs.py --configs 5 --reps 6 --no-oracle --opt trsm.flow_bk=128 > gpurun_out/r2f/cfg5_bk128.jsonl 2>> gpurun_out/r2f/cfgs.err
timeout 300 python benchmarks/configs.py --configs 5 --reps 6 --no-oracle --opt trsm.flow_sleep_ns=0 > gpurun_out/r2f/cfg5_sleep0.jsonl 2>> gpurun_out/r2f/cfgs.err
timeout 300 python benchmarks/configs.py --configs 5 --reps 6 --no-oracle --opt trsm.flow_sleep_ns=200 > gpurun_out/r2f/cfg5_sleep200.jsonl 2>> gpurun_out/r2f/cfgs.err
timeout 300 python benchmarks/configs.py --configs 5 --reps 6 --no-oracle --opt trsm.cond_check=0 > gpurun_out/r2f/cfg5_nocond.jsonl 2>> gpurun_out/r2f/cfgs.err
timeout 300 python benchmarks/configs.py --configs 5 --reps 6 --no-oracle --opt trsm.kernel=3 > gpurun_out/r2f/cfg5_chain.jsonl 2>> gpurun_out/r2f/cfgs.err
timeout 300 python benchmarks/gauss_seidel.py --n 4096 --sweeps 10 > gpurun_out/r2f/gs4096.json 2> gpurun_out/r2f/gs.err
timeout 300 python benchmarks/gauss_seidel.py --n 1000 > gpurun_out/r2f/gs1000.json 2>> gpurun_out/r2f/gs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f/cfg5_launches.csv python benchmarks/configs.py --configs 5 --reps 1 --no-oracle > gpurun_out/r2f/ncu_cfg5.log 2>&1
