mkdir -p gpurun_out/r2g
timeout 600 python -m pytest tests/test_sky_gpu.py -q > gpurun_out/r2g/pytest_sky.log 2>&1; echo "rc=$?" >> gpurun_out/r2g/pytest_sky.log
timeout 300 python benchmarks/configs.py --configs 5,qr --reps 6 > gpurun_out/r2g/cfgs.jsonl 2> gpurun_out/r2g/cfgs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g/cfg5_launches.csv python benchmarks/configs.py --configs 5 --reps 1 --no-oracle > gpurun_out/r2g/ncu_cfg5.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bbb_matmul_staged_kernel -c 1 -o gpurun_out/r2g/bbbmm_staged python benchmarks/configs.py --configs bbbmm --reps 1 --no-oracle > gpurun_out/r2g/ncu_bbbmm.log 2>&1
