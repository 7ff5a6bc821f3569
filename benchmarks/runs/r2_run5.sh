mkdir -p gpurun_out/r2e
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2e/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2e/pytest.log
timeout 300 python benchmarks/configs.py --configs 4,bbbmm,bbbtri --reps 6 > gpurun_out/r2e/cfgs.jsonl 2> gpurun_out/r2e/cfgs.err
timeout 300 python benchmarks/configs.py --configs 4 --reps 6 --no-oracle --opt sky.mm_ws=0 > gpurun_out/r2e/cfg4_nows.jsonl 2>> gpurun_out/r2e/cfgs.err
timeout 300 python benchmarks/configs.py --configs bbbmm --reps 6 --no-oracle --opt bbb.mm_staged=0 > gpurun_out/r2e/bbbmm_nostage.jsonl 2>> gpurun_out/r2e/cfgs.err
timeout 300 python benchmarks/gauss_seidel.py --n 1000 > gpurun_out/r2e/gs1000.json 2> gpurun_out/r2e/gs.err
timeout 300 python benchmarks/gauss_seidel.py --n 4096 --sweeps 10 > gpurun_out/r2e/gs4096.json 2>> gpurun_out/r2e/gs.err
timeout 300 python benchmarks/gauss_seidel.py --n 200 --check > gpurun_out/r2e/gs200.json 2>> gpurun_out/r2e/gs.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sky_trsm_flow_kernel -c 1 -o gpurun_out/r2e/flow_cfg5 python benchmarks/configs.py --configs 5 --reps 1 --no-oracle > gpurun_out/r2e/ncu_flow.log 2>&1
