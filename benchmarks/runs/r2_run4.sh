mkdir -p gpurun_out/r2d
timeout 600 python -m pytest tests/test_sky_gpu.py -x -q -k "trsm" > gpurun_out/r2d/trsm.log 2>&1; echo "rc=$?" >> gpurun_out/r2d/trsm.log
timeout 600 python benchmarks/configs.py --configs 5 --reps 6 > gpurun_out/r2d/cfg5.jsonl 2> gpurun_out/r2d/cfg5.err
timeout 300 python benchmarks/configs.py --configs 5 --reps 6 --no-oracle --opt trsm.flow_ctas_per_sm=2 > gpurun_out/r2d/cfg5_c2.jsonl 2>> gpurun_out/r2d/cfg5.err
timeout 300 python benchmarks/configs.py --configs 5 --reps 6 --no-oracle --opt trsm.flow_ctas_per_sm=3 > gpurun_out/r2d/cfg5_c3.jsonl 2>> gpurun_out/r2d/cfg5.err
timeout 300 python benchmarks/configs.py --configs 5 --reps 6 --no-oracle --opt trsm.cond_check=0 > gpurun_out/r2d/cfg5_nocond.jsonl 2>> gpurun_out/r2d/cfg5.err
timeout 300 python benchmarks/configs.py --configs 5 --reps 6 --no-oracle --opt trsm.kernel=3 > gpurun_out/r2d/cfg5_chain.jsonl 2>> gpurun_out/r2d/cfg5.err
timeout 300 python benchmarks/step_floor.py --blocks 1024 --steps 200 --variants ";bbb.stages=4;bbb.stages=3;bbb.stages=3,bbb.pdl=1;bbb.stages=4,bbb.pdl=1" > gpurun_out/r2d/floor1024.jsonl 2> gpurun_out/r2d/floor.err
timeout 300 python benchmarks/step_floor.py --blocks 4096 --steps 100 --variants ";bbb.pdl=1;bbb.pdl=0" > gpurun_out/r2d/floor4096.jsonl 2>> gpurun_out/r2d/floor.err
timeout 300 python benchmarks/step_floor.py --blocks 2048 --steps 100 --variants ";bbb.stages=3;bbb.pdl=0" > gpurun_out/r2d/floor2048.jsonl 2>> gpurun_out/r2d/floor.err
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2d/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2d/pytest.log
