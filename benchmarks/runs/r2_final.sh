mkdir -p gpurun_out/r2z
nvidia-smi -L > gpurun_out/r2z/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r2z/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2z/smoke.log
timeout 900 python -m pytest tests -q -m gpu --durations=15 > gpurun_out/r2z/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2z/pytest.log
timeout 600 python bench.py > gpurun_out/r2z/bench.json 2> gpurun_out/r2z/bench.err; echo "rc=$?" >> gpurun_out/r2z/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z/bench_reference.json 2> gpurun_out/r2z/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > gpurun_out/r2z/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bbb_walk --launch-skip 6 -c 1 -o gpurun_out/r2z/prof_walk python bench.py --steps 5 --warmup 3 --no-cpu --no-extra > gpurun_out/r2z/ncu2.log 2>&1
timeout 300 python benchmarks/configs.py --configs 4f32,qr --reps 6 > gpurun_out/r2z/configs_f32_qr.jsonl 2> gpurun_out/r2z/configs_f32_qr.err
timeout 200 python benchmarks/configs.py --configs bbbmm --reps 9 > gpurun_out/r2z/bbbmm.jsonl 2> gpurun_out/r2z/bbbmm.err
timeout 200 python benchmarks/configs.py --configs bbbmm --reps 9 --no-oracle --opt bbb.mm_ctas_per_sm=8 > gpurun_out/r2z/bbbmm_8.jsonl 2>> gpurun_out/r2z/bbbmm.err
