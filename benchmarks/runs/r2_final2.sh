mkdir -p gpurun_out/r2ab
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r2ab/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2ab/smoke.log
timeout 900 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2ab/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2ab/pytest.log
timeout 600 python bench.py > gpurun_out/r2ab/bench.json 2> gpurun_out/r2ab/bench.err; echo "rc=$?" >> gpurun_out/r2ab/bench.err
timeout 200 python benchmarks/configs.py --configs 5 --reps 9 > gpurun_out/r2ab/cfg5.jsonl 2> gpurun_out/r2ab/cfg5.err
