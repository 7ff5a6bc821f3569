mkdir -p gpurun_out/r2m
timeout 300 python -m pytest tests/test_sky_gpu.py -q -x -k "qr or ql" > gpurun_out/r2m/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2m/pytest.log
timeout 300 python benchmarks/configs.py --configs qr --reps 5 --no-oracle > gpurun_out/r2m/qr.jsonl 2> gpurun_out/r2m/qr.err
timeout 300 python benchmarks/configs.py --configs qr --reps 5 --no-oracle --opt qr.lookahead=0 > gpurun_out/r2m/qr_nolook.jsonl 2>> gpurun_out/r2m/qr.err
timeout 300 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --warp-sampling-interval 0 --clock-control none --import-source on -k regex:sky_qr_panel_reg --launch-skip 40 -c 1 -o gpurun_out/r2m/qr_panel_hires python benchmarks/configs.py --configs qr --reps 2 --no-oracle --opt qr.lookahead=0 > gpurun_out/r2m/ncu_qr.log 2>&1
