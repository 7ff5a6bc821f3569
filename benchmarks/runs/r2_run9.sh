mkdir -p gpurun_out/r2l
timeout 600 python -m pytest tests/test_sky_gpu.py tests/test_bbb_gpu.py -q -x -k "qr or ql or cached_plans" > gpurun_out/r2l/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2l/pytest.log
timeout 300 python benchmarks/configs.py --configs qr --reps 5 > gpurun_out/r2l/qr.jsonl 2> gpurun_out/r2l/qr.err
timeout 300 python benchmarks/configs.py --configs qr --reps 5 --no-oracle --opt qr.lookahead=0 > gpurun_out/r2l/qr_nolook.jsonl 2>> gpurun_out/r2l/qr.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sky_qr_panel_reg --launch-skip 40 -c 1 -o gpurun_out/r2l/qr_panel_reg python benchmarks/configs.py --configs qr --reps 2 --no-oracle --opt qr.lookahead=0 > gpurun_out/r2l/ncu_qr.log 2>&1
