mkdir -p gpurun_out/r2k
timeout 600 python -m pytest tests/test_sky_gpu.py tests/test_bbb_gpu.py -q -x -k "qr or ql or cached_plans" > gpurun_out/r2k/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2k/pytest.log
timeout 300 python benchmarks/configs.py --configs qr --reps 5 > gpurun_out/r2k/qr.jsonl 2> gpurun_out/r2k/qr.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sky_qr_panel_reg --launch-skip 40 -c 1 -o gpurun_out/r2k/qr_panel_reg python benchmarks/configs.py --configs qr --reps 2 --no-oracle > gpurun_out/r2k/ncu_qr.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --launch-skip 9000 -c 90 --log-file gpurun_out/r2k/qr_launches.csv python benchmarks/configs.py --configs qr --reps 2 --no-oracle > gpurun_out/r2k/ncu_qr2.log 2>&1
