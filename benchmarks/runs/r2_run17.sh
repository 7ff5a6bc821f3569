mkdir -p gpurun_out/r2v
timeout 400 python -m pytest tests/test_sky_gpu.py -q -x -k "qr or ql" > gpurun_out/r2v/pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2v/pytest.log
timeout 300 python benchmarks/configs.py --configs qr --reps 6 > gpurun_out/r2v/qr.jsonl 2> gpurun_out/r2v/qr.err
