mkdir -p gpurun_out/r2ac
timeout 200 ncu --set full --clock-control none --import-source on -k regex:bbb_matmul_staged --launch-skip 2 -c 1 -o gpurun_out/r2ac/bbbmm_staged python benchmarks/configs.py --configs bbbmm --reps 3 --no-oracle > gpurun_out/r2ac/ncu1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:sky_sgemm_simt --launch-skip 1 -c 1 -o gpurun_out/r2ac/sgemm python benchmarks/configs.py --configs 4f32 --reps 3 > gpurun_out/r2ac/ncu2.log 2>&1
