mkdir -p gpurun_out/r2r
for b in 512 1024 2048; do
timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu --no-extra --blocks $b > gpurun_out/r2r/blocks_$b.json 2>> gpurun_out/r2r/err.log
done
