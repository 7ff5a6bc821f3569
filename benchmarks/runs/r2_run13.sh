mkdir -p gpurun_out/r2p
for c in 8 16 32 8 16 32; do
timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu --no-extra --opt bbb.host_chunks=$c >> gpurun_out/r2p/chunks_$c.json 2>> gpurun_out/r2p/err.log
done
