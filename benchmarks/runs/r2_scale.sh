mkdir -p gpurun_out/r2s
nvidia-smi -L > gpurun_out/r2s/gpus.txt
P=29800
for N in 8 4 2; do
  P=$((P+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2s/bench_n$N.json 2> gpurun_out/r2s/bench_n$N.err; echo "rc=$?" >> gpurun_out/r2s/bench_n$N.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu > gpurun_out/r2s/bench_n1.json 2> gpurun_out/r2s/bench_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811 bench.py --workload cfg4 --gpus 8 --steps 5 --warmup 2 > gpurun_out/r2s/cfg4_n8.json 2> gpurun_out/r2s/cfg4_n8.err; echo "rc=$?" >> gpurun_out/r2s/cfg4_n8.err
timeout 300 python bench.py --workload cfg4 --steps 5 --warmup 2 > gpurun_out/r2s/cfg4_n1.json 2> gpurun_out/r2s/cfg4_n1.err
timeout 600 python -m pytest tests/test_dist_gpu.py -x -q > gpurun_out/r2s/dist.log 2>&1; echo "rc=$?" >> gpurun_out/r2s/dist.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29812 bench.py --gpus 8 --steps 200 --warmup 20 > gpurun_out/r2s/bench_n8_long.json 2> gpurun_out/r2s/bench_n8_long.err
