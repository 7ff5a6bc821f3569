mkdir -p gpurun_out/r2c
nvidia-smi -L > gpurun_out/r2c/gpus.txt
timeout 900 python -m pytest tests/test_dist_gpu.py -x -q -k "2" > gpurun_out/r2c/dist2.log 2>&1; echo "rc=$?" >> gpurun_out/r2c/dist2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c/bench_n2.json 2> gpurun_out/r2c/bench_n2.err; echo "rc=$?" >> gpurun_out/r2c/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --blocks 1024 --steps 20 --warmup 5 > gpurun_out/r2c/bench_n2_b1024.json 2> gpurun_out/r2c/bench_n2_b1024.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-extra --no-cpu > gpurun_out/r2c/bench_n1.json 2> gpurun_out/r2c/bench_n1.err
