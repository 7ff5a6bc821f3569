mkdir -p gpurun_out/r2ad
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29841 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2ad/bench_n4.json 2> gpurun_out/r2ad/bench_n4.err; echo "rc=$?" >> gpurun_out/r2ad/bench_n4.err
