mkdir -p gpurun_out/r2o
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29821 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2o/bench_n2.json 2> gpurun_out/r2o/bench_n2.err; echo "rc=$?" >> gpurun_out/r2o/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29822 bench.py --workload cfg4 --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2o/cfg4_n2.json 2> gpurun_out/r2o/cfg4_n2.err; echo "rc=$?" >> gpurun_out/r2o/cfg4_n2.err
timeout 300 python -m pytest tests/test_dist_gpu.py -x -q > gpurun_out/r2o/dist.log 2>&1; echo "rc=$?" >> gpurun_out/r2o/dist.log
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29823 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2o/ref_n2.json 2> gpurun_out/r2o/ref_n2.err; echo "rc=$?" >> gpurun_out/r2o/ref_n2.err
