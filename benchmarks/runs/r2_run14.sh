mkdir -p gpurun_out/r2q
nvidia-smi topo -m > gpurun_out/r2q/topo.txt 2>&1
python -c "import os; print(os.cpu_count(), len(os.sched_getaffinity(0)))" > gpurun_out/r2q/cpus.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29831 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2q/bench_n2.json 2> gpurun_out/r2q/bench_n2.err; echo "rc=$?" >> gpurun_out/r2q/bench_n2.err
