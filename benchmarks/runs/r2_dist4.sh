mkdir -p gpurun_out/r2t
BBM_DIST_TRACE=1 BBM_DIST_TEST_DEADLINE_S=100 timeout 160 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29704 tests/dist_gpu_worker.py > gpurun_out/r2t/w4.log 2>&1; echo rc=$? >> gpurun_out/r2t/w4.log
