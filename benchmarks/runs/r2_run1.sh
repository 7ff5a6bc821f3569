set -x
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a/smi.txt; free -g >> gpurun_out/r2a/smi.txt; nproc >> gpurun_out/r2a/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a/bench_n1.json 2> gpurun_out/r2a/bench_n1.err; echo "rc=$?" >> gpurun_out/r2a/bench_n1.err
for v in "base" "bbb.stages=3" "bbb.pdl=1" "bbb.stages=3 bbb.pdl=2 bbb.ctas_per_sm=2" "bbb.stages=3 bbb.pdl=1 bbb.ctas_per_sm=2" "bbb.stages=3 bbb.ctas_per_sm=2" "bbb.tile=128" "bbb.tile=128 bbb.pdl=2 bbb.ctas_per_sm=4" "bbb.stages=2 bbb.pdl=2 bbb.ctas_per_sm=2" "bbb.waves=2"; do
  opts=""; if [ "$v" != "base" ]; then for o in $v; do opts="$opts --opt $o"; done; fi
  for st in 20 200; do
    echo "== $v steps=$st" >> gpurun_out/r2a/floor.log
    timeout 300 python bench.py --blocks 512 --steps $st --warmup 10 --no-cpu --no-extra $opts >> gpurun_out/r2a/floor.log 2>> gpurun_out/r2a/floor.err
  done
done
