mkdir -p gpurun_out/r2b
python -m pytest tests/test_structured_add_gpu.py -x -q > gpurun_out/r2b/structadd.log 2>&1
python benchmarks/step_floor.py > gpurun_out/r2b/floor512.jsonl 2> gpurun_out/r2b/floor512.err
python benchmarks/step_floor.py --blocks 4096 --steps 100 --variants ";bbb.stages=3;bbb.stages=3,bbb.ctas_per_sm=2,bbb.pdl=2;bbb.pdl=2" > gpurun_out/r2b/floor4096.jsonl 2>> gpurun_out/r2b/floor512.err
python benchmarks/step_floor.py --blocks 1024 --steps 200 --variants ";bbb.stages=3;bbb.stages=3,bbb.ctas_per_sm=2,bbb.pdl=2;bbb.pdl=2" > gpurun_out/r2b/floor1024.jsonl 2>> gpurun_out/r2b/floor512.err
