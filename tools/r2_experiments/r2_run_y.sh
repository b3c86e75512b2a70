#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_pytest.log 2>&1; tail -3 gpurun_out/r2y_pytest.log
for W in K4 K5; do
  timeout 300 python bench.py --steps 5 --warmup 3 --workload $W --no-others --no-cpu-baseline > gpurun_out/r2y_$W.json 2> gpurun_out/r2y_$W.err
  tail -1 gpurun_out/r2y_$W.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$W', d['value'], d['ms_per_step'], d['e2e']['value'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})"
done
timeout 600 python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; tail -1 gpurun_out/r2y_bench.json | cut -c1-1500
