#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-c}
for L in 1 2; do
  RT_B200_OPTIONS=pipeline_lanes=$L timeout 300 python bench.py --steps 5 --warmup 3 --no-others --no-cpu-baseline --no-e2e > gpurun_out/r2${TAG}_K3_l$L.json 2> gpurun_out/r2${TAG}_K3_l$L.err; echo "K3 lanes=$L rc=$?"
  RT_B200_OPTIONS=pipeline_lanes=$L timeout 300 python bench.py --steps 10 --warmup 3 --slice 8 --no-others --no-cpu-baseline --no-e2e > gpurun_out/r2${TAG}_s8_l$L.json 2> gpurun_out/r2${TAG}_s8_l$L.err; echo "slice8 lanes=$L rc=$?"
  RT_B200_OPTIONS=pipeline_lanes=$L timeout 300 python bench.py --steps 20 --warmup 3 --workload K3headline --no-others --no-cpu-baseline --no-e2e > gpurun_out/r2${TAG}_head_l$L.json 2> gpurun_out/r2${TAG}_head_l$L.err; echo "head lanes=$L rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2${TAG}_*.json')):
    try:
        d=json.loads(open(f).read().strip().split('\n')[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
