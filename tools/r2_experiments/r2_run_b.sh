#!/bin/bash
# slice-of-8 tuning runs on one GPU: lanes 1/2/3, with and without the L2 flush
mkdir -p gpurun_out
for L in 1 2 3; do
  RT_B200_OPTIONS=pipeline_lanes=$L timeout 300 python bench.py --steps 10 --warmup 3 --slice 8 --no-others --no-cpu-baseline --no-e2e > gpurun_out/r2b_s8_l$L.json 2> gpurun_out/r2b_s8_l$L.err; echo "slice8 lanes=$L rc=$?"
done
BENCH_NO_FLUSH=1 RT_B200_OPTIONS=pipeline_lanes=1 timeout 300 python bench.py --steps 10 --warmup 3 --slice 8 --no-others --no-cpu-baseline --no-e2e > gpurun_out/r2b_s8_l1_noflush.json 2> gpurun_out/r2b_s8_l1_noflush.err
BENCH_NO_FLUSH=1 RT_B200_OPTIONS=pipeline_lanes=2 timeout 300 python bench.py --steps 10 --warmup 3 --slice 8 --no-others --no-cpu-baseline --no-e2e > gpurun_out/r2b_s8_l2_noflush.json 2> gpurun_out/r2b_s8_l2_noflush.err
BENCH_NO_FLUSH=1 RT_B200_OPTIONS=pipeline_lanes=1 timeout 300 python bench.py --steps 20 --warmup 3 --workload K3headline --no-others --no-cpu-baseline --no-e2e > gpurun_out/r2b_head_l1_noflush.json 2> gpurun_out/r2b_head_l1_noflush.err
RT_B200_OPTIONS=pipeline_lanes=1,sample_batch=8 timeout 300 python bench.py --steps 10 --warmup 3 --slice 8 --no-others --no-cpu-baseline --no-e2e > gpurun_out/r2b_s8_l1_b8.json 2> gpurun_out/r2b_s8_l1_b8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b_*.json')):
    try:
        d=json.loads(open(f).read().strip().split('\n')[-1])
        print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
