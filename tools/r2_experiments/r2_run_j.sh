#!/bin/bash
mkdir -p gpurun_out
run() { # tag lib options workload extra
  RT_B200_LIBNAME=$2 RT_B200_OPTIONS=$3 timeout 300 python bench.py --steps 5 --warmup 3 --workload $4 --no-others --no-cpu-baseline --no-e2e $5 > gpurun_out/r2j_$1.json 2> gpurun_out/r2j_$1.err; echo "$1 rc=$?"
}
for V in "" _c19 _c22 _c39 _s3; do
  for WL in K3 K2 K3headline K4; do run ${WL}_v$V librt_b200$V.so pipeline_lanes=0 $WL; done
done
for TV in 0 2 4; do run K3_tv$TV librt_b200.so traversal_variant=$TV K3; run K2_tv$TV librt_b200.so traversal_variant=$TV K2; done
run K3s8_v librt_b200.so pipeline_lanes=0 K3 "--slice 8"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j_K*.json')):
    try:
        d=json.loads(open(f).read().strip().split('\n')[-1])
        print(f, d['value'], d['ms_per_step'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})
    except Exception as e:
        print(f, 'ERR', e)
PY
