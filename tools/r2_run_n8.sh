#!/bin/bash
# usage: bash tools/r2_run_n8.sh <N> "<workload:exchange ...>" [tests]
N=${1:-8}; SPECS=${2:-"K3:peer K3:samples K5:peer K4:peer"}; TESTS=${3:-yes}
mkdir -p gpurun_out
if [ "$TESTS" = yes ]; then
  RT_TEST_WORLD=$N timeout 600 python -m pytest tests/test_parallel_gpu.py -m gpu -q -rA -k "K3small and torch-stream" 2>&1 | grep -E "PASSED|FAILED|passed|failed|Error" > gpurun_out/r2_parallel_gpu_n$N.log; cat gpurun_out/r2_parallel_gpu_n$N.log
fi
for S in $SPECS; do
  WL=${S%%:*}; X=${S#*:}
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --exchange $X --workload $WL > gpurun_out/r2_bench_${WL}_n${N}_$X.json 2> gpurun_out/r2_bench_${WL}_n${N}_$X.err
  echo "rc=$? $WL $X"; tail -1 gpurun_out/r2_bench_${WL}_n${N}_$X.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frame_equal', d['frame_equal'], d.get('frame_max_rel_diff'), {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})"
done
