#!/bin/bash
# usage: bash tools/r2_run_multi.sh <N>: multi-GPU correctness (tests/test_parallel_gpu.py with RT_TEST_WORLD=N) + bench lines with frame_equal
N=${1:-2}
mkdir -p gpurun_out
RT_TEST_WORLD=$N timeout 900 python -m pytest tests/test_parallel_gpu.py -m gpu -q -rA 2>&1 | grep -v "^$" > gpurun_out/r2_parallel_gpu_n$N.log; echo "pytest rc=${PIPESTATUS[0]}"; tail -14 gpurun_out/r2_parallel_gpu_n$N.log
for X in peer gather; do
  for WL in K3 K5; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --exchange $X --workload $WL > gpurun_out/r2_bench_${WL}_n${N}_$X.json 2> gpurun_out/r2_bench_${WL}_n${N}_$X.err
    echo "rc=$? $X $WL"; tail -1 gpurun_out/r2_bench_${WL}_n${N}_$X.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frame_equal', d['frame_equal'], d.get('frame_sha256'), {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})"
  done
done
