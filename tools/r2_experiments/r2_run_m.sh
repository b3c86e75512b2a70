#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sample_partition or fast_shade" 2>&1 | tail -3
RT_TEST_WORLD=$N timeout 900 python -m pytest tests/test_parallel_gpu.py -m gpu -q -k samples -rA 2>&1 | grep -E "PASSED|FAILED|passed|failed|Error" | tail -8
for X in samples peer; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --exchange $X > gpurun_out/r2m_bench_K3_n${N}_$X.json 2> gpurun_out/r2m_bench_K3_n${N}_$X.err
  echo "rc=$? $X"; tail -1 gpurun_out/r2m_bench_K3_n${N}_$X.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frame_equal', d['frame_equal'], d.get('frame_max_rel_diff'), {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})"
done
