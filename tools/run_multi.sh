# usage: bash tools/run_multi.sh <N> ["peer gather"] [workload]
N=${1:-2}; MODES=${2:-"peer gather"}; WL=${3:-K3}
for X in $MODES; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --exchange $X --workload $WL > gpurun_out/bench_${WL}_n${N}_$X.json 2> gpurun_out/bench_${WL}_n${N}_$X.err
echo "rc=$? $X $WL"; tail -1 gpurun_out/bench_${WL}_n${N}_$X.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], {k:v['ms_per_step'] for k,v in d['roofline']['kernels'].items()})"
done
