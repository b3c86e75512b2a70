N=${1:-2}
for X in peer gather; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --exchange $X > gpurun_out/bench_n${N}_$X.json 2> gpurun_out/bench_n${N}_$X.err
echo "rc=$? $X"; tail -1 gpurun_out/bench_n${N}_$X.json | cut -c1-330; tail -3 gpurun_out/bench_n${N}_$X.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 0 2>&1 | tail -1 | cut -c1-400
