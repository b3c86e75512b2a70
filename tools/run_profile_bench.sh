# Round profile: bench line, launch list of the same command, one full capture of the dominant kernels.
TAG=${1:-r1}
set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || exit 1
tail -1 gpurun_out/bench_$TAG.json | cut -c1-300
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
# full capture: skip the 3 warm-up frames' k_wf_ launches (3 x 9: generate, 4 x traverse, 3 x shade, resolve), then one frame
ncu --set full --clock-control none --import-source on -k regex:k_wf_ -s 27 -c 9 -f -o gpurun_out/prof_bench_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
