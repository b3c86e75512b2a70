# usage: bash tools/run_prof.sh <tag> [spp] [maxBounces] — launch list + full ncu capture of the wavefront kernels on a K3 frame
TAG=${1:-x}; SPP=${2:-1}; MB=${3:-2}
CMD="python tools/prof_wf.py 1 K3 $SPP $MB"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $CMD > /dev/null 2>&1
# skip frame 0's kernels (generate + 3 per segment + resolve), capture frame 1
ncu --set full --clock-control none --import-source on -k regex:k_wf_ -s $((2 + 3 * MB)) -c $((2 + 3 * MB)) -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1
tail -1 gpurun_out/ncu_$TAG.log
