# usage: bash tools/run_prof.sh <tag> [scene] [spp] [maxBounces] — full ncu capture of the wavefront kernels of frame 1
TAG=${1:-x}; SCENE=${2:-K3}; SPP=${3:-1}; MB=${4:-2}
CMD="python tools/prof_wf.py 1 $SCENE $SPP $MB"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
# frame 0's kernels: generate + (mb + 1) traverse + mb shade + resolve = 2 mb + 3; capture frame 1
ncu --set full --clock-control none --import-source on -k regex:k_wf_ -s $((2 * MB + 3)) -c $((2 * MB + 3)) -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1
tail -1 gpurun_out/ncu_$TAG.log; cat gpurun_out/plain_$TAG.log
