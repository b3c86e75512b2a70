# usage: bash tools/run_prof.sh <tag>   — launch list + full ncu capture of the wavefront kernels on the K3 headline frame
TAG=${1:-x}
CMD="python tools/prof_wf.py 1 K3 1 2"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wf_ -s 7 -c 8 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1
tail -1 gpurun_out/ncu_$TAG.log
