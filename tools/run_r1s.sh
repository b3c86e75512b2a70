python tools/tune_libs.py default:6 de4:6 de8:6 de12:6 2>&1 | tee gpurun_out/tune22.log
for t in default de8; do L=librt_b200.so; [ $t != default ] && L=librt_b200_$t.so; RT_B200_LIBNAME=$L python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --workload K4 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('K4', d['value'], d['ms_per_step'])"; done | tee gpurun_out/k4_defer.log
