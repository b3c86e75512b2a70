python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu5.log
python tools/tune_libs.py default:6 2>&1 | tee gpurun_out/tune8.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench6.log
for w in K2 K4 K5 K3headline; do python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload $w 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/bench6_others.log; done
