set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu3.log
python tools/tune_libs.py default:6 f0:6 f2:6 f1s2:6 f1s1:6 f2s2:6 2>&1 | tee gpurun_out/tune6.log
python bench.py --steps 5 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench4.log
