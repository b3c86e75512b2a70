python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu4.log
python tools/tune_libs.py default:6 m5:5 m4:4 2>&1 | tee gpurun_out/tune7.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench5.log
