python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu9.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200 | tee gpurun_out/bench8.log
