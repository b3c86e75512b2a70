python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu6.log
for b in 1 4 16; do python tools/tune.py "{\"sample_batch\": $b}" 2>&1 | tail -1 | tee -a gpurun_out/tune11.log; done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench7.log
