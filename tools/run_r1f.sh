python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest_gpu7.log
for v in 1 4 5; do python tools/tune.py "{\"traversal_variant\": $v}" 2>&1 | tail -1 | tee -a gpurun_out/tune12.log; done
