python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest_gpu8.log
python tools/bvh_stats.py K3 2>&1 | tee gpurun_out/bvh_stats.log
python tools/bvh_stats.py K4 2>&1 | tee -a gpurun_out/bvh_stats.log
for r in 0 8 16 32; do python tools/tune.py "{\"ploc_radius\": $r}" 2>&1 | tail -1 | tee -a gpurun_out/tune13.log; done
