for v in 3 2 1; do python tools/tune.py "{\"leaf_size\": $v}" 2>&1 | tail -1 | tee -a gpurun_out/tune15.log; done
