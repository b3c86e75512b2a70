python tools/tune_libs.py default:6 pf:6 2>&1 | tee gpurun_out/tune23.log
