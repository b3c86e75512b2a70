python tools/tune_libs.py default:6 sm3:6 sm5:6 2>&1 | tee gpurun_out/tune26.log
