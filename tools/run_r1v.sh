python tools/tune_libs.py default:6 pc:6 2>&1 | tee gpurun_out/tune25.log
