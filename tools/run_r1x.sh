python tools/tune_libs.py default:6 s6:6 s8:6 s16:6 2>&1 | tee gpurun_out/tune27.log
