python tools/tune_libs.py default:6 x4:6 x8:6 x12:6 x16:6 2>&1 | tee gpurun_out/tune16.log
