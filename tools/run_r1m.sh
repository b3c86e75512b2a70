python tools/tune_libs.py default:6 m7:7 m8:8 2>&1 | tee gpurun_out/tune17.log
