python tools/tune_libs.py default:6 nt:6 ntd:6 2>&1 | tee gpurun_out/tune24.log
