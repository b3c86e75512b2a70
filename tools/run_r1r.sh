python tools/tune_libs.py default:6 ss4:6 ss8:6 ss12:6 2>&1 | tee gpurun_out/tune21.log
