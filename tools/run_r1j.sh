python tools/tune_libs.py default:6 inl:6 hno:6 tno:6 hnotno:6 2>&1 | tee gpurun_out/tune14.log
