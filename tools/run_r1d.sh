python tools/tune_libs.py default:6:1 sm3:6:1 sm4:6:1 default:6:2 default:6:0 default:8:1 default:12:1 2>&1 | tee gpurun_out/tune10.log
