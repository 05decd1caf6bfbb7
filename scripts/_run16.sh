bash scripts/gpu_ab.sh "-DFWB_MBAR_SLEEP=100" "-DFWB_MBAR_SLEEP=400" "-" 2>&1 | tee gpurun_out/ab11.log
