bash scripts/gpu_ab.sh "-DFWB_PREFETCH_W" "-DFWB_TP06_MIN_BLOCKS=5" "FWB_CARVEOUT=100;" 2>&1 | tee gpurun_out/ab7.log
