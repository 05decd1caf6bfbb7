bash scripts/gpu_ab.sh - "-DFWB_NO_STAGE" "-DFWB_TP06_MIN_BLOCKS=3" 2>&1 | tee gpurun_out/ab4.log
python -m pytest tests -m gpu -q -x -k "tp06 or TP06 or lr91 or luo or court or slab or cabi" 2>&1 | tail -5 | tee gpurun_out/test5.log
