timeout 600 python -m pytest tests -m gpu -q -x -k "tp06 or TP06 or court or luo or lr91 or slab" 2>&1 | tail -4 | tee gpurun_out/test8.log
bash scripts/gpu_ab.sh - "-DFWB_NO_STAGE_W" 2>&1 | tee gpurun_out/ab9.log
