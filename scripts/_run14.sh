python -m pytest tests -m gpu -q -x -k "sym" 2>&1 | tail -3
bash scripts/gpu_ab.sh "-DFWB_MBAR_HINT=100000" "-DFWB_MBAR_HINT=2000" "-" 2>&1 | tee gpurun_out/ab10.log
