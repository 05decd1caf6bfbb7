python -m pytest tests -m gpu -q -x -k "tp06 or TP06" 2>&1 | tail -5 | tee gpurun_out/test4.log
bash scripts/gpu_ab.sh - "-DFWB_STAGE_STATE" "-DFWB_TP06_MIN_BLOCKS=3" "-DFWB_TP06_MIN_BLOCKS=3 -DFWB_STAGE_STATE" "-DFWB_TP06_POLICY_FAST" 2>&1 | tee gpurun_out/ab2.log
