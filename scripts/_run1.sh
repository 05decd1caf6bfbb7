bash scripts/gpu_ab.sh - "-DFWB_NO_STAGE" 2>&1 | tee gpurun_out/ab1.log
( time python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/test3.log
