bash scripts/gpu_ab.sh "-DFWB_L2_AHEAD=592" "-DFWB_L2_AHEAD=300" "-DFWB_L2_AHEAD=1184" "-" 2>&1 | tee gpurun_out/ab6.log
