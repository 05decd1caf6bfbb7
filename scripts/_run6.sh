bash scripts/gpu_prof_one.sh c5 r1_f_staged
