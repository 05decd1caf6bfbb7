# usage: build_variants.sh  -- builds the TP06 A/B libraries listed below HERE (no GPU needed);
# scripts/gpu_variants.sh then times each of them on the GPU box with FWB_LIB=...
set -e
build() { python -m finitewave_b200.build --variant "$1" "$2" > /dev/null && echo "built $1: $2"; }
build s1b3   "-DFWB_NO_STAGE_W -DFWB_TP06_MIN_BLOCKS=3" &
build s1b4   "-DFWB_NO_STAGE_W" &
build s1b4f  "-DFWB_NO_STAGE_W -DFWB_EXP_FASTONLY -DFWB_TP06_DIET" &
build s1b4fs "-DFWB_NO_STAGE_W -DFWB_EXP_FASTONLY -DFWB_TP06_DIET -DFWB_EXP_SMEM" &
wait
build s1b3s  "-DFWB_NO_STAGE_W -DFWB_TP06_MIN_BLOCKS=3 -DFWB_EXP_SMEM" &
build fast   "-DFWB_EXP_FASTONLY" &
build diet   "-DFWB_TP06_DIET" &
wait
