# usage: build_variants.sh  -- builds the A/B libraries listed below HERE (no GPU needed);
# scripts/gpu_variants.sh then times each of them on the GPU box with FWB_LIB=...
set -e
build() { python -m finitewave_b200.build --variant "$1" "$2" > /dev/null && echo "built $1: $2"; }
build mb4 "-DFWB_TP06_MIN_BLOCKS=4" &
build nosmem "-DFWB_NO_EXP_SMEM" &
wait
