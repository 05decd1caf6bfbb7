# usage: build_variants.sh  -- builds A/B libraries HERE (no GPU needed); scripts/gpu_variants.sh
# then times each of them on the GPU box with FWB_LIB=...  (the variants measured in round 2
# are listed in profiles/r2_history.md; add lines like the commented one)
set -e
build() { python -m finitewave_b200.build --variant "$1" "$2" > /dev/null && echo "built $1: $2"; }
# build mb4 "-DFWB_TP06_MIN_BLOCKS=4"
