#!/bin/sh
# Build an A/B variant of librrtfnd.so with extra -D flags into rrt_star_fnd_algorithm_b200/_variants/<name>.so (git-ignored,
# travels with gpurun); select it at run time with RRTFND_LIB=<path>.
#   scripts/build_variant.sh libm_pow -DRRTFND_LIBM_POW=1         (the interpreter's `**` at the reference's sites; -47 %)
#   scripts/build_variant.sh phase -DRRTFND_PHASE_CLOCKS          (scripts/phase_probe.py)
#   scripts/build_variant.sh no_rf -DRRTFND_ROOT_FILTER=0         (circle tests always through the square root and the divisions)
#   scripts/build_variant.sh ann4 -DRRTFND_ANN_FIRST=4 -DRRTFND_ANN_GROW=4   (annulus schedule of the nearest-visible search)
# (the variants measured and removed in round 2 are listed in profiles/r02_variants.md)
set -e
name=$1; shift
cd "$(dirname "$0")/../rrt_star_fnd_algorithm_b200"
mkdir -p _variants
env -u CC -u CXX nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false --prec-div=true --prec-sqrt=true \
    -Xcompiler -fPIC,-ffp-contract=off -shared "$@" -o _variants/$name.so csrc/planner.cu csrc/ops.cu csrc/fnd.cu csrc/host_api.cu
echo "$PWD/_variants/$name.so"
