#!/bin/bash
# Build the experiment variants measured by scripts/variant_ab.sh (objects are not shipped to the box).
set -e
cd "$(dirname "$0")/.."
python scripts/build_variant.py w1 -DMADRL_ONE_WARP_BLOCKS=1
python scripts/build_variant.py smem2 -DMADRL_WW_SMEM_MIN_OPL=2
python scripts/build_variant.py split -DMADRL_HW_SPLIT=1
python scripts/build_variant.py pecache -DMADRL_PE_PHILOX_CACHE=1
python scripts/build_variant.py w1lean -DMADRL_ONE_WARP_BLOCKS=1 -DMADRL_WW_LEAN_SENSE=1 -DMADRL_WW_SKIP_EMPTY_CATCH=1
python scripts/build_variant.py all -DMADRL_ONE_WARP_BLOCKS=1 -DMADRL_WW_SMEM_MIN_OPL=2 -DMADRL_WW_LEAN_SENSE=1 -DMADRL_HW_SPLIT=1 -DMADRL_HW_LEAN_SENSE=1 -DMADRL_WW_SKIP_EMPTY_CATCH=1 -DMADRL_PE_PHILOX_CACHE=1
rm -rf madrl_b200/variants/obj_*
ls -la madrl_b200/variants
