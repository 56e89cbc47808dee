#!/bin/bash
# SASS listings of the four benchmarked kernel instantiations (no encodings), from the objects of the last
# `python -m madrl_b200.build`:  bash scripts/sass_listing.sh r2  ->  profiles/r2_sass_{ww_c2,ww_c4,pe_c3,hw_c5}.txt
set -eu
TAG=${1:-r2}; B=madrl_b200/build
dump() {  # out, object, mangled name
  { echo "# cuobjdump -sass -fun $3 $2  (sm_100a, $(nvcc --version | tail -n 2 | head -n 1))"
    cuobjdump -res-usage "$2" | grep -A1 "$3" | tail -n 1
    cuobjdump -sass -fun "$3" "$2" | grep -v '^\s*/\* 0x' | sed 's#/\* 0x[0-9a-f]* \*/##; s/[[:space:]]*$//'; } > "profiles/${TAG}_sass_$1.txt"
  echo "profiles/${TAG}_sass_$1.txt: $(grep -c ';' profiles/${TAG}_sass_$1.txt) instructions"
}
dump ww_c2 $B/waterworld.o _ZN5madrl9ww_kernelIfLi1ELi1ELi30ELb0ELb0EEEvNS_8WWParamsIT_EE
dump ww_c4 $B/waterworld.o _ZN5madrl9ww_kernelIfLi4ELi1ELi30ELb0ELb0EEEvNS_8WWParamsIT_EE
dump pe_c3 $B/pursuit.o _ZN5madrl9pe_kernelILi1ELi2ELi7ELb0ELb1EEEvNS_8PEParamsE
dump hw_c5 $B/hostage.o "$(cuobjdump -res-usage $B/hostage.o | grep -o '_ZN5madrl9hw_kernelIfLi1ELi1ELi30E[A-Za-z0-9_]*' | head -n 1)"
