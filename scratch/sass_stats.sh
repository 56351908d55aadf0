#!/bin/bash
# usage: sass_stats.sh <object> <mangled kernel name>: instruction / spill counts of one kernel
f=${1:-stoch_gpmp_b200/_C/sgpmp_iterate.o}
k=${2:-_ZN5sgpmp14iterate_kernelIfLi2ELi7ELi256ELi1ELi1EEEvNS_10CostParamsIT_EENS_8IterArgsIS2_EE}
cuobjdump -sass $f -fun "$k" > /tmp/sass.txt
echo "instr $(grep -c '^\s*/\*[0-9a-f]\\+\*/' /tmp/sass.txt) LDL $(grep -c 'LDL' /tmp/sass.txt) STL $(grep -c 'STL' /tmp/sass.txt) FFMA2 $(grep -c 'FFMA2' /tmp/sass.txt) MOV $(grep -c ' MOV ' /tmp/sass.txt) CALL $(grep -c 'CALL' /tmp/sass.txt)"
cuobjdump -res-usage $f 2>/dev/null | grep -A1 "$k" | tail -1
