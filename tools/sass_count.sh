#!/bin/bash
# static SASS instruction count and opcode mix of one kernel:  tools/sass_count.sh <lib.so> <substring of the mangled name> [top]
lib=$1; pat=$2; top=${3:-14}
cuobjdump -sass "$lib" | awk -v pat="$pat" '/Function :/ {on = index($0, pat) > 0} on && /^ +\/\*[0-9a-f]{4}\*\// {op=$2; if (op ~ /^@/) op=$3; sub(/\..*/, "", op); sub(/;/, "", op); n[op]++; t++} END {printf "total %d\n", t; for (k in n) printf "%6d %s\n", n[k], k}' | sort -k1,1 -rn | head -$((top + 1))
