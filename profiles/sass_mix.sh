#!/bin/bash
# static opcode histogram of one kernel in a .so:  profiles/sass_mix.sh lib.so 'uic_iter_kernelILi8ELb1'
cuobjdump -sass "$1" | awk -v pat="$2" '
/Function :/ {on = ($0 ~ pat)}
on && /\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f]+\*\/ / { i=2; op=$i; if (op ~ /^@/) {i=3; op=$i}; sub(/\..*/,"",op); sub(/;/,"",op); c[op]++; n++ }
END { for (k in c) printf "%6d %s\n", c[k], k | "sort -rn"; close("sort -rn"); print n " total" }'
