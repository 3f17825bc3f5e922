#!/bin/sh
# static opcode histogram of the shipped library (run anywhere: no GPU needed)
cuobjdump -sass "${1:-hevc-hop_b200/libhopgpu.so}" | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | \
  awk '{op=$2; if (op ~ /^@/) op=$3; sub(/;$/,"",op); split(op,a,"."); print a[1]}' | sort | uniq -c | sort -rn
