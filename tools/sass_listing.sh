#!/bin/bash
# SASS of the two benchmarked kernels (cuobjdump -sass of rapt_b200/librapt_b200.so), instruction text only, gzip'ed into
# profiles/, plus an opcode histogram of each.   tools/sass_listing.sh [tag]
tag=${1:-r2}
cd "$(dirname "$0")/.."
for k in _ZN9rapt_fast14k_particle_rknINS_5FieldILi0EEEEEvN4rapt7AdvArgsE:particle_rkn_earthdipole _ZN9rapt_fast11k_gc_dopri5INS_5FieldILi1EEELi4EEEvN4rapt7AdvArgsE:gc_dopri5_doubledipole; do
  fn=${k%%:*}; name=${k##*:}
  cuobjdump -sass -fun "$fn" rapt_b200/librapt_b200.so 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]+)\*\/\s+/\1  /; s/\s*\/\*.*$//' > /tmp/sass_$name.txt
  gzip -9c /tmp/sass_$name.txt > profiles/${tag}_sass_$name.txt.gz
  { echo "# $fn: $(wc -l < /tmp/sass_$name.txt) SASS instructions; opcode histogram (static)"; awk '{op=$2; if (op ~ /^@/) op=$3; sub(/\..*/,"",op); c[op]++} END {for (o in c) printf "%6d %s\n", c[o], o}' /tmp/sass_$name.txt | sort -rn | head -24; } > profiles/${tag}_sass_${name}_opcodes.txt
  head -3 profiles/${tag}_sass_${name}_opcodes.txt
done
