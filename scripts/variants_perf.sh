#!/bin/bash
# usage: variants_perf.sh variants.txt [span]   -- one line per variant: ENV=VAL ENV2=VAL2 (use , instead of spaces inside CB_NVRTC_DEFS)
SPAN=${2:-6e-8}
while read -r line; do
  [ -z "$line" ] && continue
  case "$line" in \#*) continue;; esac
  echo "== $line"
  env $line CB_TIMING=1 python scripts/first_perf.py 16384 adaptive $SPAN 2>&1 | tail -1 | sed 's/.* dev /dev /'
done < "$1"
