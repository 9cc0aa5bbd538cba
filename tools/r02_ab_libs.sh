#!/bin/bash
# same-box A/B of library builds: bit-equality of outputs (tools/ab_hash.py) and interleaved forward times
# usage: tools/r02_ab_libs.sh <tag> <libA> <libB> [<libC> ...]
tag=$1; shift
mkdir -p gpurun_out
i=0
for l in "$@"; do
  RRIN_LIB=$PWD/$l timeout 120 python tools/ab_hash.py > gpurun_out/${tag}_hash$i.txt 2> gpurun_out/${tag}_hash$i.err || { echo "hash run failed for $l"; tail -5 gpurun_out/${tag}_hash$i.err; }
  if [ $i -gt 0 ]; then diff -q gpurun_out/${tag}_hash0.txt gpurun_out/${tag}_hash$i.txt > /dev/null && echo "$l: outputs bit-identical to $1" || { echo "$l: outputs DIFFER from $1"; diff gpurun_out/${tag}_hash0.txt gpurun_out/${tag}_hash$i.txt | head -6; }; fi
  i=$((i+1))
done
for rep in 1 2 3; do
  for l in "$@"; do
    echo "$l $(RRIN_LIB=$PWD/$l timeout 60 python tools/forward_time.py ${ITERS:-60} 2>&1 | tail -1)"
  done
done
