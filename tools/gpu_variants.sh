#!/bin/bash
# Times library variants (po_brax_b200/_tune/*.so next to the committed build) over the episode phases.
# usage: bash tools/gpu_variants.sh "<bench_phases args>" variant1 variant2 ...   ("default" = po_brax_b200/libpobrax.so)
ARGS="$1"; shift
for v in "$@"; do
  if [ "$v" = default ]; then unset POBRAX_LIB; else export POBRAX_LIB=$PWD/po_brax_b200/_tune/libpobrax_$v.so; fi
  echo "== variant $v"
  python tools/bench_phases.py $ARGS 2>&1 | tail -6
done
