#!/bin/bash
# A/B timing of prebuilt DIAGNOSTICS-build library variants (variants/var_*.so) on ONE box, interleaved, 3 repetitions.
# usage: tools/var_probe.sh [probe script, default tools/bwd_probe.py]
probe=${1:-tools/bwd_probe.py}
for rep in 1 2 3; do
for v in variants/var_*.so; do
  echo "== $v $(IBLN_LIB=$PWD/$v python $probe | tr '\n' ' ')"
done
done
