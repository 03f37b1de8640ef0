#!/bin/bash
# A/B timing of prebuilt library variants (variants/var_*.so) on ONE box, interleaved, 3 repetitions
cp ibl_nerf_b200/libiblnerf_b200.so /tmp/orig.so
for rep in 1 2 3; do
for v in variants/var_*.so; do
  cp $v ibl_nerf_b200/libiblnerf_b200.so
  echo "== $v $(python tools/stash_probe.py | tr '\n' ' ')"
done
done
cp /tmp/orig.so ibl_nerf_b200/libiblnerf_b200.so
