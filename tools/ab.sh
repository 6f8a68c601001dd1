#!/bin/bash
# A/B of the 2^24 sumcheck / zero-check between the in-tree library and every build under tools/_libs/ (one GPU)
for rep in 1 2; do
for l in default tools/_libs/*.so; do
  [ "$l" == "default" ] && unset QZ_LIB_PATH || export QZ_LIB_PATH=$l
  echo "== $l"; for what in ${AB_WHAT:-sumcheck}; do python tools/profile_one.py $what ${AB_N:-24} 2>&1 | tail -1; done
done
done
