#!/bin/bash
# round-2 development check on ONE GPU: gpu tests, then sumcheck / zero-check timings of the default build and of the
# A/B builds under tools/_libs/ (scratch output in gpurun_out/)
mkdir -p gpurun_out
if [ "$1" == "test" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_tests.log
fi
if [ "$1" == "sctest" ]; then
  timeout 900 python -m pytest tests/test_gpu_field.py tests/test_gpu_sumcheck.py tests/test_gpu_hyperplonk.py -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_tests.log
fi
for what in sumcheck zerocheck; do
  echo "== default $what"; python tools/profile_one.py $what 24 2>&1 | tail -1
  for l in tools/_libs/*.so; do
    [ -f "$l" ] || continue
    echo "== $l $what"; QZ_LIB_PATH=$l python tools/profile_one.py $what 24 2>&1 | tail -1
  done
done
for n in 12 16 18 20 22; do python tools/profile_one.py sumcheck $n 2>&1 | tail -1; done
python tools/profile_one.py zerocheck 20 2>&1 | tail -1
