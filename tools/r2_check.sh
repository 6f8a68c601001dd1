#!/bin/bash
# round-2 development check on ONE GPU: (optional) gpu tests, then sumcheck latency sweep and the 2^24 sumcheck /
# zero-check timings of the default build and of the A/B builds under tools/_libs/ (scratch output in gpurun_out/)
mkdir -p gpurun_out
if [ "$1" == "test" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_tests.log
fi
if [ "$1" == "sctest" ]; then
  timeout 900 python -m pytest tests/test_gpu_field.py tests/test_gpu_sumcheck.py tests/test_gpu_hyperplonk.py tests/test_gpu_c_abi.py -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_tests.log
fi
for l in default tools/_libs/libquill_base.so; do
  [ "$l" == "default" ] || [ -f "$l" ] || continue
  [ "$l" == "default" ] || export QZ_LIB_PATH=$l
  echo "== $l"
  python tools/sc_latency.py 2>&1 | tee gpurun_out/lat_$(basename $l .so).log
  for what in sumcheck zerocheck; do python tools/profile_one.py $what 24 2>&1 | tail -1; done
done
