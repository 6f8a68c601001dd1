#!/bin/bash
# quick GPU check used during development: gpu tests (optional) + bench summary
if [ "$1" == "test" ]; then timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_all.log; fi
python bench.py --steps 3 --warmup 3 --no-cpu-baseline $QB_ARGS > gpurun_out/bench_q.log 2>gpurun_out/bench_q.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_q.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_q.log").read().strip().splitlines()[-1])
print("msm ms %.2f  acc ms %.2f  frac %.3f  e2e ms %.2f  launches %d" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["gpu_launches"]))
s=d["sumcheck"]; print("sc ms %.3f  rounds ms %.3f  hbm frac %.3f  e2e ms %.2f  launches %d" % (s["ms_per_step"], s["roofline"]["kernel_ms"], s["roofline"]["frac"], s["e2e"]["ms_per_step"], s["gpu_launches"]))
print("imad peak", d["roofline"]["peak"], "clocks", d["clocks"])
PY
