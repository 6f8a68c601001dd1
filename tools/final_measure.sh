#!/bin/bash
# Round-end measurement on ONE GPU: default bench line, reference arm, ncu launch list of the same command, and one
# `--set full` capture of each dominant kernel.  Everything lands in gpurun_out/ (scratch); tools/summarise_profiles.py
# turns the ncu outputs into the text summaries committed under profiles/.
set -x
python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --mlpcs-log-n 0 --hyperplonk-log-rows 0"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
python tools/profile_one.py msm 24 pre > gpurun_out/pm24.log 2>&1 && ncu --set full --clock-control none --import-source on \
  -k regex:msm_accumulate -c 1 -o gpurun_out/prof_r1_msm -f python tools/profile_one.py msm 24 pre > gpurun_out/ncum24.log 2>&1
python tools/profile_one.py sumcheck 24 > gpurun_out/p24.log 2>&1 && ncu --set full --clock-control none --import-source on \
  -k regex:sc_round_prod -c 2 -o gpurun_out/prof_r1_sc -f python tools/profile_one.py sumcheck 24 > gpurun_out/ncu24.log 2>&1
tail -c 600 gpurun_out/bench_default.log
