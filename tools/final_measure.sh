#!/bin/bash
# Round-end measurement on ONE GPU: default bench line, reference arm, ncu launch lists of the same commands, and one
# `--set full` capture of each kernel the docs quote.  Everything lands in gpurun_out/ (scratch);
# tools/summarise_profiles.py TAG turns the ncu outputs into the text summaries committed under profiles/.
# Every ncu pass runs only after the same command has exited 0 without ncu.
set -x
mkdir -p gpurun_out /tmp/qzprof
python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --mlpcs-log-n 0 --hyperplonk-log-rows 0"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
  --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
python tools/profile_one.py msm 24 pre > gpurun_out/pm24.log 2>&1 && {
  $NCU -k regex:msm_accumulate -c 1 -o /tmp/qzprof/prof_msm -f python tools/profile_one.py msm 24 pre > gpurun_out/ncum24.log 2>&1
  $NCU -k regex:msm_bucket_reduce -c 1 -o /tmp/qzprof/prof_bucket_reduce -f python tools/profile_one.py msm 24 pre > gpurun_out/ncubr.log 2>&1
  $NCU -k regex:Onesweep -c 3 -o /tmp/qzprof/prof_sort -f python tools/profile_one.py msm 24 pre > gpurun_out/ncusort.log 2>&1
}
python tools/profile_one.py sumcheck 24 > gpurun_out/p24.log 2>&1 && {
  $NCU -k regex:sc_round_prod -c 3 -o /tmp/qzprof/prof_sc -f python tools/profile_one.py sumcheck 24 > gpurun_out/ncu24.log 2>&1
  $NCU -k regex:sc_mid -c 1 -o /tmp/qzprof/prof_sc_mid -f python tools/profile_one.py sumcheck 24 > gpurun_out/ncumid.log 2>&1
}
python tools/profile_one.py zerocheck 24 > gpurun_out/pz24.log 2>&1 && \
  $NCU -k regex:sc_round_zc -c 2 -o /tmp/qzprof/prof_zc -f python tools/profile_one.py zerocheck 24 > gpurun_out/ncuzc.log 2>&1
python tools/profile_one.py mlpcs 22 > gpurun_out/pml22.log 2>&1 && \
  $NCU -k regex:ntt_pass -c 3 -o /tmp/qzprof/prof_ntt -f python tools/profile_one.py mlpcs 22 > gpurun_out/ncuntt.log 2>&1
if [ "$1" == "hp" ]; then
  python tools/profile_hp.py 20 > gpurun_out/hp20.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv \
    --log-file gpurun_out/lhp20.csv python tools/profile_hp.py 20 > gpurun_out/ncuhp20.log 2>&1
fi
# the reports stay on the box (gpurun_out/ is capped at 64 MiB): only their text summaries travel
python tools/summarise_profiles.py ${TAG:-r02} /tmp/qzprof/ gpurun_out/profiles_${TAG:-r02}/ > gpurun_out/summarise.log 2>&1
tail -c 400 gpurun_out/bench_default.log
