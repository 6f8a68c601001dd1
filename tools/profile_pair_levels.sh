#!/bin/bash
# ncu --set full captures of the pair-level kernels (levels 1 and 2 of a 2^24 commit with QZ_MSM_PAIR_LEVELS=3), summarised
# on the box into gpurun_out/profiles_pair/ (the reports themselves are too large to travel).
mkdir -p /tmp/qzprof gpurun_out/profiles_pair
export QZ_MSM_PAIR_LEVELS=3
NCU="ncu --set full --clock-control none --import-source on"
for k in apply scan; do
  timeout ${PAIR_NCU_TIMEOUT:-50} $NCU -k regex:msm_pair_$k -c 2 -o /tmp/qzprof/prof_pair_$k -f python tools/profile_one.py msm 24 pre > gpurun_out/ncu_pair_$k.log 2>&1
  echo "ncu $k rc=$?"
done
python - <<'PY'
import os, sys
sys.argv = ["summarise"]
sys.path.insert(0, "tools")
import summarise_profiles as sp
for k in ("apply", "scan"):
    rep = f"/tmp/qzprof/prof_pair_{k}.ncu-rep"
    if os.path.exists(rep):
        out = f"gpurun_out/profiles_pair/r02_ncu_msm_pair_{k}.txt"
        sp.full(rep, out, f"r02: QZ_MSM_PAIR_LEVELS=3 ncu --set full --clock-control none -k regex:msm_pair_{k} -c 2 python "
                          "tools/profile_one.py msm 24 pre  (levels 1 and 2 of a 2^24 commit)")
        sp.hot_spots(rep, out)
        print("wrote", out)
PY
