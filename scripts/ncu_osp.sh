#!/bin/bash
# Launch list of the O-spline moment path inside the profiler range of scripts/profile_step.py (two Laplace
# evaluations + one gradient), cold (ncu's default cache flush) and warm (--cache-control none).
OUT=gpurun_out
python scripts/profile_step.py > $OUT/osp_step_plain.log 2>&1 || { tail -5 $OUT/osp_step_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/osp_launches_cold.csv \
    python scripts/profile_step.py > $OUT/osp_step_ncu1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file $OUT/osp_launches_warm.csv \
    python scripts/profile_step.py > $OUT/osp_step_ncu2.log 2>&1
tail -2 $OUT/osp_step_plain.log
