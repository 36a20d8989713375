#!/bin/bash
# Round-2 profile set of the O-spline moment path (run on the GPU box through gpurun; every program first exits 0
# without ncu):
#   1. launch list of the bench command (moment path only)      -> launches_r02_ospline.csv
#   2. --set full of the moment-path kernels + the Cholesky       -> r02_ncu_full_metrics_ospline.txt, r02_sass_stalls_ospline.txt
OUT=gpurun_out
FLAGS="--steps 2 --warmup 3 --no-cpu --no-fit --no-grad --no-predict --no-dense"
python bench.py $FLAGS > $OUT/ncu_r02b_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_r02_ospline.csv \
    python bench.py $FLAGS > $OUT/ncu_r02b_bench_ncu.log 2>&1
python scripts/profile_step.py > $OUT/ncu_r02b_step_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"osp_|chol_kernel" -c 40 -o $OUT/prof_r02_osp -f python scripts/profile_step.py > $OUT/ncu_r02b_step_ncu.log 2>&1
ncu -i $OUT/prof_r02_osp.ncu-rep --page raw --csv > $OUT/prof_r02_osp_raw.csv 2>/dev/null
python scripts/ncu_raw_summary.py $OUT/prof_r02_osp_raw.csv > $OUT/r02_ncu_full_metrics_ospline.txt 2>&1
for k in osp_pass_kernel osp_apply_kernel osp_reduce_kernel osp_levpass_kernel; do
  ncu -i $OUT/prof_r02_osp.ncu-rep --page source --csv --print-source sass -k regex:$k > $OUT/_src_$k.csv 2>/dev/null
  echo "== $k" >> $OUT/r02_sass_stalls_ospline.txt
  python scripts/ncu_sass_summary.py $OUT/_src_$k.csv 2>&1 | head -40 >> $OUT/r02_sass_stalls_ospline.txt
done
rm -f $OUT/_src_*.csv $OUT/prof_r02_osp.ncu-rep $OUT/prof_r02_osp_raw.csv
ls -la $OUT/ | head -40
