#!/bin/bash
# Round-2 profile set (run on the GPU box through gpurun; every program first exits 0 without ncu):
#   1. launch list of the bench command                       -> launches_r02.csv
#   2. --set full of the per-evaluation kernels + gradient    -> prof_r02_step.ncu-rep (+ raw / source csv)
#   3. --set full of the predict kernels                      -> prof_r02_predict.ncu-rep (+ raw / source csv)
OUT=gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu > $OUT/ncu_r02_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_r02.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $OUT/ncu_r02_bench_ncu.log 2>&1
python scripts/profile_step.py > $OUT/ncu_r02_step_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"syrk_kernel|lik_kernel|chol_kernel|leverage_kernel|syrk_reduce" -c 24 -o $OUT/prof_r02_step -f python scripts/profile_step.py > $OUT/ncu_r02_step_ncu.log 2>&1
python scripts/profile_predict.py > $OUT/ncu_r02_predict_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"kgemm_kernel|row_select_kernel" -c 4 \
    -o $OUT/prof_r02_predict -f python scripts/profile_predict.py > $OUT/ncu_r02_predict_ncu.log 2>&1
# summaries are made here: the reports and the per-instruction pages are too large to travel back (64 MiB limit)
for v in step predict; do
  ncu -i $OUT/prof_r02_$v.ncu-rep --page raw --csv > $OUT/prof_r02_${v}_raw.csv 2>/dev/null
  python scripts/ncu_raw_summary.py $OUT/prof_r02_${v}_raw.csv > $OUT/r02_ncu_full_metrics_$v.txt 2>&1
done
for k in syrk_kernel lik_kernel chol_kernel leverage_kernel; do
  ncu -i $OUT/prof_r02_step.ncu-rep --page source --csv --print-source sass -k regex:$k > $OUT/_src_$k.csv 2>/dev/null
  echo "== $k" >> $OUT/r02_sass_stalls_step.txt
  python scripts/ncu_sass_summary.py $OUT/_src_$k.csv 2>&1 | head -40 >> $OUT/r02_sass_stalls_step.txt
done
for k in kgemm_kernel row_select_kernel; do
  ncu -i $OUT/prof_r02_predict.ncu-rep --page source --csv --print-source sass -k regex:$k > $OUT/_src_$k.csv 2>/dev/null
  echo "== $k" >> $OUT/r02_sass_stalls_predict.txt
  python scripts/ncu_sass_summary.py $OUT/_src_$k.csv 2>&1 | head -40 >> $OUT/r02_sass_stalls_predict.txt
done
rm -f $OUT/_src_*.csv $OUT/prof_r02_step.ncu-rep $OUT/prof_r02_predict.ncu-rep
ls -la $OUT/
du -sh $OUT
