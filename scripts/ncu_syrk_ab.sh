#!/bin/bash
# A/B capture of the Hessian kernel: current library vs a second build (BGP_LIB_PATH); run on the GPU box
set -x
OUT=gpurun_out
python scripts/profile_step.py > $OUT/ab_plain_new.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:syrk_kernel -s 2 -c 2 -o $OUT/prof_syrk_new -f python scripts/profile_step.py > $OUT/ab_ncu_new.log 2>&1
BGP_LIB_PATH=$PWD/bayesgp_b200/libbgp_oldsyrk.so python scripts/profile_step.py > $OUT/ab_plain_old.log 2>&1 && \
BGP_LIB_PATH=$PWD/bayesgp_b200/libbgp_oldsyrk.so ncu --set full --clock-control none --import-source on -k regex:syrk_kernel -s 2 -c 2 -o $OUT/prof_syrk_old -f python scripts/profile_step.py > $OUT/ab_ncu_old.log 2>&1
for v in new old; do
  ncu -i $OUT/prof_syrk_$v.ncu-rep --page raw --csv > $OUT/prof_syrk_${v}_raw.csv 2>/dev/null
  ncu -i $OUT/prof_syrk_$v.ncu-rep --page source --csv --print-source sass > $OUT/prof_syrk_${v}_src.csv 2>/dev/null
done
ls -la $OUT/prof_syrk_*
