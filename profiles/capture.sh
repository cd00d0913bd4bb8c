#!/bin/bash
# ncu captures of one L=1024 conversion (profiles/run_once.py): launch list + one full-set capture of
# each top kernel (middle launch of the second conversion).  Run on the GPU box: bash profiles/capture.sh TAG
TAG=${1:-r01c}
OUT=gpurun_out
python profiles/run_once.py 1024 2 > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python profiles/run_once.py 1024 1 > $OUT/${TAG}_launches.log 2>&1
for k in minors_kernel gemm_grouped_kernel pivchol_kernel svd_select_kernel schur_kernel panel_cholqr_kernel enumerate_kernel sketch_scan_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 1 -f -o $OUT/${TAG}_$k \
      python profiles/run_once.py 1024 2 > $OUT/${TAG}_$k.log 2>&1
  tail -1 $OUT/${TAG}_$k.log
done
