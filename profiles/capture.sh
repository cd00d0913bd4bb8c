#!/bin/bash
# ncu captures of one L=1024 conversion: launch lists (profiles/run_once.py and bench.py itself) + one full-set
# capture of each top kernel (a launch of the second conversion).  Run on the GPU box: bash profiles/capture.sh TAG
TAG=${1:-r02}
OUT=gpurun_out
python profiles/run_once.py 1024 2 > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python profiles/run_once.py 1024 1 > $OUT/${TAG}_launches.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/${TAG}_bench_plain.json 2> $OUT/${TAG}_bench_plain.err || { echo "bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/${TAG}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/${TAG}_bench_ncu.log 2>&1
for k in minors_kernel gemm_grouped_kernel svd_select_kernel panel_cholqr_kernel enumerate_kernel sketch_scan_kernel nested_site_kernel site_plan_kernel small_modes_kernel ritz_kernel edge_vector_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o $OUT/${TAG}_$k \
      python profiles/run_once.py 1024 2 > $OUT/${TAG}_$k.log 2>&1
  tail -1 $OUT/${TAG}_$k.log
done
