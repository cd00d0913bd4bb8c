#!/bin/bash
# ncu captures of the Pfaffian path (configs[1] + an iMPS unit cell): launch list + full-set capture of the site-finish
# kernel.  Run on the GPU box: bash profiles/capture_pf.sh TAG
TAG=${1:-r02}
OUT=gpurun_out
python profiles/run_pf.py 1 > $OUT/${TAG}_pf_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_pf_plain.log; exit 1; }
tail -1 $OUT/${TAG}_pf_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_pf_launches.csv \
    python profiles/run_pf.py 1 > $OUT/${TAG}_pf_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pf_site_kernel -s 1 -c 1 -f -o $OUT/${TAG}_pf_site_kernel \
    python profiles/run_pf.py 1 > $OUT/${TAG}_pf_site_kernel.log 2>&1
tail -1 $OUT/${TAG}_pf_site_kernel.log
