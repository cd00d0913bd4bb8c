#!/bin/bash
# ncu captures of the Gutzwiller path (configs[2]): launch list + one full-set capture of the projection kernel and
# of the two canonical-sweep kernels.  Run on the GPU box: bash profiles/capture_gutz.sh TAG
TAG=${1:-r02}
OUT=gpurun_out
python profiles/run_gutz.py 256 1 > $OUT/${TAG}_gutz_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_gutz_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_gutz_launches.csv \
    python profiles/run_gutz.py 256 1 > $OUT/${TAG}_gutz_launches.log 2>&1
for k in gutz_pair_kernel block_svd_kernel block_qr_kernel; do
  SKIP=100; [ $k = gutz_pair_kernel ] && SKIP=0
  ncu --set full --clock-control none --import-source on -k regex:$k -s $SKIP -c 1 -f -o $OUT/${TAG}_$k \
      python profiles/run_gutz.py 256 1 > $OUT/${TAG}_$k.log 2>&1
  tail -1 $OUT/${TAG}_$k.log
done
