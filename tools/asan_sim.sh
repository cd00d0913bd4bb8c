#!/bin/bash
# Memory-safety check of the kernel sources on the CPU: the CTA simulator build (-DTMF_HOSTSIM) compiled with
# AddressSanitizer + UBSan, driven by the simulator test-suite.  (compute-sanitizer is closed on the GPU pool.)
#   tools/asan_sim.sh > profiles/r02_asan_sim.txt 2>&1
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
B=/tmp/simasan
mkdir -p $B
cd $ROOT/temfpy_b200/csrc
FLAGS="-O1 -g -std=c++17 -fPIC -ffp-contract=off -pthread -fsanitize=address,undefined -fno-omit-frame-pointer -DTMF_HOSTSIM"
for f in gemm modes canon pfsite siteprep siteprep_c pair_c minors minors_c chain misc gutzwiller pfaffian enumerate plan; do g++ $FLAGS -x c++ -c $f.cu -o $B/$f.o 2>/dev/null & done
g++ $FLAGS -c hostlogic.cpp -o $B/hostlogic.o
wait
g++ -shared -pthread -fsanitize=address,undefined -o $B/libtemfpy_b200_hostsim.so $B/*.o
cd $ROOT
echo "# ASan + UBSan build of the kernel simulator: $(g++ --version | head -1)"
TMF_SIM_PATH=$B/libtemfpy_b200_hostsim.so LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:abort_on_error=0 \
  python -m pytest tests/test_sim_pipeline.py tests/test_hostlogic.py tests/test_imps.py tests/test_pfaffian.py tests/test_gutzwiller.py -q -m "not gpu" -p no:cacheprovider 2>&1 | tail -25
