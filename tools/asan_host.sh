#!/bin/bash
# usage: tools/asan_host.sh [thread]      (default: address + undefined; "thread": TSan)
# Host-side C++ of libtostore_cuda.so (C ABI argument checks, WHERE program translation and
# evaluator, primary-key table, NGH loader walk + JSON parser) under AddressSanitizer +
# UBSan, driven by the CPU self-test tier. No GPU needed. The sanitised build goes to a
# scratch directory and replaces the in-tree .so only for the duration of the run.
set -e
cd "$(dirname "$0")/.."
if [ "$1" = "thread" ]; then SAN="-fsanitize=thread"; PRE="$(gcc -print-file-name=libtsan.so)"
else SAN="-fsanitize=address,-fsanitize=undefined"; PRE="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)"; fi
OUT=${TMPDIR:-/tmp}/tsc_asan
mkdir -p $OUT
SRC=tostore_b200/csrc
make -C $SRC -j8 >/dev/null
for f in tsc_api tsc_search tsc_group tsc_where tsc_pk tsc_loader; do
  /usr/local/cuda/bin/nvcc -O1 -g -std=c++17 -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ \
    -cudart static --expt-relaxed-constexpr \
    -Xcompiler -fPIC,-ffp-contract=off,$SAN,-fno-omit-frame-pointer \
    -c $SRC/$f.cu -o $OUT/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -cudart static -shared \
  -o $OUT/libtostore_cuda.so $OUT/tsc_api.o $OUT/tsc_search.o $OUT/tsc_group.o $OUT/tsc_where.o \
  $OUT/tsc_pk.o $OUT/tsc_loader.o $SRC/build/tsc_scan.o $SRC/build/tsc_scan_l2.o \
  $SRC/build/tsc_scan_ip.o $SRC/build/tsc_scan_cos.o $SRC/build/tsc_tail.o $SRC/build/tsc_gemm.o \
  -ldl -lpthread \
  -Xcompiler $SAN
cp tostore_b200/libtostore_cuda.so $OUT/libtostore_cuda.so.orig
trap 'cp $OUT/libtostore_cuda.so.orig tostore_b200/libtostore_cuda.so' EXIT
cp $OUT/libtostore_cuda.so tostore_b200/libtostore_cuda.so
LD_PRELOAD="$PRE" TSAN_OPTIONS=report_signal_unsafe=0 \
ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
  python -m pytest tests/test_where.py tests/test_where_text.py tests/test_pk_table.py tests/test_ngh_loader.py tests/test_abi.py tests/test_host_prep.py \
  -x -q -m "not gpu" -p no:cacheprovider
