#!/bin/bash
# round 2, call S (1 GPU): compute-sanitizer (memcheck, racecheck) over small GPU tests: the tail's shared-memory
# overlays, mbarrier re-use, bulk copies into the row stage, the VMM-backed row block; then the whole GPU suite
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2s}
L=gpurun_out/$T.log
nvidia-smi -L | tee $L
SAN=/usr/local/cuda/bin/compute-sanitizer
SEL="tests/test_gpu_parity.py::test_golden_fixtures tests/test_gpu_parity.py::test_k_larger_than_live_rows_and_empty_index tests/test_gpu_parity.py::test_random_rows_host_append tests/test_gpu_growth.py::test_synthetic_append_grows_too_and_rows_keep_their_address tests/test_gpu_pipeline.py::test_pipelined_search_flags_what_it_cannot_certify"
echo "== memcheck" | tee -a $L
timeout 1200 $SAN --tool memcheck --error-exitcode 9 --launch-timeout 600 python -m pytest $SEL -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/${T}_memcheck.txt 2>&1
echo "exit $?" | tee -a $L
grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" gpurun_out/${T}_memcheck.txt | head -20 | tee -a $L
echo "== memcheck: certificate / tensor path / where (a few parameters)" | tee -a $L
timeout 1200 $SAN --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_certificate.py::test_scan_path_near_ties_need_the_range_pass[10-200-0-1e-06]" "tests/test_gpu_gemm.py::test_gemm_path_large_k[1-2]" tests/test_where.py -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/${T}_memcheck2.txt 2>&1
echo "exit $?" | tee -a $L
grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" gpurun_out/${T}_memcheck2.txt | head -20 | tee -a $L
echo "== racecheck (shared memory hazards) on the fused tail" | tee -a $L
timeout 1200 $SAN --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py::test_golden_fixtures tests/test_gpu_parity.py::test_k_larger_than_live_rows_and_empty_index -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/${T}_racecheck.txt 2>&1
echo "exit $?" | tee -a $L
grep -E "passed|failed|RACECHECK SUMMARY|hazard|Error" gpurun_out/${T}_racecheck.txt | head -20 | tee -a $L
echo "== all gpu tests" | tee -a $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 | tee -a $L
