#!/bin/bash
# Round-2 final evidence pass on one B200 (after the int8 Gram engine became the default): tests, bench (both arms), ncu launch
# list of the bench command, per-kernel timings, parity report, compute-sanitizer memcheck on the int8 Gram tests.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02_pytest_gpu_final.log
python bench.py > $O/r02_bench_1gpu_final.json 2> $O/r02_bench_1gpu_final.err; echo "bench rc=$?"
python bench.py --impl reference > $O/r02_bench_ref_final.json 2> $O/r02_bench_ref_final.err; echo "bench ref rc=$?"
python tools/quick_perf.py sphere > $O/r02_quick_perf_sphere_final.txt 2>&1; echo "quick_perf rc=$?"
python tools/i8_gram_check.py --big > $O/r02_i8_gram_check.txt 2>&1; echo "i8 check rc=$?"
./tools/epi_math_probe > $O/r02_epi_math_probe.txt 2>&1; ./tools/issue_probe > $O/r02_issue_probe.txt 2>&1
python tools/parity_report.py > $O/r02_parity_errors_final.json 2> $O/r02_parity_errors_final.err; echo "parity rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r02_launches_bench_final.csv \
  python bench.py --steps 1 --warmup 3 --no-spd > $O/r02_ncu_bench_final.log 2>&1; echo "ncu list rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_gram_i8.py -m gpu -x -q \
  -k "cross or cyclic or selection or agrees or clamped" > $O/r02_sanitizer_memcheck_i8.log 2>&1; echo "memcheck rc=$?" | tee -a $O/r02_sanitizer_memcheck_i8.log
