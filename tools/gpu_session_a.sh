#!/bin/bash
# Round-2 evidence pass on one B200: tests, bench, ncu launch list, ncu --set full of the kernels changed this round,
# per-kernel timings, compute-sanitizer on the small dist/gram tests.  Everything lands in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02_pytest_gpu.log
python bench.py > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err; echo "bench ref rc=$?"
python tools/parity_report.py > $O/r02_parity_errors.json 2> $O/r02_parity_errors.err; echo "parity rc=$?"
python tools/quick_perf.py sphere spd chol solve > $O/r02_quick_perf.txt 2>&1; echo "quick_perf rc=$?"
# ncu launch list of the bench command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r02_launches_bench.csv \
  python bench.py --steps 1 --warmup 3 > $O/r02_ncu_bench.log 2>&1; echo "ncu list rc=$?"
# full captures
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spd_ai_gram_kernel -c 1 -o $O/r02_spd_ai_v4 -f \
  python tools/prof_case.py 2048 0 8192 > $O/r02_ncu_spd.log 2>&1; echo "ncu spd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:points_gram_kernel_1c -c 1 -s 1 -o $O/r02_points_gram -f \
  python tools/prof_case.py 16384 0 0 > $O/r02_ncu_gram.log 2>&1; echo "ncu gram rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"points_gram_f32_tc_kernel|trsv" -c 4 -o $O/r02_f32tc_trsv -f \
  python tools/prof_case2.py > $O/r02_ncu_f32tc.log 2>&1; echo "ncu f32tc rc=$?"
# sanitizer on small shapes
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_dist.py tests/test_gpu_edge.py -m gpu -x -q \
  > $O/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $O/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitizer_case.py \
  > $O/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $O/r02_sanitizer_racecheck.log
