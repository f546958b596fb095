#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c5_tests.log 2>&1
echo "tests rc=$?" | tee -a $O/r2c5_tests.log
timeout 600 python bench.py --no-cpu-baseline > $O/r2c5_bench.json 2> $O/r2c5_bench.err; echo "bench rc=$?" | tee -a $O/r2c5_bench.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2c5_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file $O/r2c5_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2c5_ncu.log 2>&1
echo "ncu rc=$?" | tee -a $O/r2c5_ncu.log
grep -v "^$" $O/r2c5_tests.log | tail -15; tail -c 300 $O/r2c5_bench.json
