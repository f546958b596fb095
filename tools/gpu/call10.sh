#!/usr/bin/env bash
# round-2 evidence call: full GPU test suite, N=1 bench (both arms), launch list, ncu --set full of the hot kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/r2c10_smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2c10_tests.log 2>&1
echo "tests rc=$?" | tee -a $O/r2c10_tests.log
timeout 600 python bench.py > $O/r2c10_bench.json 2> $O/r2c10_bench.err; echo "bench rc=$?" | tee -a $O/r2c10_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2c10_ref.json 2> $O/r2c10_ref.err; echo "ref rc=$?" | tee -a $O/r2c10_ref.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2c10_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/r2c10_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2c10_ncu.log 2>&1
echo "ncu list rc=$?" | tee -a $O/r2c10_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'pair_sweep_kernel|pair_gemm_kernel' -s 12 -c 6 -f -o $O/r2c10_prof_pair \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2c10_ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a $O/r2c10_ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'nxc_stream|seg_|mix_ema|bank_' -c 24 -f -o $O/r2c10_prof_proto \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2c10_ncu_full2.log 2>&1
echo "ncu full2 rc=$?" | tee -a $O/r2c10_ncu_full2.log
grep -v "^$" $O/r2c10_tests.log | tail -6; tail -c 600 $O/r2c10_bench.json; tail -c 400 $O/r2c10_ref.json
