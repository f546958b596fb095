#!/usr/bin/env bash
# GPU call 1 (round 2): parity suite, smoke, bench, prototype kernel rates, launch lists.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c1_tests.log 2>&1
echo "tests rc=$?" | tee -a $O/r2c1_tests.log
if false; then
  LATTE_B200_FP16_COPIES=1 timeout 900 python -m pytest tests/test_gpu_clip.py -x -q > $O/r2c1_tests_copies.log 2>&1
  echo "copies rc=$?" | tee -a $O/r2c1_tests_copies.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c1_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/r2c1_smoke.log
timeout 600 python bench.py > $O/r2c1_bench.json 2> $O/r2c1_bench.err; echo "bench rc=$?" | tee -a $O/r2c1_bench.err
timeout 300 python tools/proto_bench.py > $O/r2c1_proto.log 2>&1; echo "proto rc=$?" | tee -a $O/r2c1_proto.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2c1_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv \
    --log-file $O/r2c1_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2c1_ncu.log 2>&1
echo "ncu rc=$?" | tee -a $O/r2c1_ncu.log
tail -3 $O/r2c1_tests.log; tail -2 $O/r2c1_smoke.log; tail -c 600 $O/r2c1_bench.json; tail -30 $O/r2c1_proto.log
