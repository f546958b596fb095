#!/usr/bin/env bash
# round-2 final evidence call (one GPU): full GPU test suite, smoke, N=1 bench (both arms), ncu launch list of
# the bench command, ncu --set full of the three ClipLoss kernels (tools/prof_clip.py) and of the
# prototype-path kernels (tools/prof_proto.py).  Every ncu run follows the same command exiting 0 without ncu.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
O=gpurun_out
T=r2f
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $O/${T}_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_tests.log 2>&1
echo "tests rc=$?" | tee -a $O/${T}_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${T}_smoke.log
timeout 600 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?" | tee -a $O/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_ref.json 2> $O/${T}_ref.err; echo "ref rc=$?" | tee -a $O/${T}_ref.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $O/${T}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_ncu.log 2>&1
echo "ncu list rc=$?" | tee -a $O/${T}_ncu.log
timeout 200 python tools/prof_clip.py > $O/${T}_prof_clip.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'pair_sweep_kernel|pair_gemm_kernel' -s 3 -c 3 -f -o $O/${T}_prof_pair \
    python tools/prof_clip.py > $O/${T}_ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a $O/${T}_ncu_full.log
timeout 200 python tools/prof_proto.py > $O/${T}_prof_proto.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'cls_stream_kernel|nxc_stream_kernel' -s 6 -c 6 -f -o $O/${T}_prof_proto \
    python tools/prof_proto.py > $O/${T}_ncu_full2.log 2>&1
echo "ncu full2 rc=$?" | tee -a $O/${T}_ncu_full2.log
timeout 120 python tools/nxc_bench.py > $O/${T}_nxc.log 2>&1
timeout 120 python tools/proto_bench.py > $O/${T}_proto.log 2>&1
grep -v "^$" $O/${T}_tests.log | tail -4; tail -c 300 $O/${T}_ref.json
