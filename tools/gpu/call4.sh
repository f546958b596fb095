#!/usr/bin/env bash
# 2-GPU call: NCCL/peer-memory parity tests at world 2, bench at N=2, phase timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
O=gpurun_out
W=${W:-2}
LATTE_TEST_WORLDS=$W timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -s > $O/r2c4_dist$W.log 2>&1
echo "dist rc=$?" | tee -a $O/r2c4_dist$W.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $W --steps 20 --warmup 5 > $O/r2c4_bench$W.json 2> $O/r2c4_bench$W.err
echo "bench rc=$?" | tee -a $O/r2c4_bench$W.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29512 tools/dist_phases.py > $O/r2c4_phases$W.log 2>&1
echo "phases rc=$?" | tee -a $O/r2c4_phases$W.log
grep -v "^$" $O/r2c4_dist$W.log | tail -12; tail -c 1200 $O/r2c4_bench$W.json; tail -5 $O/r2c4_bench$W.err; tail -3 $O/r2c4_phases$W.log
