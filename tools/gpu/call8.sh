#!/usr/bin/env bash
# 8-GPU call: W=8 parity tests, bench at N=8 (+ A/B variants), N=4, phase table
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
O=gpurun_out
run_bench() {  # name, nproc, extra env...
  local name=$1 np=$2; shift 2
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 \
      --master-port 295$((RANDOM % 90 + 10)) bench.py --gpus $np --steps 20 --warmup 5 > $O/r2c8_$name.json 2> $O/r2c8_$name.err
  echo "$name rc=$?" | tee -a $O/r2c8_$name.err
}
run_bench n8 8 A=1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/dist_phases.py > $O/r2c8_phases8.log 2>&1
echo "phases rc=$?" | tee -a $O/r2c8_phases8.log
LATTE_TEST_WORLDS=8 timeout 600 python -m pytest tests/test_gpu_dist.py -x -q > $O/r2c8_dist8.log 2>&1
echo "dist rc=$?" | tee -a $O/r2c8_dist8.log
run_bench n8_red 8 LATTE_B200_PEER_RED=1
run_bench n8_2sweeps 8 LATTE_B200_BWD_SWEEPS=2
run_bench n4 4 A=1
run_bench n8b 8 A=1
grep -v "^$" $O/r2c8_dist8.log | tail -5; tail -12 $O/r2c8_phases8.log
for f in n8 n8_red n8_2sweeps n4 n8b; do python - <<EOF
import json
try:
    l=json.loads(open("$O/r2c8_$f.json").read().strip().splitlines()[-1])
    print("$f", "ms", round(l["ms_per_step"],4), "value", round(l["value"]/1e6,2), "e2e", round(l["e2e"]["value"]/1e6,2), "parity", l["parity"]["ok"], "gemm_ms", round(l["roofline"]["launch_ms"],4), "weak", round(l["weak_scaling"]["ms_per_step"],4), round(l["weak_scaling"]["efficiency_vs_single_gpu"],3), "clk", l["clocks"]["sm_mhz"])
except Exception as e:
    print("$f", "ERR", e)
EOF
done
