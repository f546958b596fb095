#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tools/cls_case.py > $O/r2c15_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cls_|seg_' -s 4 -c 6 -f -o $O/r2c15_cls python tools/cls_case.py > $O/r2c15_ncu.log 2>&1
echo "ncu rc=$?" | tee -a $O/r2c15_ncu.log
timeout 300 python -m pytest tests/test_gpu_prototypes.py -q -k "graphed or streaming or mix_and_ema or update_bank" > $O/r2c15_tests.log 2>&1
echo "tests rc=$?" | tee -a $O/r2c15_tests.log
tail -3 $O/r2c15_tests.log
