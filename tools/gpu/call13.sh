#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_distill.py tests/test_gpu_train_step.py -q > $O/r2c13_new_tests.log 2>&1
echo "new tests rc=$?" | tee -a $O/r2c13_new_tests.log
timeout 120 python tools/cls_case.py > $O/r2c13_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cls_' -s 4 -c 4 -f -o $O/r2c13_cls python tools/cls_case.py > $O/r2c13_ncu.log 2>&1
echo "ncu rc=$?" | tee -a $O/r2c13_ncu.log
grep -v "^$" $O/r2c13_new_tests.log | tail -12
