#!/usr/bin/env bash
# distill + train_step tests, prototype tests and rates after the TMA-ring class sums
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_distill.py tests/test_gpu_train_step.py -q > $O/r2c12_new_tests.log 2>&1
echo "new tests rc=$?" | tee -a $O/r2c12_new_tests.log
timeout 600 python -m pytest tests/test_gpu_prototypes.py -x -q > $O/r2c12_proto_tests.log 2>&1
echo "proto tests rc=$?" | tee -a $O/r2c12_proto_tests.log
timeout 300 python tools/proto_bench.py > $O/r2c12_proto.log 2>&1; echo "proto rc=$?" | tee -a $O/r2c12_proto.log
grep -v "^$" $O/r2c12_new_tests.log | tail -40; grep -v "^$" $O/r2c12_proto_tests.log | tail -8; grep -A6 "C=47" $O/r2c12_proto.log
