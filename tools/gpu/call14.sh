#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_distill.py tests/test_gpu_train_step.py tests/test_gpu_prototypes.py -q > $O/r2c14_tests.log 2>&1
echo "tests rc=$?" | tee -a $O/r2c14_tests.log
timeout 300 python tools/proto_bench.py > $O/r2c14_proto.log 2>&1; echo "proto rc=$?" | tee -a $O/r2c14_proto.log
grep -v "^$" $O/r2c14_tests.log | tail -8; grep -A6 "C=47" $O/r2c14_proto.log
