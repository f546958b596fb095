#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_prototypes.py -x -q > $O/r2c9_proto_tests.log 2>&1
echo "proto tests rc=$?" | tee -a $O/r2c9_proto_tests.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c9_tests.log 2>&1
echo "tests rc=$?" | tee -a $O/r2c9_tests.log
timeout 300 python tools/nxc_bench.py > $O/r2c9_nxc.log 2>&1; echo "nxc rc=$?" | tee -a $O/r2c9_nxc.log
grep -v "^$" $O/r2c9_proto_tests.log | tail -15; tail -3 $O/r2c9_tests.log; tail -12 $O/r2c9_nxc.log
