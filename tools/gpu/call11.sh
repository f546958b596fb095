#!/usr/bin/env bash
# prototype kernels after the 16-warp converter and the streaming class sums
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_prototypes.py -x -q > $O/r2c11_proto_tests.log 2>&1
echo "proto tests rc=$?" | tee -a $O/r2c11_proto_tests.log
timeout 300 python tools/nxc_bench.py > $O/r2c11_nxc.log 2>&1; echo "nxc rc=$?" | tee -a $O/r2c11_nxc.log
timeout 300 python tools/proto_bench.py > $O/r2c11_proto.log 2>&1; echo "proto rc=$?" | tee -a $O/r2c11_proto.log
grep -v "^$" $O/r2c11_proto_tests.log | tail -15; tail -8 $O/r2c11_nxc.log; tail -12 $O/r2c11_proto.log
