#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_clip.py -x -q -k "emulated or flushed" > $O/r2c3_emul.log 2>&1
echo "emul rc=$?" | tee -a $O/r2c3_emul.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2c3_tests.log 2>&1
echo "tests rc=$?" | tee -a $O/r2c3_tests.log
tail -25 $O/r2c3_emul.log; tail -5 $O/r2c3_tests.log
