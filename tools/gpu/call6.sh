#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/emul_phase_times.py 2 16384 2>&1 | tail -1 > gpurun_out/r2c6_emul.log
python tools/emul_phase_times.py 8 4096 2>&1 | tail -1 >> gpurun_out/r2c6_emul.log
timeout 600 python -m pytest tests/test_gpu_clip.py -x -q -k "emulated or normalize or prep or autocast" 2>&1 | tail -3 >> gpurun_out/r2c6_emul.log
cat gpurun_out/r2c6_emul.log
