#!/bin/bash
# First-contact GPU run: every group in its own process so that a faulting kernel does not
# poison the CUDA context of the others.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python tools/tc_debug.py fp32 > gpurun_out/dbg_fp32.log 2>&1; echo "fp32 rc=$?" >> gpurun_out/dbg_fp32.log
timeout 300 python tools/tc_debug.py bf16 > gpurun_out/dbg_bf16.log 2>&1; echo "bf16 rc=$?" >> gpurun_out/dbg_bf16.log
timeout 600 python -m pytest tests/test_gpu_prototypes.py -m gpu -q --timeout 300 > gpurun_out/pt_proto.log 2>&1
timeout 900 python -m pytest tests/test_gpu_clip.py -m gpu -q --timeout 300 > gpurun_out/pt_clip.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -n 30 gpurun_out/dbg_fp32.log gpurun_out/dbg_bf16.log
tail -n 25 gpurun_out/pt_proto.log gpurun_out/pt_clip.log gpurun_out/smoke.log
