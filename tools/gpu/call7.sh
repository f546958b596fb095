#!/usr/bin/env bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/emul_phase_times.py 2 16384 > gpurun_out/r2c7_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"comm_|fixup|lse_|finish|finalize|col_merge|prep_" -c 120 --csv --log-file gpurun_out/r2c7_comm_kernels.csv python tools/emul_phase_times.py 2 16384 > gpurun_out/r2c7_ncu.log 2>&1
echo "rc=$?"
tail -3 gpurun_out/r2c7_ncu.log
