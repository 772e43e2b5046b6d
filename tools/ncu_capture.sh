#!/bin/bash
# usage: tools/ncu_capture.sh <name> <kernel-regex> <launch-count> <prof_search case>   (run on the GPU box under gpurun)
# Captures `ncu --set full` for the named kernels, writes the summary table and the per-instruction stall digest as TEXT under
# gpurun_out/ncu/ and deletes the (large) report.
set -e
name=$1; regex=$2; count=$3; kase=$4; skip=${5:-0}
mkdir -p gpurun_out/ncu
rep=/tmp/ncu_$name
ncu --set full --clock-control none --import-source on -k "regex:$regex" -s $skip -c $count -o $rep python tools/prof_search.py $kase 1 > gpurun_out/ncu/$name.log 2>&1 || true
python tools/ncu_summary.py $rep.ncu-rep > gpurun_out/ncu/$name.md 2>> gpurun_out/ncu/$name.log || true
ncu -i $rep.ncu-rep --page source --csv > /tmp/$name.src.csv 2>/dev/null || true
python tools/ncu_src.py /tmp/$name.src.csv 14 >> gpurun_out/ncu/$name.md 2>> gpurun_out/ncu/$name.log || true
ncu -i $rep.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
if len(rows)>2:
    hdr,units=rows[0],rows[1]
    want=('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','smsp__inst_executed_pipe_fp64.sum','sm__warps_active.avg.pct_of_peak_sustained_active')
    for r in rows[2:]:
        print('RAW', r[hdr.index('Kernel Name')][:60], {h:(r[i],units[i]) for i,h in enumerate(hdr) if h in want})
" >> gpurun_out/ncu/$name.md 2>/dev/null || true
rm -f $rep.ncu-rep /tmp/$name.src.csv
