#!/bin/bash
# One ncu --set full capture of rp_poa_kernel + its summaries, made ON the GPU box (the .ncu-rep files are too large to
# bring back together: gpurun_out/ is capped at 64 MiB).  usage: tools/ncu_capture.sh <tag> [profile_poa.py args...]
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}" || exit 1
tag="$1"; shift
mkdir -p gpurun_out
rep="/tmp/${tag}.ncu-rep"
timeout 800 ncu --set full --clock-control none --import-source on -k regex:rp_poa_kernel -s 1 -c 1 -f -o "/tmp/${tag}" \
    python tools/profile_poa.py --windows 4736 "$@" > "gpurun_out/${tag}_ncu.log" 2>&1
[ -f "$rep" ] || { echo "no report for $tag"; tail -5 "gpurun_out/${tag}_ncu.log"; exit 1; }
python tools/ncu_by_line.py "$rep" racon_b200/lib/libracon_b200.so > "gpurun_out/${tag}_by_function.txt" 2>&1
python tools/ncu_stalls.py "$rep" racon_b200/lib/libracon_b200.so > "gpurun_out/${tag}_stalls.txt" 2>&1
ncu -i "$rep" --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ('gpu__time_duration', 'dram__bytes', 'sm__inst_executed', 'smsp__inst_executed', 'smsp__thread_inst_executed_per_inst',
        'sm__warps_active', 'launch__', 'l1tex__data_bank_conflicts', 'smsp__average_warps_issue_stalled', 'sm__inst_executed_pipe',
        'smsp__issue_active', 'sm__throughput', 'lts__t_bytes', 'l1tex__t_bytes', 'smsp__cycles_active', 'sm__cycles_elapsed',
        'smsp__warp_issue_stalled', 'sm__ctas_launched', 'local_load', 'local_store', 'smsp__inst_executed_op_local')
for h, u, v in zip(hdr, units, vals):
    if any(k in h for k in keep):
        print(h, u, v)
" > "gpurun_out/${tag}_raw_selected.txt"
ls -la "$rep"; wc -l gpurun_out/${tag}_*.txt
