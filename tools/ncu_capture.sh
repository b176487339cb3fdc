#!/bin/bash
# One `ncu --set full` capture of a command, summarised ON the GPU box (the .ncu-rep files are too large to
# bring back): per-launch table (tools/ncu_table.py) and per-source-line tables (tools/ncu_src.py) as text.
#   tools/ncu_capture.sh <out_prefix> <kernel_regex> <count> <per-line kernels, comma separated> -- <command...>
out=$1; rx=$2; cnt=$3; lines=$4; shift 5
"$@" > /dev/null 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:$rx" -c "$cnt" -o /tmp/ncu_cap "$@" > "${out}_ncu.log" 2>&1 || { echo "ncu failed"; tail -5 "${out}_ncu.log"; exit 1; }
python tools/ncu_table.py /tmp/ncu_cap.ncu-rep > "${out}_table.txt" 2>&1
IFS=',' read -ra ks <<< "$lines"
for k in "${ks[@]}"; do
    python tools/ncu_src.py /tmp/ncu_cap.ncu-rep "$k" 0 45 > "${out}_lines_${k}.txt" 2>&1
done
rm -f /tmp/ncu_cap.ncu-rep
