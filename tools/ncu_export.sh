#!/bin/bash
# ncu_export.sh <name> <kernel regex> <launch count> <lines-skip> -- <command...>
# Captures `--set full` for the first <count> launches of the kernel into /tmp, and leaves only text exports under gpurun_out/
# (details page, per-source-line aggregate of launch <lines-skip>, DRAM bytes + duration per launch): .ncu-rep files are too big to bring back.
name=$1; kern=$2; cnt=$3; skip=$4; shift 5
rep=/tmp/$name
ncu --set full --clock-control none --import-source on -k regex:$kern -c $cnt -f -o $rep "$@" > gpurun_out/${name}_ncu.log 2>&1
ncu -i $rep.ncu-rep --page details > gpurun_out/${name}_details.txt 2>&1
python tools/ncu_lines.py $rep.ncu-rep $kern $skip 60 > gpurun_out/${name}_lines.txt 2>&1
ncu -i $rep.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__thread_inst_executed_per_inst_executed.ratio','sm__inst_executed.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size']
idx=[hdr.index(w) for w in want if w in hdr]
print(','.join(hdr[i] for i in idx))
for r in rows[2:]:
    print(','.join(r[i] for i in idx))
" > gpurun_out/${name}_raw.csv
