#!/bin/bash
# usage: profiles/summarize_ncu.sh <report.ncu-rep> <mangled-kernel-substring>   (run in the build container)
# Prints the headline raw metrics, the SASS opcode mix and the per-source-line attribution of one kernel.
set -e
REP=$1; KN=$2
TMP=$(mktemp -d)
ncu -i $REP --page raw --csv 2>/dev/null > $TMP/raw.csv
python - $TMP/raw.csv <<'PY'
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem',
 'sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.max','smsp__thread_inst_executed_per_inst_executed.ratio',
 'sass__inst_executed_local_loads','sass__inst_executed_local_stores','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__average_warps_issue_stalled']
for h,u,v in zip(hdr,units,vals):
    if any(h==w or (w.endswith('stalled') and h.startswith(w)) for w in want): print('%-90s %-14s %s'%(h,u,v))
PY
ncu -i $REP --page source --csv 2>/dev/null > $TMP/src.csv
python - $TMP/src.csv <<'PY'
import csv,sys,collections,re
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
ia=hdr.index('Source'); ie=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples')
tot=sum(int(r[ie]) for r in data)
ops=collections.Counter(); samp=collections.Counter()
for r in data:
    m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ia]); op=m.group(2).split('.')[0] if m else '?'
    ops[op]+=int(r[ie]); samp[op]+=int(r[isamp])
ts=sum(samp.values())
print('--- SASS opcode mix (warp instructions executed: %d)'%tot)
for op,c in ops.most_common(16): print('%-10s %6.2f%% inst   %6.2f%% stall samples'%(op,100*c/tot,100*samp[op]/ts))
PY
D=$(mktemp -d); HERE=$(cd $(dirname $0) && pwd); (cd $D && cuobjdump -xelf all $HERE/../stoch_gpmp_b200/_C/libsgpmp.so >/dev/null 2>&1)
for f in $D/*.cubin; do if nvdisasm -g -c $f 2>/dev/null | grep -q "$KN"; then nvdisasm -g -c $f > $TMP/k.sass; fi; done
echo "--- per source line"
python $HERE/attribute_lines.py $TMP/src.csv $TMP/k.sass $KN 28
