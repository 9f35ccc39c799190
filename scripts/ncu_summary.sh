#!/bin/bash
# one-screen summary of an ncu report: scripts/ncu_summary.sh report.ncu-rep
ncu -i $1 --page details 2>&1 | grep -E "Duration|Issue Slots|Executed Instructions|Registers Per|Active Warps Per Sch|Eligible Warps|L1/TEX Hit|L2 Hit|Achieved Active|Shared Memory Config|Dynamic Shared|DRAM Throughput"
ncu -i $1 --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; vals=rows[-1]
d={h:v for h,v in zip(hdr,vals)}
st={k:float(v.replace(',','')) for k,v in d.items() if 'pcsamp_warps_issue_stalled' in k and 'not_issued' not in k and v not in ('','n/a')}
tot=sum(st.values())
for k,v in sorted(st.items(), key=lambda kv:-kv[1])[:9]: print(f'{100*v/tot:5.1f}%  {k.replace(\"smsp__pcsamp_warps_issue_stalled_\",\"\")}')
print('dram MB read/write', d.get('dram__bytes_read.sum'), d.get('dram__bytes_write.sum'))
"
