CMD="python bench.py --workload nyu_dn --steps 6 --warmup 3"
$CMD > gpurun_out/r2b_dn_plain.log 2>&1 || { echo plain failed; tail -3 gpurun_out/r2b_dn_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:dn_fwd -s 4 -c 1 -o gpurun_out/r2b_dn_fwd -f $CMD > gpurun_out/r2b_dn_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dn_bwd -s 4 -c 1 -o gpurun_out/r2b_dn_bwd -f $CMD >> gpurun_out/r2b_dn_ncu.log 2>&1
for k in fwd bwd; do ncu -i gpurun_out/r2b_dn_$k.ncu-rep --page details > gpurun_out/r2b_dn_${k}_details.txt 2>&1; done
grep -E "Duration|Issue Slots Busy|Executed Instructions  |Registers Per|L1/TEX Hit|DRAM Throughput|Achieved Occupancy|Eligible" gpurun_out/r2b_dn_fwd_details.txt gpurun_out/r2b_dn_bwd_details.txt
