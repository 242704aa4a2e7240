# ncu A/B of k_rgi on the atrium: with the degenerate-ray cut (default) and without (nodegen)
mkdir -p gpurun_out
for v in default nodegen; do
  so=zetaray_b200/libzetaray_b200.so; [ $v != default ] && so=zetaray_b200/libzetaray_b200_$v.so
  ZETARAY_B200_LIB=$PWD/$so timeout 200 ncu --set full --clock-control none -k regex:k_rgi -s 6 -c 1 -o /tmp/prof_rgi_$v python tools/bench_scenes.py atrium 2 > gpurun_out/r1i_ncu_$v.log 2>&1
  ncu -i /tmp/prof_rgi_$v.ncu-rep --page details > gpurun_out/r1i_atrium_k_rgi_${v}_details.txt 2>&1
  ncu -i /tmp/prof_rgi_$v.ncu-rep --page raw --csv 2>/dev/null | python -c "
import sys,csv
rows=list(csv.reader(sys.stdin))
if len(rows)>2:
    h=rows[0]; v=rows[-1]
    keys=('gpu__time_duration.sum','smsp__inst_executed.sum','smsp__thread_inst_executed.sum','sm__warps_active.avg.pct','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sass__inst_executed_local_loads','sass__inst_executed_local_stores','sass__inst_executed_global_loads','smsp__warp_issue_stalled','smsp__average_warp','smsp__pcsamp_warps_issue_stalled')
    print('\n'.join('%s = %s'%(a,b) for a,b in zip(h,v) if any(k in a for k in keys)))
" > gpurun_out/r1i_atrium_k_rgi_${v}_raw.txt 2>&1
  grep -E "Duration|Executed Ipc Active|Avg. Active Threads|L1/TEX Hit|L2 Hit|Issue Slots Busy|No Eligible" gpurun_out/r1i_atrium_k_rgi_${v}_details.txt | sed "s/^/$v /"
done
