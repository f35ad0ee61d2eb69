cd /root/repo; mkdir -p gpurun_out; rm -f gpurun_out/prof_*.ncu-rep
for gr in 512 1024 4096; do
  B200_GROUP_ROWS=$gr timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:gemm_tc -s 2 -c 2 --csv python tools/run_one.py f16x2 4096 4 2>/dev/null | grep -E "gemm_tc" | awk -F'","' -v g=$gr '{print "group_rows", g, $(NF-2), $(NF-1), $NF}' | tr -d '"'
done
