cd /root/repo; mkdir -p gpurun_out; rm -f gpurun_out/prof_*.ncu-rep
for sc in 0 1; do
  B200GEMM_STREAM_C=$sc timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:gemm_tc -s 2 -c 2 --csv python tools/run_one.py f16x2 4096 4 2>/dev/null | grep -E "gemm_tc" | awk -F'","' -v g=$sc '{print "stream_c", g, $(NF-2), $(NF-1), $NF}' | tr -d '"'
done
B200GEMM_STREAM_C=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_rowpanel_gpu.py -x -q -m gpu -k "f16 or packed or split or default or alpha or rowpanel or full_size" 2>&1 | tail -2
B200GEMM_STREAM_C=1 timeout 200 python tools/probe_r2.py 4096 2>&1 | head -2
B200GEMM_STREAM_C=0 timeout 200 python tools/probe_r2.py 4096 2>&1 | head -2
for spec in "f16x2 1536 gemm_tc 2 1" "bf16_obf16 2304 gemm_tc 2 1"; do set -- $spec
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:"$3" -s $4 -c $5 -f -o gpurun_out/prof_$1_$2 python tools/run_one.py $1 $2 4 > gpurun_out/ncu_$1_$2.log 2>&1; tail -1 gpurun_out/ncu_$1_$2.log | cut -c1-100
done
B200_SUMMARY_DIR=gpurun_out/summaries timeout 300 python tools/summarize_ncu.py r02 prof_f16x2_1536 prof_bf16_obf16_2304 2>&1 | tail -2
rm -f gpurun_out/prof_*.ncu-rep
