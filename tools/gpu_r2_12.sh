cd /root/repo; mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc" -s 2 -c 1 -o gpurun_out/prof_f16x2_k512 python tools/run_kslice.py 512 > gpurun_out/ncu_k512.log 2>&1; tail -2 gpurun_out/ncu_k512.log
