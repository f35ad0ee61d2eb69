cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "f16 or packed or split or default or full_size" 2>&1 | tail -8
timeout 600 python tools/probe_r2.py 4096 2>&1 | tail -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|split_f16|col_absmax" -s 4 -c 4 -o gpurun_out/prof_f16x2 python tools/run_one.py f16x2 4096 2 > gpurun_out/ncu_f16x2.log 2>&1; tail -3 gpurun_out/ncu_f16x2.log
