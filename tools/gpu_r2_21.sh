cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_rowpanel_gpu.py -x -q -m gpu -k "f16 or packed or split or default or alpha or rowpanel or full_size" 2>&1 | tail -2
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:"col_absmax|split_f16" -s 3 -c 3 --csv python tools/run_one.py f16x2 4096 4 2>/dev/null | grep -E "col_absmax|split_f16" | awk -F'","' '{print $5, $(NF-2), $NF}' | tr -d '"' | cut -c1-120
timeout 200 python tools/probe_r2.py 4096 2>&1 | head -1
