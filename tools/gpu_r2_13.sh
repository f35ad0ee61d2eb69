cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_rowpanel_gpu.py -x -q -m gpu -k "f16 or packed or split or default or alpha or rowpanel or full_size" 2>&1 | tail -5
timeout 300 python tools/probe_kslice.py 2>&1 | grep -v "dynamic': 1" | tail -24
