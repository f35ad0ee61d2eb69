cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "f16 or packed or split or default" 2>&1 | tail -8
timeout 600 python tools/probe_r2.py 4096 2>&1 | tail -40
