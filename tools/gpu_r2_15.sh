cd /root/repo; mkdir -p gpurun_out
timeout 300 python tools/probe_crossover.py 2>&1 | tail -16
