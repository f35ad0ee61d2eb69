cd /root/repo; mkdir -p gpurun_out
timeout 300 python tools/probe_kslice.py 2>&1 | tail -40
