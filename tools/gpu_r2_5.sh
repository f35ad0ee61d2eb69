# 1-GPU batch: the whole -m gpu suite (incl. the 4-bit path and the reference harness binaries), probes
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
timeout 400 python tools/probe_r2.py 4096 2>&1 | tail -22
