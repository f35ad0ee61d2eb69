cd /root/repo
mkdir -p gpurun_out
timeout 120 tools/micro/smem_base_probe.x | tee gpurun_out/smem_base_probe.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_rowpanel_gpu.py -x -q -m gpu -k "f16 or packed or split or default or alpha or rowpanel" 2>&1 | tail -8
timeout 300 python tools/probe_r2.py 4096 2>&1 | tail -24
timeout 400 python tools/probe_epi8.py 2>&1 | tail -64
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1500 gpurun_out/bench_n1.err; head -c 3000 gpurun_out/bench_n1.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|split_f16|col_absmax" -s 4 -c 4 -o gpurun_out/prof_f16x2 python tools/run_one.py f16x2 4096 2 > gpurun_out/ncu_f16x2.log 2>&1; tail -3 gpurun_out/ncu_f16x2.log
