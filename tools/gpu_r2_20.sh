cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rowpanel_gpu.py -x -q -m gpu -k "cxx or two_nccl" 2>&1 | tail -3
timeout 200 how-to-optimize-gemm_b200/harness/rowpanel_demo.x 2 4096 4096 4096 20 | tee gpurun_out/output_b200gemm_rowpanel_cxx_n2.m
timeout 200 how-to-optimize-gemm_b200/harness/rowpanel_demo.x 1 4096 4096 4096 20 | tee gpurun_out/output_b200gemm_rowpanel_cxx_n1.m
