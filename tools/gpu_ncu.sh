#!/bin/bash
# ncu --set full per kernel family, summarised on the box (the .ncu-rep files are 15 MB each; only summaries travel back)
cd /root/repo; mkdir -p gpurun_out; rm -f gpurun_out/prof_*.ncu-rep
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "f16 or default" 2>&1 | tail -2
for spec in "f16x2 4096 gemm_tc|split_f16|col_absmax 8 4" "bf16x3 4096 gemm_tc|split_planes 4 2" "bf16 4096 gemm_tc 2 1" "bf16_obf16 4096 gemm_tc 2 1" "bf16_obf16 8192 gemm_tc 2 1" \
            "s8 4096 gemm_tc 2 1" "s8_requant 4096 gemm_tc 2 1" "tf32 4096 gemm_tc 2 1" "strict 4096 gemm_ffma 2 1" "mxf4 4096 mxf4 6 3" "generic 1024 gemm_generic 2 1"; do set -- $spec
  name=$1; [ "$2" != 4096 ] && name=$1_$2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$3" -s $4 -c $5 -f -o gpurun_out/prof_$name python tools/run_one.py $1 $2 4 > gpurun_out/ncu_$name.log 2>&1; tail -1 gpurun_out/ncu_$name.log | cut -c1-120
done
names=$(ls gpurun_out/prof_*.ncu-rep | xargs -n1 basename | sed 's/.ncu-rep//' | tr '\n' ' ')
B200_SUMMARY_DIR=gpurun_out/summaries timeout 900 python tools/summarize_ncu.py r02 $names 2>&1 | tail -12
ls gpurun_out/prof_*.ncu-rep | grep -v "prof_s8.ncu-rep" | xargs rm -f
timeout 200 python tools/probe_r2.py 4096 2>&1 | head -8
