#!/bin/bash
# Round-2 final evidence run (1 GPU): tests, smoke, bench both arms, the reference's harness in every mode + its comparators,
# launch list, ncu --set full per kernel family, compute-sanitizer memcheck.  Results under gpurun_out/.
cd /root/repo; mkdir -p gpurun_out
PKG=how-to-optimize-gemm_b200
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_n1.json 2>> gpurun_out/bench_n1.err
for mode in default 0 1 2 5; do
  f=gpurun_out/output_ref_harness_b200_mode$mode.m; echo "version = 'b200gemm_mode_$mode';" > $f
  if [ $mode = default ]; then env -u B200GEMM_F32_MODE timeout 300 oracle/_ref/ref_cuda_test_MMult__b200.x >> $f 2>&1; else B200GEMM_F32_MODE=$mode timeout 300 oracle/_ref/ref_cuda_test_MMult__b200.x >> $f 2>&1; fi; tail -2 $f
done
for x in MMult_cuBLAS_1 MMult_cuBLAS_2 MMult_cuda_9 MMult_cuda_12; do
  f=gpurun_out/output_ref_harness_$x.m; echo "version = '$x';" > $f
  timeout 300 oracle/_ref/ref_cuda_test_MMult__$x.x >> $f 2>&1; tail -2 $f
done
for d in "f32 auto" "f32 strict" "bf16 auto" "s8 auto"; do set -- $d
  f=gpurun_out/output_b200gemm_$1_$2.m
  timeout 300 "$PKG/harness/test_MMult_b200.x" --dtype $1 --mode $2 --version b200gemm_$1_$2 > $f 2>&1; tail -2 $f
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-c5 > gpurun_out/bench_under_ncu.log 2>&1
for spec in "f16x2 4096 gemm_tc|split_f16|col_absmax 4 4" "bf16x3 4096 gemm_tc|split_planes 2 2" "bf16 4096 gemm_tc 2 1" "bf16_obf16 4096 gemm_tc 2 1" "bf16_obf16 8192 gemm_tc 2 1" \
            "s8 4096 gemm_tc 2 1" "s8_requant 4096 gemm_tc 2 1" "tf32 4096 gemm_tc 2 1" "strict 4096 gemm_ffma 2 1" "mxf4 4096 mxf4 3 3" "generic 1024 gemm_generic 1 1"; do set -- $spec
  name=$1; [ "$2" != 4096 ] && name=$1_$2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$3" -s $4 -c $5 -f -o gpurun_out/prof_$name python tools/run_one.py $1 $2 2 > gpurun_out/ncu_$name.log 2>&1; tail -1 gpurun_out/ncu_$name.log
done
names=$(ls gpurun_out/prof_*.ncu-rep | xargs -n1 basename | sed 's/.ncu-rep//' | tr '\n' ' ')
B200_SUMMARY_DIR=gpurun_out/summaries timeout 600 python tools/summarize_ncu.py r02 $names 2>&1 | tail -3
ls gpurun_out/prof_*.ncu-rep | grep -v "prof_f16x2.ncu-rep" | xargs rm -f        # 15 MB each: only the headline capture travels back
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/sanitizer_memcheck.log
