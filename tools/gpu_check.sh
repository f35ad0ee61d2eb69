#!/bin/bash
# Everything one gpurun call should produce; results under gpurun_out/.
# usage: tools/gpu_check.sh [tests] [probe] [bench] [harness] [ncu]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
what="${*:-tests probe bench harness ncu}"
for w in $what; do case $w in
tests)   timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log ;;
probe)   rm -f gpurun_out/probe.jsonl; timeout 900 python tests/probe_gpu.py tf32 trunc 2>&1 | grep -v tc_desc_dump | tail -30 ;;
bench)   timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
         timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json ;;
harness) for x in b200 MMult_cuda_12 MMult_cuBLAS_1; do
           echo "version = '$x';" > gpurun_out/output_ref_harness_$x.m
           B200GEMM_F32_MODE=${B200GEMM_F32_MODE:-1} timeout 600 oracle/_ref/ref_cuda_test_MMult__$x.x >> gpurun_out/output_ref_harness_$x.m 2>&1; tail -4 gpurun_out/output_ref_harness_$x.m
         done
         for d in "f32 strict" "f32 tf32" "bf16 auto" "s8 auto"; do set -- $d
           timeout 600 "how-to-optimize-gemm_b200/harness/test_MMult_b200.x" --dtype $1 --mode $2 --version b200gemm_$1_$2 > gpurun_out/output_b200gemm_$1_$2.m 2>&1; tail -3 gpurun_out/output_b200gemm_$1_$2.m
         done ;;
ncu)     timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/bench_under_ncu.log 2>&1
         for k in tf32 bf16 strict s8; do
           timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_ -s 2 -c 1 -f -o gpurun_out/prof_$k python tools/run_one.py $k 4096 3 > gpurun_out/ncu_$k.log 2>&1; tail -1 gpurun_out/ncu_$k.log
         done ;;
esac; done
