#!/bin/bash
# Everything one gpurun call should produce; results under gpurun_out/.
# usage: tools/gpu_check.sh [tests] [probe] [bench] [harness] [ncu] [quick]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
what="${*:-tests bench harness ncu}"
PKG="how-to-optimize-gemm_b200"
for w in $what; do case $w in
tests)   timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log ;;
probe)   rm -f gpurun_out/probe.jsonl; timeout 900 python tests/probe_gpu.py ${PROBE_CASES:-split} 2>&1 | grep -v tc_desc_dump | tail -60 ;;
quick)   # strict-kernel parity + timings of every mode at 4096 (cheap check after a kernel change)
         timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "strict or golden or split" 2>&1 | tail -5
         timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
         python - <<'EOF'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
print("value", round(d["value"]), d["config"]["kernel"], "e2e", round(d["e2e"]["value"]), "roofline", d.get("roofline", {}).get("achieved"))
for k, v in d.get("modes", {}).items():
    print(" ", k, round(v["gflops"]), v["kernel"], "rel", v["max_rel_err_vs_maxabs"], "bitexact", v["bit_exact_vs_REF_MMult_naive"])
print("sweep", d.get("sweep"))
EOF
         ;;
bench)   timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 7000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
         timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json ;;
harness) # the reference's own harness (unmodified sources) against our shim, in two precision modes,
         # and against its own comparators, N = 256..4096 step 256 (1024.. for MMult_cuda_12)
         for mode in 2 0 1; do
           f=gpurun_out/output_ref_harness_b200_mode$mode.m; echo "version = 'b200gemm_mode$mode';" > $f
           B200GEMM_F32_MODE=$mode timeout 600 oracle/_ref/ref_cuda_test_MMult__b200.x >> $f 2>&1; tail -3 $f
         done
         for x in MMult_cuda_12 MMult_cuBLAS_1; do
           f=gpurun_out/output_ref_harness_$x.m; echo "version = '$x';" > $f
           timeout 600 oracle/_ref/ref_cuda_test_MMult__$x.x >> $f 2>&1; tail -3 $f
         done
         for d in "f32 strict" "f32 tf32" "f32 bf16x3" "f32 bf16x2" "bf16 auto" "s8 auto"; do set -- $d
           f=gpurun_out/output_b200gemm_$1_$2.m
           timeout 600 "$PKG/harness/test_MMult_b200.x" --dtype $1 --mode $2 --version b200gemm_$1_$2 > $f 2>&1; tail -2 $f
         done ;;
ncu)     timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/bench_under_ncu.log 2>&1
         for k in ${NCU_KINDS:-bf16x3 bf16 tf32 strict s8}; do
           timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_ -s 2 -c 1 -f -o gpurun_out/prof_$k python tools/run_one.py $k 4096 3 > gpurun_out/ncu_$k.log 2>&1; tail -1 gpurun_out/ncu_$k.log
         done ;;
esac; done
