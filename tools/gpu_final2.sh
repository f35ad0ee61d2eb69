cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_n1.json 2>> gpurun_out/bench_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n1.json"))
print("value", round(d["value"]), "ms", d["ms_per_step"], "verified", d["verified"], d["max_rel_err"], "frac", d["roofline"]["frac"], "kernel_ms", d["roofline"]["kernel_ms"], "traffic", d["roofline"]["traffic"])
print("sustained", d["sustained"]["gflops"], "c5", d["c5"]["gflops"], d["c5"]["verified"], "e2e", d["e2e"]["value"])
for r in d["configs34"]: print(r["config"][:32], r.get("tflops", r.get("tops")), r.get("frac_of_bf16_burst", r.get("frac_of_2x_bf16_burst")), r["traffic"])
print(d["sweep"])
PY
