cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; tail -c 400 gpurun_out/bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n2.json"))
print("value", round(d["value"]), "ms", d["ms_per_step"], "verified", d["verified"], d["max_rel_err"], "sustained", d.get("sustained", {}).get("gflops"))
print("c5", {k: d["c5"][k] for k in ("gflops", "ms_per_step", "verified", "max_rel_err", "k_slices")})
print("e2e", d["e2e"]["value"], d["e2e"].get("verified"))
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/trace_rowpanel.py default 2>&1 | grep "rank"
