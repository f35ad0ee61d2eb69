cd /root/repo; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"; tail -c 600 gpurun_out/bench_n8.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n8.json"))
print("value", round(d["value"]), "ms", d["ms_per_step"], "verified", d["verified"], d["max_rel_err"], "sustained", d.get("sustained", {}).get("gflops"))
print("c5", {k: d["c5"][k] for k in ("gflops", "ms_per_step", "verified", "max_rel_err", "k_slices", "frac_of_n_x_bf16_burst")})
print("e2e", d["e2e"]["value"], d["e2e"].get("verified"))
PY
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 8 --steps 5 --warmup 2 > gpurun_out/bench_ref_n8.json 2> gpurun_out/bench_ref_n8.err; echo "ref n8 rc=$?"; head -c 700 gpurun_out/bench_ref_n8.json
