cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rowpanel_gpu.py -x -q -m gpu -k "two_nccl" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err; echo "bench n2 rc=$?"; tail -c 300 gpurun_out/bench_n2b.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n2b.json"))
print("value", round(d["value"]), "ms", d["ms_per_step"], "verified", d["verified"], d["max_rel_err"])
print("c5", {k: d["c5"][k] for k in ("gflops", "ms_per_step", "verified", "max_rel_err", "k_slices")})
PY
