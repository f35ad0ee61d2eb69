# 2-GPU batch: row-panel plan on two NCCL ranks, bench both arms at N=2
cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/test_rowpanel_gpu.py -x -q -m gpu -k "two_nccl" 2>&1 | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; tail -c 600 gpurun_out/bench_n2.err; head -c 2500 gpurun_out/bench_n2.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?"; cat gpurun_out/bench_ref_n2.json | head -c 900
for sl in "2048,2048" "512,1024,2560" "1024,3072"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-c5 --slices $sl > gpurun_out/bench_n2_sl.json 2>> gpurun_out/bench_n2.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_n2_sl.json')); print('slices $sl', round(d['value']), d['ms_per_step'], d['verified'])"
done
