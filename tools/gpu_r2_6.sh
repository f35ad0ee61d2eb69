# 2-GPU sweep: SMs reserved for the exchange x NCCL CTA cap x K-slices
cd /root/repo
mkdir -p gpurun_out; rm -f gpurun_out/sweep_n2.txt
run() { # reserve maxctas slices
  local env="B200_RESERVE_SMS=$1"; [ "$2" != "d" ] && env="$env NCCL_MAX_CTAS=$2"
  env $env timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-c5 --slices $3 > gpurun_out/bench_n2_sl.json 2>> gpurun_out/bench_n2.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_n2_sl.json')); print('reserve $1 nccl_max_ctas $2 slices $3 :', round(d['value']), 'GFLOP/s', round(d['ms_per_step'],4), 'ms', d['verified'])" | tee -a gpurun_out/sweep_n2.txt
}
for sl in "2048,2048" "1024,3072" "512,1536,2048"; do
  for rs in 0 8 16 32; do
    for mc in d 8 16; do run $rs $mc $sl; done
  done
done
