cd /root/repo
for p in "" "NCCL_PROTO=LL128" "NCCL_PROTO=LL" "NCCL_ALGO=Tree" "NCCL_NVLS_ENABLE=0" "NCCL_P2P_USE_CUDA_MEMCPY=1"; do
  for a in "512,1536,2048 16" "1024,3072 16"; do
    echo "== env [$p] args [$a]"
    env $p timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/trace_rowpanel.py $a 2>&1 | grep "rank 1. back"
  done
done
