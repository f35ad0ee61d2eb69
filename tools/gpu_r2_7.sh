cd /root/repo
for a in "2048,2048 0" "2048,2048 32" "512,1536,2048 16"; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/trace_rowpanel.py $a 2>&1 | grep "rank"
done
