cd /root/repo
for a in "1024,3072 0" "1024,3072 -1" "512,1536,2048 0" "2048,2048 0" "1024,3072 16" "512,1024,2560 0"; do
  echo "== args [$a]"
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/trace_rowpanel.py $a 2>&1 | grep "rank 1. back"
done
