for mode in buckets single none; do
  echo "== $mode"
  LBX_DP_MODE=$mode timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $1 --steps 100 --warmup 5 2>&1 | grep -v "^W1\|^\*\*\*\|OMP_NUM" | tail -1 | cut -c1-230
done
