run() { echo "== $1"; env $1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 10 --warmup 3 --only-default --method pointtopoint $2 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['leg_ms'], d['sections'].get('x->y',{}).get('exchange_ms'))"; }
run "A=1" ""
run "NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=16" ""
run "NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32" ""
run "NCCL_P2P_CHUNKSIZE=2097152" ""
run "A=1" "--nccl-ctas 32"
run "A=1" "--no-nccl-register"
