set -x
O=gpurun_out/r02/final8; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -30 > $O/test_multi8.log; tail -4 $O/test_multi8.log
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/scale_n1.json 2> $O/scale_n1.err
for n in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 10 --warmup 3 > $O/scale_n$n.json 2> $O/scale_n$n.err
  tail -c 300 $O/scale_n$n.err
done
CB_PROFILE_PIPELINE=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --workload cfg5 --steps 5 --warmup 3 > $O/cfg5_n8.json 2> $O/cfg5_n8.err
tail -c 300 $O/cfg5_n8.err
CB_ALLREDUCE=nccl timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 10 --warmup 3 --no-selfcheck > $O/scale_n8_nccl.json 2> $O/scale_n8_nccl.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02/final8/scale_n*.json'))+['gpurun_out/r02/final8/cfg5_n8.json']:
    try:
        d=[json.loads(l) for l in open(f) if l.startswith('{')][-1]; print(f, d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d.get('parity',{}).get('abs_diff_px'), json.dumps(d.get('selfcheck',{}))[:200], d.get('allreduce_transport'), json.dumps(d.get('stages',''))[:300])
    except Exception as e: print(f,'ERR',e)
PY
grep pipeline $O/cfg5_n8.json $O/cfg5_n8.err | tail -16
