set -x
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02/test_multi8.log; tail -4 gpurun_out/r02/test_multi8.log
for n in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02/scale_n$n.json 2> gpurun_out/r02/scale_n$n.err
  tail -c 400 gpurun_out/r02/scale_n$n.err
done
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02/scale_n1.json 2> gpurun_out/r02/scale_n1.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r02/cfg5_n8.json 2> gpurun_out/r02/cfg5_n8.err
tail -c 400 gpurun_out/r02/cfg5_n8.err
CB_ALLREDUCE=nccl timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 10 --warmup 3 --no-selfcheck > gpurun_out/r02/scale_n8_nccl.json 2> gpurun_out/r02/scale_n8_nccl.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02/scale_n*.json'))+['gpurun_out/r02/cfg5_n8.json']:
    try:
        d=[json.loads(l) for l in open(f) if l.startswith('{')][-1]; print(f, d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d.get('parity',{}).get('abs_diff_px'), json.dumps(d.get('selfcheck',{}))[:400], d.get('allreduce_transport'))
    except Exception as e: print(f,'ERR',e)
PY
