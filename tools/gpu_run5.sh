set -x
mkdir -p gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02/test_all5.log; tail -6 gpurun_out/r02/test_all5.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02/smoke5.log 2>&1; tail -3 gpurun_out/r02/smoke5.log
for w in cfg4 cfg2 cfg4_shard8; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02/bench5_${w}.json 2> gpurun_out/r02/bench5_${w}.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02/bench_n2.json 2> gpurun_out/r02/bench_n2.err; tail -c 1500 gpurun_out/r02/bench_n2.err; head -c 900 gpurun_out/r02/bench_n2.json
CB_ALLREDUCE=nccl timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 --no-selfcheck > gpurun_out/r02/bench_n2_nccl.json 2> gpurun_out/r02/bench_n2_nccl.err; tail -c 600 gpurun_out/r02/bench_n2_nccl.err
CB_PROFILE_CREATE=1 timeout 300 python tools/e2e_profile.py > gpurun_out/r02/e2e_profile5.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02/bench5_*.json'))+['gpurun_out/r02/bench_n2.json','gpurun_out/r02/bench_n2_nccl.json']:
    try:
        d=json.load(open(f))
        print(f, d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'pp', round(d['roofline']['avg_launch_ms'],4), round(d['roofline']['frac'],3), d.get('parity',{}).get('abs_diff_px'), d.get('nfev_per_step'), json.dumps(d.get('selfcheck',{}))[:300], d.get('allreduce_transport'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -12 gpurun_out/r02/e2e_profile5.log
