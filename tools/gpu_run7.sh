set -x
mkdir -p gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02/test_all7.log; tail -8 gpurun_out/r02/test_all7.log
timeout 200 python tools/h2d_probe.py > gpurun_out/r02/h2d_probe.log 2>&1; cat gpurun_out/r02/h2d_probe.log
timeout 400 python bench.py --workload bootstrap64 --steps 5 --warmup 2 > gpurun_out/r02/bench7_bootstrap64.json 2> gpurun_out/r02/bench7_bootstrap64.err; tail -3 gpurun_out/r02/bench7_bootstrap64.err
for w in sparse64 cfg4_intrinsics; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02/bench7_${w}.json 2> gpurun_out/r02/bench7_${w}.err
done
CB_PROFILE_PIPELINE=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --workload cfg5 --steps 3 --warmup 2 > gpurun_out/r02/cfg5_n2.json 2> gpurun_out/r02/cfg5_n2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02/bench7_*.json'))+['gpurun_out/r02/cfg5_n2.json']:
    try:
        d=[json.loads(l) for l in open(f) if l.startswith('{')][-1]
        print(f, d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e'].get('ms_per_step',0),3), d.get('roofline',{}).get('chosen'), round(d.get('roofline',{}).get('frac',0),3), d.get('parity'), d.get('nfev_per_step'), json.dumps(d.get('stage_ms',''))[:400], d.get('truth'))
    except Exception as e: print(f, 'ERR', e)
PY
grep pipeline gpurun_out/r02/cfg5_n2.json gpurun_out/r02/cfg5_n2.err | tail -16
