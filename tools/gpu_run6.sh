set -x
mkdir -p gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02/test_all6.log; tail -6 gpurun_out/r02/test_all6.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02/smoke6.log 2>&1; tail -3 gpurun_out/r02/smoke6.log
for w in cfg4 cfg2 cfg4_shard8 cfg3; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02/bench6_${w}.json 2> gpurun_out/r02/bench6_${w}.err
done
CB_FUSE_SMALL=0 timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02/bench6_cfg2_nofuse.json 2> gpurun_out/r02/bench6_cfg2_nofuse.err
CB_PCG_MODE=2 timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02/bench6_cfg2_pcg.json 2> gpurun_out/r02/bench6_cfg2_pcg.err
timeout 400 python bench.py --workload bootstrap64 --steps 5 --warmup 2 > gpurun_out/r02/bench6_bootstrap64.json 2> gpurun_out/r02/bench6_bootstrap64.err
CB_PROFILE_CREATE=1 timeout 300 python tools/e2e_profile.py > gpurun_out/r02/e2e_profile6.log 2>&1
CB_PROFILE_PIPELINE=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --workload cfg5 --steps 3 --warmup 2 > gpurun_out/r02/cfg5_n2.json 2> gpurun_out/r02/cfg5_n2.err
CB_LM_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02/launches6_cfg2.csv python profiles/prof_solve.py cfg2 2 > gpurun_out/r02/prof6_cfg2.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02/bench6_*.json'))+['gpurun_out/r02/cfg5_n2.json']:
    try:
        d=[json.loads(l) for l in open(f) if l.startswith('{')][-1]
        print(f, d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e'].get('ms_per_step',0),3), d.get('roofline',{}).get('chosen'), round(d.get('roofline',{}).get('frac',0),3), d.get('parity',{}).get('abs_diff_px'), d.get('nfev_per_step'), json.dumps(d.get('stage_ms',''))[:300])
    except Exception as e: print(f, 'ERR', e)
PY
tail -8 gpurun_out/r02/e2e_profile6.log
grep pipeline gpurun_out/r02/cfg5_n2.json gpurun_out/r02/cfg5_n2.err | tail -20
