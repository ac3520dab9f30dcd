set -x
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02/test_all.log; tail -4 gpurun_out/r02/test_all.log
for g in 0 2; do
  for w in cfg4 cfg2 cfg4_shard8; do
    CB_LM_GRAPH=$g timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02/bench_${w}_graph$g.json 2> gpurun_out/r02/bench_${w}_graph$g.err
  done
done
CB_LM_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02/launches_cfg4.csv python profiles/prof_solve.py cfg4 2 > gpurun_out/r02/prof_cfg4.log 2>&1
CB_LM_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02/launches_shard8.csv python profiles/prof_solve.py cfg4_shard8 2 > gpurun_out/r02/prof_shard8.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02/bench_*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['lm_loop']['trial_replay'], 'pp', d['roofline']['avg_launch_ms'], d['roofline']['frac'], 'sy', d['roofline_tensor']['avg_launch_ms'], d['roofline_tensor']['frac'], d.get('parity',{}).get('abs_diff_px'))
    except Exception as e: print(f, 'ERR', e)
PY
