set -x
O=gpurun_out/r02; mkdir -p $O
CB_PROFILE_CREATE=1 timeout 300 python tools/e2e_profile.py > $O/e2e_profile10_nt.log 2>&1; tail -9 $O/e2e_profile10_nt.log
CB_STAGE_TEMPORAL=1 CB_PROFILE_CREATE=1 timeout 300 python tools/e2e_profile.py > $O/e2e_profile10_t.log 2>&1; tail -9 $O/e2e_profile10_t.log
timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench10_cfg4.json 2> $O/bench10_cfg4.err
CB_STAGE_TEMPORAL=1 timeout 300 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench10_cfg4_t.json 2> $O/bench10_cfg4_t.err
python - <<'PY'
import json
for f in ['gpurun_out/r02/bench10_cfg4.json','gpurun_out/r02/bench10_cfg4_t.json']:
    d=[json.loads(l) for l in open(f) if l.startswith('{')][-1]
    print(f, round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))
PY
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -2
