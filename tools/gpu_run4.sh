set -x
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02/test_all4.log; tail -4 gpurun_out/r02/test_all4.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02/smoke4.log 2>&1; tail -3 gpurun_out/r02/smoke4.log
for w in cfg4 cfg2 cfg4_shard8 cfg3 sparse64 cfg4_intrinsics; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02/bench4_${w}.json 2> gpurun_out/r02/bench4_${w}.err
done
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 > gpurun_out/r02/bench4_cfg5.json 2> gpurun_out/r02/bench4_cfg5.err
timeout 300 python bench.py --workload triangulate_cfg4 --steps 10 --warmup 3 > gpurun_out/r02/bench4_triangulate.json 2> gpurun_out/r02/bench4_triangulate.err
timeout 400 python bench.py --workload bootstrap64 --steps 5 --warmup 2 > gpurun_out/r02/bench4_bootstrap64.json 2> gpurun_out/r02/bench4_bootstrap64.err
CB_PROFILE_CREATE=1 timeout 300 python - > gpurun_out/r02/e2e_profile.log 2>&1 <<'PY'
import time, sys
import numpy as np, torch
sys.path.insert(0, '.')
import caliscope_b200 as cb
from bench import make_workload, make_parameterization
from caliscope_b200 import solver, reprojection as R
rig = make_workload('cfg4'); par = make_parameterization(rig)
cam16 = rig.obs_cam.astype(np.int16); xy = np.array(rig.obs_xy); obj = np.array(rig.obs_pt, dtype=np.int32)
for it in range(4):
    t0 = time.perf_counter()
    p = cb.BAProblem(rig.cam_flags, rig.cam_const, rig.n_pts, cam16, obj, xy)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    r = p.solve(rig.x0); torch.cuda.synchronize(); t2 = time.perf_counter()
    p.close(); t3 = time.perf_counter()
    print(f"iter {it}: create {1e3*(t1-t0):.3f} ms solve {1e3*(t2-t1):.3f} ms (device {r.solve_ms:.3f}) close {1e3*(t3-t2):.3f} ms mode {r.used_graph_mode}", flush=True)
for it in range(3):
    t0 = time.perf_counter()
    res = solver.least_squares(R.joint_residuals, rig.x0, args=(par, cam16, xy, obj, None, None, None, None), jac=R.joint_jacobian, x_scale="jac", method="trf", bounds=par.bounds(), ftol=1e-8)
    print(f"least_squares total {1e3*(time.perf_counter()-t0):.3f} ms", flush=True)
PY
CB_LM_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02/launches4_cfg4.csv python profiles/prof_solve.py cfg4 2 > gpurun_out/r02/prof4_cfg4.log 2>&1
CB_LM_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02/launches4_cfg2.csv python profiles/prof_solve.py cfg2 2 > gpurun_out/r02/prof4_cfg2.log 2>&1
CB_LM_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02/launches4_shard8.csv python profiles/prof_solve.py cfg4_shard8 2 > gpurun_out/r02/prof4_shard8.log 2>&1
CB_LM_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pt_pass|pt_backsub|resjac_kernel|schur_syrk" -c 8 -o gpurun_out/r02/prof4_full python profiles/prof_solve.py cfg4 1 > gpurun_out/r02/prof4_full.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02/bench4_*.json')):
    try:
        d=json.load(open(f))
        if 'roofline_tensor' in d:
            print(f, round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['lm_loop']['trial_replay'][:12], 'pp', round(d['roofline']['avg_launch_ms'],4), round(d['roofline']['frac'],3), 'sy', round(d['roofline_tensor']['avg_launch_ms'],4), round(d['roofline_tensor']['frac'],3), d.get('parity',{}).get('abs_diff_px'), d['nfev_per_step'])
        else:
            print(f, round(d['value'],1), round(d['ms_per_step'],3), json.dumps(d.get('stage_ms', d.get('stages','')))[:300], d.get('parity'))
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/r02/e2e_profile.log | tail -40
