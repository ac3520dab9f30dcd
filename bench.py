#!/usr/bin/env python
"""Bundle-adjustment benchmark (BASELINE.json metric: LM iterations/s on the 64-cam / 50k-point /
2M-observation rig, final RMS reprojection px).

    python bench.py --gpus N --steps K --warmup W          # caliscope_b200 (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W  # scipy TRF on the CPU oracle port

A "step" is one complete bundle-adjustment solve of the synthetic rig from the same perturbed
start: W untimed solves, then exactly K timed solves bracketed by barrier + synchronize, device
time from CUDA events, max over ranks.  ``value`` has the observation list resident in HBM;
``e2e`` goes through the public API with pinned HOST buffers (problem upload + index build +
solve + result download inside the timed region).  N > 1 shards the observations by point with
one all-reduce of the reduced camera system per LM trial (strong scaling: the rig is fixed).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

# NCCL's version banner (NCCL_DEBUG=VERSION, set by some images) goes to stdout ahead of the JSON line
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "lm_iterations_per_sec"
UNIT = "LM iterations/s"

WORKLOADS = {
    # name: (n_cams, n_pts, n_obs, refine_intrinsics)
    "cfg4": (64, 50_000, 2_000_000, False),
    "cfg4_intrinsics": (64, 50_000, 2_000_000, True),
    "cfg3": (16, 10_000, 400_000, True),
    "cfg2": (8, 2_000, 40_000, False),
    # what one rank holds of cfg4 at 8 GPUs: used on ONE GPU to profile the per-iteration fixed cost
    "cfg4_shard8": (64, 6_250, 250_000, False),
    # the realistic Caliscope shape: local visibility, 8 cameras per point (synthetic.sparse64)
    "sparse64": (64, 250_000, 2_000_000, False),
}
# BASELINE.json configs[4]: cfg4 rig + 2 % outliers, solve(linear) -> solve(soft_l1) -> 2.5 % per-camera cull -> solve(linear)
PIPELINE_WORKLOADS = {"cfg5": (64, 50_000, 2_000_000, 0.02), "cfg5_small": (8, 2_000, 40_000, 0.02)}


def make_workload(name: str):
    from caliscope_b200 import synthetic

    n_cams, n_pts, n_obs, refine = WORKLOADS[name]
    if name == "sparse64":
        return synthetic.sparse64()
    return synthetic.make_rig(n_cams, n_pts, n_obs, refine_intrinsics=refine, seed=0, name=name)


def golden_scipy(name: str):
    """scipy's answer for this workload, committed by tests/golden/make_bench_golden.py (same oracle call as the live
    cpu_baseline leg): lets every GPU count print a parity block without 40-200 s of CPU per line."""
    p = ROOT / "tests" / "golden" / "bench_scipy.json"
    try:
        return json.loads(p.read_text()).get(name)
    except Exception:
        return None


def pin_to_gpu_numa(dev: int) -> str:
    """Bind this process to the CPUs local to GPU `dev` (the end-to-end arm is host-memcpy / PCIe sensitive; round 1 saw
    5.5 vs 7.1 ms per solve on the same code depending on where the process landed)."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(dev).pci_bus_id
        dom = torch.cuda.get_device_properties(dev).pci_domain_id
        devid = torch.cuda.get_device_properties(dev).pci_device_id
        path = Path(f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0/local_cpulist")
        cpus: set[int] = set()
        for part in path.read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"pinned to {len(cpus)} CPUs local to GPU {dev} ({path.read_text().strip()})"
    except Exception as e:  # best effort
        return f"not pinned ({type(e).__name__})"
    return "not pinned"


L2_MB = 126.0
FLUSH_BYTES = 256 << 20


def working_set_mb(n_cams: int, n_pts: int, n_obs: int, P: int, n_gpus: int) -> float:
    """Per-GPU working set of one LM iteration: the two observation lists (camera-major 24 B + point-major 20 B per
    observation) + the dense k-major Schur factor (3 n_pts x ceil(n_cams P / 96) 96 doubles)."""
    obs = n_obs * 44 / n_gpus
    zt = 3 * n_pts * (-(-n_cams * P // 96) * 96) * 8 / n_gpus
    return (obs + zt) / 1e6


def needs_l2_flush(n_cams: int, n_pts: int, n_obs: int, P: int, n_gpus: int) -> bool:
    return working_set_mb(n_cams, n_pts, n_obs, P, n_gpus) <= 1.5 * L2_MB


def _l2_note(n_cams: int, n_pts: int, n_obs: int, P: int, n_gpus: int) -> str:
    mb = working_set_mb(n_cams, n_pts, n_obs, P, n_gpus)
    if not needs_l2_flush(n_cams, n_pts, n_obs, P, n_gpus):
        return f"per-iteration working set (observation lists + Schur factor, {mb:.0f} MB per GPU) exceeds the 126 MB L2; no explicit flush"
    return (f"per-iteration working set is {mb:.0f} MB per GPU, so the 126 MB L2 is flushed between the timed steps by writing a "
            f"{FLUSH_BYTES >> 20} MB buffer (each step has its own CUDA-event pair; bracket_ms_per_step includes the flushes)")


def workload_config(name: str, rig, n_gpus: int) -> dict:
    n_cams, n_pts, n_obs, refine = WORKLOADS[name]
    return {
        "workload": f"{name}: synthetic {n_cams}-cam / {n_pts}-point / {n_obs}-observation ring rig, "
        + ("extrinsics + focal scale + k1 + k2" if refine else "extrinsics only")
        + ", seed 0, 0.5 px noise ("
        + {"cfg2": "BASELINE.json configs[1]", "cfg3": "BASELINE.json configs[2]", "cfg4": "BASELINE.json configs[3]",
           "cfg4_intrinsics": "BASELINE.json configs[3] with Pc = 9"}.get(name, "profiling workload, not a BASELINE config")
        + ")",
        "n_cams": n_cams,
        "n_pts": n_pts,
        "n_obs": n_obs,
        "n_params": int(len(rig.x0)),
        "loss": "linear",
        "ftol": 1e-8,
        "sharding": "single GPU" if n_gpus == 1 else f"observations sharded by point over {n_gpus} GPUs, "
        "one sum-all-reduce of the reduced camera system per LM trial (transport: see allreduce_transport)",
        "l2": _l2_note(n_cams, n_pts, n_obs, 9 if refine else 6, n_gpus),
    }


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = (
        "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
        "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, gpu_index: int = 0):
        self.rows: list[list[str]] = []
        self.proc = None
        self.gpu = gpu_index
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "10", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )  # fmt: skip
        except Exception:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(smax)) if smax else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's solver call on the oracle port
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(rig, max_nfev: int):
    """scipy.optimize.least_squares(method='trf', x_scale='jac', jac=<sparse analytic>) exactly as
    capture_volume.py:387-411, on the NumPy restatement of joint_residuals/joint_jacobian."""
    from oracle import ba_oracle as O

    orc = O.Rig(rig.cam_flags, rig.cam_const, rig.n_pts, rig.obs_cam, rig.obs_pt, rig.obs_xy)
    t0 = time.perf_counter()
    res = O.solve_scipy(orc, rig.x0, max_nfev=max_nfev if max_nfev and max_nfev > 0 else None)
    dt = time.perf_counter() - t0
    nit = max(int(res.nit), 1) if res.nfev > 1 else 1
    return {"wall_s": dt, "nit": nit, "nfev": int(res.nfev), "njev": int(res.njev), "status": int(res.status),
            "cost": float(res.cost), "rmse_px": O.overall_rmse_px(res.x, orc)}  # fmt: skip


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rig = make_workload(args.workload)
    small = make_workload("cfg2")
    for _ in range(args.warmup):
        cpu_reference_step(small, 3)
    budget_s = args.ref_budget_s
    steps, its, wall, last = 0, 0, 0.0, None
    for _ in range(args.steps):
        last = cpu_reference_step(rig, args.ref_max_nfev)
        steps += 1
        its += last["nit"]
        wall += last["wall_s"]
        if wall > budget_s:
            break
    value = its / wall
    cores = 1  # threads the solve actually uses: scipy.sparse products and LSMR are single-threaded (os.cpu_count() are visible)
    sample = (
        f"{args.workload} full rig, scipy TRF+LSMR "
        + (f"capped at max_nfev={args.ref_max_nfev}" if args.ref_max_nfev > 0 else "run to convergence (ftol 1e-8)")
        + f" per step ({last['nit']} LM iteration(s), nfev {last['nfev']}, {last['wall_s']:.1f} s each); "
        f"{steps} of {args.steps} requested steps "
        f"timed within the {budget_s:.0f} s budget; warm-up on cfg2; {os.cpu_count()} host cores visible, the SciPy path uses one"
    )
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / steps,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.workload, rig, 1),
        "obs_residuals_per_sec": rig.n_obs * last["nfev"] / last["wall_s"],
        "final_rms_px": last["rmse_px"],
        "gpu_launches": 0,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# the CUDA arm
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(rig) -> dict:
    """SURVEY.md 8(d).  `materialised`: the residual+Jacobian kernel that writes J (184 / 232 B per observation) -- kept for
    comparability with round 1; `fused`: the formulation that never writes J,
    n_obs (24 + 8*3*Pc) + 8 n_pts (6+3) + 8 n_cams (Pc (Pc+1)/2 + Pc) -- what pt_pass_kernel actually has to move
    (observation in, Z block out) and what `roofline.achieved` is computed from."""
    pc = np.where(rig.cam_flags & 1, 9, 6)
    P = int(pc.max())
    return {
        "materialised": int(rig.n_obs * (24 + 16 + 16 * P + 48) + 24 * rig.n_pts + 8 * int((pc + 9).sum())),
        "fused": int(rig.n_obs * (24 + 24 * P) + 72 * rig.n_pts + 8 * rig.n_cams * (P * (P + 1) // 2 + P)),
        "P": P,
    }


def syrk_flops(rig) -> dict:
    """Schur product S = Z Z^T.  dense: every point against every camera pair, upper triangle by 96-wide tiles with
    24-blocks on the diagonal tiles (what the kernel issues when the visibility is dense); algorithmic:
    sum_j 3 (P n_j)^2 / 2 multiply-adds over the cameras n_j that actually see point j."""
    P = int(np.where(rig.cam_flags & 1, 9, 6).max())
    nP = rig.n_cams * P
    nb = -(-nP // 96)
    dense_cols2 = (nb * (nb - 1) // 2) * 96 * 96 + nb * 10 * 24 * 24
    # distinct cameras per point
    key = np.unique(rig.obs_pt.astype(np.int64) * rig.n_cams + rig.obs_cam)
    nj = np.bincount((key // rig.n_cams).astype(np.int64), minlength=rig.n_pts).astype(np.float64)
    return {"dense_flop": 2.0 * 3 * rig.n_pts * dense_cols2, "algorithmic_flop": float(2.0 * np.sum(3 * (P * nj) ** 2 / 2))}


def load_peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_source_hash() -> str:
    import hashlib

    h = hashlib.sha256()  # the kernel sources (headers); the host engine cb_engine.cu launches them but does not change their traffic
    for f in sorted((ROOT / "caliscope_b200" / "csrc").glob("*.cuh")):
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


def load_ncu_traffic(workload: str, kernel: str):
    """DRAM bytes per launch from the committed ncu digest, only if it was taken from THIS kernel source (the digest
    carries the hash of csrc/ at capture time); a stale digest yields None instead of a wrong number."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    try:
        d = json.loads(p.read_text())
        e = d.get(workload, {}).get(kernel)
        if e and e.get("csrc_sha16") == kernel_source_hash():
            return e.get("dram_bytes_per_launch")
    except Exception:
        pass
    return None


def make_parameterization(rig):
    """The reference-shaped ``BundleParameterization`` of a synthetic rig (host mirror class; the reference's own
    class has the same fields) -- what ``CaptureVolume.optimize`` passes as ``args[0]``."""
    from caliscope_b200.bundle_parameterization import BundleParameterization, CameraBlock

    blocks = []
    for c in range(rig.n_cams):
        k = rig.cam_const[c]
        blocks.append(CameraBlock(cam_id=c, free_intrinsics=bool(rig.cam_flags[c] & 1), fx_initial=float(k[0]), fy_initial=float(k[1]),
                                  cx=float(k[2]), cy=float(k[3]), fisheye=bool(rig.cam_flags[c] & 2),
                                  dist_fixed=tuple(float(v) for v in k[6:9]), k1_initial=float(k[4]), k2_initial=float(k[5])))  # fmt: skip
    return BundleParameterization(blocks=tuple(blocks), n_points=rig.n_pts)


def selfcheck_sharded(dev: int, rank: int, world: int) -> dict:
    """N > 1 only: the sharded solve against the SAME rig solved whole on this rank's own GPU, on a small rig with
    duplicate (camera, point) rows, once plain and once with rigid-distance constraint rows; every rank checks its own
    copy and the worst case over ranks is reported.  This is the multi-GPU correctness evidence the 1-GPU test box cannot
    produce (tests/test_gpu_multi.py needs >= 2 GPUs)."""
    import torch
    import torch.distributed as dist

    import caliscope_b200 as cb
    from caliscope_b200 import distributed as D
    from caliscope_b200 import synthetic

    out = {}
    rig = synthetic.make_rig(8, 1500, 24_000, seed=5, name="selfcheck")
    rng = np.random.default_rng(11)
    # rigid pairs: 60 disjoint groups of 8 points, 3 distance rows each (noisy truth => the rows are active)
    Xt = rig.x_true[-3 * rig.n_pts:].reshape(-1, 3)
    ga, gb, dist_, w = [], [], [], []
    for g in range(60):
        pts = np.arange(8 * g, 8 * g + 8)
        for a, b in ((0, 4), (1, 5), (2, 7)):
            A, B = np.repeat(pts[a], 4), np.repeat(pts[b], 4)
            ga.append(A); gb.append(B)
            dist_.append(np.linalg.norm(Xt[pts[a]] - Xt[pts[b]]) * (1 + 1e-3 * rng.normal()))
            w.append(1.0 / (1394.6 * 0.002))
    cons = (np.array(ga, np.int32), np.array(gb, np.int32), np.array(dist_), np.array(w))
    for label, c in (("plain", None), ("constraints", cons)):
        with cb.BAProblem(rig.cam_flags, rig.cam_const, rig.n_pts, rig.obs_cam, rig.obs_pt, rig.obs_xy, constraints=c,
                          device=dev) as p1:
            r1 = p1.solve(rig.x0)
        rs, _ = D.solve_sharded(rig.cam_flags, rig.cam_const, rig.n_pts, rig.obs_cam, rig.obs_pt, rig.obs_xy, rig.x0,
                                device=dev, constraints=c)
        dx = float(np.abs(rs.x - r1.x).max())
        t = torch.tensor([dx, abs(rs.cost - r1.cost) / max(r1.cost, 1e-300), float(rs.nfev != r1.nfev)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[label] = {"max_abs_dx": float(t[0]), "rel_dcost": float(t[1]), "nfev_equal": bool(t[2] == 0),
                      "nfev": int(r1.nfev), "status": int(rs.status)}
    out["ok"] = all(v["max_abs_dx"] < 1e-9 and v["rel_dcost"] < 1e-9 for v in out.values() if isinstance(v, dict))
    out["what"] = ("sharded solve vs the same 8-cam / 1500-point / 24k-observation rig solved whole on each rank's own GPU; "
                   "max over ranks of max|x_sharded - x_single| (bar 1e-9) and relative cost difference")
    return out


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import caliscope_b200 as cb
    from caliscope_b200 import distributed as D
    from caliscope_b200 import reprojection as R
    from caliscope_b200 import solver

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)
    numa = pin_to_gpu_numa(dev)
    stream = torch.cuda.current_stream().cuda_stream
    lib = cb._lib.load()

    rig = make_workload(args.workload)
    ncp = int(np.where(rig.cam_flags & 1, 9, 6).sum())
    if world > 1:
        shard = D.shard_points(rig.obs_cam, rig.obs_pt, rig.obs_xy, rig.n_pts, rank, world)
        l_cam, l_pt, l_xy, l_npts = shard.obs_cam, shard.obs_pt, shard.obs_xy, shard.n_pts
        x0 = D.local_x(rig.x0, ncp, shard)
        transport = D.transport_kwargs(local_rank, n_camera_dims=rig.n_cams * (9 if np.any(rig.cam_flags & 1) else 6))
        cam_order = D.camera_order(rig.obs_cam, rig.obs_pt, rig.n_cams, rig.n_pts, 9 if np.any(rig.cam_flags & 1) else 6)
    else:
        shard = None
        cam_order = None
        l_cam, l_pt, l_xy, l_npts = rig.obs_cam, rig.obs_pt, rig.obs_xy, rig.n_pts
        x0 = rig.x0
        transport = {}
    solve_kw = dict(ftol=1e-8, rank=rank, world_size=world, stream=stream, **transport)
    transport_name = {"peer_group": "peer memory (fused finalize + NVLink reduce kernel, cooperative launch)", "nccl_comm": "engine-owned NCCL",
                      "allreduce": "torch.distributed callback"}.get(next(iter(transport), ""), "none (single GPU)")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm: observation list already in HBM --------------------------------
    d_cam = torch.from_numpy(l_cam).cuda()
    d_pt = torch.from_numpy(l_pt).cuda()
    d_xy = torch.from_numpy(np.ascontiguousarray(l_xy)).cuda()
    prob = cb.BAProblem(rig.cam_flags, rig.cam_const, l_npts, d_cam, d_pt, d_xy, device=dev, stream=stream, cam_order=cam_order)
    res = None
    n_c, n_p, n_o, refine = WORKLOADS[args.workload]
    flush = None
    if needs_l2_flush(n_c, n_p, n_o, 9 if refine else 6, world):
        flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device="cuda")
    barrier()
    for i in range(args.warmup):
        if flush is not None:
            flush.fill_(i & 0x7F)
        res = prob.solve(x0, **solve_kw)
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = lib.cb_ba_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # one CUDA-event pair per step on the solve stream; the L2 flush (when the working set fits L2) sits BETWEEN the pairs,
    # i.e. between timed iterations, not inside them.  The bracket over all steps (flushes included) is reported as well.
    step_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0 = time.perf_counter()
    e0.record()
    nit = nfev = trials = 0
    pp_ms = sy_ms = 0.0
    pp_n = sy_n = 0
    for i in range(args.steps):
        if flush is not None:
            flush.fill_(i & 0x7F)  # evict the previous step's data from L2 (same stream as the solve)
        step_ev[i][0].record()
        res = prob.solve(x0, **solve_kw)
        step_ev[i][1].record()
        nit += res.nit
        nfev += res.nfev
        trials += res.trials_queued
        pp_ms += res.rj_ms
        pp_n += res.rj_launches
        sy_ms += res.syrk_ms
        sy_n += res.syrk_launches
    e1.record()
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    bracket_ms = max_over_ranks(e0.elapsed_time(e1))
    dev_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in step_ev))
    launches = lib.cb_ba_launch_count() - launches0
    used_graph = res.used_graph_mode
    x_final = res.x
    if world > 1:
        x_final = D.gather_points(res.x, ncp, rig.n_pts, shard)
    # roofline pass: the graphs the timed region replays leave no place to read CUDA events back, so the same problem is
    # solved a few more times with every trial launched directly and events around the point pass and the Schur product
    # (CbBaOptions.time_kernels); the pass is itself bracketed by events so that its cost per solve can be compared with
    # the timed region's
    barrier()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_roof = min(args.steps, 5)
    r0.record()
    for i in range(n_roof):
        if flush is not None:
            flush.fill_(i & 0x7F)
        rr = prob.solve(x0, time_kernels=True, **solve_kw)
        pp_ms += rr.rj_ms; pp_n += rr.rj_launches; sy_ms += rr.syrk_ms; sy_n += rr.syrk_launches
    r1.record()
    barrier()
    roof_ms_per_step = max_over_ranks(r0.elapsed_time(r1)) / n_roof
    engine_stats = {"schur_sparse": bool(prob.stat(0)), "schur_flop_issued": prob.stat(1), "direct_reduced_solve": bool(prob.stat(2)),
                    "schur_ctas": int(prob.stat(3))}  # fmt: skip
    prob.close()

    # ---- end-to-end arm: the reference-facing call on pageable NumPy arrays -----------------------
    # N = 1: caliscope_b200.solver.least_squares(joint_residuals, x0, args=(parameterization, camera_indices int16,
    #        image_coords, image_to_world_indices, None x4), jac=..., x_scale="jac", method="trf", bounds=par.bounds(), ...)
    #        exactly as /root/reference/src/caliscope/core/capture_volume.py:387-411 calls scipy: blocks -> arrays,
    #        pageable upload, index build, solve, result download all inside the timed region.
    # N > 1: the same per rank on its shard (pageable host arrays -> BAProblem -> sharded solve).
    par = make_parameterization(rig)
    cam16 = rig.obs_cam.astype(np.int16)  # capture_volume.py:353-355 builds int16 camera indices
    xy_h = np.array(rig.obs_xy)  # pageable copies
    obj_h = np.array(rig.obs_pt, dtype=np.int32)
    lc, lp, lx = np.array(l_cam), np.array(l_pt), np.array(l_xy)

    def e2e_step():
        if world == 1:
            return solver.least_squares(R.joint_residuals, rig.x0, args=(par, cam16, xy_h, obj_h, None, None, None, None),
                                        jac=R.joint_jacobian, x_scale="jac", method="trf", bounds=par.bounds(), ftol=1e-8,
                                        loss="linear", f_scale=1.0, max_nfev=None, verbose=0)
        with cb.BAProblem(rig.cam_flags, rig.cam_const, l_npts, lc, lp, lx, device=dev, stream=stream, cam_order=cam_order) as p2:
            return p2.solve(x0, **solve_kw)

    for _ in range(min(args.warmup, 3)):
        e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t1 = time.perf_counter()
    f0.record()
    e2e_nit = 0
    e_pp_ms = e_sy_ms = 0.0
    e_pp_n = e_sy_n = 0
    for _ in range(args.steps):
        r2 = e2e_step()
        e2e_nit += r2.nit
        e_pp_ms += r2.rj_ms; e_pp_n += r2.rj_launches; e_sy_ms += r2.syrk_ms; e_sy_n += r2.syrk_launches
    f1.record()
    barrier()
    e2e_wall_ms = 1e3 * (time.perf_counter() - t1)
    e2e_ms = max_over_ranks(max(f0.elapsed_time(f1), e2e_wall_ms))
    clocks = sampler.stop() if rank == 0 else None  # sampled across both timed regions
    if world == 1:
        h2d = int(cam16.nbytes * 2 + obj_h.nbytes + xy_h.nbytes + rig.x0.nbytes + rig.cam_flags.nbytes + rig.cam_const.nbytes)
    else:
        h2d = int(lc.nbytes + lp.nbytes + lx.nbytes + x0.nbytes + rig.cam_flags.nbytes + rig.cam_const.nbytes)
    d2h = int(x0.nbytes + 4 * 160)

    selfcheck = selfcheck_sharded(dev, rank, world) if (world > 1 and not args.no_selfcheck) else None

    if rank != 0:
        if world > 1:
            D.close_comms()
            dist.destroy_process_group()
        return

    # ---- parity and baselines (rank 0) ----------------------------------------------------------
    from oracle import ba_oracle as O

    orc = O.Rig(rig.cam_flags, rig.cam_const, rig.n_pts, rig.obs_cam, rig.obs_pt, rig.obs_xy)
    rms_px = O.overall_rmse_px(x_final, orc)
    peak, peak_src = load_peaks()
    ab = algorithmic_bytes(rig)
    pp_bytes = ab["fused"] / world  # each rank's launch covers its shard
    timing_arm = (f"{n_roof} extra solves of the resident problem right after the timed region, trials launched directly with CUDA "
                  f"events around each launch on the solve stream ({roof_ms_per_step:.3f} ms per solve in that mode vs "
                  f"{dev_ms / args.steps:.3f} ms in the timed region, which replays the loop from CUDA graphs)")
    pp_avg_ms = pp_ms / max(pp_n, 1)
    achieved = pp_bytes / (pp_avg_ms * 1e-3) / 1e9 if pp_avg_ms > 0 else 0.0
    sf = syrk_flops(rig)
    sy_avg_ms = sy_ms / max(sy_n, 1)
    dmma, dfma = __import__("ctypes").c_double(), __import__("ctypes").c_double()
    if lib.cb_debug_fp64_peak(dev, __import__("ctypes").byref(dmma), __import__("ctypes").byref(dfma)) != 0:
        dmma.value = 37.1
    P = ab["P"]
    line = {
        "metric": METRIC,
        "value": nit / (dev_ms * 1e-3),
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.workload, rig, world),
        "clocks": clocks,
        "e2e": {
            "value": e2e_nit / (e2e_ms * 1e-3),
            "unit": UNIT,
            "ms_per_step": e2e_ms / args.steps,
            "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h,
            "call": ("caliscope_b200.solver.least_squares(joint_residuals, x0, args=(parameterization, camera_indices[int16], "
                     "image_coords, image_to_world_indices, None, None, None, None), jac=joint_jacobian, x_scale='jac', method='trf', "
                     "bounds=parameterization.bounds(), ftol=1e-8) on pageable NumPy arrays -- the call capture_volume.py:387-411 makes")
            if world == 1 else "per rank: BAProblem(pageable shard arrays) + sharded solve + close",
            "host": numa,
        },
        "gpu_launches": int(launches),
        "lm_loop": {"decision": "on the device (LmState)",
                    "trial_replay": {0: "direct launches", 1: "one CUDA graph per trial", 2: "device loop: one WHILE-conditional graph per solve"}[int(used_graph)],
                    "trials_queued_per_step": trials / args.steps, "host_syncs_per_trial": 0},
        "allreduce_transport": transport_name,
        "obs_residuals_per_sec": rig.n_obs * nfev / (dev_ms * 1e-3),
        "lm_iterations_per_step": nit / args.steps,
        "nfev_per_step": nfev / args.steps,
        "ms_per_lm_iteration": dev_ms / max(nit, 1),
        "wall_ms_per_step": wall_ms / args.steps,
        "bracket_ms_per_step": bracket_ms / args.steps,  # one event pair around all steps, L2 flushes included
        "final_rms_px": rms_px,
        "final_cost": res.cost,
        "status": res.status,
        "roofline_hbm": {
            "kernel": f"pt_pass_kernel<{P},...> (per observation: residual + analytic Jacobian blocks recomputed in registers, "
                      "V/g reduction, 3x3 Cholesky, Z = Jc^T Jp L^-T streamed to the Schur factor; no Jacobian is written)",
            "bound": "hbm",
            "achieved": achieved,
            "peak": peak,
            "unit": "GB/s",
            "frac": achieved / peak,
            "traffic": load_ncu_traffic(args.workload, "pt_pass_kernel") if world == 1 else None,
            "peak_source": peak_src,
            "algorithmic_bytes_per_launch": pp_bytes,
            "algorithmic_bytes_per_obs": {"fused (this kernel, SURVEY 8d minimum)": 24 + 24 * P,
                                          "materialised J (round-1 kernel, SURVEY 8d comparability figure)": 24 + 16 + 16 * P + 48},
            "avg_launch_ms": pp_avg_ms,
            "launches_timed": int(pp_n),
            "share_of_lm_iteration": pp_avg_ms / max(dev_ms / max(nit, 1), 1e-9),
            "timed_in": timing_arm,
        },
        "roofline_tensor": {
            "kernel": "schur_syrk_kernel (S = Z Z^T on mma.sync.m8n8k4.f64, TMA-staged tiles)",
            "bound": "tensor",
            "achieved": engine_stats["schur_flop_issued"] / (sy_avg_ms * 1e-3) / 1e12 if sy_avg_ms > 0 else 0.0,
            "peak": dmma.value,
            "unit": "TFLOP/s",
            "frac": (engine_stats["schur_flop_issued"] / (sy_avg_ms * 1e-3) / 1e12 / dmma.value) if sy_avg_ms > 0 else 0.0,
            "frac_algorithmic": (sf["algorithmic_flop"] / world / (sy_avg_ms * 1e-3) / 1e12 / dmma.value) if sy_avg_ms > 0 else 0.0,
            "peak_source": "measured live: cb_debug_fp64_peak (mma.sync.m8n8k4.f64, 8 warps/SM); DFMA %.1f TFLOP/s" % dfma.value,
            "flop_per_launch": {"issued by this rank's launch (engine: dense tiles, or the compacted row lists on sparse rigs)": engine_stats["schur_flop_issued"],
                                "dense tiles, whole rig / ranks": sf["dense_flop"] / world,
                                "algorithmic sum_j 3 (P n_j)^2, whole rig / ranks": sf["algorithmic_flop"] / world},
            "engine": engine_stats,
            "traffic": load_ncu_traffic(args.workload, "schur_syrk_kernel") if world == 1 else None,
            "avg_launch_ms": sy_avg_ms,
            "launches_timed": int(sy_n),
            "share_of_lm_iteration": sy_avg_ms / max(dev_ms / max(nit, 1), 1e-9),
            "timed_in": timing_arm,
        },
    }
    # "roofline" = the kernel with the largest share of an LM iteration (the Schur product on every rig with more than a
    # handful of cameras; the point pass on tiny ones); the other one stays beside it under its own key
    dom = "roofline_tensor" if line["roofline_tensor"]["share_of_lm_iteration"] >= line["roofline_hbm"]["share_of_lm_iteration"] else "roofline_hbm"
    line["roofline"] = dict(line[dom], dominant_of=["roofline_hbm", "roofline_tensor"], chosen=dom)
    gold = golden_scipy(args.workload)
    if gold is not None:
        bar = 1e-6
        line["parity"] = {
            "rms_px_gpu": rms_px, "rms_px_scipy": gold["scipy_rms_px"], "abs_diff_px": abs(rms_px - gold["scipy_rms_px"]),
            "bar_px": bar, "green": bool(abs(rms_px - gold["scipy_rms_px"]) < bar),
            "cost_gpu": res.cost, "cost_scipy": gold["scipy_cost"],
            "scipy": f"committed: tests/golden/bench_scipy.json[{args.workload}] (nfev {gold['scipy_nfev']}, nit {gold['scipy_nit']}, {gold['wall_s']} s), "
                     "generated by tests/golden/make_bench_golden.py = oracle.ba_oracle.solve_scipy",
            "note": "P = 9 (free intrinsics): scipy's own default-vs-tight runs differ by up to 3e-6 px on small rigs (DESIGN.md 2)" if P == 9 else "",
        }  # fmt: skip
    if selfcheck is not None:
        line["selfcheck"] = selfcheck
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_step(rig, args.ref_max_nfev)
        line["cpu_baseline"] = {
            "value": cpu["nit"] / cpu["wall_s"],
            "unit": UNIT,
            "cores": 1,  # scipy.sparse matvec / LSMR are single-threaded; os.cpu_count() cores are visible
            "kind": "port",
            "sample": f"{args.workload} full rig, one scipy TRF+LSMR solve on the NumPy oracle port "
            + (f"capped at max_nfev={args.ref_max_nfev}" if args.ref_max_nfev > 0 else "to convergence (ftol 1e-8)")
            + f": {cpu['nit']} LM iteration(s), nfev {cpu['nfev']}, {cpu['wall_s']:.1f} s "
            "(scipy.sparse matvec/LSMR are single-threaded)",
            "final_rms_px": cpu["rmse_px"],
            "final_cost": cpu["cost"],
            "obs_residuals_per_sec": rig.n_obs * cpu["nfev"] / cpu["wall_s"],
        }
        if args.ref_max_nfev <= 0:
            line.setdefault("parity", {})
            line["parity"].update({"rms_px_scipy_live": cpu["rmse_px"], "abs_diff_px_live": abs(rms_px - cpu["rmse_px"]),
                                   "bar_px": 1e-6, "green_live": bool(abs(rms_px - cpu["rmse_px"]) < 1e-6)})  # fmt: skip
    print(json.dumps(line), flush=True)
    if world > 1:
        D.close_comms()
        dist.destroy_process_group()


def run_pipeline(args) -> None:
    """Outlier-filter + re-solve loop (calibrate_extrinsics.py:206-250); one step = the whole loop from host buffers
    (problem upload + index build twice per step).  N > 1: observations sharded by point, global per-camera
    percentile thresholds from one all-gather of the pixel errors (pipeline.solve_filter_resolve_sharded)."""
    import torch
    import torch.distributed as dist

    from caliscope_b200 import distributed as D
    from caliscope_b200 import pipeline, synthetic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)
    numa = pin_to_gpu_numa(dev)
    n_cams, n_pts, n_obs, frac = PIPELINE_WORKLOADS[args.workload]
    rig = synthetic.make_rig(n_cams, n_pts, n_obs, seed=0, outlier_frac=frac, name=args.workload)
    if world == 1:
        call = lambda: pipeline.solve_filter_resolve(rig.cam_flags, rig.cam_const, rig.n_pts, rig.obs_cam, rig.obs_pt,
                                                     rig.obs_xy, rig.x0)  # noqa: E731
    else:
        call = lambda: pipeline.solve_filter_resolve_sharded(rig.cam_flags, rig.cam_const, rig.n_pts, rig.obs_cam,
                                                             rig.obs_pt, rig.obs_xy, rig.x0, device=dev)  # noqa: E731

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        out = call()
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    nit = 0
    for _ in range(args.steps):
        out = call()
        nit += sum(s.nit for s in out.stages)
    barrier()
    wall = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        D.close_comms()
        dist.destroy_process_group()
        return
    from oracle import ba_oracle as O
    from oracle import filtering as OF

    gold = golden_scipy("cfg5_stage1") if args.workload == "cfg5" else None
    orc3 = O.Rig(rig.cam_flags, rig.cam_const, rig.n_pts, rig.obs_cam[out.keep], rig.obs_pt[out.keep], rig.obs_xy[out.keep])
    line = {
        "metric": METRIC, "value": nit / wall, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {n_cams}-cam / {n_pts}-point / {n_obs}-observation rig with {frac:.0%} outliers, "
                   "solve(linear) -> solve(soft_l1, ftol 1e-4) -> 2.5 % per-camera cull -> solve(linear); host buffers, "
                   "problem upload + index build twice per step (BASELINE.json configs[4])",
                   "sharding": "single GPU" if world == 1 else f"observations sharded by point over {world} GPUs; global per-camera "
                   "percentile thresholds from one all-gather of the pixel errors; keep mask and compaction local to each rank",
                   "host": numa},
        "clocks": clocks,
        "e2e": {"value": nit / wall, "unit": UNIT, "ms_per_step": 1e3 * wall / args.steps,
                "h2d_bytes_per_step": int((rig.obs_cam.nbytes + rig.obs_pt.nbytes + rig.obs_xy.nbytes) / world + 3 * rig.x0.nbytes),
                "d2h_bytes_per_step": int(3 * rig.x0.nbytes + rig.n_obs / world)},
        "stages": [{"loss": l, "nfev": s.nfev, "nit": s.nit, "status": s.status, "cost": s.cost, "solve_ms": s.solve_ms,
                    "rms_px": r} for l, s, r in zip(("linear", "soft_l1", "linear after cull"), out.stages, out.rmse_px)],
        "kept_fraction": float(out.keep.mean()),
        "outliers_removed_fraction": float(1.0 - out.keep[rig.outlier_mask].mean()),
        "final_rms_px": float(O.overall_rmse_px(out.x, orc3)),
    }  # fmt: skip
    if gold is not None:
        line["parity"] = {"stage": "1 (linear, all observations)", "rms_px_gpu": out.rmse_px[0], "rms_px_scipy": gold["scipy_rms_px"],
                          "abs_diff_px": abs(out.rmse_px[0] - gold["scipy_rms_px"]), "bar_px": 1e-6,
                          "green": bool(abs(out.rmse_px[0] - gold["scipy_rms_px"]) < 1e-6),
                          "scipy": "committed: tests/golden/bench_scipy.json[cfg5_stage1]"}  # fmt: skip
    print(json.dumps(line), flush=True)
    if world > 1:
        D.close_comms()
        dist.destroy_process_group()


def run_bootstrap(args) -> None:
    """The extrinsic bootstrap (what produces bundle adjustment's start vector) on a synthetic
    64-camera board session: one step = PnP of every (camera, frame, board) group + relative poses + IQR outlier rule +
    quaternion average + stereo RMSE of every camera pair, from host arrays."""
    import torch

    from caliscope_b200 import bootstrap as B
    from caliscope_b200 import synthetic

    n_frames = 1000
    ses = synthetic.make_board_session(64, n_frames, seed=0)
    tab = B.CameraTables(ses.cam_ids, {int(c): i for i, c in enumerate(ses.cam_ids)}, ses.cam_k, ses.cam_dist, ses.cam_fisheye,
                         np.zeros(len(ses.cam_ids), bool), np.ones(len(ses.cam_ids), bool))  # fmt: skip
    numa = pin_to_gpu_numa(0)
    stats = {}

    def step():
        t0 = time.perf_counter()
        res = B.pnp_arrays(tab, ses.cam_id, ses.sync_index, ses.object_id, ses.img_xy, ses.obj_xyz)
        t1 = time.perf_counter()
        live = res.status != B.PNP_TOO_FEW
        pairs, R, t, cnt, _, nst = B.pose_network_arrays(res.keys[live], res.R[live], res.t[live], tab, 1.5)
        t2 = time.perf_counter()
        rmse, ncom = B.stereo_rmse_arrays(tab, pairs, R, t, ses.cam_id, ses.sync_index, ses.object_id, ses.keypoint_id, ses.img_xy)
        t3 = time.perf_counter()
        stats.update(pnp_s=t1 - t0, host_s=t2 - t1, stereo_s=t3 - t2, groups=len(res.keys), rel=int(cnt.sum()), pairs=len(pairs),
                     net_kernel_ms=nst.total_ms,
                     pnp_kernel_ms=res.kernel_ms, launches=res.launches, fallback=int((res.status == B.PNP_OK_FALLBACK).sum()))
        return res, pairs, R, t, rmse

    for _ in range(args.warmup):
        out = step()
    torch.cuda.synchronize()
    launches0 = B.L.load().cb_ba_launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    launches = B.L.load().cb_ba_launch_count() - launches0
    res, pairs, R, t, rmse = out
    # truth: aggregated relative poses against the generator's cameras
    err_R = err_t = 0.0
    for k, (a, b) in enumerate(pairs):
        Ra, Rb = synthetic._rot(ses.rvec[a]), synthetic._rot(ses.rvec[b])
        err_R = max(err_R, float(np.abs(R[k] - Rb @ Ra.T).max()))
        err_t = max(err_t, float(np.abs(t[k] - (ses.tvec[b] - Rb @ Ra.T @ ses.tvec[a])).max()))
    line = {
        "metric": "bootstrap_observations_per_sec", "value": ses.n_obs * args.steps / wall, "unit": "observations/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"bootstrap64: 64-camera ring, 7x5-corner board, {n_frames} frames, {ses.n_obs} observations, "
                   f"{stats['groups']} PnP groups, {stats['rel']} relative poses kept, {stats['pairs']} camera pairs; host arrays in, "
                   "poses out (uploads and sorts inside the timed region)", "host": numa},
        "e2e": {"value": ses.n_obs * args.steps / wall, "unit": "observations/s", "ms_per_step": 1e3 * wall / args.steps,
                "h2d_bytes_per_step": int(ses.n_obs * (4 + 8 + 8 + 24) + ses.n_obs * (4 + 8 + 8)),
                "d2h_bytes_per_step": int(stats["groups"] * (72 + 24 + 8 + 12) + stats["pairs"] * 16)},
        "gpu_launches": int(launches),
        "stage_ms": {"pnp (device call)": 1e3 * stats["pnp_s"], "relative + IQR + average (device call)": 1e3 * stats["host_s"],
                     "relative + IQR + average kernels alone": stats["net_kernel_ms"],
                     "stereo rmse (device call)": 1e3 * stats["stereo_s"], "pnp kernel alone": stats["pnp_kernel_ms"]},
        "fallback_groups": stats["fallback"],
        "truth": {"max_abs_dR": err_R, "max_abs_dt_m": err_t, "median_stereo_rmse": float(np.nanmedian(rmse))},
    }  # fmt: skip
    if not args.no_cpu_baseline:
        from oracle import bootstrap as OB

        nf = 40  # bounded sample: the first 40 frames through the reference's own OpenCV calls
        m = ses.sync_index < nf
        t1 = time.perf_counter()
        poses_cv, pairs_cv = OB.reference_calls_cv2(ses.cam_ids, ses.cam_k, ses.cam_dist, ses.cam_fisheye, ses.sync_index[m], ses.cam_id[m],
                                                    ses.object_id[m], ses.keypoint_id[m], ses.img_xy[m], ses.obj_xyz[m])
        dt = time.perf_counter() - t1
        # parity on the sample: device PnP against cv2's on the same groups
        rs = B.pnp_arrays(tab, ses.cam_id[m], ses.sync_index[m], ses.object_id[m], ses.img_xy[m], ses.obj_xyz[m])
        worst = 0.0
        for i, k in enumerate(rs.keys):
            kk = tuple(int(v) for v in k)
            if kk in poses_cv and rs.status[i] == B.PNP_OK and np.isfinite(poses_cv[kk][0]).all():
                worst = max(worst, float(np.abs(rs.R[i] - poses_cv[kk][0]).max()), float(np.abs(rs.t[i] - poses_cv[kk][1]).max()))
        line["cpu_baseline"] = {"value": int(m.sum()) / dt, "unit": "observations/s", "cores": 1, "kind": "reference",
                                "sample": f"first {nf} frames ({int(m.sum())} observations, {len(poses_cv)} PnP groups, {len(pairs_cv)} pairs): "
                                "the reference's own OpenCV calls (cv2.undistortPoints, solvePnP(IPPE), Rodrigues, projectPoints, "
                                "triangulatePoints) in its per-group / per-pair Python loops, on arrays instead of DataFrames"}
        line["parity"] = {"pnp_pose_max_abs_diff_vs_cv2": worst, "bar": 1e-6, "green": bool(worst < 1e-6)}
    print(json.dumps(line), flush=True)


def run_triangulation(args) -> None:
    """The step in front of bundle adjustment on the cfg4 rig: undistort 2 M pixel
    observations and DLT-triangulate the 50 000 points, host buffers in and out; one step = both calls."""
    import torch

    from caliscope_b200 import synthetic
    from caliscope_b200 import triangulation as T

    rig = synthetic.cfg4()
    proj, _ = synthetic.exact_normalized_observations(rig)
    K = np.array([[rig.cam_const[0, 0], 0, rig.cam_const[0, 2]], [0, rig.cam_const[0, 1], rig.cam_const[0, 3]], [0, 0, 1.0]])
    mats = np.tile(K, (rig.n_cams, 1, 1))
    dists = [np.array([c[4], c[5], c[6], c[7], c[8]]) for c in rig.cam_const]
    fish = np.zeros(rig.n_cams, np.int32)
    key = rig.obs_pt.astype(np.int64)
    st = T.TriangulationStats()

    def step():
        return T.triangulate_groups(proj, rig.obs_cam, key, rig.obs_xy, stats=st, undistort=(mats, dists, fish))

    for _ in range(args.warmup):
        xyz, count, _, _ = step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dlt_ms = 0.0
    launches0 = T.L.load().cb_ba_launch_count()
    for _ in range(args.steps):
        xyz, count, _, _ = step()
        dlt_ms += st.dlt_ms
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    launches = T.L.load().cb_ba_launch_count() - launches0
    truth = rig.x_true[-3 * rig.n_pts:].reshape(-1, 3)
    peak, peak_src = load_peaks()
    alg = 24 * rig.n_obs + 56 * rig.n_pts  # row index 4 + camera 4 + xy 16 per observation; start 4+4, xyz 24, count/rep 8, sig 16 per group
    line = {
        "metric": "triangulated_points_per_sec", "value": rig.n_pts * args.steps / wall, "unit": "points/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "triangulate_cfg4: undistort 2 000 000 pixel observations (64 pinhole cameras) + DLT of 50 000 "
                   "points in one cb_undistort_triangulate call; host buffers in and out, uploads inside the timed "
                   "region (inputs 56 MB + sort buffers > L2)"},
        "e2e": {"value": rig.n_pts * args.steps / wall, "unit": "points/s", "ms_per_step": 1e3 * wall / args.steps,
                "h2d_bytes_per_step": int(rig.n_obs * (16 + 4 + 8)),
                "d2h_bytes_per_step": int(rig.n_pts * (24 + 4 + 4 + 16))},
        "gpu_launches": int(launches),
        "max_abs_error_vs_truth_m": float(np.abs(xyz - truth).max()),
        "roofline": {"kernel": "tri_dlt_kernel (gather rows, 4x4 normal matrix, Jacobi eigenvector)", "bound": "hbm",
                     "achieved": alg / (dlt_ms / args.steps * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / (dlt_ms / args.steps * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(alg), "avg_launch_ms": dlt_ms / args.steps},
    }  # fmt: skip
    if not args.no_cpu_baseline:
        from oracle import triangulation as OT

        n_s = 2000
        m = rig.obs_pt < n_s
        pm = {i: proj[i] for i in range(rig.n_cams)}
        und = OT.undistort_points(rig.obs_xy[m], K, dists[0], False)
        z = np.zeros(int(m.sum()), np.int64)
        t1 = time.perf_counter()
        OT.triangulate_image_points(pm, z, rig.obs_cam[m].astype(np.int64), z, rig.obs_pt[m].astype(np.int64), und)
        dt = time.perf_counter() - t1
        line["cpu_baseline"] = {"value": n_s / dt, "unit": "points/s", "cores": 1, "kind": "port",
                                "sample": f"first {n_s} points ({int(m.sum())} observations) of the same rig, one SVD per point"}
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS) + sorted(PIPELINE_WORKLOADS) + ["triangulate_cfg4", "bootstrap64"], default="cfg4")
    ap.add_argument("--ref-max-nfev", type=int, default=0, help="cap on scipy evaluations per reference step (0: run to convergence)")
    ap.add_argument("--ref-budget-s", type=float, default=240.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-selfcheck", action="store_true", help="N > 1: skip the sharded-vs-single-GPU equality check on the small rig")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.workload == "bootstrap64":
        if args.impl != "ours" or args.gpus != 1:
            raise SystemExit("bootstrap64 runs on the CUDA arm, one GPU")
        run_bootstrap(args)
    elif args.workload == "triangulate_cfg4":
        if args.impl != "ours" or args.gpus != 1:
            raise SystemExit("triangulate_cfg4 runs on the CUDA arm, one GPU")
        run_triangulation(args)
    elif args.workload in PIPELINE_WORKLOADS:
        if args.impl != "ours":
            raise SystemExit("pipeline workloads run on the CUDA arm")
        run_pipeline(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
