#!/usr/bin/env python
"""Benchmark of the sCBCT generation loop (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2|C1] [--impl ours|reference]

A step = one whole sCBCT volume: two-channel forward projection of (planning CT, artifact
volume) with the scatter combine in its epilogue, Poisson + electronic noise, and OS-SART
(20 iterations x 33 subsets at C2) -- synthetic seeded phantom, fp32.  `value` times the
steps with inputs resident in HBM; `e2e` times the public host-buffer call
(SCBCTPipeline.run_host: pinned host volumes in, reconstructed volume back to the host).
For N > 1 the angles of each OS-SART subset are interleaved over the ranks and the partial
updates are summed with one NCCL all-reduce per subset (strong scaling: the workload is
fixed).  `--impl reference` times the float64 CPU oracle (the reference has no CPU
implementation of this path; see oracle/oracle.c) on the host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sCBCT volumes/sec (proj+scatter+OS-SART)"
UNIT = "volumes/s"
SCATTER_SCALE = 1.0


def get_config(name):
    from physics_arx_b200 import phantom
    cfg = {"C1": phantom.config_c1, "C2": phantom.config_c2}[name]()
    return cfg


def describe(cfg, n_gpus):
    """The workload, identically for both arms (the rank grid actually used is reported under `sharding`)."""
    geo = cfg["geo"]
    nsub = -(-len(cfg["angles"]) // cfg["blocksize"])
    return {
        "workload": f"{cfg['name']}: planning CT {int(geo.nVoxel[1])}x{int(geo.nVoxel[2])}x{int(geo.nVoxel[0])}, "
                    f"{len(cfg['angles'])} projections of {int(geo.nDetector[1])}x{int(geo.nDetector[0])}"
                    f"{' half-fan' if float(np.asarray(geo.offDetector)[1]) != 0 else ''}, scatter+noise, "
                    f"{cfg['niter']}-iteration OS-SART (blocksize {cfg['blocksize']}, {nsub} subsets)",
        "niter": cfg["niter"], "blocksize": cfg["blocksize"], "I0": cfg["I0"], "sigma": cfg["sigma"],
        "accuracy": float(geo.accuracy),
        "parallelism": "1 GPU" if n_gpus == 1 else (
            f"{n_gpus} GPUs, one process each: the rays of every OS-SART subset sharded over the ranks, volume "
            "replicated, partial volume updates summed over NVLink once per subset"),
        "l2": "per-step working set (projections + volumes, ~2.4 GB at C2) exceeds the 126 MB L2; no flush",
    }


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.path = None, None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 8:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for nm, val in zip(names, f[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------- CPU baseline
def cpu_sample(cfg, n_sample_angles=4):
    """Time the oracle on a bounded sample of the workload: Ax + Atb of a few angles."""
    from oracle import cpu_oracle
    from physics_arx_b200.phantom import lung_phantom
    cpu_oracle.build()
    # every host core, explicitly: under torch.distributed.run the inherited OMP_NUM_THREADS is 1
    cores = cpu_oracle.set_threads(os.cpu_count() or 1)
    geo, angles = cfg["geo"], cfg["angles"]
    n = len(angles)
    pick = np.linspace(0, n, n_sample_angles, endpoint=False).astype(int)
    vol = lung_phantom(geo.nVoxel, geo.dVoxel, seed=0).astype(np.float64)
    t0 = time.perf_counter()
    proj = cpu_oracle.ax(vol, geo, angles[pick])
    t1 = time.perf_counter()
    cpu_oracle.atb(proj, geo, angles[pick])
    t2 = time.perf_counter()
    t_ax, t_atb = (t1 - t0) / len(pick), (t2 - t1) / len(pick)
    # one volume = 1 simulation Ax + niter * (Ax + Atb) over all angles (V set-up not counted)
    t_vol = n * t_ax + cfg["niter"] * n * (t_ax + t_atb)
    return {"value": 1.0 / t_vol, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"float64 oracle (oracle/oracle.c, OpenMP) Ax+Atb of {len(pick)} of the {n} angles at full "
                      f"{cfg['name']} size: {t_ax:.2f} s/angle Ax, {t_atb:.2f} s/angle Atb; extrapolated to "
                      f"1+{cfg['niter']} Ax and {cfg['niter']} Atb passes",
            "t_ax_per_angle_s": t_ax, "t_atb_per_angle_s": t_atb}


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    vals = []
    info = None
    for i in range(args.warmup + args.steps):
        info = cpu_sample(cfg, n_sample_angles=2)
        if i >= args.warmup:
            vals.append(info["value"])
    v = float(np.mean(vals))
    info["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": describe(cfg, args.gpus), "cpu_baseline": info,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "the reference has no CPU implementation of this path (TIGRE is CUDA-only, un-vendored); "
                    "this arm times the build's float64 oracle on the host cores, each step a bounded sample"}
    print(json.dumps(line), flush=True)


# --------------------------------------------- GPU-vs-GPU anchor: the recalled TIGRE launch shapes on this GPU
def run_tigre_style(args, cfg):
    """One OS-SART subset (blocksize angles) of the workload with the TIGRE-style baseline kernels
    (csrc/tigre_style.cu: 9x9x9 hardware-trilinear texture Ax, 16x32x8z texture Atb) beside the product kernels:
    ms per launch, error of each against the float64 oracle on two angles, and the volume time the Ax / Atb
    kernels alone would add up to (1 + niter Ax passes, niter Atb passes)."""
    import ctypes
    import torch
    from oracle import cpu_oracle
    from physics_arx_b200 import _lib
    from physics_arx_b200.phantom import lung_phantom
    from physics_arx_b200.projector import Plan, _ptr, _stream
    lib = _lib.load()
    torch.cuda.set_device(0)
    geo, angles_all = cfg["geo"], cfg["angles"]
    b = cfg["blocksize"]
    angles = angles_all[np.linspace(0, len(angles_all) - 1, b).astype(int)]       # one subset's worth, spread
    plan = Plan(geo, angles)
    vol_np = lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)
    vol = torch.from_numpy(vol_np).cuda()
    res = plan.pack(vol)
    ts = ctypes.c_void_p()
    _lib.check(lib.pax_ts_create(plan._h, ctypes.byref(ts)), "pax_ts_create")
    proj_ts = torch.empty((b, plan.nv, plan.nu), dtype=torch.float32, device="cuda")
    back_ts = torch.empty((plan.nz, plan.ny, plan.nx), dtype=torch.float32, device="cuda")
    back = plan.new_volume()

    def timed(fn):
        for _ in range(max(args.warmup, 1)):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(max(args.steps, 1)):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / max(args.steps, 1)

    _lib.check(lib.pax_ts_set_volume(ts, _ptr(vol), _stream()), "pax_ts_set_volume")
    t_up = timed(lambda: _lib.check(lib.pax_ts_set_volume(ts, _ptr(vol), _stream()), "pax_ts_set_volume"))
    t_ax_ts = timed(lambda: _lib.check(lib.pax_ts_ax(ts, 0, b, _ptr(proj_ts), _stream()), "pax_ts_ax"))
    proj = plan.ax(res)
    t_ax = timed(lambda: plan.ax(res, out=proj))
    t_atb_ts = timed(lambda: _lib.check(lib.pax_ts_atb(ts, _ptr(proj), 0, b, _ptr(back_ts), _stream()), "pax_ts_atb"))
    t_atb = timed(lambda: plan.atb(proj, 0, b, back, _lib.PAX_ATB_STORE))
    # accuracy of both against the float64 oracle (two angles)
    pick = [0, b // 2]
    ref = cpu_oracle.ax(vol_np, geo, angles[pick])

    def rel(a, r):
        return float(np.linalg.norm(a.astype(np.float64) - r) / np.linalg.norm(r))

    err_ts = rel(proj_ts[pick].cpu().numpy(), ref)
    err = rel(proj[pick].cpu().numpy(), ref)
    back_ref = cpu_oracle.atb(proj.cpu().numpy(), geo, angles)
    err_atb_ts = rel(back_ts.cpu().numpy(), back_ref)
    err_atb = rel(plan.unpack(back).cpu().numpy(), back_ref)
    lib.pax_ts_destroy(ts)
    nsub = -(-len(angles_all) // b)
    niter = cfg["niter"]
    vol_ts = (1 + niter) * nsub * t_ax_ts + niter * nsub * (t_atb_ts + t_up)
    vol_us = (1 + niter) * nsub * t_ax + niter * nsub * t_atb
    line = {"impl": "tigre_style", "metric": METRIC, "unit": UNIT, "n_gpus": 1, "config": describe(cfg, 1),
            "value": 1e3 / vol_ts, "ms_per_step": vol_ts, "higher_is_better": True, "dtype": "f32",
            "data": "synthetic", "steps": args.steps, "warmup": args.warmup,
            "what": f"Ax and Atb kernels only, one {b}-angle subset timed and multiplied out to 1+{niter} Ax and "
                    f"{niter} Atb passes over {nsub} subsets; TIGRE itself also moves every array over PCIe per "
                    "call (SURVEY F8), which is not charged here",
            "tigre_style": {"ax_ms_per_subset": t_ax_ts, "atb_ms_per_subset": t_atb_ts,
                            "volume_texture_upload_ms": t_up, "ax_rel_l2_vs_oracle": err_ts,
                            "atb_rel_l2_vs_oracle": err_atb_ts, "kernels_only_ms_per_volume": vol_ts},
            "ours": {"ax_ms_per_subset": t_ax, "atb_ms_per_subset": t_atb, "ax_rel_l2_vs_oracle": err,
                     "atb_rel_l2_vs_oracle": err_atb, "kernels_only_ms_per_volume": vol_us,
                     "projector": plan.projector},
            "speedup_kernels_only": vol_ts / vol_us}
    print(json.dumps(line), flush=True)


def collective_name(st):
    if st.scheme == "gather":
        return ("residual all-gather + row-owned backprojection/update + row broadcast (multimem kernels), "
                + st.gsym["kind"])
    if st.fused is not None:
        return "fused reduce + update + broadcast kernel over " + st.fused["kind"]
    return "NCCL all-reduce + local update kernel"


# ------------------------------------------------------------- multi-GPU parity (untimed)
def parity_vs_single_gpu(group, rank, world):
    """A small case (tests/test_multigpu.py's geometry) run sharded over the group -- once as a caller gets it
    (the gather scheme), once with the reduce scheme and its fused reduce+update+broadcast kernel forced --
    against the same case on rank 0 alone."""
    import torch
    import torch.distributed as dist
    from physics_arx_b200.phantom import artifact_volume, lung_phantom, make_geometry
    from physics_arx_b200.pipeline import SCBCTPipeline
    geo = make_geometry((16, 40, 40), (10.0, 6.0, 6.0), (40, 72), (6.0, 4.0), off_detector=(0.0, -30.0))
    angles = np.linspace(0, 2 * np.pi, 23)
    ct = torch.from_numpy(lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)).cuda()
    art = torch.from_numpy(artifact_volume(geo.nVoxel, seed=1)).cuda()
    kw = dict(niter=3, blocksize=5, I0=1e8, sigma=4.0, scatter_scale=0.5, seed=3)
    single = SCBCTPipeline(geo, angles, **kw).run_device(ct, art).clone() if rank == 0 else None
    out = {}
    for name, fused, shard in (("default", None, "auto"), ("fused_reduce", True, "grid")):
        try:
            pipe = SCBCTPipeline(geo, angles, group=group, fused_update=fused, shard=shard, **kw)
            got = pipe.run_device(ct, art).clone()
            coll = collective_name(pipe.st)
            ref = got.clone()
            dist.broadcast(ref, src=0, group=group)
            same = torch.tensor([1 if torch.equal(got, ref) else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(same, op=dist.ReduceOp.MIN, group=group)
            entry = {"collective": coll, "grid": list(pipe.st.grid), "bit_identical_across_ranks": bool(same.item())}
            if rank == 0:
                hu = (got - single).abs().double() * 4095.0
                entry.update(mean_hu=float(hu.mean().item()), max_hu=float(hu.max().item()))
            out[name] = entry
            del pipe
        except Exception as e:                          # noqa: BLE001 -- report, do not hide
            out[name] = {"error": repr(e)[:200]}
    return out


# ------------------------------------------------- BASELINE.json's other configurations (untimed region)
def secondary_configs():
    """C1 / C4 / C5 of BASELINE.json on one GPU, a fraction of a second each (parity cases, not the headline)."""
    import torch
    from physics_arx_b200.algorithms import ossart
    from physics_arx_b200.drivers import (VariationBatch, config_c4, motion_projections, phase_volumes,
                                          respiratory_phases)
    from physics_arx_b200.phantom import artifact_volume, config_c1, lung_phantom
    from physics_arx_b200.pipeline import SCBCTPipeline

    def timed(fn, reps=2):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    cfg = config_c1()
    geo, angles = cfg["geo"], cfg["angles"]
    ct = lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)
    art = artifact_volume(geo.nVoxel, seed=1)
    ct_d, art_d = torch.from_numpy(ct).cuda(), torch.from_numpy(art).cuda()
    out = {}
    pipe = SCBCTPipeline(geo, angles, niter=cfg["niter"], blocksize=cfg["blocksize"], I0=cfg["I0"],
                         sigma=cfg["sigma"], scatter_scale=0.5, seed=0)
    out["c1_ms_per_volume"] = timed(lambda: pipe.run_device(ct_d, art_d))
    out["c1_projector"] = pipe.plan.projector
    c4 = config_c4()
    vb = VariationBatch(geo, angles, niter=cfg["niter"], blocksize=cfg["blocksize"], sigma=cfg["sigma"])
    n = len(c4["scatter_scales"]) * len(c4["I0s"]) * len(c4["strides"])
    t = timed(lambda: vb.run(ct_d, art_d, c4["scatter_scales"], c4["I0s"], c4["strides"], to_numpy=False), reps=1)
    out["c4_ms_per_variation"] = t / n
    out["c4_variations"] = n
    ph = respiratory_phases(len(angles))
    pv = torch.from_numpy(phase_volumes(ct, float(geo.dVoxel[0]))).cuda()
    out["c5_ms"] = timed(lambda: ossart(motion_projections(pv, ph, geo, angles), geo, angles, cfg["niter"],
                                        blocksize=cfg["blocksize"]))
    return out


# ------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="C2", choices=["C1", "C2"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "tigre_style"],
                    help="ours: the product path; reference: the float64 CPU oracle (the driver's baseline arm); "
                         "tigre_style: the recalled TIGRE launch shapes on this GPU (csrc/tigre_style.cu), one subset")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the untimed C1/C4/C5 and multi-GPU parity legs")
    ap.add_argument("--shard", default="auto", choices=["auto", "grid", "angles", "rows"],
                    help="how a subset's rays are cut over the ranks (physics_arx_b200/dist.py)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = get_config(args.config)
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return
    if args.impl == "tigre_style":
        if rank == 0:
            run_tigre_style(args, cfg)
        return
    if args.warmup < 3:
        print("warning: fewer than 3 warm-up steps", file=sys.stderr)

    import torch
    import torch.distributed as dist
    from physics_arx_b200 import _lib
    from physics_arx_b200.phantom import artifact_volume, lung_phantom
    from physics_arx_b200.pipeline import SCBCTPipeline

    _lib.load()                      # fail loudly if the CUDA library is missing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        group = dist.group.WORLD
    if args.gpus != world and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    parity = None
    if world > 1 and not args.no_secondary:
        parity = parity_vs_single_gpu(group, rank, world)

    geo, angles = cfg["geo"], cfg["angles"]
    ct_np = lung_phantom(geo.nVoxel, geo.dVoxel, seed=0)
    art_np = artifact_volume(geo.nVoxel, seed=1)
    ct_host = torch.from_numpy(ct_np).pin_memory()
    art_host = torch.from_numpy(art_np).pin_memory()
    out_host = torch.empty(ct_host.shape, dtype=torch.float32).pin_memory()
    ct_dev, art_dev = ct_host.cuda(), art_host.cuda()
    pipe = SCBCTPipeline(geo, angles, niter=cfg["niter"], blocksize=cfg["blocksize"], I0=cfg["I0"],
                         sigma=cfg["sigma"], scatter_scale=SCATTER_SCALE, seed=0, group=group, shard=args.shard)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        pipe.run_device(ct_dev, art_dev)
    # exact trilinear-sample count of one full Ax pass over this rank's angles (untimed)
    ns = torch.zeros(1, dtype=torch.int64, device="cuda")
    if pipe.n_local:
        pipe.plan.ax(pipe.ct_res, 0, pipe.n_local, out=pipe.st.b, nsamples=ns)
    barrier()
    samples_per_pass = int(ns.item())

    # ---- timed region: K steps, inputs resident in HBM
    sampler = ClockSampler(local_rank) if rank == 0 else None
    pipe.launches = 0
    pipe.st.kernel_events = []
    pipe.st.phase_events = {}
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        pipe.run_device(ct_dev, art_dev)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = pipe.launches
    events = pipe.st.kernel_events
    pipe.st.kernel_events = None
    phases = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in pipe.st.phase_events.items()}
    ax_ms = [a.elapsed_time(b) for a, b, _ in events]
    ax_counts = [c for _, _, c in events]
    clocks = sampler.stop() if sampler else None

    # ---- e2e: the public host-buffer call, copies inside the timed region
    e2e_ms = None
    if not args.no_e2e:
        pipe.run_host(ct_host, art_host, out_host)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            pipe.run_host(ct_host, art_host, out_host)
        f1.record()
        barrier()
        e2e_ms = f0.elapsed_time(f1)

    t = torch.tensor([ms, e2e_ms or 0.0, float(samples_per_pass), float(sum(ax_ms))], dtype=torch.float64,
                     device="cuda")
    tmax = t.clone()
    tsum = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ln = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(ln)
        launches = int(ln.item())
    ms = float(tmax[0].item())
    e2e_ms = float(tmax[1].item()) if e2e_ms is not None else None
    total_samples_per_pass = float(tsum[2].item())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        nvox = int(np.prod(geo.nVoxel))
        npix = int(np.prod(geo.nDetector))
        # dominant kernel: k_ax in residual mode. algorithmic bytes per launch:
        # read the volume once, read b and write W*(b-Ax) for the launch's angles.
        avg_count = float(np.mean(ax_counts)) if ax_counts else 0.0
        avg_ms = float(np.mean(ax_ms)) if ax_ms else float("nan")
        alg_bytes = 4.0 * nvox + 8.0 * avg_count * npix
        achieved = alg_bytes / (avg_ms * 1e-3) / 1e9 if ax_ms else 0.0
        traffic, wavefronts = None, None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            traffic = prof.get(f"{cfg['name']}_ax_residual_bytes_per_launch_{pipe.plan.projector}")
            wavefronts = prof.get(f"{cfg['name']}_ax_residual_l1_wavefronts_per_launch_{pipe.plan.projector}")
        except Exception:
            pass
        # the honest bound of this kernel: the L1 data pipe (shared-memory tables + global-load wavefronts), one
        # wavefront per clock and SM; wavefronts per launch from the committed ncu capture (full 20-angle launch)
        sm_clock = float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)) * 1e6
        l1_peak = 148.0 * sm_clock
        l1 = None
        if wavefronts and ax_ms and world == 1:
            full = [m for m, c in zip(ax_ms, ax_counts) if c == cfg["blocksize"]] or ax_ms
            ach = wavefronts / (float(np.mean(full)) * 1e-3)
            l1 = {"bound": "l1_data_pipe", "achieved": ach / 1e9, "peak": l1_peak / 1e9, "unit": "Gwavefronts/s",
                  "frac": ach / l1_peak, "wavefronts_per_launch": wavefronts,
                  "peak_source": "1 wavefront per clock and SM x 148 SMs x the SM clock sampled under load",
                  "source": "profiles/r2_ncu_fan_queued.json (l1tex__data_pipe_lsu_wavefronts)"}
        ax_share = float(sum(ax_ms)) / float(t[0].item()) if ax_ms else None
        samples_resid = samples_per_pass * cfg["niter"] * args.steps     # rank 0's residual passes
        line = {
            "metric": METRIC, "value": args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": describe(cfg, world),
            "sharding": (None if world == 1 else {
                "grid": list(pipe.st.grid), "mode": pipe.st.shard,
                "meaning": "ranks as (angle groups) x (detector bands): the angles of every subset interleaved over "
                           "the groups, each group splitting the detector into bands",
                "scheme": pipe.st.scheme, "collective": collective_name(pipe.st)}),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "kernel": ("k_ax_fan<8,4,4,RESIDUAL> (event-driven fan-plane projector, csrc/ax_fan.cu)"
                                    if pipe.plan.projector == "fan" else
                                    "k_ax<1,RESIDUAL,32> (ray-march projector, csrc/ax.cu)"),
                         "peak_source": peak_src, "avg_launch_ms": avg_ms,
                         "alg_bytes_per_launch": alg_bytes, "share_of_step": ax_share,
                         "note": "a gather plus prefix sums bound by per-warp latency and the L1 data pipe, not by "
                                 "HBM: see roofline_l1, ax_gsamples_per_s and profiles/"},
            "roofline_l1": l1,
            "gpu_launches": launches,
            "subset_step_ms": dict(ax=float(np.mean(ax_ms)) if ax_ms else None, **phases,
                                   total=(ms / args.steps) / max(cfg["niter"] * len(pipe.st.blocks), 1),
                                   note="rank 0, CUDA events around each phase of an OS-SART subset step; total = "
                                        "whole step / subset steps (includes the simulation pass and all gaps)"),
            "clocks": clocks,
            "ax_gsamples_per_s": (samples_resid / (sum(ax_ms) * 1e-3) / 1e9) if ax_ms else None,
            "ax_samples_per_full_pass": total_samples_per_pass,
        }
        if e2e_ms is not None:
            line["e2e"] = {"value": args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                           "h2d_bytes_per_step": int(ct_host.numel() * 4 + art_host.numel() * 4),
                           "d2h_bytes_per_step": int(out_host.numel() * 4),
                           "api": "SCBCTPipeline.run_host (pinned host volumes in, host volume out)"}
        if parity is not None:
            line["parity_vs_1gpu"] = parity
        if world == 1 and not args.no_secondary:
            try:
                line["secondary"] = secondary_configs()
            except Exception as e:                      # noqa: BLE001 -- secondary legs never sink the headline
                line["secondary"] = {"error": repr(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_sample(cfg)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
