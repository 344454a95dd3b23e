#!/usr/bin/env python3
"""bench.py -- benchmark of the ray-traced acquisition / path-tracing hot path (BASELINE.json metric: Mrays/s & Msamples/s).

  python bench.py --gpus N --steps K --warmup W [--workload W] [--also a,b,...|none] [--impl reference]

ONE JSON line on stdout.  Headline workload (`config.workload`): BASELINE.json config 2 on MitsubaScenes/Sphere_Box.xml with
the transducer looking AT the sphere ("intended" transform order, the composition the reference's own dict scene uses,
USMain.py:69-71): 5 angles x 64 elements x 209 716 samples = 67 109 120 paths per step and GPU, multi-bounce, non-zero
deposits.  (Under Mitsuba's left-multiplying XML rule the array sits INSIDE the sphere, every connection is occluded and the
channel buffer is identically zero -- SURVEY.md Q1; that degenerate case is kept as the `sphere_box` entry of `also`.)

One "step" = one complete acquisition (weak scaling: rank g traces samples g, g+N, ... of N x 209 716, then ONE NCCL sum
all-reduce of the 12.8 MB channel buffer, issued per steering-angle slice so that it overlaps the next angle's kernel).
  value     Mrays/s, everything resident in HBM, CUDA events on the launching stream, max over ranks
  e2e       the same metric through the reference-facing plugin call UltraIntegrator.simulate_acquisition_parallel(scene)
            returning a HOST numpy buffer (zero-fill, kernels, all-reduce, D2H inside the timed region)
  roofline  of the dominant kernel, timed live inside the library with CUDA events (prt_profile_begin/read).  Scenes that
            live on chip are bound by instruction issue: achieved = warp instructions/s (per-ray count from the committed
            ncu capture x this run's rays/s), peak = 4 schedulers x 148 SMs x the SM clock sampled during the run; the
            DRAM-traffic view is reported beside it (`hbm`).  Only the 10 M-triangle scene is HBM-bound (`bound: "hbm"`).
  also      the other BASELINE configs, measured briefly in the same job at the same N (compact entries, last in the line):
            sphere_box (Mitsuba rule, degenerate), ring (config 3), cbox (config 4), heightfield (config 5); at N = 1 also
            us_render: the reference driver's acquisition + beamforming + envelope + log compression (USMain.py:92-224) as one
            library call at the driver's own sizes, in delay-and-sum terms per second
  cpu_baseline / --impl reference: the C restatement of the reference path (oracle/, kind "port": the real reference needs
            mitsuba/drjit, not installable here) on all host threads.  --impl reference traces the SAME job per step as the
            GPU arm (same `config`); cpu_baseline inside the GPU arm is a bounded ~10 s sample of it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

C2_SPP = 209716          # ceil(512*512*256 / 320): BASELINE.json config 2 as (angle, element, sample) paths
C1_SPP = 3277            # ceil(256*256*16 / 320):  BASELINE.json config 1 (the reference's CPU-runnable case)
DEFAULT_WORKLOAD = "sphere_box:intended"
DEFAULT_ALSO = "sphere_box,ring,cbox,heightfield"
N_SM, SCHEDULERS = 148, 4


def is_pt(name: str) -> bool:
    return name == "cbox" or name.startswith("heightfield")


def workload_desc(name: str):
    from prt_b200 import scenes
    if name == "ring":
        return scenes.test_ring_scene(), "TestRing mesh (1152 tris, GPU LBVH) in the Sphere_Floating acquisition block"
    table = {"sphere_box": "Sphere_Box", "sphere_floating": "Sphere_Floating", "cone_box": "Cone_Box",
             "cone_floating": "Cone_FLoating", "plate_box": "Plate_Box", "plane_floating": "Plane_Floating"}
    base, _, order = name.partition(":")
    if base not in table:
        raise SystemExit(f"unknown workload {name!r}")
    order = order or "mitsuba"
    return scenes.ultrasound_scene(table[base], order), f"MitsubaScenes/{table[base]}.xml ({order} transform order)"


def acq_config(label, p, spp, world, n_tris, n_analytic):
    """The `config` object of an acquisition workload -- built identically by the GPU arm and by --impl reference."""
    return {"workload": label, "paths_per_gpu_per_step": int(p.n_angles * p.n_elements * spp), "spp_per_gpu": int(spp),
            "n_angles": int(p.n_angles), "n_elements": int(p.n_elements), "time_samples": int(p.time_samples),
            "max_depth": int(p.max_depth), "n_triangles": int(n_tris), "n_analytic": int(n_analytic),
            "parallelism": f"sample-shards x{world}, scene replicated, one sum all-reduce of the channel buffer per step",
            "l2": "flushed between timed steps (384 MiB fill, untimed); inputs are < 200 KB and live on chip by design"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thr = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thr = threading.Thread(target=self._read, daemon=True)
        self.thr.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def bvh_min_bytes(n_tris: int) -> int:
    """B_min(N) = ceil(log2(N/4))*64 + 4*48 + 48 bytes per ray (0 for analytic-only scenes)."""
    if n_tris <= 0:
        return 0
    return int(np.ceil(np.log2(max(n_tris / 4.0, 1.0)))) * 64 + 4 * 48 + 48


def ncu_entry(key: str) -> dict:
    """What the committed `ncu --set full` capture of this workload's dominant kernel says (profiles/traffic.json):
    DRAM bytes per launch, warp instructions per ray, issue-slot utilisation.  {} if not captured."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return {}
    with open(path) as fh:
        d = json.load(fh)
    return d.get(key) or d.get(key.partition(":")[0], {}) if not key.startswith("heightfield") else d.get("heightfield", {})


def roofline_block(workload, kernel, launch_ms, rays_per_launch, alg_bytes, clk, on_chip: bool):
    """roofline for the dominant kernel.  on_chip: bound by instruction issue (frac never exceeds 1: achieved is an
    instruction rate against the issue peak); else HBM.  The DRAM-traffic view sits beside it in either case."""
    peak_hbm, peak_src, sm_max = measured_peaks()
    e = ncu_entry(workload)
    traffic = e.get("dram_bytes_per_launch") if e.get("kernel", kernel) == kernel else None
    sec = launch_ms * 1e-3
    hbm = {"algorithmic_gbs": alg_bytes / sec / 1e9, "peak_gbs": peak_hbm, "peak_source": peak_src}
    if traffic is not None:
        # traffic of the captured launch scaled to this run's launch by the ray count
        scale = rays_per_launch / e["rays_per_launch"] if e.get("rays_per_launch") else 1.0
        hbm["dram_gbs"] = traffic * scale / sec / 1e9
        hbm["dram_frac"] = hbm["dram_gbs"] / peak_hbm
        if e.get("l2_hit_pct") is not None:       # of the captured launch: BVH node / triangle fetches dominate its sectors
            hbm["l2_hit_pct"], hbm["l1_hit_pct"] = e["l2_hit_pct"], e.get("l1_hit_pct")
    if on_chip:
        mhz = (clk or {}).get("sm_mhz") or sm_max
        peak = N_SM * SCHEDULERS * mhz * 1e6 / 1e9
        wipr = e.get("warp_inst_per_ray") if e.get("kernel", kernel) == kernel else None
        ach = rays_per_launch * wipr / sec / 1e9 if wipr else None
        frac = ach / peak if ach else None
        rf = {"bound": "issue", "achieved": ach, "peak": peak, "unit": "Gwarp-inst/s", "frac": frac, "traffic": traffic,
              "warp_inst_per_ray": wipr, "ncu_issue_active_pct": e.get("issue_active_pct"),
              "ncu_active_lanes": e.get("active_threads_per_warp"), "source": e.get("source"), "hbm": hbm}
        if frac is not None and frac > 1.0:      # a stale per-ray instruction count: not evidence of anything
            rf.update(frac=None, achieved=None, stale=True)
        return rf
    ach = alg_bytes / sec / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak_hbm, "unit": "GB/s", "frac": ach / peak_hbm, "traffic": traffic,
            "algorithmic_bytes_per_launch": alg_bytes, "ncu_issue_active_pct": e.get("issue_active_pct"),
            "ncu_active_lanes": e.get("active_threads_per_warp"), "source": e.get("source"), "hbm": hbm}


def cpu_oracle_rate(desc, params, spp: int, seed: int, threads: int):
    import orc_py
    sc = orc_py.OracleScene(desc)
    t0 = time.perf_counter()
    _, _, st = sc.acquire(params, seed=seed, spp=spp, prec=32, n_threads=threads)
    dt = time.perf_counter() - t0
    return st, dt


def cpu_acq_baseline(desc, p, seconds: float, extra_legs: bool):
    """Bounded sample of the same workload on all host threads (+ the single-thread / pure-Python legs of SURVEY 8(d))."""
    threads = os.cpu_count() or 1
    n_ae = p.n_angles * p.n_elements
    st, dt = cpu_oracle_rate(desc, p, C1_SPP, 0, threads)
    reps = int(min(max(seconds / max(dt, 1e-3), 1), 2000))
    rays_c, t_c = st["rays"], dt
    for k in range(1, reps):
        st, dt = cpu_oracle_rate(desc, p, C1_SPP, k, threads)
        rays_c += st["rays"]; t_c += dt
    out = {"value": rays_c / t_c / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
           "sample": f"{reps} x {C1_SPP * n_ae} paths of the same scene (BASELINE config 1 path count), oracle f32, {threads} threads, {t_c:.1f} s"}
    if extra_legs:
        st1, dt1 = cpu_oracle_rate(desc, p, C1_SPP, 0, 1)
        out["single_thread"] = {"value": st1["rays"] / dt1 / 1e6, "unit": "Mrays/s", "cores": 1}
        try:
            import pyref
            shapes_py = pyref.shapes_from_desc(desc)
            t0 = time.perf_counter()
            _, _, stp = pyref.acquire(shapes_py, p, seed=0, spp=1)
            dtp = time.perf_counter() - t0
            out["pure_python"] = {"value": stp["rays"] / dtp / 1e6, "unit": "Mrays/s", "cores": 1,
                                  "sample": f"{n_ae} paths (the reference's literal acquisition), oracle/pyref.py"}
        except ValueError:
            pass                         # pyref handles sphere / rectangle scenes only
    return out


# ----------------------------------------------------------------------------------------------------------------------
# acquisition workloads (BASELINE configs 2 and 3)
# ----------------------------------------------------------------------------------------------------------------------
def acq_measure(args, workload, steps, warmup, rank, world, device, e2e_steps, cpu_seconds, extra_legs=False):
    import torch
    import torch.distributed as dist
    from prt_b200 import mi_compat as mi
    from prt_b200.distributed import acquire_allreduce_pipelined, shard_samples

    desc, label = workload_desc(workload)
    scene = mi.Scene(desc)
    integ = scene.integrator()
    dev = scene.device()
    p = integ.acq_params(scene)
    n_ae = p.n_angles * p.n_elements
    spp_total = args.spp * world                       # weak scaling: per-GPU work fixed
    off, stride, n_s = shard_samples(spp_total, rank, world)
    buf = torch.zeros((p.n_angles, p.n_elements, p.time_samples), dtype=torch.float32, device=device)
    tx = torch.zeros((p.n_angles, p.n_elements), dtype=torch.float32, device=device)
    stats = torch.zeros(8, dtype=torch.int64, device=device)
    flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=device)   # > 126 MB L2
    stream = torch.cuda.current_stream(device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for w in range(warmup):
        buf.zero_()
        acquire_allreduce_pipelined(dev, p, buf, tx, stats, stream, 1000 + w, spp_total, off, stride, dist, world)
    barrier()
    stats.zero_()
    clocks = ClockSampler(device.index)
    clocks.start()
    time.sleep(0.4)                                 # let nvidia-smi start sampling before the timed region
    ev = []
    barrier()
    dev.ctx.profile_begin()                         # event pair around every k_acquire launch, on the launching stream
    for k in range(steps):
        flush.fill_(k & 0xff)                       # L2 flush between timed iterations (untimed)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(stream)
        buf.zero_()
        acquire_allreduce_pipelined(dev, p, buf, tx, stats, stream, k, spp_total, off, stride, dist, world)
        e[1].record(stream)
        ev.append(e)
    barrier()
    clk = clocks.stop()
    classes = dev.ctx.profile_read()
    t_local = torch.tensor([sum(e[0].elapsed_time(e[1]) for e in ev)], dtype=torch.float64, device=device)
    checksum_dev = float(buf.abs().sum().item())
    st_local = stats.cpu().numpy().copy()
    st_all = stats.clone()
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
        dist.all_reduce(st_all, op=dist.ReduceOp.SUM)
    total_ms = float(t_local.item())
    hs = st_all.cpu().numpy()
    paths, segments, rays, deposits = int(hs[0]), int(hs[1]), int(hs[2]), int(hs[3])
    value = rays / (total_ms * 1e-3) / 1e6

    # ---- end to end through the reference-facing plugin call, HOST results -----------------------------------
    integ.samples_per_element = spp_total
    integ.seed = 999
    so = sys.stdout
    sys.stdout = open(os.devnull, "w")               # the reference's method prints; keep ONE JSON line on stdout
    try:
        integ.simulate_acquisition_parallel(scene)  # warm-up (allocates the pinned result buffers: the pool needs
        integ.simulate_acquisition_parallel(scene)  # two, because the integrator still holds the previous result)
        barrier()
        e2e_rays = 0
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            integ.seed = 2000 + k
            integ.simulate_acquisition_parallel(scene)
            e2e_rays += int(integ.last_stats["rays"])
            host_checksum = float(np.abs(integ.channel_buf.ravel()[::997]).sum())   # touch the HOST result
        barrier()
        e2e_dt = time.perf_counter() - t0
    finally:
        sys.stdout.close()
        sys.stdout = so
    t_e2e = torch.tensor([e2e_dt], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)   # acquire_sharded reduces the stats: every rank holds the global ray count
    e2e_value = e2e_rays / float(t_e2e.item()) / 1e6
    del flush
    if rank != 0:
        return None
    n_tris = desc.n_triangles()
    acq = classes.get("acquire") or {"ms": total_ms, "launches": steps}
    launch_ms = acq["ms"] / max(acq["launches"], 1)
    rays_per_launch = int(st_local[2]) / max(acq["launches"], 1)
    # algorithmic bytes per launch: BVH descent per ray (0 for analytic scenes, which live in shared
    # memory) + this launch's slice of the channel buffer written once (no per-segment state is streamed)
    alg_bytes = rays_per_launch * bvh_min_bytes(n_tris) + buf.numel() * 4 * steps / max(acq["launches"], 1)
    kernel = "prt::k_acquire<%s>" % ("true" if n_tris else "false")
    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "msamples_per_s": paths / (total_ms * 1e-3) / 1e6, "msegments_per_s": segments / (total_ms * 1e-3) / 1e6,
        "config": acq_config(label, p, args.spp, world, n_tris, desc.n_analytic()),
        "segments_per_path": segments / max(paths, 1), "rays_per_path": rays / max(paths, 1),
        "deposits_per_path": deposits / max(paths, 1), "device_checksum": checksum_dev,
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(8 * p.n_angles + 512),
                "d2h_bytes_per_step": int(buf.numel() * 4 + tx.numel() * 4 + 64), "steps": e2e_steps,
                "api": "UltraIntegrator.simulate_acquisition_parallel(scene) -> numpy channel_buf", "host_checksum": host_checksum},
        "gpu_launches": acq["launches"], "kernel": kernel, "kernel_ms": launch_ms,
        "roofline": roofline_block(workload, kernel, launch_ms, rays_per_launch, alg_bytes, clk, on_chip=True),
        "clocks": clk,
    }
    if cpu_seconds > 0:
        line["cpu_baseline"] = cpu_acq_baseline(desc, p, cpu_seconds, extra_legs)
    return line


# ----------------------------------------------------------------------------------------------------------------------
# light-transport workloads (BASELINE configs 4 and 5)
# ----------------------------------------------------------------------------------------------------------------------
def pt_scene(args, workload, spp_step=None):
    from prt_b200 import scenes
    if workload.startswith("heightfield"):
        # BASELINE config 5: 9 999 392-triangle height field in a closed box, 3840x2160, 8 diffuse bounces, no RR
        n_side = int(workload.partition(":")[2] or 2237)
        width, height = (3840, 2160) if args.res == 2048 else (args.res, args.res * 9 // 16)
        spp_step = spp_step or (args.spp if args.spp != C2_SPP else 2)
        desc = scenes.heightfield_scene(n_side, (width, height), spp_step)
        label = (f"synthetic height field, {desc.n_triangles()} triangles in a closed box, {width}x{height}, all diffuse, "
                 "8 bounces (max_depth 9), no RR")
    else:
        width = height = args.res
        spp_step = spp_step or (args.spp if args.spp != C2_SPP else 16)
        desc = scenes.cbox_scene(args.res, spp_step)
        label = f"scenes/cbox.xml at {width}x{height}, path integrator max_depth 6 rr_depth 5, tent filter"
    return desc, label, width, height, spp_step


def pt_measure(args, workload, steps, warmup, rank, world, device, e2e_steps, cpu_seconds):
    """One step = one batch of `spp` samples per pixel per GPU accumulated into the film; the K timed steps are ONE job:
    film zeroed once, K batches, then ONE all-reduce of the RGBW film (BASELINE config 4: a 4096-spp job is 256 such
    batches and one all-reduce)."""
    import torch
    import torch.distributed as dist
    from prt_b200 import mi_compat as mi
    from prt_b200.distributed import shard_samples
    desc, label, width, height, spp_step = pt_scene(args, workload)
    scene = mi.Scene(desc)
    integ = scene.integrator()
    dev = scene.device()
    rp = integ.render_params(scene)
    spp_total = spp_step * world
    off, stride, n_s = shard_samples(spp_total, rank, world)
    film = torch.zeros((height, width, 4), dtype=torch.float32, device=device)
    stats = torch.zeros(8, dtype=torch.int64, device=device)
    flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=device)
    stream = torch.cuda.current_stream(device)
    bvh = dev.bvh_stats

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for w in range(warmup):
        dev.render_path_dev(rp, film.data_ptr(), stats.data_ptr(), stream.cuda_stream, seed=1000 + w, spp=spp_total,
                            sample_offset=off, sample_stride=stride)
    if world > 1:
        dist.all_reduce(film, op=dist.ReduceOp.SUM)
    barrier()
    stats.zero_()
    clocks = ClockSampler(device.index)
    clocks.start()
    time.sleep(0.4)
    ev = []
    barrier()
    dev.ctx.profile_begin()                 # event pairs around every kernel group, on the launching stream
    e_first = torch.cuda.Event(enable_timing=True)
    e_first.record(stream)
    film.zero_()
    e_zero = torch.cuda.Event(enable_timing=True)
    e_zero.record(stream)
    for k in range(steps):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        if k:
            flush.fill_(k & 0xff)           # L2 flush between timed batches; its time is taken out below
        e[0].record(stream)
        dev.render_path_dev(rp, film.data_ptr(), stats.data_ptr(), stream.cuda_stream, seed=k, spp=spp_total, sample_offset=off,
                            sample_stride=stride)
        e[1].record(stream)
        ev.append(e)
    e_red = torch.cuda.Event(enable_timing=True)
    e_red.record(stream)
    if world > 1:
        dist.all_reduce(film, op=dist.ReduceOp.SUM)
    e_last = torch.cuda.Event(enable_timing=True)
    e_last.record(stream)
    barrier()
    clk = clocks.stop()
    classes = dev.ctx.profile_read()
    job_ms = e_first.elapsed_time(e_zero) + sum(e[0].elapsed_time(e[1]) for e in ev) + e_red.elapsed_time(e_last)
    t_local = torch.tensor([job_ms], dtype=torch.float64, device=device)
    st_local = stats.cpu().numpy().copy()
    st_all = stats.clone()
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
        dist.all_reduce(st_all, op=dist.ReduceOp.SUM)
    total_ms = float(t_local.item())
    hs = st_all.cpu().numpy()
    paths, segments, rays, shadow = (int(x) for x in hs[:4])
    value = rays / (total_ms * 1e-3) / 1e6
    del flush
    # end to end: mi.render(scene) -> developed numpy image on the host (page-locked), every call a complete job
    keep = [integ.render(scene, seed=77, spp=spp_total), integ.render(scene, seed=78, spp=spp_total)]   # result pool warm-up
    del keep
    barrier()
    e2e_rays = 0
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        img = integ.render(scene, seed=2000 + k, spp=spp_total)
        e2e_rays += int(integ.last_stats["rays"])
        checksum = float(img[::8, ::8].mean())         # touch the HOST result
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    if rank != 0:
        return None
    n_tris = desc.n_triangles()
    n_launch = sum(c["launches"] for c in classes.values())
    step_kernels = sum(c["ms"] for c in classes.values())
    tc = classes.get("trace_closest", {"ms": 0.0, "launches": 0})
    mk = classes.get("megakernel", {"ms": 0.0, "launches": 0})
    if mk["launches"]:                                 # resident-scene kernel (scenes that fit shared memory) or PRT_PT_MODE=mega
        dom = "prt::k_render_resident" if n_tris <= 64 and os.environ.get("PRT_PT_MODE", "r")[0] != "m" else "prt::k_render_path"
        dom_ms = mk["ms"] / mk["launches"]
        rays_launch = int(st_local[2]) / mk["launches"]
        alg_bytes = film.numel() * 4 * 2               # the film tile read-modify-write is all that leaves the SM
        share = mk["ms"] / max(step_kernels, 1e-9)
    else:
        dom, dom_ms = "prt::k_wf_trace<false>", tc["ms"] / max(tc["launches"], 1)
        rays_launch = (int(st_local[2]) - int(st_local[3])) / max(tc["launches"], 1)
        alg_bytes = rays_launch * (bvh_min_bytes(n_tris) + WF_RAY_RECORD_BYTES)
        share = tc["ms"] / max(step_kernels, 1e-9)
    line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "msamples_per_s": paths / (total_ms * 1e-3) / 1e6,
            "config": {"workload": label, "spp_per_gpu_per_step": spp_step, "paths_per_gpu_per_step": width * height * n_s,
                       "n_triangles": n_tris, "n_analytic": desc.n_analytic(),
                       "parallelism": f"sample-shards x{world}, BVH replicated, ONE NCCL all-reduce of the {film.numel() * 4} B film per job",
                       "l2": "flushed between timed batches (384 MiB fill, untimed)"},
            "bvh": {"nodes": bvh["n_nodes"], "nodes8": bvh.get("n_nodes8"), "build_ms": bvh["build_ms"],
                    "bvh8_build_ms": bvh.get("bvh8_build_ms"), "sah_cost": bvh["sah_cost"], "device_bytes": bvh["device_bytes"]},
            "segments_per_path": segments / max(paths, 1), "rays_per_path": rays / max(paths, 1),
            "e2e": {"value": e2e_rays / float(t_e2e.item()) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 512,
                    "d2h_bytes_per_step": int(height * width * 3 * 4 + 64), "steps": e2e_steps,
                    "api": "mi.render(scene) -> developed numpy image [H,W,3]", "host_checksum": checksum},
            "gpu_launches": n_launch, "kernel": dom, "kernel_ms": dom_ms, "kernel_share_of_step": share,
            "kernel_classes": {k: {"ms_per_step": v["ms"] / steps, "launches_per_step": v["launches"] / steps}
                               for k, v in classes.items() if v["launches"]},
            "roofline": roofline_block(workload, dom, dom_ms, rays_launch, alg_bytes, clk, on_chip=n_tris < 100000),
            "clocks": clk}
    if cpu_seconds > 0:
        line["cpu_baseline"] = cpu_pt_baseline(args, workload)
    return line


WF_RAY_RECORD_BYTES = 64 + 16      # ray record read (4 x float4) + hit record written, per closest-hit query


def cpu_pt_baseline(args, workload):
    import orc_py
    from prt_b200 import scenes
    import CustomIntegrator
    import CustomSensor
    threads = os.cpu_count() or 1
    if workload.startswith("heightfield"):
        n_side, res, spp = 708, (256, 144), 4      # the oracle's BVH build over 10 M triangles takes ~40 s: a 1.0 M-triangle
        cdesc = scenes.heightfield_scene(n_side, res, spp)   # instance of the same surface bounds the rate from above
        what = f"{cdesc.n_triangles()}-triangle instance of the same height field, {res[0]}x{res[1]} x {spp} spp"
    else:
        res, spp = (256, 256), 16                   # the cbox tutorial resolution, 16 spp: ~1 M paths
        cdesc = scenes.cbox_scene(res[0], spp)
        what = f"{res[0]}x{res[1]} x {spp} spp of the same scene"
    integ = CustomIntegrator.PathIntegrator(cdesc.integrator)
    sensor = CustomSensor.PerspectiveSensor(cdesc.sensor)
    sensor._film, sensor._sampler, sensor._rfilter = cdesc.film, cdesc.sampler, cdesc.rfilter
    crp = integ.render_params(type("S", (), {"sensors": lambda self: [sensor]})())
    osc = orc_py.OracleScene(cdesc)
    t0 = time.perf_counter()
    _, cst = orc_py.render_path(osc, crp, seed=0, spp=spp, prec=32, n_threads=threads)
    dt = time.perf_counter() - t0
    return {"value": cst["rays"] / dt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
            "sample": f"{what}, oracle f32, {threads} threads, {dt:.1f} s"}


def usrender_entry(cpu_seconds: float):
    """SURVEY 8(f) rank 1, the driver's us_render() (USMain.py:92-224) at ITS sizes -- 5 angles x 64 elements x 10 000 samples
    onto 1040 x 638 pixels -- as one library call (UltraIntegrator.render_bmode -> prt_us_render: acquisition, delay-and-sum,
    envelope, log compression; only the display image leaves the device).  value = delay-and-sum terms (pixel x angle x
    element) per second of DEVICE time of the whole call; cpu = the numpy restatement of the beamformer (oracle/pyref.py) on a
    bounded column sample, one thread."""
    from prt_b200 import mi_compat as mi, scenes
    d = scenes.usmain_scene_dict()
    scene = mi.load_dict(d)
    integ = scene.integrator()
    lam = integ.sound_speed / integ.frequency
    x = np.arange(-0.04, 0.04 + lam / 4, lam / 4)
    z = np.arange(0.001, 0.05 + lam / 4, lam / 4)
    terms = x.size * z.size * integ.n_angles * integ.n_elements
    dev_ms, wall_ms = [], []
    img = None
    for it in range(13):
        t0 = time.perf_counter()
        img = integ.render_bmode(scene, x, z, dynamic_range=60.0)
        wall = (time.perf_counter() - t0) * 1e3
        if it >= 3:
            dev_ms.append(integ.last_stats["kernel_ms"])
            wall_ms.append(wall)
    ms = float(np.median(dev_ms))
    r = lambda v, n=4: float(f"{v:.{n}g}")
    out = {"value": r(terms / (ms * 1e-3) / 1e9), "unit": "Gterms/s", "ms": r(ms), "wall_ms": r(float(np.median(wall_ms))),
           "px": [int(x.size), int(z.size)], "ck": r(float(np.asarray(img, dtype=np.float64).mean()), 3), "n": 1,
           "kernel": "k_acquire+k_das+k_envelope+k_bmode"}
    if cpu_seconds > 0:
        import pyref
        integ.simulate_acquisition_parallel(scene)
        ch = np.asarray(integ.channel_buf, dtype=np.float64)
        cols = 16
        t0 = time.perf_counter()
        pyref.das_beamform(ch, integ.angles.numpy(), x[:cols], z, integ.fs, integ.sound_speed, integ.pitch, 0.0, 1.0)
        dt = time.perf_counter() - t0
        out["cpu"] = r(cols * z.size * integ.n_angles * integ.n_elements / dt / 1e9)
        out["cpu_cores"] = 1
    return out


def compact(line):
    """An `also` entry: what the driver's 1 500-character tail has room for."""
    r = lambda x, n=4: None if x is None else float(f"{x:.{n}g}")
    rf = line["roofline"]
    out = {"value": r(line["value"], 5), "msamples": r(line["msamples_per_s"], 5), "e2e": r(line["e2e"]["value"], 5),
           "ms": r(line["ms_per_step"]), "n": line["n_gpus"], "ck": r(line["e2e"]["host_checksum"], 3),
           "kernel": line["kernel"].replace("prt::", ""), "bound": rf["bound"], "frac": r(rf.get("frac"), 3),
           "dram_frac": r(rf["hbm"].get("dram_frac"), 3)}
    if rf["bound"] == "hbm" and rf["hbm"].get("l2_hit_pct") is not None:
        out["l2_hit"] = rf["hbm"]["l2_hit_pct"]      # ncu, BVH-fetch kernel (north star: "L2 hit rate for BVH fetches")
    if "cpu_baseline" in line:
        out["cpu"] = r(line["cpu_baseline"]["value"])
        out["cpu_cores"] = line["cpu_baseline"]["cores"]
    return out


def run_reference(args, rank: int):
    """--impl reference: the CPU restatement of the reference path on all host threads, the SAME job per step as the GPU arm
    (one GPU's share: the weak-scaling unit) and the same `config`."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    if is_pt(args.workload):
        t0 = time.perf_counter()
        cb = cpu_pt_baseline(args, args.workload)
        line = {"impl": "reference", "metric": "Mrays/s", "value": cb["value"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": 1,
                "warmup": 0, "ms_per_step": 1e3 * (time.perf_counter() - t0), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": args.workload},
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return
    from prt_b200.scene import AcqParams
    desc, label = workload_desc(args.workload)
    p = AcqParams.from_props(desc.integrator, desc.sensor)
    spp = args.spp
    for _ in range(args.warmup):
        cpu_oracle_rate(desc, p, max(spp // 64, 1), 0, threads)      # short: page in the library, spin up the thread pool
    rays = paths = 0
    total = 0.0
    for k in range(args.steps):
        st, dt = cpu_oracle_rate(desc, p, spp, k, threads)
        rays += st["rays"]; paths += st["paths"]; total += dt
    val = rays / total / 1e6
    sample = (f"{spp * p.n_angles * p.n_elements} paths/step = the GPU arm's per-GPU job, oracle f32 (C restatement of "
              f"CustomIntegrator.simulate_acquisition_parallel), {threads} threads")
    line = {"impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "msamples_per_s": paths / total / 1e6,
            "config": acq_config(label, p, spp, args.gpus, desc.n_triangles(), desc.n_analytic()),
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--spp", type=int, default=C2_SPP, help="samples per (angle, element) per GPU and step")
    ap.add_argument("--res", type=int, default=2048, help="film resolution of the cbox workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--also", default=None, help=f"comma list of workloads measured briefly in the same job (default for the "
                                                 f"headline workload: {DEFAULT_ALSO}; 'none' to skip)")
    ap.add_argument("--no-also", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed end-to-end steps (default: min(steps, 5))")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    cpu_s = 0.0 if (world > 1 or args.no_cpu_baseline) else 10.0
    e2e_steps = args.e2e_steps or min(args.steps, 5)
    if is_pt(args.workload):
        line = pt_measure(args, args.workload, args.steps, args.warmup, rank, world, device, min(e2e_steps, 3), cpu_s)
    else:
        line = acq_measure(args, args.workload, args.steps, args.warmup, rank, world, device, e2e_steps, cpu_s, extra_legs=True)
    if rank == 0 and not is_pt(args.workload):
        # a headline that computes nothing is not a measurement (round-1 lesson): the default workload must deposit energy
        if args.workload == DEFAULT_WORKLOAD:
            assert line["deposits_per_path"] > 0 and line["device_checksum"] > 0 and line["e2e"]["host_checksum"] > 0, \
                "headline workload deposited nothing"
    also = args.also if args.also is not None else (DEFAULT_ALSO if args.workload == DEFAULT_WORKLOAD else "none")
    entries = {}
    if not args.no_also and also != "none":
        torch.cuda.empty_cache()
        cpu_a = 0.0 if cpu_s == 0 else 3.0
        for wl in [w for w in also.split(",") if w and w != args.workload]:
            try:
                if is_pt(wl):
                    sub = pt_measure(args, wl, 3, 3, rank, world, device, 2, cpu_a)
                else:
                    sub = acq_measure(args, wl, 5, 3, rank, world, device, 2, cpu_a)
                if rank == 0:
                    entries[wl] = compact(sub)
            except Exception as ex:      # an `also` failure must not take the headline line down with it
                if rank == 0:
                    entries[wl] = {"error": f"{type(ex).__name__}: {ex}"[:160]}
            torch.cuda.empty_cache()
    if rank == 0 and world == 1 and entries and not args.no_also:
        try:          # the reference driver's other hot loop (beamforming), one GPU: SURVEY 8(f) rank 1
            entries["us_render"] = usrender_entry(cpu_s)
        except Exception as ex:
            entries["us_render"] = {"error": f"{type(ex).__name__}: {ex}"[:160]}
    if rank == 0:
        line.pop("kernel_classes", None) if entries else None
        if entries:
            line["also"] = entries              # LAST key: survives the driver's stdout tail
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
