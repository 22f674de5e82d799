#!/usr/bin/env python
"""bench.py — headline benchmark of the render hot path (BASELINE.json metric: Mrays/s = ray segments / second / 1e6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one full render of the workload (default: BASELINE.json configs[1] = config:c2, two suzanne.obj instances,
Lambert + metal, 1920x1080, 64 spp, depth 50) through the hand-written CUDA path.  N > 1: weak scaling — every rank
renders 64 spp (its share of the 64*N sample lanes of ONE render() call) and the film is combined with one NCCL
all-reduce inside the timed region.

JSON line keys: see README of the task contract; additionally `roofline` (extend kernel, HBM), `fp32` (secondary,
issue-side roofline of the same kernel), `cpu_baseline` (the oracle's restatement of the reference's multithreaded CPU
render() on this box's host cores), `e2e`, `clocks`, `gpu_launches`.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {  # BASELINE.json configs (DESIGN.md "Workload definitions")
    "c1": dict(world="config:c1", width=1200, height=800, spp=10, lanes=10, cpu_sample=(300, 200, 4)),
    "c2": dict(world="config:c2", width=1920, height=1080, spp=64, lanes=64, cpu_sample=(640, 360, 4)),
    "c3": dict(world="config:c3", width=1024, height=1024, spp=256, lanes=32, cpu_sample=(256, 256, 16)),
    "c4": dict(world="config:c4", width=3840, height=2160, spp=16, lanes=1, cpu_sample=(240, 135, 2)),
    "c5": dict(world="config:c5", width=3840, height=2160, spp=4096, lanes=8, cpu_sample=(240, 135, 2)),
}
MAX_DEPTH = 50
EXTEND_BYTES_PER_RAY = 40       # extend kernel: reads origin+direction (2 x 16 B), writes (hit ref, t) 8 B  — DESIGN.md
NODE_FLOPS, PRIM_FLOPS = 60, 41  # per node visit (two child boxes, 30 each) / per primitive test


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured"
    return 6650.0, 1965.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_sample_dims(wl, cores):
    """Bounded CPU sample: every one of the `cores` worker threads gets >= 1 sample per pixel (the reference splits SAMPLES
    across threads), and the image shrinks so the work stays ~constant (C2: ~0.9 M paths, 10-15 s on 16 cores)."""
    w, h, spp = wl["cpu_sample"]
    target = w * h * spp
    spp = max(spp, cores)
    scale = (target / float(w * h * spp)) ** 0.5
    return max(48, int(w * scale)), max(27, int(h * scale)), spp


def run_reference(args, wl):
    """--impl reference: the reference's own CPU implementation of the path.  It is Rust and cannot be built in this image
    (no rustc / cargo), so this arm times the oracle's literal C++ restatement of render() (thread-per-sample-split,
    reference node test, recursive traversal) with all host threads, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as ob
    cores = os.cpu_count() or 1
    w, h, spp = cpu_sample_dims(wl, cores)
    W = ob.OracleWorld.named(wl["world"])
    times, rays = [], []
    for i in range(args.warmup + args.steps):
        _, st = W.render(w, h, spp, MAX_DEPTH, cores, 1000 + i, "faithful")
        if i >= args.warmup:
            times.append(st["seconds"]); rays.append(st["rays"])
    val = sum(rays) / sum(times) / 1e6
    sample = f"{wl['world']} at {w}x{h}, {spp} spp, depth {MAX_DEPTH} per step (Mrays/s is resolution- and spp-invariant)"
    out = {"impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": wl["world"], "width": wl["width"], "height": wl["height"], "spp_per_gpu": wl["spp"],
                      "spp_total": wl["spp"] * args.gpus, "max_depth": MAX_DEPTH, "sample_lanes_per_gpu": wl["lanes"],
                      "note": "reference arm = CPU restatement of the Rust renderer (rustc unavailable); bounded sample per step"},
           "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def cpu_baseline(wl, budget_s=20.0):
    from oracle import binding as ob
    cores = os.cpu_count() or 1
    w, h, spp = cpu_sample_dims(wl, cores)
    W = ob.OracleWorld.named(wl["world"])
    _, st = W.render(w, h, spp, MAX_DEPTH, cores, 4242, "faithful")
    return {"value": st["rays"] / st["seconds"] / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": f"{wl['world']} at {w}x{h}, {spp} spp, depth {MAX_DEPTH}: {st['rays']} ray segments in {st['seconds']:.2f} s "
                      f"(oracle = C++ restatement of the reference's render(); rustc unavailable)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (invalidates the headline number)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.config])
    if args.spp:
        wl["spp"] = args.spp
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist
    import raytracinginaweekend_b200 as rb
    from raytracinginaweekend_b200 import _cabi
    from raytracinginaweekend_b200 import dist as rdist

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n_gpus = world_size

    world = rb.World.named(wl["world"])
    size = rb.Size2i(wl["width"], wl["height"])
    spp_total, lanes_total = wl["spp"] * n_gpus, wl["lanes"] * n_gpus     # weak scaling: fixed work per GPU
    scene = world.device_scene(local_rank)                                  # inputs resident in HBM before the timed region
    stream = torch.cuda.current_stream(dev)
    plane = torch.empty((size.height, size.width, 3), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2
    mult = float(np.float32(1.0) / np.float32(rdist.n_tasks(lanes_total, spp_total)))
    _cabi.set_option("profile", 1)   # per-launch CUDA events on the launching stream (no extra synchronisation)

    def step(seed):
        st = scene.render_planes_device(plane.data_ptr(), rank, n_gpus, size.width, size.height, spp_total, MAX_DEPTH, lanes_total,
                                        seed, 0, stream.cuda_stream)
        if n_gpus > 1:
            dist.all_reduce(plane, op=dist.ReduceOp.SUM)
        _cabi.check(_cabi.lib().rt_film_scale_device(plane.data_ptr(), plane.numel(), mult, stream.cuda_stream))
        return st

    def barrier():
        torch.cuda.synchronize(dev)
        if n_gpus > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step(100 + i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    step_ms, stats = [], []
    for i in range(args.steps):
        flush.fill_(float(i))                      # evict the previous step's working set from L2 (outside the timed events)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        stats.append(step(200 + i))
        e1.record(stream)
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    per_rank = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(n_gpus)]
    if n_gpus > 1:
        dist.all_gather(per_rank, torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev))
    t_local = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    rays_local = torch.tensor([float(sum(s["ray_segments"] for s in stats))], dtype=torch.float64, device=dev)
    if n_gpus > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays_local, op=dist.ReduceOp.SUM)
    total_ms, total_rays = float(t_local.item()), float(rays_local.item())
    value = total_rays / (total_ms * 1e-3) / 1e6

    # ---- end-to-end through the public API with HOST buffers: scene hand-off (H2D) + render + film read-back (D2H), every step
    pinned = torch.empty((size.height, size.width, 3), dtype=torch.float32).pin_memory()
    barrier()
    e2e_ms, e2e_rays = 0.0, 0.0
    e2e_steps, e2e_phases = [], []
    for i in range(-min(3, max(1, args.warmup)), max(1, args.steps)):   # i < 0: untimed warm-up of this code path (like the W steps above)
        t0 = time.perf_counter()
        world.release_device()
        t1 = time.perf_counter()
        scene_e = world.device_scene(local_rank)      # rt_scene_create: pack + rebuild + upload
        t2 = time.perf_counter()
        film, st = rdist.render_distributed(world, size, lanes_total, spp_total, MAX_DEPTH, rb.RenderMode.Default, 300 + i,
                                            device=dev, return_stats=True)
        t3 = time.perf_counter()
        pinned.copy_(film, non_blocking=False)
        torch.cuda.synchronize(dev)
        t4 = time.perf_counter()
        e2e_phases.append([round((b - a) * 1e3, 2) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4))] + [round(st["render_ms"], 2)])
        if i < 0:
            continue
        e2e_steps.append((time.perf_counter() - t0) * 1e3)
        e2e_ms += e2e_steps[-1]
        e2e_rays += st["ray_segments"]
    scene = world.device_scene(local_rank)
    t_e = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    r_e = torch.tensor([e2e_rays], dtype=torch.float64, device=dev)
    if n_gpus > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        dist.all_reduce(r_e, op=dist.ReduceOp.SUM)
    e2e_value = float(r_e.item()) / (float(t_e.item()) * 1e-3) / 1e6

    if rank == 0:
        hbm_peak, sm_mhz, peak_kind = peaks()
        ext_ms = sum(s["extend_ms"] for s in stats)
        n_ext = sum(s["iterations"] for s in stats)
        rays_rank0 = sum(s["ray_segments"] for s in stats)
        achieved = rays_rank0 * EXTEND_BYTES_PER_RAY / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
        # one extra (untimed) step with a single wavefront pipeline: kernels then run back to back, so per-launch CUDA-event times
        # are comparable with the (serialised) ncu launch list and extend's share of the step is well defined
        _cabi.set_option("pipes", 1)
        pst = scene.render_planes_device(plane.data_ptr(), rank, n_gpus, size.width, size.height, spp_total, MAX_DEPTH, lanes_total,
                                         200, 0, stream.cuda_stream)
        _cabi.set_option("pipes", 2)
        torch.cuda.synchronize(dev)
        serial = {"pipes": 1, "step_ms": pst["render_ms"], "extend_ms": pst["extend_ms"], "shade_ms": pst["shade_ms"],
                  "avg_launch_ms": pst["extend_ms"] / max(1, pst["iterations"]), "share_of_step": pst["extend_ms"] / max(1e-9, pst["render_ms"]),
                  "achieved_gbs": pst["ray_segments"] * EXTEND_BYTES_PER_RAY / (pst["extend_ms"] * 1e-3) / 1e9}
        # one extra (untimed) step with the counting kernel variant: node visits / primitive tests per ray -> FP32 issue roofline
        _cabi.set_option("count", 1)
        cst = scene.render_planes_device(plane.data_ptr(), rank, n_gpus, size.width, size.height, spp_total, MAX_DEPTH, lanes_total,
                                         200, 0, stream.cuda_stream)
        _cabi.set_option("count", 0)
        torch.cuda.synchronize(dev)
        nodes_per_ray = cst["node_visits"] / max(1, cst["ray_segments"])
        prims_per_ray = cst["prim_tests"] / max(1, cst["ray_segments"])
        flops_per_ray = NODE_FLOPS * nodes_per_ray + PRIM_FLOPS * prims_per_ray
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        fp32_peak = sms * 128 * sm_mhz * 1e6 / 1e12      # TFLOP/s without FMA (parity code is --fmad=false)
        # algorithmic flops of one step / the extend time of the single-pipeline probe step (kernels back to back, no overlap)
        fp32_ach = pst["ray_segments"] * flops_per_ray / (pst["extend_ms"] * 1e-3) / 1e12 if pst["extend_ms"] > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "extend_traffic.json")
        traffic_src = None
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            # measured DRAM bytes per ray of the captured launch x this run's average rays per launch ("per launch like achieved")
            traffic = tj["dram_bytes_per_ray"] * rays_rank0 / max(1, n_ext)
            traffic_src = tj.get("capture")
        out = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": wl["world"], "width": size.width, "height": size.height, "spp_per_gpu": wl["spp"],
                       "spp_total": spp_total, "max_depth": MAX_DEPTH, "sample_lanes_per_gpu": wl["lanes"],
                       "parallelism": f"sample-lane sharding x{n_gpus}, NCCL all-reduce of the film",
                       "l2": "256 MB flush between steps; path-state working set %.1f GB > L2" % (wl["lanes"] * size.count() * 136 / 1e9),
                       "ray_segments_per_step": total_rays / args.steps,
                       "per_rank_ms_per_step": [float(t.item()) / args.steps for t in per_rank] if n_gpus > 1 else [total_ms / args.steps]},
            "roofline": {"kernel": "k_extend", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
                         "bytes_per_ray": EXTEND_BYTES_PER_RAY, "launches": n_ext, "avg_launch_ms": ext_ms / max(1, n_ext),
                         "share_of_step": serial["share_of_step"], "serial_probe": serial,
                         "note": "scene (nodes+primitives) is L2-resident, so the kernel is FP32-issue / latency bound: see fp32. The timed region "
                                 "runs two wavefront pipelines on separate streams: avg_launch_ms / achieved are per-launch CUDA-event times "
                                 "under that overlap; serial_probe repeats one step with a single pipeline (kernels back to back), which is "
                                 "what share_of_step and the ncu launch list describe"},
            "fp32": {"kernel": "k_extend", "achieved": fp32_ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": fp32_ach / fp32_peak,
                     "nodes_per_ray": nodes_per_ray, "prims_per_ray": prims_per_ray, "flops_per_ray": flops_per_ray,
                     "note": "algorithmic flops (SURVEY 8d: 30 per box test, 41 per primitive test); peak = SMs x 128 lanes x max SM clock, no FMA"},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "ms_per_step": [round(x, 2) for x in e2e_steps], "h2d_bytes_per_step": scene.device_bytes(),
                    "scene_desc_bytes": world.upload_bytes(),
                    "phases_ms": {"columns": ["release", "scene_create", "render_call", "readback", "render_device"], "steps": e2e_phases[-len(e2e_steps):]},
                    "d2h_bytes_per_step": size.count() * 12, "note": "rt_scene_create (scene hand-off incl. building and uploading the packed scene) + render + film read-back to pinned host memory, every step"},
            "clocks": clocks, "gpu_launches": int(sum(s["kernel_launches"] for s in stats)) + args.steps,
        }
        if n_gpus == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(wl)
        print(json.dumps(out), flush=True)
    if n_gpus > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
