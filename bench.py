#!/usr/bin/env python
"""bench.py — headline benchmark of the render hot path (BASELINE.json metric: Mrays/s = ray segments / second / 1e6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2] [--only-headline]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one full render of the workload through the hand-written CUDA path.  The headline workload is BASELINE.json
configs[1] = config:c2 (two suzanne.obj instances, Lambert + metal, 1920x1080, 64 spp, depth 50).  N > 1: weak scaling —
every rank renders 64 spp (its share of the 64*N sample lanes of ONE render() call) and the film is combined with one NCCL
all-reduce (pre-multiplied sum: the 1/T of merge_planes is folded into the collective) inside the timed region.

The JSON line also carries
  roofline      the dominant kernel (k_extend) against its binding roof: FP32 issue for scenes that live in L2, HBM for C4; both
                fractions are computed from BACK-TO-BACK kernel time (a probe step with one pipeline), never from event deltas of
                overlapping streams;
  e2e           the same metric through the reference-facing call with HOST buffers: rt_scene_create (scene hand-off, H2D) +
                rt_render into a caller buffer (film D2H) every step;
  configs       every other BASELINE config (C1, C3, C4 full size; C5 at 3840x2160 with a reduced spp, stated) with value, e2e,
                node / primitive tests per ray, both roofline fractions and — at N = 1 — the image RMSE against the reference CPU
                render (the oracle's FAITHFUL mode) at equal spp and depth with the Monte-Carlo bound of SURVEY §8(d);
  strong        a FIXED amount of work (C5, 64 spp total) split over the N GPUs: what the north star's "4096 spp split
                across 1/2/4/8 GPUs" measures; the weak headline cannot show per-render fixed costs, this does;
  cpu_baseline  the oracle's restatement of the reference's multithreaded CPU render() on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {  # BASELINE.json configs (DESIGN.md "Workload definitions"); rmse = (w, h) of the CPU-side comparison render at equal spp
    "c1": dict(world="config:c1", width=1200, height=800, spp=10, lanes=10, cpu_sample=(300, 200, 4), rmse=(120, 80)),
    "c2": dict(world="config:c2", width=1920, height=1080, spp=64, lanes=64, cpu_sample=(640, 360, 4), rmse=(64, 36)),
    "c3": dict(world="config:c3", width=1024, height=1024, spp=256, lanes=32, cpu_sample=(256, 256, 16), rmse=(64, 64)),
    "c4": dict(world="config:c4", width=3840, height=2160, spp=16, lanes=16, cpu_sample=(240, 135, 2), rmse=None,
               rmse_note="not run: the reference's node rule visits ~0.4 N of the 10^6 nodes per ray, one 16-spp pixel costs ~1 s of CPU"),
    "c5": dict(world="config:c5", width=3840, height=2160, spp=16, lanes=16, cpu_sample=(240, 135, 2), rmse=(32, 18),
               spp_note="BASELINE spp is 4096 (34 G paths, ~4 min per step on one GPU); 16 spp here, 64 spp in `strong`; Mrays/s is spp-invariant"),
}
STRONG = dict(world="config:c5", width=3840, height=2160, spp_total=64, lanes_total=64)
MAX_DEPTH = 50
EXTEND_BYTES_PER_RAY = 40        # extend kernel: reads origin+direction (2 x 16 B), writes (hit ref, t) 8 B  — DESIGN.md
RECORD_BYTES = 64                # one node / primitive record
NODE_FLOPS, PRIM_FLOPS = 60, 41  # per node visit (two child boxes, 30 each) / per primitive test
L2_BYTES = 126e6


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
            self.proc.wait()
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


# ------------------------------------------------------------------------------------------------ CPU side (oracle: checker / baseline only)
def cpu_sample_dims(wl, cores):
    """Bounded CPU sample: every one of the `cores` worker threads gets >= 1 sample per pixel (the reference splits SAMPLES
    across threads), and the image shrinks so the work stays ~constant (C2: ~0.9 M paths, 10-15 s on 16 cores)."""
    w, h, spp = wl["cpu_sample"]
    target = w * h * spp
    spp = max(spp, cores)
    scale = (target / float(w * h * spp)) ** 0.5
    return max(48, int(w * scale)), max(27, int(h * scale)), spp


CPU_NOTE = ("oracle = C++ restatement of the reference's render() (thread-per-sample-split, reference node rule, recursive traversal), "
            "g++ -O3 -ffp-contract=off like cargo --release (opt-level 3); rustc is not in this image")


def run_reference(args, wl):
    """--impl reference: the reference's own CPU implementation of the path.  It is Rust and cannot be built in this image
    (no rustc / cargo; oracle/ref_harness/build.sh builds it the day a toolchain exists), so this arm times the oracle's literal
    C++ restatement of render() with all host threads, each step a bounded sample of the workload."""
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
    sample = f"{wl['world']} at {w}x{h}, {spp} spp, depth {MAX_DEPTH} per step (Mrays/s is resolution- and spp-invariant); {CPU_NOTE}"
    out = {"impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": wl["world"], "width": wl["width"], "height": wl["height"], "spp_per_gpu": wl["spp"],
                      "spp_total": wl["spp"] * args.gpus, "max_depth": MAX_DEPTH, "sample_lanes_per_gpu": wl["lanes"],
                      "note": "reference arm = CPU restatement of the Rust renderer (rustc unavailable); bounded sample per step"},
           "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def cpu_baseline(wl):
    from oracle import binding as ob
    cores = os.cpu_count() or 1
    w, h, spp = cpu_sample_dims(wl, cores)
    W = ob.OracleWorld.named(wl["world"])
    _, st = W.render(w, h, spp, MAX_DEPTH, cores, 4242, "faithful")
    return {"value": st["rays"] / st["seconds"] / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": f"{wl['world']} at {w}x{h}, {spp} spp, depth {MAX_DEPTH}: {st['rays']} ray segments in {st['seconds']:.2f} s ({CPU_NOTE})"}


def mc_gate(gpu, sumsq, ref, n):
    """SURVEY §8(d) acceptance: RMSE(R, G) <= 1.25 sqrt(2 mean(sigma^2) / N) over pixels finite in both; per-channel mean error
    <= 4 sqrt(2 mean(sigma^2) / (N P)); NaN-mask IoU.  sigma^2 per channel from the GPU's own first / second moments."""
    nan_g, nan_r = np.isnan(gpu).any(axis=2), np.isnan(ref).any(axis=2)
    union = int((nan_g | nan_r).sum())
    iou = 1.0 if union == 0 else float((nan_g & nan_r).sum()) / union
    ok = ~(nan_g | nan_r) & np.isfinite(sumsq)
    g64, r64 = gpu[ok].astype(np.float64), ref[ok].astype(np.float64)
    var = np.maximum(sumsq[ok].astype(np.float64) / n - (g64 ** 2).sum(axis=1), 0.0) * (n / max(n - 1, 1)) / 3.0
    mean_var = float(var.mean())
    rmse = math.sqrt(float(((g64 - r64) ** 2).mean()))
    bound = 1.25 * math.sqrt(2.0 * mean_var / n)
    bias = max(abs(float((g64[:, c] - r64[:, c]).mean())) for c in range(3))
    bias_bound = 4.0 * math.sqrt(2.0 * mean_var / (n * max(int(ok.sum()), 1)))
    return {"rmse": rmse, "bound": bound, "within_bound": bool(rmse <= bound), "bias": bias, "bias_bound": bias_bound,
            "nan_pixels_gpu": int(nan_g.sum()), "nan_pixels_ref": int(nan_r.sum()), "nan_mask_iou": iou, "pixels": int(ok.sum())}


# ------------------------------------------------------------------------------------------------ GPU side
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world_size > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.stream = torch.cuda.current_stream(self.dev)
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=self.dev)   # > 126 MB L2
        self.hbm_peak, self.sm_mhz, self.peak_kind = peaks()
        self.sms = torch.cuda.get_device_properties(self.dev).multi_processor_count
        self.fp32_peak = self.sms * 128 * self.sm_mhz * 1e6 / 1e12   # TFLOP/s without FMA (parity code is --fmad=false)
        self.launches = 0

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world_size > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_sum(self, ms, rays):
        """(max over ranks of the summed step time, sum over ranks of the ray segments)."""
        t = self.torch.tensor([ms], dtype=self.torch.float64, device=self.dev)
        r = self.torch.tensor([rays], dtype=self.torch.float64, device=self.dev)
        if self.world_size > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(r, op=self.dist.ReduceOp.SUM)
        return float(t.item()), float(r.item())

    def step_fn(self, scene, plane, width, height, spp_total, lanes_total):
        """One device-resident step: this rank's lanes -> film combine (NCCL pre-multiplied sum, or the 1/T scale kernel at N = 1)."""
        from raytracinginaweekend_b200 import _cabi
        from raytracinginaweekend_b200 import dist as rdist
        mult = float(np.float32(1.0) / np.float32(rdist.n_tasks(lanes_total, spp_total)))

        def step(seed):
            st = scene.render_planes_device(plane.data_ptr(), self.rank, self.world_size, width, height, spp_total, MAX_DEPTH, lanes_total,
                                            seed, 0, self.stream.cuda_stream)
            rdist.combine_film(plane, mult, self.world_size)
            return st
        return step

    def timed(self, step, warmup, steps, seed0=200):
        for i in range(warmup):
            step(100 + i)
        self.barrier()
        ms, stats = [], []
        for i in range(steps):
            self.flush.fill_(float(i))          # evict the previous step's working set from L2 (outside the timed events)
            e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            stats.append(step(seed0 + i))
            e1.record(self.stream)
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        self.barrier()
        self.launches += int(sum(s["kernel_launches"] for s in stats)) + steps
        total_ms, total_rays = self.max_sum(sum(ms), float(sum(s["ray_segments"] for s in stats)))
        return {"value": total_rays / (total_ms * 1e-3) / 1e6, "ms_per_step": total_ms / steps, "rays_per_step": total_rays / steps,
                "local_ms": ms, "stats": stats}

    def probes(self, scene, plane, width, height, spp_total, lanes_total, scene_bytes):
        """Two extra untimed steps on this rank: (1) ONE pipeline + per-launch events -> kernels run back to back, so the summed
        extend / shade times are real device time and comparable with the serialised ncu launch list; (2) the counting kernel
        variant -> node visits / primitive tests per ray.  From these: both roofline fractions of k_extend."""
        from raytracinginaweekend_b200 import _cabi
        _cabi.set_option("pipes", 1); _cabi.set_option("profile", 1)
        p = scene.render_planes_device(plane.data_ptr(), self.rank, self.world_size, width, height, spp_total, MAX_DEPTH, lanes_total, 200, 0,
                                       self.stream.cuda_stream)
        _cabi.set_option("pipes", 0); _cabi.set_option("profile", 0); _cabi.set_option("count", 1)
        c = scene.render_planes_device(plane.data_ptr(), self.rank, self.world_size, width, height, spp_total, MAX_DEPTH, lanes_total, 200, 0,
                                       self.stream.cuda_stream)
        _cabi.set_option("count", 0)
        self.torch.cuda.synchronize(self.dev)
        rays = max(1, p["ray_segments"])
        npr, ppr = c["node_visits"] / max(1, c["ray_segments"]), c["prim_tests"] / max(1, c["ray_segments"])
        ext_s = max(p["extend_ms"], 1e-9) * 1e-3
        flops = NODE_FLOPS * npr + PRIM_FLOPS * ppr
        fp32 = rays * flops / ext_s / 1e12
        in_l2 = scene_bytes < L2_BYTES   # scene bytes count against HBM only when nodes + primitives do not fit L2
        bytes_per_ray = EXTEND_BYTES_PER_RAY + (0 if in_l2 else RECORD_BYTES * (npr + ppr))
        hbm = rays * bytes_per_ray / ext_s / 1e9
        lanes = c["busy_lane_rounds"] / max(1, c["warp_rounds"])
        return {"kernel": "k_extend", "extend_ms": p["extend_ms"], "shade_ms": p["shade_ms"], "step_ms": p["render_ms"], "launches": int(p["iterations"]),
                "avg_launch_ms": p["extend_ms"] / max(1, p["iterations"]), "share_of_step": p["extend_ms"] / max(1e-9, p["render_ms"]),
                "nodes_per_ray": npr, "prims_per_ray": ppr, "flops_per_ray": flops, "bytes_per_ray": bytes_per_ray, "scene_in_l2": in_l2,
                "fp32_tflops": fp32, "fp32_frac": fp32 / self.fp32_peak, "hbm_gbs": hbm, "hbm_frac": hbm / self.hbm_peak,
                "lanes_holding_a_ray": lanes, "slots": int(p["slots"]), "rays": int(rays)}

    def roofline(self, pr, config="c2"):
        """The JSON `roofline` object from a probe: the BINDING roof first (FP32 issue when the scene lives in L2, HBM otherwise)."""
        traffic, src = None, None
        tpath = os.path.join(ROOT, "profiles", "extend_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            tj = tj.get(config, tj if config == "c2" and "dram_bytes_per_ray" in tj else None)
            if tj:
                traffic = tj["dram_bytes_per_ray"] * pr["rays"] / max(1, pr["launches"])   # per launch, like `achieved`
                src = tj.get("capture")
        common = {"kernel": "k_extend", "launches": pr["launches"], "avg_launch_ms": pr["avg_launch_ms"], "share_of_step": pr["share_of_step"],
                  "traffic": traffic, "traffic_source": src, "timing": "one probe step with a single pipeline: extend launches back to back, CUDA events "
                  "on the launching stream around every launch (the timed steps overlap two pipelines, whose event deltas would double-count)",
                  "nodes_per_ray": pr["nodes_per_ray"], "prims_per_ray": pr["prims_per_ray"], "lanes_holding_a_ray": pr["lanes_holding_a_ray"]}
        fp32 = {"bound": "fp32-issue", "achieved": pr["fp32_tflops"], "peak": self.fp32_peak, "unit": "TFLOP/s", "frac": pr["fp32_frac"],
                "flops_per_ray": pr["flops_per_ray"], "peak_source": f"{self.sms} SMs x 128 lanes x {self.sm_mhz:.0f} MHz (MEASURED_PEAKS.json {self.peak_kind}), no FMA: parity code is --fmad=false",
                "algorithmic": "30 flops per box test (2 per node visit), 41 per primitive test (SURVEY 8d) x measured tests per ray"}
        hbm = {"bound": "hbm", "achieved": pr["hbm_gbs"], "peak": self.hbm_peak, "unit": "GB/s", "frac": pr["hbm_frac"], "bytes_per_ray": pr["bytes_per_ray"],
               "peak_source": f"MEASURED_PEAKS.json ({self.peak_kind})",
               "algorithmic": "40 B per ray (origin + direction in, hit out)" + ("" if pr["scene_in_l2"] else " + 64 B per node visit and primitive test (scene > L2)")}
        # Which roof binds?  FP32 issue while the scene lives in L2.  For a scene larger than L2 (C4, 384 MB) the ALGORITHMIC bytes
        # say HBM — but the measured DRAM traffic (ncu, `traffic`) shows whether L2 absorbs the node / primitive re-reads anyway:
        # when the measured traffic is below a quarter of the algorithmic bytes, HBM is not what limits the kernel.
        measured_bpr = traffic * pr["launches"] / max(1, pr["rays"]) if traffic else None
        hbm_binds = (not pr["scene_in_l2"]) and not (measured_bpr is not None and measured_bpr < 0.25 * pr["bytes_per_ray"])
        if measured_bpr is not None:
            hbm["measured_dram_bytes_per_ray"] = measured_bpr
            hbm["measured_dram_gbs"] = measured_bpr * pr["rays"] / (pr["extend_ms"] * 1e-3) / 1e9
            hbm["measured_frac_of_peak"] = hbm["measured_dram_gbs"] / self.hbm_peak
        first, second = (hbm, fp32) if hbm_binds else (fp32, hbm)
        return {**first, **common, "secondary": second,
                "note": "no tensor-core work exists on this path. FP32 issue binds while nodes + primitives are served from L2: C1, C2, C3, C5 fit it, and for C4 "
                        "(384 MB > 126 MB L2) ncu measures 238 B of DRAM traffic per ray against 4.5 kB algorithmic (L2 hit 92 %), so HBM does not bind there either; "
                        "the algorithmic-bytes HBM fraction is kept as `secondary`"}

    def e2e_host(self, world, wl, spp_total, lanes_total, steps):
        """End to end through the reference-facing call with HOST buffers (what INTEGRATION.md's shim does): every step hands the
        scene over again (rt_scene_create: pack + accelerator build + H2D) and renders into a caller-owned host buffer (film D2H)."""
        import raytracinginaweekend_b200 as rb
        size = rb.Size2i(wl["width"], wl["height"])
        ms, rays, phases = [], 0.0, []
        for i in range(-1, steps):
            t0 = time.perf_counter()
            world.release_device()
            world.device_scene(self.local_rank)          # rt_scene_create
            t1 = time.perf_counter()
            _, st = rb.render(size, lanes_total, spp_total, MAX_DEPTH, world, seed=300 + i, device=self.local_rank, return_stats=True)   # rt_render (host out_rgb)
            t2 = time.perf_counter()
            if i < 0:
                continue
            ms.append((t2 - t0) * 1e3); rays += st["ray_segments"]
            phases.append([round((t1 - t0) * 1e3, 2), round((t2 - t1) * 1e3, 2), round(st["render_ms"], 2)])
            self.launches += int(st["kernel_launches"])
        return {"value": rays / (sum(ms) * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": [round(x, 2) for x in ms],
                "h2d_bytes_per_step": int(world.upload_bytes()), "d2h_bytes_per_step": size.count() * 12,
                "phases_ms": {"columns": ["rt_scene_create", "rt_render (host film)", "of which device render"], "steps": phases},
                "note": "rt_scene_create (scene hand-off: pack, accelerator build, upload) + rt_render into a pageable caller buffer, every step"}

    def e2e_dist(self, world, wl, spp_total, lanes_total, steps):
        """N > 1: the public multi-process API (render_distributed): scene hand-off + this rank's lanes + NCCL combine + film to pinned host memory."""
        import raytracinginaweekend_b200 as rb
        from raytracinginaweekend_b200 import dist as rdist
        size = rb.Size2i(wl["width"], wl["height"])
        pinned = self.torch.empty((size.height, size.width, 3), dtype=self.torch.float32).pin_memory()
        self.barrier()
        ms, rays = [], 0.0
        for i in range(-1, steps):
            t0 = time.perf_counter()
            world.release_device()
            world.device_scene(self.local_rank)
            film, st = rdist.render_distributed(world, size, lanes_total, spp_total, MAX_DEPTH, rb.RenderMode.Default, 300 + i, device=self.dev, return_stats=True)
            pinned.copy_(film, non_blocking=False)
            self.torch.cuda.synchronize(self.dev)
            if i < 0:
                continue
            ms.append((time.perf_counter() - t0) * 1e3); rays += st["ray_segments"]
            self.launches += int(st["kernel_launches"]) + 1
        total_ms, total_rays = self.max_sum(sum(ms), rays)
        return {"value": total_rays / (total_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": [round(x, 2) for x in ms],
                "h2d_bytes_per_step": int(world.upload_bytes()), "d2h_bytes_per_step": size.count() * 12,
                "note": "per rank: rt_scene_create + rt_render_planes_device + NCCL all-reduce (pre-multiplied sum) + film copy to pinned host memory, every step"}

    def rmse(self, world, wl):
        """GPU film (+ second moments) vs the reference CPU render (oracle FAITHFUL mode) at equal spp and depth, reduced resolution."""
        from oracle import binding as ob
        w, h = wl["rmse"]
        spp, T = wl["spp"], min(wl["lanes"], wl["spp"])
        t0 = time.perf_counter()
        ref, ost = ob.OracleWorld.named(wl["world"]).render(w, h, spp, MAX_DEPTH, os.cpu_count() or 1, 4242, "faithful")
        cpu_s = time.perf_counter() - t0
        scene = world.device_scene(self.local_rank)
        film = self.torch.empty((h, w, 3), dtype=self.torch.float32, device=self.dev)
        sumsq = self.torch.empty((h, w), dtype=self.torch.float32, device=self.dev)
        scene.render_moments_device(film.data_ptr(), sumsq.data_ptr(), w, h, spp, MAX_DEPTH, T, 99, 0, self.stream.cuda_stream)
        self.torch.cuda.synchronize(self.dev)
        g = mc_gate(film.cpu().numpy(), sumsq.cpu().numpy(), ref, spp)
        g.update({"size": [w, h], "spp": spp, "cpu_seconds": round(cpu_s, 2), "reference": "oracle FAITHFUL render (sequential Xoroshiro streams, rejection samplers, "
                  "reference traversal) — independent random numbers, equal spp and depth; bound = 1.25 sqrt(2 mean(sigma^2) / N), SURVEY 8(d)"})
        return g

    def run_workload(self, name, wl, warmup, steps, with_rmse, e2e_steps):
        import raytracinginaweekend_b200 as rb
        n = self.world_size
        t0 = time.perf_counter()
        world = rb.World.named(wl["world"])
        t1 = time.perf_counter()
        scene = world.device_scene(self.local_rank)     # inputs resident in HBM before the timed region
        t2 = time.perf_counter()
        scene_bytes = int(scene.device_bytes())
        spp_total, lanes_total = wl["spp"] * n, wl["lanes"] * n          # weak scaling: fixed work per GPU
        plane = self.torch.empty((wl["height"], wl["width"], 3), dtype=self.torch.float32, device=self.dev)
        step = self.step_fn(scene, plane, wl["width"], wl["height"], spp_total, lanes_total)
        r = self.timed(step, warmup, steps)
        pr = self.probes(scene, plane, wl["width"], wl["height"], spp_total, lanes_total, scene_bytes)
        e2e = (self.e2e_host if n == 1 else self.e2e_dist)(world, wl, spp_total, lanes_total, e2e_steps)
        out = {"workload": wl["world"], "width": wl["width"], "height": wl["height"], "spp_per_gpu": wl["spp"], "spp_total": spp_total,
               "sample_lanes_per_gpu": wl["lanes"], "path_slots": pr["slots"], "value": r["value"], "unit": "Mrays/s", "ms_per_step": r["ms_per_step"],
               "steps": steps, "ray_segments_per_step": r["rays_per_step"], "e2e": e2e, "roofline": self.roofline(pr, name),
               "scene_device_bytes": scene_bytes, "world_build_s": round(t1 - t0, 3), "scene_create_s": round(t2 - t1, 3)}
        for k in ("spp_note",):
            if k in wl:
                out[k] = wl[k]
        if with_rmse:
            out["rmse"] = self.rmse(world, wl) if wl.get("rmse") else {"skipped": wl.get("rmse_note", "")}
        world.close()
        return out, r

    def strong(self, steps):
        """A fixed job (C5, 64 spp in total, 64 lanes) split over the N GPUs: the north star's sample split with the film gather."""
        import raytracinginaweekend_b200 as rb
        world = rb.World.named(STRONG["world"])
        scene = world.device_scene(self.local_rank)
        plane = self.torch.empty((STRONG["height"], STRONG["width"], 3), dtype=self.torch.float32, device=self.dev)
        step = self.step_fn(scene, plane, STRONG["width"], STRONG["height"], STRONG["spp_total"], STRONG["lanes_total"])
        r = self.timed(step, 1, steps)
        # split of the step on this rank: render vs the film combine (all-reduce of 99.5 MB + nothing else)
        self.barrier()
        e0, e1, e2 = (self.torch.cuda.Event(enable_timing=True) for _ in range(3))
        from raytracinginaweekend_b200 import dist as rdist
        e0.record(self.stream)
        st = scene.render_planes_device(plane.data_ptr(), self.rank, self.world_size, STRONG["width"], STRONG["height"], STRONG["spp_total"], MAX_DEPTH,
                                        STRONG["lanes_total"], 7, 0, self.stream.cuda_stream)
        e1.record(self.stream)
        rdist.combine_film(plane, 1.0 / STRONG["lanes_total"], self.world_size)
        e2.record(self.stream)
        e2.synchronize()
        out = {"workload": STRONG["world"], "width": STRONG["width"], "height": STRONG["height"], "spp_total": STRONG["spp_total"],
               "lanes_total": STRONG["lanes_total"], "lanes_per_gpu": STRONG["lanes_total"] // self.world_size, "scaling": "strong", "value": r["value"], "unit": "Mrays/s",
               "ms_per_step": r["ms_per_step"], "steps": steps, "ray_segments_per_step": r["rays_per_step"],
               "rank0_render_ms": e0.elapsed_time(e1), "rank0_combine_ms": e1.elapsed_time(e2), "rank0_iterations": int(st["iterations"]),
               "note": "total work fixed, per-GPU work = 1/N: what limits it is the per-render fixed cost (the >= 50 wavefront iterations of a depth-50 path, "
                       "whose duration stops shrinking with the ray count; round 2 parks the rays that outlast a launch's ordinary tail instead of "
                       "waiting for them, DESIGN 9.2) — the film all-reduce is < 1 ms"}
        world.close()
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(WORKLOADS), help="headline workload (anything but c2 invalidates the headline)")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (invalidates the headline number)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only-headline", action="store_true", help="skip the other configs, the strong record and the RMSE renders")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.config])
    if args.spp:
        wl["spp"] = args.spp
    if args.impl == "reference":
        return run_reference(args, wl)

    B = Bench(args)
    n = B.world_size
    sampler = ClockSampler(B.local_rank)
    if B.rank == 0:
        sampler.start()
    head, r = B.run_workload(args.config, wl, max(3, args.warmup), args.steps, with_rmse=(n == 1 and not args.only_headline), e2e_steps=max(1, args.steps))
    clocks = sampler.stop() if B.rank == 0 else None
    per_rank = [B.torch.zeros(1, dtype=B.torch.float64, device=B.dev) for _ in range(n)]
    if n > 1:
        B.dist.all_gather(per_rank, B.torch.tensor([sum(r["local_ms"])], dtype=B.torch.float64, device=B.dev))
    configs, strong = {}, None
    if not args.only_headline:
        for name in sorted(WORKLOADS):
            if name == args.config:
                continue
            configs[name], _ = B.run_workload(name, WORKLOADS[name], 1, 2, with_rmse=(n == 1), e2e_steps=2)
        strong = B.strong(2)

    if B.rank == 0:
        out = {
            "metric": "Mrays/s", "value": head["value"], "unit": "Mrays/s", "n_gpus": n, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": wl["world"], "width": wl["width"], "height": wl["height"], "spp_per_gpu": wl["spp"],
                       "spp_total": wl["spp"] * n, "max_depth": MAX_DEPTH, "sample_lanes_per_gpu": wl["lanes"], "path_slots": head["path_slots"],
                       "parallelism": f"sample-lane sharding x{n}, NCCL all-reduce (pre-multiplied sum) of the film",
                       "l2": "256 MB flush between steps; path-state working set %.1f GB > L2" % (head["path_slots"] * 120 / 1e9),
                       "ray_segments_per_step": head["ray_segments_per_step"],
                       "per_rank_ms_per_step": [float(t.item()) / args.steps for t in per_rank] if n > 1 else [head["ms_per_step"]]},
            "roofline": head["roofline"], "e2e": head["e2e"], "clocks": clocks, "gpu_launches": int(B.launches),
        }
        if "rmse" in head:
            out["rmse"] = head["rmse"]
        if configs:
            out["configs"] = configs
        if strong:
            out["strong"] = strong
        if n == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(wl)
        print(json.dumps(out), flush=True)
    # leave the device clean: no kernels in flight, the library's workspaces (up to 19 GB) and the flush buffer released
    from raytracinginaweekend_b200 import _cabi
    B.torch.cuda.synchronize(B.dev)
    _cabi.shutdown()
    B.flush = None
    B.torch.cuda.empty_cache()
    if n > 1:
        B.dist.barrier()
        B.dist.destroy_process_group()


if __name__ == "__main__":
    main()
