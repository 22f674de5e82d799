"""Scratch GPU check: trace parity, replay-render parity, timing.  Run under gpurun."""
import sys, time, json, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import binding as ob
from raytracinginaweekend_b200 import _cabi as rt

def cmp_trace(name, W, S, org, dir, label):
    t0 = time.time(); ref = W.trace(org, dir); t1 = time.time()
    res = {}
    for mode in (0, 1, 2):
        rt.set_option("traversal", mode)
        got = S.trace(org, dir)
        ids = int((got["hit_id"] != ref["hit_id"]).sum())
        hitm = ref["hit_id"] != 0xFFFFFFFF
        both = hitm & (got["hit_id"] == ref["hit_id"])
        tb = int((got["t"][both].view(np.uint32) != ref["t"][both].view(np.uint32)).sum())
        nb = int((got["normal"][both].view(np.uint32) != ref["normal"][both].view(np.uint32)).any(axis=1).sum())
        ffb = int((got["front_face"][both] != ref["front_face"][both]).sum())
        uvd = float(np.abs(got["uv"][both] - ref["uv"][both]).max()) if both.any() else 0.0
        res[mode] = dict(id_mismatch=ids, t_bit_mismatch=tb, normal_bit_mismatch=nb, ff_mismatch=ffb, uv_maxabs=uvd)
    print(name, label, "n=", len(org), "hits=", int(hitm.sum()), "oracle_s=%.2f" % (t1 - t0), json.dumps(res), flush=True)

def rand_rays(fs, n, seed):
    rng = np.random.default_rng(seed)
    if len(fs.nodes_box):
        lo, hi = fs.nodes_box[0, :3], fs.nodes_box[0, 3:]
    else:
        lo, hi = np.array([-1, -1, -1.0]), np.array([1, 1, 1.0])
    lo = np.maximum(lo, -60); hi = np.minimum(hi, 60)
    org = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    # re-normalise in f32 the way the reference does for camera rays is not required: any direction is admissible
    return org, d

print("devices", rt.device_count(), flush=True)
for name, (pw, ph) in {"config:c1": (300, 200), "config:c2": (320, 180), "config:c3": (256, 256)}.items():
    W = ob.OracleWorld.named(name); fs = W.export()
    S = rt.DeviceScene.from_flat(fs)
    org, dir = W.primary_rays(pw, ph)
    cmp_trace(name, W, S, org, dir, "primary")
    org, dir = rand_rays(fs, 60000, 42)
    cmp_trace(name, W, S, org, dir, "random")
    rt.set_option("traversal", 2)
    # replay render parity (bit-exact expected)
    for (w, h, spp, T) in [(64, 48, 4, 1), (64, 48, 5, 2)]:
        ref, st = W.render(w, h, spp, 50, T, 7, "replay")
        got, gst = S.render(w, h, spp, 50, T, 7)
        a = ref.view(np.uint32); b = got.view(np.uint32)
        nan_ref = np.isnan(ref); nan_got = np.isnan(got)
        neq = (a != b) & ~(nan_ref & nan_got)
        print(name, "replay", (w, h, spp, T), "pix_mismatch", int(neq.any(axis=2).sum()), "of", w * h,
              "maxabs", float(np.nanmax(np.abs(ref - got))), "rays ref/gpu", st["rays"], gst["ray_segments"],
              "nan ref/gpu", int(nan_ref.any(axis=2).sum()), int(nan_got.any(axis=2).sum()), flush=True)
    S.close(); W.close()

# timing
for name, spp in [("config:c2", 16), ("config:c1", 10), ("config:c3", 32)]:
    W = ob.OracleWorld.named(name); fs = W.export(); w, h = fs.default_size
    S = rt.DeviceScene.from_flat(fs)
    rt.set_option("traversal", 2)
    S.render(w, h, 2, 50, 1, 3)
    for T in (1, 4):
        for refill in (1, 12, 16, 20, 24, 28, 32):
            rt.set_option("refill", refill)
            img, st = S.render(w, h, spp, 50, T, 3)
            print(name, "T", T, "refill", refill, "ms %.1f" % st["render_ms"], "Mrays/s %.1f" % (st["ray_segments"] / st["render_ms"] / 1e3),
                  "rays", st["ray_segments"], "iters", st["iterations"], flush=True)
    rt.set_option("refill", 20)
    for trav in (1, 0):
        rt.set_option("traversal", trav)
        img, st = S.render(w, h, spp, 50, 4, 3)
        print(name, "T 4 trav", trav, "ms %.1f" % st["render_ms"], "Mrays/s %.1f" % (st["ray_segments"] / st["render_ms"] / 1e3), flush=True)
    rt.set_option("traversal", 2); rt.set_option("count", 1)
    img, st = S.render(w, h, spp, 50, 4, 3)
    print(name, "counts: nodes/ray %.2f prims/ray %.2f exact/node %.5f" % (st["node_visits"] / st["ray_segments"], st["prim_tests"] / st["ray_segments"], st["exact_steps"] / max(1, st["node_visits"])), flush=True)
    rt.set_option("count", 0); rt.set_option("profile", 1)
    img, st = S.render(w, h, spp, 50, 4, 3)
    print(name, "profile T4: extend_ms %.1f shade_ms %.1f total %.1f" % (st["extend_ms"], st["shade_ms"], st["render_ms"]), flush=True)
    rt.set_option("profile", 0)
    S.close(); W.close()
