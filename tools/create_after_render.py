"""Scene hand-off times after other configs have rendered in the same process (what bench.py's e2e leg of the later configs sees).
    RT_PACK_LOG=1 python tools/create_after_render.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb

def render(name, w, h, lanes, spp):
    W = rb.World.named(name)
    t0 = time.perf_counter(); rb.render(rb.Size2i(w, h), lanes, spp, 50, W, seed=1, device=0); t1 = time.perf_counter()
    print(name, "render %.1f ms" % ((t1 - t0) * 1e3), flush=True)
    return W

def creates(W, name, n=3):
    for i in range(n):
        t0 = time.perf_counter(); W.release_device(); t1 = time.perf_counter(); W.device_scene(0); t2 = time.perf_counter()
        print(name, "release %.2f ms  create %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)

W2 = render("config:c2", 1920, 1080, 64, 64); creates(W2, "c2"); W2.close()
W1 = render("config:c1", 1200, 800, 10, 10); creates(W1, "c1"); W1.close()
W3 = render("config:c3", 1024, 1024, 32, 64); creates(W3, "c3")
rb.render(rb.Size2i(1024, 1024), 32, 64, 50, W3, seed=2, device=0); creates(W3, "c3 after its own render")
W3.close()
