import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginaweekend_b200 as rb
for name in (sys.argv[1:] or ["config:c5", "config:c2", "config:c1", "config:c3"]):
    t0 = time.time(); W = rb.World.named(name); print(name, "world build %.3fs" % (time.time() - t0), flush=True)
    for i in range(3):
        W.release_device(); t0 = time.time(); W.device_scene(0); print(name, "scene create %.3fs" % (time.time() - t0), flush=True)
