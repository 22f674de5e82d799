import sys, os, time
sys.path.insert(0, '/root/repo')
import raytracinginaweekend_b200 as rb
t0=time.time(); W = rb.World.named("config:c4"); print("world build %.2fs" % (time.time()-t0), flush=True)
for i in range(2):
    W.release_device(); t0=time.time(); W.device_scene(0); print("scene create %.2fs" % (time.time()-t0), flush=True)
