"""bench.py's run_workload for a few configs in one process (diagnosis of cross-config effects), with RT_PACK_LOG phase times on stderr.
    RT_PACK_LOG=1 python tools/bench_two.py c1 c3"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

names = sys.argv[1:] or ["c1", "c3"]
B = bench.Bench(argparse.Namespace(gpus=1, steps=2, warmup=1))
for name in names:
    rmse = not name.endswith("-")
    name = name.rstrip("-")
    print("==== " + name, file=sys.stderr, flush=True)
    out, _ = B.run_workload(name, bench.WORKLOADS[name], 1, 2, with_rmse=rmse, e2e_steps=2)
    print(name, out["scene_create_s"], out["e2e"]["phases_ms"]["steps"], flush=True)
from raytracinginaweekend_b200 import _cabi
B.torch.cuda.synchronize(B.dev)
_cabi.shutdown()
