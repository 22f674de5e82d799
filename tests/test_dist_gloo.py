"""world_size-2 gloo test of the multi-GPU host logic (lane sharding + film all-reduce + 1/T scale) on CPU.
The per-rank plane producer is substituted by the oracle (the product path has no CPU implementation); the collective,
the lane arithmetic and the final scaling are the product's own code (raytracinginaweekend_b200/dist.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, case, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle import binding as ob
    from raytracinginaweekend_b200 import dist as rdist
    from raytracinginaweekend_b200.api import RenderMode, Size2i
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    name, w, h, spp, T, seed = case
    W = ob.OracleWorld.named(name)

    def plane_fn(lane_first, lane_stride):
        plane, _ = W.render_planes(w, h, spp, 50, T, seed, "replay", lane_first, lane_stride)
        return torch.from_numpy(plane)
    film = rdist.render_distributed(None, Size2i(w, h), T, spp, 50, RenderMode.Default, seed, plane_fn=plane_fn)
    np.save(os.path.join(out_dir, f"film_{rank}.npy"), film.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case", [("demo:defocus_blur", 24, 16, 6, 2, 5), ("demo:simple_plane", 20, 12, 7, 4, 9), ("demo:defocus_blur", 16, 12, 1, 4, 3)])
def test_two_rank_film_combine_matches_single_process_render(case, tmp_path, oracle_world):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, case, str(tmp_path)), nprocs=2, join=True)
    name, w, h, spp, T, seed = case
    ref, _ = oracle_world(name).render(w, h, spp, 50, T, seed, "replay")
    f0, f1 = np.load(tmp_path / "film_0.npy"), np.load(tmp_path / "film_1.npy")
    assert np.array_equal(f0.view(np.uint32), f1.view(np.uint32))          # every rank holds the same film
    nan = np.isnan(ref)
    assert np.array_equal(np.isnan(f0), nan)
    if min(T, spp) <= 2:
        assert np.array_equal(f0.view(np.uint32)[~nan], ref.view(np.uint32)[~nan])   # a+b is commutative: bit-exact
    else:
        assert np.allclose(f0[~nan], ref[~nan], rtol=2e-6, atol=1e-7)      # summation order differs from merge_planes
