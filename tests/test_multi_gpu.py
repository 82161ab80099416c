"""N>1 path on CPU: ray sharding + image all-gather over gloo, world_size 2 (the CUDA render itself is stubbed by a
deterministic per-ray function - the sharding logic is what is under test)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_render(o, d):
    # exact element-wise arithmetic only (vectorised transcendental kernels may differ in the last ulp between a
    # slice and the full array, which is not what this test is about)
    rgb = o * 0.25 + d
    depth = o[:, 0] * 2.0 - d[:, 1]
    return {"rgb": rgb, "depth_volume": depth, "mask_volume": depth * 0.5, "normals_volume": d * 2}


def _worker(rank, world, port, n_rays, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from neumesh_b200 import parallel
    g = torch.Generator().manual_seed(0)
    o, d = torch.randn(n_rays, 3, generator=g), torch.randn(n_rays, 3, generator=g)
    sl = parallel.shard_indices(n_rays, rank, world)
    part = _fake_render(o[sl], d[sl])
    assert part["rgb"].shape[0] == parallel.shard_count(n_rays, rank, world)
    if sl.numel() == 0:
        # an empty shard goes through the real render_fused (which must not touch the library: null pointers) - the rank
        # still takes part in the all-gather below, so nobody hangs
        from neumesh_b200.renderer import render_fused
        empty = render_fused(o[sl], d[sl], None, calc_normal=True)
        assert all(empty[k].shape[0] == 0 for k in part)
        part = {k: empty[k] for k in part}
    full = parallel.gather_image(part, n_rays, rank, world)
    ref = _fake_render(o, d)
    ok = all(torch.equal(full[k], ref[k]) for k in ref)
    q.put((rank, ok, parallel.shard_count(n_rays, rank, world), 0))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rays", [1000, 1001, 3, 300])
def test_ray_sharding_all_gather_gloo(n_rays):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rays, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    assert res[0][2] + res[1][2] == n_rays  # the block-cyclic slices cover every ray exactly once


def test_block_cyclic_shards_cover_all_rays():
    from neumesh_b200 import parallel
    for n in (0, 1, 7, 1001, 640000):
        for w in (1, 2, 4, 8):
            idx = torch.arange(n)
            parts = [parallel.shard_indices(n, r, w) for r in range(w)]
            assert [len(p) for p in parts] == [parallel.shard_count(n, r, w) for r in range(w)]
            assert torch.equal(torch.sort(torch.cat(parts))[0], idx)
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= parallel.SHARD_BLOCK


def test_shard_range_partitions():
    from neumesh_b200 import parallel
    for n in (0, 1, 7, 640000):
        for w in (1, 2, 4, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker_frames(rank, world, port, rays_per_frame, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from neumesh_b200 import parallel
    g = torch.Generator().manual_seed(0)
    n = rays_per_frame * world                       # one whole frame per rank
    o, d = torch.randn(n, 3, generator=g), torch.randn(n, 3, generator=g)
    mine = slice(rank * rays_per_frame, (rank + 1) * rays_per_frame)
    full = parallel.gather_image_contiguous(_fake_render(o[mine], d[mine]), world)
    ref = _fake_render(o, d)
    q.put((rank, all(torch.equal(full[k], ref[k]) for k in ref)))
    dist.barrier()
    dist.destroy_process_group()


def test_whole_frame_sharding_all_gather_gloo():
    """The weak-scaling mode of bench.py (one frame per rank per step): contiguous slices, one all_gather_into_tensor."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_frames, args=(r, 2, port, 777, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
