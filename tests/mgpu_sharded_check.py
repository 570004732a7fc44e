#!/usr/bin/env python
"""Multi-GPU check of the sharded-map update (run under torchrun with >= 2 GPUs; tests/test_sharded.py launches it when
the box has them):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_sharded_check.py
Every rank holds an x-slab of the map.  (a) NCCL path: pass -> all-reduce -> step per pass.  (b) fused path: one
persistent kernel per rank, blobs through NVLink peer mailboxes.  Both must give bit-identical states on all ranks,
bit-identical to each other (same rank-order sum), and agree with the unsharded update on rank 0 to 1e-9."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    import torch
    import torch.distributed as dist

    from agi_lidar_slam_b200 import _cabi, sharded, synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = synth.small_config()
    mp = cfg["map"]
    p4 = lambda a: np.concatenate([a, np.zeros((len(a), 1), np.float32)], 1)  # noqa: E731
    kw = dict(max_scan_points=1 << 16, max_down_points=1 << 15, max_map_points=1 << 17)
    b = sharded.slab_bounds(mp[:, 0], world)
    keep = sharded.shard_indices(mp[:, 0], b, rank)
    ctx = _cabi.Context(local, **kw)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.map_build(p4(mp[keep]))
    body, _, _ = ctx.scan_preprocess(cfg["scan"], None, None, cfg["leaf"])
    ctx.scan_upload(body)
    own = (float(b[rank]), float(b[rank + 1]))
    out = {}
    for ext in (False, True):
        red, _t = sharded.nccl_reduce(ctx, dev)
        a = sharded.ShardedUpdate(ctx, own, reduce=red).update(cfg["x_prior"], cfg["P"], 0.001, 4, ext)
        ctx.blob_bind(None)
        if "peers" not in out:
            sharded.connect_peers(ctx, rank, world)
            out["peers"] = True
        f = None
        for _ in range(3):  # repeated launches: stamps are epochs, slots alternate
            f = sharded.PeerShardedUpdate(ctx, own).update(cfg["x_prior"], cfg["P"], 0.001, 4, ext)
        assert not ctx.peer_timed_out()
        assert a[2:] == f[2:], (a[2:], f[2:])
        assert np.array_equal(a[0], f[0]) and np.array_equal(a[1], f[1]), "fused exchange != NCCL exchange"
        xs = torch.tensor(f[0], dtype=torch.float64, device=dev)
        x0 = xs.clone()
        dist.broadcast(x0, 0)
        assert torch.equal(xs, x0), "ranks disagree"
        if rank == 0:
            with _cabi.Context(local, **kw) as full:
                full.map_build(p4(mp))
                full.scan_upload(body)
                ref = full.update_scan(cfg["x_prior"], cfg["P"], 0.001, 4, ext)
            assert ref[2:] == f[2:]
            assert np.abs(ref[0] - f[0]).max() < 1e-9
            print(f"ext={ext}: fused == nccl bitwise on {world} ranks; |x - unsharded| = {np.abs(ref[0] - f[0]).max():.1e}; "
                  f"valid {f[2]}, passes {f[3]}", flush=True)
    # striped ownership (lio_set_shard_stripes): the same checks with the rows dealt out in 8 m stripes
    x0, width = float(mp[:, 0].min()), 8.0
    keep = sharded.stripe_indices(mp[:, 0], x0, width, world, rank)
    ctx.map_build(p4(mp[keep]))
    ctx.scan_upload(body)
    ctx.set_shard_stripes(x0, width, world, rank)
    for ext in (False, True):
        red, _t = sharded.nccl_reduce(ctx, dev)
        a = sharded.ShardedUpdate(ctx, own, reduce=red).update(cfg["x_prior"], cfg["P"], 0.001, 4, ext)
        ctx.blob_bind(None)
        f = sharded.PeerShardedUpdate(ctx, own).update(cfg["x_prior"], cfg["P"], 0.001, 4, ext)
        assert not ctx.peer_timed_out()
        assert a[2:] == f[2:] and np.array_equal(a[0], f[0]) and np.array_equal(a[1], f[1]), "stripes: fused != NCCL"
        xs = torch.tensor(f[0], dtype=torch.float64, device=dev)
        x0t = xs.clone()
        dist.broadcast(x0t, 0)
        assert torch.equal(xs, x0t), "stripes: ranks disagree"
        if rank == 0:
            with _cabi.Context(local, **kw) as full:
                full.map_build(p4(mp))
                full.scan_upload(body)
                ref = full.update_scan(cfg["x_prior"], cfg["P"], 0.001, 4, ext)
            assert ref[2:] == f[2:] and np.abs(ref[0] - f[0]).max() < 1e-9
            print(f"stripes ext={ext}: fused == nccl bitwise; |x - unsharded| = {np.abs(ref[0] - f[0]).max():.1e}", flush=True)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_SHARDED_OK", flush=True)


if __name__ == "__main__":
    main()
