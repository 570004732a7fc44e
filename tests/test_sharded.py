"""Sharded-map path: host logic on CPU with a 2-rank gloo group; device logic on one GPU with the two
ranks emulated by two contexts (the blob exchange goes through the host), against the unsharded update."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_slabs_partition_and_halo_cover_neighbours(small_cfg):
    from agi_lidar_slam_b200 import sharded

    mp = small_cfg["map"]
    for world in (2, 3, 8):
        b = sharded.slab_bounds(mp[:, 0], world)
        assert len(b) == world + 1 and np.isinf(b[0]) and np.isinf(b[-1]) and np.all(np.diff(b[1:-1]) >= 0)
        core = [np.nonzero((mp[:, 0] >= b[r]) & (mp[:, 0] < b[r + 1]))[0] for r in range(world)]
        assert sum(len(c) for c in core) == len(mp)  # every point in exactly one core slab
        assert max(len(c) for c in core) - min(len(c) for c in core) <= 2 + len(mp) // 500  # equal-count
        for r in range(world):
            keep = sharded.shard_indices(mp[:, 0], b, r)
            assert np.isin(core[r], keep).all()
            # every map point within sqrt(5) m of a point of the core slab is in the rank's local map
            lo, hi = mp[core[r], 0].min(), mp[core[r], 0].max()
            need = np.nonzero((mp[:, 0] >= lo - np.sqrt(5.0)) & (mp[:, 0] <= hi + np.sqrt(5.0)))[0]
            assert np.isin(need, keep).all()


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    blob = rng.normal(size=92)
    blob[90] = 10 + rank  # n_valid of this rank
    t = torch.from_numpy(blob.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    q.put((rank, blob, t.numpy().copy()))
    dist.destroy_process_group()


def test_blob_allreduce_two_ranks_gloo():
    """The only collective of the path: sum of the 92-double blob, identical bits on every rank."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted([q.get(timeout=120) for _ in ps], key=lambda t: t[0])
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    total = out[0][1] + out[1][1]
    assert np.array_equal(out[0][2], out[1][2])  # same bits on both ranks -> identical Kalman steps
    assert np.allclose(out[0][2], total, rtol=0, atol=1e-15) and out[0][2][90] == 21


@pytest.mark.gpu
@pytest.mark.parametrize("ext,stripes", [(False, False), (True, False), (False, True), (True, True)])
def test_two_shards_equal_unsharded_update(orc, small_cfg, ext, stripes):
    from agi_lidar_slam_b200 import _cabi, sharded

    cfg = small_cfg
    mp = cfg["map"]
    s = cfg["scan"]
    pts5 = np.concatenate([s[:, :3], np.zeros((len(s), 1), np.float32), s[:, 3:4]], 1)
    body = np.ascontiguousarray(orc.voxel_grid(pts5, 0.5)[0][:, :4])
    p4 = lambda a: np.concatenate([a, np.zeros((len(a), 1), np.float32)], 1)
    kw = dict(max_scan_points=1 << 16, max_down_points=1 << 15, max_map_points=1 << 17)
    with _cabi.Context(0, **kw) as full:
        full.map_build(p4(mp))
        full.scan_upload(body)
        ref = full.update_scan(cfg["x_prior"], cfg["P"], 0.001, 4, ext)
    world = 2
    b = sharded.slab_bounds(mp[:, 0], world)
    ranks = []
    x0, width = float(mp[:, 0].min()), 8.0  # striped ownership: 8 m stripes dealt round-robin (lio_set_shard_stripes)
    for r in range(world):
        c = _cabi.Context(0, **kw)
        if stripes:
            keep = sharded.stripe_indices(mp[:, 0], x0, width, world, r)
            c.set_shard_stripes(x0, width, world, r)
        else:
            keep = sharded.shard_indices(mp[:, 0], b, r)
        assert len(keep) < (0.9 if stripes else 0.8) * len(mp)  # a real shard, not a replica
        c.map_build(p4(mp[keep]))
        c.scan_upload(body)
        ranks.append(c)
    # lock-step emulation of the two ranks: pass on both, sum the blobs on the host, identical step on both
    for c in ranks:
        c.state_upload(cfg["x_prior"], cfg["P"])
        c.update_begin(4, ext, True)
    for _ in range(5):
        blobs = []
        for r, c in enumerate(ranks):
            c.update_pass_enqueue(ext, float(b[r]), float(b[r + 1]))
            blobs.append(c.blob_download())
        tot = blobs[0] + blobs[1]
        tot[91] = blobs[0][91]
        for c in ranks:
            c.blob_upload(tot)
            c.update_step_enqueue(0.001, ext)
    outs = [c.state_download() for c in ranks]
    for c in ranks:
        c.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])  # ranks bit-identical
    assert outs[0][2:] == ref[2:]  # matched-point count and pass count of the unsharded update
    assert np.abs(outs[0][0] - ref[0]).max() < 1e-9  # cross-rank sum order differs: 1e-9, not bitwise
    assert np.abs(outs[0][1] - ref[1]).max() < 1e-9 * np.abs(ref[1]).max()


@pytest.mark.gpu
def test_fused_peer_exchange_on_two_gpus():
    """NCCL path and in-kernel NVLink mailbox path on real GPUs (skipped on a one-GPU box)."""
    import subprocess
    import sys
    from pathlib import Path

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = Path(__file__).resolve().parents[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(_free_port()), str(root / "tests" / "mgpu_sharded_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_SHARDED_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
