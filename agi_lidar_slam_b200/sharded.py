"""Spatially sharded map (BASELINE.json config 5, SURVEY.md §8e case 2): host side.

The map is cut over `world` ranks along x -- one equal-count slab per rank (`slab_bounds` / `shard_indices`), or stripes
of a few metres dealt round-robin (`stripe_indices` + `Context.set_shard_stripes`: every rank then gets a share of a
scan wherever the robot is).  Rank r holds the points of its region plus a halo of sqrt(knn_max_d2) + margin metres on
both sides, so every neighbour set that can pass the validity gate (d2[4] <= 5, esekfom.hpp:147) of a query inside
the region is complete in the rank's local map.  During an update a rank searches and contributes only the rows it
OWNS: those whose p_world.x at their last search pass falls into its region (the same bits on every rank; the
ownership arguments of lio_update_pass_enqueue / lio_update_enqueue_sharded).  The 92-double blob {HtH 78, Hth 12,
n_valid, searched} is summed over the ranks in rank order (inside the kernel over NVLink peer mailboxes, or by an
NCCL all-reduce / any `reduce` callable between the pass and the step) and every rank performs the identical Kalman
step, so the states stay bit-identical across ranks.  Nothing else crosses NVLink.
"""
from __future__ import annotations

import numpy as np

HALO_MARGIN = 0.5  # metres on top of sqrt(5): the pose moves by centimetres between the passes of one update


def slab_bounds(x_coords: np.ndarray, world: int) -> np.ndarray:
    """Equal-count cut points along x: (world + 1,) float32 with -inf / +inf at the ends.  Deterministic on every rank
    (computed from the same array)."""
    xs = np.sort(np.asarray(x_coords, np.float32))
    cuts = [np.float32(-np.inf)]
    for r in range(1, world):
        cuts.append(xs[(len(xs) * r) // world])
    cuts.append(np.float32(np.inf))
    return np.asarray(cuts, np.float32)


def shard_indices(x_coords: np.ndarray, bounds: np.ndarray, rank: int, knn_max_d2: float = 5.0) -> np.ndarray:
    """Indices of the map points rank `rank` keeps: core slab [bounds[r], bounds[r+1]) plus the halo."""
    halo = np.float32(np.sqrt(knn_max_d2) + HALO_MARGIN)
    x = np.asarray(x_coords, np.float32)
    return np.nonzero((x >= bounds[rank] - halo) & (x < bounds[rank + 1] + halo))[0]


def stripe_of(x_coords, x_origin: float, width: float) -> np.ndarray:
    """Stripe number floor((x - origin) / width) in FP32, the arithmetic of the kernels (owns_row, csrc/lio_pass.cu)."""
    x = np.asarray(x_coords, np.float32)
    return np.floor((x - np.float32(x_origin)) * (np.float32(1.0) / np.float32(width))).astype(np.int64)


def stripe_indices(x_coords: np.ndarray, x_origin: float, width: float, world: int, rank: int,
                   knn_max_d2: float = 5.0) -> np.ndarray:
    """Indices of the map points rank `rank` keeps under striped ownership (lio_set_shard_stripes): the points of its
    stripes (stripe mod world == rank) plus the halo on both sides of every one of them.  width must exceed the halo."""
    halo = np.float32(np.sqrt(knn_max_d2) + HALO_MARGIN)
    assert width > halo
    x = np.asarray(x_coords, np.float32)
    keep = np.zeros(len(x), bool)
    for dx in (np.float32(0), -halo, halo):  # own stripe, or within the halo of a neighbouring stripe that is ours
        keep |= np.mod(stripe_of(x + dx, x_origin, width), world) == rank
    return np.nonzero(keep)[0]


class ShardedUpdate:
    """update_iterated_dyn_share_modified (esekfom.hpp:270-346) over a sharded map.

    ctx     : this rank's context, whose map was built from map_xyz[shard_indices(...)]
    own     : (x_min, x_max) core slab of this rank
    reduce  : callable that sums the 92-double blob over the ranks IN PLACE on the device (see `nccl_reduce`) -- or
              None together with `host_reduce`, a callable blob92 -> summed blob92 on the host (tests, gloo).
    """

    def __init__(self, ctx, own, reduce=None, host_reduce=None):
        self.ctx = ctx
        self.own = (float(own[0]), float(own[1]))
        self.reduce = reduce
        self.host_reduce = host_reduce
        assert (reduce is None) != (host_reduce is None)

    def update(self, x, P, R=0.001, max_iter=4, extrinsic_est=False):
        c = self.ctx
        c.state_upload(x, P)
        c.update_begin(max_iter, extrinsic_est, True)
        for _ in range(max_iter + 1):  # the kernels exit at once when the loop has finished on the device
            c.update_pass_enqueue(extrinsic_est, self.own[0], self.own[1])
            if self.reduce is not None:
                self.reduce()
            else:
                c.blob_upload(self.host_reduce(c.blob_download()))
            c.update_step_enqueue(R, extrinsic_est)
        return c.state_download()


def nccl_reduce(ctx, device):
    """Binds the context's blob to a torch tensor and returns (reduce_callable, tensor): an in-place NCCL all-reduce
    (sum, 92 doubles = 736 bytes: pure latency over NVLink) on torch's current stream."""
    import torch
    import torch.distributed as dist

    t = torch.zeros(92, dtype=torch.float64, device=device)
    ctx.blob_bind(t.data_ptr())

    def reduce():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    return reduce, t


def connect_peers(ctx, rank: int, world: int):
    """One-time setup of the in-kernel exchange: all-gather the ranks' 64-byte mailbox handles (cudaIpc) and map them."""
    import torch.distributed as dist

    mine = ctx.peer_handle().tobytes()
    allh = [None] * world
    dist.all_gather_object(allh, mine)
    ctx.peer_connect(rank, world, np.frombuffer(b"".join(allh), np.uint8))


class PeerShardedUpdate:
    """The sharded update as ONE persistent kernel per rank: the blobs travel through peer mailboxes over NVLink inside
    the kernel (lio_update_enqueue_sharded); no NCCL call and no launch between the passes."""

    def __init__(self, ctx, own):
        self.ctx = ctx
        self.own = (float(own[0]), float(own[1]))

    def update(self, x, P, R=0.001, max_iter=4, extrinsic_est=False):
        c = self.ctx
        c.state_upload(x, P)
        c.update_enqueue_sharded(R, max_iter, extrinsic_est, True, self.own[0], self.own[1])
        return c.state_download()
