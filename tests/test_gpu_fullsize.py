"""Parity at BASELINE.json's full size (config 3): one Ouster OS1-128 scan (128 x 1024 rays) against the 2,000,000-point
city map, the workload bench.py times.  Stage by stage against the oracle (and the reference's own ikd-Tree for the
search), plus size-independent properties: Build -> flatten returns the input set, the update is deterministic, and
restarting it from its own posterior moves nothing."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    from agi_lidar_slam_b200 import synth

    scene, mp = synth.city_map(2_000_000, 3003)
    d, col = synth.spinning_dirs(128, 1024, -22.5, 22.5)
    rng = np.random.default_rng(3003 + 18)
    pos = np.array([rng.uniform(-30, 30), rng.uniform(-30, 30), 2.0])
    R = synth.rot_zyx(rng.uniform(-np.pi, np.pi), rng.normal(0, 0.02), rng.normal(0, 0.02))
    scan = synth.static_scan(scene, d, col / 1024 * 100.0, pos, R, 120.0, 3004)
    x_true = synth.make_state(pos=pos, R=R)
    return dict(map=mp, scan=scan, x_true=x_true, x_prior=synth.perturbed_prior(x_true, 3035), P=synth.init_P())


@pytest.fixture(scope="module")
def big_ctx(big):
    from agi_lidar_slam_b200 import _cabi

    c = _cabi.Context(0, max_scan_points=1 << 17, max_down_points=100000, max_map_points=1 << 21)
    mp = big["map"]
    c.map_build(np.concatenate([mp, np.zeros((len(mp), 1), np.float32)], 1))
    yield c
    c.close()


def test_build_flatten_round_trip_2m(big, big_ctx):
    xyz, ids = big_ctx.map_dump()
    assert len(ids) == 2_000_000 == big_ctx.map_size()[1]
    order = np.argsort(ids)
    assert np.array_equal(ids[order], np.arange(2_000_000, dtype=ids.dtype))
    assert np.array_equal(xyz[order].view(np.uint32), big["map"].view(np.uint32))


def test_os1_128_scan_against_2m_map(big, big_ctx, orc):
    ctx = big_ctx
    scan = big["scan"]
    assert len(scan) > 100_000  # rays without a return within 120 m are dropped by the sensor model
    # --- voxel filter: leaf assignment, M and centroids bit-exact
    body, _, keys = ctx.scan_preprocess(scan, None, None, 0.5, want_keys=True)
    cen, _, okeys = orc.voxel_grid(np.concatenate([scan[:, :3], np.zeros((len(scan), 1), np.float32), scan[:, 3:4]], 1),
                                   0.5)
    assert np.array_equal(keys, okeys)
    assert len(body) == len(cen) and 2000 < len(body) < 100000
    assert np.array_equal(body[:, :3].view(np.uint32), cen[:, :3].view(np.uint32))
    # --- 5-NN of every downsampled point: ids / d2 bit-exact against the oracle map and the reference ikd-Tree
    q = orc.body_to_world(big["x_prior"], body[:, :3])
    idx, d2, nbr = ctx.knn5(q)
    om = orc.Map(1.0)
    om.build(big["map"])
    oi, od, on = om.knn(q, 5, 5.0, threads=16)
    assert np.array_equal(idx, oi) and np.array_equal(d2.view(np.uint32), od.view(np.uint32))
    assert np.array_equal(nbr.view(np.uint32), on.view(np.uint32))
    assert (idx[:, 4] >= 0).mean() > 0.8
    if orc.ikd_available():
        t = orc.IkdTree()
        t.build(big["map"])
        sub = np.arange(0, len(q), 4)
        ti, td, _ = t.knn(q[sub], 5, threads=16)
        td = np.where(td <= 5.0, td, np.inf).astype(np.float32)
        assert np.array_equal(d2[sub].view(np.uint32), td.view(np.uint32))
        assert np.array_equal(idx[sub], np.where(np.isfinite(td), ti, -1))
    # --- the whole update: pass count, matched count, posterior
    x, P, nv, npass = ctx.update_scan(big["x_prior"], big["P"], 0.001, 4, False)
    xr, Pr, trace, nvr = orc.Scan(body[:, :3]).update(big["x_prior"], big["P"], om.knn_backend(), 0.001, 4, False)
    assert npass == len(trace) and nv == nvr and nv > 1000
    dx = orc.boxminus(x, xr)
    assert np.abs(dx[0:3]).max() < 1e-4 and np.abs(dx[3:6]).max() < 1e-4  # north_star: 1e-4 m / 1e-4 rad
    assert np.abs(dx).max() < 1e-8 and np.abs(P - Pr).max() < 1e-10       # what is actually reached
    assert np.linalg.norm(x[0:3] - big["x_true"][0:3]) < 0.02
    # --- properties: deterministic; a restart from the posterior stays put
    x2, P2, nv2, np2 = ctx.update_scan(big["x_prior"], big["P"], 0.001, 4, False)
    assert np.array_equal(x, x2) and np.array_equal(P, P2) and (nv, npass) == (nv2, np2)
    x3, _, _, _ = ctx.update_scan(x, big["P"], 0.001, 4, False)
    d3 = orc.boxminus(x3, x)
    assert np.abs(d3[0:3]).max() < 2e-3 and np.abs(d3[3:6]).max() < 2e-3


def test_full_size_update_with_extrinsic_estimation(big, big_ctx, orc):
    """The same scan with extrinsic_est = true (12 x 12 normal equations, 91 outputs per row): whole update against the
    oracle at full size."""
    ctx = big_ctx
    body, _, _ = ctx.scan_preprocess(big["scan"], None, None, 0.5)
    om = orc.Map(1.0)
    om.build(big["map"])
    x, P, nv, npass = ctx.update_scan(big["x_prior"], big["P"], 0.001, 4, True)
    xr, Pr, trace, nvr = orc.Scan(body[:, :3]).update(big["x_prior"], big["P"], om.knn_backend(), 0.001, 4, True)
    assert npass == len(trace) and nv == nvr and nv > 1000
    dx = orc.boxminus(x, xr)
    assert np.abs(dx[0:3]).max() < 1e-4 and np.abs(dx[3:6]).max() < 1e-4  # north_star
    # (12 x 12 normal equations with the extrinsic rotation nearly degenerate against the body rotation: the covariance
    # agrees to 1e-6 of its largest entry, the tolerance of tests/test_gpu_update.py)
    assert np.abs(dx).max() < 1e-7 and np.abs(P - Pr).max() < 1e-6 * np.abs(Pr).max()
    assert not np.array_equal(x[7:14], big["x_prior"][7:14])  # the extrinsic block took part
    x2, P2, nv2, np2 = ctx.update_scan(big["x_prior"], big["P"], 0.001, 4, True)
    assert np.array_equal(x, x2) and np.array_equal(P, P2) and (nv, npass) == (nv2, np2)


def test_dense_scan_many_tiles_per_block(big, orc):
    """The scan downsampled at 0.15 m: M ~ 40k points (the size SURVEY.md 8d's worked example assumes), i.e. several
    search tiles per block and 8-lane search groups.  Voxel filter bit-exact, update against the oracle."""
    from agi_lidar_slam_b200 import _cabi

    scan, mp = big["scan"], big["map"]
    with _cabi.Context(0, max_scan_points=1 << 17, max_down_points=150000, max_map_points=1 << 21) as ctx:
        ctx.map_build(np.concatenate([mp, np.zeros((len(mp), 1), np.float32)], 1))
        body, _, _ = ctx.scan_preprocess(scan, None, None, 0.15)
        cen, _, _ = orc.voxel_grid(np.concatenate([scan[:, :3], np.zeros((len(scan), 1), np.float32), scan[:, 3:4]], 1),
                                   0.15)
        assert len(body) == len(cen) and len(body) > 30000
        assert np.array_equal(body[:, :3].view(np.uint32), cen[:, :3].view(np.uint32))
        om = orc.Map(1.0)
        om.build(mp)
        for ext in (False, True):
            x, P, nv, npass = ctx.update_scan(big["x_prior"], big["P"], 0.001, 4, ext)
            xr, Pr, trace, nvr = orc.Scan(body[:, :3]).update(big["x_prior"], big["P"], om.knn_backend(), 0.001, 4, ext)
            assert npass == len(trace) and nv == nvr and nv > 10000
            dx = orc.boxminus(x, xr)
            assert np.abs(dx[0:3]).max() < 1e-4 and np.abs(dx[3:6]).max() < 1e-4
            assert np.abs(dx).max() < 1e-7 and np.abs(P - Pr).max() < 1e-6 * np.abs(Pr).max()
