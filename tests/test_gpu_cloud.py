"""GPU tests of the PointCloud2 path (north-star extensions, PARITY UNPINNED: the reference has
no such code; oracle/cloud_oracle.cpp is the self-authored definition).  Every step of the
definition is order independent, so the CUDA result is compared bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    import rplidar_ros2_driver_b200 as R

    return R


@pytest.fixture(scope="module")
def ctx(R):
    c = R.Context(0, 40000, 64)
    yield c
    c.close()


def room_scans(oracle, n_scans, n, seed, ties=False):
    """A square room seen from an off-centre sensor, + noise, 5% unmeasured, rotated start."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n_scans, n), oracle.NODE_DTYPE)
    for s in range(n_scans):
        keys = np.sort(rng.choice(65536, size=n, replace=ties))
        th = keys * (2 * np.pi / 65536)
        ox, oy = rng.uniform(-1, 1, 2)
        half = 3.0
        with np.errstate(divide="ignore"):
            tx = np.where(np.cos(th) > 0, (half - ox) / np.cos(th), (-half - ox) / np.cos(th))
            ty = np.where(np.sin(th) > 0, (half - oy) / np.sin(th), (-half - oy) / np.sin(th))
        r = np.minimum(np.abs(tx), np.abs(ty)) + rng.normal(0, 0.004, n)
        r[rng.random(n) < 0.01] += rng.uniform(0.3, 1.5)  # outliers
        dist = np.clip(r * 4000.0, 1, 2**31).astype(np.uint32)
        dist[rng.random(n) < 0.05] = 0
        nodes = oracle.make_nodes(keys, dist, rng.integers(0, 256, n), 2)
        out[s] = np.roll(nodes, -int(rng.integers(0, n)))
    return out


def check_cloud(R, O, ctx, nodes, counts, **kw):
    """Both implementations against the definition: flags 0 = the shared-memory kernel with SOR / voxel grid
    fused (revolutions of at most 4096 nodes; larger ones take the TMA kernel + post passes either way),
    CLOUD_NO_FUSED = scan kernel + separate post passes."""
    counts = np.asarray(counts, np.uint32)
    exp = [O.cloud(nodes[s, : counts[s]], O.cloud_params(**kw)) for s in range(nodes.shape[0])]
    for flags in (0, R.CLOUD_NO_FUSED):
        xyzi, pc = ctx.cloud_batch(nodes.view(R.NODE_DTYPE), counts, R.cloud_params(flags=flags, **kw))
        for s in range(nodes.shape[0]):
            assert pc[s] == exp[s].shape[0], (s, kw, flags, pc[s], exp[s].shape[0])
            assert (xyzi[s, : pc[s]].view(np.uint32) == exp[s].view(np.uint32)).all(), (s, kw, flags)
    return xyzi, pc


@pytest.mark.parametrize("n", [1, 2, 33, 360, 3200, 32768])
def test_polar_to_xyz_window_bit_exact(R, oracle, ctx, n):
    nodes = oracle.synth_batch(700 + n, 4, n, 1)
    counts = np.full(4, n, np.uint32)
    check_cloud(R, oracle, ctx, nodes, counts, range_min=0.0, range_max=1e9)
    check_cloud(R, oracle, ctx, nodes, counts, range_min=0.15, range_max=25.0, intensity_min=20.0, is_new_protocol=0)
    check_cloud(R, oracle, ctx, nodes, counts, range_min=5.0, range_max=12.0, intensity_min=100.0, is_new_protocol=1)


def test_cloud_with_duplicate_keys_uses_stable_order(R, oracle, ctx):
    nodes = oracle.synth_batch(31, 3, 2048, 2)
    check_cloud(R, oracle, ctx, nodes, np.full(3, 2048, np.uint32), range_min=0.15, range_max=40.0)


@pytest.mark.parametrize("n", [5, 30, 34, 3200])
def test_statistical_outlier_removal(R, oracle, ctx, n):
    nodes = room_scans(oracle, 6, n, 11 + n)
    counts = np.full(6, n, np.uint32)
    _, pc0 = check_cloud(R, oracle, ctx, nodes, counts, range_min=0.15, range_max=40.0)
    _, pc = check_cloud(R, oracle, ctx, nodes, counts, range_min=0.15, range_max=40.0, sor_k=8, sor_alpha=1.0)
    if n >= 3200:
        assert (pc < pc0).all() and (pc > 0.8 * pc0).all()  # outliers go, walls stay
    check_cloud(R, oracle, ctx, nodes, counts, range_min=0.15, range_max=40.0, sor_k=3, sor_alpha=0.0)
    check_cloud(R, oracle, ctx, nodes, counts, range_min=0.15, range_max=40.0, sor_k=32, sor_alpha=2.5)


def test_voxel_cells_around_the_origin(R, oracle, ctx):
    """Voxels larger than range_min: points fall into the cells (-1,-1), (-1,0), (0,-1), (0,0) around the
    sensor (cell (-1,-1) has the all-ones key: it must not be mistaken for an empty table slot)."""
    nodes = oracle.synth_batch(50, 6, 3200, variant=0)
    nodes["dist_mm_q2"] = (nodes["dist_mm_q2"] % 3000) + 40  # 1 cm .. 76 cm
    counts = np.full(6, 3200, np.uint32)
    for voxel in (1.0, 0.4):
        _, pc = check_cloud(R, oracle, ctx, nodes, counts, range_min=0.0, range_max=40.0, voxel_size=voxel)
        assert (pc <= 16).all() and (pc >= 4).all()


@pytest.mark.parametrize("voxel", [0.05, 0.5])
def test_voxel_grid(R, oracle, ctx, voxel):
    nodes = room_scans(oracle, 6, 3200, 5)
    counts = np.full(6, 3200, np.uint32)
    _, pc = check_cloud(R, oracle, ctx, nodes, counts, range_min=0.15, range_max=40.0, voxel_size=voxel)
    assert (pc > 0).all() and (pc < 3200).all()
    check_cloud(R, oracle, ctx, nodes, counts, range_min=0.15, range_max=40.0, voxel_size=voxel, sor_k=8, sor_alpha=1.0)


def test_full_chain_on_shuffled_scans(R, oracle, ctx):
    """BASELINE.json configs[3]: angle order forced by the kernel (variant 3 = shuffled), range
    and intensity window, SOR k=8 alpha=1, 5 cm voxels."""
    nodes = oracle.synth_batch(2, 4, 8192, 3)
    check_cloud(R, oracle, ctx, nodes, np.full(4, 8192, np.uint32), range_min=0.15, range_max=40.0,
                intensity_min=10.0, sor_k=8, sor_alpha=1.0, voxel_size=0.05)


def test_ragged_and_empty_scans(R, oracle, ctx):
    stride = 400
    counts = np.array([0, 1, 399, 400, 7], np.uint32)
    nodes = np.zeros((5, stride), oracle.NODE_DTYPE)
    src = room_scans(oracle, 5, stride, 77)
    for s, n in enumerate(counts):
        nodes[s, :n] = src[s, :n]
    nodes[3]["dist_mm_q2"] = 0  # nothing measured
    check_cloud(R, oracle, ctx, nodes, counts, range_min=0.15, range_max=40.0, sor_k=8, sor_alpha=1.0, voxel_size=0.05)


def test_fuse_packs_per_scan_clouds(R, oracle, ctx):
    import torch

    n_scans, n = 16, 3200
    host = room_scans(oracle, n_scans, n, 3)
    dev = torch.device("cuda")
    nodes = torch.from_numpy(host.view(np.uint8).reshape(n_scans, n, 8)).to(dev)
    counts = torch.full((n_scans,), n, dtype=torch.int32, device=dev)
    xyzi = torch.zeros((n_scans, n, 4), dtype=torch.float32, device=dev)
    pc = torch.zeros(n_scans, dtype=torch.int32, device=dev)
    fused = torch.zeros((n_scans * n, 4), dtype=torch.float32, device=dev)
    offs = torch.zeros(n_scans, dtype=torch.int32, device=dev)
    total = torch.zeros(1, dtype=torch.int32, device=dev)
    prm = R.cloud_params(range_min=0.15, range_max=40.0, voxel_size=0.05)
    torch.cuda.synchronize()  # buffers were filled on torch's stream; the library runs on its own
    ctx.cloud_batch_dev(nodes.data_ptr(), counts.data_ptr(), n_scans, n, prm, xyzi.data_ptr(), pc.data_ptr())
    torch.cuda.synchronize()  # buffers were filled on torch's stream; the library runs on its own
    ctx.cloud_fuse_dev(xyzi.data_ptr(), pc.data_ptr(), n_scans, n, fused.data_ptr(), offs.data_ptr(), total.data_ptr())
    ctx.synchronize()
    torch.cuda.synchronize()
    pcs = pc.cpu().numpy()
    assert int(total.item()) == int(pcs.sum())
    assert (offs.cpu().numpy() == np.concatenate([[0], np.cumsum(pcs)[:-1]])).all()
    exp = np.concatenate([oracle.cloud(host[s], oracle.cloud_params(range_min=0.15, range_max=40.0, voxel_size=0.05))
                          for s in range(n_scans)])
    got = fused[: int(total.item())].cpu().numpy()
    assert (got.view(np.uint32) == exp.view(np.uint32)).all()


@pytest.mark.parametrize("stride,n", [(64, 64), (65, 61), (1000, 999), (3200, 3200), (4096, 4096), (4095, 4001)])
def test_shared_memory_chain_strides(R, oracle, ctx, stride, n):
    """The fused chain at the edges of its range: smallest and largest capacity, odd strides (no TMA staging:
    scan bases are only 8-byte aligned), counts below the stride."""
    src = room_scans(oracle, 5, n, 1000 + stride)
    nodes = np.zeros((5, stride), oracle.NODE_DTYPE)
    nodes[:, :n] = src
    counts = np.array([n, n - 1, n, max(n // 2, 1), n], np.uint32)
    for kw in (dict(), dict(voxel_size=0.05), dict(sor_k=8, sor_alpha=1.0), dict(sor_k=8, sor_alpha=1.0, voxel_size=0.05),
               dict(sor_k=20, sor_alpha=0.5, voxel_size=0.3)):
        check_cloud(R, oracle, ctx, nodes, counts, range_min=0.15, range_max=40.0, **kw)


def test_shared_memory_chain_ties_and_dense_cells(R, oracle, ctx):
    """Duplicate keys inside a batch (those scans go through the general kernel and the list-restricted post
    passes), every point in one voxel, every point its own voxel."""
    n = 2048
    nodes = room_scans(oracle, 6, n, 4242)
    nodes[1] = room_scans(oracle, 1, n, 4243, ties=True)[0]
    nodes[4] = room_scans(oracle, 1, n, 4244, ties=True)[0]
    counts = np.full(6, n, np.uint32)
    check_cloud(R, oracle, ctx, nodes, counts, range_min=0.15, range_max=40.0, sor_k=8, sor_alpha=1.0, voxel_size=0.05)
    one = oracle.synth_batch(77, 3, 4096, 0)
    one["dist_mm_q2"] = np.where(one["dist_mm_q2"] != 0, 2 + (one["dist_mm_q2"] % 30), 0)  # 0.5 .. 8 mm from the sensor
    check_cloud(R, oracle, ctx, one, np.full(3, 4096, np.uint32), range_min=0.0, range_max=40.0, voxel_size=3.9)
    far = oracle.synth_batch(78, 3, 4096, 0)
    far["dist_mm_q2"] = np.where(far["dist_mm_q2"] != 0, 150000 + (far["dist_mm_q2"] % 8000), 0)  # 37.5 .. 39.5 m
    _, pc = check_cloud(R, oracle, ctx, far, np.full(3, 4096, np.uint32), range_min=0.15, range_max=40.0, voxel_size=0.002)
    assert (pc > 3500).all()


def test_large_voxels_take_the_separate_passes(R, oracle, ctx):
    """voxel > 4 m or more than 32000 cells per axis: the 32-bit cell keys / accumulators of the fused kernel
    are not exact any more, the library must fall back to the separate passes (same results)."""
    nodes = room_scans(oracle, 4, 3200, 99)
    counts = np.full(4, 3200, np.uint32)
    check_cloud(R, oracle, ctx, nodes, counts, range_min=0.15, range_max=40.0, voxel_size=6.0)
    check_cloud(R, oracle, ctx, nodes, counts, range_min=0.15, range_max=900.0, voxel_size=0.001)


@pytest.mark.parametrize("variant,n", [(0, 3200), (4, 3200), (1, 32768)])
def test_cuda_projection_within_1e6_of_float64(R, oracle, ctx, variant, n):
    """The CUDA path against float64 numpy directly (no oracle in between): BASELINE.json's 1e-6 relative
    tolerance on the polar -> Cartesian path, laser_geometry::projectLaser semantics (tests/test_cloud_semantics.py)."""
    from test_cloud_semantics import check_projection

    nodes = oracle.synth_batch(9100 + variant, 3, n, variant)
    counts = np.full(3, n, np.uint32)
    xyzi, pc = ctx.cloud_batch(nodes.view(R.NODE_DTYPE), counts, R.cloud_params(range_min=0.15, range_max=40.0))
    for s in range(3):
        assert check_projection(xyzi[s, : pc[s]], nodes[s], range_min=0.15, range_max=40.0) < 1e-6
