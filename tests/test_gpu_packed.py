"""Packed (ragged) ray layout: per-ray sample counts with an offsets[R+1] array
(BASELINE.json config 4).  The oracle buckets rays by count and runs the dense restatement
of the reference per bucket (per-ray results do not depend on batch composition)."""
import pytest
import torch

import avr_oracle as O
from conftest import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["auto", "generic"])
def family(request):
    """Warp-per-ray kernels (what the library picks for packed rays) AND the thread-per-ray generic kernels."""
    import avr_b200
    lib = avr_b200.load_library()
    lib.avr_set_force_generic(1 if request.param == "generic" else 0)
    yield request.param
    lib.avr_set_force_generic(0)


def _ragged(r, lo, hi, seed, zero_some=True):
    g = torch.Generator().manual_seed(seed)
    counts = torch.randint(lo, hi + 1, (r,), generator=g)
    if zero_some:
        counts[::97] = 0
        counts[1::89] = 1
    offsets = torch.zeros(r + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(counts, 0)
    return counts, offsets, g


def _packed_inputs(counts, offsets, g):
    s = int(offsets[-1])
    d = 0.9 + 0.8 * torch.rand(counts.numel(), generator=g)
    near, far = d - 0.15, d + 0.15
    u = torch.rand(s, generator=g)
    x = torch.cat([torch.sigmoid(torch.randn(s, 3, generator=g)), torch.relu(torch.randn(s, 1, generator=g)) * 30], -1)
    return near, far, u, x


def test_packed_pipeline_vs_oracle(dev, family):
    from avr_b200 import ops
    r = 3000
    counts, offsets, g = _ragged(r, 8, 256, seed=0)
    # a few rays longer than what the warp-per-ray backward keeps in registers (256 samples)
    counts[[7, 1500, 2999]] = torch.tensor([257, 700, 1030])
    offsets[1:] = torch.cumsum(counts, 0)
    near, far, u, x = _packed_inputs(counts, offsets, g)
    od = offsets.to(dev)
    # coarse sampling: each ray stratified over its own count -> bit-exact per bucket
    z = ops.coarse_sample_packed(near.to(dev), far.to(dev), u.to(dev), od).cpu()
    for k, rays, idx in O.bucketed(offsets):
        if k == 0:
            continue
        want = O.coarse_z(near[rays].unsqueeze(0), far[rays].unsqueeze(0), k, u[idx].unsqueeze(0))[0]
        assert torch.equal(z[idx], want), k
    # compositing forward
    xd = x.to(dev).requires_grad_(True)
    zd = z.to(dev).requires_grad_(True)
    rgb, depth, w = ops.composite_packed(xd, zd, od, True, 1.8)
    want_rgb, want_depth, want_w = O.composite_packed(z, x, offsets, True, 1.8)
    assert_close(rgb, want_rgb, what="rgb")
    assert_close(depth, want_depth, what="depth")
    assert_close(w, want_w, what="w")
    empty = counts == 0
    assert (rgb.cpu()[empty] == 1).all() and (depth.cpu()[empty] == 0).all()
    # backward (incl. gradients into weights and z)
    g_rgb, g_d, g_w = torch.randn(r, 3, generator=g), torch.randn(r, generator=g), torch.randn(int(offsets[-1]), generator=g)
    torch.autograd.backward([rgb, depth, w], [g_rgb.to(dev), g_d.to(dev), g_w.to(dev)])
    dx, dz = xd.grad.cpu(), zd.grad.cpu()
    for k, rays, idx in O.bucketed(offsets):
        if k == 0:
            continue
        zz, xx = z[idx].unsqueeze(0), x[idx].unsqueeze(0)
        wdx, _ = O.composite_grads(zz, xx, g_rgb[rays].unsqueeze(0), g_d[rays].reshape(1, -1, 1), g_w[idx].reshape(1, -1, k, 1), True)
        assert_close(dx[idx][..., :3], wdx[0][..., :3], what=f"d_rgb k={k}")
        assert_close(dx[idx][..., 3][:, :-1], wdx[0][..., 3][:, :-1], what=f"d_sigma k={k}")
        assert_close(dx[idx][..., 3][:, -1] / 1e10, wdx[0][..., 3][:, -1] / 1e10, what=f"d_sigma last k={k}")
        _, wdz = O.composite_grads(zz.double(), xx.double(), g_rgb[rays].unsqueeze(0).double(), g_d[rays].reshape(1, -1, 1).double(),
                                   g_w[idx].reshape(1, -1, k, 1).double(), True, want_dz=True)
        scale = wdz.abs().amax(-1, keepdim=True).clamp_min(1.0)
        assert_close(dz[idx] / scale[0], (wdz[0] / scale[0]).float(), rtol=1e-5, atol=2e-5, what=f"d_z k={k}")


@pytest.mark.parametrize("kernels", ["bins", "networks"])
@pytest.mark.parametrize("r,lo,hi", [(2000, 8, 128),     # few rays: one launch at the maximum shape
                                     (6000, 1, 256),     # per-class ragged kernels (8..32 lanes per ray), n = 0 for k = 1
                                     (5000, 8, 300)])    # rays beyond the largest class box take the warp-per-ray classes
def test_packed_importance_and_merge(r, lo, hi, kernels, dev):
    """Both kernel families (bucket ranking, importance_bins.cu / sorting networks).  With the cdf and
    the indices exported: indices bit-exact against searchsorted on the kernel's own cdf, per ray."""
    from avr_b200 import _lib, ops
    _lib.set_option("AVR_IMPORTANCE_BINS", 1 if kernels == "bins" else 0)
    counts, offsets, g = _ragged(r, lo, hi, seed=1, zero_some=False)
    fine_counts = counts // 2
    fine_offsets = torch.zeros(r + 1, dtype=torch.int64)
    fine_offsets[1:] = torch.cumsum(fine_counts, 0)
    near, far, u, _ = _packed_inputs(counts, offsets, g)
    zc = ops.coarse_sample_packed(near.to(dev), far.to(dev), u.to(dev), offsets.to(dev)).cpu()
    w = torch.rand(int(offsets[-1]), generator=g) ** 6
    sf = int(fine_offsets[-1])
    uf, uf2 = torch.rand(sf, generator=g), torch.rand(sf, generator=g)
    args = (w.to(dev), zc.to(dev), near.to(dev), far.to(dev), uf.to(dev), uf2.to(dev), offsets.to(dev), fine_offsets.to(dev),
            hi, hi // 2)
    zf, zs = ops.importance_sample_packed(*args)
    zf2, zs2, kcdf, kidx = ops.importance_sample_packed(*args, want_cdf=True, want_idx=True)
    _lib.set_option("AVR_IMPORTANCE_BINS", None)
    if kernels == "bins" or r < 4096:
        assert torch.equal(zf, zf2) and torch.equal(zs, zs2)      # exporting the cdf changes nothing
    else:
        # the network family serves a cdf request with its warp-per-ray classes (they export the packed cdf), whose
        # scan groups the pdf differently from the 8/16-lane classes: the same rare one-bin flips as against the oracle
        assert int((zf != zf2).sum()) <= max(2, int(2e-5 * sf))
        zf, zs = zf2, zs2                                         # check the run whose cdf / indices we hold
    zf, zs, kcdf, kidx = zf.cpu(), zs.cpu(), kcdf.cpu(), kidx.cpu()
    out_offsets = offsets + fine_offsets
    mismatched = 0
    for k, rays, idx in O.bucketed(offsets):
        n = k // 2
        fidx = fine_offsets[rays].unsqueeze(-1) + torch.arange(n)
        oidx = out_offsets[rays].unsqueeze(-1) + torch.arange(k + n)
        want, cdf, bins = O.fine_z(near[rays].unsqueeze(0), far[rays].unsqueeze(0), w[idx].unsqueeze(0),
                                   uf[fidx].unsqueeze(0), uf2[fidx].unsqueeze(0), return_aux=True)
        same = zf[fidx] == want[0]
        mismatched += int((~same).sum())
        # north_star: indices bit-exact given an identical CDF — the kernel's own table, entry by entry
        cidx = (offsets[rays] + rays).unsqueeze(-1) + torch.arange(k + 1)
        assert torch.equal(kidx[fidx].long(), O.cdf_search(kcdf[cidx], uf[fidx])), k
        assert (kcdf[cidx][:, 0] == 0).all() and (kcdf[cidx][:, 1:] >= kcdf[cidx][:, :-1]).all()
        assert_close(kcdf[cidx], cdf[0], rtol=1e-5, atol=1e-6, what=f"cdf k={k}")
        # the merge is the exact sort of the kernel's own samples
        assert torch.equal(zs[oidx], torch.sort(torch.cat([zc[idx], zf[fidx]], -1), -1).values), k
    assert mismatched <= max(2, int(2e-5 * sf))          # rare one-bin flips from the CDF's summation order


def test_packed_full_size_smoke(dev):
    """2^22 rays with 8..256 samples would need ~11 GB of rgbs; this runs 2^18 rays end to end
    (coarse -> composite -> importance -> merge -> composite) and checks invariants."""
    from avr_b200 import ops
    r = 1 << 18
    g = torch.Generator(device=dev).manual_seed(0)
    counts = torch.randint(8, 257, (r,), device=dev, generator=g)
    offsets = torch.zeros(r + 1, dtype=torch.int64, device=dev)
    offsets[1:] = torch.cumsum(counts, 0)
    s = int(offsets[-1])
    d = 0.9 + 0.8 * torch.rand(r, device=dev, generator=g)
    near, far = d - 0.15, d + 0.15
    z = ops.coarse_sample_packed(near, far, torch.rand(s, device=dev, generator=g), offsets)
    x = torch.cat([torch.sigmoid(torch.randn(s, 3, device=dev, generator=g)),
                   torch.relu(torch.randn(s, 1, device=dev, generator=g)) * 30], -1)
    rgb, depth, w = ops.composite_packed(x, z, offsets, True, 1.8)
    seg = torch.repeat_interleave(torch.arange(r, device=dev), counts)
    acc = torch.zeros(r, device=dev).index_add_(0, seg, w)
    assert (w >= 0).all() and acc.max() <= 1 + 1e-6
    fine_offsets = torch.zeros(r + 1, dtype=torch.int64, device=dev)
    fine_offsets[1:] = torch.cumsum(counts // 2, 0)
    sf = int(fine_offsets[-1])
    zf, zs = ops.importance_sample_packed(w, z, near, far, torch.rand(sf, device=dev, generator=g),
                                          torch.rand(sf, device=dev, generator=g), offsets, fine_offsets, 256, 128)
    seg2 = torch.repeat_interleave(torch.arange(r, device=dev), counts + counts // 2)
    same_ray = seg2[1:] == seg2[:-1]
    assert (zs[1:][same_ray] >= zs[:-1][same_ray]).all()
    assert abs(zs.double().sum().item() - (z.double().sum() + zf.double().sum()).item()) < 1e-3


@pytest.mark.parametrize("name", ["medium", "short_only", "capacity_edges", "single_ray", "few_rays", "with_empty_tail",
                                  "misaligned_end", "rays_across_tiles", "thirty_two_ends", "long_short_mix"])
def test_packed_span_tiling_matches_generic(name, dev):
    """The dynamically tiled span kernels against the thread-per-ray generic kernels on count
    distributions that hit every tiling decision (tile / too short / too long / partial quads)."""
    import avr_b200
    from avr_b200 import ops
    lib = avr_b200.load_library()
    g = torch.Generator().manual_seed(hash(name) % 1000)
    counts = {
        "medium": torch.randint(14, 300, (4000,), generator=g),
        "short_only": torch.randint(0, 14, (3000,), generator=g),
        "capacity_edges": torch.tensor([416, 417, 415, 14, 13, 402, 14, 208, 208, 1, 416] * 50),
        "single_ray": torch.tensor([97]),
        "few_rays": torch.tensor([20, 3, 150, 416, 33]),
        "with_empty_tail": torch.cat([torch.randint(14, 200, (500,), generator=g), torch.zeros(40, dtype=torch.int64)]),
        "misaligned_end": torch.tensor([15, 17, 19, 21, 23, 30, 14, 15]),   # S % 4 != 0, tiles end inside the last quad
        # the streaming forward kernel: rays that continue over two, three, a dozen tiles (the carried transmittance
        # and sums), tiles that end at the 32nd ray end, open rays cut off by short and empty rays
        "rays_across_tiles": torch.tensor([1030, 4097, 700, 20, 5000, 14, 416, 417, 832, 833, 1248, 415, 430] * 7),
        "thirty_two_ends": torch.randint(14, 21, (2500,), generator=g),
        "long_short_mix": torch.tensor([900, 3, 900, 0, 13, 500, 14, 1, 2000, 0, 0, 417, 12, 416] * 20),
    }[name]
    r = counts.numel()
    offsets = torch.zeros(r + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(counts, 0)
    s = int(offsets[-1])
    z = torch.zeros(s)
    for k, rays, idx in O.bucketed(offsets):
        if k:
            z[idx] = torch.sort(0.8 + torch.rand(rays.numel(), k, generator=g), -1).values
    x = torch.cat([torch.sigmoid(torch.randn(s, 3, generator=g)), torch.relu(torch.randn(s, 1, generator=g)) * 30], -1)
    g_rgb, g_d = torch.randn(r, 3, generator=g), torch.randn(r, generator=g)
    res = {}
    for fam in (0, 1):
        lib.avr_set_force_generic(fam)
        try:
            xd = x.to(dev).requires_grad_(True)
            rgb, depth, w = ops.composite_packed(xd, z.to(dev), offsets.to(dev), True, 1.8)
            rgb2, depth2, none = ops.composite_packed(x.to(dev), z.to(dev), offsets.to(dev), False, 1.8, want_w=False)
            torch.autograd.backward([rgb, depth], [g_rgb.to(dev), g_d.to(dev)])
            res[fam] = (rgb.detach().cpu(), depth.detach().cpu(), w.detach().cpu(), xd.grad.cpu(), rgb2.cpu(), depth2.cpu())
        finally:
            lib.avr_set_force_generic(0)
    a, b = res[0], res[1]
    # two fp32 summation orders of the same ray: the bound grows with the ray (the sequential kernel adds up to 5000
    # terms one by one in the long-ray cases); a lost or doubled carry between tiles would be an error of order one
    tight = dict(rtol=3e-6, atol=3e-7) if int(counts.max()) <= 512 else dict(rtol=1e-5, atol=1e-6)
    for i, what in enumerate(["rgb", "depth", "w"]):
        assert_close(a[i], b[i], what=what, **tight)
    assert_close(a[3][:, :3], b[3][:, :3], what="d_rgb", **tight)
    last = offsets[1:][counts > 0] - 1
    mask = torch.ones(s, dtype=torch.bool)
    mask[last] = False
    assert_close(a[3][mask, 3], b[3][mask, 3], rtol=1e-5, atol=1e-6, what="d_sigma")
    assert_close(a[3][last, 3] / 1e10, b[3][last, 3] / 1e10, rtol=1e-5, atol=1e-6, what="d_sigma last / 1e10")
    assert_close(a[4], b[4], what="rgb (no white background, no weights)", **tight)
    assert_close(a[5], b[5], what="depth (no weights)", **tight)
    want = O.composite_packed(z, x, offsets, True, 1.8)
    assert_close(a[0], want[0], what="rgb vs oracle")
    assert_close(a[2], want[2], what="w vs oracle")
    # against the same computation in fp64, the span kernels must not be further off than the sequential ones by
    # more than rounding (this is what separates an ordering difference from a wrong carry)
    want64 = O.composite_packed(z.double(), x.double(), offsets, True, 1.8)
    err_span = (a[0].double() - want64[0]).abs().max().item()
    err_seq = (b[0].double() - want64[0]).abs().max().item()
    assert err_span <= max(4.0 * err_seq, 2e-6), (err_span, err_seq)


@pytest.mark.parametrize("variant", ["flat", "ray"])
@pytest.mark.parametrize("case", ["c4_like", "empty_and_single", "empty_segments", "long", "scalar_bounds", "unaligned"])
def test_packed_coarse_sampler_bit_exact(case, variant, dev):
    """Packed coarse sampler (segment-of-32-rays flat kernel and the one-warp-per-ray kernel) against
    the oracle per bucket, bit for bit: ragged counts, empty rays, whole empty segments, rays longer
    than a warp step, scalar bounds, and u/z views off the 16-byte grid (the CPU suite walks the
    same per-lane code on the host: tests/test_host_kernel_cores.py)."""
    from avr_b200 import _lib, ops
    _lib.set_option("AVR_COARSE_PACKED", 1 if variant == "ray" else 0)
    g = torch.Generator().manual_seed(3)
    counts = {
        "c4_like": lambda: torch.randint(8, 257, (4099,), generator=g),
        "empty_and_single": lambda: torch.tensor([0, 0, 1, 2, 3, 5, 31, 32, 33, 64, 200])[torch.randint(0, 11, (1500,), generator=g)],
        "empty_segments": lambda: torch.cat([torch.zeros(70, dtype=torch.int64), torch.tensor([5]), torch.zeros(40, dtype=torch.int64), torch.tensor([1, 0, 0, 9])]),
        "long": lambda: torch.tensor([1030, 0, 4097, 7, 700, 257]),
        "scalar_bounds": lambda: torch.randint(0, 300, (777,), generator=g),
        "unaligned": lambda: torch.randint(1, 12, (333,), generator=g) * 2 + 1,
    }[case]().to(torch.int64)
    r = counts.numel()
    offsets = torch.zeros(r + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(counts, 0)
    s = int(offsets[-1])
    if case == "scalar_bounds":
        near, far = torch.tensor([0.8]), torch.tensor([1.8])
    else:
        d = 0.9 + 0.8 * torch.rand(r, generator=g)
        near, far = d - 0.15, d + 0.15
    u = torch.rand(s, generator=g)
    u[::13] = torch.tensor([0.0, 2.0 ** -24, 1 - 2.0 ** -24, 1e-30, 1e-38, 1e-44])[torch.arange(u[::13].numel()) % 6]
    ud = u.to(dev)
    if case == "unaligned":
        buf = torch.zeros(s + 1, device=dev)
        buf[1:] = ud
        ud = buf[1:]                                     # 4 bytes off the 16-byte grid
        assert ud.data_ptr() % 16 == 4
    z = ops.coarse_sample_packed(near.to(dev), far.to(dev), ud, offsets.to(dev)).cpu()
    for k, rays, idx in O.bucketed(offsets):
        if k == 0:
            continue
        n = near[rays] if near.numel() > 1 else near.expand(len(rays))
        f = far[rays] if far.numel() > 1 else far.expand(len(rays))
        want = O.coarse_z(n.unsqueeze(0), f.unsqueeze(0), k, u[idx].unsqueeze(0))[0]
        assert torch.equal(z[idx], want), (case, variant, k)
    _lib.set_option("AVR_COARSE_PACKED", None)
