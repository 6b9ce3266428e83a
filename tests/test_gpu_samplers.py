"""Parity of the sampling kernels: coarse depths bit-exact, CDF indices bit-exact given the
kernel's own CDF, fine depths bit-exact given equal indices, merge = exact sort."""
import pytest
import torch

import avr_oracle as O
from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["auto", "bins", "warp", "generic"])
def family(request):
    """Run through the kernel the library picks (8/16-lanes-per-ray sorting networks for the hot
    shapes), the bucket-ranking kernel (importance_bins.cu, opt-in), the warp-per-ray register
    kernel, AND the general shared-memory kernel."""
    import avr_b200
    from avr_b200 import _lib
    lib = avr_b200.load_library()
    lib.avr_set_force_generic(1 if request.param == "generic" else 0)
    if request.param == "warp":
        _lib.set_option("AVR_IMPORTANCE_GRP", 0)
        _lib.set_option("AVR_IMPORTANCE_BINS", 0)
    if request.param == "bins":
        _lib.set_option("AVR_IMPORTANCE_BINS", 1)
    yield request.param
    _lib.set_option("AVR_IMPORTANCE_GRP", None)
    _lib.set_option("AVR_IMPORTANCE_BINS", None)
    lib.avr_set_force_generic(0)


def test_coarse_golden_bit_exact(dev):
    import avr_b200
    g = load_golden("coarse")
    near = torch.tensor([0.8], device=dev).expand(2, 37)
    far = torch.tensor([1.8], device=dev).expand(2, 37)
    for k in (64, 20, 1):
        z = avr_b200.sample_coarse(near, far, k, device=dev, u=g[f"u_k{k}"].to(dev))
        assert torch.equal(z.cpu(), g[f"ref_z_k{k}"]), k
    d = g["avr_d"].to(dev)
    z = avr_b200.sample_coarse(d - 0.15, d + 0.15, 20, device=dev, u=g["avr_u"].to(dev))
    assert torch.equal(z.cpu(), g["ref_avr_z"])


def test_coarse_draws_inside_like_reference(dev):
    import avr_b200
    near = torch.tensor([0.8], device=dev).expand(3, 50)
    far = torch.tensor([1.8], device=dev).expand(3, 50)
    torch.manual_seed(4)
    z = avr_b200.sample_coarse(near, far, 64, device=dev)
    torch.manual_seed(4)
    u = torch.rand_like(torch.empty(3, 50, 64, device=dev))   # the reference's draw call (renderers.py:14)
    assert torch.equal(z, avr_b200.sample_coarse(near, far, 64, device=dev, u=u))
    assert (z[..., 1:] > z[..., :-1]).all() and z.min() >= 0.8 and z.max() < 1.8 + 1e-6
    zi = avr_b200.sample_coarse(near, far, 64, device=dev, infinity=1.8, u=u)
    assert torch.equal(zi[..., :-1], z[..., 1:]) and (zi[..., -1] == 1.8).all()


def test_coarse_backward_matches_autograd(dev):
    import avr_b200
    g = torch.Generator().manual_seed(2)
    d = (0.9 + 0.8 * torch.rand(2, 300, generator=g))
    u = torch.rand(2, 300, 20, generator=g)
    gz = torch.randn(2, 300, 20, generator=g)
    near = (d - 0.15).requires_grad_(True)
    far = (d + 0.15).requires_grad_(True)
    O.coarse_z(near, far, 20, u).backward(gz)
    n2 = (d - 0.15).to(dev).requires_grad_(True)
    f2 = (d + 0.15).to(dev).requires_grad_(True)
    avr_b200.sample_coarse(n2, f2, 20, device=dev, u=u.to(dev)).backward(gz.to(dev))
    assert_close(n2.grad, near.grad, rtol=1e-5, atol=2e-6, what="d_near")
    assert_close(f2.grad, far.grad, rtol=1e-5, atol=2e-6, what="d_far")
    # gradient through a shared distance d (how AdaptiveVolumeRenderer uses it)
    dd = d.clone().to(dev).requires_grad_(True)
    avr_b200.sample_coarse(dd - 0.15, dd + 0.15, 20, device=dev, u=u.to(dev)).backward(gz.to(dev))
    assert_close(dd.grad, gz.sum(-1), rtol=1e-5, atol=5e-6, what="d_distance")


def _check_importance(w, near, far, u, u2, dev, ref_z=None, min_same=0.9995):
    from avr_b200 import ops
    kc = w.shape[-2] if w.dim() == 4 else w.shape[-1]
    out = ops.importance_sample(w.to(dev), near.to(dev), far.to(dev), u.to(dev), u2.to(dev),
                                want_fine=True, want_cdf=True, want_idx=True)
    cdf, idx, zf = out["cdf"].cpu(), out["idx"].cpu().long(), out["z_fine"].cpu()
    # the CDF itself: same function as the reference's up to summation order, and monotone
    ocdf = O.cdf_from_weights(w.squeeze(-1) if w.dim() == 4 else w)
    assert_close(cdf, ocdf, rtol=1e-6, atol=5e-7, what="cdf")
    assert (cdf[..., 1:] >= cdf[..., :-1]).all() and (cdf[..., 0] == 0).all()
    # indices: BIT-EXACT with torch's searchsorted run on the kernel's CDF (renderers.py:42-43)
    want_idx = O.cdf_search(cdf, u)
    assert torch.equal(idx, want_idx)
    assert idx.min() >= 0 and idx.max() <= kc
    # depths: bit-exact given the indices (renderers.py:45-46)
    t = (idx.float() + u2) / kc
    want_z = near.unsqueeze(-1) + torch.einsum("bs,bsj->bsj", far - near, t)
    assert torch.equal(zf, want_z)
    if ref_z is not None:
        # end to end against the reference: identical wherever its own CDF gave the same bin
        ridx = O.cdf_search(ocdf, u)
        same = ridx == idx
        assert same.float().mean() > min_same
        assert torch.equal(zf[same], ref_z[same])
    return out


def test_importance_golden(dev, family):
    g = load_golden("fine")
    r = g["w"].shape[1]
    near = torch.tensor([0.8]).expand(1, r).contiguous()
    far = torch.tensor([1.8]).expand(1, r).contiguous()
    _check_importance(g["w"], near, far, g["u"], g["u2"], dev, g["ref_z"])
    # adversarial draws sit exactly on CDF edges (u = 1 - 2^-24 vs a cdf[-1] that rounds to
    # 1 -/+ 1 ulp), where the summation order of the CDF decides the bin: flips are expected there
    out = _check_importance(g["w"], near, far, g["u_adv"], g["u2_adv"], dev, g["ref_z_adv"], min_same=0.99)
    idx = out["idx"].cpu()
    assert (idx[..., 0] == 0).all() and (idx[..., 2] >= 63).all()
    d = g["pr_d"]
    _check_importance(g["w"], d - 0.15, d + 0.15, g["pr_u"], g["pr_u2"], dev, g["ref_pr_z"])


def test_sample_fine_signature_and_seed(dev):
    import avr_b200
    g = load_golden("fine")
    r = g["w"].shape[1]
    near = torch.tensor([0.8], device=dev).expand(1, r)
    far = torch.tensor([1.8], device=dev).expand(1, r)
    torch.manual_seed(11)
    z = avr_b200.sample_fine(near, far, 128, g["w"].to(dev), device=dev)
    torch.manual_seed(11)
    u = torch.rand(1, r, 128, device=dev)
    u2 = torch.rand_like(u)
    assert torch.equal(z, avr_b200.sample_fine(near, far, 128, g["w"].to(dev), device=dev, u=u, u2=u2))
    assert z.shape == (1, r, 128)


@pytest.mark.parametrize("kc,n,nd", [(64, 16, 16), (64, 128, 0), (32, 8, 8), (20, 5, 3), (200, 300, 12), (1, 1, 1),
                                     (128, 64, 0), (96, 250, 6), (300, 100, 20), (33, 31, 0), (64, 0, 16), (32, 8, 8)])
def test_merge_is_exact_sort(kc, n, nd, dev, family):
    from avr_b200 import ops
    g = torch.Generator().manual_seed(kc * 7 + n)
    r = 257
    near = torch.tensor([0.8]).expand(1, r).contiguous()
    far = torch.tensor([1.8]).expand(1, r).contiguous()
    w = torch.rand(1, r, kc, generator=g) ** 6
    u, u2 = torch.rand(1, r, n, generator=g), torch.rand(1, r, n, generator=g)
    normals = torch.randn(1, r, nd, generator=g) if nd else None
    zc = O.coarse_z(near, far, kc, torch.rand(1, r, kc, generator=g))
    if kc >= 20:
        # arbitrary caller input: a few rays whose coarse depths are NOT ascending
        zc[0, 5] = zc[0, 5].flip(-1)
        zc[0, 77, [3, 11]] = zc[0, 77, [11, 3]]
        zc[0, 200] = zc[0, 200][torch.randperm(kc, generator=g)]
    std = 50.0 if kc == 32 else 0.01        # std=50 puts the clamped depth samples all over [near, far]
    out = ops.importance_sample(w.to(dev), near.to(dev), far.to(dev), u.to(dev), u2.to(dev), z_coarse=zc.to(dev),
                                normals=None if normals is None else normals.to(dev), depth_std=std,
                                want_fine=True, want_sorted=True)
    zf = out["z_fine"].cpu()
    parts = [zc, zf] + ([O.depth_z(normals, std, torch.tensor([0.8]), torch.tensor([1.8]))] if nd else [])
    assert torch.equal(out["z_sorted"].cpu(), O.merge_sorted(*parts))


def test_sort_rays_and_grad(dev):
    from avr_b200 import ops
    g = torch.Generator().manual_seed(5)
    for k in (1, 20, 33, 96, 1024):
        z = torch.rand(3, 40, k, generator=g)
        z[0, 0, : k // 2] = 0.5                                  # ties: stable order expected
        z[1] = torch.sort(z[1], -1).values                       # already ascending rays (the adaptive renderer's case)
        z[2, ::3] = torch.sort(z[2, ::3], -1).values             # sorted and unsorted rays mixed inside a warp
        if k >= 8:
            z[1, 7, 3:6] = z[1, 7, 3]                            # ascending with ties
        zd = z.to(dev).requires_grad_(True)
        out, perm = ops.SortRays.apply(zd)
        want = torch.sort(z, dim=-1, stable=True)
        assert torch.equal(out.cpu(), want.values) and torch.equal(perm.cpu().long(), want.indices)
        gz = torch.randn(3, 40, k, generator=g)
        out.backward(gz.to(dev))
        zz = z.clone().requires_grad_(True)
        torch.sort(zz, dim=-1, stable=True).values.backward(gz)
        assert torch.equal(zd.grad.cpu(), zz.grad)


def test_importance_full_size_properties(dev):
    """BASELINE.json config 3: 2^20 rays, 64 coarse weights -> 128 fine samples."""
    from avr_b200 import ops
    r, kc, n = 1 << 20, 64, 128
    g = torch.Generator(device=dev).manual_seed(0)
    w = torch.rand(1, r, kc, device=dev, generator=g) ** 6
    u = torch.rand(1, r, n, device=dev, generator=g)
    u2 = torch.rand(1, r, n, device=dev, generator=g)
    u[0, :, 0] = 0.0
    u[0, :, 1] = 2.0 ** -24
    u[0, :, 2] = 1.0 - 2.0 ** -24
    near = torch.tensor([0.8], device=dev)
    far = torch.tensor([1.8], device=dev)
    zc = ops.coarse_sample_raw(near, far, 0, torch.rand(1, r, kc, device=dev, generator=g))
    out = ops.importance_sample(w, near, far, u, u2, z_coarse=zc, want_fine=True, want_sorted=True,
                                want_cdf=True, want_idx=True)
    cdf, idx = out["cdf"], out["idx"].long()
    assert torch.equal(idx, torch.clamp_min(torch.searchsorted(cdf, u, right=True) - 1, 0))
    assert (idx[..., 0] == 0).all() and idx.max() <= kc
    zs = out["z_sorted"]
    assert (zs[..., 1:] >= zs[..., :-1]).all()
    # exactly the sorted multiset of the inputs
    assert torch.equal(zs, torch.sort(torch.cat([zc, out["z_fine"]], -1), -1).values)


@pytest.mark.parametrize("kc,n,nd", [(64, 128, 0), (64, 16, 16), (32, 8, 8)])
def test_importance_unaligned_inputs_take_the_scalar_paths(kc, n, nd, dev):
    """The hot dense shapes move their rows with 16-byte accesses when the pointers allow it; inputs that
    start 4 bytes off a 16-byte boundary take the scalar loads / stores (and the lane-blocked split of the
    draws) and must give the same bits."""
    from avr_b200 import ops
    g = torch.Generator().manual_seed(kc + n)
    r = 301
    near = torch.tensor([0.8], device=dev)
    far = torch.tensor([1.8], device=dev)

    def off(t):  # the same values in a contiguous tensor whose storage starts one float later
        flat = torch.empty(t.numel() + 1, dtype=t.dtype, device=dev)
        flat[1:] = t.to(dev).reshape(-1)
        v = flat[1:].view(t.shape)
        assert v.data_ptr() % 16 == 4 and v.is_contiguous()
        return v

    w = torch.rand(1, r, kc, generator=g) ** 6
    u, u2 = torch.rand(1, r, n, generator=g), torch.rand(1, r, n, generator=g)
    nrm = torch.randn(1, r, nd, generator=g) if nd else None
    zc = O.coarse_z(near.cpu().expand(1, r).contiguous(), far.cpu().expand(1, r).contiguous(), kc, torch.rand(1, r, kc, generator=g))
    kw = dict(depth_std=0.01, want_fine=True, want_sorted=True, want_cdf=True, want_idx=True)
    ref = ops.importance_sample(w.to(dev), near, far, u.to(dev), u2.to(dev), z_coarse=zc.to(dev),
                                normals=None if nrm is None else nrm.to(dev), **kw)
    for which in ("w", "u", "z", "all"):
        a = ops.importance_sample(off(w) if which in ("w", "all") else w.to(dev), near, far,
                                  off(u) if which in ("u", "all") else u.to(dev), off(u2) if which in ("u", "all") else u2.to(dev),
                                  z_coarse=off(zc) if which in ("z", "all") else zc.to(dev),
                                  normals=None if nrm is None else (off(nrm) if which in ("z", "all") else nrm.to(dev)), **kw)
        for k in ("z_sorted", "z_fine", "idx", "cdf"):
            assert torch.equal(a[k], ref[k]), (which, k)
