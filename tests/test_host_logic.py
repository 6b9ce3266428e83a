"""Host-side logic that needs no GPU: how the reference's tensors are mapped onto the kernels'
layouts (zero-copy recovery of the radiance field's buffer, scalar vs per-ray bounds), ray sharding,
the ray classes of the packed importance sampler, and the refusal to run on CPU."""
import pytest
import torch


def test_as_rgbs_recovers_the_field_buffer_without_a_copy():
    from avr_b200.renderers import _as_rgbs
    out = torch.rand(2, 5 * 7, 4)                       # the field's (SB, R*K, 4) output
    x = out.view(2, 5, 7, 4)
    sigma, rad = x[..., 3:4], x[..., :3]                # how the reference slices it (renderers.py:177-178)
    rgbs = _as_rgbs(sigma, rad)
    assert rgbs.data_ptr() == out.data_ptr() and rgbs.shape == (2, 5, 7, 4)
    # anything else is interleaved into a fresh buffer with the same values
    rgbs2 = _as_rgbs(sigma.clone(), rad.clone())
    assert rgbs2.data_ptr() != out.data_ptr() and torch.equal(rgbs2, x)
    # slices of two different buffers must not be mistaken for one
    other = torch.rand(2, 5, 7, 4)
    assert torch.equal(_as_rgbs(other[..., 3:4], rad), torch.cat([rad, other[..., 3:4]], -1))


def test_bounds_scalar_and_per_ray():
    from avr_b200 import ops
    near = torch.tensor([0.8]).expand(3, 11)            # VolumeRenderer: stride-0 expand of one element (renderers.py:169)
    far = torch.tensor([1.8]).expand(3, 11)
    n, f, stride = ops._bounds(near, far, 33)
    assert stride == 0 and n.numel() == 1 and f.numel() == 1 and float(n) == pytest.approx(0.8)
    d = torch.rand(3, 11)
    n, f, stride = ops._bounds(d - 0.15, d + 0.15, 33)  # AdaptiveVolumeRenderer: per-ray bounds
    assert stride == 1 and n.shape == (33,) and n.is_contiguous()
    n, f, stride = ops._bounds(near, d + 0.15, 33)       # mixed: the scalar side is broadcast
    assert stride == 1 and torch.equal(n, torch.full((33,), 0.8))
    with pytest.raises(Exception):
        ops._bounds(d[:, :5], d, 33)


def test_cpu_tensors_are_refused():
    import avr_b200
    from avr_b200 import ops
    with pytest.raises(avr_b200.AvrError):
        ops.composite(torch.rand(4, 8, 4), torch.rand(4, 8))
    with pytest.raises(avr_b200.AvrError):
        avr_b200.sample_coarse(torch.rand(1, 4), torch.rand(1, 4) + 1, 8, u=torch.rand(1, 4, 8))
    with pytest.raises(avr_b200.AvrError):
        ops.ray_points(torch.rand(1, 4, 3), torch.rand(1, 4, 3), torch.rand(1, 4, 8))
    ren = avr_b200.VolumeRenderer(0.8, 1.8, 8, 4, 2, 0.01)
    assert list(ren.state_dict()) == []
    with pytest.raises(avr_b200.AvrError):
        ren(torch.eye(4).expand(1, 4, 4, 4), torch.eye(3)[None], torch.rand(1, 4, 2), lambda *a, **k: None)


def test_fp64_is_refused():
    from avr_b200 import ops, AvrError
    with pytest.raises(AvrError):
        ops._f32c(torch.rand(3, dtype=torch.float64))


def test_shard_bounds_packed_balances_samples():
    from avr_b200.dist import shard_bounds, shard_bounds_packed
    counts = torch.tensor([1] * 50 + [100] * 50)
    offsets = torch.zeros(101, dtype=torch.int64)
    offsets[1:] = torch.cumsum(counts, 0)
    cuts = [shard_bounds_packed(offsets, 4, r) for r in range(4)]
    assert cuts[0][0] == 0 and cuts[-1][1] == 100
    assert all(cuts[i][1] == cuts[i + 1][0] for i in range(3))
    per_rank = [int(offsets[b] - offsets[a]) for a, b in cuts]
    assert max(per_rank) - min(per_rank) <= 200           # within two long rays of each other
    assert [shard_bounds(10, 3, r) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]


def test_accelerate_switches_a_reference_model():
    """avr_b200.accelerate on the reference's own RadFieldAndRenderer (train.py:255-276): the renderers
    are replaced by their avr_b200 equivalents with the same configuration, the adaptive renderer's
    lstm / out_layer stay the same objects, the field's forward is rebound, the state_dict is unchanged."""
    import ref_shim
    if not ref_shim.available():
        pytest.skip("reference checkout not present")
    ref = ref_shim.load()
    import models
    import avr_b200
    from ref_shim import Conf
    conf = Conf(use_encoder=True, use_global_encoder=False, use_xyz=True, canon_xyz=False, use_code=True,
                code=dict(num_freqs=6, freq_factor=1.5, include_input=True), use_viewdirs=True, use_code_viewdirs=False,
                mlp_coarse=dict(type="resnet", n_blocks=1, d_hidden=16), mlp_fine=dict(type="resnet", n_blocks=1, d_hidden=16),
                encoder=dict(backbone="resnet34", pretrained=False, num_layers=4))
    net = models.make_new_model(conf)
    ren = ref.VolumeRenderer.from_conf(Conf(near=0.8, far=1.8, n_coarse=64, n_fine=32, n_fine_depth=16, depth_std=0.01))
    model = models.RadFieldAndRenderer(net, ren)
    keys = set(model.state_dict())
    stock_forward = net.forward
    out = avr_b200.accelerate(model)
    assert out is model and isinstance(model.renderer, avr_b200.VolumeRenderer) and set(model.state_dict()) == keys
    r = model.renderer
    assert (r.n_coarse, r.n_fine, r.n_fine_depth, r.depth_std, r.white_back) == (64, 32, 16, 0.01, True)
    assert torch.equal(r.near.cpu(), ren.near.cpu()) and torch.equal(r.far.cpu(), ren.far.cpu())
    assert model.rf is net and net.forward != stock_forward and net.forward.__func__.__name__ == "_fused_forward"
    # adaptive renderer: parameters are shared, not copied
    aren = ref.AdaptiveVolumeRenderer(512, 10, 0.15, 20, True)
    model2 = avr_b200.accelerate(models.RadFieldAndRenderer(models.make_new_model(conf), aren), fuse_field=False)
    a = model2.renderer
    assert isinstance(a, avr_b200.AdaptiveVolumeRenderer) and a.lstm is aren.lstm and a.out_layer is aren.out_layer
    assert (a.steps, a.epsilon, a.n_coarse, a.white_back, a.n_feature_channels) == (10, 0.15, 20, True, 512)
    assert model2.rf.forward.__func__.__name__ == "forward"             # untouched with fuse_field=False
    # the third renderer, Raymarcher (renderers.py:292-358): same treatment
    rm = ref.Raymarcher(512, 10)
    m3 = avr_b200.convert_renderer(rm)
    assert isinstance(m3, avr_b200.Raymarcher) and m3.lstm is rm.lstm and m3.out_layer is rm.out_layer and m3.steps == 10
    assert sorted(m3.state_dict()) == sorted(rm.state_dict())
    # a field outside the kernels' family keeps its torch forward; the renderer is still switched
    other = models.make_new_model(conf)
    other.use_global_encoder = True
    model3 = avr_b200.accelerate(models.RadFieldAndRenderer(other, ref.VolumeRenderer(0.8, 1.8, 8, 4, 4, 0.01)))
    assert isinstance(model3.renderer, avr_b200.VolumeRenderer) and other.forward.__func__.__name__ == "forward"
    with pytest.raises(avr_b200.AvrError):
        avr_b200.convert_renderer(torch.nn.Linear(2, 2))


def test_lane_local_sorting_networks_sort_every_zero_one_input():
    """csrc/sort_net.cuh presort_lane: the 5 / 19 / 60-comparator networks, parsed out of the header."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import check_sort_nets
    found = dict(check_sort_nets.networks())
    assert sorted(found) == [4, 8, 16] and [len(found[n]) for n in (4, 8, 16)] == [5, 19, 60]
    for n, net in found.items():
        assert check_sort_nets.sorts(n, net)
