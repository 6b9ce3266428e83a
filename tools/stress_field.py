"""Stress: does a forward launch ever differ from the first forward on the same inputs when feature-map gradient
launches (AVR_FIELD_BWD_ASYNC variants) run in between?  Diagnosis tool."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402

from avr_b200 import _lib, field, field_inputs  # noqa: E402
from field_stub import ray_ordered_case  # noqa: E402

dev = torch.device("cuda:0")
d = ray_ordered_case(sb=2, ns=2, rays=40, k=64, ch=512, h=16, w=12, seed=512)
scale = (d["latent_scaling"] / d["image_shape"]).tolist()
cfg = field.FieldConfig(ns=int(d["ns"]), scale=(scale[0], scale[1]), freqs=tuple(d["freqs"].reshape(-1).tolist()),
                        phases=tuple(d["phases"].reshape(-1).tolist()), include_input=True,
                        normalize_z=bool(int(d.get("normalize_z", 1))), use_viewdirs=True)
poses, focal, c = d["poses"].to(dev), d["focal"].to(dev), d["c"].to(dev)
g_out = d["g_out"].to(dev)


def fwd_bwd(do_bwd):
    xyz = d["xyz"].to(dev).requires_grad_(True)
    vd = d["viewdirs"].to(dev).requires_grad_(True)
    lat = d["latent"].to(dev).requires_grad_(True)
    out = field_inputs(xyz, vd, lat.permute(0, 2, 3, 1).contiguous(), poses, focal, c, cfg)
    res = out.detach().clone()
    if do_bwd:
        out.backward(g_out)
        return res, lat.grad.clone()
    return res, None


_lib.set_option("AVR_FIELD_BWD_ASYNC", 0)
ref_out, ref_lat = fwd_bwd(True)
torch.cuda.synchronize()
for mode in [int(x) for x in sys.argv[1:]] or [14, 16, 4]:
    _lib.set_option("AVR_FIELD_BWD_ASYNC", mode)
    bad_f = bad_b = 0
    for it in range(300):
        out, latg = fwd_bwd(True)
        if not torch.equal(out, ref_out):
            bad_f += 1
            if bad_f <= 2:
                diff = (out != ref_out)
                rows = diff.any(1).nonzero().flatten()
                cols = diff.any(0).nonzero().flatten()
                print(f"mode {mode} it {it}: forward differs in {int(diff.sum())} elements, rows {rows[:8].tolist()}..{int(rows[-1])} ({rows.numel()}), "
                      f"cols {cols[:6].tolist()}..{int(cols[-1])} ({cols.numel()})", flush=True)
        err = float((latg - ref_lat).abs().max() / ref_lat.abs().max())
        if err > 1e-5:
            bad_b += 1
            if bad_b <= 2:
                print(f"mode {mode} it {it}: d_latent rel err {err:.3e}", flush=True)
    print(f"mode {mode}: forward mismatches {bad_f}/300, d_latent mismatches {bad_b}/300", flush=True)
