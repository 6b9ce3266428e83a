"""Kernel cores that compile for the host as well as for the device, walked on the CPU.

Some kernels keep their per-lane work in headers that are plain C++ under g++ and device code
under nvcc (csrc/coarse_packed_core.h).  Here the same functions are built with g++
(-ffp-contract=off, so every operation is one IEEE rounding, like the kernel's __f*_rn
intrinsics) and driven warp by warp, lane by lane, through the launch's control flow; the result
is held bit-for-bit against the oracle.  This checks the index arithmetic (segment tables,
ray search, 16-byte groups that straddle rays and segments, empty rays) and the division
shortcut without a GPU; the `-m gpu` tests then confirm the launch itself.

The host build is test infrastructure: nothing in the package loads it."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

import avr_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "adaptive-volume-rendering_b200", "csrc")
SRC_DIR = os.path.join(ROOT, "tests", "host_cores")


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("host_cores") / "libhostcores.so")
    srcs = [os.path.join(SRC_DIR, f) for f in sorted(os.listdir(SRC_DIR)) if f.endswith(".cpp")]
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I", CSRC,
                    "-I", os.path.join(ROOT, "include"), *srcs, "-o", out], check=True)
    lib = ctypes.CDLL(out)
    P = ctypes.c_void_p
    lib.host_coarse_packed.restype = ctypes.c_int64
    lib.host_coarse_packed.argtypes = [P, P, ctypes.c_int, P, P, ctypes.c_int64, P, ctypes.c_int, ctypes.c_int]
    lib.host_markstein_mismatches.restype = ctypes.c_int64
    lib.host_markstein_mismatches.argtypes = [P, ctypes.c_int64, ctypes.c_int, ctypes.c_int]
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _run_coarse(host, near, far, u, offsets, per_ray=True, misalign=0, force_slow=False):
    """Runs the host walk; `misalign` shifts the u/z base addresses off the 16-byte grid."""
    s = int(offsets[-1])
    ubuf = np.zeros(s + 8, dtype=np.float32)
    zbuf = np.full(s + 8, np.float32(-7.0))
    base = (-(ubuf.ctypes.data // 4) % 4 + misalign) % 4  # element index of a 16-byte boundary + misalign
    zbase = (-(zbuf.ctypes.data // 4) % 4 + misalign) % 4
    uv, zv = ubuf[base:base + s], zbuf[zbase:zbase + s]
    uv[:] = u
    vec_ok = int(uv.ctypes.data % 16 == 0 and zv.ctypes.data % 16 == 0)
    assert vec_ok == (misalign == 0)
    slow = host.host_coarse_packed(_ptr(near), _ptr(far), int(per_ray), _ptr(uv), _ptr(offsets), len(offsets) - 1,
                                   _ptr(zv), vec_ok, int(force_slow))
    # nothing outside the stream was written
    assert np.all(zbuf[:zbase] == -7.0) and np.all(zbuf[zbase + s:] == -7.0)
    return zv.copy(), slow


def _reference(near, far, u, offsets, per_ray=True):
    want = np.empty_like(u)
    off_t = torch.from_numpy(offsets)
    nt, ft, ut = torch.from_numpy(near), torch.from_numpy(far), torch.from_numpy(u)
    for k, rays, idx in O.bucketed(off_t):
        if k == 0:
            continue
        n = nt[rays] if per_ray else nt.expand(len(rays))
        f = ft[rays] if per_ray else ft.expand(len(rays))
        want[idx.numpy()] = O.coarse_z(n.unsqueeze(0), f.unsqueeze(0), k, ut[idx].unsqueeze(0))[0].numpy()
    return want


def _case(counts, seed, per_ray=True):
    rng = np.random.default_rng(seed)
    counts = np.asarray(counts, dtype=np.int64)
    offsets = np.zeros(len(counts) + 1, dtype=np.int64)
    offsets[1:] = np.cumsum(counts)
    s = int(offsets[-1])
    nb = len(counts) if per_ray else 1
    d = (0.9 + 0.8 * rng.random(nb)).astype(np.float32)
    near, far = (d - np.float32(0.15)).astype(np.float32), (d + np.float32(0.15)).astype(np.float32)
    u = rng.random(s, dtype=np.float32)
    return near, far, u, offsets


COUNT_CASES = {
    "c4_like_8_256": lambda rng: rng.integers(8, 257, 1000),
    "with_empty_and_single": lambda rng: rng.choice([0, 0, 1, 2, 3, 5, 31, 32, 33, 64, 200], 777),
    "all_empty_segments": lambda rng: np.concatenate([np.zeros(70, int), [5], np.zeros(40, int), [1, 0, 0, 9]]),
    "one_ray": lambda rng: [13],
    "long_rays": lambda rng: [1030, 0, 4097, 7, 700, 257],
    "odd_counts_unaligned": lambda rng: rng.integers(1, 12, 333) * 2 + 1,
    "no_rays_in_last_segment_slot": lambda rng: rng.integers(8, 40, 33),
}


@pytest.mark.parametrize("name", sorted(COUNT_CASES))
@pytest.mark.parametrize("misalign", [0, 1])
def test_packed_coarse_core_matches_the_oracle_bit_for_bit(host, name, misalign):
    counts = COUNT_CASES[name](np.random.default_rng(1))
    near, far, u, offsets = _case(counts, seed=2)
    got, slow = _run_coarse(host, near, far, u, offsets, misalign=misalign)
    assert slow == 0
    want = _reference(near, far, u, offsets)
    assert got.tobytes() == want.tobytes(), int(np.sum(got != want))


def test_packed_coarse_core_scalar_bounds_and_adversarial_draws(host):
    counts = np.random.default_rng(5).integers(0, 300, 500)
    near, far, u, offsets = _case(counts, seed=6, per_ray=False)
    # draws at the ends of [0,1), tiny ones, and denormal products that must leave the shortcut
    special = np.array([0.0, 2.0 ** -24, 1 - 2.0 ** -24, 1e-30, 1e-38, 1e-44, 0.5], dtype=np.float32)
    u[:: 11] = np.resize(special, len(u[:: 11]))
    got, _ = _run_coarse(host, near, far, u, offsets, per_ray=False)
    want = _reference(near, far, u, offsets, per_ray=False)
    assert got.tobytes() == want.tobytes()
    # the per-ray loop (taken for segments beyond 2^31 samples) computes the same
    got_slow, slow = _run_coarse(host, near, far, u, offsets, per_ray=False, force_slow=True)
    assert slow == (len(counts) + 31) // 32 and got_slow.tobytes() == want.tobytes()


def test_division_shortcut_is_the_ieee_quotient(host):
    """q0 = a*y, r = fma(-K, q0, a), q = fma(r, y, q0) with y = RN(1/K) against a / K: every K up to
    4096 (and a band below 2^24) against 20k numerators spread over the admitted range
    [2^-90, 2^100) with both signs, plus every bin numerator j < K."""
    rng = np.random.default_rng(0)
    mant = rng.random(20000, dtype=np.float32) + np.float32(1.0)
    expo = rng.integers(-90, 100, 20000)
    a = np.ldexp(mant, expo).astype(np.float32) * rng.choice(np.float32([-1, 1]), 20000)
    a[:64] = np.ldexp(np.float32(1.0), np.arange(-90, -26)).astype(np.float32)   # exact powers of two
    a[64:128] = np.nextafter(a[:64], np.float32(0))                               # all-ones significands
    assert host.host_markstein_mismatches(_ptr(a), len(a), 1, 4096) == 0
    assert host.host_markstein_mismatches(_ptr(a), 2000, (1 << 24) - 40, (1 << 24) - 2) == 0
