"""Host model of the bucket-ranking merge of csrc/importance_bins.cu (steps 1-5 of its header):
levels by a monotone fp32 function of the VALUE, 8-bit counters three to a word, in-place prefix by
the two-multiply trick, placement by arrival order, value ranking inside shared levels.  The model
uses the kernel's arithmetic (numpy float32 / uint32) with a RANDOM arrival order — the kernel's
atomics may serve the elements of a level in any order — and must reproduce np.sort exactly."""
import numpy as np
import pytest


def bucket_rank(values, near, span, kc, words, near_level, rng):
    f32 = np.float32
    v = values.astype(f32)
    t_count = v.size
    assert t_count <= 255
    q_levels = 3 * words
    qtop = f32(q_levels - 1 - (2 if near_level else 0))
    kcf = f32(kc)
    scale = f32(qtop * kcf) / f32(span * (kcf + f32(1))) if span > 0 else f32(0)
    t = ((v - f32(near)).astype(f32) * scale).astype(f32)
    t = np.minimum(np.maximum(np.nan_to_num(t, nan=0.0), f32(0)), qtop)
    q = ((t + f32(8388608.0)).astype(f32).view(np.uint32) & 0x7FFFFF).astype(np.int64)
    if near_level:
        q = q + np.where(v > f32(near), 2, np.where(v == f32(near), 1, 0))
    word = (q * 0x55555556) >> 32
    k = q - 3 * word
    cnt = np.zeros(words, dtype=np.uint32)
    arrival = np.zeros(t_count, dtype=np.int64)
    for e in rng.permutation(t_count):                      # atomicAdd in any order
        old = cnt[word[e]]
        arrival[e] = (int(old) >> (8 * k[e])) & 0xFF
        cnt[word[e]] = np.uint32(int(old) + (1 << (8 * k[e])))
    r = 0
    for i in range(words):                                   # prefix pass, two multiplies per word
        x = (int(cnt[i]) * 0x01010100 + r * 0x01010101) & 0xFFFFFFFF
        cnt[i] = x
        r = x >> 24
    assert r == t_count
    out = np.full(t_count, np.nan, dtype=f32)
    tmp = np.full(t_count, np.nan, dtype=f32)
    p = np.array([(int(cnt[word[e]]) >> (8 * k[e])) & 0xFF for e in range(t_count)])
    c = np.array([((int(cnt[word[e]]) >> (8 * k[e] + 8)) & 0xFF) for e in range(t_count)]) - p
    for e in range(t_count):
        if c[e] == 1:
            out[p[e]] = v[e]
        else:
            tmp[p[e] + arrival[e]] = v[e]
    for e in range(t_count):
        if c[e] > 1:
            if near_level and v[e] == f32(near):
                out[p[e] + arrival[e]] = v[e]
                continue
            rank = 0
            for i in range(c[e]):
                f = tmp[p[e] + i]
                rank += (f < v[e]) or (f == v[e] and i < arrival[e])
            out[p[e] + rank] = v[e]
    return out, int(c.max())


@pytest.mark.parametrize("case", ["uniform", "peaky", "one_bin", "all_equal", "unsorted_coarse_outside", "near_quirk", "tiny_span"])
def test_bucket_ranking_is_an_exact_sort(case):
    rng = np.random.default_rng(7)
    f32 = np.float32
    kc, n = 64, 128
    worst = 0
    for trial in range(40):
        near = f32(0.8) if trial % 2 == 0 else f32(rng.uniform(0.5, 3.0))
        span = f32(1.0) if trial % 2 == 0 else f32(rng.uniform(0.05, 4.0))
        if case == "tiny_span":
            span = f32(near * 2.0 ** -20)
        j = np.arange(kc, dtype=f32)
        coarse = (near + (span * (j / f32(kc))).astype(f32) + ((rng.random(kc).astype(f32) * span) / f32(kc)).astype(f32)).astype(f32)
        if case == "peaky":
            pdf = rng.random(kc) ** 6
        elif case == "one_bin":
            pdf = np.zeros(kc); pdf[rng.integers(kc)] = 1.0
        else:
            pdf = np.ones(kc)
        bins = rng.choice(kc, size=n, p=pdf / pdf.sum()).astype(f32)
        bins[::17] = kc                                       # the index may equal Kc (renderers.py:42-43)
        u2 = rng.random(n).astype(f32)
        if case == "all_equal":
            bins[:] = 5; u2[:] = f32(0.25)
        fine = (near + (span * ((bins + u2).astype(f32) / f32(kc)).astype(f32)).astype(f32)).astype(f32)
        parts = [coarse, fine]
        near_level = case == "near_quirk"
        if near_level:                                        # 16 clamped "depth" samples, all == near
            parts = [coarse, fine[:47], np.full(16, near, dtype=f32)]
            parts[0][0] = near
        if case == "unsorted_coarse_outside":
            rng.shuffle(parts[0])
            parts[0][:3] = [near - f32(0.5), near + span * f32(3), near - f32(1e-7)]
        vals = np.concatenate(parts).astype(f32)
        got, cmax = bucket_rank(vals, near, span, kc, 2 * kc, near_level or case == "unsorted_coarse_outside", rng)
        worst = max(worst, cmax)
        assert np.array_equal(got, np.sort(vals)), (case, trial)
    assert worst >= 1
