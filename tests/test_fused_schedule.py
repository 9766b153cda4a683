"""CPU checks of the one-launch cross-attention kernel's job lists (csrc/xattn_fused.cuh: FxWalk, FxJobs, fx_range,
fx_cta_has_image), replayed on the host by the library itself (the same code the kernel compiles).  What must hold:
every (image, head, row tile) is a main job exactly once; every unit of an image with a weight map is a stat job exactly
once, before any main job of its CTA; the unbiased main jobs of a CTA precede its biased ones (they overlap the grid
barrier); the set of CTAs the barrier of image b waits for is exactly the set of CTAs that publish a partial for b."""
import ctypes
import itertools

import numpy as np
import pytest

from paint_with_words_sd_b200 import _native

K_MAX_LOCAL = 4


def _jobs(B, H, tiles, grid, widx):
    L = _native.lib()
    L.pww_debug_fused_schedule.restype = ctypes.c_int
    L.pww_debug_fused_schedule.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    w = np.asarray(widx, dtype=np.int32)
    cap = 2 * B * H * tiles + 8
    out = np.full((cap, 10), -7, dtype=np.int32)
    n = L.pww_debug_fused_schedule(B, H, tiles, grid, w.ctypes.data, out.ctypes.data, cap)
    assert n >= 0
    return out[:n]


def _has_image(cta, grid, B, H, tiles, widx, b):
    L = _native.lib()
    L.pww_debug_fused_cta_has_image.restype = ctypes.c_int
    L.pww_debug_fused_cta_has_image.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p, ctypes.c_int]
    w = np.asarray(widx, dtype=np.int32)
    return L.pww_debug_fused_cta_has_image(cta, grid, B, H, tiles, w.ctypes.data, b)


CASES = [
    (2, 8, 32, 148, [0, -1]),                      # the workload: cond + uncond, SD1.5 64x64 latents
    (2, 8, 32, 148, [-1, 0]),
    (16, 8, 32, 148, list(range(8)) + [-1] * 8),   # batched CFG, conditional half first
    (16, 8, 32, 148, [v for i in range(8) for v in (i, -1)]),
    (3, 5, 72, 148, [0, 1, -1]),                   # odd head count (SD2.1), one biased image without a partner
    (4, 20, 5, 148, [-1, -1, -1, 0]),
    (1, 8, 32, 148, [0]),                          # reference-style single call
    (1, 8, 1, 148, [-1]),                          # fewer units than CTAs
    (5, 10, 3, 7, [0, 1, 2, 3, 4]),                # all biased, capped grid
    (32, 8, 2, 148, [(-1 if i % 3 else i // 3) for i in range(32)]),
    (4, 8, 8, 8, [0, -1, 1, -1]),                  # capped grid: long job lists
    (2, 8, 8, 3, [0, -1]),
]


@pytest.mark.parametrize("B,H,tiles,grid,widx", CASES)
def test_job_lists(B, H, tiles, grid, widx):
    j = _jobs(B, H, tiles, grid, widx)
    main = j[j[:, 2] == 1]
    stat = j[j[:, 2] == 0]
    units = set(itertools.product(range(B), range(H), range(tiles)))
    assert len(main) == len(units) and set(map(tuple, main[:, 4:7].tolist())) == units
    biased_units = {u for u in units if widx[u[0]] >= 0}
    assert len(stat) == len(biased_units) and set(map(tuple, stat[:, 4:7].tolist())) == biased_units
    for cta in np.unique(j[:, 0]):
        rows = j[j[:, 0] == cta]
        assert rows[:, 1].tolist() == list(range(len(rows)))               # global job index counts up
        kinds = rows[:, 2].tolist()
        assert kinds == sorted(kinds)                                      # stat jobs first
        m = rows[rows[:, 2] == 1]
        assert m[:, 3].tolist() == list(range(len(m)))                     # main index counts up
        flags = m[:, 7].tolist()
        assert flags == sorted(flags)                                      # unbiased main jobs before biased ones
        # stat and biased-main lists visit the same units in the same order with the same local image index
        sb = rows[rows[:, 2] == 0][:, [4, 5, 6, 8]].tolist()
        mb = m[m[:, 7] == 1][:, [4, 5, 6, 8]].tolist()
        assert sb == mb
        if sb:
            li = [r[3] for r in sb]
            assert li[0] == 0 and all(b - a in (0, 1) for a, b in zip(li, li[1:])) and max(li) < K_MAX_LOCAL
        assert all(r[8] == -1 for r in m[m[:, 7] == 0])
        # balanced: a CTA's biased share is within one pair of its unbiased share when the batch is CFG-shaped
    # grid barrier membership
    for b in range(B):
        if widx[b] < 0:
            continue
        publishers = set(stat[stat[:, 4] == b][:, 0].tolist())
        expected = {c for c in range(grid) if _has_image(c, grid, B, H, tiles, widx, b)}
        assert publishers == expected, (b, sorted(publishers ^ expected))


@pytest.mark.parametrize("order", ["cond_first", "uncond_first", "interleaved"])
def test_cfg_batches_are_balanced_whatever_the_image_order(order):
    B, H, tiles, grid = 16, 8, 32, 148
    widx = {"cond_first": list(range(8)) + [-1] * 8, "uncond_first": [-1] * 8 + list(range(8)),
            "interleaved": [v for i in range(8) for v in (i, -1)]}[order]
    j = _jobs(B, H, tiles, grid, widx)
    main = j[j[:, 2] == 1]
    for cta in range(grid):
        rows = main[main[:, 0] == cta]
        nb = int(rows[:, 7].sum())
        assert abs(2 * nb - len(rows)) <= 1, (cta, nb, len(rows))
