"""Debug driver (GPU box): fused kernel vs the dense pair on the scenarios of the failing tests; prints where they differ."""
import json
import math
import sys
import os
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from paint_with_words_sd_b200 import _native            # noqa: E402
from paint_with_words_sd_b200 import attention as A     # noqa: E402
from tests.test_xattn_gpu import _inputs, _set_fused_grid  # noqa: E402


def run(q, k, v, H, scale, w, g, stat, idx, impl):
    dev = "cuda"
    gs = torch.tensor([g], dtype=torch.float32, device=dev)
    A.XATTN_IMPL = impl
    out, st = A.cross_attention(q.to(dev), k.to(dev), v.to(dev), H, scale, None if w is None else w.to(dev),
                                None if idx is None else idx.to(dev),
                                _native.PWW_STAT_MAX if stat == "max" else _native.PWW_STAT_STD, gs, return_stats=True)
    torch.cuda.synchronize()
    A.XATTN_IMPL = "fused"
    return out.float().cpu(), None if st is None else st.cpu()


def compare(name, q, k, v, H, D, w, g, stat, idx, grid=0, reps=3):
    B, N, C = q.shape
    ref, st_ref = run(q, k, v, H, D ** -0.5, w, g, stat, idx, "dense")
    amax = ref.abs().max().item()
    for r in range(reps):
        _set_fused_grid(grid)
        try:
            got, st = run(q, k, v, H, D ** -0.5, w, g, stat, idx, "fused")
        finally:
            _set_fused_grid(0)
        d = (got - ref).abs().reshape(B, N, H, D)
        bad = d > 3e-3 * amax
        rec = {"case": name, "rep": r, "rel_err": round(d.max().item() / amax, 5), "bad_elems": int(bad.sum()),
               "stats_fused": None if st is None else [round(float(x), 4) for x in st[:6]],
               "stats_dense": None if st_ref is None else [round(float(x), 4) for x in st_ref[:6]]}
        if bad.any():
            per_bth = bad.reshape(B, (N + 127) // 128 if N % 128 == 0 else -1, 128, H, D).any(-1).any(2) if N % 128 == 0 else None
            if per_bth is not None:
                units = torch.nonzero(per_bth).tolist()          # (image, tile, head)
                rec["bad_units"] = len(units)
                rec["bad_units_first"] = units[:24]
                rows_bad = bad.any(-1).any(-1)                   # [B, N]
                rec["bad_rows_per_image"] = rows_bad.sum(1).tolist()
                u = units[0]
                sub = bad[u[0], u[1] * 128:(u[1] + 1) * 128, u[2]]   # [128, D]
                rec["first_unit_bad_rows"] = int(sub.any(-1).sum())
                rec["first_unit_bad_cols"] = torch.nonzero(sub.any(0)).flatten().tolist()
                rec["first_unit_rows_sample"] = torch.nonzero(sub.any(-1)).flatten().tolist()[:40]
        print(json.dumps(rec), flush=True)


def main():
    golden = np.load(os.path.join(ROOT, "tests", "golden", "mask_builder.npz"))
    # 1. aurora map
    w = torch.from_numpy(golden["aurora_512_w8"])[None]
    q, k, v, _ = _inputs(1, 4096, 8, 40, 77, seed=2026)
    compare("aurora", q, k, v, 8, 40, w, 0.4 * math.log(1 + 14.6146), "max", None)
    # 2. large bias
    q, k, v, _ = _inputs(1, 1024, 8, 40, 77, seed=21)
    gen = torch.Generator().manual_seed(9)
    w = torch.zeros(1, 1024, 77)
    base = torch.rand(1024, 3, generator=gen) * 8.0
    w[0, :, 4] = base[:, 0]; w[0, :, 5] = base[:, 0]; w[0, :, 20] = base[:, 1]; w[0, :, 33] = base[:, 2]
    w[0, :, 20] += base[:, 0]
    compare("large_bias", q, k, v, 8, 40, w, 3.0, "max", None)
    compare("large_bias_g0.3", q, k, v, 8, 40, w, 0.3, "max", None)
    w2 = w.clone(); w2[0, :, 5] = 0
    compare("large_bias_noshare", q, k, v, 8, 40, w2, 3.0, "max", None)
    # 3. all biased
    q, k, v, w = _inputs(3, 1024, 8, 40, 77, seed=77)
    compare("all_biased", q, k, v, 8, 40, w, 0.9, "max", None)
    compare("one_biased_two_un", q, k, v, 8, 40, w, 0.9, "max", torch.tensor([2, -1, -1], dtype=torch.int32))
    # 4. long job lists
    for (N, H, D, B, grid) in [(1024, 8, 40, 4, 8), (1024, 8, 40, 2, 3), (333, 3, 40, 5, 4), (1024, 8, 80, 6, 12)]:
        q, k, v, w = _inputs(B, N, H, D, 77, seed=B * 1000 + N + D)
        nb = (B + 1) // 2
        idx = torch.tensor([(i // 2 if i % 2 == 0 else -1) for i in range(B)], dtype=torch.int32)
        compare(f"long_{N}_{H}_{D}_{B}_g{grid}", q, k, v, H, D, w[:nb].contiguous(), 0.4 * math.log(6.0), "max", idx, grid)
        compare(f"long_{N}_{H}_{D}_{B}_full", q, k, v, H, D, w[:nb].contiguous(), 0.4 * math.log(6.0), "max", idx, 0, reps=2)


if __name__ == "__main__":
    main()
