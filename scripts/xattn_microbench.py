"""Kernel-only timing of the PwW cross-attention op (the roofline leg of bench.py, standalone):
python scripts/xattn_microbench.py [sd21] [dense] [quick]   -- `dense` also times the round-1 two-launch path on the same inputs."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda", 0)
peak, src = bench.measured_peaks()
shapes = [(4096, 8, 40), (1024, 8, 80), (256, 8, 160), (64, 8, 160)]
if "sd21" in sys.argv[1:]:
    shapes = [(9216, 5, 64), (2304, 10, 64), (576, 20, 64), (144, 20, 64)]
dense = "dense" in sys.argv[1:]
if "quick" in sys.argv[1:]:
    shapes = shapes[:1]
for (N, H, D) in shapes:
    for (B, biased) in [(2, 1), (16, 8)]:
        r = bench.xattn_roofline(dev, B=B, biased=biased, N=N, H=H, D=D, iters=32 if B > 2 else 64, dense_pair=dense)
        gbs = r["alg_bytes"] / (r["us_op"] * 1e-6) / 1e9
        row = {"N": N, "H": H, "D": D, "B": B, "biased": biased, "us_op": round(r["us_op"], 2),
               "alg_MB": round(r["alg_bytes"] / 1e6, 2), "GBs": round(gbs, 1), "frac": round(gbs / peak, 3),
               "frac_dense_map_bytes": round(r["alg_bytes_dense_map"] / (r["us_op"] * 1e-6) / 1e9 / peak, 3)}
        if dense:
            row["us_dense_stats"], row["us_dense_fwd"] = round(r["us_dense_stats"], 2), round(r["us_dense_fwd"], 2)
        print(json.dumps(row), flush=True)
