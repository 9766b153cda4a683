"""Kernel-only timing of the PwW cross-attention op (the roofline leg of bench.py, standalone)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda", 0)
peak, src = bench.measured_peaks()
shapes = [(4096, 8, 40), (1024, 8, 80), (256, 8, 160), (64, 8, 160)]
if len(sys.argv) > 1 and sys.argv[1] == "sd21":
    shapes = [(9216, 5, 64), (2304, 10, 64), (576, 20, 64), (144, 20, 64)]
for (N, H, D) in shapes:
    for (B, biased) in [(2, 1), (16, 8)]:
        r = bench.xattn_roofline(dev, B=B, biased=biased, N=N, H=H, D=D, iters=32 if B > 2 else 64)
        gbs = r["alg_bytes"] / (r["us_fwd"] * 1e-6) / 1e9
        print(json.dumps({"N": N, "H": H, "D": D, "B": B, "biased": biased, "us_fwd": round(r["us_fwd"], 2),
                          "us_stats": round(r["us_stats"], 2), "alg_MB": round(r["alg_bytes"] / 1e6, 2),
                          "fwd_GBs": round(gbs, 1), "frac": round(gbs / peak, 3),
                          "op_frac": round(r["alg_bytes"] / ((r["us_fwd"] + r["us_stats"]) * 1e-6) / 1e9 / peak, 3)}))
