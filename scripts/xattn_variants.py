"""A/B the structural variants of the tcgen05 cross-attention forward kernel on one device, in one process:
for each variant run the GPU parity tests of the op, then the kernel-only timing of bench.py's roofline leg.
usage: python scripts/xattn_variants.py [variants...]   (default 1 3)
variants: 0 per-thread stores | 1 TMA-store epilogue at D = 40 (the default build) |
3 TMA-store epilogue at every head dim"""
import ctypes
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)
import bench  # noqa: E402
from paint_with_words_sd_b200 import _native  # noqa: E402

L = _native.lib()
L.pww_debug_set_variant.argtypes = [ctypes.c_int]
dev = torch.device("cuda", 0)
peak, _ = bench.measured_peaks()
variants = [int(a) for a in sys.argv[1:]] or [1, 3]
shapes = [(4096, 8, 40, 2, 1), (4096, 8, 40, 16, 8), (1024, 8, 80, 16, 8), (256, 8, 160, 16, 8), (9216, 5, 64, 16, 8),
          (2304, 10, 64, 16, 8)]
for var in variants:
    assert L.pww_debug_set_variant(var) == 0
    rc = pytest.main(["tests/test_xattn_gpu.py", "-q", "-m", "gpu", "-x", "-p", "no:cacheprovider"])
    print(json.dumps({"variant": var, "parity_tests_rc": int(rc)}), flush=True)
    if int(rc) != 0:
        continue
    for (N, H, D, B, biased) in shapes:
        r = bench.xattn_roofline(dev, B=B, biased=biased, N=N, H=H, D=D, iters=32 if B > 2 else 64)
        gbs = r["alg_bytes"] / (r["us_fwd"] * 1e-6) / 1e9
        print(json.dumps({"variant": var, "N": N, "H": H, "D": D, "B": B, "us_fwd": round(r["us_fwd"], 2),
                          "fwd_GBs": round(gbs, 1), "frac": round(gbs / peak, 3)}), flush=True)
L.pww_debug_set_variant(1)
