"""First hardware contact for the experimental four-group forward kernel (csrc/xattn_tc_g4.cuh, variant 2):
outputs compared with the default kernel's (variant 1) on the same inputs, then kernel-only timing of both, then the
op's GPU parity tests under variant 2.  Every result line is flushed to gpurun_out/g4_try.log as it is produced."""
import ctypes
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.chdir(ROOT)
import bench  # noqa: E402
from paint_with_words_sd_b200 import _native  # noqa: E402
from paint_with_words_sd_b200 import attention as A  # noqa: E402

os.makedirs("gpurun_out", exist_ok=True)
LOG = open("gpurun_out/g4_try.log", "a")


def say(d):
    line = json.dumps(d)
    print(line, flush=True)
    LOG.write(line + "\n"); LOG.flush(); os.fsync(LOG.fileno())


L = _native.lib()
L.pww_debug_set_variant.argtypes = [ctypes.c_int]
dev = torch.device("cuda", 0)
peak, _ = bench.measured_peaks()
H, D, T = 8, 40, 77
TINY = os.environ.get("G4_TINY") == "1"     # one small launch and exit: for compute-sanitizer
VAR = int(os.environ.get("G4_VARIANT", "2"))  # 2 = four groups, 4 = the same code with two groups (bisection)
for (B, biased, N) in ([(1, 1, 256)] if TINY else [(1, 1, 256), (2, 1, 4096), (16, 8, 4096), (3, 2, 1000)]):
    g = torch.Generator().manual_seed(B * 1000 + N)
    C = H * D
    q = (torch.randn(B, N, C, generator=g) * 0.5).half().to(dev)
    k = (torch.randn(B, T, C, generator=g) * 0.5).half().to(dev)
    v = (torch.randn(B, T, C, generator=g) * 0.5).half().to(dev)
    w = (torch.rand(biased, N, T, generator=g) > 0.8).float().to(dev)
    idx = torch.tensor(list(range(biased)) + [-1] * (B - biased), dtype=torch.int32, device=dev)
    gs = torch.full((1,), 0.4 * math.log(8.0), dtype=torch.float32, device=dev)
    outs = {}
    for var in (1, VAR):
        assert L.pww_debug_set_variant(var) == 0
        outs[var] = A.cross_attention(q, k, v, H, D ** -0.5, w, idx, _native.PWW_STAT_MAX, gs).float()
        torch.cuda.synchronize()
    diff = (outs[1] - outs[VAR]).abs().max().item()
    say({"check": "g4 vs default", "variant": VAR, "B": B, "biased": biased, "N": N, "max_abs_diff": diff,
         "ref_amax": outs[1].abs().max().item(), "nan": bool(torch.isnan(outs[VAR]).any())})
if TINY:
    sys.exit(0)
for var in (1, VAR):
    L.pww_debug_set_variant(var)
    for (B, biased) in [(2, 1), (16, 8)]:
        r = bench.xattn_roofline(dev, B=B, biased=biased, N=4096, H=8, D=40, iters=32 if B > 2 else 64)
        gbs = r["alg_bytes"] / (r["us_fwd"] * 1e-6) / 1e9
        say({"variant": var, "B": B, "us_fwd": round(r["us_fwd"], 2), "fwd_GBs": round(gbs, 1), "frac": round(gbs / peak, 3)})
L.pww_debug_set_variant(VAR)
import pytest  # noqa: E402

rc = pytest.main(["tests/test_xattn_gpu.py", "-q", "-m", "gpu", "-x", "-p", "no:cacheprovider"])
say({"variant": VAR, "parity_tests_rc": int(rc)})
L.pww_debug_set_variant(1)
