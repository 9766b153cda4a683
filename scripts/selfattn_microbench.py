"""Self-attention kernel timing vs torch SDPA (comparison only)."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paint_with_words_sd_b200 import attention as A  # noqa: E402

dev = "cuda"
shapes = [(2, 4096, 8, 40), (2, 1024, 8, 80), (2, 256, 8, 160), (2, 64, 8, 160), (2, 9216, 5, 64), (2, 2304, 10, 64)]
for (B, N, H, D) in shapes:
    g = torch.Generator().manual_seed(0)
    C = H * D
    q, k, v = [(torch.randn(B, N, C, generator=g) * 0.5).half().to(dev) for _ in range(3)]
    res = {}
    for impl in ("native", "torch-sdpa"):
        A.SELF_ATTN_IMPL = impl
        for _ in range(3):
            A.self_attention(q, k, v, H, D ** -0.5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for _ in range(iters):
            A.self_attention(q, k, v, H, D ** -0.5)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        res[impl] = us
    flops = 4.0 * B * N * N * C
    print(json.dumps({"B": B, "N": N, "H": H, "D": D, "native_us": round(res["native"], 1),
                      "sdpa_us": round(res["torch-sdpa"], 1), "native_TFLOPs": round(flops / res["native"] / 1e6, 1),
                      "sdpa_TFLOPs": round(flops / res["torch-sdpa"] / 1e6, 1)}))
