"""Drive the one-launch PwW cross-attention kernel for ncu: python scripts/profile_xattn.py [B] [biased] [N] [H] [D] [iters]
Rotating Q / packed-map / output buffers larger than L2, the golden aurora_1 map (or a region-structured synthetic one)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from paint_with_words_sd_b200 import _native  # noqa: E402
from paint_with_words_sd_b200 import attention as A  # noqa: E402
from paint_with_words_sd_b200.conditioning import pack_weight_map  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
biased = int(sys.argv[2]) if len(sys.argv) > 2 else 1
N = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
H = int(sys.argv[4]) if len(sys.argv) > 4 else 8
D = int(sys.argv[5]) if len(sys.argv) > 5 else 40
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 5
T, C = 77, H * D
dev = "cuda"
g = torch.Generator().manual_seed(0)
nsets = 24
qs = [(torch.randn(B, N, C, generator=g) * 0.5).half().to(dev) for _ in range(nsets)]
k = (torch.randn(B, T, C, generator=g) * 0.5).half().to(dev)
v = (torch.randn(B, T, C, generator=g) * 0.5).half().to(dev)
base = bench.golden_weight_map(N)
if base is None:
    base = bench.region_weight_map(N, T)
mp0, ci0 = pack_weight_map(torch.stack([base] * biased, 0))
packs = [(mp0.to(dev).clone(), ci0.to(dev)) for _ in range(nsets)]
idx = torch.tensor(list(range(biased)) + [-1] * (B - biased), dtype=torch.int32, device=dev)
gs = torch.full((1,), 0.4 * math.log(8.0), dtype=torch.float32, device=dev)
for i in range(iters):
    A.cross_attention(qs[i % nsets], k, v, H, D ** -0.5, None, idx, _native.PWW_STAT_MAX, gs, packed=packs[i % nsets])
torch.cuda.synchronize()
print("done", B, biased, N, H, D)
