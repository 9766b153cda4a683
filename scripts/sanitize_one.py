"""One launch of each cross-attention path under compute-sanitizer (GPU box):
    compute-sanitizer --tool racecheck python scripts/sanitize_one.py [N H D max|std]
Shapes: the bench launch (B=2, one biased image, N=4096, H=8, D=40) shrunk to N=1024 rows so the instrumented run ends in
seconds but still puts several units (and > 1 mask group) on a CTA when the grid is limited by PWW_DEBUG_GRID."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from paint_with_words_sd_b200 import _native  # noqa: E402
from paint_with_words_sd_b200 import attention as A  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
B, N, H, D, T = 4, 1024, 8, 40, 77      # 4 images: with PWW_DEBUG_GRID=6 every CTA runs ring mode over ~10 units
if len(sys.argv) > 3:
    N, H, D = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
stat = _native.PWW_STAT_STD if (len(sys.argv) > 4 and sys.argv[4] == "std") else _native.PWW_STAT_MAX
q = (torch.randn(B, N, H * D) * 0.5).half().to(dev)
k = (torch.randn(B, T, H * D) * 0.5).half().to(dev)
v = (torch.randn(B, T, H * D) * 0.5).half().to(dev)
w = torch.zeros(2, N, T)
for i in range(2):
    w[i, :, 3 + i] = (torch.rand(N) > 0.5).float() * 1.5
    w[i, :, 9 + i] = (torch.rand(N) > 0.5).float() * 0.7
idx = torch.tensor([0, -1, 1, -1], dtype=torch.int32, device=dev)
if os.environ.get("PWW_DEBUG_GRID"):
    import ctypes
    L = _native.lib()
    L.pww_debug_set_fused_grid.argtypes = [ctypes.c_int]
    L.pww_debug_set_fused_grid(int(os.environ["PWW_DEBUG_GRID"]))
gs = torch.tensor([0.4 * math.log(8.0)], dtype=torch.float32, device=dev)
out = A.cross_attention(q, k, v, H, D ** -0.5, w.to(dev), idx, stat, gs)
torch.cuda.synchronize()
print("ok", N, H, D, "std" if stat == _native.PWW_STAT_STD else "max", float(out.float().abs().max()))
if len(sys.argv) > 5 and sys.argv[5] == "self":       # the native self-attention kernel on the same q (keys = queries)
    A.SELF_ATTN_IMPL = "native"
    o2 = A.self_attention(q, q, q, H, D ** -0.5)
    torch.cuda.synchronize()
    print("ok self", float(o2.float().abs().max()))
