"""clock64 timeline of one CTA of the one-launch kernel: python scripts/fused_timeline.py [B] [biased] [cta] [N] [H] [D]"""
import ctypes
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
cta = int(sys.argv[3]) if len(sys.argv) > 3 else 0
N = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
H = int(sys.argv[5]) if len(sys.argv) > 5 else 8
D = int(sys.argv[6]) if len(sys.argv) > 6 else 40
T, C = 77, H * D
dev = "cuda"
g = torch.Generator().manual_seed(0)
q = (torch.randn(B, N, C, generator=g) * 0.5).half().to(dev)
k = (torch.randn(B, T, C, generator=g) * 0.5).half().to(dev)
v = (torch.randn(B, T, C, generator=g) * 0.5).half().to(dev)
base = bench.golden_weight_map(N)
if base is None:
    base = bench.region_weight_map(N, T)
mp0, ci0 = pack_weight_map(torch.stack([base] * max(1, biased), 0))
pk = (mp0.to(dev), ci0.to(dev))
idx = torch.tensor(list(range(biased)) + [-1] * (B - biased), dtype=torch.int32, device=dev)
gs = torch.full((1,), 0.4 * math.log(8.0), dtype=torch.float32, device=dev)
L = _native.lib()
L.pww_debug_set_fused_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(20):          # warm clocks
    A.cross_attention(q, k, v, H, D ** -0.5, None, idx, _native.PWW_STAT_MAX, gs, packed=pk if biased else None)
torch.cuda.synchronize()
TAGS, ITS = 24, 64
buf = torch.zeros(TAGS * ITS, dtype=torch.int64, device=dev)
# L2 of the previous launches' Q / K / V is evicted by READING a large buffer (clean lines: a fill would leave the L2 full of
# dirty lines whose write-back the timed launch then pays for)
if os.environ.get("PWW_TL_FLUSH", "read") == "read":
    flush.view(torch.int32).sum().item()
else:
    flush.fill_(1)
torch.cuda.synchronize()
assert L.pww_debug_set_fused_timeline(buf.data_ptr(), cta) == 0
A.cross_attention(q, k, v, H, D ** -0.5, None, idx, _native.PWW_STAT_MAX, gs, packed=pk if biased else None)
torch.cuda.synchronize()
L.pww_debug_set_fused_timeline(None, 0)
tab = buf.cpu().view(TAGS, ITS)
t0 = int(tab[13, 0])
f = lambda tag, it: (int(tab[tag, it]) - t0) if tab[tag, it] > 0 else -99999999   # noqa: E731
print(f"B={B} biased={biased} cta={cta} N={N} H={H} D={D}; cycles since the post-prologue __syncthreads")
print(f"kernel entry {f(21, 0)}; job table built {f(22, 0)}; TMEM allocated {f(23, 0)} (cycles relative to the post-prologue sync)")
print(f"first Q load about to issue {f(15, 1)} {f(15, 2)}")
print(f"grid barrier: start {f(10, 0)} end {f(11, 0)}; publish {f(12, 0)}; softmax groups done {f(14, 0)} {f(14, 1)}; kernel end {f(15, 0)}")
print("S issuer: q_issued(TMA) | q_ready k_ready all_ready s_issued || softmax: s_seen p_ready epi_done || PV: v_ready all_ready pv_issued "
      "|| loaders: k_stage_free k_copied v_stage_free v_copied")
g = lambda tag, it: f"{f(tag, it):7d}" if tab[tag, it] > 0 else "      ."   # noqa: E731
for it in range(ITS):
    if tab[0, it] > 0 or tab[3, it] > 0 or tab[16, it] > 0:
        print(f"{it:2d} | {g(0, it)} | {g(1, it)} {g(7, it)} {g(2, it)} {g(3, it)} || {g(4, it)} {g(5, it)} {g(6, it)} || "
              f"{g(20, it)} {g(8, it)} {g(9, it)} || {g(16, it)} {g(17, it)} {g(18, it)} {g(19, it)}")
