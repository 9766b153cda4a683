"""Dump the debug timeline of CTA 0 for one stats + fwd launch."""
import ctypes
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paint_with_words_sd_b200 import _native  # noqa: E402
from paint_with_words_sd_b200 import attention as A  # noqa: E402

B, biased, N, H, D, T = 16, 8, 4096, 8, 40, 77
dev = "cuda"
g = torch.Generator().manual_seed(0)
C = H * D
q = (torch.randn(B, N, C, generator=g) * 0.5).half().to(dev)
k = (torch.randn(B, T, C, generator=g) * 0.5).half().to(dev)
v = (torch.randn(B, T, C, generator=g) * 0.5).half().to(dev)
w = (torch.rand(biased, N, T, generator=g) > 0.8).float().to(dev)
idx = torch.tensor(list(range(biased)) + [-1] * (B - biased), dtype=torch.int32, device=dev)
gs = torch.full((1,), 0.4 * math.log(8.0), dtype=torch.float32, device=dev)
L = _native.lib()
L.pww_debug_set_timeline.argtypes = [ctypes.c_void_p]
for _ in range(3):
    A.cross_attention(q, k, v, H, D ** -0.5, w, idx, _native.PWW_STAT_MAX, gs)
torch.cuda.synchronize()
TAGS, ITS = 14, 40
which = sys.argv[1] if len(sys.argv) > 1 else "stats"
buf = torch.zeros(TAGS * ITS, dtype=torch.int64, device=dev)
st = A._state(torch.device("cuda", 0))
stream = torch.cuda.current_stream().cuda_stream
ws = L.pww_xattn_workspace_bytes(B, H, N, T, D)
assert L.pww_debug_set_timeline(buf.data_ptr()) == 0
if which == "stats":
    rc = L.pww_xattn_stats_f16(q.data_ptr(), k.data_ptr(), B, H, N, T, D, q.stride(0), q.stride(1), k.stride(0), k.stride(1),
                               0, idx.data_ptr(), st.stats.data_ptr(), st.workspace.data_ptr(), st.workspace.numel(), stream)
else:
    out = torch.empty_like(q)
    rc = L.pww_xattn_fwd_f16(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), B, H, N, T, D, q.stride(0), q.stride(1),
                             k.stride(0), k.stride(1), out.stride(0), out.stride(1), w.data_ptr(), w.stride(0), idx.data_ptr(),
                             st.stats.data_ptr(), gs.data_ptr(), D ** -0.5, stream)
assert rc == 0
torch.cuda.synchronize()
L.pww_debug_set_timeline(None)
tab = buf.cpu().view(TAGS, ITS)
t0 = int(tab[tab > 0].min())
names = {0: "wg_tmem_loaded", 1: "prod_load", 2: "qk_waits_done", 3: "pv_waits_done", 4: "qk_issued", 5: "wg_sready", 6: "wg_pready_arr",
         7: "wg_math_done_q2", 8: "pv_issued", 9: "wg_math_done_q3", 10: "wg_math_done_q0", 11: "wg_math_done_q1", 12: "wg_epi_done"}
rows = []
for tag in range(TAGS):
    for it in range(ITS):
        if tab[tag, it] > 0:
            rows.append((int(tab[tag, it]) - t0, tag, it))
for t, tag, it in sorted(rows)[:150]:
    print(f"{t:8d}  {names.get(tag, tag):20s} it={it}")
print("---- per-iteration stamps (cycles since start)")
print("it  prod_load qk_wait qk_issued | sready tmem_ld math_done(q2,q3,q0,q1) pready_arr epi_done | pv_wait pv_issued")
for it in range(ITS):
    if tab[1, it] > 0:
        f = lambda tag: (int(tab[tag, it]) - t0) if tab[tag, it] > 0 else -1
        print(f"{it:2d} {f(1):7d} {f(2):7d} {f(4):7d} | {f(5):7d} {f(0):7d}  {f(7):7d} {f(9):7d} {f(10):7d} {f(11):7d}  {f(6):7d} {f(12):7d} | {f(3):7d} {f(8):7d}")
