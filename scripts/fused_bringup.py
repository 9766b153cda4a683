"""Bring-up driver for the one-launch cross-attention kernel (GPU box): every case runs in its own process (a trap or
a barrier time-out kills only that case) and prints max error vs the fp32 oracle, the statistic and the launch time.
    python scripts/fused_bringup.py            # all cases
    python scripts/fused_bringup.py 3          # one case (in-process)"""
import json
import math
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

#        N     H   D   B  biased pattern      stat   grid
CASES = [(128, 1, 40, 1, [-1], "max", 0),
         (128, 1, 40, 1, [0], "max", 0),
         (128, 2, 64, 1, [0], "max", 0),
         (128, 2, 80, 1, [0], "max", 0),
         (128, 2, 160, 1, [0], "max", 0),
         (4096, 8, 40, 2, [-1, -1], "max", 0),
         (4096, 8, 40, 2, [0, -1], "max", 0),
         (4096, 8, 40, 2, [0, -1], "std", 0),
         (1024, 8, 40, 4, [0, -1, 1, -1], "max", 8),
         (1024, 8, 80, 2, [0, -1], "max", 0),
         (256, 8, 160, 2, [0, -1], "max", 0),
         (9216, 5, 64, 2, [0, -1], "max", 0),
         (4096, 8, 40, 16, [v for i in range(8) for v in (i, -1)], "max", 0)]


def run_case(ci):
    import ctypes
    import torch
    from oracle import pww_oracle as O
    from paint_with_words_sd_b200 import _native
    from paint_with_words_sd_b200 import attention as A
    N, H, D, B, idx, stat, grid = CASES[ci]
    T = 77
    g = torch.Generator().manual_seed(ci)
    C = H * D
    q = (torch.randn(B, N, C, generator=g) * 0.5).half()
    k = (torch.randn(B, T, C, generator=g) * 0.5).half()
    v = (torch.randn(B, T, C, generator=g) * 0.5).half()
    nbw = max(idx) + 1
    w = None
    if nbw > 0:
        w = torch.zeros(nbw, N, T)
        for b in range(nbw):
            cols = torch.randperm(T, generator=g)[:7]
            for c in cols:
                w[b, :, c] += (torch.rand(N, generator=g) > 0.6).float() * float(torch.rand(1, generator=g) * 2)
    gg = 0.4 * math.log(8.0)
    L = _native.lib()
    L.pww_debug_set_fused_grid.argtypes = [ctypes.c_int]
    L.pww_debug_set_fused_grid(grid)
    dev = "cuda"
    gs = torch.tensor([gg], dtype=torch.float32, device=dev)
    st_id = _native.PWW_STAT_MAX if stat == "max" else _native.PWW_STAT_STD
    args = (q.to(dev), k.to(dev), v.to(dev), H, D ** -0.5, None if w is None else w.to(dev),
            None if w is None else torch.tensor(idx, dtype=torch.int32, device=dev), st_id, gs)
    out, st = A.cross_attention(*args, return_stats=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        A.cross_attention(*args)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    # oracle per image
    refs, stats = [], []
    for b in range(B):
        box = {}

        def bias_fn(s, b=b):
            m = s.max() if stat == "max" else s.std()
            box["m"] = float(m)
            return gg * w[idx[b]] * m.float()
        refs.append(O.attention_core(q[b:b + 1].float(), k[b:b + 1].float(), v[b:b + 1].float(), H, D ** -0.5,
                                     bias_fn if idx[b] >= 0 else None, emulate_fp16=False))
        stats.append(box.get("m", 0.0))
    ref = torch.cat(refs, 0)
    got = out.float().cpu()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    per_image = [round((got[b] - ref[b]).abs().max().item() / ref.abs().max().item(), 5) for b in range(B)]
    print(json.dumps({"case": ci, "shape": CASES[ci][:4], "idx": idx[:4], "stat": stat, "grid": grid, "rel_err": round(err, 6),
                      "per_image": per_image[:4], "stats_got": None if st is None else [round(float(x), 4) for x in st[:4]],
                      "stats_ref": [round(x, 4) for x in stats[:4]], "us_per_call_eager": round(us, 1),
                      "nan": bool(torch.isnan(got).any())}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(int(sys.argv[1]))
    else:
        for ci in range(len(CASES)):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), str(ci)], capture_output=True, text=True,
                                   timeout=120)
                tail = (r.stdout.strip().splitlines() or [""])[-1]
                if r.returncode != 0:
                    err = [l for l in (r.stdout + r.stderr).splitlines() if "pww:" in l or "Error" in l or "error" in l]
                    print(json.dumps({"case": ci, "shape": CASES[ci][:4], "rc": r.returncode, "msg": err[:6]}), flush=True)
                else:
                    print(tail, flush=True)
            except subprocess.TimeoutExpired:
                print(json.dumps({"case": ci, "timeout": True}), flush=True)
