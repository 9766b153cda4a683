"""Repeat the self-attention kernel on multi-wave grids; report the first failing iteration (each config in its own process)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CASES = [(1, 9216, 5, 64, 150), (2, 9216, 5, 64, 80), (4, 4096, 8, 40, 60), (8, 1024, 8, 80, 60), (2, 2304, 10, 64, 100), (16, 256, 8, 160, 60)]
if os.environ.get("PWW_STRESS_ONLY"):
    CASES = [CASES[int(i)] for i in os.environ["PWW_STRESS_ONLY"].split(",")]


def run(ci):
    import torch
    from paint_with_words_sd_b200 import attention as A
    B, N, H, D, iters = CASES[ci]
    g = torch.Generator().manual_seed(ci)
    C = H * D
    q, k, v = [(torch.randn(B, N, C, generator=g) * 0.5).half().cuda() for _ in range(3)]
    ref = None
    for it in range(iters):
        o = A.self_attention(q, k, v, H, D ** -0.5)
        torch.cuda.synchronize()
        if ref is None:
            ref = o.clone()
        elif not torch.equal(o, ref):
            print(json.dumps({"case": CASES[ci], "iter": it, "mismatch": float((o.float() - ref.float()).abs().max())}), flush=True)
    print(json.dumps({"case": CASES[ci], "ok": True}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(int(sys.argv[1]))
    else:
        for ci in range(len(CASES)):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), str(ci)], capture_output=True, text=True, timeout=200)
                lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
                to = sorted(set(l for l in (r.stdout + r.stderr).splitlines() if "pww-attn: timeout" in l))
                print(json.dumps({"case": CASES[ci], "rc": r.returncode, "out": lines[-3:], "timeouts": to[:16]}), flush=True)
            except subprocess.TimeoutExpired:
                print(json.dumps({"case": CASES[ci], "timeout": True}), flush=True)
