"""Drive the self-attention kernel for ncu: python scripts/profile_selfattn.py [B] [N] [H] [D] [native|sdpa]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paint_with_words_sd_b200 import attention as A  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
H = int(sys.argv[3]) if len(sys.argv) > 3 else 8
D = int(sys.argv[4]) if len(sys.argv) > 4 else 40
A.SELF_ATTN_IMPL = sys.argv[5] if len(sys.argv) > 5 else "native"     # "sdpa": the library path the shim uses above 512 keys
g = torch.Generator().manual_seed(0)
q, k, v = [(torch.randn(B, N, H * D, generator=g) * 0.5).half().cuda() for _ in range(3)]
for _ in range(4):
    A.self_attention(q, k, v, H, D ** -0.5)
torch.cuda.synchronize()
print("done")
