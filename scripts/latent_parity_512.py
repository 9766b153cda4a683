"""SURVEY 8c tolerance (iv): latents of the FULL-SIZE workload (SD1.5-shaped UNet, 512x512, aurora_1 map, CFG 7.5, LMS) after
K steps -- this repo's sampler (fp16 UNet, batched CFG, CUDA graph, native attention) against the restated reference loop
with the oracle attention (fp32, CPU, two batch-1 forwards per step) on the same seeded weights and inputs.
    python scripts/latent_parity_512.py [K=2]        (about 2 minutes of host CPU per step on the GPU box)"""
import json
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import paint_with_words_sd_b200 as P  # noqa: E402
from oracle import loop as oracle_loop  # noqa: E402
from paint_with_words_sd_b200.conditioning import _encode_text_color_inputs  # noqa: E402
from paint_with_words_sd_b200.pipeline import PwWSampler  # noqa: E402
from paint_with_words_sd_b200.scheduler import LMSDiscreteScheduler  # noqa: E402
from paint_with_words_sd_b200.synthetic import RandomTextEncoder, SimpleWordTokenizer  # noqa: E402
from paint_with_words_sd_b200.unet import UNetConfig, attention_modules, build_unet  # noqa: E402
from tests.fixtures import SETTINGS, color_map_image  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
WF = lambda w, sigma, qk: 0.4 * w * math.log(1 + sigma) * qk.max()   # noqa: E731


def setup(device):
    s = SETTINGS["aurora"]
    _, _, cond, uncond = _encode_text_color_inputs(RandomTextEncoder(768).to(device), SimpleWordTokenizer(), device,
                                                   color_map_image("aurora", 512), dict(s["ctx"]), s["prompt"], "")
    sch = LMSDiscreteScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear")
    sch.set_timesteps(30)
    lat = torch.randn(1, 4, 64, 64, generator=torch.manual_seed(0)) * sch.init_noise_sigma
    return cond, uncond, sch, lat


torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
t0 = time.time()
unet = build_unet(UNetConfig.sd15(), seed=0, dtype=torch.float16, device="cuda")
cond, uncond, sch, lat = setup("cuda")
P.patch_unet(unet)
try:
    got = PwWSampler(unet, sch, [cond], [uncond], lat.cuda(), WF, 7.5).run(K).float().cpu()
finally:
    P.unpatch_all()
t_gpu = time.time() - t0
t0 = time.time()
ref_unet = build_unet(UNetConfig.sd15(), seed=0)
cond, uncond, sch, lat = setup("cpu")
oracle_loop.patch_with_oracle(ref_unet)
try:
    ref = oracle_loop.reference_denoise_loop(ref_unet, sch, cond, uncond, lat, WF, 7.5, max_steps=K)
finally:
    cls = attention_modules(ref_unet)[0].__class__
    if "__call__" in cls.__dict__:
        delattr(cls, "__call__")
d = got - ref
print(json.dumps({"steps": K, "latent_shape": list(ref.shape), "max_abs": float(d.abs().max()), "rmse": float(d.pow(2).mean().sqrt()),
                  "ref_rms": float(ref.pow(2).mean().sqrt()), "rel_rmse": float(d.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()),
                  "seconds_gpu_arm": round(t_gpu, 1), "seconds_cpu_oracle": round(time.time() - t0, 1),
                  "threads": torch.get_num_threads()}))
