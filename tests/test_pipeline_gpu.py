"""Loop-level parity on the GPU: PwWSampler (batched CFG, CUDA graph, native attention, fp16 UNet) against the
restated reference loop (two batch-1 forwards, oracle attention, fp32 CPU) on the same seeded tiny UNet."""
import math

import pytest
import torch

import paint_with_words_sd_b200 as P
from oracle import loop as oracle_loop
from paint_with_words_sd_b200 import conditioning as C
from paint_with_words_sd_b200.pipeline import PwWSampler
from paint_with_words_sd_b200.scheduler import LMSDiscreteScheduler
from paint_with_words_sd_b200.synthetic import RandomTextEncoder, SimpleWordTokenizer
from paint_with_words_sd_b200.unet import UNetConfig, attention_modules, build_unet
from tests.fixtures import SETTINGS, color_map_image, moon_mask_image

pytestmark = pytest.mark.gpu
WF = lambda w, sigma, qk: 0.4 * w * math.log(1 + sigma) * qk.max()   # runner.py:104


def _setup(cfg, size, steps, device):
    tok, enc = SimpleWordTokenizer(), RandomTextEncoder(cfg.cross_attention_dim)
    s = SETTINGS["aurora"]
    _, _, cond, uncond = C._encode_text_color_inputs(enc.to(device), tok, device, color_map_image("aurora", size),
                                                     dict(s["ctx"]), s["prompt"], "")
    sch = LMSDiscreteScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear")
    sch.set_timesteps(steps)
    lat = torch.randn(1, 4, size // 8, size // 8, generator=torch.manual_seed(0)) * sch.init_noise_sigma
    return cond, uncond, sch, lat


def _cpu_reference(cfg, size, steps, extra=None):
    unet = build_unet(cfg, seed=0)
    cond, uncond, sch, lat = _setup(cfg, size, steps, "cpu")
    try:
        oracle_loop.patch_with_oracle(unet)
        return oracle_loop.reference_denoise_loop(unet, sch, cond, uncond, lat, WF, extra_input=extra)
    finally:
        cls = attention_modules(unet)[0].__class__
        if "__call__" in cls.__dict__:
            delattr(cls, "__call__")


@pytest.mark.parametrize("use_graph", [False, True])
def test_sampler_matches_reference_loop(use_graph):
    cfg = UNetConfig.tiny()          # widths (160,320,640,640), 4 heads -> head dims 40/80/160 (kernel family)
    size, steps = 128, 4
    ref = _cpu_reference(cfg, size, steps)
    unet = build_unet(cfg, seed=0, dtype=torch.float16, device="cuda")
    cond, uncond, sch, lat = _setup(cfg, size, steps, "cuda")
    try:
        P.patch_unet(unet)
        out = PwWSampler(unet, sch, [cond], [uncond], lat.cuda(), WF, 7.5, use_graph=use_graph).run()
    finally:
        P.unpatch_all()
    out = out.float().cpu()
    rel_rmse = ((out - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    assert torch.isfinite(out).all() and rel_rmse < 3e-2, rel_rmse


def test_public_api_txt2img_and_inpaint():
    tools = list(P.pww_load_tools("cuda:0", hf_model_path="synthetic:tiny"))
    try:
        s = SETTINGS["aurora"]
        img = P.paint_with_words(color_context=dict(s["ctx"]), color_map_image=color_map_image("aurora", 128),
                                 input_prompt=s["prompt"], num_inference_steps=3, device="cuda:0",
                                 weight_function=WF, preloaded_utils=tuple(tools))
        assert img.size == (128, 128)
        tools[1] = build_unet(UNetConfig.tiny(in_channels=9), seed=0, dtype=torch.float16, device="cuda")
        init = color_map_image("aurora", 128)
        img = P.paint_with_words_inpaint(color_context=dict(s["ctx"]), color_map_image=color_map_image("aurora", 128),
                                         mask_image=moon_mask_image(128), init_image=init, input_prompt=s["prompt"],
                                         num_inference_steps=3, device="cuda:0", weight_function=WF,
                                         preloaded_utils=tuple(tools))
        assert img.size == (128, 128)
    finally:
        P.unpatch_all()
