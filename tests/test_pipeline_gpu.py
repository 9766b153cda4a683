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


def test_inpaint_loop_matches_reference_loop():
    """paint_with_words_inpaint.py:230-266: the 9-channel input cat[latents, mask, masked-image latents] through the
    sampler (`extra_input`) against the restated reference loop on the same seeded tiny 9-channel UNet (fp32 CPU, two
    batch-1 forwards per step, oracle attention).  Weight 0.15 as runner_inpaint.py:72."""
    cfg = UNetConfig.tiny(in_channels=9)
    size, steps = 128, 4
    wf = lambda w, sigma, qk: 0.15 * w * math.log(1 + sigma) * qk.max()   # noqa: E731
    g = torch.Generator().manual_seed(5)
    mask = (torch.rand(1, 1, size // 8, size // 8, generator=g) > 0.5).float()
    masked = torch.randn(1, 4, size // 8, size // 8, generator=g) * 0.18215 * (1 - mask)
    extra = torch.cat([mask, masked], 1)
    unet_cpu = build_unet(cfg, seed=0)
    cond, uncond, sch, lat = _setup(cfg, size, steps, "cpu")
    try:
        oracle_loop.patch_with_oracle(unet_cpu)
        ref = oracle_loop.reference_denoise_loop(unet_cpu, sch, cond, uncond, lat, wf, extra_input=extra)
    finally:
        cls = attention_modules(unet_cpu)[0].__class__
        if "__call__" in cls.__dict__:
            delattr(cls, "__call__")
    unet = build_unet(cfg, seed=0, dtype=torch.float16, device="cuda")
    cond, uncond, sch, lat = _setup(cfg, size, steps, "cuda")
    try:
        P.patch_unet(unet)
        out = PwWSampler(unet, sch, [cond], [uncond], lat.cuda(), wf, 7.5, extra_input=extra.cuda()).run()
    finally:
        P.unpatch_all()
    out = out.float().cpu()
    rel_rmse = ((out - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    assert torch.isfinite(out).all() and rel_rmse < 3e-2, rel_rmse


def test_batched_images_match_solo_runs():
    """the per-image statistic makes an image's result independent of how images are grouped on
    a GPU (and therefore of how they are sharded over GPUs): two images in one sampler == each image alone.  The
    attention op itself is bit-identical across groupings (tests/test_xattn_gpu.py::test_batched_cfg_semantics_...);
    cuDNN / cuBLAS may pick different algorithms for batch 2 and batch 4, so the loop-level comparison allows fp16
    noise -- a statistic that leaked across images would be off by far more."""
    cfg = UNetConfig.tiny()
    size, steps = 128, 3
    unet = build_unet(cfg, seed=0, dtype=torch.float16, device="cuda")
    tok, enc = SimpleWordTokenizer(), RandomTextEncoder(cfg.cross_attention_dim).to("cuda")
    sch = LMSDiscreteScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear")
    sch.set_timesteps(steps)
    conds, unconds, lats = [], [], []
    for i, name in enumerate(("aurora", "cat_dog")):
        s = SETTINGS[name]
        _, _, c, u = C._encode_text_color_inputs(enc, tok, "cuda", color_map_image(name, size), dict(s["ctx"]),
                                                 s["prompt"], "")
        conds.append(c); unconds.append(u)
        lats.append(torch.randn(1, 4, size // 8, size // 8, generator=torch.manual_seed(i)) * sch.init_noise_sigma)
    try:
        P.patch_unet(unet)
        both = PwWSampler(unet, sch, conds, unconds, torch.cat(lats, 0).cuda(), WF, 7.5, use_graph=False).run().clone()
        solo = [PwWSampler(unet, sch, [conds[i]], [unconds[i]], lats[i].cuda(), WF, 7.5, use_graph=False).run().clone()
                for i in range(2)]
    finally:
        P.unpatch_all()
    for i in range(2):
        d = (both[i] - solo[i][0]).abs().max().item()
        # convolution / GEMM kernels may be chosen per batch size; anything beyond fp16 noise means the statistic leaked
        assert d <= 2e-2 * solo[i].abs().max().item(), (i, d)


def test_attn_processor_hook_matches_class_patch():
    """diffusers>=0.12 style: `PwWAttnProcessor()(attn, hidden_states, encoder_hidden_states=ctx_dict)` is the same
    computation as the class-level patch (paint_with_words.py:193-195)."""
    from paint_with_words_sd_b200.attention import PwWAttnProcessor, inj_forward
    from paint_with_words_sd_b200.unet import CrossAttention
    torch.manual_seed(0)
    attn = CrossAttention(320, 768, 8, 40).half().cuda()
    attn_self = CrossAttention(320, None, 8, 40).half().cuda()
    x = (torch.randn(1, 1024, 320) * 0.5).half().cuda()
    ctx = (torch.randn(1, 77, 768) * 0.5).cuda()
    w = torch.zeros(1024, 77)
    w[:512, 3] = 1.5
    w[256:, 9] = 0.4
    d = {"CONTEXT_TENSOR": ctx, "CROSS_ATTENTION_WEIGHT_1024": w.cuda(), "CROSS_ATTENTION_WEIGHT_ORIG": 0,
         "SIGMA": torch.tensor(3.0), "WEIGHT_FUNCTION": WF}
    with torch.no_grad():
        a = PwWAttnProcessor()(attn, x, encoder_hidden_states=d)
        b = inj_forward(attn, x, d)
        c = PwWAttnProcessor()(attn_self, x)                 # self-attention through the hook
        e = inj_forward(attn, x, ctx)                        # tensor context: no bias
    assert torch.equal(a, b) and a.shape == x.shape and torch.isfinite(c).all()
    assert not torch.equal(a, e)                             # the bias did something


def test_pipeline_classes_match_the_functional_api():
    """paint_with_words.py:513-842 / paint_with_words_inpaint.py:273-575: the pipeline classes are a thin surface over the
    functional API -- same latents for the same inputs."""
    s = SETTINGS["aurora"]
    tools = P.pww_load_tools("cuda:0", hf_model_path="synthetic:tiny")
    try:
        vae, unet, enc, tok, sch = tools
        pipe = P.PaintWithWord_StableDiffusionPipeline(vae=vae, text_encoder=enc, tokenizer=tok, unet=unet, scheduler=sch)
        out = pipe(s["prompt"], color_map_image=color_map_image("aurora", 128), color_context=dict(s["ctx"]),
                   weight_function=WF, num_inference_steps=3, seed=3, output_type="latent")
        ref = P.paint_with_words(color_context=dict(s["ctx"]), color_map_image=color_map_image("aurora", 128),
                                 input_prompt=s["prompt"], num_inference_steps=3, seed=3, device="cuda:0",
                                 weight_function=WF, preloaded_utils=(vae, unet, enc, tok, pipe.scheduler), return_latents=True)
        assert torch.equal(out.images, ref) and out.nsfw_content_detected is False
        img = pipe(s["prompt"], color_map_image=color_map_image("aurora", 128), color_context=dict(s["ctx"]),
                   weight_function=WF, num_inference_steps=2).images[0]
        assert img.size == (128, 128)
        unet9 = build_unet(UNetConfig.tiny(in_channels=9), seed=0, dtype=torch.float16, device="cuda")
        ipipe = P.PaintWithWord_StableDiffusionInpaintPipeline(vae=vae, text_encoder=enc, tokenizer=tok, unet=unet9)
        res = ipipe(s["prompt"], image=color_map_image("aurora", 128), mask_image=moon_mask_image(128),
                    color_map_image=color_map_image("aurora", 128), color_context=dict(s["ctx"]), weight_function=WF,
                    num_inference_steps=2, return_dict=False)
        assert res[0][0].size == (128, 128) and res[1] is False
    finally:
        P.unpatch_all()


def test_sampler_feeds_modules_their_own_dtype_and_follows_weight_updates():
    """the reference runs under torch.autocast; a diffusers fp16 UNet does not cast its input itself, so the
    sampler must hand it fp16 (a module that asserts its input dtype stands in for one).  And the fused projection
    weights the shim caches per module must follow an in-place weight update."""
    cfg = UNetConfig.tiny()
    unet = build_unet(cfg, seed=0, dtype=torch.float16, device="cuda")

    class Strict(torch.nn.Module):
        def __init__(self, inner):
            super().__init__()
            self.inner, self.in_channels = inner, inner.in_channels

        def forward(self, sample, t, encoder_hidden_states=None):
            assert sample.dtype == torch.float16, sample.dtype
            return self.inner(sample, t, encoder_hidden_states=encoder_hidden_states)

    cond, uncond, sch, lat = _setup(cfg, 128, 3, "cuda")
    try:
        P.patch_unet(unet)
        a = PwWSampler(Strict(unet), sch, [cond], [uncond], lat.cuda(), WF, 7.5, use_graph=False).run().clone()
        with torch.no_grad():
            for m in attention_modules(unet):
                m.to_q.weight.mul_(0.5)                       # in place: same storage, new version
        cond, uncond, sch, lat = _setup(cfg, 128, 3, "cuda")
        b = PwWSampler(Strict(unet), sch, [cond], [uncond], lat.cuda(), WF, 7.5, use_graph=False).run().clone()
    finally:
        P.unpatch_all()
    assert torch.isfinite(a).all() and torch.isfinite(b).all() and not torch.equal(a, b)
