"""Fused channels-last UNet ops (GroupNorm[+add][+SiLU], GEGLU) against plain PyTorch fp32 on the same inputs."""
import pytest
import torch
import torch.nn.functional as F

from paint_with_words_sd_b200 import fused_ops
from paint_with_words_sd_b200.unet import UNetConfig, build_unet

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C,HW,G", [(320, 64, 32), (640, 32, 32), (1280, 16, 32), (1280, 8, 32), (2560, 8, 32),
                                    (1920, 16, 32), (960, 32, 32), (160, 16, 8), (480, 7, 8)])
@pytest.mark.parametrize("silu,with_add", [(True, False), (True, True), (False, False)])
def test_group_norm_nhwc(C, HW, G, silu, with_add):
    g = torch.Generator().manual_seed(C + HW)
    B = 2
    x = (torch.randn(B, C, HW, HW, generator=g) * 1.5 + 0.3).half()
    gn = torch.nn.GroupNorm(G, C, eps=1e-5)
    gn.weight.data = torch.randn(C, generator=g) * 0.5 + 1.0
    gn.bias.data = torch.randn(C, generator=g) * 0.2
    add = (torch.randn(B, C, generator=g) * 0.5).half() if with_add else None
    xin = x.float() + (add.float()[:, :, None, None] if with_add else 0.0)
    ref = F.group_norm(xin, G, gn.weight.float(), gn.bias.float(), 1e-5)
    if silu:
        ref = F.silu(ref)
    gn_h = gn.half().cuda()
    got = fused_ops.group_norm_nhwc(x.cuda().contiguous(memory_format=torch.channels_last), gn_h,
                                    None if add is None else add.cuda(), silu=silu)
    assert got.is_contiguous(memory_format=torch.channels_last)
    # reference with fp16-rounded affine parameters, like the kernel sees them
    ref = F.group_norm(xin, G, gn_h.weight.float().cpu(), gn_h.bias.float().cpu(), 1e-5)
    if silu:
        ref = F.silu(ref)
    err = (got.float().cpu() - ref).abs().max().item()
    assert err <= 4e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("M,I", [(2 * 4096, 1280), (2 * 64, 5120), (3, 8), (77, 2560)])
def test_geglu(M, I):
    g = torch.Generator().manual_seed(M + I)
    h = (torch.randn(M, 2 * I, generator=g) * 2.0).half()
    ref = h[:, :I].float() * F.gelu(h[:, I:].float())
    got = fused_ops.geglu(h.cuda()).float().cpu()
    assert (got - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("M,C", [(2 * 4096, 320), (2 * 1024, 640), (2 * 256, 1280), (5, 1280), (3, 8), (7, 2048)])
@pytest.mark.parametrize("with_res", [True, False])
def test_add_layer_norm(M, C, with_res):
    g = torch.Generator().manual_seed(M + C)
    x = (torch.randn(M, C, generator=g) * 2.0).half()
    res = (torch.randn(M, C, generator=g) * 2.0).half() if with_res else None
    ln = torch.nn.LayerNorm(C)
    ln.weight.data = torch.randn(C, generator=g) * 0.5 + 1.0
    ln.bias.data = torch.randn(C, generator=g) * 0.2
    ln_h = ln.half().cuda()
    s_ref = (x.float() + res.float()).half() if with_res else x
    y_ref = F.layer_norm(s_ref.float(), (C,), ln_h.weight.float().cpu(), ln_h.bias.float().cpu(), ln.eps)
    s, y = fused_ops.add_layer_norm(x.cuda(), None if res is None else res.cuda(), ln_h)
    assert torch.equal(s.cpu(), s_ref)                      # the residual stream is bit-identical to an fp16 add
    assert (y.float().cpu() - y_ref).abs().max().item() <= 4e-3 * max(1.0, y_ref.abs().max().item())


def test_unet_fast_route_matches_plain_route():
    """Same fp16 weights: channels-last fused route vs the module-by-module PyTorch route (stock attention)."""
    cfg = UNetConfig.tiny()
    unet = build_unet(cfg, seed=0, dtype=torch.float16, device="cuda")
    x = torch.randn(2, 4, 16, 16, generator=torch.manual_seed(0)).cuda()
    ctx = torch.randn(2, 77, cfg.cross_attention_dim, generator=torch.manual_seed(1)).cuda().half()
    t = torch.tensor([500.0], device="cuda")
    with torch.no_grad():
        fast = unet(x, t, encoder_hidden_states=ctx).sample.float()
        orig = fused_ops.is_fast
        fused_ops.is_fast = lambda _x: False
        try:
            plain = unet(x, t, encoder_hidden_states=ctx).sample.float()
        finally:
            fused_ops.is_fast = orig
    rel = ((fast - plain).pow(2).mean().sqrt() / plain.pow(2).mean().sqrt()).item()
    assert rel < 1e-2, rel
    # leave no per-forward state behind for other tests
    from paint_with_words_sd_b200.unet import _resnets
    assert all(hasattr(r, "_pww_t") for r in _resnets(unet))
