"""Kineto/CUPTI kernel-time table of eager denoising steps (warm, back-to-back) -- complements the ncu launch list."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import paint_with_words_sd_b200 as P  # noqa: E402
from paint_with_words_sd_b200.conditioning import _encode_text_color_inputs  # noqa: E402
from paint_with_words_sd_b200.pipeline import PwWSampler  # noqa: E402
from paint_with_words_sd_b200.scheduler import LMSDiscreteScheduler  # noqa: E402
from paint_with_words_sd_b200.synthetic import RandomTextEncoder, SimpleWordTokenizer  # noqa: E402
from paint_with_words_sd_b200.unet import UNetConfig, build_unet  # noqa: E402
from tests.fixtures import SETTINGS, color_map_image  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
unet = build_unet(UNetConfig.sd15(), seed=0, dtype=torch.float16, device=dev)
P.patch_unet(unet)
s = SETTINGS["aurora"]
_, _, cond, uncond = _encode_text_color_inputs(RandomTextEncoder(768).to(dev), SimpleWordTokenizer(), dev,
                                               color_map_image("aurora", 512), dict(s["ctx"]), s["prompt"], "")
sch = LMSDiscreteScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear")
sch.set_timesteps(30)
lat = (torch.randn(1, 4, 64, 64, generator=torch.manual_seed(0)) * sch.init_noise_sigma).to(dev)
smp = PwWSampler(unet, sch, [cond], [uncond], lat, bench.weight_function, 7.5, use_graph=False)
with torch.no_grad():
    for _ in range(4):
        smp.step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            smp.step()
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
