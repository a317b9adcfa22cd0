"""Which kernel variant every convolution of one CelebA score-net forward (or, with `train`, one DSM training step:
forward + data gradients) runs (BN, CTA pair, staged epilogue, pixel-major tiling, statically compiled epilogue mode):
python tools/list_conv_variants.py [batch] [train|openai]"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from score_based_multimodal_autoencoder_b200 import _lib as L, ops  # noqa: E402
from score_based_multimodal_autoencoder_b200 import unet_model  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
seen = collections.Counter()
orig = ops.conv_igemm


def traced(x, wpk, **kw):
    y = orig(x, wpk, **kw)
    v = L.lib().sbm_conv_last_variant()
    flags = "+".join(k + ("(pre)" if k == "out2" and kw.get("out2_preact") else "") for k in ("bias", "act", "residual", "stats", "out2", "rowbias", "gn_tab") if kw.get(k) is not None and not (isinstance(kw.get(k), int) and kw.get(k) == 0))
    od = "bf16" if (kw.get("out") is not None and kw["out"].dtype == torch.bfloat16) or kw.get("out_dtype") == torch.bfloat16 else "f32"
    key = (x.shape[1], kw["cin"], kw["cout"], kw["kh"], flags, od, v & 0xFFFF, bool(v & (1 << 16)), bool(v & (1 << 17)),
           bool(v & (1 << 18)), bool(v & (1 << 19)))
    seen[key] += 1
    return y


ops.conv_igemm = traced
unet_model.ops.conv_igemm = traced
torch.manual_seed(0)
m = unet_model.Unet(dim=256, channels=3, dim_mults=(1, 2, 2, 2, 2)).cuda().eval()
x = torch.randn(B, 3, 16, 16, device="cuda")
t = torch.rand(B, device="cuda")
if len(sys.argv) > 2 and sys.argv[2] == "openai":
    from score_based_multimodal_autoencoder_b200 import unet_openai
    unet_openai.ops.conv_igemm = traced
    mo = unet_openai.UNetModel(in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2, attention_resolutions=(),
                               dropout=0.1, channel_mult=(1, 2, 4, 8), num_heads=1, use_z=True, z_dim=512).cuda().eval()
    with torch.no_grad():
        mo(x, t * 999, torch.randn(B, 512, device="cuda"))
elif len(sys.argv) > 2 and sys.argv[2] == "train":
    from score_based_multimodal_autoencoder_b200 import autograd, sde_helper2 as sh
    autograd.ops.conv_igemm = traced
    m.train()
    loss = sh.loss_fn(x, m, sh.VPSDE(0.1, 20.0, 1000))
    loss.backward()
else:
    with torch.no_grad():
        m(x, t)
torch.cuda.synchronize()
print("H cin cout k flags out BN pair staged pixel-major static-epilogue count")
for key, n in sorted(seen.items(), key=lambda kv: (-kv[0][0], kv[0][1:])):
    print(*key, n)
