"""ORACLE (test infrastructure, NOT product code).

Functional CPU fp32 restatement of the two score networks of the reference:
  * `unet_forward`        <- unet_model.py:189-323  (ConvNeXt `Unet`, the net every shipped command uses)
  * `unet_openai_forward` <- unet_openai.py:361-575 (guided-diffusion `UNetModel`)
Both take a plain `state_dict` (reference key names, Appendix D of SURVEY.md) so that the same
weights can be loaded into the reference modules, this oracle and the B200 modules.
Pinned against the real reference by oracle/gen_golden.py -> tests/golden/.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- ConvNeXt Unet
def _gn(x, sd, key, groups=1):
    return F.group_norm(x, groups, sd[key + ".weight"], sd[key + ".bias"], eps=1e-5)


def sinusoidal_embedding(time: torch.Tensor, dim: int) -> torch.Tensor:
    """unet_model.py:40-47 — sin block first, then cos; frequencies exp(-k*log(1e4)/(half-1))."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
    arg = time[:, None] * freq[None, :]
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


def convnext_block(sd, p, x, t_emb):
    """unet_model.py:115-124.  p = key prefix (e.g. 'downs.0.0')."""
    c = x.shape[1]
    h = F.conv2d(x, sd[p + ".ds_conv.weight"], sd[p + ".ds_conv.bias"], padding=3, groups=c)
    if (p + ".mlp.1.weight") in sd and t_emb is not None:
        cond = F.linear(F.gelu(t_emb), sd[p + ".mlp.1.weight"], sd[p + ".mlp.1.bias"])
        h = h + cond[:, :, None, None]
    h = _gn(h, sd, p + ".net.0")
    h = F.conv2d(h, sd[p + ".net.1.weight"], sd[p + ".net.1.bias"], padding=1)
    h = F.gelu(h)
    h = _gn(h, sd, p + ".net.3")
    h = F.conv2d(h, sd[p + ".net.4.weight"], sd[p + ".net.4.bias"], padding=1)
    if (p + ".res_conv.weight") in sd:
        res = F.conv2d(x, sd[p + ".res_conv.weight"], sd[p + ".res_conv.bias"])
    else:
        res = x
    return h + res


def resnet_block(sd, p, x, t_emb, groups=8):
    """unet_model.py:49-65, 67-90 (`Unet(use_convnext=False)`): Block = conv3x3 -> GroupNorm(groups) -> SiLU; the time
    projection (SiLU -> Linear) is added to Block 1's OUTPUT."""
    h = F.silu(_gn(F.conv2d(x, sd[p + ".block1.proj.weight"], sd[p + ".block1.proj.bias"], padding=1), sd,
                   p + ".block1.norm", groups))
    if (p + ".mlp.1.weight") in sd and t_emb is not None:
        h = F.linear(F.silu(t_emb), sd[p + ".mlp.1.weight"], sd[p + ".mlp.1.bias"])[:, :, None, None] + h
    h = F.silu(_gn(F.conv2d(h, sd[p + ".block2.proj.weight"], sd[p + ".block2.proj.bias"], padding=1), sd,
                   p + ".block2.norm", groups))
    if (p + ".res_conv.weight") in sd:
        return h + F.conv2d(x, sd[p + ".res_conv.weight"], sd[p + ".res_conv.bias"])
    return h + x


def linear_attention(sd, p, x, heads=4):
    """unet_model.py:162-177 (p = prefix of the LinearAttention module)."""
    b, c, hh, ww = x.shape
    qkv = F.conv2d(x, sd[p + ".to_qkv.weight"])
    q, k, v = qkv.chunk(3, dim=1)
    d = q.shape[1] // heads
    q, k, v = (a.reshape(b, heads, d, hh * ww) for a in (q, k, v))
    q = q.softmax(dim=-2) * d ** -0.5
    k = k.softmax(dim=-1)
    context = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", context, q).reshape(b, heads * d, hh, ww)
    out = F.conv2d(out, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    return _gn(out, sd, p + ".to_out.1")


def softmax_attention(sd, p, x, heads=4):
    """unet_model.py:135-149."""
    b, c, hh, ww = x.shape
    qkv = F.conv2d(x, sd[p + ".to_qkv.weight"])
    q, k, v = qkv.chunk(3, dim=1)
    d = q.shape[1] // heads
    q, k, v = (a.reshape(b, heads, d, hh * ww) for a in (q, k, v))
    q = q * d ** -0.5
    sim = torch.einsum("bhdi,bhdj->bhij", q, k)
    sim = sim - sim.amax(dim=-1, keepdim=True)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bhdj->bhid", attn, v)
    out = out.permute(0, 1, 3, 2).reshape(b, heads * d, hh, ww)
    return F.conv2d(out, sd[p + ".to_out.weight"], sd[p + ".to_out.bias"])


def _residual_prenorm(sd, p, x, fn):
    """Residual(PreNorm(dim, fn)) — unet_model.py:21-27, 179-187."""
    return fn(sd, p + ".fn.fn", _gn(x, sd, p + ".fn.norm")) + x


def unet_forward(sd: dict, x: torch.Tensor, time: torch.Tensor, *, dim: int, dim_mults=(1, 2, 4, 8),
                 use_convnext: bool = True, groups: int = 8) -> torch.Tensor:
    """unet_model.py:275-323, including the zero padding of non-power-of-two extents (:276-284) and the final crop
    (:318-322).  use_convnext=False: ResnetBlock(groups) in place of every ConvNextBlock (unet_model.py:214-217)."""
    block = convnext_block if use_convnext else (lambda sd_, p, x_, t_: resnet_block(sd_, p, x_, t_, groups))
    n_levels = len(dim_mults)
    pw = int((2 ** math.ceil(math.log2(x.shape[-1])) - x.shape[-1]) // 2)
    ph = int((2 ** math.ceil(math.log2(x.shape[-2])) - x.shape[-2]) // 2)
    if pw or ph:
        y = unet_forward(sd, F.pad(x, (pw, pw, ph, ph)), time, dim=dim, dim_mults=dim_mults,
                         use_convnext=use_convnext, groups=groups)
        y = y[..., pw:-pw] if pw else y
        return y[..., ph:-ph, :] if ph else y
    x = F.conv2d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=3)
    t = sinusoidal_embedding(time, dim)
    t = F.linear(t, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"])
    t = F.gelu(t)
    t = F.linear(t, sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])
    skips = []
    for lv in range(n_levels):
        x = block(sd, f"downs.{lv}.0", x, t)
        x = block(sd, f"downs.{lv}.1", x, t)
        x = _residual_prenorm(sd, f"downs.{lv}.2", x, linear_attention)
        skips.append(x)
        if lv < n_levels - 1:
            x = F.conv2d(x, sd[f"downs.{lv}.3.weight"], sd[f"downs.{lv}.3.bias"], stride=2, padding=1)
    x = block(sd, "mid_block1", x, t)
    x = _residual_prenorm(sd, "mid_attn", x, softmax_attention)
    x = block(sd, "mid_block2", x, t)
    for u in range(n_levels - 1):
        x = torch.cat((x, skips.pop()), dim=1)
        x = block(sd, f"ups.{u}.0", x, t)
        x = block(sd, f"ups.{u}.1", x, t)
        x = _residual_prenorm(sd, f"ups.{u}.2", x, linear_attention)
        # `is_last` at unet_model.py:257 is never true -> every up level upsamples
        x = F.conv_transpose2d(x, sd[f"ups.{u}.3.weight"], sd[f"ups.{u}.3.bias"], stride=2, padding=1)
    x = block(sd, "final_conv.0", x, None)
    return F.conv2d(x, sd["final_conv.1.weight"], sd["final_conv.1.bias"])


# --------------------------------------------------------------------------- OpenAI UNetModel
def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period=10000) -> torch.Tensor:
    """unet_openai.py:66-83 — cos block first, then sin; frequencies exp(-log(max_period)*k/half)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn32(x, sd, key):
    return F.group_norm(x.float(), 32, sd[key + ".weight"], sd[key + ".bias"], eps=1e-5)


def res_block(sd, p, x, emb):
    """unet_openai.py:291-305 (eval mode: dropout is the identity).  use_scale_shift_norm (:296-300) shows in the
    width of the embedding projection: 2 * out_channels = [scale | shift]."""
    h = F.conv2d(F.silu(_gn32(x, sd, p + ".in_layers.0")), sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"],
                 padding=1)
    e = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    if e.shape[1] == 2 * h.shape[1]:
        scale, shift = torch.chunk(e[:, :, None, None], 2, dim=1)
        h = F.silu(_gn32(h, sd, p + ".out_layers.0") * (1 + scale) + shift)
    else:
        h = F.silu(_gn32(h + e[:, :, None, None], sd, p + ".out_layers.0"))
    h = F.conv2d(h, sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if (p + ".skip_connection.weight") in sd:
        w = sd[p + ".skip_connection.weight"]
        x = F.conv2d(x, w, sd[p + ".skip_connection.bias"], padding=w.shape[-1] // 2)
    return x + h


def attention_block(sd, p, x, num_heads):
    """unet_openai.py:329-358."""
    b, c, hh, ww = x.shape
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(_gn32(xf, sd, p + ".norm"), sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    qkv = qkv.reshape(b * num_heads, -1, qkv.shape[2])
    ch = qkv.shape[1] // 3
    q, k, v = torch.split(qkv, ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    w = torch.softmax(w.float(), dim=-1)
    h = torch.einsum("bts,bcs->bct", w, v).reshape(b, -1, xf.shape[-1])
    h = F.conv1d(h, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return (xf + h).reshape(b, c, hh, ww)


def unet_openai_forward(sd: dict, x, timesteps, *, model_channels, num_res_blocks, attention_resolutions,
                        channel_mult=(1, 2, 4, 8), num_heads=1, z=None, y=None):
    """unet_openai.py:538-575 (dims=2, conv_resample=True; y: class labels when the net has a `label_emb`)."""
    emb = timestep_embedding(timesteps, model_channels)
    emb = F.linear(F.silu(F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])),
                   sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    if z is not None:
        zp = F.linear(F.silu(F.linear(z, sd["proj.0.weight"], sd["proj.0.bias"])), sd["proj.2.weight"],
                      sd["proj.2.bias"])
        emb = emb + zp
    if y is not None:
        emb = emb + sd["label_emb.weight"][y]   # unet_openai.py:561-564
    hs = []
    h = F.conv2d(x, sd["input_blocks.0.0.weight"], sd["input_blocks.0.0.bias"], padding=1)
    hs.append(h)
    idx, ds = 1, 1
    for level, _ in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            h = res_block(sd, f"input_blocks.{idx}.0", h, emb)
            if ds in attention_resolutions:
                h = attention_block(sd, f"input_blocks.{idx}.1", h, num_heads)
            hs.append(h)
            idx += 1
        if level != len(channel_mult) - 1:
            h = F.conv2d(h, sd[f"input_blocks.{idx}.0.op.weight"], sd[f"input_blocks.{idx}.0.op.bias"], stride=2,
                         padding=1)
            hs.append(h)
            idx += 1
            ds *= 2
    h = res_block(sd, "middle_block.0", h, emb)
    h = attention_block(sd, "middle_block.1", h, num_heads)
    h = res_block(sd, "middle_block.2", h, emb)
    idx = 0
    for level, _ in list(enumerate(channel_mult))[::-1]:
        for i in range(num_res_blocks + 1):
            h = torch.cat([h, hs.pop()], dim=1)
            h = res_block(sd, f"output_blocks.{idx}.0", h, emb)
            nxt = 1
            if ds in attention_resolutions:
                h = attention_block(sd, f"output_blocks.{idx}.1", h, num_heads)
                nxt = 2
            if level and i == num_res_blocks:
                h = F.interpolate(h, scale_factor=2, mode="nearest")
                h = F.conv2d(h, sd[f"output_blocks.{idx}.{nxt}.conv.weight"], sd[f"output_blocks.{idx}.{nxt}.conv.bias"],
                             padding=1)
                ds //= 2
            idx += 1
    h = F.silu(_gn32(h, sd, "out.0"))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


def unet_param_shapes(dim: int, channels: int, dim_mults=(1, 2, 4, 8), init_dim=None, out_dim=None, mult=2,
                      heads=4, dim_head=32) -> dict:
    """Parameter names/shapes of unet_model.Unet (unet_model.py:203-273), so the oracle can be driven without
    instantiating any module."""
    init_dim = init_dim if init_dim is not None else dim // 3 * 2
    out_dim = out_dim if out_dim is not None else channels
    tdim = dim * 4
    hid = heads * dim_head
    s = {"init_conv.weight": (init_dim, channels, 7, 7), "init_conv.bias": (init_dim,),
         "time_mlp.1.weight": (tdim, dim), "time_mlp.1.bias": (tdim,),
         "time_mlp.3.weight": (tdim, tdim), "time_mlp.3.bias": (tdim,)}

    def block(p, cin, cout, temb=True):
        if temb:
            s[p + ".mlp.1.weight"], s[p + ".mlp.1.bias"] = (cin, tdim), (cin,)
        s[p + ".ds_conv.weight"], s[p + ".ds_conv.bias"] = (cin, 1, 7, 7), (cin,)
        s[p + ".net.0.weight"], s[p + ".net.0.bias"] = (cin,), (cin,)
        s[p + ".net.1.weight"], s[p + ".net.1.bias"] = (cout * mult, cin, 3, 3), (cout * mult,)
        s[p + ".net.3.weight"], s[p + ".net.3.bias"] = (cout * mult,), (cout * mult,)
        s[p + ".net.4.weight"], s[p + ".net.4.bias"] = (cout, cout * mult, 3, 3), (cout,)
        if cin != cout:
            s[p + ".res_conv.weight"], s[p + ".res_conv.bias"] = (cout, cin, 1, 1), (cout,)

    def lin_attn(p, c):
        s[p + ".fn.norm.weight"], s[p + ".fn.norm.bias"] = (c,), (c,)
        s[p + ".fn.fn.to_qkv.weight"] = (3 * hid, c, 1, 1)
        s[p + ".fn.fn.to_out.0.weight"], s[p + ".fn.fn.to_out.0.bias"] = (c, hid, 1, 1), (c,)
        s[p + ".fn.fn.to_out.1.weight"], s[p + ".fn.fn.to_out.1.bias"] = (c,), (c,)

    dims = [init_dim] + [dim * m for m in dim_mults]
    in_out = list(zip(dims[:-1], dims[1:]))
    for lv, (ci, co) in enumerate(in_out):
        block(f"downs.{lv}.0", ci, co)
        block(f"downs.{lv}.1", co, co)
        lin_attn(f"downs.{lv}.2", co)
        if lv < len(in_out) - 1:
            s[f"downs.{lv}.3.weight"], s[f"downs.{lv}.3.bias"] = (co, co, 4, 4), (co,)
    mid = dims[-1]
    block("mid_block1", mid, mid)
    s["mid_attn.fn.norm.weight"], s["mid_attn.fn.norm.bias"] = (mid,), (mid,)
    s["mid_attn.fn.fn.to_qkv.weight"] = (3 * hid, mid, 1, 1)
    s["mid_attn.fn.fn.to_out.weight"], s["mid_attn.fn.fn.to_out.bias"] = (mid, hid, 1, 1), (mid,)
    block("mid_block2", mid, mid)
    for u, (ci, co) in enumerate(reversed(in_out[1:])):
        block(f"ups.{u}.0", co * 2, ci)
        block(f"ups.{u}.1", ci, ci)
        lin_attn(f"ups.{u}.2", ci)
        s[f"ups.{u}.3.weight"], s[f"ups.{u}.3.bias"] = (ci, ci, 4, 4), (ci,)
    block("final_conv.0", dim, dim, temb=False)
    s["final_conv.1.weight"], s["final_conv.1.bias"] = (out_dim, dim, 1, 1), (out_dim,)
    return s
