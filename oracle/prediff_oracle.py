"""CPU oracle for the PreDiff sampling path - TEST INFRASTRUCTURE ONLY.

A plain PyTorch fp32 (CPU) functional restatement of the reference algorithm for the hot path of
gaozhihan/PreDiff. It is NOT part of the product: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it, and only as the checker / reported CPU baseline.
The product path (prediff_b200/) never imports this module and has no CPU fallback.

Pinning: the reference has no tests or golden vectors of its own, so this restatement is pinned
against outputs of the UNMODIFIED reference modules run in the build container - the fixtures under
tests/golden/*.npz, produced by tests/golden/gen_golden.py (which imports /root/reference/src with a sys.modules
stub for the absent `lightning`/`diffusers` packages) - plus the schedule known-answers of SURVEY.md section 4.
tests/test_oracle.py checks every function here against those fixtures.

Every function takes a `sd` dict (reference state_dict key -> torch fp32 tensor) and cites the reference lines
it restates. Paths are relative to /root/reference/.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------- schedule


def make_schedule(timesteps=1000, linear_start=1e-4, linear_end=2e-2):
    """register_schedule (src/prediff/diffusion/latent_diffusion.py:228-278) with the "linear" beta schedule
    (src/prediff/diffusion/utils.py:17-22): float64 math, fp32 buffers. v_posterior = 0."""
    betas = np.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=np.float64) ** 2
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    buf = {
        "betas": betas,
        "alphas_cumprod": ac,
        "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": np.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": np.sqrt(1.0 - ac),
        "log_one_minus_alphas_cumprod": np.log(1.0 - ac),
        "sqrt_recip_alphas_cumprod": np.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": np.sqrt(1.0 / ac - 1),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": np.log(np.maximum(post_var, 1e-20)),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    }
    return {k: torch.tensor(v, dtype=torch.float32) for k, v in buf.items()}


def ddim_timesteps(num_ddim, num_ddpm=1000):
    """make_ddim_timesteps('uniform', ...) (src/prediff/diffusion/utils.py:42-56): range(0, T, T//n) + 1."""
    c = num_ddpm // num_ddim
    return np.asarray(list(range(0, num_ddpm, c))) + 1


def ddim_params(alphas_cumprod, ts, eta):
    """make_ddim_sampling_parameters (src/prediff/diffusion/utils.py:59-70). alphas_cumprod: fp32 numpy."""
    a = alphas_cumprod[ts]
    a_prev = np.asarray([alphas_cumprod[0]] + alphas_cumprod[ts[:-1]].tolist())
    sig = eta * np.sqrt((1 - a_prev) / (1 - a) * (1 - a / a_prev))
    return sig, a, a_prev


# ----------------------------------------------------------------------------------------------- UNet pieces


def timestep_embedding(t, dim, max_period=10000):
    """src/prediff/models/utils.py:68-83 (cos first, then sin)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _res_block3d(sd, p, x, t_emb, groups_in, groups_out):
    """TimeEmbedResBlock.forward (src/prediff/models/time_embed.py:134-169), dims=3, no scale-shift, no up/down.
    x: (B, C, T, H, W)."""
    h = F.group_norm(x, groups_in, sd[f"{p}.in_layers.0.weight"], sd[f"{p}.in_layers.0.bias"], 1e-5)
    h = F.conv3d(F.silu(h), sd[f"{p}.in_layers.2.weight"], sd[f"{p}.in_layers.2.bias"], padding=1)
    if f"{p}.emb_layers.1.weight" in sd:
        e = F.linear(F.silu(t_emb), sd[f"{p}.emb_layers.1.weight"], sd[f"{p}.emb_layers.1.bias"])
        h = h + e[:, :, None, None, None]
    h = F.group_norm(h, groups_out, sd[f"{p}.out_layers.0.weight"], sd[f"{p}.out_layers.0.bias"], 1e-5)
    h = F.conv3d(F.silu(h), sd[f"{p}.out_layers.3.weight"], sd[f"{p}.out_layers.3.bias"], padding=1)
    if f"{p}.skip_connection.weight" in sd:
        x = F.conv3d(x, sd[f"{p}.skip_connection.weight"], sd[f"{p}.skip_connection.bias"])
    return x + h


def _gn_groups(c, g=32):
    # time_embed.py:90,116: norm_groups if channels % norm_groups == 0 else channels
    return g if c % g == 0 else c


def axial_attention(sd, p, x, heads, axis):
    """CuboidSelfAttentionLayer.forward (src/prediff/models/cuboid_transformer/cuboid_transformer.py:812-966)
    for an axial cuboid ((T,1,1), (1,H,1) or (1,1,W); strategy 'l', no shift, no padding, no global vectors):
    LN -> qkv -> per-line softmax(q k^T / sqrt(hd) + rel-pos bias) v -> proj. x: (B, T, H, W, C)."""
    B, T, H, W, C = x.shape
    hd = C // heads
    y = F.layer_norm(x, (C,), sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], 1e-5)
    qkv = F.linear(y, sd[f"{p}.qkv.weight"]).view(B, T, H, W, 3, heads, hd)
    dim = 1 + axis
    L = x.shape[dim]
    q, k, v = (qkv[:, :, :, :, i].movedim(dim, 4) for i in range(3))  # (B, o1, o2, heads, L, hd)
    s = (q * hd ** -0.5) @ k.transpose(-1, -2)
    idx = torch.arange(L)
    rel = idx[:, None] - idx[None, :] + L - 1  # cuboid_transformer.py:719-734 reduced to one axis
    s = s + sd[f"{p}.relative_position_bias_table"][rel].permute(2, 0, 1)
    o = (torch.softmax(s, dim=-1) @ v).movedim(4, dim).reshape(B, T, H, W, C)
    return F.linear(o, sd[f"{p}.proj.weight"], sd[f"{p}.proj.bias"])


def ffn(sd, p, x):
    """PositionwiseFFN.forward, pre-norm, GELU(erf) (cuboid_transformer.py:182-208)."""
    C = x.shape[-1]
    y = F.layer_norm(x, (C,), sd[f"{p}.layer_norm.weight"], sd[f"{p}.layer_norm.bias"], 1e-5)
    y = F.gelu(F.linear(y, sd[f"{p}.ffn_1.weight"], sd[f"{p}.ffn_1.bias"]))
    return F.linear(y, sd[f"{p}.ffn_2.weight"], sd[f"{p}.ffn_2.bias"]) + x


def _to_cuboids(x, size, strategy):
    """cuboid_reorder (cuboid_transformer.py:388-429): (B, T, H, W, C) -> (B, num_cuboids, volume, C). Per axis the
    padded length splits as (blocks, in-block) for 'l' and (in-block, blocks) for 'd'."""
    B, C = x.shape[0], x.shape[-1]
    shape, blk, inn = [B], [], []
    for a in range(3):
        n = x.shape[1 + a] // size[a]
        shape += [n, size[a]] if strategy[a] == "l" else [size[a], n]
        blk.append(1 + 2 * a + (0 if strategy[a] == "l" else 1))
        inn.append(1 + 2 * a + (1 if strategy[a] == "l" else 0))
    y = x.reshape(shape + [C]).permute([0] + blk + inn + [7])
    return y.reshape(B, -1, size[0] * size[1] * size[2], C)


def _from_cuboids(y, size, strategy, padded):
    """cuboid_reorder_reverse (cuboid_transformer.py:432-467)."""
    B, C = y.shape[0], y.shape[-1]
    n = [padded[a] // size[a] for a in range(3)]
    y = y.reshape([B] + n + list(size) + [C])
    perm = [0]
    for a in range(3):
        perm += [1 + a, 4 + a] if strategy[a] == "l" else [4 + a, 1 + a]
    return y.permute(perm + [7]).reshape(B, padded[0], padded[1], padded[2], C)


def cuboid_attention_core(qkv, table, heads, size0, strategy, shift0, padding_type="zeros", gqkv=None,
                          global_self_attn=False, sep=None):
    """The attention core of CuboidSelfAttentionLayer.forward (cuboid_transformer.py:812-966; global vectors: below)
    from the per-token q|k|v rows: qkv (B, T, H, W, 3C), taken BEFORE padding (qkv has no bias, so the zero rows the
    reference pads after its LayerNorm map to zero q, k, v) -> (B, T, H, W, C) before the final projection.
    Cuboid size / shift update :563-592; zero padding at the end of each axis; roll by -shift; reorder; mask =
    same shifted-window region (and, for 'ignore', both tokens real) :470-528; bias = table[index[:vol, :vol]]
    with the index built from the constructor's cuboid size :714-734, 855-861; masked_softmax :531-560.
    gqkv (B, K, 3C): q|k|v rows of the K global vectors (use_global_vector with the shared global_qkv net, :893-901) ->
    returns (o, new_global (B, K, C) before global_proj). Local queries see the K global keys as extra, never-masked
    columns (:902-913); the global queries attend over ALL num_cuboids * volume slots in cuboid order (+ the global keys
    with global_self_attn), and under 'ignore' padding the slot mask is the padded, rolled validity grid flattened in
    RASTER order - the reference applies it to the cuboid-ordered keys as it is (:915-945), so does this.
    sep (separate_global_qkv=True, :866-891): (tok (B, T, H, W, 3C) = l2g_q | g2l_k | g2l_v rows of the tokens,
    grow (B, K, 3C or 6C) = l2g_k | l2g_v | g2l_q [| g2g_q | g2g_k | g2g_v] rows of the global vectors) instead of gqkv: the
    tokens meet the global keys with their own query net, the global queries meet the tokens' own key / value nets, and the
    global self-attention has a third q|k|v set."""
    B, T, H, W, C3 = qkv.shape
    C, hd = C3 // 3, C3 // 3 // heads
    dims = (T, H, W)
    size, shift = list(size0), list(shift0)
    for a in range(3):
        if strategy[a] == "d":
            shift[a] = 0
        if dims[a] <= size[a]:
            size[a], shift[a] = dims[a], 0
    pad = [(size[a] - dims[a] % size[a]) % size[a] for a in range(3)]
    padded = [dims[a] + pad[a] for a in range(3)]
    real = torch.ones(1, T, H, W, 1)
    region = torch.zeros(1, *padded, 1)
    for a in range(3):  # three slices per axis, last assignment wins (a zero shift leaves a single region)
        idx = torch.arange(padded[a])
        lab = torch.zeros(padded[a])
        lab[idx >= padded[a] - size[a]] = 1
        if shift[a] > 0:
            lab[idx >= padded[a] - shift[a]] = 2
        else:
            lab[:] = 2
        view = [1, 1, 1, 1, 1]
        view[1 + a] = padded[a]
        region = region * 3 + lab.view(view)
    if sep is not None:   # the tokens' extra rows travel through the same padding / roll / reorder
        qkv = torch.cat([qkv, sep[0]], dim=-1)
    if padding_type == "nearest" and any(pad):   # _generalize_padding (models/utils.py:228-258): resample to the padded size
        y = F.interpolate(qkv.permute(0, 4, 1, 2, 3), size=tuple(padded)).permute(0, 2, 3, 4, 1)
    else:
        y = F.pad(qkv, (0, 0, 0, pad[2], 0, pad[1], 0, pad[0]))
    real = F.pad(real, (0, 0, 0, pad[2], 0, pad[1], 0, pad[0]))
    if any(s > 0 for s in shift):
        y = torch.roll(y, shifts=[-s for s in shift], dims=(1, 2, 3))
        real = torch.roll(real, shifts=[-s for s in shift], dims=(1, 2, 3))
    y = _to_cuboids(y, size, strategy)                      # (B, nc, vol, 3C)
    nc, vol = y.shape[1], y.shape[2]
    region = _to_cuboids(region, size, strategy)[0, :, :, 0]  # (nc, vol)
    mask = region[:, :, None] == region[:, None, :]
    if padding_type == "ignore":
        ok = _to_cuboids(real, size, strategy)[0, :, :, 0] > 0
        mask = mask & ok[:, :, None] & ok[:, None, :]
    if sep is not None:
        q, k, v, q_lg, k_gl, v_gl = y.reshape(B, nc, vol, 6, heads, hd).permute(3, 0, 4, 1, 2, 5)
    else:
        q, k, v = y.reshape(B, nc, vol, 3, heads, hd).permute(3, 0, 4, 1, 2, 5)
        q_lg, k_gl, v_gl = q, k, v
    s = (q * hd ** -0.5) @ k.transpose(-1, -2)                # (B, heads, nc, vol, vol)
    bt, bh, bw = size0
    i = torch.arange(vol)
    ct, ch, cw = i // (bh * bw), (i // bw) % bh, i % bw
    rel = ((ct[:, None] - ct[None] + bt - 1) * (2 * bh - 1) + (ch[:, None] - ch[None] + bh - 1)) * (2 * bw - 1) \
        + (cw[:, None] - cw[None] + bw - 1)
    s = s + table[rel].permute(2, 0, 1)[:, None]
    new_g = None
    if gqkv is not None or sep is not None:
        if sep is not None:
            grow = sep[1]
            K = grow.shape[1]
            parts = grow.reshape(B, 1, K, grow.shape[-1] // C, heads, hd).permute(3, 0, 4, 1, 2, 5)   # (n, B, heads, 1, K, hd)
            gk, gv, gq = parts[0], parts[1], parts[2]
            gq_g, gk_g, gv_g = (parts[3], parts[4], parts[5]) if global_self_attn else (None, None, None)
        else:
            K = gqkv.shape[1]
            gq, gk, gv = gqkv.reshape(B, 1, K, 3, heads, hd).permute(3, 0, 4, 1, 2, 5)   # (B, heads, 1, K, hd)
            gq_g, gk_g, gv_g = gq, gk, gv
        s = torch.cat([s, (q_lg * hd ** -0.5) @ gk.transpose(-1, -2)], dim=-1)       # (B, heads, nc, vol, vol + K)
        mask = F.pad(mask, (0, K), value=True)
        v_lg = torch.cat([v, gv.expand(B, heads, nc, K, hd)], dim=3)
        # global queries over every slot (:928-945)
        s2 = (gq.squeeze(2) * hd ** -0.5) @ k_gl.reshape(B, heads, nc * vol, hd).transpose(-1, -2)   # (B, heads, K, nc * vol)
        m2 = real.reshape(-1) > 0 if padding_type == "ignore" else None            # raster order (see the docstring)
        v2 = v_gl.reshape(B, heads, nc * vol, hd)
        if global_self_attn:
            s2 = torch.cat([s2, (gq_g.squeeze(2) * hd ** -0.5) @ gk_g.squeeze(2).transpose(-1, -2)], dim=-1)
            m2 = F.pad(m2, (0, K), value=True) if m2 is not None else None
            v2 = torch.cat([v2, gv_g.squeeze(2)], dim=2)
        if m2 is not None:
            p2 = torch.softmax(s2.masked_fill(~m2, -1e18), dim=-1) * m2
        else:
            p2 = torch.softmax(s2, dim=-1)
        new_g = (p2 @ v2).permute(0, 2, 1, 3).reshape(B, K, C)
    else:
        v_lg = v
    s = s.masked_fill(~mask, -1e18)
    p = torch.softmax(s, dim=-1) * mask
    o = (p @ v_lg).permute(0, 2, 3, 1, 4).reshape(B, nc, vol, C)
    o = _from_cuboids(o, size, strategy, padded)
    if any(s_ > 0 for s_ in shift):
        o = torch.roll(o, shifts=list(shift), dims=(1, 2, 3))
    if padding_type == "nearest" and any(pad):   # _generalize_unpadding (:261-270): resample back
        o = F.interpolate(o.permute(0, 4, 1, 2, 3), size=(T, H, W)).permute(0, 2, 3, 4, 1).contiguous()
    else:
        o = o[:, :T, :H, :W].contiguous()
    return o if new_g is None else (o, new_g)


def cuboid_attention_gv(sd, p, x, g, heads, size0, strategy, shift0, padding_type="zeros", global_self_attn=False,
                        separate=False):
    """CuboidSelfAttentionLayer.forward with use_global_vector (cuboid_transformer.py:812-966), global_dim_ratio = 1:
    x (B, T, H, W, C), g (B, K, C) -> (x_out, g_out), both BEFORE the residual adds of the stack block.
    separate: separate_global_qkv=True - the six extra Linear nets of :770-795 (no bias, like qkv)."""
    C = x.shape[-1]
    y = F.layer_norm(x, (C,), sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], 1e-5)
    gn = F.layer_norm(g, (C,), sd[f"{p}.global_vec_norm.weight"], sd[f"{p}.global_vec_norm.bias"], 1e-5)
    if separate:
        tok = torch.cat([F.linear(y, sd[f"{p}.l2g_q_net.weight"]), F.linear(y, sd[f"{p}.g2l_k_net.weight"]),
                         F.linear(y, sd[f"{p}.g2l_v_net.weight"])], dim=-1)
        grow = [F.linear(gn, sd[f"{p}.l2g_global_kv_net.weight"]), F.linear(gn, sd[f"{p}.g2l_global_q_net.weight"])]
        if global_self_attn:
            grow.append(F.linear(gn, sd[f"{p}.g2g_global_qkv_net.weight"]))
        kw = dict(sep=(tok, torch.cat(grow, dim=-1)))
    else:
        kw = dict(gqkv=F.linear(gn, sd[f"{p}.global_qkv.weight"]))
    o, ng = cuboid_attention_core(F.linear(y, sd[f"{p}.qkv.weight"]), sd[f"{p}.relative_position_bias_table"], heads,
                                  size0, strategy, shift0, padding_type, global_self_attn=global_self_attn, **kw)
    return (F.linear(o, sd[f"{p}.proj.weight"], sd[f"{p}.proj.bias"]),
            F.linear(ng, sd[f"{p}.global_proj.weight"], sd[f"{p}.global_proj.bias"]))


def cuboid_attention(sd, p, x, heads, size0, strategy, shift0, padding_type="zeros"):
    """CuboidSelfAttentionLayer.forward (cuboid_transformer.py:812-966) for any cuboid size / strategy / shift with
    'zeros', 'ignore' or 'nearest' padding: LN -> qkv (no bias) -> cuboid_attention_core -> proj. x: (B, T, H, W, C).
    ('nearest' resamples AFTER the LayerNorm; qkv is per token without bias, so resampling its output is the same thing.)"""
    C = x.shape[-1]
    y = F.layer_norm(x, (C,), sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], 1e-5)
    o = cuboid_attention_core(F.linear(y, sd[f"{p}.qkv.weight"]), sd[f"{p}.relative_position_bias_table"], heads,
                              size0, strategy, shift0, padding_type)
    return F.linear(o, sd[f"{p}.proj.weight"], sd[f"{p}.proj.bias"])


def stack_block(sd, p, x, heads, layers=None, padding_type="zeros", g=None, global_ffn=True, global_self_attn=False,
                separate=False):
    """StackCuboidSelfAttentionBlock.forward with use_inter_ffn (cuboid_transformer.py:1147-1156). `layers`: list of
    (cuboid_size, strategy, shift_size) (None = the axial pattern of the shipped config). g (B, K, C): global vectors
    (:1130-1145) -> returns (x, g)."""
    if g is not None:
        for i, (size, strategy, shift) in enumerate(layers):
            xo, go = cuboid_attention_gv(sd, f"{p}.attn_l.{i}", x, g, heads, size, strategy, shift, padding_type,
                                         global_self_attn, separate)
            x, g = x + xo, g + go
            x = ffn(sd, f"{p}.ffn_l.{i}", x)
            if global_ffn:
                g = ffn(sd, f"{p}.global_ffn_l.{i}", g)
        return x, g
    if layers is None:
        for i in range(3):
            x = x + axial_attention(sd, f"{p}.attn_l.{i}", x, heads, i)
            x = ffn(sd, f"{p}.ffn_l.{i}", x)
        return x
    for i, (size, strategy, shift) in enumerate(layers):
        x = x + cuboid_attention(sd, f"{p}.attn_l.{i}", x, heads, size, strategy, shift, padding_type)
        x = ffn(sd, f"{p}.ffn_l.{i}", x)
    return x


def patch_merge(sd, p, x):
    """PatchMerging3D.forward, downsample (1,2,2) (cuboid_transformer.py:261-296)."""
    B, T, H, W, C = x.shape
    x = x.reshape(B, T, H // 2, 2, W // 2, 2, C).permute(0, 1, 2, 4, 3, 5, 6).reshape(B, T, H // 2, W // 2, 4 * C)
    x = F.layer_norm(x, (4 * C,), sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], 1e-5)
    return F.linear(x, sd[f"{p}.reduction.weight"])


def upsample3d(sd, p, x):
    """Upsample3DLayer.forward, no temporal upsample (cuboid_transformer.py:353-375)."""
    B, T, H, W, C = x.shape
    y = x.reshape(B * T, H, W, C).permute(0, 3, 1, 2)
    y = F.interpolate(y, scale_factor=2, mode="nearest")
    y = F.conv2d(y, sd[f"{p}.conv.weight"], sd[f"{p}.conv.bias"], padding=1)
    return y.permute(0, 2, 3, 1).reshape(B, T, 2 * H, 2 * W, -1)


def unet_forward(sd, cfg, x, t, cond):
    """CuboidTransformerUNet.forward (src/prediff/models/cuboid_transformer/cuboid_transformer_unet.py:406-493).
    x (B,T_out,H,W,C), t (B,) int64, cond (B,T_in,H,W,C) -> (B,T_out,H,W,C)."""
    heads = cfg.num_heads
    pats = tuple(getattr(cfg, "patterns", ("axial", "axial")))
    K = getattr(cfg, "num_global_vectors", 0)
    layers = [None if pats[lvl] == "axial" and not K else cfg.layers(lvl) for lvl in range(2)]
    x = torch.cat([cond, x], dim=1)
    ind = torch.ones_like(x[..., :1])
    ind[:, cfg.t_in:] = 0.0
    x = torch.cat([x, ind], dim=-1).permute(0, 4, 1, 2, 3)
    cin = cfg.c + 1
    x = _res_block3d(sd, "first_proj", x, None, _gn_groups(cin), _gn_groups(cfg.units[0]))
    x = x.permute(0, 2, 3, 4, 1)
    B, T, H, W, C = x.shape
    # PosEmbed 't+h+w' (cuboid_transformer.py:78-85)
    x = x + sd["pos_embed.T_embed.weight"].reshape(T, 1, 1, C) + sd["pos_embed.H_embed.weight"].reshape(1, H, 1, C) \
        + sd["pos_embed.W_embed.weight"].reshape(1, 1, W, C)
    # TimeEmbedLayer (time_embed.py:16-24)
    e = timestep_embedding(t, cfg.units[0])
    e = F.linear(e, sd["time_embed.layer.0.weight"], sd["time_embed.layer.0.bias"])
    t_emb = F.linear(F.silu(e), sd["time_embed.layer.2.weight"], sd["time_embed.layer.2.bias"])

    # global vectors (cuboid_transformer_unet.py:432-434, 449-450, 489-490): carried beside x, projected between levels
    gvec = [sd["init_global_vectors"].expand(B, K, C) if K else None]

    def level(name_t, name_s, lvl, x):
        g = _gn_groups(cfg.units[lvl])
        for d in range(cfg.depth[lvl]):
            x = _res_block3d(sd, f"{name_t}.{lvl}", x.permute(0, 4, 1, 2, 3), t_emb, g, g).permute(0, 2, 3, 4, 1)
            r = stack_block(sd, f"{name_s}.{lvl}.{d}", x, heads, layers[lvl], getattr(cfg, "padding_type", "zeros"),
                            gvec[0], getattr(cfg, "use_global_vector_ffn", True), getattr(cfg, "use_global_self_attn", False),
                            getattr(cfg, "separate_global_qkv", False))
            x, gvec[0] = r if K else (r, None)
        return x

    x = level("down_time_embed_blocks", "down_self_blocks", 0, x)
    skip = x
    x = patch_merge(sd, "downsample_layers.0", x)
    if K:
        gvec[0] = F.linear(gvec[0], sd["down_layer_global_proj.0.weight"], sd["down_layer_global_proj.0.bias"])
    x = level("down_time_embed_blocks", "down_self_blocks", 1, x)
    x = level("up_time_embed_blocks", "up_self_blocks", 1, x)
    x = upsample3d(sd, "upsample_layers.0", x)
    if K:
        gvec[0] = F.linear(gvec[0], sd["up_layer_global_proj.0.weight"], sd["up_layer_global_proj.0.bias"])
    x = x + skip
    x = level("up_time_embed_blocks", "up_self_blocks", 0, x)
    return F.linear(x[:, cfg.t_in:], sd["final_proj.weight"], sd["final_proj.bias"])


# ----------------------------------------------------------------------------------------------- VAE


def _resnet2d(sd, p, x, groups):
    """ResnetBlock2D.forward without temb (src/prediff/taming/resnet.py:454-495), eps 1e-6."""
    h = F.silu(F.group_norm(x, groups, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], 1e-6))
    h = F.conv2d(h, sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], padding=1)
    h = F.silu(F.group_norm(h, groups, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], 1e-6))
    h = F.conv2d(h, sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"], padding=1)
    if f"{p}.conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[f"{p}.conv_shortcut.weight"], sd[f"{p}.conv_shortcut.bias"])
    return x + h


def _attn_block(sd, p, x, groups):
    """AttentionBlock.forward, single head (src/prediff/taming/attention.py:136-189)."""
    N, C, H, W = x.shape
    h = F.group_norm(x, groups, sd[f"{p}.group_norm.weight"], sd[f"{p}.group_norm.bias"], 1e-6)
    h = h.view(N, C, H * W).transpose(1, 2)
    q = F.linear(h, sd[f"{p}.query.weight"], sd[f"{p}.query.bias"])
    k = F.linear(h, sd[f"{p}.key.weight"], sd[f"{p}.key.bias"])
    v = F.linear(h, sd[f"{p}.value.weight"], sd[f"{p}.value.bias"])
    s = torch.softmax((q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(C)), dim=-1)
    o = F.linear(s @ v, sd[f"{p}.proj_attn.weight"], sd[f"{p}.proj_attn.bias"])
    return o.transpose(-1, -2).reshape(N, C, H, W) + x


def _mid_block(sd, p, x, groups):
    """UNetMidBlock2D.forward (src/prediff/taming/unet_2d_blocks.py:158-165)."""
    x = _resnet2d(sd, f"{p}.resnets.0", x, groups)
    x = _attn_block(sd, f"{p}.attentions.0", x, groups)
    return _resnet2d(sd, f"{p}.resnets.1", x, groups)


def vae_encode_moments(sd, cfg, x):
    """AutoencoderKL.encode up to the moments tensor (autoencoder_kl.py:80-84; Encoder.forward vae.py:70-86;
    DownEncoderBlock2D unet_2d_blocks.py:217-225; Downsample2D resnet.py:181-190). x (N,1,H,W) -> (N,2*latent,h,w)."""
    g = cfg.norm_num_groups
    boc = cfg.block_out_channels
    h = F.conv2d(x, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    for i in range(len(boc)):
        for j in range(cfg.layers_per_block):
            h = _resnet2d(sd, f"encoder.down_blocks.{i}.resnets.{j}", h, g)
        if i != len(boc) - 1:
            p = f"encoder.down_blocks.{i}.downsamplers.0.conv"
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[f"{p}.weight"], sd[f"{p}.bias"], stride=2)
    h = _mid_block(sd, "encoder.mid_block", h, g)
    h = F.silu(F.group_norm(h, g, sd["encoder.conv_norm_out.weight"], sd["encoder.conv_norm_out.bias"], 1e-6))
    h = F.conv2d(h, sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], padding=1)
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])


def vae_encode_mode(sd, cfg, x):
    """DiagonalGaussianDistribution.mode() = mean = first half of the channels
    (src/prediff/utils/distributions.py:26-33,70-71)."""
    return vae_encode_moments(sd, cfg, x)[:, :cfg.latent_channels]


def vae_decode(sd, cfg, z):
    """AutoencoderKL.decode (autoencoder_kl.py:86-113; Decoder.forward vae.py:150-166; UpDecoderBlock2D
    unet_2d_blocks.py:271-279; Upsample2D resnet.py:108-143). z (N,latent,h,w) -> (N,1,H,W)."""
    g = cfg.norm_num_groups
    boc = cfg.block_out_channels
    h = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    h = F.conv2d(h, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    h = _mid_block(sd, "decoder.mid_block", h, g)
    for i in range(len(boc)):
        for j in range(cfg.layers_per_block + 1):
            h = _resnet2d(sd, f"decoder.up_blocks.{i}.resnets.{j}", h, g)
        if i != len(boc) - 1:
            p = f"decoder.up_blocks.{i}.upsamplers.0.conv"
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"{p}.weight"], sd[f"{p}.bias"], padding=1)
    h = F.silu(F.group_norm(h, g, sd["decoder.conv_norm_out.weight"], sd["decoder.conv_norm_out.bias"], 1e-6))
    return F.conv2d(h, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)


# ----------------------------------------------------------------------------------------------- sampler


def p_sample_ddpm(sched, eps, z, t, noise, guide=None):
    """One ancestral step given the model output eps: predict_start_from_noise + q_posterior + p_sample
    (latent_diffusion.py:553-566,598-631; clip_denoised False, temperature 1); optional knowledge-alignment shift
    `guide` (aligned_mean, latent_diffusion.py:592-596). t: python int shared by the batch."""
    z0 = sched["sqrt_recip_alphas_cumprod"][t] * z - sched["sqrt_recipm1_alphas_cumprod"][t] * eps
    mean = sched["posterior_mean_coef1"][t] * z0 + sched["posterior_mean_coef2"][t] * z
    logvar = sched["posterior_log_variance_clipped"][t]
    if guide is not None:
        mean = mean - (0.5 * logvar).exp() * guide
    nonzero = 0.0 if t == 0 else 1.0
    return mean + nonzero * (0.5 * logvar).exp() * noise


def cond_schedule_ids(num_timesteps_cond, timesteps=1000):
    """make_cond_schedule (latent_diffusion.py:295-299)."""
    ids = torch.full((timesteps,), timesteps - 1, dtype=torch.long)
    ids[:num_timesteps_cond] = torch.round(torch.linspace(0, timesteps - 1, num_timesteps_cond)).long()
    return ids


def sample_loop_ddpm(sd, cfg, sched, z, cond, noise, n_steps, cond_ids=None, cond_noise=None):
    """p_sample_loop (latent_diffusion.py:633-684) with `timesteps=n_steps`, x_T = z, and the per-step
    torch.randn replaced by the pre-generated noise[k] (k-th executed step). cond_ids / cond_noise: the
    shorten_cond_schedule branch (:665-667) - the context is re-noised, cumulatively, before every step."""
    for k, i in enumerate(reversed(range(n_steps))):
        t = torch.full((z.shape[0],), i, dtype=torch.long)
        if cond_ids is not None:
            cond = q_sample(sched, cond, cond_ids[t], cond_noise[k])
        eps = unet_forward(sd, cfg, z, t, cond)
        z = p_sample_ddpm(sched, eps, z, i, noise[k])
    return z


def ddim_coefficients(sched, n_steps, eta=0.0):
    """Per-step (t, a_t, a_prev, sigma) for the S6 DDIM definition built from the
    reference helpers diffusion/utils.py:42-70, iterated from the largest timestep down."""
    ac = sched["alphas_cumprod"].numpy()
    ts = ddim_timesteps(n_steps, ac.shape[0])
    sig, a, a_prev = ddim_params(ac, ts, eta)
    return [(int(ts[i]), float(a[i]), float(a_prev[i]), float(sig[i])) for i in reversed(range(len(ts)))]


def sample_loop_ddim(sd, cfg, sched, z, cond, n_steps, eta=0.0, noise=None, guide_fn=None):
    """50-step DDIM (eta=0 deterministic): z0 = (z - sqrt(1-a_t) eps)/sqrt(a_t);
    z <- sqrt(a_prev) z0 + sqrt(1 - a_prev - sigma^2) eps + sigma noise.
    guide_fn(z, t) -> g (knowledge alignment): eps_hat = eps + sqrt(1 - a_t) * g, i.e.
    guided-diffusion's condition_score with the sign of the reference's aligned_mean (latent_diffusion.py:594-595)."""
    for k, (t, a_t, a_prev, sigma) in enumerate(ddim_coefficients(sched, n_steps, eta)):
        tt = torch.full((z.shape[0],), t, dtype=torch.long)
        eps = unet_forward(sd, cfg, z, tt, cond)
        if guide_fn is not None:
            eps = eps + math.sqrt(1.0 - a_t) * guide_fn(z, tt)
        z0 = (z - math.sqrt(1.0 - a_t) * eps) / math.sqrt(a_t)
        z = math.sqrt(a_prev) * z0 + math.sqrt(max(1.0 - a_prev - sigma ** 2, 0.0)) * eps
        if noise is not None and sigma > 0:
            z = z + sigma * noise[k]
    return z


# ----------------------------------------------------------------------------------------------- forward loss


def lvlb_weights(sched, timesteps=1000, linear_start=1e-4, linear_end=2e-2):
    """eps-parameterization branch of register_schedule (latent_diffusion.py:270-277): fp32 tensor arithmetic on the
    registered buffers, alphas = fp32(1 - betas) from the float64 schedule, entry 0 replaced by entry 1."""
    betas64 = np.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=np.float64) ** 2
    alphas = torch.tensor(1.0 - betas64, dtype=torch.float32)
    w = sched["betas"] ** 2 / (2 * sched["posterior_variance"] * alphas * (1 - sched["alphas_cumprod"]))
    w[0] = w[1]
    return w


def q_sample(sched, x_start, t, noise):
    """latent_diffusion.py:489-492."""
    shape = [t.shape[0]] + [1] * (x_start.dim() - 1)
    return sched["sqrt_alphas_cumprod"][t].reshape(shape) * x_start \
        + sched["sqrt_one_minus_alphas_cumprod"][t].reshape(shape) * noise


def p_losses(sd, cfg, sched, x_start, cond, t, noise, loss_type="l2", original_elbo_weight=0.0, l_simple_weight=1.0,
             logvar_init=0.0):
    """LatentDiffusion.p_losses (latent_diffusion.py:517-551), eps-parameterization, learn_logvar = False.
    Returns dict(loss_simple, loss_vlb, loss, per_sample)."""
    eps = unet_forward(sd, cfg, q_sample(sched, x_start, t, noise), t, cond)
    d = (noise - eps).abs() if loss_type == "l1" else (noise - eps) ** 2
    per_sample = d.mean(dim=tuple(range(1, d.dim())))
    logvar_t = torch.full((t.shape[0],), float(logvar_init))
    loss = l_simple_weight * (per_sample / torch.exp(logvar_t) + logvar_t).mean()
    loss_vlb = (lvlb_weights(sched)[t] * per_sample).mean()
    return dict(loss_simple=per_sample.mean(), loss_vlb=loss_vlb, loss=loss + original_elbo_weight * loss_vlb,
                per_sample=per_sample)


# ----------------------------------------------------------------------------------------------- knowledge alignment


def ka_forward(sd, cfg, x, t):
    """NoisyCuboidTransformerEncoder.forward (src/prediff/diffusion/knowledge_alignment/models.py:459-528) with
    pool="attention", readout_seq=True, out_len=T: x (B,T,H,W,C), t (B,) -> (B,T,1)."""
    heads = cfg.num_heads
    u0, u1 = cfg.units
    B, T = x.shape[0], x.shape[1]
    h = _res_block3d(sd, "first_proj", x.permute(0, 4, 1, 2, 3), None, _gn_groups(cfg.c), _gn_groups(u0)).permute(0, 2, 3, 4, 1)
    _, _, H, W, C = h.shape
    h = h + sd["pos_embed.T_embed.weight"].reshape(T, 1, 1, C) + sd["pos_embed.H_embed.weight"].reshape(1, H, 1, C) \
        + sd["pos_embed.W_embed.weight"].reshape(1, 1, W, C)
    e = timestep_embedding(t, u0)
    e = F.linear(e, sd["time_embed.layer.0.weight"], sd["time_embed.layer.0.bias"])
    t_emb = F.linear(F.silu(e), sd["time_embed.layer.2.weight"], sd["time_embed.layer.2.bias"])
    for lvl in range(2):
        if lvl > 0:
            h = patch_merge(sd, "downsample_layers.0", h)
        g = _gn_groups(cfg.units[lvl])
        for d in range(cfg.depth[lvl]):
            h = _res_block3d(sd, f"down_time_embed_blocks.{lvl}", h.permute(0, 4, 1, 2, 3), t_emb, g, g).permute(0, 2, 3, 4, 1)
            h = stack_block(sd, f"down_self_blocks.{lvl}.{d}", h, heads)
    # read-out per frame: GroupNorm(32) + SiLU + AttentionPool3d (models.py:49-104), token 0 of the pooled sequence
    o = h.permute(0, 1, 4, 2, 3).reshape(B * T, u1, -1)                       # (b t) c (h w)
    o = F.silu(F.group_norm(o, min(u1, 32), sd["out.0.weight"], sd["out.0.bias"], 1e-5))
    o = torch.cat([o.mean(dim=-1, keepdim=True), o], dim=-1) + sd["out.2.positional_embedding"][None]
    qkv = F.conv1d(o, sd["out.2.qkv_proj.weight"], sd["out.2.qkv_proj.bias"])  # (bt, 3c, L)
    bs, width, L = qkv.shape
    ch = width // (3 * heads)
    q, k, v = qkv.chunk(3, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    wgt = torch.einsum("bct,bcs->bts", (q * scale).reshape(bs * heads, ch, L), (k * scale).reshape(bs * heads, ch, L))
    wgt = torch.softmax(wgt, dim=-1)
    a = torch.einsum("bts,bcs->bct", wgt, v.reshape(bs * heads, ch, L)).reshape(bs, -1, L)
    out = F.conv1d(a, sd["out.2.c_proj.weight"], sd["out.2.c_proj.bias"])[:, :, 0]
    return out.reshape(B, T, -1)


def ka_alignment_value(sd, cfg, zt, t, avg_x_gt):
    """SEVIRAvgIntensityAlignment.alignment_fn (knowledge_alignment/sevir.py:55-83): || mean_T U(zt,t) - target ||_2
    with the norm taken over the whole (B,1) tensor."""
    pred = ka_forward(sd, cfg, zt, t).mean(dim=1)
    return torch.linalg.vector_norm(pred - avg_x_gt, ord=2)


def ka_mean_shift(sd, cfg, zt, t, avg_x_gt, guide_scale=50.0):
    """get_mean_shift (sevir.py:85-104) = guide_scale * d alignment_fn / d zt (alignment_pl.py:423-446)."""
    with torch.enable_grad():
        z = zt.detach().clone().requires_grad_(True)
        val = ka_alignment_value(sd, cfg, z, t, avg_x_gt)
        grad = torch.autograd.grad(val.sum(), z)[0]
    return guide_scale * grad


def to_torch_sd(np_sd):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in np_sd.items()}
