"""Configs, state_dict key/shape specs (the reference's key names) and a portable seeded weight generator.

The reference ships no weights offline and its default init zeroes most projections, so parity
tests and the benchmark use weights drawn by `seeded_state_dict`: numpy PCG64 streams, identical on every
machine, every tensor non-degenerate (norm weights 1 + 0.1 N, biases 0.02 N, matrices N(0, gain^2/fan_in)).
The same dict is loaded into the reference modules (golden generation), the oracle and the CUDA path.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np


@dataclass
class UNetConfig:
    """Subset of cfg.yaml `model.latent_model` the CUDA path is built for (axial pattern, 2 levels)."""
    t_in: int = 7
    t_out: int = 6
    h: int = 16
    w: int = 16
    c: int = 64
    base_units: int = 256
    depth: Tuple[int, int] = (4, 4)
    num_heads: int = 4
    # block_attn_patterns per level (names of cuboid_transformer_patterns.py) and padding_type ('zeros' | 'ignore' | 'nearest')
    patterns: Tuple[str, str] = ("axial", "axial")
    padding_type: str = "zeros"
    # block_attn_patterns=None in the reference: explicit per-level lists of (cuboid_size, strategy, shift_size)
    # (block_cuboid_size / block_cuboid_strategy / block_cuboid_shift_size, cuboid_transformer_unet.py:215-232)
    explicit_layers: Tuple = None
    # global vectors (cuboid_transformer_unet.py:55-60, cuboid_transformer.py:864-945): K learned vectors per sample that
    # every cuboid attends to and that attend to the whole grid; shared q|k|v net (separate_global_qkv=False), ratio 1
    num_global_vectors: int = 0
    use_global_vector_ffn: bool = True
    use_global_self_attn: bool = False
    separate_global_qkv: bool = False   # True: the six extra Linear nets of cuboid_transformer.py:770-795 instead of global_qkv

    @property
    def T(self):
        return self.t_in + self.t_out

    @property
    def units(self):
        return (self.base_units, 2 * self.base_units)

    @property
    def temb_channels(self):
        return 4 * self.base_units

    def layers(self, level):
        """(cuboid_size, strategy, shift_size) of every attention layer of a level's StackCuboidSelfAttentionBlock:
        the level's pattern evaluated on its mem_shape (cuboid_transformer_unet.py:201-214, 387-404)."""
        from . import patterns as P
        if self.explicit_layers is not None:
            return [(tuple(a), tuple(b), tuple(c)) for a, b, c in self.explicit_layers[level]]
        return P.resolve(self.patterns[level], (self.T, self.h >> level, self.w >> level, self.units[level]))

    def cuboids(self, level):
        """Constructor cuboid sizes of a level's attention layers (axial: (T,1,1), (1,H,1), (1,1,W))."""
        return [size for size, _, _ in self.layers(level)]


@dataclass
class VAEConfig:
    """cfg.yaml `model.vae` (diffusers-style AutoencoderKL vendored in src/prediff/taming)."""
    in_channels: int = 1
    out_channels: int = 1
    latent_channels: int = 64
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    h: int = 128
    w: int = 128


@dataclass
class KAConfig:
    """cfg.yaml `model.align.model_args`: NoisyCuboidTransformerEncoder (knowledge-alignment network U(z_t, t))."""
    t: int = 6
    h: int = 16
    w: int = 16
    c: int = 64
    base_units: int = 128
    depth: Tuple[int, int] = (1, 1)
    num_heads: int = 4
    guide_scale: float = 50.0

    @property
    def units(self):
        return (self.base_units, 2 * self.base_units)

    @property
    def temb_channels(self):
        return 4 * self.base_units

    def cuboids(self, level):
        h, w = self.h >> level, self.w >> level
        return [(self.t, 1, 1), (1, h, 1), (1, 1, w)]


TINY_UNET = UNetConfig(base_units=64, depth=(1, 1))
TINY_VAE = VAEConfig(latent_channels=64, block_out_channels=(64, 64, 128, 128), layers_per_block=1, h=128, w=128)

Spec = List[Tuple[str, Tuple[int, ...]]]


def _resblock3d(prefix, cin, cout, temb) -> Spec:
    s = [(f"{prefix}.in_layers.0.weight", (cin,)), (f"{prefix}.in_layers.0.bias", (cin,)),
         (f"{prefix}.in_layers.2.weight", (cout, cin, 3, 3, 3)), (f"{prefix}.in_layers.2.bias", (cout,))]
    if temb:
        s += [(f"{prefix}.emb_layers.1.weight", (cout, temb)), (f"{prefix}.emb_layers.1.bias", (cout,))]
    s += [(f"{prefix}.out_layers.0.weight", (cout,)), (f"{prefix}.out_layers.0.bias", (cout,)),
          (f"{prefix}.out_layers.3.weight", (cout, cout, 3, 3, 3)), (f"{prefix}.out_layers.3.bias", (cout,))]
    if cin != cout:
        s += [(f"{prefix}.skip_connection.weight", (cout, cin, 1, 1, 1)), (f"{prefix}.skip_connection.bias", (cout,))]
    return s


def _stack_block(prefix, dim, heads, cuboids, gv=False, gv_ffn=False, gv_sep=False, gv_sa=False) -> Spec:
    """gv: the block carries global vectors (global_qkv / global_proj / global_vec_norm per attention layer,
    cuboid_transformer.py:777-810); gv_ffn: and a PositionwiseFFN for them per layer (global_ffn_l, :1054-1068)."""
    s: Spec = []
    for name in ("ffn_l",) + (("global_ffn_l",) if gv and gv_ffn else ()):
        for i in range(len(cuboids)):
            p = f"{prefix}.{name}.{i}"
            s += [(f"{p}.ffn_1.weight", (4 * dim, dim)), (f"{p}.ffn_1.bias", (4 * dim,)),
                  (f"{p}.ffn_2.weight", (dim, 4 * dim)), (f"{p}.ffn_2.bias", (dim,)),
                  (f"{p}.layer_norm.weight", (dim,)), (f"{p}.layer_norm.bias", (dim,))]
    for i, (bt, bh, bw) in enumerate(cuboids):
        p = f"{prefix}.attn_l.{i}"
        s += [(f"{p}.relative_position_bias_table", ((2 * bt - 1) * (2 * bh - 1) * (2 * bw - 1), heads)),
              (f"{p}.qkv.weight", (3 * dim, dim))]
        if gv and gv_sep:   # separate_global_qkv (registration order of :770-795)
            s += [(f"{p}.l2g_q_net.weight", (dim, dim)), (f"{p}.l2g_global_kv_net.weight", (2 * dim, dim)),
                  (f"{p}.g2l_global_q_net.weight", (dim, dim)), (f"{p}.g2l_k_net.weight", (dim, dim)),
                  (f"{p}.g2l_v_net.weight", (dim, dim))]
            if gv_sa:
                s += [(f"{p}.g2g_global_qkv_net.weight", (3 * dim, dim))]
        elif gv:
            s += [(f"{p}.global_qkv.weight", (3 * dim, dim))]
        s += [(f"{p}.proj.weight", (dim, dim)), (f"{p}.proj.bias", (dim,))]
        if gv:
            s += [(f"{p}.global_proj.weight", (dim, dim)), (f"{p}.global_proj.bias", (dim,))]
        s += [(f"{p}.norm.weight", (dim,)), (f"{p}.norm.bias", (dim,))]
        if gv:
            s += [(f"{p}.global_vec_norm.weight", (dim,)), (f"{p}.global_vec_norm.bias", (dim,))]
    return s


def unet_param_spec(cfg: UNetConfig) -> Spec:
    """Parameter names/shapes of the reference CuboidTransformerUNet.state_dict() (minus the int64
    `relative_position_index` buffers, which are derived), in the reference's registration order."""
    u0, u1 = cfg.units
    te = cfg.temb_channels
    s: Spec = []
    gv = cfg.num_global_vectors > 0
    if gv:
        s += [("init_global_vectors", (cfg.num_global_vectors, u0))]
    s += _resblock3d("first_proj", cfg.c + 1, u0, 0)
    s += [("pos_embed.T_embed.weight", (cfg.T, u0)), ("pos_embed.H_embed.weight", (cfg.h, u0)),
          ("pos_embed.W_embed.weight", (cfg.w, u0))]
    s += [("time_embed.layer.0.weight", (te, u0)), ("time_embed.layer.0.bias", (te,)),
          ("time_embed.layer.2.weight", (te, te)), ("time_embed.layer.2.bias", (te,))]
    s += [("downsample_layers.0.reduction.weight", (u1, 4 * u0)), ("downsample_layers.0.norm.weight", (4 * u0,)),
          ("downsample_layers.0.norm.bias", (4 * u0,))]
    if gv:
        s += [("down_layer_global_proj.0.weight", (u1, u0)), ("down_layer_global_proj.0.bias", (u1,))]
    s += [("upsample_layers.0.conv.weight", (u0, u1, 3, 3)), ("upsample_layers.0.conv.bias", (u0,))]
    if gv:
        s += [("up_layer_global_proj.0.weight", (u0, u1)), ("up_layer_global_proj.0.bias", (u0,))]
    for name in ("down_self_blocks", "up_self_blocks"):
        for lvl, dim in enumerate((u0, u1)):
            for d in range(cfg.depth[lvl]):
                s += _stack_block(f"{name}.{lvl}.{d}", dim, cfg.num_heads, cfg.cuboids(lvl), gv, cfg.use_global_vector_ffn,
                                  cfg.separate_global_qkv, cfg.use_global_self_attn)
    for name in ("down_time_embed_blocks", "up_time_embed_blocks"):
        for lvl, dim in enumerate((u0, u1)):
            s += _resblock3d(f"{name}.{lvl}", dim, dim, te)
    s += [("final_proj.weight", (cfg.c, u0)), ("final_proj.bias", (cfg.c,))]
    return s


def _resnet2d(prefix, cin, cout) -> Spec:
    s = [(f"{prefix}.norm1.weight", (cin,)), (f"{prefix}.norm1.bias", (cin,)),
         (f"{prefix}.conv1.weight", (cout, cin, 3, 3)), (f"{prefix}.conv1.bias", (cout,)),
         (f"{prefix}.norm2.weight", (cout,)), (f"{prefix}.norm2.bias", (cout,)),
         (f"{prefix}.conv2.weight", (cout, cout, 3, 3)), (f"{prefix}.conv2.bias", (cout,))]
    if cin != cout:
        s += [(f"{prefix}.conv_shortcut.weight", (cout, cin, 1, 1)), (f"{prefix}.conv_shortcut.bias", (cout,))]
    return s


def _mid_block(prefix, c) -> Spec:
    a = f"{prefix}.attentions.0"
    s = [(f"{a}.group_norm.weight", (c,)), (f"{a}.group_norm.bias", (c,))]
    for n in ("query", "key", "value", "proj_attn"):
        s += [(f"{a}.{n}.weight", (c, c)), (f"{a}.{n}.bias", (c,))]
    s += _resnet2d(f"{prefix}.resnets.0", c, c)
    s += _resnet2d(f"{prefix}.resnets.1", c, c)
    return s


def vae_param_spec(cfg: VAEConfig) -> Spec:
    """Parameter names/shapes of the reference AutoencoderKL.state_dict() (taming/autoencoder_kl.py)."""
    boc = cfg.block_out_channels
    s: Spec = [("encoder.conv_in.weight", (boc[0], cfg.in_channels, 3, 3)), ("encoder.conv_in.bias", (boc[0],))]
    cin = boc[0]
    for i, cout in enumerate(boc):
        for j in range(cfg.layers_per_block):
            s += _resnet2d(f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
        if i != len(boc) - 1:
            s += [(f"encoder.down_blocks.{i}.downsamplers.0.conv.weight", (cout, cout, 3, 3)),
                  (f"encoder.down_blocks.{i}.downsamplers.0.conv.bias", (cout,))]
        cin = cout
    s += _mid_block("encoder.mid_block", boc[-1])
    s += [("encoder.conv_norm_out.weight", (boc[-1],)), ("encoder.conv_norm_out.bias", (boc[-1],)),
          ("encoder.conv_out.weight", (2 * cfg.latent_channels, boc[-1], 3, 3)),
          ("encoder.conv_out.bias", (2 * cfg.latent_channels,))]
    s += [("decoder.conv_in.weight", (boc[-1], cfg.latent_channels, 3, 3)), ("decoder.conv_in.bias", (boc[-1],))]
    rev = list(reversed(boc))
    cin = rev[0]
    for i, cout in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            s += _resnet2d(f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
        if i != len(boc) - 1:
            s += [(f"decoder.up_blocks.{i}.upsamplers.0.conv.weight", (cout, cout, 3, 3)),
                  (f"decoder.up_blocks.{i}.upsamplers.0.conv.bias", (cout,))]
        cin = cout
    s += _mid_block("decoder.mid_block", boc[-1])
    s += [("decoder.conv_norm_out.weight", (boc[0],)), ("decoder.conv_norm_out.bias", (boc[0],)),
          ("decoder.conv_out.weight", (cfg.out_channels, boc[0], 3, 3)), ("decoder.conv_out.bias", (cfg.out_channels,))]
    s += [("quant_conv.weight", (2 * cfg.latent_channels, 2 * cfg.latent_channels, 1, 1)),
          ("quant_conv.bias", (2 * cfg.latent_channels,)),
          ("post_quant_conv.weight", (cfg.latent_channels, cfg.latent_channels, 1, 1)),
          ("post_quant_conv.bias", (cfg.latent_channels,))]
    return s


def ka_param_spec(cfg: KAConfig) -> Spec:
    """Parameter names/shapes of the reference NoisyCuboidTransformerEncoder.state_dict()
    (src/prediff/diffusion/knowledge_alignment/models.py), minus the derived int64 buffers."""
    u0, u1 = cfg.units
    te = cfg.temb_channels
    s: Spec = []
    s += _resblock3d("first_proj", cfg.c, u0, 0)
    s += [("pos_embed.T_embed.weight", (cfg.t, u0)), ("pos_embed.H_embed.weight", (cfg.h, u0)),
          ("pos_embed.W_embed.weight", (cfg.w, u0))]
    s += [("time_embed.layer.0.weight", (te, u0)), ("time_embed.layer.0.bias", (te,)),
          ("time_embed.layer.2.weight", (te, te)), ("time_embed.layer.2.bias", (te,))]
    s += [("downsample_layers.0.reduction.weight", (u1, 4 * u0)), ("downsample_layers.0.norm.weight", (4 * u0,)),
          ("downsample_layers.0.norm.bias", (4 * u0,))]
    for lvl, dim in enumerate((u0, u1)):
        for d in range(cfg.depth[lvl]):
            s += _stack_block(f"down_self_blocks.{lvl}.{d}", dim, cfg.num_heads, cfg.cuboids(lvl))
    for lvl, dim in enumerate((u0, u1)):
        s += _resblock3d(f"down_time_embed_blocks.{lvl}", dim, dim, te)
    tokens = (cfg.h // 2) * (cfg.w // 2) + 1
    s += [("out.0.weight", (u1,)), ("out.0.bias", (u1,)), ("out.2.positional_embedding", (u1, tokens)),
          ("out.2.qkv_proj.weight", (3 * u1, u1, 1)), ("out.2.qkv_proj.bias", (3 * u1,)),
          ("out.2.c_proj.weight", (1, u1, 1)), ("out.2.c_proj.bias", (1,))]
    return s


def _is_norm_weight(name: str) -> bool:
    parts = name.split(".")
    if parts[-1] != "weight":
        return False
    owner = parts[-2]
    if owner in ("norm", "layer_norm", "norm1", "norm2", "group_norm", "conv_norm_out", "global_vec_norm"):
        return True
    if name == "out.0.weight":  # GroupNorm of the knowledge-alignment read-out head
        return True
    # GroupNorm inside the nn.Sequential of TimeEmbedResBlock: in_layers.0 / out_layers.0
    return len(parts) >= 3 and parts[-3] in ("in_layers", "out_layers") and owner == "0"


def seeded_state_dict(spec: Spec, seed: int, gain: float = 1.0) -> Dict[str, np.ndarray]:
    """Deterministic, platform-independent fp32 weights for a spec (one PCG64 stream per tensor)."""
    out: Dict[str, np.ndarray] = {}
    for idx, (name, shape) in enumerate(spec):
        rng = np.random.Generator(np.random.PCG64([seed, idx]))
        x = rng.standard_normal(shape, dtype=np.float32)
        if _is_norm_weight(name):
            x = 1.0 + 0.1 * x
        elif name.endswith(".bias"):
            x = 0.02 * x
        elif name.endswith("relative_position_bias_table"):
            x = 0.3 * x
        elif "embed.weight" in name and name.startswith("pos_embed"):
            x = 0.1 * x
        elif name.endswith("positional_embedding"):
            x = x * np.float32(shape[0] ** -0.5)
        else:
            fan_in = int(np.prod(shape[1:]))
            x = x * np.float32(gain / np.sqrt(fan_in))
        out[name] = np.ascontiguousarray(x, dtype=np.float32)
    return out


def relative_position_index(cuboid) -> np.ndarray:
    """The derived int64 buffer of CuboidSelfAttentionLayer (cuboid_transformer.py:719-734)."""
    bt, bh, bw = cuboid
    t, h, w = np.meshgrid(np.arange(bt), np.arange(bh), np.arange(bw), indexing="ij")
    c = np.stack([t.ravel(), h.ravel(), w.ravel()])  # (3, vol)
    rel = c[:, :, None] - c[:, None, :]
    rel = rel.transpose(1, 2, 0).copy()
    rel[..., 0] += bt - 1
    rel[..., 1] += bh - 1
    rel[..., 2] += bw - 1
    rel[..., 0] *= (2 * bh - 1) * (2 * bw - 1)
    rel[..., 1] *= (2 * bw - 1)
    return rel.sum(-1).astype(np.int64)
