"""CuboidTransformerUNet - host-side mirror of the reference denoiser
(src/prediff/models/cuboid_transformer/cuboid_transformer_unet.py:23-493) over the CUDA implementation.

Same constructor argument names, `forward(x, t, cond, verbose=False)` contract, `state_dict()` key names and
shapes. Built: two levels, patch-merge / upsample, GELU FFN, relative position bias, global vectors (num_global_vectors
<= 32, shared global_qkv net or separate_global_qkv=True, global_dim_ratio=1), and every
registered `block_attn_patterns` name (axial - the shipped SEVIR-LR config, on its own fast path - full, divided_st,
video_swin_PxM, spatial_lg_M, axial_space_dilate_K; prediff_b200/patterns.py) with 'zeros', 'ignore' or 'nearest' padding;
anything else raises NotImplementedError at construction - there is no fallback path.
"""
import ctypes
from typing import Sequence

import torch
from torch import nn

from . import _lib as L
from .module_tree import build_param_tree
from . import patterns as _patterns
from .weights import UNetConfig, relative_position_index, unet_param_spec


class _CUnetConfig(ctypes.Structure):
    _fields_ = [("t_in", ctypes.c_int32), ("t_out", ctypes.c_int32), ("h", ctypes.c_int32), ("w", ctypes.c_int32),
                ("c", ctypes.c_int32), ("base_units", ctypes.c_int32), ("depth", ctypes.c_int32 * 2),
                ("num_heads", ctypes.c_int32), ("max_batch", ctypes.c_int32)]


_MAX_LAYERS = 8  # PD_MAX_ATTN_LAYERS


class _CUnetPattern(ctypes.Structure):
    _fields_ = [("n_layers", ctypes.c_int32 * 2), ("cuboid_size", ctypes.c_int32 * 3 * _MAX_LAYERS * 2),
                ("strategy", ctypes.c_int32 * 3 * _MAX_LAYERS * 2), ("shift_size", ctypes.c_int32 * 3 * _MAX_LAYERS * 2),
                ("padding_type", ctypes.c_int32)]


def _unsupported(what):
    raise NotImplementedError(f"prediff_b200.CuboidTransformerUNet: {what} is not built (only the SEVIR-LR "
                              "configuration family: registered self-attention patterns, 2 levels, global vectors with "
                              "global_dim_ratio=1)")


class CuboidTransformerUNet(nn.Module):

    def __init__(self, input_shape, target_shape, base_units=128, block_units=None, scale_alpha=1.0,
                 depth=(4, 4), downsample=2, downsample_type="patch_merge", upsample_type="upsample",
                 upsample_kernel_size=3, block_attn_patterns="axial", block_cuboid_size=((4, 4, 4), (4, 4, 4)),
                 block_cuboid_strategy=(("l", "l", "l"), ("d", "d", "d")),
                 block_cuboid_shift_size=((0, 0, 0), (0, 0, 0)), num_heads=4, attn_drop=0.0, proj_drop=0.0,
                 ffn_drop=0.0, ffn_activation="gelu", gated_ffn=False, norm_layer="layer_norm", use_inter_ffn=True,
                 hierarchical_pos_embed=False, pos_embed_type="t+h+w", padding_type="zeros", checkpoint_level=0,
                 use_relative_pos=True, self_attn_use_final_proj=True, num_global_vectors=0, use_global_vector_ffn=True,
                 use_global_self_attn=False, separate_global_qkv=False, global_dim_ratio=1,
                 time_embed_channels_mult=4, time_embed_use_scale_shift_norm=False, time_embed_dropout=0.0,
                 unet_res_connect=True, max_batch=32, precision=None, streamk_ctas_per_sample=0, **ignored_init_modes):
        """`precision` (not a reference argument): "bf16" (default; env PD_PRECISION overrides the default) or "tf32" -
        the operand precision of the tensor-core GEMMs, see pd_unet_set_precision in include/prediff_b200.h.
        `streamk_ctas_per_sample` (not a reference argument): 0 = default cut (batch-invariant results); e.g. 72 for a model
        that only serves single samples (pd_unet_set_streamk_ctas)."""
        super().__init__()
        import os
        precision = precision or os.environ.get("PD_PRECISION", "bf16")
        if precision not in ("bf16", "tf32"):
            raise ValueError(f"precision must be 'bf16' or 'tf32', got {precision!r}")
        self.precision = precision
        self.streamk_ctas_per_sample = int(streamk_ctas_per_sample)
        T_in, H, W, C = input_shape
        T_out, H2, W2, C2 = target_shape
        assert (H, W, C) == (H2, W2, C2)
        patterns = block_attn_patterns if isinstance(block_attn_patterns, (list, tuple)) else [block_attn_patterns] * len(depth)
        if len(depth) != 2:
            _unsupported(f"depth={list(depth)} (needs exactly two levels)")
        explicit = None
        if block_attn_patterns is None:
            # explicit lists, shared by all blocks or one list per block (cuboid_transformer_unet.py:215-232)
            def per_block(v):
                return [list(v)] * 2 if not isinstance(v[0][0], (list, tuple)) else [list(b) for b in v]
            sizes, strats, shifts = per_block(block_cuboid_size), per_block(block_cuboid_strategy), per_block(block_cuboid_shift_size)
            if not (len(sizes) == len(strats) == len(shifts) == 2):
                _unsupported("block_cuboid_* lists must describe exactly two blocks")
            explicit = tuple(tuple((tuple(int(x) for x in a), tuple(b), tuple(int(x) for x in c))
                                   for a, b, c in zip(sizes[i], strats[i], shifts[i])) for i in range(2))
            patterns = ["explicit", "explicit"]
        else:
            for name in patterns:
                try:
                    _patterns.get(name)
                except KeyError:
                    _unsupported(f"block_attn_patterns={patterns}")
        if padding_type not in ("zeros", "ignore", "nearest"):
            _unsupported(f"padding_type='{padding_type}'")
        if block_units is not None and list(block_units) != [base_units, 2 * base_units]:
            _unsupported(f"block_units={block_units}")
        checks = [(scale_alpha == 1.0, "scale_alpha != 1"), (downsample in (2, (1, 2, 2), [1, 2, 2]), "downsample != 2"),
                  (downsample_type == "patch_merge", "downsample_type"), (upsample_type == "upsample", "upsample_type"),
                  (upsample_kernel_size == 3, "upsample_kernel_size"), (ffn_activation == "gelu", "ffn_activation"),
                  (not gated_ffn, "gated_ffn"), (norm_layer == "layer_norm", "norm_layer"), (use_inter_ffn, "use_inter_ffn"),
                  (not hierarchical_pos_embed, "hierarchical_pos_embed"), (pos_embed_type == "t+h+w", "pos_embed_type"),
                  (use_relative_pos, "use_relative_pos=False"), (self_attn_use_final_proj, "self_attn_use_final_proj"),
                  (0 <= int(num_global_vectors or 0) <= 32, "num_global_vectors > 32"),
                  (not num_global_vectors or global_dim_ratio == 1, "global_dim_ratio != 1"),
                  (not num_global_vectors or precision == "bf16", "global vectors with precision='tf32'"),
                  (time_embed_channels_mult == 4, "time_embed_channels_mult"),
                  (not time_embed_use_scale_shift_norm, "scale-shift norm"), (unet_res_connect, "unet_res_connect=False")]
        for ok, what in checks:
            if not ok:
                _unsupported(what)
        self.cfg = UNetConfig(t_in=T_in, t_out=T_out, h=H, w=W, c=C, base_units=base_units, depth=tuple(depth),
                              num_heads=num_heads, patterns=tuple(patterns), padding_type=padding_type,
                              explicit_layers=explicit, num_global_vectors=int(num_global_vectors or 0),
                              use_global_vector_ffn=bool(use_global_vector_ffn), use_global_self_attn=bool(use_global_self_attn),
                              separate_global_qkv=bool(separate_global_qkv) and bool(num_global_vectors))
        self.num_global_vectors = self.cfg.num_global_vectors
        self.use_global_vector = self.num_global_vectors > 0
        for lvl in range(2):
            if len(self.cfg.layers(lvl)) > _MAX_LAYERS:
                _unsupported(f"{len(self.cfg.layers(lvl))} attention layers per block")
        self.input_shape, self.target_shape = list(input_shape), list(target_shape)
        self.in_len, self.out_len = T_in, T_out
        self.max_batch = max_batch
        bufs = {}
        for name in ("down_self_blocks", "up_self_blocks"):
            for lvl in range(2):
                for d in range(self.cfg.depth[lvl]):
                    for i, cub in enumerate(self.cfg.cuboids(lvl)):
                        bufs[f"{name}.{lvl}.{d}.attn_l.{i}.relative_position_index"] = \
                            torch.from_numpy(relative_position_index(cub))
        build_param_tree(self, unet_param_spec(self.cfg), bufs)
        self._handle = None
        self._dirty = True
        self.register_load_state_dict_post_hook(type(self)._mark_dirty)

    # ---- shape properties of the reference class (cuboid_transformer_unet.py:377-404) -------------------------
    @property
    def data_shape(self):
        c = self.cfg
        return (c.t_in + c.t_out, c.h, c.w, c.c + 1)   # + the observed / target indicator channel

    @property
    def mem_shapes(self):
        c = self.cfg
        return [(c.T, c.h, c.w, c.units[0]), (c.T, c.h // 2, c.w // 2, c.units[1])]

    # ---- C++ handle management -----------------------------------------------------------------------------
    def _ensure_handle(self):
        if self._handle is None:
            c = self.cfg
            cc = _CUnetConfig(c.t_in, c.t_out, c.h, c.w, c.c, c.base_units, (ctypes.c_int32 * 2)(*c.depth), c.num_heads,
                              self.max_batch)
            h = ctypes.c_void_p()
            if tuple(c.patterns) == ("axial", "axial") and c.padding_type == "zeros" and not c.num_global_vectors:
                L.check(L.lib().pd_unet_create(ctypes.byref(cc), ctypes.byref(h)))
            else:
                pt = _CUnetPattern()
                pt.padding_type = {"zeros": 0, "ignore": 1, "nearest": 2}[c.padding_type]
                for lvl in range(2):
                    layers = c.layers(lvl)
                    pt.n_layers[lvl] = len(layers)
                    for i, (size, strategy, shift) in enumerate(layers):
                        for a in range(3):
                            pt.cuboid_size[lvl][i][a] = size[a]
                            pt.strategy[lvl][i][a] = 0 if strategy[a] == "l" else 1
                            pt.shift_size[lvl][i][a] = shift[a]
                L.check(L.lib().pd_unet_create_gv_ex(ctypes.byref(cc), ctypes.byref(pt), c.num_global_vectors,
                                                     int(c.use_global_vector_ffn), int(c.use_global_self_attn),
                                                     int(c.separate_global_qkv), ctypes.byref(h)))
            self._handle = h
            self._dirty = True
        return self._handle

    def refresh(self):
        """Pushes the current parameter values to the CUDA side and repacks them (bf16, K-major, tap-major)."""
        h = self._ensure_handle()
        lib = L.lib()
        for name, p in self.named_parameters():
            t = p.detach().contiguous().float()
            shape = (ctypes.c_int64 * t.dim())(*t.shape)
            L.check(lib.pd_unet_load_weight(h, name.encode(), L.ptr(t), shape, t.dim()))
        L.check(lib.pd_unet_set_precision(h, 1 if self.precision == "tf32" else 0))
        L.check(lib.pd_unet_set_streamk_ctas(h, self.streamk_ctas_per_sample))
        L.check(lib.pd_unet_finalize(h))
        self._dirty = False

    def _mark_dirty(self, *unused):
        """Parameters changed: the packed CUDA copies are stale. Registered as a load_state_dict post-hook, which torch
        runs for this module also when a PARENT's load_state_dict recurses through it (nn.Module.load_state_dict never
        calls a child's overridden load_state_dict)."""
        self._dirty = True

    def _apply(self, fn, *a, **kw):
        r = super()._apply(fn, *a, **kw)
        self._dirty = True
        return r

    def __del__(self):
        try:
            if self._handle is not None:
                L.lib().pd_unet_destroy(self._handle)
        except Exception:
            pass

    def weight_spec_from_library(self):
        """(name, shape) list as declared by the C++ model - must equal weights.unet_param_spec."""
        h = self._ensure_handle()
        lib = L.lib()
        out = []
        for i in range(lib.pd_unet_num_weights(h)):
            name = ctypes.c_char_p()
            shape = (ctypes.c_int64 * 5)()
            nd = lib.pd_unet_weight_info(h, i, ctypes.byref(name), shape)
            out.append((name.value.decode(), tuple(shape[:nd])))
        return out

    @property
    def handle(self):
        if self._dirty:
            self.refresh()
        return self._handle

    # ---- reference call surface -----------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, t, cond, verbose=False):
        """x (B, T_out, H, W, C) fp32, t (B,) int64, cond (B, T_in, H, W, C) fp32 -> (B, T_out, H, W, C).
        Reference: cuboid_transformer_unet.py:406-493."""
        c = self.cfg
        B = x.shape[0]
        if not x.is_cuda:
            raise L.PDError("prediff_b200.CuboidTransformerUNet runs on a CUDA (sm_100) device only; got a CPU tensor")
        assert tuple(x.shape[1:]) == (c.t_out, c.h, c.w, c.c), f"x shape {tuple(x.shape)}"
        assert tuple(cond.shape) == (B, c.t_in, c.h, c.w, c.c), f"cond shape {tuple(cond.shape)}"
        assert t.shape == (B,)
        x = x.contiguous().float()
        cond = cond.contiguous().float()
        t = t.to(device=x.device, dtype=torch.int64).contiguous()
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            L.check(L.lib().pd_unet_forward(self.handle, L.ptr(x), L.ptr(t), L.ptr(cond), L.ptr(out), B, L.stream_ptr()))
        return out
