"""Knowledge alignment - host-side mirrors of the reference's guidance objects over the CUDA implementation.

* `NoisyCuboidTransformerEncoder`  <- src/prediff/diffusion/knowledge_alignment/models.py:106-528 (U(z_t, t))
* `SEVIRAvgIntensityAlignment`     <- src/prediff/diffusion/knowledge_alignment/sevir.py:7-104

Same constructor argument names, `state_dict()` key names / shapes and call signatures, so
`LatentDiffusion.set_alignment(alignment_obj.get_mean_shift)` (train_sevirlr_prediff.py:190-206) works unchanged.
The reference obtains the guidance with `torch.autograd.grad` through the network (alignment_pl.py:423-446); here
`get_mean_shift` is one C-ABI call (`pd_ka_mean_shift`) that runs the forward and a hand-written input-gradient
backward on the device. Only the shipped configuration family is built (axial pattern, two levels, patch-merge,
attention-pool read-out with readout_seq); anything else raises NotImplementedError - there is no fallback.
"""
import ctypes
from typing import Any, Dict

import torch
from torch import nn

from . import _lib as L
from .module_tree import build_param_tree
from .weights import KAConfig, ka_param_spec, relative_position_index


class _CKAConfig(ctypes.Structure):
    _fields_ = [("t", ctypes.c_int32), ("h", ctypes.c_int32), ("w", ctypes.c_int32), ("c", ctypes.c_int32),
                ("base_units", ctypes.c_int32), ("depth", ctypes.c_int32 * 2), ("num_heads", ctypes.c_int32),
                ("max_batch", ctypes.c_int32)]


def _unsupported(what):
    raise NotImplementedError(f"prediff_b200.NoisyCuboidTransformerEncoder: {what} is not built (only the shipped "
                              "SEVIR-LR alignment configuration: axial pattern, 2 levels, attention-pool read-out)")


class NoisyCuboidTransformerEncoder(nn.Module):

    def __init__(self, input_shape, out_channels=1, base_units=128, block_units=None, scale_alpha=1.0, depth=(1, 1),
                 downsample=2, downsample_type="patch_merge", block_attn_patterns="axial", num_heads=4, attn_drop=0.0,
                 proj_drop=0.0, ffn_drop=0.0, ffn_activation="gelu", gated_ffn=False, norm_layer="layer_norm",
                 use_inter_ffn=True, hierarchical_pos_embed=False, pos_embed_type="t+h+w", padding_type="zeros",
                 checkpoint_level=0, use_relative_pos=True, self_attn_use_final_proj=True, num_global_vectors=0,
                 use_global_vector_ffn=True, use_global_self_attn=False, separate_global_qkv=False, global_dim_ratio=1,
                 time_embed_channels_mult=4, time_embed_use_scale_shift_norm=False, time_embed_dropout=0.0,
                 pool="attention", readout_seq=True, out_len=None, max_batch=32, **ignored_init_modes):
        super().__init__()
        T, H, W, C = input_shape
        patterns = block_attn_patterns if isinstance(block_attn_patterns, (list, tuple)) else [block_attn_patterns] * len(depth)
        if len(depth) != 2:
            _unsupported(f"depth={list(depth)} (needs exactly two levels)")
        if any(p != "axial" for p in patterns):
            _unsupported(f"block_attn_patterns={patterns}")
        if block_units is not None and list(block_units) != [base_units, 2 * base_units]:
            _unsupported(f"block_units={block_units}")
        checks = [(out_channels == 1, "out_channels != 1"), (scale_alpha == 1.0, "scale_alpha != 1"),
                  (downsample in (2, (1, 2, 2), [1, 2, 2]), "downsample != 2"),
                  (downsample_type == "patch_merge", "downsample_type"), (ffn_activation == "gelu", "ffn_activation"),
                  (not gated_ffn, "gated_ffn"), (norm_layer == "layer_norm", "norm_layer"), (use_inter_ffn, "use_inter_ffn"),
                  (not hierarchical_pos_embed, "hierarchical_pos_embed"), (pos_embed_type == "t+h+w", "pos_embed_type"),
                  (use_relative_pos, "use_relative_pos=False"), (self_attn_use_final_proj, "self_attn_use_final_proj"),
                  (not num_global_vectors, "global vectors"), (time_embed_channels_mult == 4, "time_embed_channels_mult"),
                  (not time_embed_use_scale_shift_norm, "scale-shift norm"), (pool == "attention", f"pool={pool}"),
                  (readout_seq, "readout_seq=False"), (out_len in (None, T), f"out_len={out_len} != T")]
        for ok, what in checks:
            if not ok:
                _unsupported(what)
        self.cfg = KAConfig(t=T, h=H, w=W, c=C, base_units=base_units, depth=tuple(depth), num_heads=num_heads)
        self.input_shape, self.out_channels, self.out_len = list(input_shape), out_channels, T
        self.max_batch = max_batch
        bufs = {}
        for lvl in range(2):
            for d in range(self.cfg.depth[lvl]):
                for i, cub in enumerate(self.cfg.cuboids(lvl)):
                    bufs[f"down_self_blocks.{lvl}.{d}.attn_l.{i}.relative_position_index"] = \
                        torch.from_numpy(relative_position_index(cub))
        build_param_tree(self, ka_param_spec(self.cfg), bufs)
        self._handle = None
        self._dirty = True
        self.register_load_state_dict_post_hook(type(self)._mark_dirty)

    # ---- C++ handle management ------------------------------------------------------------------------------
    def _ensure_handle(self):
        if self._handle is None:
            c = self.cfg
            cc = _CKAConfig(c.t, c.h, c.w, c.c, c.base_units, (ctypes.c_int32 * 2)(*c.depth), c.num_heads, self.max_batch)
            h = ctypes.c_void_p()
            L.check(L.lib().pd_ka_create(ctypes.byref(cc), ctypes.byref(h)))
            self._handle = h
            self._dirty = True
        return self._handle

    def refresh(self):
        """Pushes the parameters to the CUDA side and repacks them (forward and dgrad layouts)."""
        h = self._ensure_handle()
        lib = L.lib()
        for name, p in self.named_parameters():
            t = p.detach().contiguous().float()
            shape = (ctypes.c_int64 * t.dim())(*t.shape)
            L.check(lib.pd_ka_load_weight(h, name.encode(), L.ptr(t), shape, t.dim()))
        L.check(lib.pd_ka_finalize(h))
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
                L.lib().pd_ka_destroy(self._handle)
        except Exception:
            pass

    def weight_spec_from_library(self):
        h = self._ensure_handle()
        lib = L.lib()
        out = []
        for i in range(lib.pd_ka_num_weights(h)):
            name = ctypes.c_char_p()
            shape = (ctypes.c_int64 * 5)()
            nd = lib.pd_ka_weight_info(h, i, ctypes.byref(name), shape)
            out.append((name.value.decode(), tuple(shape[:nd])))
        return out

    @property
    def handle(self):
        if self._dirty:
            self.refresh()
        return self._handle

    def _check(self, x, t):
        c = self.cfg
        if not x.is_cuda:
            raise L.PDError("prediff_b200.NoisyCuboidTransformerEncoder runs on a CUDA (sm_100) device only")
        assert tuple(x.shape[1:]) == (c.t, c.h, c.w, c.c), f"x shape {tuple(x.shape)}"
        assert t.shape == (x.shape[0],)
        return x.detach().contiguous().float(), t.to(device=x.device, dtype=torch.int64).contiguous()

    # ---- reference call surface -----------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, t, verbose=False, **kwargs):
        """x (B, T, H, W, C), t (B,) -> (B, T, 1)   (models.py:459-528)."""
        x, t = self._check(x, t)
        B = x.shape[0]
        out = torch.empty(B, self.cfg.t, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            L.check(L.lib().pd_ka_forward(self.handle, L.ptr(x), L.ptr(t), L.ptr(out), B, L.stream_ptr()))
        return out.unsqueeze(-1)

    @torch.no_grad()
    def mean_shift(self, zt, t, avg_x_gt, guide_scale, return_value=False):
        """guide_scale * d || mean_T U(zt, t) - avg_x_gt ||_2 / d zt, norm over the whole batch (sevir.py:76-104)."""
        zt, t = self._check(zt, t)
        B = zt.shape[0]
        target = torch.as_tensor(avg_x_gt, dtype=torch.float32, device=zt.device).reshape(-1)
        if target.numel() == 1:
            target = target.expand(B)
        assert target.numel() == B, f"avg_x_gt must have one value per sample, got {tuple(target.shape)}"
        target = target.contiguous()
        grad = torch.empty_like(zt)
        val = torch.empty(1, device=zt.device, dtype=torch.float32)
        with torch.cuda.device(zt.device):
            L.check(L.lib().pd_ka_mean_shift(self.handle, L.ptr(zt), L.ptr(t), L.ptr(target), ctypes.c_float(guide_scale),
                                             L.ptr(grad), L.ptr(val), B, L.stream_ptr()))
        return (grad, val) if return_value else grad


class SEVIRAvgIntensityAlignment:

    def __init__(self, alignment_type: str = "avg_x", guide_scale: float = 1.0, model_type: str = "cuboid",
                 model_args: Dict[str, Any] = None, model_ckpt_path: str = None):
        assert alignment_type in ["avg_x"], f"alignment_type {alignment_type} is not supported"
        self.alignment_type = alignment_type
        self.guide_scale = guide_scale
        if model_type != "cuboid":
            raise NotImplementedError(f"model_type={model_type} is not implemented")
        self.model = NoisyCuboidTransformerEncoder(**(model_args or {}))
        if model_ckpt_path is not None:
            self.model.load_state_dict(torch.load(model_ckpt_path, map_location="cpu"))

    @classmethod
    def model_objective(cls, x, y=None, **kwargs):
        """(b t h w c) -> (b t 1): the quantity U is trained to predict (sevir.py:41-53)."""
        return torch.mean(x, dim=[2, 3, 4], keepdim=False).unsqueeze(-1)

    def alignment_fn(self, zt, t, y=None, zc=None, **kwargs):
        """|| mean_T U(zt, t) - avg_x_gt ||_2 (sevir.py:55-83); value only - the gradient comes from get_mean_shift."""
        pred = self.model(zt, t).mean(dim=1)
        target = torch.as_tensor(kwargs.get("avg_x_gt"), dtype=torch.float32, device=pred.device)
        return torch.linalg.vector_norm(pred - target, ord=2)

    def get_mean_shift(self, zt, t, y=None, zc=None, **kwargs):
        """guide_scale * grad_zt alignment_fn (sevir.py:85-104)."""
        if "avg_x_gt" not in kwargs:
            raise KeyError("get_mean_shift: alignment kwarg `avg_x_gt` is required (sevir.py:78)")
        return self.model.mean_shift(zt, t, kwargs["avg_x_gt"], self.guide_scale)
