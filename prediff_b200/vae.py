"""AutoencoderKL - host-side mirror of the reference first-stage model
(src/prediff/taming/autoencoder_kl.py:9-113) over the CUDA implementation: same constructor argument names,
`encode(x) -> posterior` / `decode(z) -> tensor` contract in the reference's NCHW layout, same state_dict keys."""
import ctypes

import torch
from torch import nn

from . import _lib as L
from .distributions import reference_distribution_class
from .module_tree import build_param_tree
from .weights import VAEConfig, vae_param_spec


class _CVaeConfig(ctypes.Structure):
    _fields_ = [("in_channels", ctypes.c_int32), ("out_channels", ctypes.c_int32), ("latent_channels", ctypes.c_int32),
                ("block_out_channels", ctypes.c_int32 * 4), ("layers_per_block", ctypes.c_int32),
                ("norm_num_groups", ctypes.c_int32), ("h", ctypes.c_int32), ("w", ctypes.c_int32),
                ("max_frames", ctypes.c_int32)]


class AutoencoderKL(nn.Module):

    def __init__(self, in_channels=1, out_channels=1, down_block_types=("DownEncoderBlock2D",) * 4,
                 up_block_types=("UpDecoderBlock2D",) * 4, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 act_fn="silu", latent_channels=64, norm_num_groups=32, sample_size=128, scaling_factor=0.18215,
                 max_frames=64):
        super().__init__()
        if list(down_block_types) != ["DownEncoderBlock2D"] * 4 or list(up_block_types) != ["UpDecoderBlock2D"] * 4:
            raise NotImplementedError("prediff_b200.AutoencoderKL: only 4 x DownEncoderBlock2D / UpDecoderBlock2D is built")
        if act_fn != "silu" or len(block_out_channels) != 4:
            raise NotImplementedError("prediff_b200.AutoencoderKL: act_fn must be 'silu' with four resolution levels")
        if isinstance(sample_size, int):
            sample_size = (sample_size, sample_size)
        self.cfg = VAEConfig(in_channels=in_channels, out_channels=out_channels, latent_channels=latent_channels,
                             block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                             norm_num_groups=norm_num_groups, h=sample_size[0], w=sample_size[1])
        self.max_frames = max_frames
        self.use_slicing = False
        build_param_tree(self, vae_param_spec(self.cfg))
        self._handle = None
        self._dirty = True
        self.register_load_state_dict_post_hook(type(self)._mark_dirty)

    def _ensure_handle(self):
        if self._handle is None:
            c = self.cfg
            cc = _CVaeConfig(c.in_channels, c.out_channels, c.latent_channels, (ctypes.c_int32 * 4)(*c.block_out_channels),
                             c.layers_per_block, c.norm_num_groups, c.h, c.w, self.max_frames)
            h = ctypes.c_void_p()
            L.check(L.lib().pd_vae_create(ctypes.byref(cc), ctypes.byref(h)))
            self._handle = h
            self._dirty = True
        return self._handle

    def refresh(self):
        h = self._ensure_handle()
        lib = L.lib()
        for name, p in self.named_parameters():
            t = p.detach().contiguous().float()
            shape = (ctypes.c_int64 * t.dim())(*t.shape)
            L.check(lib.pd_vae_load_weight(h, name.encode(), L.ptr(t), shape, t.dim()))
        L.check(lib.pd_vae_finalize(h))
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
                L.lib().pd_vae_destroy(self._handle)
        except Exception:
            pass

    def weight_spec_from_library(self):
        h = self._ensure_handle()
        lib = L.lib()
        out = []
        for i in range(lib.pd_vae_num_weights(h)):
            name = ctypes.c_char_p()
            shape = (ctypes.c_int64 * 5)()
            nd = lib.pd_vae_weight_info(h, i, ctypes.byref(name), shape)
            out.append((name.value.decode(), tuple(shape[:nd])))
        return out

    @property
    def handle(self):
        if self._dirty:
            self.refresh()
        return self._handle

    def _check(self, t):
        if not t.is_cuda:
            raise L.PDError("prediff_b200.AutoencoderKL runs on a CUDA (sm_100) device only; got a CPU tensor")

    # ---- reference call surface -----------------------------------------------------------------------------
    @torch.no_grad()
    def encode_moments(self, x):
        """x (N, 1, H, W) -> moments (N, 2*latent, H/8, W/8) = quant_conv(Encoder(x)) (autoencoder_kl.py:80-82)."""
        self._check(x)
        c = self.cfg
        N = x.shape[0]
        assert tuple(x.shape[1:]) == (c.in_channels, c.h, c.w), f"x shape {tuple(x.shape)}"
        x = x.contiguous().float()  # (N,1,H,W) is bit-identical to channels-last [N][H][W]
        out = torch.empty(N, c.h // 8, c.w // 8, 2 * c.latent_channels, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            for i in range(0, N, self.max_frames):
                n = min(self.max_frames, N - i)
                L.check(L.lib().pd_vae_encode(self.handle, L.ptr(x[i:i + n]), L.ptr(out[i:i + n]), n, L.stream_ptr()))
        return out.permute(0, 3, 1, 2)

    def enable_slicing(self):
        """autoencoder_kl.py:91-98. Decoding already walks the batch in slices of `max_frames` frames with identical
        results, so the flag only mirrors the reference attribute."""
        self.use_slicing = True

    def disable_slicing(self):
        self.use_slicing = False

    def encode(self, x):
        """Returns the posterior (mode() = mean = first latent_channels channels; distributions.py:70-71)."""
        return reference_distribution_class()(self.encode_moments(x))

    @torch.no_grad()
    def decode(self, z, out=None):
        """z (N, latent, H/8, W/8) -> (N, 1, H, W) = Decoder(post_quant_conv(z)) (autoencoder_kl.py:86-113).
        `out` (optional, contiguous fp32 (N, 1, H, W)): the decoder's last kernel writes the frames there - e.g. this
        rank's slice of the ensemble's all-gather buffer, so the collective needs no staging copy."""
        self._check(z)
        c = self.cfg
        N = z.shape[0]
        assert tuple(z.shape[1:]) == (c.latent_channels, c.h // 8, c.w // 8), f"z shape {tuple(z.shape)}"
        zl = z.permute(0, 2, 3, 1).contiguous().float()
        if out is None:
            out = torch.empty(N, 1, c.h, c.w, device=z.device, dtype=torch.float32)
        else:
            assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and \
                tuple(out.shape) == (N, 1, c.h, c.w), f"decode(out=): bad buffer {tuple(out.shape)} {out.dtype}"
        with torch.cuda.device(z.device):
            for i in range(0, N, self.max_frames):
                n = min(self.max_frames, N - i)
                L.check(L.lib().pd_vae_decode(self.handle, L.ptr(zl[i:i + n]), L.ptr(out[i:i + n]), n, L.stream_ptr()))
        return out

    def forward(self, sample, sample_posterior=False, return_posterior=False, generator=None):
        posterior = self.encode(sample)
        z = posterior.sample(generator=generator) if sample_posterior else posterior.mode()
        dec = self.decode(z)
        return (dec, posterior) if return_posterior else dec
