"""LatentDiffusion - host-side mirror of the reference sampling runtime
(src/prediff/diffusion/latent_diffusion.py:25-724) over the device-resident CUDA loop.

Keeps the reference call surface used by the Lightning script: constructor injection of
`torch_nn_module` / `first_stage_model`, `sample`, `p_sample_loop`, `p_sample`, `apply_model`,
`encode_first_stage` / `decode_first_stage`, `cond_stage_forward`, `set_alignment`, the registered schedule
buffers - plus `ddim_sample_loop`, the 50-step DDIM the benchmark is quoted on (the reference has no DDIM
sampler; SURVEY.md D1 / section 8 row S6 define it from the reference's own helper functions).

Also mirrored, because the inference script's validation_step uses them: `forward` / `p_losses` (forward loss, no
gradients) and `ema_scope` / `model_ema` (EMA weights swapped in for evaluation). Optimizers, backward passes and
Lightning hooks are out of scope.
"""
import contextlib
import ctypes
from typing import Any, Callable, Dict, Optional, Sequence

import numpy as np
import torch
from torch import nn

from . import _lib as L
from .distributions import DiagonalGaussianDistribution

PD_MODE_DDPM, PD_MODE_DDIM = 0, 1
_SCHEDULE_BUFFERS = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
                     "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
                     "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                     "posterior_mean_coef1", "posterior_mean_coef2"]


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=False):
    """diffusion/utils.py:42-56, 'uniform' only."""
    if ddim_discr_method != "uniform":
        raise NotImplementedError(ddim_discr_method)
    c = num_ddpm_timesteps // num_ddim_timesteps
    return np.asarray(list(range(0, num_ddpm_timesteps, c))) + 1


class LatentDiffusion(nn.Module):

    def __init__(self, torch_nn_module: nn.Module, layout: str = "NTHWC", data_shape: Sequence[int] = (6, 128, 128, 1),
                 timesteps=1000, beta_schedule="linear", loss_type="l2", monitor="val/loss", use_ema=False,
                 log_every_t=100, clip_denoised=False, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3,
                 given_betas=None, original_elbo_weight=0., v_posterior=0., l_simple_weight=1., parameterization="eps",
                 learn_logvar=False, logvar_init=0., latent_shape: Sequence[int] = (6, 16, 16, 64),
                 first_stage_model: nn.Module = None, cond_stage_model=None, num_timesteps_cond=None,
                 cond_stage_trainable=False, cond_stage_forward=None, scale_by_std=False, scale_factor=1.0):
        super().__init__()
        if parameterization != "eps":
            raise NotImplementedError("prediff_b200.LatentDiffusion: only eps-parameterization is built")
        if beta_schedule != "linear" or given_betas is not None or v_posterior != 0.:
            raise NotImplementedError("prediff_b200.LatentDiffusion: only the 'linear' beta schedule with v_posterior=0")
        if layout != "NTHWC":
            raise NotImplementedError("prediff_b200.LatentDiffusion: layout must be 'NTHWC'")
        if scale_by_std:   # the reference registers a 'scale_factor' buffer (a state_dict key) and rescales on the first batch
            raise NotImplementedError("prediff_b200.LatentDiffusion: scale_by_std=True is not built")
        if cond_stage_forward not in (None, "encode"):
            raise NotImplementedError(f"prediff_b200.LatentDiffusion: cond_stage_forward='{cond_stage_forward}' is not built "
                                      "(the context is always encoded with first_stage_model.encode)")
        self.parameterization = parameterization
        self.clip_denoised = clip_denoised
        self.log_every_t = log_every_t
        self.torch_nn_module = torch_nn_module
        self.layout = layout
        self.data_shape = tuple(data_shape)
        self.latent_shape = tuple(latent_shape)
        self.batch_axis, self.t_axis, self.h_axis, self.w_axis, self.c_axis = 0, 1, 2, 3, 4
        # EMA shadow weights (latent_diffusion.py:128-131): kept under the reference's buffer names so its checkpoints
        # load; evaluation swaps them in through ema_scope()
        self.use_ema = bool(use_ema)
        if self.use_ema:
            from .ema import LitEma
            self.model_ema = LitEma(self.torch_nn_module, include_frozen=hasattr(self.torch_nn_module, "_dirty"))
        self.scale_factor = scale_factor
        self.alignment_fn = None
        # latent_diffusion.py:153-157: with num_timesteps_cond > 1 the sampling loop re-noises the context every step
        self.num_timesteps_cond = 1 if num_timesteps_cond is None else int(num_timesteps_cond)
        assert self.num_timesteps_cond <= timesteps
        self.shorten_cond_schedule = self.num_timesteps_cond > 1
        self.cond_stage_trainable = False   # forced off for '__is_first_stage__' in the reference too (:338-341)
        # loss hyper-parameters of p_losses (latent_diffusion.py:134-150); only the forward (validation) loss is built
        if loss_type not in ("l1", "l2"):
            raise NotImplementedError(f"prediff_b200.LatentDiffusion: unknown loss type '{loss_type}'")
        self.loss_type = loss_type
        self.original_elbo_weight = original_elbo_weight
        self.l_simple_weight = l_simple_weight
        self.learn_logvar = bool(learn_logvar)   # the shipped config sets it (cfg.yaml:95); the values come from the checkpoint
        self.logvar_init = float(logvar_init)
        self.register_schedule(timesteps=timesteps, linear_start=linear_start, linear_end=linear_end)
        if self.clip_denoised:   # latent_diffusion.py:580-581, applied inside the fused update kernel
            L.check(L.lib().pd_sampler_set_clip_denoised(self._sampler, 1))
        if self.shorten_cond_schedule:
            self.make_cond_schedule()
        logvar = torch.full(fill_value=logvar_init, size=(self.num_timesteps,))
        if self.learn_logvar:   # latent_diffusion.py:146-150: a parameter when learned, a buffer otherwise (same key)
            self.logvar = nn.Parameter(logvar, requires_grad=True)
        else:
            self.register_buffer("logvar", logvar)
        self.first_stage_model = first_stage_model
        if first_stage_model is not None:
            first_stage_model.eval()
        if cond_stage_model == "__is_first_stage__":
            self.cond_stage_model = first_stage_model
            self._cond_is_first_stage = True
        elif cond_stage_model is None:
            self.cond_stage_model = None
            self._cond_is_first_stage = False
        else:
            raise NotImplementedError("prediff_b200.LatentDiffusion: cond_stage_model must be '__is_first_stage__' or None")

    def make_cond_schedule(self):
        """latent_diffusion.py:295-299 (buffer `cond_ids`, a state_dict key of the reference)."""
        cond_ids = torch.full(size=(self.num_timesteps,), fill_value=self.num_timesteps - 1, dtype=torch.long)
        ids = torch.round(torch.linspace(0, self.num_timesteps - 1, self.num_timesteps_cond)).long()
        cond_ids[:self.num_timesteps_cond] = ids
        self.register_buffer("cond_ids", cond_ids)

    # ---- schedule -------------------------------------------------------------------------------------------
    def register_schedule(self, given_betas=None, beta_schedule="linear", timesteps=1000, linear_start=1e-4,
                          linear_end=2e-2, cosine_s=8e-3):
        """latent_diffusion.py:228-278; the table is computed (float64) and owned by the C++ sampler."""
        h = ctypes.c_void_p()
        L.check(L.lib().pd_sampler_create(int(timesteps), ctypes.c_double(linear_start), ctypes.c_double(linear_end),
                                          ctypes.byref(h)))
        self._sampler = h
        self.num_timesteps = int(timesteps)
        self.linear_start, self.linear_end = linear_start, linear_end
        for name in _SCHEDULE_BUFFERS:
            buf = torch.empty(self.num_timesteps, dtype=torch.float32)
            L.check(L.lib().pd_sampler_get_buffer(self._sampler, name.encode(), L.ptr(buf)))
            self.register_buffer(name, buf)
        w = torch.empty(self.num_timesteps, dtype=torch.float32)
        L.check(L.lib().pd_sampler_get_buffer(self._sampler, b"lvlb_weights", L.ptr(w)))
        self.register_buffer("lvlb_weights", w, persistent=False)   # latent_diffusion.py:277

    def __del__(self):
        try:
            if getattr(self, "_sampler", None) is not None:
                L.lib().pd_sampler_destroy(self._sampler)
        except Exception:
            pass

    def set_alignment(self, alignment_fn: Callable = None):
        """latent_diffusion.py:169-180. Signature `alignment_fn(zt, t, zc=None, y=None, **kwargs)`."""
        self.alignment_fn = alignment_fn

    def _weights_changed(self):
        """The CUDA denoiser repacks its weights lazily; parameter writes through `.data` must be announced."""
        if hasattr(self.torch_nn_module, "_dirty"):
            self.torch_nn_module._dirty = True

    @contextlib.contextmanager
    def ema_scope(self, context=None):
        """latent_diffusion.py:280-293: run the body with the EMA weights in the denoiser, then restore."""
        if self.use_ema:
            self.model_ema.store(self.torch_nn_module.parameters())
            self.model_ema.copy_to(self.torch_nn_module)
            self._weights_changed()
            if context is not None:
                print(f"{context}: Switched to EMA weights")
        try:
            yield None
        finally:
            if self.use_ema:
                self.model_ema.restore(self.torch_nn_module.parameters())
                self._weights_changed()
                if context is not None:
                    print(f"{context}: Restored training weights")

    def on_train_batch_end(self, *args, **kwargs):
        """latent_diffusion.py:484-486 (the update itself is plain tensor arithmetic; training is not built)."""
        if self.use_ema:
            self.model_ema(self.torch_nn_module)

    @property
    def einops_layout(self):
        return " ".join(self.layout)

    def extract_into_tensor(self, a, t, x_shape):
        out = a.gather(-1, t)
        return out.reshape([t.shape[0]] + [1] * (len(x_shape) - 1))

    def get_batch_latent_shape(self, batch_size=1):
        return (batch_size,) + tuple(self.latent_shape)

    def get_batch_data_shape(self, batch_size=1):
        """latent_diffusion.py:204-214."""
        return (batch_size,) + tuple(self.data_shape)

    @property
    def einops_spatial_layout(self):
        return "(N T) C H W"   # latent_diffusion.py:394-399 for the 5-d 'NTHWC' layout

    @property
    def device(self):
        """Where the denoiser's work runs (LightningModule.device in the reference)."""
        return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")

    def get_first_stage_encoding(self, encoder_posterior):
        """latent_diffusion.py:381-391: posterior sample (or the tensor itself), times scale_factor."""
        z = encoder_posterior.sample() if hasattr(encoder_posterior, "sample") else encoder_posterior
        if not torch.is_tensor(z):
            raise NotImplementedError(f"encoder_posterior of type '{type(encoder_posterior)}' not yet implemented")
        return self.scale_factor * z

    # ---- the reference's step helpers (latent_diffusion.py:553-596), plain tensor expressions on the buffers. The
    # device-resident loop evaluates the same formulas inside sampler_update_kernel; these exist for callers that use
    # them directly (score correctors, visualisation) ----
    def _buf(self, name, like):
        return getattr(self, name).to(like.device)

    def predict_start_from_noise(self, x_t, t, noise):
        return (self.extract_into_tensor(self._buf("sqrt_recip_alphas_cumprod", x_t), t, x_t.shape) * x_t -
                self.extract_into_tensor(self._buf("sqrt_recipm1_alphas_cumprod", x_t), t, x_t.shape) * noise)

    def q_posterior(self, x_start, x_t, t):
        mean = (self.extract_into_tensor(self._buf("posterior_mean_coef1", x_t), t, x_t.shape) * x_start +
                self.extract_into_tensor(self._buf("posterior_mean_coef2", x_t), t, x_t.shape) * x_t)
        var = self.extract_into_tensor(self._buf("posterior_variance", x_t), t, x_t.shape)
        log_var = self.extract_into_tensor(self._buf("posterior_log_variance_clipped", x_t), t, x_t.shape)
        return mean, var, log_var

    def p_mean_variance(self, zt, zc, t, clip_denoised: bool = False, return_x0=False, score_corrector=None,
                        corrector_kwargs=None):
        model_out = self.apply_model(zt, t, zc)
        if score_corrector is not None:
            model_out = score_corrector.modify_score(self, model_out, zt, t, zc, **(corrector_kwargs or {}))
        z_recon = self.predict_start_from_noise(zt, t=t, noise=model_out)
        if clip_denoised:
            z_recon = z_recon.clamp(-1., 1.)
        mean, var, log_var = self.q_posterior(x_start=z_recon, x_t=zt, t=t)
        return (mean, var, log_var, z_recon) if return_x0 else (mean, var, log_var)

    def aligned_mean(self, zt, t, zc, y, orig_mean, orig_log_var, **kwargs):
        align_gradient = self.alignment_fn(zt, t, zc=zc, y=y, **kwargs)
        return orig_mean - (0.5 * orig_log_var).exp() * align_gradient

    # ---- first stage glue (latent_diffusion.py:361-432) ------------------------------------------------------
    @torch.no_grad()
    def cond_stage_forward(self, c: Dict[str, Any]):
        """{"y": (N,T,H,W,C)} -> context latents (N,T,h,w,c): per-frame encode, posterior mode."""
        if self._cond_is_first_stage:
            c = c.get("y")
            N, T = c.shape[0], c.shape[1]
            frames = c.permute(0, 1, 4, 2, 3).reshape(N * T, c.shape[4], c.shape[2], c.shape[3])
            post = self.cond_stage_model.encode(frames)
            z = post.mode() if hasattr(post, "mode") else post
            return z.reshape(N, T, *z.shape[1:]).permute(0, 1, 3, 4, 2).contiguous()
        return c

    @torch.no_grad()
    def encode_first_stage(self, x):
        post = self.first_stage_model.encode(x)
        z = post.sample() if hasattr(post, "sample") else post
        return (self.scale_factor * z).detach()

    @torch.no_grad()
    def decode_first_stage(self, z, out=None):
        """(N,T,h,w,c) latents -> (N,T,H,W,C) pixels (latent_diffusion.py:423-432). `out` (optional): a contiguous fp32
        (N,T,H,W,1) buffer the CUDA decoder writes into directly (one pixel channel: 'NTHWC' and '(NT)CHW' are the same
        bytes) - used for the in-place ensemble all-gather (prediff_b200/dist.py)."""
        z = 1. / self.scale_factor * z
        N, T = z.shape[0], z.shape[1]
        frames = z.permute(0, 1, 4, 2, 3).reshape(N * T, z.shape[4], z.shape[2], z.shape[3])
        if out is not None:
            assert out.is_contiguous() and out.shape[0] == N and out.shape[1] == T and out.shape[4] == 1
            self.first_stage_model.decode(frames, out=out.view(N * T, 1, out.shape[2], out.shape[3]))
            return out
        out = self.first_stage_model.decode(frames)
        return out.reshape(N, T, *out.shape[1:]).permute(0, 1, 3, 4, 2).contiguous()

    # ---- denoiser --------------------------------------------------------------------------------------------
    def apply_model(self, x_noisy, t, cond):
        out = self.torch_nn_module(x_noisy, t, cond)
        return out[0] if isinstance(out, tuple) else out

    def _native(self):
        """True when the denoiser is the CUDA UNet, i.e. the loop can stay resident on the device."""
        from .unet import CuboidTransformerUNet
        return isinstance(self.torch_nn_module, CuboidTransformerUNet)

    def q_sample(self, x_start, t, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        return (self.extract_into_tensor(self.sqrt_alphas_cumprod.to(x_start.device), t, x_start.shape) * x_start +
                self.extract_into_tensor(self.sqrt_one_minus_alphas_cumprod.to(x_start.device), t, x_start.shape) * noise)

    # ---- forward (validation) loss: latent_diffusion.py:447-478, 494-551 ---------------------------------------
    @property
    def loss_mean_dim(self):
        return tuple(i for i in range(len(self.layout)) if i != self.batch_axis)

    def get_loss(self, pred, target, mean=True):
        """latent_diffusion.py:494-509 (plain torch on whatever device the tensors live on)."""
        loss = (target - pred).abs() if self.loss_type == "l1" else (target - pred) ** 2
        return loss.mean() if mean else loss

    def get_input(self, batch, **kwargs):
        """Dataset dependent (latent_diffusion.py:419-435): returns (target sequence, {"y": context sequence})."""
        return batch

    @torch.no_grad()
    def p_losses(self, x_start, cond, t, noise=None):
        """latent_diffusion.py:517-551, forward only: (loss, {prefix/loss_simple, prefix/loss_vlb, prefix/loss}).
        With the CUDA UNet the whole evaluation (q_sample, UNet, reductions) is one call into the library and the four
        scalars stay on the device; with any other denoiser module the reference's arithmetic runs in torch."""
        noise = torch.randn_like(x_start) if noise is None else noise
        prefix = "train" if self.training else "val"
        B = x_start.shape[0]
        if self._native() and x_start.is_cuda:
            x_start, cond, noise = (v.contiguous().float() for v in (x_start, cond, noise))
            t = t.to(device=x_start.device, dtype=torch.int64).contiguous()
            per_sample = torch.empty(B, device=x_start.device, dtype=torch.float32)
            out4 = torch.empty(4, device=x_start.device, dtype=torch.float32)
            with torch.cuda.device(x_start.device):
                L.check(L.lib().pd_diffusion_losses(
                    self._sampler, self.torch_nn_module.handle, L.ptr(x_start), L.ptr(cond), L.ptr(t), L.ptr(noise), B,
                    1 if self.loss_type == "l1" else 0, ctypes.c_float(self.logvar_init), ctypes.c_float(self.l_simple_weight),
                    ctypes.c_float(self.original_elbo_weight), L.ptr(per_sample), L.ptr(out4), L.stream_ptr()))
            self.last_loss_per_sample = per_sample
            self.last_loss_fused = out4   # the kernel's own scalars (computed with the scalar logvar_init)
            # the reference reads logvar[t] from the persistent buffer / parameter (a checkpoint may hold values that
            # differ from logvar_init), so the logged terms are B-element tensor arithmetic on the per-sample losses
            return self._loss_terms(per_sample, t, prefix)
        x_noisy = self.q_sample(x_start=x_start, t=t, noise=noise)
        model_output = self.apply_model(x_noisy, t, cond)
        loss_simple = self.get_loss(model_output, noise, mean=False).mean(dim=self.loss_mean_dim)
        return self._loss_terms(loss_simple, t, prefix)

    def _loss_terms(self, loss_simple, t, prefix):
        """latent_diffusion.py:534-549 from the per-sample loss: logvar weighting, vlb term, the logged dict."""
        logvar_t = self.logvar.detach().to(t.device)[t]
        loss = loss_simple / torch.exp(logvar_t) + logvar_t
        loss_dict = {f"{prefix}/loss_simple": loss_simple.mean()}
        if self.learn_logvar:
            loss_dict[f"{prefix}/loss_gamma"] = loss.mean()
            loss_dict["logvar"] = self.logvar.data.mean()
        loss = self.l_simple_weight * loss.mean()
        loss_vlb = (self.lvlb_weights.to(t.device)[t] * loss_simple).mean()
        loss_dict[f"{prefix}/loss_vlb"] = loss_vlb
        loss = loss + self.original_elbo_weight * loss_vlb
        loss_dict[f"{prefix}/loss"] = loss
        return loss, loss_dict

    @torch.no_grad()
    def forward(self, batch, verbose=False, t=None, noise=None):
        """latent_diffusion.py:447-478: target -> latents (posterior sample), t ~ U{0..T-1}, context -> latents,
        p_losses. `t` / `noise` may be injected (parity tests); by default they are drawn where the reference draws."""
        x, c = self.get_input(batch)
        B = x.shape[self.batch_axis]
        N, T = x.shape[0], x.shape[1]
        frames = x.permute(0, 1, 4, 2, 3).reshape(N * T, x.shape[4], x.shape[2], x.shape[3])
        z = self.encode_first_stage(frames)
        z = z.reshape(N, T, *z.shape[1:]).permute(0, 1, 3, 4, 2).contiguous()
        if t is None:
            t = torch.randint(0, self.num_timesteps, (B,), device=z.device).long()
        if self.shorten_cond_schedule and self.cond_stage_model is not None:
            # latent_diffusion.py:469-471 calls torch.randn_like(c.float()) on the context DICT: the reference raises here
            raise NotImplementedError("prediff_b200.LatentDiffusion.forward: shorten_cond_schedule (the reference's own "
                                      "forward fails with a dict context; only the sampling loop uses the option)")
        zc = self.cond_stage_forward(c) if self.cond_stage_model is not None else (c if torch.is_tensor(c) else c.get("y"))
        return self.p_losses(z, zc, t, noise=noise)

    def _native_alignment(self, use_alignment, alignment_kwargs):
        """(alignment object, avg_x_gt [B] device tensor factory) when the registered alignment_fn is the
        `get_mean_shift` of a prediff_b200 SEVIRAvgIntensityAlignment - then the guided loop stays on the device."""
        if not use_alignment:
            return None
        from .alignment import SEVIRAvgIntensityAlignment
        owner = getattr(self.alignment_fn, "__self__", None)
        if isinstance(owner, SEVIRAvgIntensityAlignment) and alignment_kwargs and "avg_x_gt" in alignment_kwargs \
                and getattr(self.alignment_fn, "__name__", "") == "get_mean_shift":
            return owner
        return None

    def _run_range(self, z, cond, noise, mode, n_total, eta, k0, k1, align=None, avg_x_gt=None):
        B = z.shape[0]
        with torch.cuda.device(z.device):
            if align is None:
                L.check(L.lib().pd_sample_loop_range(self._sampler, self.torch_nn_module.handle, L.ptr(z), L.ptr(cond),
                                                     L.ptr(noise), B, mode, n_total, ctypes.c_float(eta), k0, k1,
                                                     L.stream_ptr()))
            else:
                L.check(L.lib().pd_sample_loop_aligned(self._sampler, self.torch_nn_module.handle, align.model.handle,
                                                       L.ptr(z), L.ptr(cond), L.ptr(noise), L.ptr(avg_x_gt),
                                                       ctypes.c_float(align.guide_scale), B, mode, n_total,
                                                       ctypes.c_float(eta), k0, k1, L.stream_ptr()))

    @staticmethod
    def _target_vector(alignment_kwargs, B, device):
        t = torch.as_tensor(alignment_kwargs["avg_x_gt"], dtype=torch.float32, device=device).reshape(-1)
        if t.numel() == 1:
            t = t.expand(B)
        assert t.numel() == B, "avg_x_gt must hold one value per sample"
        return t.contiguous()

    def _coef_row(self, t_int: int, clip: bool, device):
        """The update kernel's coefficient row for ancestral timestep t (csrc/sampler_host.cu Sampler::coefficients):
        z0 = c0 z - c1 eps (clamped if c7); z <- c2 z0 + c3 z + c5 noise - c6 guide."""
        lv = float(self.posterior_log_variance_clipped[t_int])
        sigma = float(np.exp(np.float32(0.5) * np.float32(lv)))
        row = [float(self.sqrt_recip_alphas_cumprod[t_int]), float(self.sqrt_recipm1_alphas_cumprod[t_int]),
               float(self.posterior_mean_coef1[t_int]), float(self.posterior_mean_coef2[t_int]), 0.0,
               0.0 if t_int == 0 else sigma, sigma, 1.0 if clip else 0.0]
        return torch.tensor(row, dtype=torch.float32, device=device)

    @torch.no_grad()
    def p_sample(self, zt, zc, t, y=None, use_alignment=False, alignment_kwargs=None, clip_denoised=False,
                 return_x0=False, temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                 noise=None, _t_int=None):
        """One ancestral step (latent_diffusion.py:598-631). `noise` may be injected for parity tests; otherwise it
        is drawn with torch.randn exactly where the reference draws it.

        With the CUDA UNet the step runs through the C ABI: `pd_sample_step_ddpm` (UNet + fused update in one call), or -
        with a foreign `alignment_fn` / a score corrector - the UNet, the caller's function and then the fused update
        kernel (`pd_op_sampler_update`). Only `return_x0=True` and non-CUDA denoisers evaluate the reference's tensor
        expressions in torch."""
        B = zt.shape[0]
        dev = zt.device
        if noise is None:
            noise = torch.randn(zt.shape, device=dev)   # noise_like (diffusion/utils.py:118-119)
        noise = noise * temperature
        if noise_dropout > 0.:
            noise = torch.nn.functional.dropout(noise, p=noise_dropout)
        clip = bool(clip_denoised)   # the reference's p_sample uses its argument only; p_sample_loop passes self.clip_denoised
        native = self._native() and zt.is_cuda and not return_x0
        if native:
            ti = int(t[0]) if _t_int is None else int(_t_int)   # the reference's loop passes torch.full((B,), i)
            native = _t_int is not None or bool((t == ti).all())
        if native:
            z = zt.contiguous().float().clone()   # the caller's zt must stay intact (SURVEY 8b ownership)
            zc_, nz = zc.contiguous().float(), noise.contiguous().float()
            align = self._native_alignment(use_alignment, alignment_kwargs)
            lib = L.lib()
            with torch.cuda.device(dev):
                L.check(lib.pd_sampler_set_clip_denoised(self._sampler, 1 if clip else 0))
                try:
                    if score_corrector is None and (not use_alignment or align is not None):
                        if align is None:
                            L.check(lib.pd_sample_step_ddpm(self._sampler, self.torch_nn_module.handle, L.ptr(z), L.ptr(zc_),
                                                            L.ptr(nz), B, ti, L.stream_ptr()))
                        else:   # timestep ti = executed step 0 of a (ti + 1)-step ancestral schedule
                            target = self._target_vector(alignment_kwargs, B, dev)
                            self._run_range(z, zc_, nz[None].contiguous(), PD_MODE_DDPM, ti + 1, 0.0, 0, 1, align, target)
                    else:
                        eps = self.apply_model(zt, t, zc)
                        if score_corrector is not None:
                            eps = score_corrector.modify_score(self, eps, zt, t, zc, **(corrector_kwargs or {}))
                        g = None
                        if use_alignment:
                            g = self.alignment_fn(zt, t, zc=zc, y=y, **(alignment_kwargs or {})).contiguous().float()
                        coef = self._coef_row(ti, clip, dev)
                        eps = eps.contiguous().float()
                        L.check(lib.pd_op_sampler_update(L.ptr(z), L.ptr(eps), L.ptr(nz), L.ptr(g), L.ptr(coef),
                                                         ctypes.c_int64(z.numel()), L.stream_ptr()))
                finally:
                    L.check(lib.pd_sampler_set_clip_denoised(self._sampler, 1 if self.clip_denoised else 0))
            return z
        outputs = self.p_mean_variance(zt=zt, zc=zc, t=t, clip_denoised=clip, return_x0=return_x0,
                                       score_corrector=score_corrector, corrector_kwargs=corrector_kwargs)
        mean, _, logvar = outputs[:3]
        if use_alignment:
            mean = self.aligned_mean(zt=zt, t=t, zc=zc, y=y, orig_mean=mean, orig_log_var=logvar,
                                     **(alignment_kwargs or {}))
        nonzero = (1 - (t == 0).float()).reshape(B, *((1,) * (zt.dim() - 1)))
        out = mean + nonzero * (0.5 * logvar).exp() * noise
        return (out, outputs[3]) if return_x0 else out

    @torch.no_grad()
    def p_sample_loop(self, cond, shape, y=None, use_alignment=False, alignment_kwargs=None,
                      return_intermediates=False, x_T=None, verbose=False, callback=None, timesteps=None, mask=None,
                      x0=None, img_callback=None, start_T=None, log_every_t=None, noise=None, cond_noise=None):
        """DDPM ancestral loop, t = timesteps-1 .. 0 (latent_diffusion.py:633-684).

        With `shorten_cond_schedule` (num_timesteps_cond > 1) the context is re-noised before every step,
        cond <- q_sample(cond, cond_ids[t]) - cumulatively, as the reference does (:665-667); the draws are
        torch.randn_like(cond) in the reference's order, or `cond_noise` [steps, *cond.shape] (parity tests).

        With the CUDA UNet and no alignment / inpainting the stretches between logging points run as one
        device-resident loop; the per-step noise is pre-drawn with the same torch.randn calls, in the same order,
        that the reference makes (or taken from `noise` [steps, *shape])."""
        log_every_t = log_every_t or self.log_every_t
        device = cond.device
        B = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T.clone()
        img = img.contiguous().float()
        intermediates = [img.clone()]
        timesteps = self.num_timesteps if timesteps is None else timesteps
        if start_T is not None:
            timesteps = min(timesteps, start_T)
        if mask is not None:
            assert x0 is not None and x0.shape[2:3] == mask.shape[2:3]
        cond = cond.contiguous().float()
        align = self._native_alignment(use_alignment, alignment_kwargs) if self._native() else None
        host_driven = (use_alignment and align is None) or mask is not None or not self._native() or \
            self.shorten_cond_schedule
        target = self._target_vector(alignment_kwargs, B, device) if align is not None else None
        if host_driven:
            for k, i in enumerate(reversed(range(timesteps))):
                ts = torch.full((B,), i, device=device, dtype=torch.long)
                if self.shorten_cond_schedule:
                    tc = self.cond_ids.to(device)[ts]
                    cond = self.q_sample(x_start=cond, t=tc,
                                         noise=torch.randn_like(cond) if cond_noise is None else cond_noise[k])
                img = self.p_sample(zt=img, zc=cond, t=ts, y=y, use_alignment=use_alignment,
                                    alignment_kwargs=alignment_kwargs, clip_denoised=self.clip_denoised,
                                    noise=None if noise is None else noise[k], _t_int=i)
                if mask is not None:
                    img = self.q_sample(x0, ts) * mask + (1. - mask) * img
                if i % log_every_t == 0 or i == timesteps - 1:
                    intermediates.append(img.clone())
                if callback:
                    callback(i)
                if img_callback:
                    img_callback(img, i)
        else:
            # stop points: after every step the reference would log / call back on
            per_step_host = bool(callback or img_callback)
            stops = [k + 1 for k, i in enumerate(reversed(range(timesteps)))
                     if per_step_host or (return_intermediates and (i % log_every_t == 0 or i == timesteps - 1))]
            if not stops or stops[-1] != timesteps:
                stops.append(timesteps)
            chunk = max(1, (256 << 20) // max(1, img.numel() * 4))  # <= 256 MiB of noise resident at a time
            k = 0
            for stop in stops:
                while k < stop:
                    k1 = min(stop, k + chunk)
                    if noise is not None:
                        nz = noise[k:k1].contiguous().float()
                    else:  # the reference's per-step torch.randn(shape) calls, drawn up front in order
                        nz = torch.stack([torch.randn(shape, device=device) for _ in range(k1 - k)])
                    self._run_range(img, cond, nz, PD_MODE_DDPM, timesteps, 0.0, k, k1, align, target)
                    k = k1
                i = timesteps - stop  # timestep just executed
                if return_intermediates and (i % log_every_t == 0 or i == timesteps - 1):
                    intermediates.append(img.clone())
                if per_step_host:
                    if callback:
                        callback(i)
                    if img_callback:
                        img_callback(img, i)
        if return_intermediates:
            return img, intermediates
        return img

    @torch.no_grad()
    def ddim_sample_loop(self, cond, shape, x_T=None, ddim_steps=50, eta=0.0, noise=None, use_alignment=False,
                         alignment_kwargs=None):
        """DDIM over make_ddim_timesteps('uniform', ddim_steps, T); eta = 0 is
        deterministic. Runs entirely on the device (one CUDA-graph replay per step). With `use_alignment` the
        knowledge-alignment guidance enters as eps_hat = eps + sqrt(1 - a_t) * g (S6), g from the CUDA KA network."""
        if not self._native():
            raise L.PDError("ddim_sample_loop needs the prediff_b200 CuboidTransformerUNet as torch_nn_module")
        device = cond.device
        img = torch.randn(shape, device=device) if x_T is None else x_T.clone()
        img = img.contiguous().float()
        cond = cond.contiguous().float()
        if eta != 0.0 and noise is None:
            noise = torch.stack([torch.randn(shape, device=device) for _ in range(ddim_steps)])
        nz = None if noise is None else noise.contiguous().float()
        align = self._native_alignment(use_alignment, alignment_kwargs)
        if use_alignment and align is None:
            raise L.PDError("ddim_sample_loop(use_alignment=True) needs set_alignment(<prediff_b200 "
                            "SEVIRAvgIntensityAlignment>.get_mean_shift) and alignment_kwargs={'avg_x_gt': ...}")
        target = self._target_vector(alignment_kwargs, img.shape[0], device) if align is not None else None
        self._run_range(img, cond, nz, PD_MODE_DDIM, ddim_steps, float(eta), 0, ddim_steps, align, target)
        return img

    @torch.no_grad()
    def sample(self, cond, batch_size=16, use_alignment=False, alignment_kwargs=None, return_intermediates=False,
               x_T=None, verbose=False, timesteps=None, mask=None, x0=None, shape=None, return_decoded=True,
               sampler="ddpm", ddim_steps=50, ddim_eta=0.0, out=None, **kwargs):
        """latent_diffusion.py:686-724: encode the context, run the loop, decode. `sampler="ddim"` selects the
        DDIM loop (not in the reference; default stays the reference's ancestral sampler)."""
        if use_alignment:
            assert self.alignment_fn is not None, "Alignment function not set."
        if shape is None:
            shape = self.get_batch_latent_shape(batch_size=batch_size)
        if self.cond_stage_model is not None:
            assert cond is not None
            if isinstance(cond, dict):
                zc = {k: (v[:batch_size] if not isinstance(v, list) else [e[:batch_size] for e in v])
                      for k, v in cond.items()}
            else:
                zc = cond[:batch_size]
            zc = self.cond_stage_forward(zc if isinstance(zc, dict) else {"y": zc})
        else:
            zc = cond if isinstance(cond, torch.Tensor) else cond.get("y", None)
        y = cond if isinstance(cond, torch.Tensor) else cond.get("y", None)
        if sampler == "ddim":
            assert mask is None and not return_intermediates
            output = self.ddim_sample_loop(cond=zc, shape=shape, x_T=x_T, ddim_steps=ddim_steps, eta=ddim_eta,
                                           use_alignment=use_alignment, alignment_kwargs=alignment_kwargs)
        else:
            output = self.p_sample_loop(cond=zc, shape=shape, y=y, use_alignment=use_alignment,
                                        alignment_kwargs=alignment_kwargs, return_intermediates=return_intermediates,
                                        x_T=x_T, verbose=verbose, timesteps=timesteps, mask=mask, x0=x0, **kwargs)
        if return_decoded:
            if return_intermediates:
                samples, inter = output
                output = [self.decode_first_stage(samples), [self.decode_first_stage(e) for e in inter]]
            else:
                output = self.decode_first_stage(output) if out is None else self.decode_first_stage(output, out=out)
        return output
