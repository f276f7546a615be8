"""LitEma - exponential moving average of the denoiser's parameters, host-side mirror of the reference class
(src/prediff/utils/ema.py:6-76): same buffer names (`decay`, `num_updates`, parameter names with the dots removed), so
a checkpoint written by the reference's LatentDiffusion(use_ema=True) loads, and same update / store / copy_to / restore
semantics. The shipped config sets `use_ema: true` (scripts/prediff/sevirlr/cfg.yaml:83) and the inference script
evaluates under `ema_scope()` (train_sevirlr_prediff.py:815). Pure tensor plumbing: the CUDA model re-reads its
parameters (`refresh()`) after a swap.
"""
import torch
from torch import nn


class LitEma(nn.Module):
    def __init__(self, model, decay=0.9999, use_num_upates=True, include_frozen=False):
        """include_frozen: also shadow parameters with requires_grad == False. The reference tracks trainable parameters
        only; the CUDA denoiser mirrors keep all their parameters frozen (no backward exists), although every one of them
        is trainable in the reference - LatentDiffusion passes True for them so the reference's `model_ema.*` keys exist."""
        super().__init__()
        if decay < 0.0 or decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.m_name2s_name = {}
        self.register_buffer("decay", torch.tensor(decay, dtype=torch.float32))
        self.register_buffer("num_updates", torch.tensor(0 if use_num_upates else -1, dtype=torch.int))
        for name, p in model.named_parameters():
            if p.requires_grad or include_frozen:
                s_name = name.replace(".", "")   # '.' is not allowed in buffer names
                self.m_name2s_name[name] = s_name
                self.register_buffer(s_name, p.clone().detach().data)
        self.collected_params = []

    @torch.no_grad()
    def forward(self, model):
        """One EMA update: shadow -= (1 - decay_t) * (shadow - param), decay_t = min(decay, (1 + n) / (10 + n))."""
        decay = self.decay
        if self.num_updates >= 0:
            self.num_updates += 1
            decay = min(self.decay, (1 + self.num_updates) / (10 + self.num_updates))
        one_minus_decay = 1.0 - decay
        shadow = dict(self.named_buffers())
        for key, p in model.named_parameters():
            if key in self.m_name2s_name:
                s = shadow[self.m_name2s_name[key]]
                s.sub_(one_minus_decay * (s - p.to(s.dtype)))

    @torch.no_grad()
    def copy_to(self, model):
        shadow = dict(self.named_buffers())
        for key, p in model.named_parameters():
            if key in self.m_name2s_name:
                p.data.copy_(shadow[self.m_name2s_name[key]].data)

    def store(self, parameters):
        self.collected_params = [param.clone() for param in parameters]

    @torch.no_grad()
    def restore(self, parameters):
        for c_param, param in zip(self.collected_params, parameters):
            param.data.copy_(c_param.data)
