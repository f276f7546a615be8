"""SEVIRSkillScore - host-side mirror of the reference's evaluation metric
(src/prediff/datasets/sevir/evaluation.py:88-285) over the on-device evaluation kernel (`pd_sevir_eval_update`).

Same constructor arguments, `update(pred, target)` / `compute()` / `reset()` contract and result dictionary as the
reference's torchmetrics `Metric`, so `self.test_score.update(pred_seq, target_seq)` (train_sevirlr_prediff.py:962) takes
the decoded frames straight from `LatentDiffusion.sample()` without the `.cpu()` round trip. The kernel also accumulates
the squared / absolute error sums of the torchmetrics MeanSquaredError / MeanAbsoluteError that run beside it (:960-961):
`mse()` / `mae()`. States are summed across ranks in `compute()` (the reference's `dist_reduce_fx="sum"`) when
torch.distributed is initialised. There is no CPU fallback."""
import ctypes
import re
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib as L


class SEVIRSkillScore:

    def __init__(self, layout: str = "NHWT", mode: str = "0", seq_len: Optional[int] = None,
                 preprocess_type: str = "sevir", threshold_list: Sequence[int] = (16, 74, 133, 160, 181, 219),
                 metrics_list: Sequence[str] = ("csi", "bias", "sucr", "pod"), eps: float = 1e-4):
        assert preprocess_type == "sevir" or preprocess_type.startswith("sevir_pool")
        if mode not in ("0", "1", "2"):
            raise NotImplementedError(f"mode {mode} not supported!")
        if mode in ("1", "2"):
            assert isinstance(seq_len, int), "seq_len must be provided when we need to keep seq_len dim."
        if len(threshold_list) > 8:
            raise NotImplementedError("prediff_b200.SEVIRSkillScore: at most 8 thresholds")
        self.layout, self.mode, self.seq_len = layout, mode, seq_len
        self.preprocess_type = preprocess_type
        self.pool_scale = 1 if preprocess_type == "sevir" else int(re.search(r"\d+", preprocess_type).group())
        self.threshold_list, self.metrics_list, self.eps = list(threshold_list), list(metrics_list), eps
        self.keep_seq_len_dim = mode in ("1", "2")
        self._thr = (ctypes.c_float * len(self.threshold_list))(*[float(t) for t in self.threshold_list])
        self._counts = None   # int64 [n_thr][T][3] on the device of the first update
        self._sums = None     # double [2]
        self._numel = 0

    def reset(self):
        self._counts, self._sums, self._numel = None, None, 0

    def _to_nthw(self, x):
        """Any permutation of N, T, H, W (+ optional C of size 1) -> contiguous fp32 (N, T, H, W)."""
        lay = self.layout
        if "C" in lay:
            assert x.shape[lay.find("C")] == 1, "single-channel frames expected"
            x = x.squeeze(lay.find("C"))
            lay = lay.replace("C", "")
        return x.permute(*[lay.find(a) for a in "NTHW"]).contiguous().float()

    @torch.no_grad()
    def update(self, pred: torch.Tensor, target: torch.Tensor):
        if not pred.is_cuda:
            raise L.PDError("prediff_b200.SEVIRSkillScore runs on a CUDA (sm_100) device only; got a CPU tensor")
        p, t = self._to_nthw(pred.detach()), self._to_nthw(target.detach().to(pred.device))
        assert p.shape == t.shape
        N, T, H, W = p.shape
        if self.keep_seq_len_dim:
            assert T == self.seq_len, f"seq_len {self.seq_len} != T {T}"
        if self._counts is None:
            self._counts = torch.zeros(len(self.threshold_list), T, 3, dtype=torch.int64, device=p.device)
            self._sums = torch.zeros(2, dtype=torch.float64, device=p.device)
        assert self._counts.shape[1] == T
        with torch.cuda.device(p.device):
            L.check(L.lib().pd_sevir_eval_update(L.ptr(p), L.ptr(t), L.ptr(self._counts), L.ptr(self._sums), N, T, H, W,
                                                 self.pool_scale, self._thr, len(self.threshold_list), L.stream_ptr()))
        self._numel += p.numel()

    def _reduced_state(self):
        assert self._counts is not None, "compute() before any update()"
        counts, sums, numel = self._counts.clone(), self._sums.clone(), self._numel
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            n = torch.tensor([numel], dtype=torch.int64, device=counts.device)
            for t in (counts, sums, n):
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
            numel = int(n.item())
        return counts, sums, numel

    @property
    def hits_misses_fas(self):
        """int64 (n_thr, T, 3) numpy array of the accumulated (rank-local) state."""
        return self._counts.cpu().numpy()

    def mse(self):
        _, sums, numel = self._reduced_state()
        return sums[0].item() / max(numel, 1)

    def mae(self):
        _, sums, numel = self._reduced_state()
        return sums[1].item() / max(numel, 1)

    def compute(self):
        """evaluation.py:247-285 (fp32 state arithmetic like the reference's float states)."""
        counts, _, _ = self._reduced_state()
        c = counts.cpu().numpy().astype(np.float32)
        if not self.keep_seq_len_dim:
            c = c.sum(axis=1)
        hits, misses, fas = c[..., 0], c[..., 1], c[..., 2]
        eps = np.float32(self.eps)
        fn = {"pod": lambda: hits / (hits + misses + eps), "sucr": lambda: hits / (hits + fas + eps),
              "csi": lambda: hits / (hits + misses + fas + eps),
              "bias": lambda: np.power(((hits + fas) / (hits + misses + eps)) / np.log(np.float32(2.0)), np.float32(2.0))}
        ret = {th: {} for th in self.threshold_list}
        ret["avg"] = {}
        for m in self.metrics_list:
            sc = fn[m]().astype(np.float32)
            score_avg = np.zeros((self.seq_len,)) if self.keep_seq_len_dim else 0
            for i, th in enumerate(self.threshold_list):
                score = sc[i] if self.keep_seq_len_dim else sc[i].item()
                ret[th][m] = score if self.mode in ("0", "1") else np.mean(score).item()
                score_avg = score_avg + score
            score_avg = score_avg / len(self.threshold_list)
            ret["avg"][m] = score_avg if self.mode in ("0", "1") else np.mean(score_avg).item()
        return ret


class StructuralSimilarityIndexMeasure:
    """Mirror of `torchmetrics.image.StructuralSimilarityIndexMeasure()` as the inference script uses it
    (train_sevirlr_prediff.py:229-230: default arguments; :964-965: `self.test_ssim(pred_seq_bchw, target_seq_bchw)` with
    "(b t) c h w" frames) over the on-device kernel `pd_ssim_update` (csrc/eval.cu): 11 x 11 Gaussian window (sigma 1.5),
    k1 = 0.01, k2 = 0.03, `data_range=None` (taken from each update's batch), per-image mean over the interior, reduction
    'elementwise_mean'. One pixel channel (SEVIR VIL frames). The state (sum of per-image SSIM, image count) stays on the
    device and is summed across ranks in `compute()` like the reference's `dist_reduce_fx="sum"` states."""

    def __init__(self, gaussian_kernel=True, sigma=1.5, kernel_size=11, reduction="elementwise_mean", data_range=None,
                 k1=0.01, k2=0.03, **unused):
        if not gaussian_kernel or sigma != 1.5 or kernel_size != 11 or reduction != "elementwise_mean" or (k1, k2) != (0.01, 0.03):
            raise NotImplementedError("prediff_b200.StructuralSimilarityIndexMeasure: only the torchmetrics defaults are built")
        if isinstance(data_range, (tuple, list)):
            raise NotImplementedError("prediff_b200.StructuralSimilarityIndexMeasure: clamping data_range tuples are not built")
        self.data_range = data_range
        self._state = None

    def reset(self):
        self._state = None

    @torch.no_grad()
    def update(self, preds: torch.Tensor, target: torch.Tensor):
        if not preds.is_cuda:
            raise L.PDError("prediff_b200.StructuralSimilarityIndexMeasure runs on a CUDA (sm_100) device only; got a CPU tensor")
        assert preds.dim() == 4 and preds.shape[1] == 1 and preds.shape == target.shape, "expected (B, 1, H, W) frames"
        p = preds.detach().contiguous().float()
        t = target.detach().to(p.device).contiguous().float()
        if self._state is None:
            self._state = torch.zeros(2, dtype=torch.float64, device=p.device)
        with torch.cuda.device(p.device):
            L.check(L.lib().pd_ssim_update(L.ptr(p), L.ptr(t), p.shape[0], p.shape[2], p.shape[3],
                                           ctypes.c_float(-1.0 if self.data_range is None else float(self.data_range)),
                                           L.ptr(self._state), L.stream_ptr()))

    __call__ = update   # the script calls the metric object (forward = update + batch value); only the state is kept here

    def compute(self):
        assert self._state is not None, "compute() before any update()"
        st = self._state.clone()
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.all_reduce(st, op=torch.distributed.ReduceOp.SUM)
        return (st[0] / st[1]).float()
