"""Input side of the sampling path: raw SEVIR-LR VIL events -> the (B, T, H, W, 1) fp32 batches `LatentDiffusion.sample`
takes as `cond["y"]` / evaluation targets.

Mirrors the reference's `SEVIRDataLoader` in 'sequent' mode for the 'vil' data type
(src/prediff/datasets/sevir/sevir_dataloader.py:87-300 constructor arguments, :310-358 shard properties, :517-521 __len__,
:834-891 `_idx_sample`, :610-650 `preprocess_data_dict`): same window enumeration, same rescaling, same output layout and
bit-identical values - but the events cross PCIe as uint8 from pinned memory on a copy stream (4x fewer bytes than the
fp32 batch the reference moves) and are windowed / rescaled / transposed by one kernel (`pd_sevir_windows`, csrc/io.cu),
double-buffered so the copy of batch i+1 runs under the sampling of batch i.

The HDF5 / catalog layer of the reference (h5py, pandas filters, `_load_event_batch`) is not rebuilt: `events` is any
uint8 array of shape (num_events, H, W, raw_seq_len) - a numpy array, an `np.load(..., mmap_mode="r")` memory map of an
exported event file, or an `h5py.Dataset` (`f["vil"]`), which has the same indexing interface.
"""
import numpy as np
import torch

from . import _lib as L

PREPROCESS_SCALE = {"sevir": 1 / 47.54, "01": 1 / 255}      # sevir_dataloader.py:25-44, 'vil' entries
PREPROCESS_OFFSET = {"sevir": -33.44, "01": 0.0}


class SEVIRDataLoader:
    def __init__(self, events, seq_len=13, raw_seq_len=None, sample_mode="sequent", stride=6, batch_size=4,
                 layout="NTHWC", num_shard=1, rank=0, split_mode="uneven", preprocess=True, rescale_method="01",
                 data_types=("vil",), device=None, prefetch=2):
        if tuple(data_types) != ("vil",):
            raise NotImplementedError("prediff_b200.data.SEVIRDataLoader: only the 'vil' data type is on the PreDiff path")
        if sample_mode != "sequent":
            raise NotImplementedError("prediff_b200.data.SEVIRDataLoader: only sample_mode='sequent' (evaluation order)")
        if layout not in ("NTHWC", "NTHW"):
            raise NotImplementedError(f"layout {layout!r}: the sampling path uses 'NTHWC' (cfg.yaml:12)")
        if not preprocess or rescale_method not in PREPROCESS_SCALE:
            raise NotImplementedError("preprocess=False / unknown rescale_method")
        if split_mode not in ("ceil", "floor", "uneven"):
            raise ValueError(f"Invalid split_mode: {split_mode}")
        if len(events.shape) != 4 or events.dtype != np.uint8:
            raise ValueError("events must be uint8 (num_events, H, W, raw_seq_len) - the raw SEVIR 'NHWT' layout")
        self.events = events
        self.total_num_event, self.H, self.W = int(events.shape[0]), int(events.shape[1]), int(events.shape[2])
        self.raw_seq_len = int(events.shape[3]) if raw_seq_len is None else int(raw_seq_len)
        if self.raw_seq_len != events.shape[3]:
            raise ValueError("raw_seq_len does not match the event array")
        self.seq_len, self.stride, self.batch_size = int(seq_len), int(stride), int(batch_size)
        assert self.seq_len <= self.raw_seq_len and self.stride > 0 and self.batch_size > 0
        self.layout, self.rescale_method = layout, rescale_method
        self.num_shard, self.rank, self.split_mode = int(num_shard), int(rank), split_mode
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.prefetch = max(1, int(prefetch))
        self._copy_stream = None
        self._slots = None

    # ---- the reference's bookkeeping (sevir_dataloader.py:310-358, :517-521) ----
    @property
    def num_seq_per_event(self):
        return 1 + (self.raw_seq_len - self.seq_len) // self.stride

    @property
    def start_event_idx(self):
        return self.total_num_event // self.num_shard * self.rank

    @property
    def end_event_idx(self):
        if self.split_mode == "ceil":
            last_start = self.total_num_event // self.num_shard * (self.num_shard - 1)
            return self.start_event_idx + self.total_num_event - last_start
        if self.split_mode == "floor" or self.rank != self.num_shard - 1:
            return self.total_num_event // self.num_shard * (self.rank + 1)
        return self.total_num_event

    @property
    def num_event(self):
        return self.end_event_idx - self.start_event_idx

    @property
    def total_num_seq(self):
        return int(self.num_seq_per_event * self.num_event)

    def __len__(self):
        return self.total_num_seq // self.batch_size

    # ---- staging ----
    def _max_events_per_batch(self):
        n = self.num_seq_per_event
        return (self.batch_size + n - 2) // n + 1

    def _make_slots(self):
        if self._slots is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            ev_shape = (self._max_events_per_batch(), self.H, self.W, self.raw_seq_len)
            out_shape = (self.batch_size, self.seq_len, self.H, self.W) + ((1,) if self.layout == "NTHWC" else ())
            self._slots = [dict(host=torch.empty(ev_shape, dtype=torch.uint8).pin_memory(),
                                dev=torch.empty(ev_shape, dtype=torch.uint8, device=self.device),
                                out=torch.empty(out_shape, dtype=torch.float32, device=self.device),
                                ready=torch.cuda.Event(), free=torch.cuda.Event())
                           for _ in range(self.prefetch + 1)]
        return self._slots

    def _issue(self, slot, first_seq):
        """Pinned staging + H2D + window kernel for sequences [first_seq, first_seq + batch_size) on the copy stream."""
        n = self.num_seq_per_event
        e0, e1 = first_seq // n, (first_seq + self.batch_size - 1) // n
        if e1 >= self.total_num_event:
            raise IndexError(f"sequences [{first_seq}, {first_seq + self.batch_size}) run past the last event")
        ne = e1 - e0 + 1
        if slot.get("h2d_bytes"):
            slot["ready"].synchronize()   # the previous copy out of this pinned buffer has really happened
        slot["host"][:ne].numpy()[...] = np.asarray(self.events[e0:e1 + 1])   # the HDF5 / memory-map read
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(slot["free"])           # the consumer is done with this slot's previous batch
            slot["dev"][:ne].copy_(slot["host"][:ne], non_blocking=True)
            L.check(L.lib().pd_sevir_windows(L.ptr(slot["dev"]), e0, ne, self.H, self.W, self.raw_seq_len,
                                             L.c_i64(first_seq), self.batch_size, self.seq_len, self.stride,
                                             L.c_float(np.float32(PREPROCESS_SCALE[self.rescale_method])),
                                             L.c_float(np.float32(PREPROCESS_OFFSET[self.rescale_method])),
                                             L.ptr(slot["out"]), L.stream_ptr()))
            slot["ready"].record(self._copy_stream)
        slot["h2d_bytes"] = ne * self.H * self.W * self.raw_seq_len

    def _idx_sample(self, index):
        """Batch `index` of the sequent enumeration, as the reference returns it: {'vil': (B, T, H, W, 1) fp32} - here a
        device tensor (sevir_dataloader.py:834-891). Indices count from event 0 of the array, like the reference's."""
        slot = self._make_slots()[0]
        self._issue(slot, int(index) * self.batch_size)
        torch.cuda.current_stream(self.device).wait_event(slot["ready"])
        out = slot["out"].clone()
        slot["free"].record(torch.cuda.current_stream(self.device))
        return {"vil": out}

    def __iter__(self):
        """Batches of this rank's shard in order, prefetched: while the caller works on batch i, batches i+1 .. i+prefetch
        are being read, copied and preprocessed on the copy stream. The yielded tensor is only valid until the next
        iteration step (its buffer is recycled) - clone it to keep it."""
        slots = self._make_slots()
        first = self.start_event_idx * self.num_seq_per_event
        nb = len(self)
        cur = torch.cuda.current_stream(self.device)
        for s in slots:
            s["free"].record(cur)
        for i in range(min(self.prefetch, nb)):
            self._issue(slots[i % len(slots)], first + i * self.batch_size)
        for i in range(nb):
            slot = slots[i % len(slots)]
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(slot["ready"])
            yield slot["out"]
            slot["free"].record(torch.cuda.current_stream(self.device))
            j = i + self.prefetch
            if j < nb:
                self._issue(slots[j % len(slots)], first + j * self.batch_size)
