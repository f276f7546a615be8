"""Input side of the sampling path: raw SEVIR-LR VIL events -> the (B, T, H, W, 1) fp32 batches `LatentDiffusion.sample`
takes as `cond["y"]` / evaluation targets.

Mirrors the reference's `SEVIRDataLoader` in 'sequent' mode for the 'vil' data type
(src/prediff/datasets/sevir/sevir_dataloader.py:87-300 constructor arguments, :310-358 shard properties, :517-521 __len__,
:834-891 `_idx_sample`, :610-650 `preprocess_data_dict`): same window enumeration, same rescaling, same output layout and
bit-identical values - but the events cross PCIe as uint8 from pinned memory on a copy stream (4x fewer bytes than the
fp32 batch the reference moves) and are windowed / rescaled / transposed by one kernel (`pd_sevir_windows`, csrc/io.cu),
double-buffered so the copy of batch i+1 runs under the sampling of batch i.

`events` is any uint8 array of shape (num_events, H, W, raw_seq_len) - a numpy array, an `np.load(..., mmap_mode="r")`
memory map of an exported event file, an `h5py.Dataset` (`f["vil"]`), or a `SEVIRCatalogEvents`: the reference's catalog
layer (sevir_dataloader.py:212-300, 360-390 - catalog CSV, date / datetime / catalog filters, co-located image types,
duplicate-id removal, optional shuffle, one open file per HDF5 name) restated over pandas, which presents the selected
events, in the reference's order, behind the same indexing interface. `SEVIRDataLoader.from_catalog(...)` takes the
reference constructor's catalog arguments. Files are opened with `h5py.File(path, "r")`; h5py is absent from this image, so
the opener is injectable (`open_file=`: any callable path -> mapping with a "vil" dataset) and a missing h5py without an
opener is an error at construction, not a fallback.
"""
import numpy as np
import torch

from . import _lib as L

PREPROCESS_SCALE = {"sevir": 1 / 47.54, "01": 1 / 255}      # sevir_dataloader.py:25-44, 'vil' entries
PREPROCESS_OFFSET = {"sevir": -33.44, "01": 0.0}


def _default_open_file(path):
    try:
        import h5py
    except ImportError as e:
        raise L.PDError("prediff_b200.data.SEVIRCatalogEvents: h5py is not importable - pass open_file= (a callable "
                        "path -> mapping with a 'vil' dataset of shape (n, H, W, raw_seq_len) uint8)") from e
    return h5py.File(path, "r")


class SEVIRCatalogEvents:
    """The catalog layer of the reference's SEVIRDataLoader for the 'vil' data type (sevir_dataloader.py:212-300):
    `catalog` (CSV path or DataFrame) filtered by start_date < time_utc <= end_date (:239-242), `datetime_filter` (:243-244)
    and `catalog_filter` ('default' = pct_missing == 0, :246-249); `_compute_samples` (:256-271): rows of the requested image
    types, ids that have every requested type exactly once (repeated ids are dropped entirely, as the reference does),
    one (file name, file index) per id in the reference's order (ids sorted by `groupby`, or `DataFrame.sample(frac=1,
    random_state=shuffle_seed)` when shuffled); `_open_files` (:287-299): one handle per distinct file name.
    Indexing (`events[e0:e1]`, `events[i]`) = `_read_data` (:360-390): `file["vil"][index:index + 1]` per event,
    concatenated -> uint8 (n, H, W, raw_seq_len)."""

    dtype = np.dtype(np.uint8)

    def __init__(self, sevir_catalog, sevir_data_dir, data_types=("vil",), start_date=None, end_date=None, datetime_filter=None,
                 catalog_filter="default", shuffle=False, shuffle_seed=1, open_file=None, verbose=False):
        import pandas as pd
        if "vil" not in tuple(data_types):
            raise NotImplementedError("prediff_b200.data.SEVIRCatalogEvents: the PreDiff path reads the 'vil' data type")
        if "lght" in tuple(data_types):
            raise NotImplementedError("prediff_b200.data.SEVIRCatalogEvents: lightning ('lght') gridding is not on the path")
        self.data_types = list(data_types)
        cat = pd.read_csv(sevir_catalog, parse_dates=["time_utc"], low_memory=False) if isinstance(sevir_catalog, str) \
            else sevir_catalog
        if start_date is not None:
            cat = cat[cat.time_utc > start_date]
        if end_date is not None:
            cat = cat[cat.time_utc <= end_date]
        if datetime_filter:
            cat = cat[datetime_filter(cat.time_utc)]
        if catalog_filter is not None:
            if isinstance(catalog_filter, str) and catalog_filter == "default":
                catalog_filter = lambda c: c.pct_missing == 0   # noqa: E731
            cat = cat[catalog_filter(cat)]
        self.catalog = cat
        self.sevir_data_dir = sevir_data_dir
        self.shuffle, self.shuffle_seed = bool(shuffle), int(shuffle_seed)
        self._compute_samples()
        self.reset()   # the reference constructor ends with reset(): a shuffled loader starts from the SECOND permutation
        opener = open_file or _default_open_file
        self._files = {}
        for f in np.unique(self._samples["vil_filename"].values):   # _open_files
            if verbose:
                print("Opening HDF5 file for reading", f)
            self._files[f] = opener(f"{self.sevir_data_dir}/{f}")
        first = self._files[self._samples["vil_filename"].iloc[0]]["vil"] if len(self._samples) else None
        self._event_shape = tuple(int(v) for v in first.shape[1:]) if first is not None else (0, 0, 0)

    def _compute_samples(self):
        import pandas as pd
        imgt = self.data_types
        cat = self.catalog
        filt = cat[np.logical_or.reduce([cat.img_type == i for i in imgt])]
        n_types = filt.groupby("id")["img_type"].nunique()
        n_rows = filt.groupby("id").size()
        keep = n_rows.index[(n_types == len(imgt)) & (n_rows == len(imgt))]   # every requested type present, no repeated id
        filt = filt[filt.id.isin(keep)]
        rows = filt[filt.img_type == "vil"].sort_values("id", kind="stable")   # groupby('id') yields ids in sorted order
        self._samples = pd.DataFrame({"id": rows.id.values, "vil_filename": rows.file_name.values,
                                      "vil_index": rows.file_index.values.astype(np.int64)})
        if self.shuffle:
            self.shuffle_samples()

    def shuffle_samples(self):
        self._samples = self._samples.sample(frac=1, random_state=self.shuffle_seed)

    def reset(self, shuffle=None):
        """Start of an epoch (sevir_dataloader.py:508-515): reshuffles the CURRENT order with the same seed."""
        if self.shuffle if shuffle is None else shuffle:
            self.shuffle_samples()

    def close(self):
        for f in self._files.values():
            if hasattr(f, "close"):
                f.close()
        self._files = {}

    @property
    def shape(self):
        return (len(self._samples),) + self._event_shape

    def __len__(self):
        return len(self._samples)

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            k = int(key) + (len(self) if int(key) < 0 else 0)
            if not 0 <= k < len(self):
                raise IndexError(f"event {int(key)} out of range for {len(self)} events")
            return self[k:k + 1][0]
        if not isinstance(key, slice):
            raise TypeError("SEVIRCatalogEvents is indexed by an event number or a slice of event numbers")
        rows = self._samples.iloc[key]
        out = np.empty((len(rows),) + self._event_shape, np.uint8)
        for k, (fname, idx) in enumerate(zip(rows["vil_filename"].values, rows["vil_index"].values)):
            out[k] = self._files[fname]["vil"][int(idx):int(idx) + 1][0]
        return out


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

    @classmethod
    def from_catalog(cls, sevir_catalog, sevir_data_dir, data_types=("vil",), start_date=None, end_date=None,
                     datetime_filter=None, catalog_filter="default", shuffle=False, shuffle_seed=1, open_file=None,
                     verbose=False, **loader_kwargs):
        """The reference constructor's catalog arguments (sevir_dataloader.py:99-122) -> a loader over the selected events."""
        events = SEVIRCatalogEvents(sevir_catalog, sevir_data_dir, data_types=data_types, start_date=start_date, end_date=end_date,
                                    datetime_filter=datetime_filter, catalog_filter=catalog_filter, shuffle=shuffle,
                                    shuffle_seed=shuffle_seed, open_file=open_file, verbose=verbose)
        return cls(events, data_types=("vil",), **loader_kwargs)

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


class SEVIRTorchDataset(torch.utils.data.Dataset):
    """Mirror of the reference's torch Dataset wrapper (src/prediff/datasets/sevir/sevir_torch_wrap.py:73-163) for
    aug_mode "0" (evaluation): item `index` = sequence `index` of the sequent enumeration as a (T, H, W, 1) fp32 tensor in
    `layout` - here a device tensor produced by the window kernel (the reference returns a host tensor that the Lightning
    loop then copies to the GPU). Same constructor argument names; the HDF5 opener is injectable like SEVIRCatalogEvents'."""

    def __init__(self, seq_len=25, raw_seq_len=49, sample_mode="sequent", stride=12, layout="THWC", split_mode="uneven",
                 sevir_catalog=None, sevir_data_dir=None, start_date=None, end_date=None, datetime_filter=None,
                 catalog_filter="default", shuffle=False, shuffle_seed=1, output_type=np.float32, preprocess=True,
                 rescale_method="01", verbose=False, aug_mode="0", ret_contiguous=True, open_file=None, device=None):
        super().__init__()
        if aug_mode != "0":
            raise NotImplementedError("prediff_b200.data.SEVIRTorchDataset: augmentation (aug_mode '1' / '2') is a training feature")
        if output_type not in (np.float32, torch.float32):
            raise NotImplementedError("prediff_b200.data.SEVIRTorchDataset: fp32 output only")
        if sorted(layout) != sorted("THWC"):
            raise ValueError(f"layout {layout!r} must be a permutation of 'THWC'")
        self.layout, self.ret_contiguous = layout, ret_contiguous
        self.sevir_dataloader = SEVIRDataLoader.from_catalog(
            sevir_catalog, sevir_data_dir, start_date=start_date, end_date=end_date, datetime_filter=datetime_filter,
            catalog_filter=catalog_filter, shuffle=shuffle, shuffle_seed=shuffle_seed, open_file=open_file, verbose=verbose,
            seq_len=seq_len, raw_seq_len=raw_seq_len, sample_mode=sample_mode, stride=stride, batch_size=1, layout="NTHWC",
            num_shard=1, rank=0, split_mode=split_mode, preprocess=preprocess, rescale_method=rescale_method, device=device)

    def __getitem__(self, index):
        data = self.sevir_dataloader._idx_sample(index=index)["vil"].squeeze(0)        # (T, H, W, 1)
        data = data.permute(*["THWC".index(a) for a in self.layout])
        return data.contiguous() if self.ret_contiguous else data

    def __len__(self):
        return len(self.sevir_dataloader)
