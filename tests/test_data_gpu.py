"""Input side: prediff_b200.data.SEVIRDataLoader / pd_sevir_windows against the CPU oracle and the
goldens of the unmodified reference SEVIRDataLoader._idx_sample - bit-exact (integer -> fp32 scale, no tolerance)."""
import os

import numpy as np
import pytest
import torch

from oracle import data_oracle as DO

GOLD = os.path.join(os.path.dirname(__file__), "golden", "loader.npz")
CASES = {"lr": (4, 13, 6, "01", "NTHWC"), "b3": (3, 13, 6, "01", "NTHWC"), "sevir": (2, 10, 5, "sevir", "NTHWC"),
         "nthw": (2, 13, 12, "01", "NTHW")}


def golden_events():
    rng = np.random.Generator(np.random.PCG64(7171))   # tests/golden/gen_golden.py::loader_events
    return rng.integers(0, 256, size=(5, 16, 24, 25), dtype=np.uint8)


@pytest.mark.parametrize("tag", sorted(CASES))
def test_oracle_matches_reference_batches(tag):
    """CPU: the numpy restatement reproduces every batch the reference's _idx_sample produced (tests/golden/loader.npz)."""
    bs, seq_len, stride, rescale, layout = CASES[tag]
    g = np.load(GOLD)
    ev = golden_events()
    assert int(g[f"{tag}_n"]) == (ev.shape[0] * DO.num_seq_per_event(25, seq_len, stride)) // bs
    for i in range(int(g[f"{tag}_n"])):
        o = DO.idx_sample(ev, i, bs, seq_len, stride, rescale, layout)
        assert o.dtype == np.float32 and o.shape == g[f"{tag}_{i}"].shape
        assert np.array_equal(o, g[f"{tag}_{i}"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(CASES))
def test_loader_matches_reference_batches(tag):
    from prediff_b200.data import SEVIRDataLoader
    bs, seq_len, stride, rescale, layout = CASES[tag]
    g = np.load(GOLD)
    ev = golden_events()
    dl = SEVIRDataLoader(ev, seq_len=seq_len, stride=stride, batch_size=bs, layout=layout, rescale_method=rescale)
    assert len(dl) == int(g[f"{tag}_n"])
    for i in range(len(dl)):
        out = dl._idx_sample(i)["vil"]
        assert out.is_cuda and out.dtype == torch.float32
        assert np.array_equal(out.cpu().numpy(), g[f"{tag}_{i}"])
    got = [b.clone() for b in dl]          # prefetching iterator: same batches, same order
    assert len(got) == len(dl)
    for i, b in enumerate(got):
        assert np.array_equal(b.cpu().numpy(), g[f"{tag}_{i}"])


@pytest.mark.gpu
def test_loader_full_size_and_shards():
    """SEVIR-LR shapes (128 x 128 x 25 events, 13-frame windows, stride 6), two shards, prefetch under work."""
    from prediff_b200.data import SEVIRDataLoader
    rng = np.random.Generator(np.random.PCG64(99))
    ev = rng.integers(0, 256, size=(11, 128, 128, 25), dtype=np.uint8)
    seen = 0
    for rank in (0, 1):
        dl = SEVIRDataLoader(ev, seq_len=13, stride=6, batch_size=4, num_shard=2, rank=rank, prefetch=2)
        assert dl.num_seq_per_event == 3
        assert (dl.start_event_idx, dl.end_event_idx) == ((0, 5) if rank == 0 else (5, 11))   # 'uneven' split
        first = dl.start_event_idx * 3
        for i, batch in enumerate(dl):
            busy = torch.randn(2048, 2048, device="cuda") @ torch.randn(2048, 2048, device="cuda")   # consumer work
            want = np.concatenate([
                DO.idx_sample(ev, (first + i * 4 + b), 1, 13, 6, "01", "NTHWC") for b in range(4)], 0)
            assert batch.shape == (4, 13, 128, 128, 1)
            assert np.array_equal(batch.cpu().numpy(), want)
            assert torch.isfinite(busy).all()
            seen += 1
        assert i + 1 == len(dl) == dl.total_num_seq // 4
    assert seen == (5 * 3) // 4 + (6 * 3) // 4


@pytest.mark.gpu
def test_windows_abi_rejects_out_of_range():
    from prediff_b200 import _lib as L
    ev = torch.zeros(2, 8, 8, 25, dtype=torch.uint8, device="cuda")
    out = torch.empty(4, 13, 8, 8, device="cuda")
    # 3 windows per event: sequences 4..7 need events 1..2, the buffer holds 0..1
    rc = L.lib().pd_sevir_windows(L.ptr(ev), 0, 2, 8, 8, 25, L.c_i64(4), 4, 13, 6, L.c_float(1 / 255), L.c_float(0.0),
                                  L.ptr(out), L.stream_ptr())
    assert rc != 0 and b"events" in L.lib().pd_last_error()
    with pytest.raises(NotImplementedError):
        from prediff_b200.data import SEVIRDataLoader
        SEVIRDataLoader(ev.cpu().numpy(), sample_mode="random")


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["default", "dates", "filters", "nofilter", "shuffle", "colocated"])
def test_loader_from_catalog_matches_reference_loader(tag):
    """SEVIRDataLoader.from_catalog (catalog filters -> event order -> per-event reads -> pinned staging -> window kernel)
    against the batches of the unmodified reference SEVIRDataLoader built from the same synthetic catalog
    (tests/golden/catalog.npz) - bit-exact."""
    from prediff_b200.data import SEVIRDataLoader
    from tests.golden import catalog_cases as CC
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "catalog.npz"))
    files = CC.catalog_files()
    ckw, lkw = CC.CASES[tag]
    dl = SEVIRDataLoader.from_catalog(CC.catalog_frame(), "/data", open_file=lambda p: files[p[len("/data/"):]], **ckw, **lkw)
    assert len(dl) == int(g[f"{tag}_len"])
    for i in range(len(dl)):
        assert np.array_equal(dl._idx_sample(i)["vil"].cpu().numpy(), g[f"{tag}_b{i}"])
    for i, b in enumerate(dl):   # the prefetching iterator reads through the catalog layer too
        assert np.array_equal(b.cpu().numpy(), g[f"{tag}_b{i}"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag,layout", [("thwc", "THWC"), ("cthw", "CTHW")])
def test_torch_dataset_wrapper_matches_reference(tag, layout):
    """SEVIRTorchDataset (sevir_torch_wrap.py:73-163, aug_mode "0") item by item against the unmodified reference class."""
    import datetime
    from prediff_b200.data import SEVIRTorchDataset
    from tests.golden import catalog_cases as CC
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "catalog.npz"))
    files = CC.catalog_files()
    ds = SEVIRTorchDataset(seq_len=13, raw_seq_len=CC.T_RAW, stride=6, layout=layout, sevir_catalog=CC.catalog_frame(),
                           sevir_data_dir="/data", rescale_method="01", start_date=datetime.datetime(2019, 2, 1),
                           open_file=lambda p: files[p[len("/data/"):]])
    assert len(ds) == int(g[f"ds_{tag}_len"])
    for i in range(len(ds)):
        item = ds[i]
        assert item.is_cuda and item.is_contiguous()
        assert np.array_equal(item.cpu().numpy(), g[f"ds_{tag}_{i}"])
    with pytest.raises(NotImplementedError):
        SEVIRTorchDataset(sevir_catalog=CC.catalog_frame(), sevir_data_dir="/data", aug_mode="1")
