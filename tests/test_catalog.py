"""CPU tests of the loader's catalog layer: prediff_b200.data.SEVIRCatalogEvents against the
unmodified reference SEVIRDataLoader run on a synthetic catalog (tests/golden/catalog.npz, written by
tests/golden/gen_golden.py::gen_catalog with an in-memory stand-in for h5py.File): event selection and order under every
filter, shuffling across epochs, the per-event reads, and the windows the reference makes of them (bit-exact, through the
numpy restatement of the window rule in oracle/data_oracle.py)."""
import os

import numpy as np
import pytest

from oracle import data_oracle as DO
from prediff_b200 import _lib as L
from prediff_b200.data import SEVIRCatalogEvents
from tests.golden import catalog_cases as CC

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "catalog.npz"))
FILES = CC.catalog_files()


def opener(path):
    assert path.startswith("/data/")
    return FILES[path[len("/data/"):]]


def make(tag):
    ckw, lkw = CC.CASES[tag]
    ckw = dict(ckw)
    return SEVIRCatalogEvents(CC.catalog_frame(), "/data", data_types=ckw.pop("data_types", ("vil",)), open_file=opener, **ckw), lkw


@pytest.mark.parametrize("tag", list(CC.CASES))
def test_event_selection_and_order_equal_the_reference(tag):
    ev, _ = make(tag)
    assert [str(v) for v in ev._samples["vil_filename"].values] == [str(v) for v in G[f"{tag}_files"]]
    assert np.array_equal(ev._samples["vil_index"].values, G[f"{tag}_index"])
    assert ev.shape == (len(G[f"{tag}_index"]), CC.H, CC.W, CC.T_RAW) and ev.dtype == np.uint8


def test_shuffle_is_reapplied_every_epoch_like_the_reference():
    ev, _ = make("shuffle")
    plain, _ = make("default")
    assert not np.array_equal(ev._samples["vil_index"].values, plain._samples["vil_index"].values) or \
        [str(v) for v in ev._samples["vil_filename"].values] != [str(v) for v in plain._samples["vil_filename"].values]
    ev.reset()
    assert [str(v) for v in ev._samples["vil_filename"].values] == [str(v) for v in G["shuffle_files_epoch2"]]
    assert np.array_equal(ev._samples["vil_index"].values, G["shuffle_index_epoch2"])


@pytest.mark.parametrize("tag", list(CC.CASES))
def test_reads_and_windows_equal_the_reference_batches(tag):
    """events[e0:e1] = the reference's _read_data rows; windows of them (oracle rule, pinned by loader.npz) = dl[i]."""
    ev, lkw = make(tag)
    for k in range(len(ev)):
        f, i = str(G[f"{tag}_files"][k]), int(G[f"{tag}_index"][k])
        assert np.array_equal(ev[k], FILES[f]["vil"][i])
    assert np.array_equal(ev[-1], ev[len(ev) - 1]) and ev[2:2].shape == (0, CC.H, CC.W, CC.T_RAW)
    with pytest.raises(IndexError):
        ev[len(ev)]
    n = int(G[f"{tag}_len"])
    n_seq = 1 + (CC.T_RAW - lkw["seq_len"]) // lkw["stride"]
    assert n == (len(ev) * n_seq) // lkw["batch_size"]
    full = ev[0:len(ev)]
    for b in range(n):
        want = G[f"{tag}_b{b}"]
        got = DO.idx_sample(full, b, lkw["batch_size"], lkw["seq_len"], lkw["stride"], lkw["rescale_method"], lkw["layout"])
        assert got.dtype == want.dtype and np.array_equal(got, want), (tag, b)


def test_missing_h5py_without_an_opener_is_an_error():
    try:
        import h5py  # noqa: F401
        pytest.skip("h5py is importable here")
    except ImportError:
        pass
    with pytest.raises(L.PDError):
        SEVIRCatalogEvents(CC.catalog_frame(), "/data")
    with pytest.raises(NotImplementedError):
        SEVIRCatalogEvents(CC.catalog_frame(), "/data", data_types=("vil", "lght"), open_file=opener)
