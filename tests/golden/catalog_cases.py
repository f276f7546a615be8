"""Synthetic SEVIR catalog + event files for the catalog-layer goldens (gen_golden.py writes catalog.npz from the unmodified
reference SEVIRDataLoader with an in-memory stand-in for h5py.File; tests/test_catalog.py reads it)."""
import datetime

import numpy as np

H, W, T_RAW = 16, 24, 25
FILES = {"vil/2019/SEVIR_VIL_A.h5": 7, "vil/2019/SEVIR_VIL_B.h5": 6}   # file name -> number of stored events
IR_FILE = "ir069/2019/SEVIR_IR069_A.h5"


def catalog_files(seed=8181):
    """file name -> {"vil" | "ir069": uint8 (n, H, W, T_RAW)} ('NHWT' like the SEVIR-LR HDF5 datasets)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {f: {"vil": rng.integers(0, 256, size=(n, H, W, T_RAW), dtype=np.uint8)} for f, n in FILES.items()}
    out[IR_FILE] = {"ir069": rng.integers(0, 256, size=(13, H, W, T_RAW), dtype=np.uint8)}
    return out


def catalog_frame():
    """13 'vil' rows over two files (one id listed twice = the SEVIR bug the reference drops, two rows with missing data),
    'ir069' rows for most but not all ids, times spread over 2019."""
    import pandas as pd
    ids = ["S840001", "R19010112345678", "S835002", "R19020323456789", "S851003", "S799004", "R19040534567890", "S812005",
           "S812005", "R19060745678901", "S860006", "S801007", "R19080956789012"]
    rows = []
    per_file = {f: 0 for f in FILES}
    names = list(FILES)
    for k, ev in enumerate(ids):
        f = names[0] if per_file[names[0]] < FILES[names[0]] and k % 2 == 0 or per_file[names[1]] >= FILES[names[1]] else names[1]
        t = datetime.datetime(2019, 1, 5, 3, 0) + datetime.timedelta(days=26 * k, hours=5 * k)
        rows.append(dict(id=ev, file_name=f, file_index=per_file[f], img_type="vil", time_utc=t,
                         pct_missing=0.25 if k in (3, 10) else 0.0))
        per_file[f] += 1
        if k not in (5, 11):   # two events without a co-located ir069 image
            rows.append(dict(id=ev, file_name=IR_FILE, file_index=k, img_type="ir069", time_utc=t, pct_missing=0.0))
    df = pd.DataFrame(rows)
    df["time_utc"] = pd.to_datetime(df["time_utc"])
    return df.sample(frac=1, random_state=4).reset_index(drop=True)   # catalog rows come in no particular order


def day_hours(t):
    return np.logical_and(t.dt.hour >= 6, t.dt.hour <= 20)


def stored_in_file_a(c):
    return [("_A" in f) or ("IR069" in f) for f in c.file_name]


# tag -> (catalog kwargs of the reference constructor, loader kwargs)
CASES = {
    "default": (dict(), dict(batch_size=4, seq_len=13, stride=6, layout="NTHWC", rescale_method="01")),
    "dates": (dict(start_date=datetime.datetime(2019, 2, 1), end_date=datetime.datetime(2019, 9, 1)),
              dict(batch_size=3, seq_len=13, stride=6, layout="NTHWC", rescale_method="01")),
    "filters": (dict(datetime_filter=day_hours, catalog_filter=stored_in_file_a),
                dict(batch_size=2, seq_len=10, stride=5, layout="NTHWC", rescale_method="sevir")),
    "nofilter": (dict(catalog_filter=None), dict(batch_size=4, seq_len=13, stride=12, layout="NTHW", rescale_method="01")),
    "shuffle": (dict(shuffle=True, shuffle_seed=3), dict(batch_size=4, seq_len=13, stride=6, layout="NTHWC", rescale_method="01")),
    "colocated": (dict(data_types=["vil", "ir069"]), dict(batch_size=4, seq_len=13, stride=6, layout="NTHWC", rescale_method="01")),
}
