"""Cuboid self-attention patterns and the cuboid geometry they imply - host-side mirror of
src/prediff/models/cuboid_transformer/cuboid_transformer_patterns.py:10-118 (pattern registry) and of the index
arithmetic of CuboidSelfAttentionLayer (cuboid_transformer.py:388-429 cuboid_reorder, :470-528 attention mask,
:563-592 size/shift update, :714-734 relative position index).

A pattern maps a level's (T, H, W, C) to three per-layer lists: cuboid sizes, strategies ('l' local / 'd' dilated)
and shift sizes. `resolve(name, shape)` returns them; `layer_geometry` turns one layer into the flat tables the CUDA
kernel consumes (the same tables are built in C++ by `build_cuboid_tables`, csrc/cuboid.cu - tests compare both
against the unmodified reference functions).
"""
import re
from typing import List, Sequence, Tuple

import numpy as np

Layer = Tuple[Tuple[int, int, int], Tuple[str, str, str], Tuple[int, int, int]]


def _lll(n):
    return [("l", "l", "l")] * n


def _zeros(n):
    return [(0, 0, 0)] * n


def full_attention(shape):
    T, H, W, _ = shape
    return [(T, H, W)], _lll(1), _zeros(1)


def self_axial(shape):
    T, H, W, _ = shape
    return [(T, 1, 1), (1, H, 1), (1, 1, W)], _lll(3), _zeros(3)


def self_video_swin(shape, P=2, M=4):
    T, H, W, _ = shape
    P = min(P, T)
    M = min(M, H, W)
    return [(P, M, M), (P, M, M)], _lll(2), [(0, 0, 0), (P // 2, M // 2, M // 2)]


def self_divided_space_time(shape):
    T, H, W, _ = shape
    return [(T, 1, 1), (1, H, W)], _lll(2), _zeros(2)


def self_spatial_lg_v1(shape, M=4):
    T, H, W, _ = shape
    if H <= M and W <= M:
        return [(T, 1, 1), (1, H, W)], _lll(2), _zeros(2)
    return [(T, 1, 1), (1, M, M), (1, M, M)], [("l", "l", "l"), ("l", "l", "l"), ("d", "d", "d")], _zeros(3)


def self_axial_space_dilate_K(shape, K=2):
    T, H, W, _ = shape
    K = min(K, H, W)
    sizes = [(T, 1, 1), (1, H // K, 1), (1, H // K, 1), (1, 1, W // K), (1, 1, W // K)]
    strat = [("l", "l", "l"), ("d", "d", "d"), ("l", "l", "l"), ("d", "d", "d"), ("l", "l", "l")]
    return sizes, strat, _zeros(5)


_PARAM = [(re.compile(r"^video_swin_(\d+)x(\d+)$"), lambda m: (lambda s: self_video_swin(s, P=int(m[1]), M=int(m[2]))),
           lambda m: int(m[1]) in (1, 2, 4, 8, 10) and int(m[2]) in (1, 2, 4, 8, 16, 32)),
          (re.compile(r"^spatial_lg_(\d+)$"), lambda m: (lambda s: self_spatial_lg_v1(s, M=int(m[1]))),
           lambda m: int(m[1]) in (1, 2, 4, 8, 16, 32)),
          (re.compile(r"^axial_space_dilate_(\d+)$"), lambda m: (lambda s: self_axial_space_dilate_K(s, K=int(m[1]))),
           lambda m: int(m[1]) in (2, 4, 8))]
_BASIC = {"full": full_attention, "axial": self_axial, "video_swin": self_video_swin,
          "divided_st": self_divided_space_time, "spatial_lg_v1": self_spatial_lg_v1}


def get(name: str):
    """CuboidSelfAttentionPatterns.get(name) (same registered names, cuboid_transformer_patterns.py:60-118)."""
    if name in _BASIC:
        return _BASIC[name]
    for rx, make, ok in _PARAM:
        m = rx.match(name)
        if m and ok(m):
            return make(m)
    raise KeyError(f"cuboid self-attention pattern '{name}' is not registered")


def resolve(name: str, shape: Sequence[int]) -> List[Layer]:
    sizes, strat, shift = get(name)(tuple(shape))
    return [(tuple(int(v) for v in a), tuple(b), tuple(int(v) for v in c)) for a, b, c in zip(sizes, strat, shift)]


AXIAL_LAYERS = None  # sentinel: UNetConfig.patterns == None means "axial" at both levels


def effective_size_shift(data_shape, size, shift, strategy):
    """update_cuboid_size_shift_size (cuboid_transformer.py:563-592)."""
    size, shift = list(size), list(shift)
    for i in range(3):
        if strategy[i] == "d":
            shift[i] = 0
        if data_shape[i] <= size[i]:
            size[i] = data_shape[i]
            shift[i] = 0
    return tuple(size), tuple(shift)


def layer_geometry(data_shape, size0, strategy, shift0, padding_type="zeros"):
    """Flat tables of one CuboidSelfAttentionLayer on a (T, H, W) grid.

    Returns dict(size, shift, pad, num_cuboids, volume,
                 tok  int32 (num_cuboids * volume): row of the token inside one sample's [T*H*W] block, -1 = padding,
                 lab  int32 (num_cuboids * volume): shifted-window region label; -1 = masked out (padding, 'ignore'),
                 rel  int32 (volume): code of the in-cuboid position such that
                      relative_position_index[i, j] == rel[i] - rel[j] + rel_off,
                 rel_off int,
                 dst  int32 (num_cuboids * volume) or None: 'nearest' padding only - the token the slot's result is written to,
                 gmask int32 (num_cuboids * volume) or None: 'ignore' padding only - 1 = the slot is visible to the queries of
                      the global vectors. The reference flattens the validity grid of the padded, rolled frame in RASTER
                      order and applies it to the cuboid-ordered keys as it is (cuboid_transformer.py:915-945); so is this).
    The label / validity rules restate compute_cuboid_self_attention_mask (cuboid_transformer.py:470-528):
    mask[c, i, j] = lab[c, i] == lab[c, j] and both >= 0.
    """
    assert padding_type in ("zeros", "ignore", "nearest")
    dims = tuple(int(v) for v in data_shape)
    size, shift = effective_size_shift(dims, size0, shift0, strategy)
    pad = tuple((size[a] - dims[a] % size[a]) % size[a] for a in range(3))
    padded = tuple(dims[a] + pad[a] for a in range(3))
    n = tuple(padded[a] // size[a] for a in range(3))
    nc, vol = n[0] * n[1] * n[2], size[0] * size[1] * size[2]
    tok = np.empty((nc, vol), np.int32)
    lab = np.empty((nc, vol), np.int32)
    # 'nearest' (models/utils.py:228-270): padded position o copies token floor(o * dims / padded) (F.interpolate to the
    # padded size) and token t takes the result of position floor(t * padded / dims) (F.interpolate back), both in torch's
    # float32 index arithmetic; `dst` is then the token a slot's result goes to (-1: nobody) and `tok` the token it copies
    nearest = padding_type == "nearest" and any(pad)
    dst = np.full((nc, vol), -1, np.int32) if nearest else None
    if nearest:
        near_src, near_dst = [], []
        for a in range(3):
            up, down = np.float32(dims[a]) / np.float32(padded[a]), np.float32(padded[a]) / np.float32(dims[a])
            near_src.append([min(int(np.floor(np.float32(o) * up)), dims[a] - 1) for o in range(padded[a])])
            d = [-1] * padded[a]
            for t in range(dims[a]):
                d[min(int(np.floor(np.float32(t) * down)), padded[a] - 1)] = t
            near_dst.append(d)
    any_shift = any(s > 0 for s in shift)
    for c in range(nc):
        cub = (c // (n[1] * n[2]), (c // n[2]) % n[1], c % n[2])
        for i in range(vol):
            inn = (i // (size[1] * size[2]), (i // size[2]) % size[1], i % size[2])
            valid, label, src = True, 0, []
            for a in range(3):
                # cuboid_reorder: 'l' -> (block, in-block), 'd' -> (in-block, block) split of the padded axis
                p = cub[a] * size[a] + inn[a] if strategy[a] == "l" else inn[a] * n[a] + cub[a]
                # region label in the rolled frame (3 slices per axis; a zero shift leaves one region, label 2)
                if shift[a] == 0:
                    la = 2
                else:
                    la = 0 if p < padded[a] - size[a] else (1 if p < padded[a] - shift[a] else 2)
                label = label * 3 + la
                o = (p + shift[a]) % padded[a] if any_shift else p   # torch.roll(x, -shift): rolled[p] = x[p + shift]
                valid = valid and o < dims[a]
                src.append(o)
            if nearest:   # src = position in the padded (un-rolled) grid
                st = [near_src[a][src[a]] for a in range(3)]
                dt = [near_dst[a][src[a]] for a in range(3)]
                tok[c, i] = (st[0] * dims[1] + st[1]) * dims[2] + st[2]
                dst[c, i] = (dt[0] * dims[1] + dt[1]) * dims[2] + dt[2] if min(dt) >= 0 else -1
                lab[c, i] = label
                continue
            tok[c, i] = (src[0] * dims[1] + src[1]) * dims[2] + src[2] if valid else -1
            lab[c, i] = -1 if (not valid and padding_type == "ignore") else label
    gmask = None
    if padding_type == "ignore":
        g = np.zeros(padded, np.int32)
        g[:dims[0], :dims[1], :dims[2]] = 1
        if any_shift:
            g = np.roll(g, [-v for v in shift], (0, 1, 2))
        gmask = np.ascontiguousarray(g).reshape(-1)
    b0 = tuple(int(v) for v in size0)   # the index buffer is built from the constructor's cuboid size and sliced
    s1, s2 = (2 * b0[1] - 1) * (2 * b0[2] - 1), 2 * b0[2] - 1
    i = np.arange(vol)
    rel = ((i // (b0[1] * b0[2])) * s1 + ((i // b0[2]) % b0[1]) * s2 + i % b0[2]).astype(np.int32)
    rel_off = (b0[0] - 1) * s1 + (b0[1] - 1) * s2 + (b0[2] - 1)
    return dict(size=size, shift=shift, pad=pad, num_cuboids=nc, volume=vol, tok=tok.reshape(-1), lab=lab.reshape(-1),
                rel=rel, rel_off=int(rel_off), dst=None if dst is None else dst.reshape(-1), gmask=gmask)
