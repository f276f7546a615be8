"""Shared case lists for the cuboid-pattern goldens (gen_golden.py writes them, tests read them)."""

# (tag, (T, H, W), C, heads, cuboid_size, strategy, shift_size, padding_type)
LAYER_CASES = [
    ("swin_pad_shift_z", (13, 8, 8), 32, 2, (4, 4, 4), "lll", (2, 2, 2), "zeros"),
    ("swin_pad_shift_i", (13, 8, 8), 32, 2, (4, 4, 4), "lll", (2, 2, 2), "ignore"),
    ("dilated_i", (13, 8, 8), 32, 2, (1, 4, 4), "ddd", (0, 0, 0), "ignore"),
    ("mixed_ragged_z", (6, 7, 9), 32, 2, (4, 3, 4), "ldl", (2, 1, 2), "zeros"),
    ("mixed_ragged_i", (6, 7, 9), 32, 2, (4, 3, 4), "ldl", (2, 1, 2), "ignore"),
    ("clipped_i", (5, 6, 6), 32, 2, (8, 4, 8), "lll", (4, 2, 4), "ignore"),
    ("full_z", (5, 8, 8), 32, 2, (5, 8, 8), "lll", (0, 0, 0), "zeros"),
    ("plane_hd32", (3, 8, 8), 64, 2, (1, 8, 8), "lll", (0, 0, 0), "zeros"),
]
LAYER_SEED = 4004

# padding_type='nearest' (models/utils.py:228-270: the grid is resampled with F.interpolate before and after the attention);
# goldens in patterns_nearest.npz. Same tuple layout as LAYER_CASES.
NEAREST_LAYER_CASES = [
    ("swin_pad_shift_n", (13, 8, 8), 32, 2, (4, 4, 4), "lll", (2, 2, 2), "nearest"),
    ("dilated_n", (13, 8, 8), 32, 2, (2, 4, 4), "ddd", (0, 0, 0), "nearest"),
    ("mixed_ragged_n", (6, 7, 9), 32, 2, (4, 3, 4), "ldl", (2, 1, 2), "nearest"),
    ("ragged_all_n", (5, 7, 9), 16, 2, (3, 4, 5), "lld", (1, 2, 0), "nearest"),
    ("clipped_n", (5, 6, 6), 32, 2, (8, 4, 8), "lll", (4, 2, 4), "nearest"),
    ("nopad_n", (4, 8, 8), 32, 2, (2, 4, 4), "lll", (1, 2, 2), "nearest"),
    ("swin2x8_hd64_n", (13, 8, 8), 128, 2, (2, 8, 8), "lll", (1, 4, 4), "nearest"),
]
# tiny UNet with a pattern that pads T = 13 (video_swin_2x8: cuboids of 2 frames), B = 1, t = 500
NEAREST_UNET_CASES = [("swin_n", ("video_swin_2x8", "video_swin_2x8"), "nearest")]

# (tag, block_attn_patterns, padding_type): tiny UNet (base_units 64, depth (1, 1)), B = 1, t = 500
UNET_CASES = [
    ("swin_lg", ("video_swin_4x4", "spatial_lg_4"), "zeros"),
    ("dst_dilate", ("divided_st", "axial_space_dilate_2"), "ignore"),
    ("full_swin", ("full", "video_swin_2x8"), "ignore"),
]

# block_attn_patterns=None: the reference constructor's default block_cuboid_size / strategy / shift_size
# (cuboid_transformer_unet.py:35-40), the same lists for both blocks
_DEFAULT_BLOCK = (((4, 4, 4), ("l", "l", "l"), (0, 0, 0)), ((4, 4, 4), ("d", "d", "d"), (0, 0, 0)))
EXPLICIT_LAYERS = (_DEFAULT_BLOCK, _DEFAULT_BLOCK)


# global vectors (cuboid_transformer.py:864-945), goldens in global_vectors.npz:
# (tag, (T, H, W), C, heads, cuboid_size, strategy, shift_size, padding_type, num_global, use_global_self_attn)
GV_LAYER_CASES = [
    ("gv_axial_t", (13, 8, 8), 32, 2, (13, 1, 1), "lll", (0, 0, 0), "zeros", 4, False),
    ("gv_swin_z", (13, 8, 8), 32, 2, (4, 4, 4), "lll", (2, 2, 2), "zeros", 4, False),       # padded slots take part (zero rows)
    ("gv_swin_i", (13, 8, 8), 32, 2, (4, 4, 4), "lll", (2, 2, 2), "ignore", 4, True),        # raster-order slot mask, shifted
    ("gv_dilated_i", (13, 8, 8), 32, 2, (2, 4, 4), "ddd", (0, 0, 0), "ignore", 8, False),    # mask order != slot order
    ("gv_ragged_z", (6, 7, 9), 32, 2, (4, 3, 4), "ldl", (2, 1, 2), "zeros", 3, True),
    ("gv_ragged_i", (6, 7, 9), 32, 2, (4, 3, 4), "ldl", (2, 1, 2), "ignore", 5, False),
    ("gv_swin_n", (13, 8, 8), 32, 2, (4, 4, 4), "lll", (2, 2, 2), "nearest", 4, True),       # resampled copies take part
    ("gv_plane_hd32", (3, 8, 8), 64, 2, (1, 8, 8), "lll", (0, 0, 0), "zeros", 8, True),
    ("gv_full_hd64", (5, 8, 8), 128, 2, (5, 8, 8), "lll", (0, 0, 0), "ignore", 16, True),    # 320-slot cuboid: 5 key chunks + globals
]
# (tag, block_attn_patterns, padding_type, num_global_vectors, use_global_vector_ffn, use_global_self_attn): tiny UNet, B = 2
GV_UNET_CASES = [
    ("gv_axial", ("axial", "axial"), "zeros", 4, True, False),
    ("gv_swin_sa", ("video_swin_2x8", "spatial_lg_4"), "ignore", 8, True, True),
    ("gv_dst_noffn", ("divided_st", "axial_space_dilate_2"), "nearest", 2, False, True),
]


# separate_global_qkv=True (the shipped cfg.yaml's value for the UNet): same tuple layouts, goldens in global_vectors_sep.npz
GV_SEP_LAYER_CASES = [
    ("gvs_axial_t", (13, 8, 8), 32, 2, (13, 1, 1), "lll", (0, 0, 0), "zeros", 4, True),
    ("gvs_swin_i", (13, 8, 8), 32, 2, (4, 4, 4), "lll", (2, 2, 2), "ignore", 4, True),
    ("gvs_dilated_z", (13, 8, 8), 32, 2, (2, 4, 4), "ddd", (0, 0, 0), "zeros", 8, False),
    ("gvs_ragged_n", (6, 7, 9), 32, 2, (4, 3, 4), "ldl", (2, 1, 2), "nearest", 3, True),
    ("gvs_full_hd64", (5, 8, 8), 128, 2, (5, 8, 8), "lll", (0, 0, 0), "ignore", 16, True),
]
GV_SEP_UNET_CASES = [
    ("gvs_axial", ("axial", "axial"), "zeros", 4, False, True),          # the shipped flags: no global FFN, global self-attention
    ("gvs_swin", ("video_swin_2x8", "spatial_lg_4"), "ignore", 8, True, False),
]


def gv_sep_layer_spec(C, heads, size, self_attn):
    n_rel = (2 * size[0] - 1) * (2 * size[1] - 1) * (2 * size[2] - 1)
    s = [("a.relative_position_bias_table", (n_rel, heads)), ("a.qkv.weight", (3 * C, C)),
         ("a.l2g_q_net.weight", (C, C)), ("a.l2g_global_kv_net.weight", (2 * C, C)), ("a.g2l_global_q_net.weight", (C, C)),
         ("a.g2l_k_net.weight", (C, C)), ("a.g2l_v_net.weight", (C, C))]
    if self_attn:
        s += [("a.g2g_global_qkv_net.weight", (3 * C, C))]
    return s + [("a.proj.weight", (C, C)), ("a.proj.bias", (C,)), ("a.global_proj.weight", (C, C)), ("a.global_proj.bias", (C,)),
                ("a.norm.weight", (C,)), ("a.norm.bias", (C,)), ("a.global_vec_norm.weight", (C,)), ("a.global_vec_norm.bias", (C,))]


def gv_layer_spec(C, heads, size):
    n_rel = (2 * size[0] - 1) * (2 * size[1] - 1) * (2 * size[2] - 1)
    return [("a.relative_position_bias_table", (n_rel, heads)), ("a.qkv.weight", (3 * C, C)),
            ("a.global_qkv.weight", (3 * C, C)), ("a.proj.weight", (C, C)), ("a.proj.bias", (C,)),
            ("a.global_proj.weight", (C, C)), ("a.global_proj.bias", (C,)), ("a.norm.weight", (C,)), ("a.norm.bias", (C,)),
            ("a.global_vec_norm.weight", (C,)), ("a.global_vec_norm.bias", (C,))]


def layer_spec(C, heads, size):
    n_rel = (2 * size[0] - 1) * (2 * size[1] - 1) * (2 * size[2] - 1)
    return [("a.relative_position_bias_table", (n_rel, heads)), ("a.qkv.weight", (3 * C, C)),
            ("a.proj.weight", (C, C)), ("a.proj.bias", (C,)), ("a.norm.weight", (C,)), ("a.norm.bias", (C,))]


def sweep_cases(n=32, seed=9009):
    """Deterministic random sweep of small CuboidSelfAttentionLayer configurations (grid <= 6 x 7 x 7, C = 16, 2 heads):
    random cuboid sizes (some larger than the grid -> clipped), strategies, shifts and both padding types."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for k in range(n):
        dims = tuple(int(v) for v in rng.integers(2, [7, 8, 8]))
        size = tuple(int(v) for v in rng.integers(1, [6, 6, 6]))
        strat = "".join(rng.choice(["l", "d"]) for _ in range(3))
        shift = tuple(int(rng.integers(0, max(1, s // 2 + 1))) for s in size)
        pad = "zeros" if rng.integers(0, 2) == 0 else "ignore"
        out.append((f"sweep{k}", dims, 16, 2, size, strat, shift, pad))
    return out
