"""CPU test: the mirrors accept exactly the keyword arguments the reference inference script passes for the shipped
SEVIR-LR config (scripts/prediff/sevirlr/cfg.yaml:78-220 -> train_sevirlr_prediff.py:91-206) - the values below are the
cfg.yaml values, copied here because the GPU box has no /root/reference. This is the "option B" drop-in of
INTEGRATION.md: `class PreDiffSEVIRPLModule(prediff_b200.diffusion.LatentDiffusion)`."""
import torch

from prediff_b200.alignment import SEVIRAvgIntensityAlignment
from prediff_b200.diffusion import LatentDiffusion
from prediff_b200.unet import CuboidTransformerUNet
from prediff_b200.vae import AutoencoderKL

LATENT_MODEL = dict(   # cfg.yaml:157-206 as passed at train_sevirlr_prediff.py:91-137
    input_shape=[7, 16, 16, 64], target_shape=[6, 16, 16, 64], base_units=256, scale_alpha=1.0, num_heads=4,
    attn_drop=0.1, proj_drop=0.1, ffn_drop=0.1, downsample=2, downsample_type="patch_merge", upsample_type="upsample",
    upsample_kernel_size=3, depth=[4, 4], block_attn_patterns=["axial", "axial"], num_global_vectors=0,
    use_global_vector_ffn=False, use_global_self_attn=True, separate_global_qkv=True, global_dim_ratio=1,
    ffn_activation="gelu", gated_ffn=False, norm_layer="layer_norm", padding_type="zeros", checkpoint_level=0,
    pos_embed_type="t+h+w", use_relative_pos=True, self_attn_use_final_proj=True, attn_linear_init_mode="0",
    ffn_linear_init_mode="0", ffn2_linear_init_mode="2", attn_proj_linear_init_mode="2", conv_init_mode="0",
    down_linear_init_mode="0", up_linear_init_mode="0", global_proj_linear_init_mode="2", norm_init_mode="0",
    time_embed_channels_mult=4, time_embed_use_scale_shift_norm=False, time_embed_dropout=0.0, unet_res_connect=True)
VAE = dict(   # cfg.yaml:207-218 as passed at train_sevirlr_prediff.py:140-149
    down_block_types=["DownEncoderBlock2D"] * 4, in_channels=1, block_out_channels=[128, 256, 512, 512], act_fn="silu",
    latent_channels=64, up_block_types=["UpDecoderBlock2D"] * 4, norm_num_groups=32, layers_per_block=2, out_channels=1)
DIFFUSION = dict(   # cfg.yaml:79-103 as passed at train_sevirlr_prediff.py:159-188 (loss_type / monitor: cfg.yaml optim)
    layout="NTHWC", data_shape=[6, 128, 128, 1], timesteps=1000, beta_schedule="linear", loss_type="l2",
    monitor="val/loss", use_ema=True, log_every_t=100, clip_denoised=False, linear_start=1e-4, linear_end=2e-2,
    cosine_s=8e-3, given_betas=None, original_elbo_weight=0., v_posterior=0., l_simple_weight=1.,
    parameterization="eps", learn_logvar=True, logvar_init=0., latent_shape=[6, 16, 16, 64],
    cond_stage_model="__is_first_stage__", num_timesteps_cond=None, cond_stage_trainable=False, cond_stage_forward=None,
    scale_by_std=False, scale_factor=1.0)
ALIGN_MODEL_ARGS = dict(   # cfg.yaml:110-155
    input_shape=[6, 16, 16, 64], out_channels=1, base_units=128, scale_alpha=1.0, depth=[1, 1], downsample=2,
    downsample_type="patch_merge", block_attn_patterns="axial", num_heads=4, attn_drop=0.1, proj_drop=0.1, ffn_drop=0.1,
    ffn_activation="gelu", gated_ffn=False, norm_layer="layer_norm", use_inter_ffn=True, hierarchical_pos_embed=False,
    pos_embed_type="t+h+w", padding_type="zeros", checkpoint_level=0, use_relative_pos=True,
    self_attn_use_final_proj=True, num_global_vectors=0, use_global_vector_ffn=True, use_global_self_attn=False,
    separate_global_qkv=False, global_dim_ratio=1, attn_linear_init_mode="0", ffn_linear_init_mode="0",
    ffn2_linear_init_mode="2", attn_proj_linear_init_mode="2", conv_init_mode="0", down_linear_init_mode="0",
    global_proj_linear_init_mode="2", norm_init_mode="0", time_embed_channels_mult=4,
    time_embed_use_scale_shift_norm=False, time_embed_dropout=0.0, pool="attention", readout_seq=True, out_len=6)


def test_mirrors_accept_the_shipped_config_as_the_script_passes_it():
    unet = CuboidTransformerUNet(**LATENT_MODEL)
    vae = AutoencoderKL(**VAE)

    class PreDiffSEVIRModule(LatentDiffusion):          # what the script's PL module does, minus Lightning
        def get_input(self, batch, **kwargs):           # train_sevirlr_prediff.py:718-758 (in_len 7, out_len 6)
            return batch[:, 7:], {"y": batch[:, :7]}

    ldm = PreDiffSEVIRModule(torch_nn_module=unet, first_stage_model=vae, **DIFFUSION)
    assert ldm.use_ema and ldm.learn_logvar and isinstance(ldm.logvar, torch.nn.Parameter)
    assert sum(p.numel() for p in unet.parameters()) == 136817538 and sum(p.numel() for p in vae.parameters()) == 84499393
    sd = ldm.state_dict()
    for key in ("logvar", "model_ema.decay", "model_ema.num_updates", "betas", "posterior_mean_coef2",
                "torch_nn_module.final_proj.weight", "first_stage_model.quant_conv.weight",
                "cond_stage_model.quant_conv.weight", "model_ema.final_projweight"):
        assert key in sd, key
    with ldm.ema_scope():
        pass
    al = SEVIRAvgIntensityAlignment(alignment_type="avg_x", guide_scale=50.0, model_type="cuboid",
                                    model_args=ALIGN_MODEL_ARGS, model_ckpt_path=None)
    ldm.set_alignment(alignment_fn=al.get_mean_shift)
    assert ldm._native_alignment(True, {"avg_x_gt": torch.zeros(4, 1)}) is al
    x, c = ldm.get_input(torch.zeros(2, 13, 128, 128, 1))
    assert tuple(x.shape) == (2, 6, 128, 128, 1) and tuple(c["y"].shape) == (2, 7, 128, 128, 1)
