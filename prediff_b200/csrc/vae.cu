// AutoencoderKL encode / decode as launch plans over the shared kernels.
// Reference: src/prediff/taming/autoencoder_kl.py:80-113, vae.py:70-86 (Encoder.forward), :150-166
// (Decoder.forward), unet_2d_blocks.py:158-165,217-225,271-279, resnet.py:454-495 (ResnetBlock2D), :108-143
// (Upsample2D), :181-190 (Downsample2D), attention.py:136-189 (AttentionBlock).
// Activations are channels-last [N][H][W][C]; fp32 residual stream, bf16 GEMM operands.
#include "vae.cuh"
#include <algorithm>
#include <cmath>

namespace pd {

// ---- plan building ------------------------------------------------------------------------------------------
struct VAE::Ctx {
    Plan* pl;
    int N;
    float* f[3];      // fp32 pool: cur / conv1-out / spare
    bf16 *a, *cast;   // GN output; casts, upsampled / parity-split tensors, V^T
    bf16* qkv;        // attention q|k|v, later the attention output
    double* gn_sums;
    int gn_slot = 0, G = 32;
    int cur = 0;      // index of the stream buffer in f[]
    float* stream() const { return f[cur]; }
    float* other(int k) const { return f[(cur + k) % 3]; }
    double* next_sums() { return gn_sums + (size_t)(gn_slot++) * N * 128 * 2; }
};

struct VAE::DirPlan {
    Arena arena;
    Ctx ctx;
    Plan plan;
    GemmOp last_op;
    float* first_buf = nullptr;
    size_t in_slot = 0, out_slot = 0;
    Ctx& ctx_storage() { return ctx; }
};

VAE::VAE(const pd_vae_config& c) : cfg(c) { declare_weights(); }
VAE::~VAE() = default;

void VAE::declare_resnet(const std::string& p, int cin, int cout) {
    ws.declare(p + ".norm1.weight", {cin});
    ws.declare(p + ".norm1.bias", {cin});
    ws.declare(p + ".conv1.weight", {cout, cin, 3, 3});
    ws.declare(p + ".conv1.bias", {cout});
    ws.declare(p + ".norm2.weight", {cout});
    ws.declare(p + ".norm2.bias", {cout});
    ws.declare(p + ".conv2.weight", {cout, cout, 3, 3});
    ws.declare(p + ".conv2.bias", {cout});
    if (cin != cout) {
        ws.declare(p + ".conv_shortcut.weight", {cout, cin, 1, 1});
        ws.declare(p + ".conv_shortcut.bias", {cout});
    }
}

void VAE::declare_mid(const std::string& p, int c) {
    const std::string a = p + ".attentions.0";
    ws.declare(a + ".group_norm.weight", {c});
    ws.declare(a + ".group_norm.bias", {c});
    for (const char* n : {"query", "key", "value", "proj_attn"}) {
        ws.declare(a + "." + n + ".weight", {c, c});
        ws.declare(a + "." + n + ".bias", {c});
    }
    declare_resnet(p + ".resnets.0", c, c);
    declare_resnet(p + ".resnets.1", c, c);
}

// Same names / shapes / order as the reference AutoencoderKL.state_dict().
void VAE::declare_weights() {
    const int* boc = cfg.block_out_channels;
    const int L = cfg.latent_channels;
    ws.declare("encoder.conv_in.weight", {boc[0], cfg.in_channels, 3, 3});
    ws.declare("encoder.conv_in.bias", {boc[0]});
    int cin = boc[0];
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < cfg.layers_per_block; ++j)
            declare_resnet(strf("encoder.down_blocks.%d.resnets.%d", i, j), j == 0 ? cin : boc[i], boc[i]);
        if (i != 3) {
            ws.declare(strf("encoder.down_blocks.%d.downsamplers.0.conv.weight", i), {boc[i], boc[i], 3, 3});
            ws.declare(strf("encoder.down_blocks.%d.downsamplers.0.conv.bias", i), {boc[i]});
        }
        cin = boc[i];
    }
    declare_mid("encoder.mid_block", boc[3]);
    ws.declare("encoder.conv_norm_out.weight", {boc[3]});
    ws.declare("encoder.conv_norm_out.bias", {boc[3]});
    ws.declare("encoder.conv_out.weight", {2 * L, boc[3], 3, 3});
    ws.declare("encoder.conv_out.bias", {2 * L});
    ws.declare("decoder.conv_in.weight", {boc[3], L, 3, 3});
    ws.declare("decoder.conv_in.bias", {boc[3]});
    cin = boc[3];
    for (int i = 0; i < 4; ++i) {
        const int cout = boc[3 - i];
        for (int j = 0; j < cfg.layers_per_block + 1; ++j)
            declare_resnet(strf("decoder.up_blocks.%d.resnets.%d", i, j), j == 0 ? cin : cout, cout);
        if (i != 3) {
            ws.declare(strf("decoder.up_blocks.%d.upsamplers.0.conv.weight", i), {cout, cout, 3, 3});
            ws.declare(strf("decoder.up_blocks.%d.upsamplers.0.conv.bias", i), {cout});
        }
        cin = cout;
    }
    declare_mid("decoder.mid_block", boc[3]);
    ws.declare("decoder.conv_norm_out.weight", {boc[0]});
    ws.declare("decoder.conv_norm_out.bias", {boc[0]});
    ws.declare("decoder.conv_out.weight", {cfg.out_channels, boc[0], 3, 3});
    ws.declare("decoder.conv_out.bias", {cfg.out_channels});
    ws.declare("quant_conv.weight", {2 * L, 2 * L, 1, 1});
    ws.declare("quant_conv.bias", {2 * L});
    ws.declare("post_quant_conv.weight", {L, L, 1, 1});
    ws.declare("post_quant_conv.bias", {L});
}

int VAE::validate() const {
    PD_CHECK(cfg.in_channels == 1 && cfg.out_channels == 1, PD_ERR_SHAPE,
             "vae: only single-channel frames are built (SEVIR VIL); got in=%d out=%d", cfg.in_channels, cfg.out_channels);
    PD_CHECK(cfg.h % 8 == 0 && cfg.w % 8 == 0, PD_ERR_SHAPE, "vae: H, W must be multiples of 8");
    const int hl = cfg.h / 8, wl = cfg.w / 8;
    PD_CHECK(cfg.w <= 128 && 128 % wl == 0 && (hl * wl) % 128 == 0, PD_ERR_SHAPE,
             "vae: frame %dx%d unsupported (needs W <= 128, (H/8)*(W/8) a multiple of 128)", cfg.h, cfg.w);
    for (int i = 0; i < 4; ++i)
        PD_CHECK(cfg.block_out_channels[i] % 64 == 0 && cfg.block_out_channels[i] % cfg.norm_num_groups == 0 &&
                     cfg.block_out_channels[i] <= 1024,
                 PD_ERR_SHAPE, "vae: block_out_channels[%d]=%d must be a multiple of 64 and of the group count", i,
                 cfg.block_out_channels[i]);
    PD_CHECK(cfg.latent_channels % 64 == 0, PD_ERR_SHAPE, "vae: latent_channels must be a multiple of 64");
    PD_CHECK(cfg.norm_num_groups >= 1 && cfg.norm_num_groups <= 128, PD_ERR_SHAPE, "vae: norm_num_groups");
    PD_CHECK(cfg.layers_per_block >= 1 && cfg.max_frames >= 1, PD_ERR_SHAPE, "vae: layers_per_block / max_frames");
    return PD_OK;
}

#define PD_GETW(dst, name)                \
    do {                                  \
        (dst) = ws.get(name);             \
        if (!(dst)) return PD_ERR_WEIGHT; \
    } while (0)

int VAE::pack_conv_w(const std::string& name, int co, int ci, int taps, bf16** out) {
    const float* w = ws.get(name);
    if (!w) return PD_ERR_WEIGHT;
    packed.emplace_back(new DevMem());
    PD_TRY(packed.back()->alloc((size_t)co * taps * ci * sizeof(bf16)));
    *out = packed.back()->as<bf16>();
    return pack_conv(w, *out, co, ci, taps, ci, 0);
}

int VAE::finalize_resnet(const std::string& p, int cin, int cout, Res2W* r) {
    r->cin = cin;
    r->cout = cout;
    PD_GETW(r->gn1_w, p + ".norm1.weight");
    PD_GETW(r->gn1_b, p + ".norm1.bias");
    PD_GETW(r->conv1_b, p + ".conv1.bias");
    PD_GETW(r->gn2_w, p + ".norm2.weight");
    PD_GETW(r->gn2_b, p + ".norm2.bias");
    PD_GETW(r->conv2_b, p + ".conv2.bias");
    PD_TRY(pack_conv_w(p + ".conv1.weight", cout, cin, 9, &r->conv1_w));
    PD_TRY(pack_conv_w(p + ".conv2.weight", cout, cout, 9, &r->conv2_w));
    r->sc_w = nullptr;
    r->sc_b = nullptr;
    if (cin != cout) {
        PD_TRY(pack_conv_w(p + ".conv_shortcut.weight", cout, cin, 1, &r->sc_w));
        PD_GETW(r->sc_b, p + ".conv_shortcut.bias");
    }
    return PD_OK;
}

int VAE::finalize_mid(const std::string& p, int c, MidW* m) {
    PD_TRY(finalize_resnet(p + ".resnets.0", c, c, &m->r0));
    PD_TRY(finalize_resnet(p + ".resnets.1", c, c, &m->r1));
    const std::string a = p + ".attentions.0";
    PD_GETW(m->gn_w, a + ".group_norm.weight");
    PD_GETW(m->gn_b, a + ".group_norm.bias");
    PD_GETW(m->proj_b, a + ".proj_attn.bias");
    PD_TRY(pack_conv_w(a + ".proj_attn.weight", c, c, 1, &m->proj_w));
    // q, k, v Linears share their input: one GEMM with the three weight matrices stacked ([3c][c]) and biases
    packed.emplace_back(new DevMem());
    PD_TRY(packed.back()->alloc((size_t)3 * c * c * sizeof(bf16)));
    m->qkv_w = packed.back()->as<bf16>();
    packed.emplace_back(new DevMem());
    PD_TRY(packed.back()->alloc((size_t)3 * c * sizeof(float)));
    float* qb = packed.back()->as<float>();
    const char* names[3] = {"query", "key", "value"};
    for (int i = 0; i < 3; ++i) {
        const float *w, *b;
        PD_GETW(w, a + "." + names[i] + ".weight");
        PD_GETW(b, a + "." + names[i] + ".bias");
        PD_TRY(pack_linear(w, m->qkv_w + (size_t)i * c * c, c, c, c, 0));
        PD_CUDA(cudaMemcpy(qb + (size_t)i * c, b, c * sizeof(float), cudaMemcpyDeviceToDevice));
    }
    m->qkv_b = qb;
    return PD_OK;
}

int VAE::finalize() {
    PD_TRY(gemm_init());
    PD_TRY(validate());
    PD_TRY(ws.check_complete());
    packed.clear();
    enc_plans.clear();
    dec_plans.clear();
    const int* boc = cfg.block_out_channels;
    const int L = cfg.latent_channels;
    PD_GETW(enc_in_w, "encoder.conv_in.weight");
    PD_GETW(enc_in_b, "encoder.conv_in.bias");
    enc_res.clear();
    int cin = boc[0];
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < cfg.layers_per_block; ++j) {
            Res2W r;
            PD_TRY(finalize_resnet(strf("encoder.down_blocks.%d.resnets.%d", i, j), j == 0 ? cin : boc[i], boc[i], &r));
            enc_res.push_back(r);
        }
        if (i != 3) {
            PD_TRY(pack_conv_w(strf("encoder.down_blocks.%d.downsamplers.0.conv.weight", i), boc[i], boc[i], 9, &down_w[i]));
            PD_GETW(down_b[i], strf("encoder.down_blocks.%d.downsamplers.0.conv.bias", i));
        }
        cin = boc[i];
    }
    PD_TRY(finalize_mid("encoder.mid_block", boc[3], &enc_mid));
    PD_GETW(enc_no_w, "encoder.conv_norm_out.weight");
    PD_GETW(enc_no_b, "encoder.conv_norm_out.bias");
    PD_TRY(pack_conv_w("encoder.conv_out.weight", 2 * L, boc[3], 9, &enc_out_w));
    PD_GETW(enc_out_b, "encoder.conv_out.bias");
    PD_TRY(pack_conv_w("quant_conv.weight", 2 * L, 2 * L, 1, &quant_w));
    PD_GETW(quant_b, "quant_conv.bias");
    PD_TRY(pack_conv_w("post_quant_conv.weight", L, L, 1, &pquant_w));
    PD_GETW(pquant_b, "post_quant_conv.bias");
    PD_TRY(pack_conv_w("decoder.conv_in.weight", boc[3], L, 9, &dec_in_w));
    PD_GETW(dec_in_b, "decoder.conv_in.bias");
    PD_TRY(finalize_mid("decoder.mid_block", boc[3], &dec_mid));
    dec_res.clear();
    cin = boc[3];
    for (int i = 0; i < 4; ++i) {
        const int cout = boc[3 - i];
        for (int j = 0; j < cfg.layers_per_block + 1; ++j) {
            Res2W r;
            PD_TRY(finalize_resnet(strf("decoder.up_blocks.%d.resnets.%d", i, j), j == 0 ? cin : cout, cout, &r));
            dec_res.push_back(r);
        }
        if (i != 3) {
            PD_TRY(pack_conv_w(strf("decoder.up_blocks.%d.upsamplers.0.conv.weight", i), cout, cout, 9, &up_w[i]));
            PD_GETW(up_b[i], strf("decoder.up_blocks.%d.upsamplers.0.conv.bias", i));
        }
        cin = cout;
    }
    PD_GETW(dec_no_w, "decoder.conv_norm_out.weight");
    PD_GETW(dec_no_b, "decoder.conv_norm_out.bias");
    {   // conv_out (C0 -> 1): fp32 [9][C0] (tap-major) for the direct kernel
        const int c0 = boc[0];
        const float* w;
        PD_GETW(w, "decoder.conv_out.weight");
        std::vector<float> hw((size_t)c0 * 9), tw((size_t)c0 * 9);
        PD_CUDA(cudaMemcpy(hw.data(), w, hw.size() * sizeof(float), cudaMemcpyDeviceToHost));
        for (int c = 0; c < c0; ++c)
            for (int t = 0; t < 9; ++t) tw[(size_t)t * c0 + c] = hw[(size_t)c * 9 + t];
        PD_TRY(dec_out_w.alloc(tw.size() * sizeof(float)));
        PD_CUDA(cudaMemcpy(dec_out_w.p, tw.data(), tw.size() * sizeof(float), cudaMemcpyHostToDevice));
        const float* b;
        PD_GETW(b, "decoder.conv_out.bias");
        PD_CUDA(cudaMemcpy(&dec_out_b, b, sizeof(float), cudaMemcpyDeviceToHost));
    }
    PD_CUDA(cudaDeviceSynchronize());
    finalized = true;
    return PD_OK;
}

int VAE::add_gn(Ctx& c, const float* x, const float* w, const float* b, bf16* y, int R, int C, int silu) {
    double* s = c.next_sums();
    const int N = c.N, G = c.G;
    c.pl->add([=](cudaStream_t st) { return gn_stats(x, s, N, R, C, G, st); });
    c.pl->add([=](cudaStream_t st) { return gn_apply(x, s, w, b, y, N, R, C, G, 1e-6f, silu, st); });
    return PD_OK;
}

int VAE::add_resnet(Ctx& c, const Res2W& r, int h, int w) {
    float* x = c.stream();
    float* hbuf = c.other(1);
    PD_TRY(add_gn(c, x, r.gn1_w, r.gn1_b, c.a, h * w, r.cin, 1));
    {
        GemmEpilogue e;
        e.bias = r.conv1_b;
        e.out_f32 = hbuf;
        GemmOp op;
        PD_TRY(gemm_make(&op, c.a, GemmGeom::conv(c.N, 1, h, w, r.cin, 1, 3, 3), r.conv1_w, r.cout, e));
        c.pl->add_gemm(op);
    }
    PD_TRY(add_gn(c, hbuf, r.gn2_w, r.gn2_b, c.a, h * w, r.cout, 1));
    float* res = x;
    float* out = x;
    if (r.sc_w) {  // 1x1 conv_shortcut on the raw input (resnet.py:490-491)
        float* sc = c.other(2);
        const int N = c.N;
        bf16* cast = c.cast;
        const int64_t rc = (int64_t)h * w * r.cin;
        c.pl->add([=](cudaStream_t st) { return cast_bf16(x, cast, N, rc, rc, st); });
        GemmEpilogue e;
        e.bias = r.sc_b;
        e.out_f32 = sc;
        GemmOp op;
        PD_TRY(gemm_make(&op, cast, GemmGeom::conv(c.N, 1, h, w, r.cin, 1, 1, 1), r.sc_w, r.cout, e));
        c.pl->add_gemm(op);
        res = sc;
        out = sc;
        c.cur = (c.cur + 2) % 3;
    }
    GemmEpilogue e;
    e.bias = r.conv2_b;
    e.residual = res;
    e.out_f32 = out;
    GemmOp op;
    PD_TRY(gemm_make(&op, c.a, GemmGeom::conv(c.N, 1, h, w, r.cout, 1, 3, 3), r.conv2_w, r.cout, e));
    c.pl->add_gemm(op);
    return PD_OK;
}

int VAE::add_mid(Ctx& c, const MidW& m, int h, int w, int ch) {
    PD_TRY(add_resnet(c, m.r0, h, w));
    // ---- AttentionBlock, single head over the h*w tokens (attention.py:136-189) ----
    const int tok = h * w, N = c.N;
    float* x = c.stream();
    float* sbuf = c.other(1);
    PD_TRY(add_gn(c, x, m.gn_w, m.gn_b, c.a, tok, ch, 0));
    {
        GemmEpilogue e;
        e.bias = m.qkv_b;
        e.out_bf16 = c.qkv;
        GemmOp op;
        PD_TRY(gemm_make(&op, c.a, GemmGeom::linear(N * tok, ch), m.qkv_w, 3 * ch, e));
        c.pl->add_gemm(op);
    }
    {   // S = Q K^T per frame: A = Q view of qkv, B = K view (per-sample operand)
        GemmGeom g;
        g.samples = N; g.W = tok; g.C = ch;
        g.sW = 3 * ch; g.sH = (int64_t)3 * ch * tok; g.sD = g.sH; g.sN = g.sH;
        g.ldb = 3 * ch; g.b_sample_stride = (int64_t)3 * ch * tok;
        GemmEpilogue e;
        e.out_f32 = sbuf;
        GemmOp op;
        PD_TRY(gemm_make(&op, c.qkv, g, c.qkv + ch, tok, e));
        c.pl->add_gemm(op);
    }
    {
        bf16* p = c.a;
        const float scale = 1.0f / sqrtf((float)ch);
        c.pl->add([=](cudaStream_t st) { return softmax_rows(sbuf, p, N * tok, tok, scale, st); });
        const bf16* v = c.qkv + 2 * ch;
        bf16* vt = c.cast;
        c.pl->add([=](cudaStream_t st) { return transpose_bf16(v, vt, N, tok, ch, 3 * ch, st); });
    }
    {   // O = P V per frame: B = V^T [ch][tok]
        GemmGeom g;
        g.samples = N; g.W = tok; g.C = tok;
        g.b_sample_stride = (int64_t)ch * tok;
        GemmEpilogue e;
        e.out_bf16 = c.qkv;  // q/k/v are dead (V^T lives in `cast`)
        GemmOp op;
        PD_TRY(gemm_make(&op, c.a, g, c.cast, ch, e));
        c.pl->add_gemm(op);
    }
    {
        GemmEpilogue e;
        e.bias = m.proj_b;
        e.residual = x;
        e.out_f32 = x;
        GemmOp op;
        PD_TRY(gemm_make(&op, c.qkv, GemmGeom::linear(N * tok, ch), m.proj_w, ch, e));
        c.pl->add_gemm(op);
    }
    return add_resnet(c, m.r1, h, w);
}

int64_t VAE::max_elems(int N) const {
    // largest [N][h][w][c] tensor of either direction: 2x-upsampled decoder tensors dominate
    const int* boc = cfg.block_out_channels;
    int64_t m = 0;
    int h = cfg.h, w = cfg.w;
    for (int i = 0; i < 4; ++i) {
        const int cmax = boc[i] > boc[i < 3 ? i + 1 : 3] ? boc[i] : boc[i < 3 ? i + 1 : 3];
        m = std::max<int64_t>(m, (int64_t)N * h * w * cmax);
        h /= 2; w /= 2;
    }
    return m;
}

template <class A>
void VAE::carve(A& ar, int N, Ctx* c, int n_gn) const {
    const int64_t me = max_elems(N);
    const int tok = (cfg.h / 8) * (cfg.w / 8), ch = cfg.block_out_channels[3];
    for (int i = 0; i < 3; ++i) c->f[i] = ar.template take<float>((size_t)std::max<int64_t>(me, (int64_t)N * tok * tok));
    c->a = ar.template take<bf16>((size_t)std::max<int64_t>(me, (int64_t)N * tok * tok));
    c->cast = ar.template take<bf16>((size_t)me);
    c->qkv = ar.template take<bf16>((size_t)N * tok * 3 * ch);
    c->gn_sums = ar.template take<double>((size_t)n_gn * N * 128 * 2);
}

int VAE::build(int N, bool encode, DirPlan* dp) {
    const int* boc = cfg.block_out_channels;
    const int L = cfg.latent_channels;
    const int n_gn = encode ? 2 * cfg.layers_per_block * 4 + 5 + 1 : 2 * (cfg.layers_per_block + 1) * 4 + 5 + 1;
    ArenaSizer sz;
    Ctx tmp;
    carve(sz, N, &tmp, n_gn);
    PD_TRY(dp->arena.reserve(sz.used() + 4096));
    Ctx& c = dp->ctx_storage();
    carve(dp->arena, N, &c, n_gn);
    PD_CHECK(!dp->arena.overflowed(), PD_ERR_STATE, "vae: arena overflow");
    c.pl = &dp->plan;
    c.N = N;
    c.G = cfg.norm_num_groups;
    Plan& pl = dp->plan;
    {
        double* gs = c.gn_sums;
        const size_t bytes = (size_t)n_gn * N * 128 * 2 * sizeof(double);
        pl.add([=](cudaStream_t st) {
            PD_CUDA(cudaMemsetAsync(gs, 0, bytes, st));
            return PD_OK;
        }, STEP_NONE);
    }
    int h = cfg.h, w = cfg.w;
    if (encode) {
        dp->in_slot = pl.steps.size();
        pl.add([](cudaStream_t) { return PD_OK; });  // conv_in, bound per call (user input pointer)
        dp->first_buf = c.stream();
        size_t ri = 0;
        for (int i = 0; i < 4; ++i) {
            for (int j = 0; j < cfg.layers_per_block; ++j) PD_TRY(add_resnet(c, enc_res[ri++], h, w));
            if (i != 3) {  // Downsample2D: pad (0,1,0,1) + 3x3 stride 2 (resnet.py:183-188)
                const float* x = c.stream();
                bf16* planes = c.cast;
                const int N_ = N, hh = h, ww = w, ch = boc[i];
                pl.add([=](cudaStream_t st) { return parity_split_cast(x, planes, N_, hh, ww, ch, st); });
                GemmEpilogue e;
                e.bias = down_b[i];
                e.out_f32 = c.other(1);
                GemmOp op;
                PD_TRY(gemm_make(&op, planes, GemmGeom::conv_s2_planes(N, h / 2, w / 2, boc[i]), down_w[i], boc[i], e));
                pl.add_gemm(op);
                c.cur = (c.cur + 1) % 3;
                h /= 2; w /= 2;
            }
        }
        PD_TRY(add_mid(c, enc_mid, h, w, boc[3]));
        PD_TRY(add_gn(c, c.stream(), enc_no_w, enc_no_b, c.a, h * w, boc[3], 1));
        {
            GemmEpilogue e;
            e.bias = enc_out_b;
            e.out_bf16 = c.cast;  // feeds quant_conv directly
            GemmOp op;
            PD_TRY(gemm_make(&op, c.a, GemmGeom::conv(N, 1, h, w, boc[3], 1, 3, 3), enc_out_w, 2 * L, e));
            pl.add_gemm(op);
        }
        {
            GemmEpilogue e;
            e.bias = quant_b;
            e.out_f32 = c.stream();  // placeholder; re-bound to the caller's tensor per call
            PD_TRY(gemm_make(&dp->last_op, c.cast, GemmGeom::conv(N, 1, h, w, 2 * L, 1, 1, 1), quant_w, 2 * L, e));
            pl.gemm_flops += dp->last_op.flops;
            dp->out_slot = pl.steps.size();
            pl.add([](cudaStream_t) { return PD_OK; });
        }
    } else {
        h = cfg.h / 8; w = cfg.w / 8;
        dp->in_slot = pl.steps.size();
        pl.add([](cudaStream_t) { return PD_OK; });  // cast of z, bound per call
        {
            GemmEpilogue e;
            e.bias = pquant_b;
            e.out_bf16 = c.a;
            GemmOp op;
            PD_TRY(gemm_make(&op, c.cast, GemmGeom::conv(N, 1, h, w, L, 1, 1, 1), pquant_w, L, e));
            pl.add_gemm(op);
        }
        {
            GemmEpilogue e;
            e.bias = dec_in_b;
            e.out_f32 = c.stream();
            GemmOp op;
            PD_TRY(gemm_make(&op, c.a, GemmGeom::conv(N, 1, h, w, L, 1, 3, 3), dec_in_w, boc[3], e));
            pl.add_gemm(op);
        }
        PD_TRY(add_mid(c, dec_mid, h, w, boc[3]));
        size_t ri = 0;
        for (int i = 0; i < 4; ++i) {
            const int cout = boc[3 - i];
            for (int j = 0; j < cfg.layers_per_block + 1; ++j) PD_TRY(add_resnet(c, dec_res[ri++], h, w));
            if (i != 3) {  // Upsample2D: nearest 2x + conv 3x3 (resnet.py:128,137-139)
                const float* x = c.stream();
                bf16* up = c.cast;
                const int N_ = N, hh = h, ww = w;
                pl.add([=](cudaStream_t st) { return upsample2x_cast(x, up, N_, hh, ww, cout, st); });
                h *= 2; w *= 2;
                GemmEpilogue e;
                e.bias = up_b[i];
                e.out_f32 = c.other(1);
                GemmOp op;
                PD_TRY(gemm_make(&op, up, GemmGeom::conv(N, 1, h, w, cout, 1, 3, 3), up_w[i], cout, e));
                pl.add_gemm(op);
                c.cur = (c.cur + 1) % 3;
            }
        }
        PD_TRY(add_gn(c, c.stream(), dec_no_w, dec_no_b, c.a, h * w, boc[0], 1));
        dp->out_slot = pl.steps.size();
        pl.add([](cudaStream_t) { return PD_OK; });  // conv_out, bound per call (user output pointer)
    }
    PD_CHECK(c.gn_slot == n_gn, PD_ERR_STATE, "vae: gn slot accounting (%d vs %d)", c.gn_slot, n_gn);
    return PD_OK;
}

int VAE::get_plan(int N, bool encode, DirPlan** out) {
    PD_CHECK(finalized, PD_ERR_STATE, "vae: call pd_vae_finalize() after loading all weights");
    PD_CHECK(N >= 1 && N <= cfg.max_frames, PD_ERR_SHAPE, "vae: %d frames outside [1, max_frames=%d]", N, cfg.max_frames);
    auto& plans = encode ? enc_plans : dec_plans;
    auto it = plans.find(N);
    if (it == plans.end()) {
        std::unique_ptr<DirPlan> dp(new DirPlan());
        PD_TRY(build(N, encode, dp.get()));
        it = plans.emplace(N, std::move(dp)).first;
    }
    *out = it->second.get();
    return PD_OK;
}

int VAE::encode(const float* x, float* moments, int N, cudaStream_t st) {
    PD_CHECK(x && moments, PD_ERR_ARG, "vae encode: null pointer");
    DirPlan* dp = nullptr;
    PD_TRY(get_plan(N, true, &dp));
    const int H = cfg.h, W = cfg.w, c0 = cfg.block_out_channels[0];
    const float *w = enc_in_w, *b = enc_in_b;
    float* first = dp->first_buf;
    dp->plan.steps[dp->in_slot] = [=](cudaStream_t s) { return conv3x3_c1_in(x, w, b, first, N, H, W, c0, s); };
    GemmOp op = dp->last_op;
    PD_TRY(gemm_bind_output(&op, moments, nullptr, nullptr));
    dp->plan.steps[dp->out_slot] = [op](cudaStream_t s) { return gemm_launch(op, s); };
    return dp->plan.run(st);
}

int VAE::decode(const float* z, float* out, int N, cudaStream_t st) {
    PD_CHECK(z && out, PD_ERR_ARG, "vae decode: null pointer");
    DirPlan* dp = nullptr;
    PD_TRY(get_plan(N, false, &dp));
    const Ctx& c = dp->ctx_storage();
    const int64_t rc = (int64_t)(cfg.h / 8) * (cfg.w / 8) * cfg.latent_channels;
    bf16* cast = c.cast;
    dp->plan.steps[dp->in_slot] = [=](cudaStream_t s) { return cast_bf16(z, cast, N, rc, rc, s); };
    const bf16* a = c.a;
    const float* w = dec_out_w.as<float>();
    const float bias = dec_out_b;
    const int H = cfg.h, W = cfg.w, c0 = cfg.block_out_channels[0];
    dp->plan.steps[dp->out_slot] = [=](cudaStream_t s) { return conv3x3_c1_out(a, w, bias, out, N, H, W, c0, s); };
    return dp->plan.run(st);
}

}  // namespace pd
