// Knowledge-alignment network: forward + hand-written input-gradient backward as static launch plans (see ka.cuh).
// Reference: knowledge_alignment/models.py:459-528 (NoisyCuboidTransformerEncoder.forward), :49-104
// (AttentionPool3d), sevir.py:55-104 (alignment_fn / get_mean_shift), alignment_pl.py:423-446 (autograd.grad).
// The trunk is the same kernel set as the UNet encoder half (gemm.cu / norm.cu / attention.cu); the backward
// runs every GEMM as a dgrad with transposed (Linear) or tap-reversed (Conv3d) weights packed once at finalize.
#include <cstdlib>
#include "ka.cuh"

namespace pd {

struct KABlock {       // one (TimeEmbedResBlock, StackCuboidSelfAttentionBlock) pair and what its backward needs
    float *x_in, *h, *xs[7];
    bf16* qkv[3];
    float* pre[3];
    double *st1, *st2, *bst1, *bst2;
};

struct KANet::Bufs {
    float *z, *hf, *x_first, *x1a, *qkvp, *wsave, *outv, *dout, *loss, *dtok, *gout;
    float *g32[2], *dx[2];
    bf16 *zb, *a[2], *ln[2], *att[2], *mid[2], *dbig[2], *dxb[2], *pm, *tok, *dqkvp;
    float *e0, *e1, *temb, *embs;
    double *gn_sums, *st_f1, *st_f2, *bst_f1, *bst_f2, *st_out, *bst_out;
    std::vector<KABlock> blk[2];
};

struct KANet::BatchPlan {
    Arena arena;
    Bufs bufs;
    Plan fwd, bwd;
    size_t in_slot = 0, t_slot = 0, loss_slot = 0;
    size_t gn_bytes = 0;
};

KANet::~KANet() = default;

KANet::KANet(const pd_ka_config& c) : cfg(c) {
    C0 = cfg.base_units;
    C1 = 2 * cfg.base_units;
    T = cfg.t;
    TE = 4 * cfg.base_units;
    declare_weights();
}

void KANet::declare_resblock(const std::string& p, int cin, int cout, bool emb) {
    ws.declare(p + ".in_layers.0.weight", {cin});
    ws.declare(p + ".in_layers.0.bias", {cin});
    ws.declare(p + ".in_layers.2.weight", {cout, cin, 3, 3, 3});
    ws.declare(p + ".in_layers.2.bias", {cout});
    if (emb) {
        ws.declare(p + ".emb_layers.1.weight", {cout, TE});
        ws.declare(p + ".emb_layers.1.bias", {cout});
    }
    ws.declare(p + ".out_layers.0.weight", {cout});
    ws.declare(p + ".out_layers.0.bias", {cout});
    ws.declare(p + ".out_layers.3.weight", {cout, cout, 3, 3, 3});
    ws.declare(p + ".out_layers.3.bias", {cout});
    if (cin != cout) {
        ws.declare(p + ".skip_connection.weight", {cout, cin, 1, 1, 1});
        ws.declare(p + ".skip_connection.bias", {cout});
    }
}

void KANet::declare_stack(const std::string& p, int dim, int lvl) {
    const int L[3] = {T, cfg.h >> lvl, cfg.w >> lvl};
    for (int i = 0; i < 3; ++i) {
        const std::string f = p + strf(".ffn_l.%d", i);
        ws.declare(f + ".ffn_1.weight", {4 * dim, dim});
        ws.declare(f + ".ffn_1.bias", {4 * dim});
        ws.declare(f + ".ffn_2.weight", {dim, 4 * dim});
        ws.declare(f + ".ffn_2.bias", {dim});
        ws.declare(f + ".layer_norm.weight", {dim});
        ws.declare(f + ".layer_norm.bias", {dim});
    }
    for (int i = 0; i < 3; ++i) {
        const std::string a = p + strf(".attn_l.%d", i);
        ws.declare(a + ".relative_position_bias_table", {2 * L[i] - 1, cfg.num_heads});
        ws.declare(a + ".qkv.weight", {3 * dim, dim});
        ws.declare(a + ".proj.weight", {dim, dim});
        ws.declare(a + ".proj.bias", {dim});
        ws.declare(a + ".norm.weight", {dim});
        ws.declare(a + ".norm.bias", {dim});
    }
}

// Same names, shapes and order as the reference module's state_dict() (minus derived int64 buffers).
void KANet::declare_weights() {
    declare_resblock("first_proj", cfg.c, C0, false);
    ws.declare("pos_embed.T_embed.weight", {T, C0});
    ws.declare("pos_embed.H_embed.weight", {cfg.h, C0});
    ws.declare("pos_embed.W_embed.weight", {cfg.w, C0});
    ws.declare("time_embed.layer.0.weight", {TE, C0});
    ws.declare("time_embed.layer.0.bias", {TE});
    ws.declare("time_embed.layer.2.weight", {TE, TE});
    ws.declare("time_embed.layer.2.bias", {TE});
    ws.declare("downsample_layers.0.reduction.weight", {C1, 4 * C0});
    ws.declare("downsample_layers.0.norm.weight", {4 * C0});
    ws.declare("downsample_layers.0.norm.bias", {4 * C0});
    for (int lvl = 0; lvl < 2; ++lvl)
        for (int d = 0; d < cfg.depth[lvl]; ++d) declare_stack(strf("down_self_blocks.%d.%d", lvl, d), lvl ? C1 : C0, lvl);
    for (int lvl = 0; lvl < 2; ++lvl) declare_resblock(strf("down_time_embed_blocks.%d", lvl), lvl ? C1 : C0, lvl ? C1 : C0, true);
    const int tokens = (cfg.h / 2) * (cfg.w / 2) + 1;
    ws.declare("out.0.weight", {C1});
    ws.declare("out.0.bias", {C1});
    ws.declare("out.2.positional_embedding", {C1, tokens});
    ws.declare("out.2.qkv_proj.weight", {3 * C1, C1, 1});
    ws.declare("out.2.qkv_proj.bias", {3 * C1});
    ws.declare("out.2.c_proj.weight", {1, C1, 1});
    ws.declare("out.2.c_proj.bias", {1});
}

int KANet::validate() const {
    PD_CHECK(T >= 1 && T <= 16, PD_ERR_SHAPE, "ka: T = %d must be in [1, 16]", T);
    PD_CHECK(cfg.h <= 16 && cfg.w <= 16 && cfg.h % 2 == 0 && cfg.w % 2 == 0, PD_ERR_SHAPE,
             "ka: latent H, W must be even and <= 16 (axial lines of <= 16 tokens)");
    for (int lvl = 0; lvl < 2; ++lvl) {
        const int hh = cfg.h >> lvl, ww = cfg.w >> lvl;
        PD_CHECK(128 % ww == 0 && ((hh * ww) % 128 == 0 || 128 % (hh * ww) == 0), PD_ERR_SHAPE,
                 "ka: H x W = %d x %d incompatible with 128-row tiles", hh, ww);
    }
    PD_CHECK(cfg.c % 64 == 0, PD_ERR_SHAPE, "ka: latent channels %d must be a multiple of 64", cfg.c);
    PD_CHECK(C0 % 128 == 0 && C0 <= 256, PD_ERR_SHAPE, "ka: base_units %d must be 128 or 256", C0);
    PD_CHECK(cfg.num_heads >= 1 && cfg.num_heads <= 8 && C0 % cfg.num_heads == 0, PD_ERR_SHAPE, "ka: heads");
    const int hd0 = C0 / cfg.num_heads;
    PD_CHECK(hd0 == 16 || hd0 == 32 || hd0 == 64, PD_ERR_SHAPE, "ka: head dim %d unsupported", hd0);
    PD_CHECK(C1 / cfg.num_heads <= 128 && (cfg.h / 2) * (cfg.w / 2) + 1 <= 96, PD_ERR_SHAPE, "ka: read-out head size");
    PD_CHECK(cfg.depth[0] >= 1 && cfg.depth[1] >= 1, PD_ERR_SHAPE, "ka: depth");
    PD_CHECK(cfg.max_batch >= 1 && cfg.max_batch <= 1024, PD_ERR_SHAPE, "ka: max_batch");
    return PD_OK;
}

// ---- weight repacking ---------------------------------------------------------------------------------------
int KANet::pack_conv_w(const std::string& name, int co, int ci, int taps, bf16** fwd, bf16** dgrad) {
    const float* w = ws.get(name);
    if (!w) return PD_ERR_WEIGHT;
    const size_t n = (size_t)co * taps * ci;
    packed.emplace_back(new DevMem());
    PD_TRY(packed.back()->alloc(n * sizeof(bf16)));
    *fwd = packed.back()->as<bf16>();
    PD_TRY(pack_conv(w, *fwd, co, ci, taps, ci, 0));
    packed.emplace_back(new DevMem());
    PD_TRY(packed.back()->alloc(n * sizeof(bf16)));
    *dgrad = packed.back()->as<bf16>();
    return pack_conv_dgrad(w, *dgrad, co, ci, taps, 0);
}
int KANet::pack_linear_w(const std::string& name, int n, int k, bf16** fwd, bf16** tr) {
    const float* w = ws.get(name);
    if (!w) return PD_ERR_WEIGHT;
    packed.emplace_back(new DevMem());
    PD_TRY(packed.back()->alloc((size_t)n * k * sizeof(bf16)));
    *fwd = packed.back()->as<bf16>();
    PD_TRY(pack_linear(w, *fwd, n, k, k, 0));
    packed.emplace_back(new DevMem());
    PD_TRY(packed.back()->alloc((size_t)n * k * sizeof(bf16)));
    *tr = packed.back()->as<bf16>();
    return pack_linear_t(w, *tr, n, k, 0);
}

#define PD_GETW(dst, name)                      \
    do {                                        \
        (dst) = ws.get(name);                   \
        if (!(dst)) return PD_ERR_WEIGHT;       \
    } while (0)

int KANet::finalize_resblock(const std::string& p, int cin, int cout, ResW* r, ResBwdW* rb) {
    PD_GETW(r->gn1_w, p + ".in_layers.0.weight");
    PD_GETW(r->gn1_b, p + ".in_layers.0.bias");
    PD_GETW(r->conv1_b, p + ".in_layers.2.bias");
    PD_GETW(r->gn2_w, p + ".out_layers.0.weight");
    PD_GETW(r->gn2_b, p + ".out_layers.0.bias");
    PD_GETW(r->conv2_b, p + ".out_layers.3.bias");
    PD_TRY(pack_conv_w(p + ".in_layers.2.weight", cout, cin, 27, &r->conv1_w, &rb->conv1_d));
    PD_TRY(pack_conv_w(p + ".out_layers.3.weight", cout, cout, 27, &r->conv2_w, &rb->conv2_d));
    return PD_OK;
}

int KANet::finalize_stack(const std::string& p, int dim, StackW* s, StackBwdW* sb) {
    s->a.assign(3, AttnW{});   // the KA network is built for the axial pattern only (three layers per stack block)
    s->f.assign(3, FfnW{});
    for (int i = 0; i < 3; ++i) {
        const std::string a = p + strf(".attn_l.%d", i), f = p + strf(".ffn_l.%d", i);
        PD_GETW(s->a[i].ln_w, a + ".norm.weight");
        PD_GETW(s->a[i].ln_b, a + ".norm.bias");
        PD_GETW(s->a[i].table, a + ".relative_position_bias_table");
        PD_GETW(s->a[i].proj_b, a + ".proj.bias");
        PD_TRY(pack_linear_w(a + ".qkv.weight", 3 * dim, dim, &s->a[i].qkv_w, &sb->qkv_t[i]));
        PD_TRY(pack_linear_w(a + ".proj.weight", dim, dim, &s->a[i].proj_w, &sb->proj_t[i]));
        PD_GETW(s->f[i].ln_w, f + ".layer_norm.weight");
        PD_GETW(s->f[i].ln_b, f + ".layer_norm.bias");
        PD_GETW(s->f[i].b1, f + ".ffn_1.bias");
        PD_GETW(s->f[i].b2, f + ".ffn_2.bias");
        PD_TRY(pack_linear_w(f + ".ffn_1.weight", 4 * dim, dim, &s->f[i].w1, &sb->w1_t[i]));
        PD_TRY(pack_linear_w(f + ".ffn_2.weight", dim, 4 * dim, &s->f[i].w2, &sb->w2_t[i]));
    }
    return PD_OK;
}

int KANet::finalize() {
    PD_TRY(gemm_init());
    PD_TRY(validate());
    PD_TRY(ws.check_complete());
    ++generation;
    packed.clear();
    plans.clear();
    PD_TRY(finalize_resblock("first_proj", cfg.c, C0, &first, &first_b));
    PD_TRY(pack_linear_w("first_proj.skip_connection.weight", C0, cfg.c, &skip_w, &skip_t));
    PD_GETW(skip_b, "first_proj.skip_connection.bias");
    PD_GETW(pos_T, "pos_embed.T_embed.weight");
    PD_GETW(pos_H, "pos_embed.H_embed.weight");
    PD_GETW(pos_W, "pos_embed.W_embed.weight");
    PD_GETW(te_w0, "time_embed.layer.0.weight");
    PD_GETW(te_b0, "time_embed.layer.0.bias");
    PD_GETW(te_w2, "time_embed.layer.2.weight");
    PD_GETW(te_b2, "time_embed.layer.2.bias");
    emb_total = C0 + C1;
    PD_TRY(emb_cat.alloc(((size_t)emb_total * TE + emb_total) * sizeof(float)));
    {
        const int dims[2] = {C0, C1};
        int off = 0;
        float* wcat = emb_cat.as<float>();
        float* bcat = wcat + (size_t)emb_total * TE;
        for (int i = 0; i < 2; ++i) {
            const float *w, *b;
            PD_GETW(w, strf("down_time_embed_blocks.%d.emb_layers.1.weight", i));
            PD_GETW(b, strf("down_time_embed_blocks.%d.emb_layers.1.bias", i));
            PD_CUDA(cudaMemcpy(wcat + (size_t)off * TE, w, (size_t)dims[i] * TE * sizeof(float), cudaMemcpyDeviceToDevice));
            PD_CUDA(cudaMemcpy(bcat + off, b, dims[i] * sizeof(float), cudaMemcpyDeviceToDevice));
            emb_off[i] = off;
            off += dims[i];
        }
    }
    for (int lvl = 0; lvl < 2; ++lvl) {
        const int dim = lvl ? C1 : C0;
        PD_TRY(finalize_resblock(strf("down_time_embed_blocks.%d", lvl), dim, dim, &res[lvl], &res_b[lvl]));
        stack[lvl].resize(cfg.depth[lvl]);
        stack_b[lvl].resize(cfg.depth[lvl]);
        for (int d = 0; d < cfg.depth[lvl]; ++d)
            PD_TRY(finalize_stack(strf("down_self_blocks.%d.%d", lvl, d), dim, &stack[lvl][d], &stack_b[lvl][d]));
    }
    PD_GETW(pm_ln_w, "downsample_layers.0.norm.weight");
    PD_GETW(pm_ln_b, "downsample_layers.0.norm.bias");
    PD_TRY(pack_linear_w("downsample_layers.0.reduction.weight", C1, 4 * C0, &pm_w, &pm_t));
    PD_GETW(out_gn_w, "out.0.weight");
    PD_GETW(out_gn_b, "out.0.bias");
    PD_TRY(pack_linear_w("out.2.qkv_proj.weight", 3 * C1, C1, &hq_w, &hq_t));
    PD_GETW(hq_b, "out.2.qkv_proj.bias");
    PD_GETW(cproj_w, "out.2.c_proj.weight");
    {
        const float* cb;
        PD_GETW(cb, "out.2.c_proj.bias");
        PD_CUDA(cudaMemcpy(&cproj_b, cb, sizeof(float), cudaMemcpyDeviceToHost));
        const float* pe;
        PD_GETW(pe, "out.2.positional_embedding");
        const int tokens = (cfg.h / 2) * (cfg.w / 2) + 1;
        PD_TRY(head_pos.alloc((size_t)tokens * C1 * sizeof(float)));
        PD_TRY(transpose_f32(pe, head_pos.as<float>(), C1, tokens, 0));   // [C1][tokens] -> [tokens][C1]
    }
    PD_CUDA(cudaDeviceSynchronize());
    finalized = true;
    return PD_OK;
}

// ---- plan construction --------------------------------------------------------------------------------------
template <class A>
void KANet::carve(A& ar, int B, Bufs* b) const {
    const size_t P0 = (size_t)B * T * cfg.h * cfg.w, P1 = P0 / 4;
    const size_t P[2] = {P0, P1};
    const int C[2] = {C0, C1};
    const size_t F = (size_t)B * T, L = (size_t)(cfg.h / 2) * (cfg.w / 2) + 1;
    const int cin = cfg.c;
    b->z = ar.template take<float>(P0 * cin);
    b->zb = ar.template take<bf16>(P0 * cin);
    b->hf = ar.template take<float>(P0 * C0);
    b->x_first = ar.template take<float>(P0 * C0);
    b->x1a = ar.template take<float>(P1 * C1);
    b->gout = ar.template take<float>(P0 * cin);
    for (int l = 0; l < 2; ++l) {
        const size_t wide = (size_t)(l == 1 ? (C1 > 4 * C0 ? C1 : 4 * C0) : (C0 > cin ? C0 : cin));
        b->a[l] = ar.template take<bf16>(P[l] * (size_t)(l == 0 && cin > C0 ? cin : C[l]));
        b->ln[l] = ar.template take<bf16>(P[l] * C[l]);
        b->att[l] = ar.template take<bf16>(P[l] * C[l]);
        b->mid[l] = ar.template take<bf16>(P[l] * 4 * C[l]);
        b->dbig[l] = ar.template take<bf16>(P[l] * 4 * C[l]);
        b->g32[l] = ar.template take<float>(P[l] * wide);
        b->dx[l] = ar.template take<float>(P[l] * C[l]);
        b->dxb[l] = ar.template take<bf16>(P[l] * C[l]);
    }
    b->pm = ar.template take<bf16>(P1 * 4 * C0);
    b->tok = ar.template take<bf16>(F * L * C1);
    b->qkvp = ar.template take<float>(F * L * 3 * C1);
    b->dqkvp = ar.template take<bf16>(F * L * 3 * C1);
    b->dtok = ar.template take<float>(F * L * C1);
    b->wsave = ar.template take<float>(F * cfg.num_heads * L);
    b->outv = ar.template take<float>(F);
    b->dout = ar.template take<float>(F);
    b->loss = ar.template take<float>(4);
    b->e0 = ar.template take<float>((size_t)B * C0);
    b->e1 = ar.template take<float>((size_t)B * TE);
    b->temb = ar.template take<float>((size_t)B * TE);
    b->embs = ar.template take<float>((size_t)B * emb_total);
    // GroupNorm statistics: forward slots then backward slots, one memset per call
    const size_t slot = F * 128 * 2;   // doubles; covers [B][G<=128][2] and the read-out's [B*T][32][2]
    b->gn_sums = ar.template take<double>((size_t)num_gn_slots() * 2 * slot);
    double* s = b->gn_sums;
    auto next = [&]() { double* r = s; if (s) s += slot; return r; };
    b->st_f1 = next(); b->st_f2 = next(); b->bst_f1 = next(); b->bst_f2 = next();
    b->st_out = next(); b->bst_out = next();
    for (int l = 0; l < 2; ++l) {
        b->blk[l].resize(cfg.depth[l]);
        for (int d = 0; d < cfg.depth[l]; ++d) {
            KABlock& k = b->blk[l][d];
            k.st1 = next(); k.st2 = next(); k.bst1 = next(); k.bst2 = next();
            k.h = ar.template take<float>(P[l] * C[l]);
            for (int i = 0; i < 7; ++i) k.xs[i] = ar.template take<float>(P[l] * C[l]);
            for (int i = 0; i < 3; ++i) {
                k.qkv[i] = ar.template take<bf16>(P[l] * 3 * C[l]);
                k.pre[i] = ar.template take<float>(P[l] * 4 * C[l]);
            }
            k.x_in = nullptr;
        }
    }
}

int KANet::build_plan(int B, BatchPlan* bp) {
    // wide GEMM tiles: the guidance runs beside the UNet and should hold as few SMs as it can (gemm.cu)
    struct TilePref {
        TilePref() { gemm_set_tile_preference(getenv("PD_KA_NARROW_TILES") ? 0 : 1); }
        ~TilePref() { gemm_set_tile_preference(0); }
    } tile_pref;
    ArenaSizer sz;
    {
        Bufs tmp;
        carve(sz, B, &tmp);
    }
    PD_TRY(bp->arena.reserve(sz.used() + 4096));
    Bufs& b = bp->bufs;
    carve(bp->arena, B, &b);
    PD_CHECK(!bp->arena.overflowed(), PD_ERR_STATE, "ka: arena overflow");
    const int H = cfg.h, W = cfg.w, HW = H * W, cin = cfg.c, heads = cfg.num_heads, Tn = T;
    const int R0 = T * HW;
    const int gin = cin % 32 == 0 ? 32 : cin;   // time_embed.py:90
    const int F = B * T, Rh = HW / 4, Ltok = Rh + 1;
    const int Cs[2] = {C0, C1};
    bp->gn_bytes = (size_t)num_gn_slots() * 2 * ((size_t)F * 128 * 2) * sizeof(double);

    // ================================================= forward =================================================
    Plan& pl = bp->fwd;
    {
        double* all = b.gn_sums;
        const size_t bytes = bp->gn_bytes;
        pl.add([=](cudaStream_t st) {
            PD_CUDA(cudaMemsetAsync(all, 0, bytes, st));
            return PD_OK;
        }, STEP_NONE);
    }
    bp->in_slot = pl.steps.size();
    pl.add([](cudaStream_t) { return PD_OK; }, STEP_NONE);   // placeholder: z_t -> arena copy, bound per call
    {   // time embedding (models/utils.py:68-83, time_embed.py:16-24,108-114)
        const float *w0 = te_w0, *b0 = te_b0, *w2 = te_w2, *b2 = te_b2;
        const float* wc = emb_cat.as<float>();
        const float* bc = wc + (size_t)emb_total * TE;
        const int c0 = C0, te = TE, et = emb_total;
        float *e0 = b.e0, *e1 = b.e1, *temb = b.temb, *embs = b.embs;
        bp->t_slot = pl.steps.size();
        pl.add([](cudaStream_t) { return PD_OK; });   // placeholder: timestep_embedding, bound per call
        pl.add([=](cudaStream_t st) { return small_linear(e0, w0, b0, e1, B, c0, te, 0, 1, st); });
        pl.add([=](cudaStream_t st) { return small_linear(e1, w2, b2, temb, B, te, te, 0, 0, st); });
        pl.add([=](cudaStream_t st) { return small_linear(temb, wc, bc, embs, B, te, et, 1, 0, st); });
    }
    {   // first_proj: TimeEmbedResBlock(cin -> C0) without embedding, 1x1x1 skip (models.py first_proj; time_embed.py:134-169)
        const float* z = b.z;
        bf16 *zb = b.zb, *a = b.a[0];
        float *hf = b.hf, *x = b.x_first;
        double *s1 = b.st_f1, *s2 = b.st_f2;
        const ResW r = first;
        const int c0 = C0;
        pl.add([=](cudaStream_t st) { return gn_stats(z, s1, B, R0, cin, gin, st); });
        pl.add([=](cudaStream_t st) { return gn_apply(z, s1, r.gn1_w, r.gn1_b, a, B, R0, cin, gin, 1e-5f, 1, st); });
        {
            GemmEpilogue e;
            e.bias = r.conv1_b;
            e.out_f32 = hf;
            GemmOp op;
            PD_TRY(gemm_make(&op, a, GemmGeom::conv(B, T, H, W, cin, 3, 3, 3), r.conv1_w, C0, e));
            pl.add_gemm(op);
        }
        pl.add([=](cudaStream_t st) { return cast_bf16(z, zb, 1, (int64_t)B * R0 * cin, 0, st); });
        {
            GemmEpilogue e;
            e.bias = skip_b;
            e.out_f32 = x;
            GemmOp op;
            PD_TRY(gemm_make(&op, zb, GemmGeom::conv(B, T, H, W, cin, 1, 1, 1), skip_w, C0, e));
            pl.add_gemm(op);
        }
        pl.add([=](cudaStream_t st) { return gn_stats(hf, s2, B, R0, c0, 32, st); });
        pl.add([=](cudaStream_t st) { return gn_apply(hf, s2, r.gn2_w, r.gn2_b, a, B, R0, c0, 32, 1e-5f, 1, st); });
        {
            GemmEpilogue e;
            e.bias = r.conv2_b;
            e.residual = x;
            e.out_f32 = x;
            GemmOp op;
            PD_TRY(gemm_make(&op, a, GemmGeom::conv(B, T, H, W, C0, 3, 3, 3), r.conv2_w, C0, e));
            pl.add_gemm(op);
        }
        const float *pt = pos_T, *ph = pos_H, *pw = pos_W;
        pl.add([=](cudaStream_t st) { return pos_embed_add(x, pt, ph, pw, B, Tn, H, W, c0, st); });
    }
    float* level_in = b.x_first;
    for (int l = 0; l < 2; ++l) {
        const int Hl = H >> l, Wl = W >> l, C = Cs[l];
        const int R = T * Hl * Wl, P = B * R;
        if (l == 1) {   // PatchMerging3D (cuboid_transformer.py:261-296)
            const float *x0 = level_in, *lw = pm_ln_w, *lb = pm_ln_b;
            bf16* pm = b.pm;
            const int BT = B * T, c0 = C0;
            pl.add([=](cudaStream_t st) { return patch_merge_ln(x0, lw, lb, pm, BT, H, W, c0, 1e-5f, st); });
            GemmEpilogue e;
            e.out_f32 = b.x1a;
            GemmOp op;
            PD_TRY(gemm_make(&op, pm, GemmGeom::linear(P, 4 * C0), pm_w, C1, e));
            pl.add_gemm(op);
            level_in = b.x1a;
        }
        bf16 *a = b.a[l], *ln = b.ln[l], *att = b.att[l], *mid = b.mid[l];
        for (int d = 0; d < cfg.depth[l]; ++d) {
            KABlock& k = b.blk[l][d];
            k.x_in = level_in;
            const ResW r = res[l];
            const StackW& s = stack[l][d];
            {   // TimeEmbedResBlock
                const float* xin = k.x_in;
                float* h = k.h;
                double *s1 = k.st1, *s2 = k.st2;
                pl.add([=](cudaStream_t st) { return gn_stats(xin, s1, B, R, C, 32, st); });
                pl.add([=](cudaStream_t st) { return gn_apply(xin, s1, r.gn1_w, r.gn1_b, a, B, R, C, 32, 1e-5f, 1, st); });
                {
                    GemmEpilogue e;
                    e.bias = r.conv1_b;
                    e.rowvec = b.embs + emb_off[l];
                    e.rowvec_ld = emb_total;
                    e.out_f32 = h;
                    GemmOp op;
                    PD_TRY(gemm_make(&op, a, GemmGeom::conv(B, T, Hl, Wl, C, 3, 3, 3), r.conv1_w, C, e));
                    pl.add_gemm(op);
                }
                pl.add([=](cudaStream_t st) { return gn_stats(h, s2, B, R, C, 32, st); });
                pl.add([=](cudaStream_t st) { return gn_apply(h, s2, r.gn2_w, r.gn2_b, a, B, R, C, 32, 1e-5f, 1, st); });
                {
                    GemmEpilogue e;
                    e.bias = r.conv2_b;
                    e.residual = xin;
                    e.out_f32 = k.xs[0];
                    GemmOp op;
                    PD_TRY(gemm_make(&op, a, GemmGeom::conv(B, T, Hl, Wl, C, 3, 3, 3), r.conv2_w, C, e));
                    pl.add_gemm(op);
                }
            }
            for (int i = 0; i < 3; ++i) {   // StackCuboidSelfAttentionBlock (cuboid_transformer.py:1147-1156)
                const AttnW aw = s.a[i];
                const FfnW fw = s.f[i];
                const float *xa = k.xs[2 * i], *xb = k.xs[2 * i + 1];
                bf16* qkv = k.qkv[i];
                float* pre = k.pre[i];
                pl.add([=](cudaStream_t st) { return layer_norm(xa, aw.ln_w, aw.ln_b, ln, P, C, 1e-5f, st); });
                {
                    GemmEpilogue e;
                    e.out_bf16 = qkv;
                    GemmOp op;
                    PD_TRY(gemm_make(&op, ln, GemmGeom::linear(P, C), aw.qkv_w, 3 * C, e));
                    pl.add_gemm(op);
                }
                pl.add([=](cudaStream_t st) { return axial_attention(qkv, aw.table, att, B, Tn, Hl, Wl, C, heads, i, st); });
                {
                    GemmEpilogue e;
                    e.bias = aw.proj_b;
                    e.residual = xa;
                    e.out_f32 = k.xs[2 * i + 1];
                    GemmOp op;
                    PD_TRY(gemm_make(&op, att, GemmGeom::linear(P, C), aw.proj_w, C, e));
                    pl.add_gemm(op);
                }
                pl.add([=](cudaStream_t st) { return layer_norm(xb, fw.ln_w, fw.ln_b, ln, P, C, 1e-5f, st); });
                {
                    GemmEpilogue e;
                    e.bias = fw.b1;
                    e.out_f32 = pre;   // pre-activation kept in fp32 for the backward
                    GemmOp op;
                    PD_TRY(gemm_make(&op, ln, GemmGeom::linear(P, C), fw.w1, 4 * C, e));
                    pl.add_gemm(op);
                }
                pl.add([=](cudaStream_t st) { return gelu_fwd(pre, mid, (int64_t)P * 4 * C, st); });
                {
                    GemmEpilogue e;
                    e.bias = fw.b2;
                    e.residual = xb;
                    e.out_f32 = k.xs[2 * i + 2];
                    GemmOp op;
                    PD_TRY(gemm_make(&op, mid, GemmGeom::linear(P, 4 * C), fw.w2, C, e));
                    pl.add_gemm(op);
                }
            }
            level_in = k.xs[6];
        }
    }
    {   // read-out (models.py:500-528): per-frame GN + SiLU, attention pool, token 0 -> c_proj
        const float* x1 = level_in;
        double* so = b.st_out;
        const float *gw = out_gn_w, *gb = out_gn_b, *pos = head_pos.as<float>(), *cw = cproj_w;
        const float cb = cproj_b;
        bf16* tok = b.tok;
        float *qkvp = b.qkvp, *wsave = b.wsave, *outv = b.outv;
        const int c1 = C1;
        pl.add([=](cudaStream_t st) { return gn_stats(x1, so, F, Rh, c1, 32, st); });
        pl.add([=](cudaStream_t st) { return ka_tokens(x1, so, gw, gb, pos, tok, F, Rh, c1, 32, 1e-5f, st); });
        {
            GemmEpilogue e;
            e.bias = hq_b;
            e.out_f32 = qkvp;
            GemmOp op;
            PD_TRY(gemm_make(&op, tok, GemmGeom::linear(F * Ltok, C1), hq_w, 3 * C1, e));
            pl.add_gemm(op);
        }
        pl.add([=](cudaStream_t st) { return ka_pool(qkvp, cw, cb, wsave, outv, F, Ltok, c1, heads, st); });
    }

    // ================================================= backward ================================================
    Plan& bw = bp->bwd;
    bp->loss_slot = bw.steps.size();
    bw.add([](cudaStream_t) { return PD_OK; });   // placeholder: ka_loss_grad (target pointer, guide scale)
    {
        const float *qkvp = b.qkvp, *cw = cproj_w, *wsave = b.wsave, *dout = b.dout;
        bf16* dqkvp = b.dqkvp;
        const int c1 = C1;
        bw.add([=](cudaStream_t st) { return ka_pool_bwd(qkvp, cw, wsave, dout, dqkvp, F, Ltok, c1, heads, st); });
        {
            GemmEpilogue e;
            e.out_f32 = b.dtok;
            GemmOp op;
            PD_TRY(gemm_make(&op, dqkvp, GemmGeom::linear(F * Ltok, 3 * C1), hq_t, C1, e));
            bw.add_gemm(op);
        }
        const float* dtok = b.dtok;
        float* dout_sp = b.g32[1];
        bw.add([=](cudaStream_t st) { return ka_tokens_bwd(dtok, dout_sp, F, Rh, c1, st); });
        const float* x1 = level_in;
        const double* so = b.st_out;
        double* bso = b.bst_out;
        const float *gw = out_gn_w, *gb = out_gn_b;
        float* dx = b.dx[1];
        bf16* dxb = b.dxb[1];
        bw.add([=](cudaStream_t st) { return gn_bwd(x1, dout_sp, so, bso, gw, gb, dx, dxb, F, Rh, c1, 32, 1e-5f, 1, 0, st); });
    }
    for (int l = 1; l >= 0; --l) {
        const int Hl = H >> l, Wl = W >> l, C = Cs[l];
        const int R = T * Hl * Wl, P = B * R;
        bf16 *ln = b.ln[l], *att = b.att[l], *mid = b.mid[l], *dbig = b.dbig[l], *dxb = b.dxb[l];
        float *g32 = b.g32[l], *dx = b.dx[l];
        for (int d = cfg.depth[l] - 1; d >= 0; --d) {
            const KABlock& k = b.blk[l][d];
            const StackW& s = stack[l][d];
            const StackBwdW& sb = stack_b[l][d];
            for (int i = 2; i >= 0; --i) {
                const AttnW aw = s.a[i];
                const FfnW fw = s.f[i];
                const float *xa = k.xs[2 * i], *xb = k.xs[2 * i + 1], *pre = k.pre[i];
                const bf16* qkv = k.qkv[i];
                // ---- FFN: x_out = x + W2 gelu(W1 LN(x) + b1) + b2 ----
                {
                    GemmEpilogue e;
                    e.out_bf16 = mid;   // d gelu-output
                    GemmOp op;
                    PD_TRY(gemm_make(&op, dxb, GemmGeom::linear(P, C), sb.w2_t[i], 4 * C, e));
                    bw.add_gemm(op);
                }
                bw.add([=](cudaStream_t st) { return gelu_bwd(pre, mid, dbig, (int64_t)P * 4 * C, st); });
                {
                    GemmEpilogue e;
                    e.out_f32 = g32;    // d LN-output
                    GemmOp op;
                    PD_TRY(gemm_make(&op, dbig, GemmGeom::linear(P, 4 * C), sb.w1_t[i], C, e));
                    bw.add_gemm(op);
                }
                bw.add([=](cudaStream_t st) { return layer_norm_bwd(xb, fw.ln_w, g32, dx, dxb, P, C, 1e-5f, 1, st); });
                // ---- attention: x_out = x + proj(attn(qkv(LN(x)))) ----
                {
                    GemmEpilogue e;
                    e.out_bf16 = att;   // d attention-output
                    GemmOp op;
                    PD_TRY(gemm_make(&op, dxb, GemmGeom::linear(P, C), sb.proj_t[i], C, e));
                    bw.add_gemm(op);
                }
                bw.add([=](cudaStream_t st) {
                    return axial_attention_bwd(qkv, aw.table, att, dbig, B, Tn, Hl, Wl, C, heads, i, st);
                });
                {
                    GemmEpilogue e;
                    e.out_f32 = g32;
                    GemmOp op;
                    PD_TRY(gemm_make(&op, dbig, GemmGeom::linear(P, 3 * C), sb.qkv_t[i], C, e));
                    bw.add_gemm(op);
                }
                bw.add([=](cudaStream_t st) { return layer_norm_bwd(xa, aw.ln_w, g32, dx, dxb, P, C, 1e-5f, 1, st); });
            }
            {   // ---- TimeEmbedResBlock: out = x + conv2(silu(GN2(conv1(silu(GN1(x))) + emb))) ----
                const ResW r = res[l];
                const ResBwdW rb = res_b[l];
                const float *xin = k.x_in, *h = k.h;
                const double *s1 = k.st1, *s2 = k.st2;
                double *b1 = k.bst1, *b2 = k.bst2;
                {
                    GemmEpilogue e;
                    e.out_f32 = g32;
                    GemmOp op;
                    PD_TRY(gemm_make(&op, dxb, GemmGeom::conv(B, T, Hl, Wl, C, 3, 3, 3), rb.conv2_d, C, e));
                    bw.add_gemm(op);
                }
                bw.add([=](cudaStream_t st) {
                    return gn_bwd(h, g32, s2, b2, r.gn2_w, r.gn2_b, nullptr, ln, B, R, C, 32, 1e-5f, 1, 0, st);
                });
                {
                    GemmEpilogue e;
                    e.out_f32 = g32;
                    GemmOp op;
                    PD_TRY(gemm_make(&op, ln, GemmGeom::conv(B, T, Hl, Wl, C, 3, 3, 3), rb.conv1_d, C, e));
                    bw.add_gemm(op);
                }
                bw.add([=](cudaStream_t st) {
                    return gn_bwd(xin, g32, s1, b1, r.gn1_w, r.gn1_b, dx, dxb, B, R, C, 32, 1e-5f, 1, 1, st);
                });
            }
        }
        if (l == 1) {   // PatchMerging3D backward: reduction^T, then LayerNorm backward scattered to the 2x2 sources
            {
                GemmEpilogue e;
                e.out_f32 = g32;
                GemmOp op;
                PD_TRY(gemm_make(&op, dxb, GemmGeom::linear(P, C1), pm_t, 4 * C0, e));
                bw.add_gemm(op);
            }
            const float *x0 = b.blk[0][cfg.depth[0] - 1].xs[6], *lw = pm_ln_w;
            float* dx0 = b.dx[0];
            bf16* dxb0 = b.dxb[0];
            const int BT = B * T, c0 = C0;
            bw.add([=](cudaStream_t st) { return patch_merge_ln_bwd(x0, lw, g32, dx0, dxb0, BT, H, W, c0, 1e-5f, st); });
        }
    }
    {   // ---- first_proj backward (pos-embed is additive: identity) ----
        const ResW r = first;
        const ResBwdW rb = first_b;
        bf16 *dxb = b.dxb[0], *ln = b.ln[0];
        float *g32 = b.g32[0], *gout = b.gout;
        const float *hf = b.hf, *z = b.z;
        const double *s1 = b.st_f1, *s2 = b.st_f2;
        double *b1 = b.bst_f1, *b2 = b.bst_f2;
        const int c0 = C0;
        {
            GemmEpilogue e;
            e.out_f32 = g32;
            GemmOp op;
            PD_TRY(gemm_make(&op, dxb, GemmGeom::conv(B, T, H, W, C0, 3, 3, 3), rb.conv2_d, C0, e));
            bw.add_gemm(op);
        }
        bw.add([=](cudaStream_t st) {
            return gn_bwd(hf, g32, s2, b2, r.gn2_w, r.gn2_b, nullptr, ln, B, R0, c0, 32, 1e-5f, 1, 0, st);
        });
        {
            GemmEpilogue e;
            e.out_f32 = g32;
            GemmOp op;
            PD_TRY(gemm_make(&op, ln, GemmGeom::conv(B, T, H, W, C0, 3, 3, 3), rb.conv1_d, cin, e));
            bw.add_gemm(op);
        }
        {   // 1x1x1 skip connection: its dgrad initialises the result
            GemmEpilogue e;
            e.out_f32 = gout;
            GemmOp op;
            PD_TRY(gemm_make(&op, dxb, GemmGeom::conv(B, T, H, W, C0, 1, 1, 1), skip_t, cin, e));
            bw.add_gemm(op);
        }
        bw.add([=](cudaStream_t st) {
            return gn_bwd(z, g32, s1, b1, r.gn1_w, r.gn1_b, gout, nullptr, B, R0, cin, gin, 1e-5f, 1, 1, st);
        });
    }
    return PD_OK;
}

int KANet::get_plan(int B, BatchPlan** out) {
    PD_CHECK(finalized, PD_ERR_STATE, "ka: call pd_ka_finalize() after loading all weights");
    PD_CHECK(B >= 1 && B <= cfg.max_batch, PD_ERR_SHAPE, "ka: batch %d outside [1, max_batch=%d]", B, cfg.max_batch);
    auto it = plans.find(B);
    if (it == plans.end()) {
        std::unique_ptr<BatchPlan> bp(new BatchPlan());
        PD_TRY(build_plan(B, bp.get()));
        it = plans.emplace(B, std::move(bp)).first;
    }
    *out = it->second.get();
    return PD_OK;
}

int KANet::bind(BatchPlan* bp, const float* zt, const int64_t* t, const int* step, int t_stride, int B) {
    Bufs& b = bp->bufs;
    float* z = b.z;
    const size_t bytes = (size_t)B * T * cfg.h * cfg.w * cfg.c * sizeof(float);
    bp->fwd.steps[bp->in_slot] = [=](cudaStream_t s) {
        PD_CUDA(cudaMemcpyAsync(z, zt, bytes, cudaMemcpyDeviceToDevice, s));
        return PD_OK;
    };
    float* e0 = b.e0;
    const int c0 = C0;
    bp->fwd.steps[bp->t_slot] = [=](cudaStream_t s) { return timestep_embedding(t, step, t_stride, e0, B, c0, s); };
    return PD_OK;
}

int KANet::forward(const float* zt, const int64_t* t, const int* step, int t_stride, float* pred, int B,
                   cudaStream_t st) {
    PD_CHECK(zt && t && pred, PD_ERR_ARG, "ka forward: null pointer");
    BatchPlan* bp = nullptr;
    PD_TRY(get_plan(B, &bp));
    PD_TRY(bind(bp, zt, t, step, t_stride, B));
    PD_TRY(bp->fwd.run(st));
    PD_CUDA(cudaMemcpyAsync(pred, bp->bufs.outv, (size_t)B * T * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return PD_OK;
}

int KANet::mean_shift(const float* zt, const int64_t* t, const int* step, int t_stride, const float* avg_x_gt,
                      float guide_scale, float* grad_out, int B, cudaStream_t st) {
    PD_CHECK(zt && t && avg_x_gt, PD_ERR_ARG, "ka mean_shift: null pointer");
    BatchPlan* bp = nullptr;
    PD_TRY(get_plan(B, &bp));
    PD_TRY(bind(bp, zt, t, step, t_stride, B));
    {
        Bufs& b = bp->bufs;
        const float* outv = b.outv;
        float *dout = b.dout, *loss = b.loss;
        const int Tn = T;
        bp->bwd.steps[bp->loss_slot] = [=](cudaStream_t s) {
            return ka_loss_grad(outv, avg_x_gt, dout, loss, B, Tn, guide_scale, s);
        };
    }
    PD_TRY(bp->fwd.run(st));
    PD_TRY(bp->bwd.run(st));
    if (grad_out)
        PD_CUDA(cudaMemcpyAsync(grad_out, bp->bufs.gout, (size_t)B * T * cfg.h * cfg.w * cfg.c * sizeof(float),
                                cudaMemcpyDeviceToDevice, st));
    return PD_OK;
}

int KANet::guide_buffer(int B, float** g, float** loss_dev) {
    BatchPlan* bp = nullptr;
    PD_TRY(get_plan(B, &bp));
    if (g) *g = bp->bufs.gout;
    if (loss_dev) *loss_dev = bp->bufs.loss;
    return PD_OK;
}

int KANet::kernels(int B, int* n_fwd, int* n_bwd) {
    BatchPlan* bp = nullptr;
    PD_TRY(get_plan(B, &bp));
    // gn_bwd is two kernels per plan step: 1 (read-out) + 2 per block + 2 (first_proj)
    if (n_fwd) *n_fwd = bp->fwd.num_kernels();
    if (n_bwd) *n_bwd = bp->bwd.num_kernels() + 3 + 2 * num_blocks();
    return PD_OK;
}

}  // namespace pd
