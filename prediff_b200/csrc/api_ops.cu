// extern "C" kernel-level entry points (include/prediff_b200.h): the same launchers the model programs use,
// exposed so the parity tests can drive every kernel in isolation through the C ABI.
#include <cstdlib>
#include <vector>
#include "../../include/prediff_b200.h"
#include "gemm.cuh"
#include "ops.cuh"

using namespace pd;

namespace {
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
}  // namespace

extern "C" {

int pd_init(void) { return gemm_init(); }
const char* pd_last_error(void) { return last_error(); }
const char* pd_version(void) { return "prediff_b200 0.1 (sm_100a)"; }

int pd_op_conv_gemm(const void* A, const void* Wt, int samples, int D, int H, int W, int C, int kt, int kh, int kw, int N,
                    const float* bias, const float* rowvec, const float* residual, float* out_f32, void* out_bf16,
                    int act, int block_n, void* stream) {
    PD_TRY(gemm_init());
    PD_CHECK(kt * kh * kw <= kMaxTaps, PD_ERR_SHAPE, "pd_op_conv_gemm: window too large");
    GemmGeom g = GemmGeom::conv(samples, D, H, W, C, kt, kh, kw);
    GemmEpilogue e;
    e.bias = bias; e.rowvec = rowvec; e.residual = residual; e.out_f32 = out_f32;
    e.out_bf16 = static_cast<bf16*>(out_bf16); e.act = act;
    // long reductions exercise the split-K path: give it its (zeroed) per-tile handshake flags
    int* flags = nullptr;
    const int nflags = block_n == 0 || block_n == 256 ? gemm_split_flags_needed(g, N) : 0;
    if (nflags > 0) {
        PD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&flags), nflags * sizeof(int), S(stream)));
        PD_CUDA(cudaMemsetAsync(flags, 0, nflags * sizeof(int), S(stream)));
        e.split_flags = flags;
    }
    GemmOp op;
    int rc = gemm_make(&op, static_cast<const bf16*>(A), g, static_cast<const bf16*>(Wt), N, e, block_n);
    if (rc == PD_OK) rc = gemm_launch(op, S(stream));
    if (flags) cudaFreeAsync(flags, S(stream));
    return rc;
}

// ---- TF32-class precision (PD_PRECISION_TF32): the same launchers with fp32 operands -----------------------------------
int pd_op_conv_gemm_tf32(const float* A, const float* Wt, int samples, int D, int H, int W, int C, int kt, int kh, int kw,
                         int N, const float* bias, const float* rowvec, const float* residual, float* out_f32, int act,
                         int round_out, int block_n, int streamk_ctas_per_sample, void* stream) {
    PD_TRY(gemm_init());
    PD_CHECK(kt * kh * kw <= kMaxTaps, PD_ERR_SHAPE, "pd_op_conv_gemm_tf32: window too large");
    GemmGeom g = GemmGeom::conv(samples, D, H, W, C, kt, kh, kw);
    g.tf32 = 1;
    GemmEpilogue e;
    e.bias = bias; e.rowvec = rowvec; e.residual = residual; e.out_f32 = out_f32; e.act = act; e.round_tf32 = round_out;
    if (streamk_ctas_per_sample > 0) {
        GemmOp op;
        PD_TRY(gemm_make(&op, A, g, Wt, N, e, 256));
        std::vector<SkSeg> segs;
        int n_slots = 0, n_flags = 0;
        PD_TRY(gemm_streamk_schedule(op, streamk_ctas_per_sample, &segs, &n_slots, &n_flags));
        SkSeg* dsegs = nullptr;
        float* partials = nullptr;
        int* flags = nullptr;
        PD_CUDA(cudaMalloc(reinterpret_cast<void**>(&dsegs), segs.size() * sizeof(SkSeg)));
        PD_CUDA(cudaMalloc(reinterpret_cast<void**>(&partials), (size_t)(n_slots + 1) * kGemmBlockM * 256 * sizeof(float)));
        PD_CUDA(cudaMalloc(reinterpret_cast<void**>(&flags), (size_t)n_flags * sizeof(int)));
        PD_CUDA(cudaMemcpy(dsegs, segs.data(), segs.size() * sizeof(SkSeg), cudaMemcpyHostToDevice));
        PD_CUDA(cudaMemset(flags, 0, (size_t)n_flags * sizeof(int)));
        int rc = gemm_streamk_attach(&op, dsegs, (int)segs.size() / 2, partials, flags);
        if (rc == PD_OK) rc = gemm_launch(op, S(stream));
        cudaStreamSynchronize(S(stream));
        cudaFree(dsegs); cudaFree(partials); cudaFree(flags);
        return rc;
    }
    int* flags = nullptr;
    const int nflags = block_n == 0 || block_n == 256 ? gemm_split_flags_needed(g, N) : 0;
    if (nflags > 0) {
        PD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&flags), nflags * sizeof(int), S(stream)));
        PD_CUDA(cudaMemsetAsync(flags, 0, nflags * sizeof(int), S(stream)));
        e.split_flags = flags;
    }
    GemmOp op;
    int rc = gemm_make(&op, A, g, Wt, N, e, block_n);
    if (rc == PD_OK) rc = gemm_launch(op, S(stream));
    if (flags) cudaFreeAsync(flags, S(stream));
    return rc;
}
int pd_op_pack_tf32(const float* w, float* out, int Co, int Ci, int taps, int Cipad, void* stream) {
    return taps == 0 ? pack_linear(w, out, Co, Ci, Cipad, S(stream), 1) : pack_conv(w, out, Co, Ci, taps, Cipad, S(stream), 1);
}
int pd_op_axial_attention_f32(const float* qkv, const float* bias_table, float* out, int B, int T, int H, int W, int C,
                              int heads, int axis, void* stream) {
    PD_TRY(gemm_init());
    return axial_attention(qkv, bias_table, out, B, T, H, W, C, heads, axis, S(stream), 1);
}
int pd_op_norm_tf32(int kind, const float* x, const float* gamma, const float* beta, float* y, int Sn, int R, int C, int G,
                    float eps, int silu, void* stream) {
    PD_TRY(gemm_init());
    if (kind == 1) return layer_norm(x, gamma, beta, y, Sn * R, C, eps, S(stream), 1);
    PD_CHECK(kind == 0, PD_ERR_ARG, "pd_op_norm_tf32: kind 0 = GroupNorm(+SiLU), 1 = LayerNorm");
    double* sums = nullptr;
    const size_t bytes = (size_t)Sn * G * 2 * sizeof(double);
    PD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&sums), bytes, S(stream)));
    PD_CUDA(cudaMemsetAsync(sums, 0, bytes, S(stream)));
    int rc = gn_stats(x, sums, Sn, R, C, G, S(stream));
    if (rc == PD_OK) rc = gn_apply(x, sums, gamma, beta, y, Sn, R, C, G, eps, silu, S(stream), 1);
    cudaFreeAsync(sums, S(stream));
    return rc;
}

int pd_op_linear_residual_ln_n(const void* A, const void* Wt, int M, int K, int N, const float* bias, float* x_inout,
                               const float* ln_gamma, const float* ln_beta, void* ln_out_bf16, void* stream) {
    PD_TRY(gemm_init());
    GemmGeom g = GemmGeom::linear(M, K);
    GemmEpilogue e;
    e.bias = bias; e.residual = x_inout; e.out_f32 = x_inout;
    e.ln_gamma = ln_gamma; e.ln_beta = ln_beta; e.ln_out = static_cast<bf16*>(ln_out_bf16);
    GemmOp op;
    PD_TRY(gemm_make(&op, static_cast<const bf16*>(A), g, static_cast<const bf16*>(Wt), N, e));
    return gemm_launch(op, S(stream));
}

int pd_op_linear_residual_ln(const void* A, const void* Wt, int M, int K, const float* bias, float* x_inout,
                             const float* ln_gamma, const float* ln_beta, void* ln_out_bf16, void* stream) {
    return pd_op_linear_residual_ln_n(A, Wt, M, K, 256, bias, x_inout, ln_gamma, ln_beta, ln_out_bf16, stream);
}

int pd_op_conv_gemm_phases(const void* A, const void* Wt, int samples, int D, int H, int W, int C, int kt, int kh, int kw,
                           int N, const float* bias, const float* residual, float* out_f32, void* out_bf16, int act,
                           int block_n, int dbg_block, unsigned long long* stamps9, void* stream) {
    PD_TRY(gemm_init());
    GemmGeom g = GemmGeom::conv(samples, D, H, W, C, kt, kh, kw);
    GemmEpilogue e;
    e.bias = bias; e.residual = residual; e.out_f32 = out_f32;
    e.out_bf16 = static_cast<bf16*>(out_bf16); e.act = act;
    e.dbg = stamps9; e.dbg_block = dbg_block;
    if (getenv("PD_PHASE_SPLIT")) {   // profiling aid: let the long-K convs split like they do inside the UNet plan
        static int* flags = nullptr;
        if (!flags) {
            PD_CUDA(cudaMalloc(reinterpret_cast<void**>(&flags), 8192 * sizeof(int)));
            PD_CUDA(cudaMemset(flags, 0, 8192 * sizeof(int)));
        }
        if (gemm_split_flags_needed(g, N) <= 8192) e.split_flags = flags;
    }
    GemmOp op;
    PD_TRY(gemm_make(&op, static_cast<const bf16*>(A), g, static_cast<const bf16*>(Wt), N, e, block_n));
    return gemm_launch(op, S(stream));
}

int pd_op_conv_s2_gemm(const void* planes, const void* Wt, int F, int Ho, int Wo, int C, int N, const float* bias,
                       float* out_f32, void* stream) {
    PD_TRY(gemm_init());
    GemmGeom g = GemmGeom::conv_s2_planes(F, Ho, Wo, C);
    GemmEpilogue e;
    e.bias = bias; e.out_f32 = out_f32;
    GemmOp op;
    PD_TRY(gemm_make(&op, static_cast<const bf16*>(planes), g, static_cast<const bf16*>(Wt), N, e));
    return gemm_launch(op, S(stream));
}

int pd_op_group_norm(const float* x, const float* gamma, const float* beta, void* y, int Sn, int R, int C, int G,
                     float eps, int silu, void* stream) {
    PD_TRY(gemm_init());
    double* sums = nullptr;
    const size_t bytes = (size_t)Sn * G * 2 * sizeof(double);
    PD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&sums), bytes, S(stream)));
    PD_CUDA(cudaMemsetAsync(sums, 0, bytes, S(stream)));
    int rc = gn_stats(x, sums, Sn, R, C, G, S(stream));
    if (rc == PD_OK) rc = gn_apply(x, sums, gamma, beta, static_cast<bf16*>(y), Sn, R, C, G, eps, silu, S(stream));
    cudaFreeAsync(sums, S(stream));
    return rc;
}

int pd_op_layer_norm(const float* x, const float* gamma, const float* beta, void* y, int P, int C, float eps,
                     void* stream) {
    PD_TRY(gemm_init());
    return layer_norm(x, gamma, beta, static_cast<bf16*>(y), P, C, eps, S(stream));
}

int pd_op_patch_merge_ln(const float* x, const float* gamma, const float* beta, void* y, int BT, int H, int W, int C,
                         float eps, void* stream) {
    PD_TRY(gemm_init());
    return patch_merge_ln(x, gamma, beta, static_cast<bf16*>(y), BT, H, W, C, eps, S(stream));
}

int pd_op_axial_attention(const void* qkv, const float* bias_table, void* out, int B, int T, int H, int W, int C,
                          int heads, int axis, void* stream) {
    PD_TRY(gemm_init());
    return axial_attention(static_cast<const bf16*>(qkv), bias_table, static_cast<bf16*>(out), B, T, H, W, C, heads, axis,
                           S(stream));
}

int pd_op_axial_attention_gv(const void* qkv, const float* bias_table, const void* gqkv_bf16, void* out, int B, int T, int H,
                             int W, int C, int heads, int axis, int n_global, void* stream) {
    PD_TRY(gemm_init());
    PD_CHECK(n_global >= 1 && gqkv_bf16, PD_ERR_ARG, "pd_op_axial_attention_gv: needs 1..16 global vectors");
    const GvKeys gk = gv_keys(static_cast<const bf16*>(gqkv_bf16), C, n_global, false, 3 * C, nullptr);
    return axial_attention(static_cast<const bf16*>(qkv), bias_table, static_cast<bf16*>(out), B, T, H, W, C, heads, axis,
                           S(stream), 0, &gk);
}

int pd_cuboid_tables(int T, int H, int W, const int32_t size[3], const int32_t strategy[3], const int32_t shift[3],
                     int padding_type, int32_t meta[12], int32_t* tok, int32_t* lab, int32_t* rel, int64_t capacity) {
    CuboidLayerSpec sp;
    for (int a = 0; a < 3; ++a) {
        sp.size[a] = size[a];
        sp.strategy[a] = strategy[a];
        sp.shift[a] = shift[a];
    }
    CuboidTables g;
    PD_TRY(build_cuboid_tables(T, H, W, sp, padding_type, &g));
    for (int a = 0; a < 3; ++a) {
        meta[a] = g.size[a];
        meta[3 + a] = g.shift[a];
        meta[6 + a] = g.pad[a];
    }
    meta[9] = g.num_cuboids;
    meta[10] = g.volume;
    meta[11] = g.rel_off;
    if (tok || lab || rel) {
        PD_CHECK((int64_t)g.tok.size() <= capacity, PD_ERR_ARG, "pd_cuboid_tables: capacity %lld < %zu", (long long)capacity,
                 g.tok.size());
        if (tok) memcpy(tok, g.tok.data(), g.tok.size() * sizeof(int));
        if (lab) memcpy(lab, g.lab.data(), g.lab.size() * sizeof(int));
        if (rel) memcpy(rel, g.rel.data(), g.rel.size() * sizeof(int));
    }
    return g.axial_axis >= 0 ? 1 + g.axial_axis : 0;
}

int pd_cuboid_tables_dst(int T, int H, int W, const int32_t size[3], const int32_t strategy[3], const int32_t shift[3],
                         int padding_type, int32_t* dst, int64_t capacity) {
    CuboidLayerSpec sp;
    for (int a = 0; a < 3; ++a) {
        sp.size[a] = size[a];
        sp.strategy[a] = strategy[a];
        sp.shift[a] = shift[a];
    }
    CuboidTables g;
    PD_TRY(build_cuboid_tables(T, H, W, sp, padding_type, &g));
    PD_CHECK(dst && (int64_t)g.dst.size() <= capacity, PD_ERR_ARG, "pd_cuboid_tables_dst: capacity %lld < %zu",
             (long long)capacity, g.dst.size());
    if (!g.dst.empty()) memcpy(dst, g.dst.data(), g.dst.size() * sizeof(int));
    return (int)(g.dst.empty() ? 0 : 1);
}

int pd_cuboid_tables_gmask(int T, int H, int W, const int32_t size[3], const int32_t strategy[3], const int32_t shift[3],
                           int padding_type, int32_t* gmask, int64_t capacity) {
    CuboidLayerSpec sp;
    for (int a = 0; a < 3; ++a) {
        sp.size[a] = size[a];
        sp.strategy[a] = strategy[a];
        sp.shift[a] = shift[a];
    }
    CuboidTables g;
    PD_TRY(build_cuboid_tables(T, H, W, sp, padding_type, &g));
    PD_CHECK(gmask && (int64_t)g.gmask.size() <= capacity, PD_ERR_ARG, "pd_cuboid_tables_gmask: capacity %lld < %zu",
             (long long)capacity, g.gmask.size());
    if (!g.gmask.empty()) memcpy(gmask, g.gmask.data(), g.gmask.size() * sizeof(int));
    return (int)(g.gmask.empty() ? 0 : 1);
}

int pd_op_gv_linear(const float* in, const float* ln_gamma, const float* ln_beta, const float* W, const float* bias,
                    const float* res, float* out_f32, void* out_bf16, int M, int K, int N, int act, void* stream) {
    return gv_linear(in, ln_gamma, ln_beta, W, bias, res, out_f32, static_cast<bf16*>(out_bf16), M, K, N, act, S(stream));
}

int pd_op_cuboid_attention_gv(const void* qkv, const float* bias_table, const float* gqkv_f32, const void* gqkv_bf16, void* out,
                              float* gout, int B, int T, int H, int W, int C, int heads, const int32_t size[3],
                              const int32_t strategy[3], const int32_t shift[3], int padding_type, int n_global, int self_attn,
                              void* stream) {
    return pd_op_cuboid_attention_gv2(qkv, bias_table, nullptr, gqkv_f32, gqkv_bf16, 3 * C, out, gout, B, T, H, W, C, heads, size,
                                      strategy, shift, padding_type, n_global, self_attn, 0, stream);
}

int pd_op_cuboid_attention_gv2(const void* qkv, const float* bias_table, const void* tok2_bf16, const float* grow_f32,
                               const void* grow_bf16, int grow_ld, void* out, float* gout, int B, int T, int H, int W, int C,
                               int heads, const int32_t size[3], const int32_t strategy[3], const int32_t shift[3], int padding_type,
                               int n_global, int self_attn, int line_kernel, void* stream) {
    PD_TRY(gemm_init());
    PD_CHECK(qkv && bias_table && grow_f32 && grow_bf16 && out && gout, PD_ERR_ARG, "pd_op_cuboid_attention_gv: null pointer");
    PD_CHECK(C % heads == 0 && heads >= 1, PD_ERR_SHAPE, "pd_op_cuboid_attention_gv: C=%d heads=%d", C, heads);
    const bool separate = tok2_bf16 != nullptr;
    PD_CHECK(grow_ld == (separate && self_attn ? 6 * C : 3 * C), PD_ERR_ARG,
             "pd_op_cuboid_attention_gv2: grow_ld %d (3C for the shared net or separate nets without self-attention, else 6C)", grow_ld);
    CuboidLayerSpec sp;
    for (int a = 0; a < 3; ++a) {
        sp.size[a] = size[a];
        sp.strategy[a] = strategy[a];
        sp.shift[a] = shift[a];
    }
    CuboidTables g;
    PD_TRY(build_cuboid_tables(T, H, W, sp, padding_type, &g));
    PD_CHECK(!line_kernel || (g.axial_axis >= 0 && n_global <= 16), PD_ERR_ARG,
             "pd_op_cuboid_attention_gv2: the line kernel takes axial layers with at most 16 global vectors");
    CuboidTablesDev d;
    PD_TRY(d.upload(g));
    cudaStream_t st = S(stream);
    const int n_keys = g.num_cuboids * g.volume + (self_attn ? n_global : 0);
    float* ws = nullptr;
    PD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws),
                            global_attention_workspace_floats(B, heads, n_global, C / heads, n_keys) * sizeof(float) + 16, st));
    const bf16* g16 = static_cast<const bf16*>(grow_bf16);
    const bf16* tok2 = static_cast<const bf16*>(tok2_bf16);
    const GvKeys gk = gv_keys(g16, C, n_global, separate, grow_ld, tok2);
    const GvQuery gq = gv_query(grow_f32, g16, static_cast<const bf16*>(qkv), tok2, C, separate, self_attn != 0, grow_ld);
    int rc = line_kernel ? axial_attention(qkv, bias_table, out, B, T, H, W, C, heads, g.axial_axis, st, 0, &gk)
                         : cuboid_attention(static_cast<const bf16*>(qkv), bias_table, static_cast<bf16*>(out), B, T * H * W, C,
                                            heads, d.dev, st, 1, &gk);
    if (rc == PD_OK) rc = global_attention(gq, gout, ws, B, T * H * W, C, heads, n_global, d.dev, st);
    cudaFreeAsync(ws, st);
    PD_CUDA(cudaStreamSynchronize(st));   // the tables are freed on return
    return rc;
}

int pd_op_cuboid_attention(const void* qkv, const float* bias_table, void* out, int B, int T, int H, int W, int C, int heads,
                           const int32_t size[3], const int32_t strategy[3], const int32_t shift[3], int padding_type,
                           void* stream) {
    return pd_op_cuboid_attention_impl(qkv, bias_table, out, B, T, H, W, C, heads, size, strategy, shift, padding_type, 0,
                                       stream);
}

int pd_op_cuboid_attention_impl(const void* qkv, const float* bias_table, void* out, int B, int T, int H, int W, int C,
                                int heads, const int32_t size[3], const int32_t strategy[3], const int32_t shift[3],
                                int padding_type, int impl, void* stream) {
    PD_TRY(gemm_init());
    PD_CHECK(impl >= 0 && impl <= 2, PD_ERR_ARG, "pd_op_cuboid_attention_impl: impl %d (0 auto, 1 mma.sync, 2 tcgen05)", impl);
    CuboidLayerSpec sp;
    for (int a = 0; a < 3; ++a) {
        sp.size[a] = size[a];
        sp.strategy[a] = strategy[a];
        sp.shift[a] = shift[a];
    }
    CuboidTables g;
    PD_TRY(build_cuboid_tables(T, H, W, sp, padding_type, &g));
    CuboidTablesDev d;
    PD_TRY(d.upload(g));
    PD_TRY(cuboid_attention(static_cast<const bf16*>(qkv), bias_table, static_cast<bf16*>(out), B, T * H * W, C, heads, d.dev,
                            S(stream), impl));
    PD_CUDA(cudaStreamSynchronize(S(stream)));   // the tables are freed on return
    return PD_OK;
}

// one launch with a stream-ordered temporary workspace (the models keep one per plan instead)
static int ffn_cluster_once(const void* ln_in_bf16, const void* W1_bf16, const float* b1, const void* W2_bf16, const float* b2,
                            float* x_inout, const float* ln_gamma, const float* ln_beta, void* ln_out_bf16, double* gn_sums,
                            int gn_groups, int gn_rows, int M, unsigned long long* stamps32, void* stream,
                            const FfnProjArgs* proj = nullptr) {
    PD_TRY(gemm_init());
    PD_CHECK(M >= 1, PD_ERR_ARG, "ffn_cluster: M must be positive");
    cudaStream_t st = S(stream);
    float* ws = nullptr;
    PD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws), ffn_cluster_workspace_bytes(M), st));
    FfnClusterOp op;
    int rc = ffn_cluster_make(&op, static_cast<const bf16*>(ln_in_bf16), M, static_cast<const bf16*>(W1_bf16), b1,
                              static_cast<const bf16*>(W2_bf16), b2, x_inout, ln_gamma, ln_beta,
                              static_cast<bf16*>(ln_out_bf16), 1e-5f, ws, proj);
    if (rc == PD_OK && gn_sums) rc = ffn_cluster_set_gn(&op, gn_sums, gn_groups, gn_rows);
    if (rc == PD_OK) {
        ffn_cluster_set_dbg(&op, stamps32);
        rc = ffn_cluster_launch(op, st);
    }
    cudaFreeAsync(ws, st);
    return rc;
}

int pd_op_ffn_cluster(const void* ln_in_bf16, const void* W1_bf16, const float* b1, const void* W2_bf16, const float* b2,
                      float* x_inout, const float* ln_gamma, const float* ln_beta, void* ln_out_bf16, double* gn_sums,
                      int gn_groups, int gn_rows, int M, void* stream) {
    return ffn_cluster_once(ln_in_bf16, W1_bf16, b1, W2_bf16, b2, x_inout, ln_gamma, ln_beta, ln_out_bf16, gn_sums, gn_groups,
                            gn_rows, M, nullptr, stream);
}

int pd_op_ffn_cluster_phases(const void* ln_in_bf16, const void* W1_bf16, const float* b1, const void* W2_bf16,
                             const float* b2, float* x_inout, const float* ln_gamma, const float* ln_beta, void* ln_out_bf16,
                             int M, unsigned long long* stamps32, void* stream) {
    return ffn_cluster_once(ln_in_bf16, W1_bf16, b1, W2_bf16, b2, x_inout, ln_gamma, ln_beta, ln_out_bf16, nullptr, 0, 0, M,
                            stamps32, stream);
}

int pd_op_proj_ffn_cluster(const void* att_bf16, const void* Wp_bf16, const float* bp, const float* ln1_gamma,
                           const float* ln1_beta, void* ln_scratch_bf16, const void* W1_bf16, const float* b1,
                           const void* W2_bf16, const float* b2, float* x_inout, const float* ln_gamma, const float* ln_beta,
                           void* ln_out_bf16, double* gn_sums, int gn_groups, int gn_rows, int M, unsigned long long* stamps32,
                           void* stream) {
    PD_CHECK(att_bf16 && Wp_bf16 && bp && ln1_gamma && ln1_beta && ln_scratch_bf16, PD_ERR_ARG,
             "proj_ffn_cluster: null projection argument");
    FfnProjArgs pa;
    pa.att = static_cast<const bf16*>(att_bf16); pa.wp = static_cast<const bf16*>(Wp_bf16); pa.bp = bp;
    pa.ln1_gamma = ln1_gamma; pa.ln1_beta = ln1_beta;
    return ffn_cluster_once(ln_scratch_bf16, W1_bf16, b1, W2_bf16, b2, x_inout, ln_gamma, ln_beta, ln_out_bf16, gn_sums,
                            gn_groups, gn_rows, M, stamps32, stream, &pa);
}

int pd_ssim_update(const float* pred, const float* target, int N, int H, int W, float data_range, double* state,
                   void* stream) {
    PD_TRY(gemm_init());
    return ssim_update(pred, target, N, H, W, data_range, state, S(stream));
}

int pd_op_q_sample(const float* x_start, const float* noise, const int64_t* t, const float* sqrt_alphas_cumprod,
                   const float* sqrt_one_minus_alphas_cumprod, float* out, int B, int64_t n_per_sample, void* stream) {
    PD_TRY(gemm_init());
    return q_sample(x_start, noise, t, sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod, out, B, n_per_sample, S(stream));
}

int pd_op_sampler_update(float* z, const float* eps, const float* noise, const float* guide, const float* coef8,
                         int64_t n, void* stream) {
    PD_TRY(gemm_init());
    return sampler_update(z, eps, noise, guide, coef8, nullptr, n, 0, S(stream));
}

int pd_op_timestep_embedding(const int64_t* t, float* out, int B, int dim, void* stream) {
    PD_TRY(gemm_init());
    return timestep_embedding(t, nullptr, 0, out, B, dim, S(stream));
}

int pd_op_small_linear(const float* in, const float* W, const float* bias, float* out, int B, int K, int N, int in_silu,
                       int out_silu, void* stream) {
    PD_TRY(gemm_init());
    return small_linear(in, W, bias, out, B, K, N, in_silu, out_silu, S(stream));
}

int pd_op_pack_conv(const float* w, void* out, int Co, int Ci, int taps, int Cipad, void* stream) {
    PD_TRY(gemm_init());
    return pack_conv(w, static_cast<bf16*>(out), Co, Ci, taps, Cipad, S(stream));
}

int pd_op_pack_linear(const float* w, void* out, int N, int K, int Kpad, void* stream) {
    PD_TRY(gemm_init());
    return pack_linear(w, static_cast<bf16*>(out), N, K, Kpad, S(stream));
}

int pd_op_upsample2x_cast(const float* x, void* y, int F, int H, int W, int C, void* stream) {
    PD_TRY(gemm_init());
    return upsample2x_cast(x, static_cast<bf16*>(y), F, H, W, C, S(stream));
}

int pd_op_parity_split_cast(const float* x, void* y, int F, int H, int W, int C, void* stream) {
    PD_TRY(gemm_init());
    return parity_split_cast(x, static_cast<bf16*>(y), F, H, W, C, S(stream));
}

// ---- input-gradient kernels (backward.cu) ------------------------------------------------------------------
int pd_op_group_norm_bwd(const float* x, const float* dy, const float* gamma, const float* beta, float* dx_io,
                         void* dx_bf16, int Sn, int R, int C, int G, float eps, int silu, int accumulate, void* stream) {
    PD_TRY(gemm_init());
    double* sums = nullptr;
    const size_t bytes = (size_t)Sn * G * 2 * sizeof(double);
    PD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&sums), 2 * bytes, S(stream)));
    PD_CUDA(cudaMemsetAsync(sums, 0, 2 * bytes, S(stream)));
    double* bsums = sums + (size_t)Sn * G * 2;
    int rc = gn_stats(x, sums, Sn, R, C, G, S(stream));
    if (rc == PD_OK)
        rc = gn_bwd(x, dy, sums, bsums, gamma, beta, dx_io, static_cast<bf16*>(dx_bf16), Sn, R, C, G, eps, silu,
                    accumulate, S(stream));
    cudaFreeAsync(sums, S(stream));
    return rc;
}

int pd_op_layer_norm_bwd(const float* x, const float* gamma, const float* dy, float* dx_io, void* dx_bf16, int P, int C,
                         float eps, int accumulate, void* stream) {
    PD_TRY(gemm_init());
    return layer_norm_bwd(x, gamma, dy, dx_io, static_cast<bf16*>(dx_bf16), P, C, eps, accumulate, S(stream));
}

int pd_op_patch_merge_ln_bwd(const float* x, const float* gamma, const float* dy, float* dx, void* dx_bf16, int BT, int H,
                             int W, int C, float eps, void* stream) {
    PD_TRY(gemm_init());
    return patch_merge_ln_bwd(x, gamma, dy, dx, static_cast<bf16*>(dx_bf16), BT, H, W, C, eps, S(stream));
}

int pd_op_gelu(const float* pre, void* y_bf16, int64_t n, void* stream) {
    PD_TRY(gemm_init());
    return gelu_fwd(pre, static_cast<bf16*>(y_bf16), n, S(stream));
}

int pd_op_gelu_bwd(const float* pre, const void* dy_bf16, void* dpre_bf16, int64_t n, void* stream) {
    PD_TRY(gemm_init());
    return gelu_bwd(pre, static_cast<const bf16*>(dy_bf16), static_cast<bf16*>(dpre_bf16), n, S(stream));
}

int pd_op_qkv_attn(const void* ln_bf16, const void* Wqkv_bf16, const float* bias_table, void* out_bf16, int B, int T, int H,
                   int W, int C, int heads, int axis, unsigned long long* stamps32, void* stream) {
    PD_TRY(gemm_init());
    QkvAttnOp op;
    PD_TRY(qkv_attn_make(&op, static_cast<const bf16*>(ln_bf16), static_cast<const bf16*>(Wqkv_bf16), bias_table,
                         static_cast<bf16*>(out_bf16), B, T, H, W, C, heads, axis));
    // PD_QKV_DBG_CTA="x,y": the CTA that writes the stamps (default 0,0)
    int cx = 0, cy = 0;
    if (const char* e = getenv("PD_QKV_DBG_CTA")) sscanf(e, "%d,%d", &cx, &cy);
    qkv_attn_set_dbg(&op, stamps32, cx, cy);
    return qkv_attn_launch(op, S(stream));
}

int pd_op_axial_attention_bwd(const void* qkv, const float* bias_table, const void* dout, void* dqkv, int B, int T, int H,
                              int W, int C, int heads, int axis, void* stream) {
    PD_TRY(gemm_init());
    return axial_attention_bwd(static_cast<const bf16*>(qkv), bias_table, static_cast<const bf16*>(dout),
                               static_cast<bf16*>(dqkv), B, T, H, W, C, heads, axis, S(stream));
}

int pd_op_pack_linear_t(const float* w, void* out, int N, int K, void* stream) {
    PD_TRY(gemm_init());
    return pack_linear_t(w, static_cast<bf16*>(out), N, K, S(stream));
}

int pd_op_pack_conv_dgrad(const float* w, void* out, int Co, int Ci, int taps, void* stream) {
    PD_TRY(gemm_init());
    return pack_conv_dgrad(w, static_cast<bf16*>(out), Co, Ci, taps, S(stream));
}

int pd_sevir_eval_update(const float* pred, const float* target, int64_t* counts, double* sums, int N, int T, int H,
                         int W, int pool, const float* thresholds_host, int n_thresholds, void* stream) {
    PD_TRY(gemm_init());
    return sevir_eval_update(pred, target, reinterpret_cast<long long*>(counts), sums, N, T, H, W, pool, thresholds_host,
                             n_thresholds, S(stream));
}

int pd_op_ffn_fused_phases(const void* ln_in_bf16, const void* W1_bf16, const float* b1, const void* W2_bf16,
                           const float* b2, float* x_inout, const float* ln_gamma, const float* ln_beta, void* ln_out_bf16,
                           int M, unsigned long long* stamps32, void* stream) {
    PD_TRY(gemm_init());
    FfnFusedOp op;
    PD_TRY(ffn_fused_make(&op, static_cast<const bf16*>(ln_in_bf16), M, static_cast<const bf16*>(W1_bf16), b1,
                          static_cast<const bf16*>(W2_bf16), b2, x_inout, ln_gamma, ln_beta,
                          static_cast<bf16*>(ln_out_bf16), 1e-5f, stamps32));
    return ffn_fused_launch(op, S(stream));
}

int pd_op_ffn_fused(const void* ln_in_bf16, const void* W1_bf16, const float* b1, const void* W2_bf16, const float* b2,
                    float* x_inout, const float* ln_gamma, const float* ln_beta, void* ln_out_bf16, int M, void* stream) {
    return pd_op_ffn_fused_phases(ln_in_bf16, W1_bf16, b1, W2_bf16, b2, x_inout, ln_gamma, ln_beta, ln_out_bf16, M,
                                  nullptr, stream);
}

int pd_op_proj_ffn_fused(const void* att_bf16, const void* Wp_bf16, const float* bp, const float* ln1_gamma,
                         const float* ln1_beta, void* ln_scratch_bf16, const void* W1_bf16, const float* b1,
                         const void* W2_bf16, const float* b2, float* x_inout, const float* ln_gamma, const float* ln_beta,
                         void* ln_out_bf16, int M, unsigned long long* stamps32, void* stream) {
    PD_TRY(gemm_init());
    FfnProjArgs pa;
    pa.att = static_cast<const bf16*>(att_bf16);
    pa.wp = static_cast<const bf16*>(Wp_bf16);
    pa.bp = bp;
    pa.ln1_gamma = ln1_gamma;
    pa.ln1_beta = ln1_beta;
    FfnFusedOp op;
    PD_TRY(ffn_fused_make(&op, static_cast<const bf16*>(ln_scratch_bf16), M, static_cast<const bf16*>(W1_bf16), b1,
                          static_cast<const bf16*>(W2_bf16), b2, x_inout, ln_gamma, ln_beta,
                          static_cast<bf16*>(ln_out_bf16), 1e-5f, stamps32, &pa));
    return ffn_fused_launch(op, S(stream));
}

int pd_op_conv_gemm_streamk(const void* A, const void* Wt, int samples, int D, int H, int W, int C, int kt, int kh, int kw,
                            int N, const float* bias, const float* rowvec, const float* residual, float* out_f32,
                            const float* ln_gamma, const float* ln_beta, void* ln_out_bf16, int ctas_per_sample,
                            void* stream) {
    return pd_op_conv_gemm_streamk_phases(A, Wt, samples, D, H, W, C, kt, kh, kw, N, bias, rowvec, residual, out_f32, ln_gamma,
                                          ln_beta, ln_out_bf16, ctas_per_sample, -1, nullptr, stream);
}

int pd_op_conv_gemm_streamk_phases(const void* A, const void* Wt, int samples, int D, int H, int W, int C, int kt, int kh,
                                   int kw, int N, const float* bias, const float* rowvec, const float* residual,
                                   float* out_f32, const float* ln_gamma, const float* ln_beta, void* ln_out_bf16,
                                   int ctas_per_sample, int dbg_cta, unsigned long long* stamps13, void* stream) {
    PD_TRY(gemm_init());
    GemmGeom g = GemmGeom::conv(samples, D, H, W, C, kt, kh, kw);
    GemmEpilogue e;
    e.bias = bias; e.rowvec = rowvec; e.residual = residual; e.out_f32 = out_f32;
    e.ln_gamma = ln_gamma; e.ln_beta = ln_beta; e.ln_out = static_cast<bf16*>(ln_out_bf16);
    e.dbg = stamps13; e.dbg_block = dbg_cta;
    GemmOp op;
    PD_TRY(gemm_make(&op, static_cast<const bf16*>(A), g, static_cast<const bf16*>(Wt), N, e, 256));
    std::vector<SkSeg> segs;
    int n_slots = 0, n_flags = 0;
    PD_TRY(gemm_streamk_schedule(op, ctas_per_sample, &segs, &n_slots, &n_flags));
    cudaStream_t st = S(stream);
    SkSeg* segs_dev = nullptr;
    float* partials = nullptr;
    int* flags = nullptr;
    PD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&segs_dev), segs.size() * sizeof(SkSeg), st));
    PD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&partials), (size_t)(n_slots + 1) * 128 * 256 * sizeof(float), st));
    PD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&flags), (size_t)n_flags * sizeof(int), st));
    PD_CUDA(cudaMemcpyAsync(segs_dev, segs.data(), segs.size() * sizeof(SkSeg), cudaMemcpyHostToDevice, st));
    PD_CUDA(cudaMemsetAsync(flags, 0, (size_t)n_flags * sizeof(int), st));
    PD_CUDA(cudaStreamSynchronize(st));   // the host vector dies at return
    PD_TRY(gemm_streamk_attach(&op, segs_dev, (int)segs.size() / 2, partials, flags));
    int rc = gemm_launch(op, st);
    if (rc == PD_OK && stamps13 && dbg_cta < -1) {   // back-to-back launches timed with events: ns per launch -> stamps13[15]
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        for (int i = 0; i < -dbg_cta && rc == PD_OK; ++i) rc = gemm_launch(op, st);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const unsigned long long ns = (unsigned long long)(ms * 1e6 / -dbg_cta);
        cudaMemcpy(stamps13 + 15, &ns, sizeof(ns), cudaMemcpyHostToDevice);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    cudaFreeAsync(segs_dev, st);
    cudaFreeAsync(partials, st);
    cudaFreeAsync(flags, st);
    return rc;
}

int pd_op_conv_gemm_gnstats(const void* A, const void* Wt, int samples, int D, int H, int W, int C, int kt, int kh, int kw,
                            int N, const float* bias, const float* residual, float* out_f32, double* gn_sums, int groups,
                            int streamk_ctas_per_sample, void* stream) {
    PD_TRY(gemm_init());
    GemmGeom g = GemmGeom::conv(samples, D, H, W, C, kt, kh, kw);
    GemmEpilogue e;
    e.bias = bias; e.residual = residual; e.out_f32 = out_f32;
    e.gn_sums = gn_sums; e.gn_groups = groups; e.gn_rows = D * H * W;
    GemmOp op;
    PD_TRY(gemm_make(&op, static_cast<const bf16*>(A), g, static_cast<const bf16*>(Wt), N, e, 256));
    cudaStream_t st = S(stream);
    if (streamk_ctas_per_sample <= 0) return gemm_launch(op, st);
    std::vector<SkSeg> segs;
    int n_slots = 0, n_flags = 0;
    PD_TRY(gemm_streamk_schedule(op, streamk_ctas_per_sample, &segs, &n_slots, &n_flags));
    SkSeg* segs_dev = nullptr;
    float* partials = nullptr;
    int* flags = nullptr;
    PD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&segs_dev), segs.size() * sizeof(SkSeg), st));
    PD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&partials), (size_t)(n_slots + 1) * 128 * 256 * sizeof(float), st));
    PD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&flags), (size_t)n_flags * sizeof(int), st));
    PD_CUDA(cudaMemcpyAsync(segs_dev, segs.data(), segs.size() * sizeof(SkSeg), cudaMemcpyHostToDevice, st));
    PD_CUDA(cudaMemsetAsync(flags, 0, (size_t)n_flags * sizeof(int), st));
    PD_CUDA(cudaStreamSynchronize(st));   // the host vector dies at return
    PD_TRY(gemm_streamk_attach(&op, segs_dev, (int)segs.size() / 2, partials, flags));
    const int rc = gemm_launch(op, st);
    cudaFreeAsync(segs_dev, st);
    cudaFreeAsync(partials, st);
    cudaFreeAsync(flags, st);
    return rc;
}

int pd_sevir_windows(const unsigned char* events_u8, int event_base, int n_events, int H, int W, int T_raw,
                     long long first_seq, int batch, int seq_len, int stride, float scale, float offset, float* out,
                     void* stream) {
    return sevir_windows(events_u8, event_base, n_events, H, W, T_raw, first_seq, batch, seq_len, stride, scale, offset, out,
                         S(stream));
}

}  // extern "C"
