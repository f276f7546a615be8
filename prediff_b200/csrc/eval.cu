// On-device evaluation step right after the decoder: the SEVIR skill-score contingency
// counts + squared / absolute error sums in ONE pass over the forecast and the target, accumulated into device-resident
// state, so decoded frames never round-trip to the host between `sample()` and the metrics.
// Reference: SEVIRSkillScore.update / calc_seq_hits_misses_fas / preprocess (src/prediff/datasets/sevir/evaluation.py:
// 197-245), _threshold (:12-37), process_data_dict_back (sevir_dataloader.py:652-681, 'vil' scale 1/255), and the
// torchmetrics MeanSquaredError / MeanAbsoluteError sums used beside it (train_sevirlr_prediff.py:960-962).
// HBM-bound: 8 B per pixel read once, all thresholds evaluated from registers; integer counts are exact.
#include "ops.cuh"

namespace pd {
namespace {

constexpr int kMaxThr = 8;
constexpr int kEvalThreads = 256;

struct EvalParams {
    float thr[kMaxThr];
    int n_thr;
};

// pred / target: [N][T][H][W] fp32 in [0, 1]. pool: max-pool window (1 = none; H, W multiples of pool).
// counts: int64 [n_thr][T][3] (hits, misses, false alarms); sums: double [2] (sum sq err, sum abs err) over raw pixels.
__global__ void __launch_bounds__(kEvalThreads) sevir_eval_kernel(const float* __restrict__ pred,
                                                                  const float* __restrict__ target,
                                                                  unsigned long long* __restrict__ counts,
                                                                  double* __restrict__ sums, int N, int T, int H, int W,
                                                                  int pool, const EvalParams prm) {
    // one (n, t) frame slice per blockIdx.y; blockIdx.x strides over the pooled pixels of that frame
    const int nt = blockIdx.y;
    const int t = nt % T;
    const int Hp = H / pool, Wp = W / pool;
    const int cells = Hp * Wp;
    const float* pf = pred + (size_t)nt * H * W;
    const float* tf = target + (size_t)nt * H * W;
    const float scale = 1.0f / 255.0f;   // PREPROCESS_SCALE_01['vil'] as fp32; back-transform is x / scale
    unsigned int hit[kMaxThr], mis[kMaxThr], fa[kMaxThr];
#pragma unroll
    for (int i = 0; i < kMaxThr; ++i) hit[i] = mis[i] = fa[i] = 0u;
    double se = 0.0, ae = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += gridDim.x * blockDim.x) {
        const int y0 = (c / Wp) * pool, x0 = (c % Wp) * pool;
        // max-pool of the back-transformed values; like F.max_pool2d a NaN in the window propagates (max with NaN
        // is NaN in torch) and then fails both comparisons
        float pm = -INFINITY, tm = -INFINITY;
        bool nan_p = false, nan_t = false;
        for (int dy = 0; dy < pool; ++dy)
            for (int dx = 0; dx < pool; ++dx) {
                const float pv = __ldg(pf + (size_t)(y0 + dy) * W + x0 + dx);
                const float tv = __ldg(tf + (size_t)(y0 + dy) * W + x0 + dx);
                const float d = pv - tv;
                se += (double)(d * d);
                ae += (double)fabsf(d);
                const float pb = __fdiv_rn(pv, scale), tb = __fdiv_rn(tv, scale);
                nan_p |= isnan(pb);
                nan_t |= isnan(tb);
                pm = fmaxf(pm, pb);
                tm = fmaxf(tm, tb);
            }
        const bool bad = nan_p || nan_t;
#pragma unroll
        for (int i = 0; i < kMaxThr; ++i) {
            if (i < prm.n_thr) {
                const bool tt = !bad && tm >= prm.thr[i];
                const bool pp = !bad && pm >= prm.thr[i];
                hit[i] += (tt && pp) ? 1u : 0u;
                mis[i] += (tt && !pp) ? 1u : 0u;
                fa[i] += (!tt && pp) ? 1u : 0u;
            }
        }
    }
    // block reduction: warp shuffles, then one atomic per (warp, counter)
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < kMaxThr; ++i) {
        if (i < prm.n_thr) {
            unsigned int h = hit[i], m = mis[i], f = fa[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                h += __shfl_xor_sync(0xffffffffu, h, o);
                m += __shfl_xor_sync(0xffffffffu, m, o);
                f += __shfl_xor_sync(0xffffffffu, f, o);
            }
            if (lane == 0) {
                unsigned long long* c3 = counts + ((size_t)i * T + t) * 3;
                if (h) atomicAdd(c3 + 0, (unsigned long long)h);
                if (m) atomicAdd(c3 + 1, (unsigned long long)m);
                if (f) atomicAdd(c3 + 2, (unsigned long long)f);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        se += __shfl_xor_sync(0xffffffffu, se, o);
        ae += __shfl_xor_sync(0xffffffffu, ae, o);
    }
    if (lane == 0) {
        atomicAdd(sums + 0, se);
        atomicAdd(sums + 1, ae);
    }
}

// ---- SSIM (torchmetrics.image.StructuralSimilarityIndexMeasure with its defaults, as the inference script builds it:
// train_sevirlr_prediff.py:229-230, updated per sampled batch at :964-965) ------------------------------------------
// torchmetrics 1.2.0 (setup.py:33) functional/image/ssim.py::_ssim_update, C = 1: 11 x 11 Gaussian window (sigma 1.5,
// size int(3.5 sigma + 0.5) * 2 + 1), reflect padding by 5 that the final crop [5:-5, 5:-5] removes again - so only the
// (H - 10) x (W - 10) interior is ever used and it only reads real pixels; data_range = None -> max(pred.max() -
// pred.min(), target.max() - target.min()) over the update batch; c1 = (0.01 dr)^2, c2 = (0.03 dr)^2;
//   ssim = ((2 mu_p mu_t + c1)(2 s_pt + c2)) / ((mu_p^2 + mu_t^2 + c1)(s_p^2 + s_t^2 + c2));
// per image: mean over the interior; state: sum of the per-image means, number of images (reduction elementwise_mean).
constexpr int kSsimR = 5, kSsimK = 11, kSsimTile = 32, kSsimWin = kSsimTile + 2 * kSsimR;   // 42

__global__ void __launch_bounds__(256) minmax_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                             int64_t n, float* __restrict__ part /* [grid][4] */) {
    float amin = INFINITY, amax = -INFINITY, bmin = INFINITY, bmax = -INFINITY;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = __ldg(a + i), y = __ldg(b + i);
        amin = fminf(amin, x); amax = fmaxf(amax, x);
        bmin = fminf(bmin, y); bmax = fmaxf(bmax, y);
    }
    __shared__ float red[8][4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        amin = fminf(amin, __shfl_xor_sync(0xffffffffu, amin, o)); amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        bmin = fminf(bmin, __shfl_xor_sync(0xffffffffu, bmin, o)); bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        float* r = red[threadIdx.x >> 5];
        r[0] = amin; r[1] = amax; r[2] = bmin; r[3] = bmax;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            amin = fminf(amin, red[w][0]); amax = fmaxf(amax, red[w][1]);
            bmin = fminf(bmin, red[w][2]); bmax = fmaxf(bmax, red[w][3]);
        }
        float* o = part + blockIdx.x * 4;
        o[0] = amin; o[1] = amax; o[2] = bmin; o[3] = bmax;
    }
}

struct SsimWindow { float g[kSsimK]; };

// one block = one 32 x 32 tile of the interior of one image; separable Gaussian: rows first (42 x 32 x 5 maps in shared
// memory), then columns; per-tile sum of the SSIM map in double -> partial[image][tile]
__global__ void __launch_bounds__(256) ssim_tile_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                        int H, int W, float data_range, const float* __restrict__ mm_part,
                                                        int mm_n, const SsimWindow win, double* __restrict__ partial) {
    __shared__ float sp[kSsimWin][kSsimWin + 1], stt[kSsimWin][kSsimWin + 1];
    __shared__ float sh[5][kSsimWin][kSsimTile + 1];
    __shared__ double red[8];
    __shared__ float s_dr;
    const int n = blockIdx.z;
    const int IH = H - 2 * kSsimR, IW = W - 2 * kSsimR;   // interior
    const int y0 = blockIdx.y * kSsimTile, x0 = blockIdx.x * kSsimTile;   // interior coordinates of the tile
    const float* pf = pred + (size_t)n * H * W;
    const float* tf = target + (size_t)n * H * W;
    if (threadIdx.x == 0) {
        float dr = data_range;
        if (!(dr > 0.f)) {   // data_range = None: from the batch's value ranges
            float amin = INFINITY, amax = -INFINITY, bmin = INFINITY, bmax = -INFINITY;
            for (int i = 0; i < mm_n; ++i) {
                amin = fminf(amin, mm_part[4 * i]); amax = fmaxf(amax, mm_part[4 * i + 1]);
                bmin = fminf(bmin, mm_part[4 * i + 2]); bmax = fmaxf(bmax, mm_part[4 * i + 3]);
            }
            dr = fmaxf(amax - amin, bmax - bmin);
        }
        s_dr = dr;
    }
    for (int i = threadIdx.x; i < kSsimWin * kSsimWin; i += blockDim.x) {
        const int r = i / kSsimWin, c = i - r * kSsimWin;
        const int y = y0 + r, x = x0 + c;   // image coordinates (interior (0,0) = image (5,5), window starts 5 earlier)
        const bool in = y < H && x < W;
        sp[r][c] = in ? __ldg(pf + (size_t)y * W + x) : 0.f;
        stt[r][c] = in ? __ldg(tf + (size_t)y * W + x) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSsimWin * kSsimTile; i += blockDim.x) {
        const int r = i / kSsimTile, c = i - r * kSsimTile;
        float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
        for (int k = 0; k < kSsimK; ++k) {
            const float p = sp[r][c + k], t = stt[r][c + k], g = win.g[k];
            a = fmaf(g, p, a); b = fmaf(g, t, b);
            aa = fmaf(g, p * p, aa); bb = fmaf(g, t * t, bb); ab = fmaf(g, p * t, ab);
        }
        sh[0][r][c] = a; sh[1][r][c] = b; sh[2][r][c] = aa; sh[3][r][c] = bb; sh[4][r][c] = ab;
    }
    __syncthreads();
    const float dr = s_dr;
    const float c1 = (0.01f * dr) * (0.01f * dr), c2 = (0.03f * dr) * (0.03f * dr);
    double acc = 0.0;
    for (int i = threadIdx.x; i < kSsimTile * kSsimTile; i += blockDim.x) {
        const int r = i / kSsimTile, c = i - r * kSsimTile;
        if (y0 + r >= IH || x0 + c >= IW) continue;
        float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < kSsimK; ++k) {
            const float g = win.g[k];
#pragma unroll
            for (int q = 0; q < 5; ++q) m[q] = fmaf(g, sh[q][r + k][c], m[q]);
        }
        const float mu_pp = m[0] * m[0], mu_tt = m[1] * m[1], mu_pt = m[0] * m[1];
        const float s_pp = m[2] - mu_pp, s_tt = m[3] - mu_tt, s_pt = m[4] - mu_pt;
        const float upper = 2.f * s_pt + c2, lower = s_pp + s_tt + c2;
        acc += (double)(((2.f * mu_pt + c1) * upper) / ((mu_pp + mu_tt + c1) * lower));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        partial[((size_t)n * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
    }
}

// state[0] += sum over images of (tile sums in fixed order / interior pixels); state[1] += N
__global__ void ssim_finish_kernel(const double* __restrict__ partial, int N, int tiles, double inv_pixels,
                                   double* __restrict__ state) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double tot = 0.0;
    for (int n = 0; n < N; ++n) {
        double s = 0.0;
        for (int t = 0; t < tiles; ++t) s += partial[(size_t)n * tiles + t];
        tot += (double)(float)(s * inv_pixels);   // the per-image mean is an fp32 tensor in the reference
    }
    state[0] += tot;
    state[1] += (double)N;
}

}  // namespace

int ssim_update(const float* pred, const float* target, int N, int H, int W, float data_range, double* state,
                cudaStream_t st) {
    if (N == 0) return PD_OK;
    PD_CHECK(pred && target && state, PD_ERR_ARG, "ssim_update: null argument");
    PD_CHECK(H > 2 * kSsimR && W > 2 * kSsimR && N <= 65535, PD_ERR_SHAPE, "ssim_update: frames must exceed the 11 x 11 window");
    SsimWindow win;
    {
        double g[kSsimK], sum = 0.0;   // _gaussian(11, 1.5): exp(-(d / sigma)^2 / 2), normalised (fp32 like the reference)
        for (int k = 0; k < kSsimK; ++k) {
            const float d = (float)(k - kSsimR);
            const float e = expf(-powf(d / 1.5f, 2.f) / 2.f);
            g[k] = e;
            sum += e;
        }
        float fsum = 0.f;
        for (int k = 0; k < kSsimK; ++k) fsum += (float)g[k];
        (void)sum;
        for (int k = 0; k < kSsimK; ++k) win.g[k] = (float)g[k] / fsum;
    }
    const int IH = H - 2 * kSsimR, IW = W - 2 * kSsimR;
    const dim3 grid(ceil_div(IW, kSsimTile), ceil_div(IH, kSsimTile), N);
    const int tiles = grid.x * grid.y;
    constexpr int kMmBlocks = 148;
    uint8_t* ws = nullptr;
    const size_t bytes = (size_t)N * tiles * sizeof(double) + kMmBlocks * 4 * sizeof(float);
    PD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&ws), bytes, st));
    double* partial = reinterpret_cast<double*>(ws);
    float* mm = reinterpret_cast<float*>(ws + (size_t)N * tiles * sizeof(double));
    if (!(data_range > 0.f)) {
        minmax_partial_kernel<<<kMmBlocks, 256, 0, st>>>(pred, target, (int64_t)N * H * W, mm);
        PD_LAUNCH_CHECK();
    }
    ssim_tile_kernel<<<grid, 256, 0, st>>>(pred, target, H, W, data_range, mm, kMmBlocks, win, partial);
    PD_LAUNCH_CHECK();
    ssim_finish_kernel<<<1, 32, 0, st>>>(partial, N, tiles, 1.0 / ((double)IH * IW), state);
    PD_LAUNCH_CHECK();
    PD_CUDA(cudaFreeAsync(ws, st));
    return PD_OK;
}

int sevir_eval_update(const float* pred, const float* target, long long* counts, double* sums, int N, int T, int H, int W,
                      int pool, const float* thresholds, int n_thr, cudaStream_t st) {
    if (N == 0) return PD_OK;   // empty batch: nothing to add (an empty tensor has no storage)
    PD_CHECK(pred && target && counts && sums && thresholds, PD_ERR_ARG, "sevir_eval_update: null argument");
    PD_CHECK(n_thr >= 1 && n_thr <= kMaxThr, PD_ERR_SHAPE, "sevir_eval_update: 1..%d thresholds (got %d)", kMaxThr, n_thr);
    PD_CHECK(pool >= 1 && H % pool == 0 && W % pool == 0, PD_ERR_SHAPE, "sevir_eval_update: pool %d must divide H, W", pool);
    PD_CHECK(N >= 0 && T >= 1 && (long long)N * T <= 65535, PD_ERR_SHAPE, "sevir_eval_update: N * T = %lld too large",
             (long long)N * T);
    EvalParams prm;
    prm.n_thr = n_thr;
    for (int i = 0; i < kMaxThr; ++i) prm.thr[i] = i < n_thr ? thresholds[i] : 0.f;
    const int cells = (H / pool) * (W / pool);
    int bx = ceil_div(cells, kEvalThreads * 4);
    if (bx < 1) bx = 1;
    sevir_eval_kernel<<<dim3(bx, N * T), kEvalThreads, 0, st>>>(pred, target, reinterpret_cast<unsigned long long*>(counts),
                                                               sums, N, T, H, W, pool, prm);
    PD_LAUNCH_CHECK();
    return PD_OK;
}

}  // namespace pd
