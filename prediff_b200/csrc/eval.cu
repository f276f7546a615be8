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

}  // namespace

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
