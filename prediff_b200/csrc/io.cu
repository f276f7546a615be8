// Input side of the sampling path: raw SEVIR-LR VIL events (uint8, 'NHWT' as stored in the
// HDF5 files) -> the (B, T, H, W, 1) fp32 batch LatentDiffusion.sample receives. Reference:
// src/prediff/datasets/sevir/sevir_dataloader.py:834-877 (_idx_sample: sequent windows), :610-650 (preprocess_data_dict:
// scale * (float(x) + offset)), :71-84 (change_layout NHWT -> NTHWC). The events travel host -> device as uint8 (4x fewer
// bytes than the fp32 batch the reference moves) and are windowed, rescaled and transposed in one pass here.
#include "ops.cuh"

namespace pd {
namespace {

// One block per (sample, image row): the row's W x T_raw bytes are contiguous in the event; they are staged in shared
// memory and written back time-major, W floats per frame row (coalesced both ways).
__global__ void __launch_bounds__(256) sevir_windows_kernel(const uint8_t* __restrict__ events, int event_base, int H, int W,
                                                            int T_raw, long long first_seq, int n_per_event, int stride,
                                                            int T, float scale, float offset, float* __restrict__ out) {
    extern __shared__ uint8_t row[];
    grid_dep_launch();
    grid_dep_wait();
    const int b = blockIdx.y, h = blockIdx.x;
    const long long s = first_seq + b;
    const int e = (int)(s / n_per_event) - event_base;
    const int t0 = (int)(s % n_per_event) * stride;
    const uint8_t* src = events + ((size_t)e * H + h) * (size_t)W * T_raw;
    const int nbytes = W * T_raw;
    if ((nbytes & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0) {
        for (int i = threadIdx.x; i < nbytes / 4; i += blockDim.x)
            reinterpret_cast<uint32_t*>(row)[i] = __ldg(reinterpret_cast<const uint32_t*>(src) + i);
    } else {
        for (int i = threadIdx.x; i < nbytes; i += blockDim.x) row[i] = __ldg(src + i);
    }
    __syncthreads();
    float* dst = out + (((size_t)b * T) * H + h) * W;   // + t * H * W + w
    for (int i = threadIdx.x; i < T * W; i += blockDim.x) {
        const int t = i / W, w = i - t * W;
        const float x = (float)row[w * T_raw + t0 + t];
        dst[(size_t)t * H * W + w] = __fmul_rn(scale, __fadd_rn(x, offset));   // exactly scale * (x + offset), no contraction
    }
}

}  // namespace

int sevir_windows(const uint8_t* events, int event_base, int n_events, int H, int W, int T_raw, long long first_seq, int batch,
                  int seq_len, int stride, float scale, float offset, float* out, cudaStream_t st) {
    PD_CHECK(events && out, PD_ERR_ARG, "sevir_windows: null pointer");
    PD_CHECK(H > 0 && W > 0 && T_raw > 0 && batch > 0 && seq_len > 0 && stride > 0 && seq_len <= T_raw && first_seq >= 0,
             PD_ERR_SHAPE, "sevir_windows: bad shape");
    const int n_per_event = 1 + (T_raw - seq_len) / stride;   // sevir_dataloader.py:310-311
    const long long e_first = first_seq / n_per_event, e_last = (first_seq + batch - 1) / n_per_event;
    PD_CHECK(e_first >= event_base && e_last < (long long)event_base + n_events, PD_ERR_SHAPE,
             "sevir_windows: sequences [%lld, %lld) need events [%lld, %lld], buffer holds [%d, %d)", first_seq,
             first_seq + batch, e_first, e_last, event_base, event_base + n_events);
    const size_t smem = (size_t)W * T_raw;
    PD_CHECK(smem <= 48 * 1024, PD_ERR_SHAPE, "sevir_windows: W * T_raw = %zu bytes exceed one block's staging row", smem);
    PD_LAUNCH(sevir_windows_kernel, dim3(H, batch), 256, smem, st, events, event_base, H, W, T_raw, first_seq, n_per_event,
              stride, seq_len, scale, offset, out);
    return PD_OK;
}

}  // namespace pd
