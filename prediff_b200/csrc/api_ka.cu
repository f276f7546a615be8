// extern "C" entry points of the knowledge-alignment network (include/prediff_b200.h).
#include "ka_api.cuh"

using namespace pd;

namespace {
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
}

extern "C" {

int pd_ka_create(const pd_ka_config* cfg, pd_ka** out) {
    PD_CHECK(cfg && out, PD_ERR_ARG, "pd_ka_create: null argument");
    pd_ka* m = new (std::nothrow) pd_ka(*cfg);
    PD_CHECK(m, PD_ERR_CUDA, "pd_ka_create: out of host memory");
    const int rc = m->impl.validate();
    if (rc != PD_OK) {
        delete m;
        return rc;
    }
    *out = m;
    return PD_OK;
}
void pd_ka_destroy(pd_ka* m) { delete m; }
int pd_ka_num_weights(const pd_ka* m) { return m ? m->impl.ws.size() : 0; }
int pd_ka_weight_info(const pd_ka* m, int i, const char** name, int64_t shape[5]) {
    PD_CHECK(m && name && shape && i >= 0 && i < m->impl.ws.size(), PD_ERR_ARG, "pd_ka_weight_info: bad argument");
    const WeightEntry& e = m->impl.ws.at(i);
    *name = e.name.c_str();
    for (size_t d = 0; d < 5; ++d) shape[d] = d < e.shape.size() ? e.shape[d] : 0;
    return (int)e.shape.size();
}
int pd_ka_load_weight(pd_ka* m, const char* name, const float* data, const int64_t* shape, int ndim) {
    PD_CHECK(m, PD_ERR_ARG, "pd_ka_load_weight: null model");
    m->impl.finalized = false;
    return m->impl.ws.load(name, data, shape, ndim);
}
int pd_ka_finalize(pd_ka* m) {
    PD_CHECK(m, PD_ERR_ARG, "pd_ka_finalize: null model");
    return m->impl.finalize();
}
int pd_ka_forward(pd_ka* m, const float* zt, const int64_t* t, float* pred, int batch, void* stream) {
    PD_CHECK(m, PD_ERR_ARG, "pd_ka_forward: null model");
    return m->impl.forward(zt, t, nullptr, 0, pred, batch, S(stream));
}
int pd_ka_mean_shift(pd_ka* m, const float* zt, const int64_t* t, const float* avg_x_gt, float guide_scale, float* grad,
                     float* loss_out, int batch, void* stream) {
    PD_CHECK(m && grad, PD_ERR_ARG, "pd_ka_mean_shift: null argument");
    PD_TRY(m->impl.mean_shift(zt, t, nullptr, 0, avg_x_gt, guide_scale, grad, batch, S(stream)));
    if (loss_out) {
        float* l = nullptr;
        PD_TRY(m->impl.guide_buffer(batch, nullptr, &l));
        PD_CUDA(cudaMemcpyAsync(loss_out, l, sizeof(float), cudaMemcpyDeviceToDevice, S(stream)));
    }
    return PD_OK;
}
int pd_ka_kernels(pd_ka* m, int batch, int* n_forward, int* n_backward) {
    PD_CHECK(m, PD_ERR_ARG, "pd_ka_kernels: null model");
    return m->impl.kernels(batch, n_forward, n_backward);
}

}  // extern "C"
