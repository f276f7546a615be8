// extern "C" entry points for the AutoencoderKL (include/prediff_b200.h).
#include "vae.cuh"

using namespace pd;

struct pd_vae {
    VAE impl;
    explicit pd_vae(const pd_vae_config& c) : impl(c) {}
};

extern "C" {

int pd_vae_create(const pd_vae_config* cfg, pd_vae** out) {
    PD_CHECK(cfg && out, PD_ERR_ARG, "pd_vae_create: null argument");
    pd_vae* m = new (std::nothrow) pd_vae(*cfg);
    PD_CHECK(m, PD_ERR_CUDA, "pd_vae_create: out of host memory");
    const int rc = m->impl.validate();
    if (rc != PD_OK) {
        delete m;
        return rc;
    }
    *out = m;
    return PD_OK;
}
void pd_vae_destroy(pd_vae* m) { delete m; }
int pd_vae_num_weights(const pd_vae* m) { return m ? m->impl.ws.size() : 0; }
int pd_vae_weight_info(const pd_vae* m, int i, const char** name, int64_t shape[5]) {
    PD_CHECK(m && name && shape && i >= 0 && i < m->impl.ws.size(), PD_ERR_ARG, "pd_vae_weight_info: bad argument");
    const WeightEntry& e = m->impl.ws.at(i);
    *name = e.name.c_str();
    for (size_t d = 0; d < 5; ++d) shape[d] = d < e.shape.size() ? e.shape[d] : 0;
    return (int)e.shape.size();
}
int pd_vae_load_weight(pd_vae* m, const char* name, const float* data, const int64_t* shape, int ndim) {
    PD_CHECK(m, PD_ERR_ARG, "pd_vae_load_weight: null model");
    m->impl.finalized = false;
    return m->impl.ws.load(name, data, shape, ndim);
}
int pd_vae_finalize(pd_vae* m) {
    PD_CHECK(m, PD_ERR_ARG, "pd_vae_finalize: null model");
    return m->impl.finalize();
}
int pd_vae_encode(pd_vae* m, const float* x, float* moments, int n, void* stream) {
    PD_CHECK(m, PD_ERR_ARG, "pd_vae_encode: null model");
    return m->impl.encode(x, moments, n, reinterpret_cast<cudaStream_t>(stream));
}
int pd_vae_decode(pd_vae* m, const float* z, float* out, int n, void* stream) {
    PD_CHECK(m, PD_ERR_ARG, "pd_vae_decode: null model");
    return m->impl.decode(z, out, n, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
