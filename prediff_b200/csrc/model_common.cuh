// Host-side plumbing shared by the UNet / VAE programs: named fp32 weight store (reference state_dict keys),
// device buffers, a bump arena for activations, and the "plan" - the list of kernel launches one forward is.
#pragma once
#include "common.cuh"
#include "gemm.cuh"
#include "ops.cuh"
#include <cstdarg>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace pd {

struct DevMem {
    void* p = nullptr;
    size_t bytes = 0;
    DevMem() = default;
    DevMem(const DevMem&) = delete;
    DevMem& operator=(const DevMem&) = delete;
    ~DevMem() { release(); }
    int alloc(size_t n) {
        release();
        if (n == 0) n = 16;
        PD_CUDA(cudaMalloc(&p, n));
        bytes = n;
        return PD_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

// Bump allocator over one device block; every allocation is 1 KiB aligned (TMA needs 16 B, keep tiles apart).
class Arena {
public:
    int reserve(size_t bytes) { used_ = 0; return mem_.alloc(bytes); }
    template <class T>
    T* take(size_t count) {
        const size_t b = (size_t)round_up64((int64_t)(count * sizeof(T)), 1024);
        if (used_ + b > mem_.bytes) { overflow_ = true; return nullptr; }
        T* r = reinterpret_cast<T*>(static_cast<uint8_t*>(mem_.p) + used_);
        used_ += b;
        return r;
    }
    bool overflowed() const { return overflow_; }
    size_t used() const { return used_; }
    size_t capacity() const { return mem_.bytes; }
private:
    DevMem mem_;
    size_t used_ = 0;
    bool overflow_ = false;
};

// Sizing pass: same take() calls against a null arena to learn the byte count.
class ArenaSizer {
public:
    template <class T>
    T* take(size_t count) {
        used_ += (size_t)round_up64((int64_t)(count * sizeof(T)), 1024);
        return nullptr;
    }
    size_t used() const { return used_; }
private:
    size_t used_ = 0;
};

struct WeightEntry {
    std::string name;
    std::vector<int64_t> shape;
    int64_t numel = 0;
    std::unique_ptr<DevMem> data;  // fp32, reference layout; null until loaded
};

class WeightStore {
public:
    void declare(const std::string& name, std::vector<int64_t> shape) {
        WeightEntry e;
        e.name = name;
        e.numel = 1;
        for (auto d : shape) e.numel *= d;
        e.shape = std::move(shape);
        index_[name] = (int)entries_.size();
        entries_.push_back(std::move(e));
    }
    int size() const { return (int)entries_.size(); }
    const WeightEntry& at(int i) const { return entries_[i]; }
    int load(const char* name, const float* src, const int64_t* shape, int ndim) {
        PD_CHECK(name && src && shape, PD_ERR_ARG, "load_weight: null argument");
        auto it = index_.find(name);
        PD_CHECK(it != index_.end(), PD_ERR_WEIGHT, "load_weight: unknown key '%s'", name);
        WeightEntry& e = entries_[it->second];
        bool same = ndim == (int)e.shape.size();
        for (int i = 0; same && i < ndim; ++i) same = shape[i] == e.shape[i];
        PD_CHECK(same, PD_ERR_WEIGHT, "load_weight: shape mismatch for '%s'", name);
        if (!e.data) {
            e.data.reset(new DevMem());
            PD_TRY(e.data->alloc((size_t)e.numel * sizeof(float)));
        }
        PD_CUDA(cudaMemcpy(e.data->p, src, (size_t)e.numel * sizeof(float), cudaMemcpyDefault));
        return PD_OK;
    }
    // nullptr + error if missing
    const float* get(const std::string& name) const {
        auto it = index_.find(name);
        if (it == index_.end() || !entries_[it->second].data) {
            set_error("weight '%s' was not loaded", name.c_str());
            return nullptr;
        }
        return entries_[it->second].data->as<float>();
    }
    int check_complete() const {
        for (const auto& e : entries_)
            PD_CHECK(e.data != nullptr, PD_ERR_WEIGHT, "finalize: weight '%s' was never loaded", e.name.c_str());
        return PD_OK;
    }
private:
    std::vector<WeightEntry> entries_;
    std::map<std::string, int> index_;
};

using Step = std::function<int(cudaStream_t)>;

enum StepKind : uint8_t { STEP_KERNEL = 0, STEP_GEMM = 1, STEP_NONE = 2 /* memset / placeholder without a kernel */ };

struct PlanProfile {
    double gemm_ms = 0, other_ms = 0, gemm_flops = 0;
    int n_gemm = 0, n_other = 0;
};

// Writes %globaltimer (ns) to *slot: the per-launch trace of Plan::run_traced (tools/trace_unet.py).
int stamp_globaltimer(unsigned long long* slot, cudaStream_t st);

struct Plan {
    Plan() = default;
    Plan(const Plan&) = delete;
    Plan& operator=(const Plan&) = delete;
    ~Plan() {   // plans are rebuilt by every finalize() (weight refresh): their streams and events must not pile up
        if (aux) cudaStreamDestroy(aux);
        if (side) cudaStreamDestroy(side);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (ev_side_join) cudaEventDestroy(ev_side_join);
        for (cudaEvent_t e : mark_ev)
            if (e) cudaEventDestroy(e);
    }
    std::vector<Step> steps;
    std::vector<uint8_t> kinds;
    std::vector<std::string> labels;   // one per step: kernel class + site, for the per-launch trace
    std::vector<double> flops;         // one per step: algorithmic FLOPs of a GEMM-family step (0 otherwise)
    std::string scope;                 // prefix applied to labels added from now on (e.g. "L0.res")
    double gemm_flops = 0;
    int n_gemm = 0;
    // Optional side branch: steps [fork_begin, fork_end) do not depend on the steps that follow them until step
    // `join_before` (the time-embedding MLP at the head of the UNet plan: its output is first read by a conv epilogue
    // ~10 launches later). They run on `aux` between two events, so inside a captured graph they become a parallel branch
    // instead of 36 us at the head of the dependent chain.
    size_t fork_begin = 0, fork_end = 0, join_before = 0;
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int enable_fork(size_t begin, size_t end, size_t join) {
        if (end <= begin || join < end || join >= steps.size()) return PD_OK;
        PD_CUDA(cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking));
        PD_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        PD_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        fork_begin = begin; fork_end = end; join_before = join;
        return PD_OK;
    }
    // Second lane: steps added while `cur_lane` is 1 run on the stream `side` - a chain that only meets the main chain at
    // explicit points (the global vectors' projections / FFN beside the token grid's, unet.cu). mark() names the step added
    // last, wait(m) makes the step added next wait for mark m (an event recorded after the marked step on its lane's stream);
    // the side lane joins the main one at the end of run(). Inside a captured graph these become parallel branches; the
    // traced / profiled passes run every step on one stream in plan order, which satisfies every dependency.
    std::vector<uint8_t> lanes;                    // per step
    std::vector<int> mark_of;                      // per step: mark recorded after it, or -1
    std::vector<std::vector<int>> waits;           // per step: marks it waits for
    std::vector<cudaEvent_t> mark_ev;              // per mark
    std::vector<int> pending_waits;                // for the step added next
    uint8_t cur_lane = 0;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_side_join = nullptr;
    void lane(int l) { cur_lane = (uint8_t)l; }
    int mark() {
        if (mark_of.back() < 0) {
            mark_of.back() = (int)mark_ev.size();
            mark_ev.push_back(nullptr);
        }
        return mark_of.back();
    }
    void wait(int m) { if (m >= 0) pending_waits.push_back(m); }
    int enable_lanes() {
        bool any = false;
        for (auto l : lanes) any = any || l;
        if (!any) return PD_OK;
        PD_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
        PD_CUDA(cudaEventCreateWithFlags(&ev_side_join, cudaEventDisableTiming));
        for (auto& e : mark_ev) PD_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        return PD_OK;
    }
    int run(cudaStream_t st) const {
        if (!aux && !side) {
            for (const auto& s : steps) PD_TRY(s(st));
            return PD_OK;
        }
        bool side_used = false;
        for (size_t i = 0; i < steps.size(); ++i) {
            if (aux && i == fork_begin) {
                PD_CUDA(cudaEventRecord(ev_fork, st));
                PD_CUDA(cudaStreamWaitEvent(aux, ev_fork, 0));
            }
            if (aux && i >= fork_begin && i < fork_end) {
                PD_TRY(steps[i](aux));
                if (i + 1 == fork_end) PD_CUDA(cudaEventRecord(ev_join, aux));
                continue;
            }
            if (aux && i == join_before) PD_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
            cudaStream_t s = (side && lanes[i]) ? side : st;
            if (side) {
                for (int m : waits[i]) PD_CUDA(cudaStreamWaitEvent(s, mark_ev[m], 0));
                side_used = side_used || lanes[i];
            }
            PD_TRY(steps[i](s));
            if (side && mark_of[i] >= 0) PD_CUDA(cudaEventRecord(mark_ev[mark_of[i]], s));
        }
        if (side_used) {
            PD_CUDA(cudaEventRecord(ev_side_join, side));
            PD_CUDA(cudaStreamWaitEvent(st, ev_side_join, 0));
        }
        return PD_OK;
    }
    void add(Step s, StepKind k = STEP_KERNEL, const char* label = "") {
        steps.push_back(std::move(s));
        kinds.push_back(k);
        labels.push_back(scope.empty() ? std::string(label) : scope + "." + label);
        flops.push_back(0.0);
        lanes.push_back(cur_lane);
        mark_of.push_back(-1);
        waits.push_back(std::move(pending_waits));
        pending_waits.clear();
    }
    void add(Step s, const char* label) { add(std::move(s), STEP_KERNEL, label); }
    // GEMM-family steps in launch order, for the L2 weight prefetch chain (common.cuh): `own` = the step's weight ranges,
    // `set_next` stores the ranges the step should request for its successor.
    struct WeightUser {
        std::function<void(const WRange&)> set_next;
        WRange own;
    };
    std::vector<WeightUser> weight_users;
    void add_gemm(const GemmOp& op, const char* label = "gemm") {
        gemm_flops += op.flops;
        ++n_gemm;
        auto sp = std::make_shared<GemmOp>(op);
        add([sp](cudaStream_t st) { return gemm_launch(*sp, st); }, STEP_GEMM, label);
        flops.back() = op.flops;
        weight_users.push_back({[sp](const WRange& r) { sp->p.pf = r; }, op.own_w});
    }
    void add_ffn_fused(const FfnFusedOp& op, const char* label) {
        auto sp = std::make_shared<FfnFusedOp>(op);
        add([sp](cudaStream_t st) { return ffn_fused_launch(*sp, st); }, STEP_GEMM, label);
        weight_users.push_back({[sp](const WRange& r) { ffn_fused_set_prefetch(sp.get(), r); }, ffn_fused_weights(op)});
    }
    void add_ffn_cluster(const FfnClusterOp& op, const char* label) {
        auto sp = std::make_shared<FfnClusterOp>(op);
        add([sp](cudaStream_t st) { return ffn_cluster_launch(*sp, st); }, STEP_GEMM, label);
        weight_users.push_back({[](const WRange&) {}, ffn_cluster_weights(op)});   // it is prefetched for, it prefetches nothing
    }
    void add_qkv_attn(const QkvAttnOp& op, const char* label) {
        const double fl = qkv_attn_flops(op);
        gemm_flops += fl;
        ++n_gemm;
        auto sp = std::make_shared<QkvAttnOp>(op);
        add([sp](cudaStream_t st) { return qkv_attn_launch(*sp, st); }, STEP_GEMM, label);
        flops.back() = fl;
        weight_users.push_back({[sp](const WRange& r) { qkv_attn_set_prefetch(sp.get(), r); }, qkv_attn_weights(op)});
    }
    // Every GEMM-family step requests the weights of the next one (the last wraps around to the first of the next pass).
    void link_prefetch() {
        const size_t n = weight_users.size();
        for (size_t i = 0; i < n; ++i) weight_users[i].set_next(weight_users[(i + 1) % n].own);
    }
    // One eager pass with a %globaltimer stamp kernel after every step: ns[i] = stamp after step i (ns[0] = start),
    // so ns[i+1] - ns[i] = duration of step i + one (constant) stamp-kernel slot. `ns` has steps.size() + 1 slots.
    int run_traced(cudaStream_t st, unsigned long long* ns_dev) const {
        PD_TRY(stamp_globaltimer(ns_dev, st));
        for (size_t i = 0; i < steps.size(); ++i) {
            PD_TRY(steps[i](st));
            PD_TRY(stamp_globaltimer(ns_dev + i + 1, st));
        }
        return PD_OK;
    }
    int num_kernels() const {
        int n = 0;
        for (auto k : kinds) n += k != STEP_NONE;
        return n;
    }
    // One eager pass with every launch bracketed by CUDA events on the launching stream; per-class device time.
    int run_profiled(cudaStream_t st, PlanProfile* out) const {
        const size_t n = steps.size();
        std::vector<cudaEvent_t> ev(n + 1);
        for (auto& e : ev) PD_CUDA(cudaEventCreate(&e));
        int rc = PD_OK;
        PD_CUDA(cudaEventRecord(ev[0], st));
        for (size_t i = 0; i < n && rc == PD_OK; ++i) {
            rc = steps[i](st);
            cudaEventRecord(ev[i + 1], st);
        }
        cudaStreamSynchronize(st);
        if (rc == PD_OK) {
            *out = PlanProfile();
            out->gemm_flops = gemm_flops;
            for (size_t i = 0; i < n; ++i) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
                if (kinds[i] == STEP_GEMM) { out->gemm_ms += ms; ++out->n_gemm; }
                else if (kinds[i] == STEP_KERNEL) { out->other_ms += ms; ++out->n_other; }
            }
        }
        for (auto& e : ev) cudaEventDestroy(e);
        return rc;
    }
};

inline std::string strf(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
inline std::string strf(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return buf;
}

}  // namespace pd
