// solarosc_comm.hpp -- multi-GPU sharding behind the C ABI: one process per GPU, every rank evaluates its contiguous slice of
// the parameter batch and the per-point lnL (and flag) vectors are all-gathered device-to-device with NCCL over NVLink.
//
// Every parameter point is independent (total_llh has no cross-point state, likelihood_main.jl:30-35), so the gather is the
// only collective of the path: 12 bytes per point.  A host that runs one process per GPU without
// torch (Julia with MPI.jl, a plain C++ driver) gets it from the library: rank 0 asks for a unique id, distributes its 128
// bytes by whatever means it has, every rank calls solarosc_comm_init.  NCCL is bound at run time (dlopen of libnccl.so.2):
// the library links against the CUDA runtime only and loads where NCCL is absent.  Included into solarosc_b200.cu.
#pragma once
#include <dlfcn.h>
#include <memory>
#include <nccl.h>

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
        auto sym = [&](const char* s) { return dlsym(api.lib, s); };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.CommAbort = reinterpret_cast<decltype(api.CommAbort)>(sym("ncclCommAbort"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.GroupStart && api.GroupEnd && api.CommDestroy && api.GetErrorString;
    });
    return api;
}

#define NCK(call)                                                                                             \
    do {                                                                                                      \
        ncclResult_t r_ = (call);                                                                             \
        if (r_ != ncclSuccess) {                                                                              \
            char buf_[512];                                                                                   \
            snprintf(buf_, sizeof buf_, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, nccl_api().GetErrorString(r_)); \
            throw CudaFail{buf_};                                                                             \
        }                                                                                                     \
    } while (0)

// rows [lo, hi) of rank r: whole groups, dealt as evenly as possible, earlier ranks first
void shard_bounds(size_t n_total, size_t group, int rank, int world, size_t* lo, size_t* hi) {
    const size_t n_groups = n_total / group, base = n_groups / world, extra = n_groups % world;
    const size_t a = (size_t)rank * base + std::min<size_t>(rank, extra);
    *lo = a * group;
    *hi = (a + base + ((size_t)rank < extra ? 1 : 0)) * group;
}

// gathered[world][width] (padded slices) -> out[n_total]
// (any_flags, when given, receives the OR of all flags: the host then knows about a negative-rate error without walking
// n_total words)
__global__ void unpad_gather_kernel(const double* __restrict__ g_lnl, const uint32_t* __restrict__ g_fl, size_t width, size_t n_total,
                                    size_t group, int world, double* __restrict__ out_lnl, uint32_t* __restrict__ out_fl,
                                    uint32_t* __restrict__ any_flags) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (any_flags) {
        uint32_t f = 0;
        if (i < n_total) {
            const size_t n_groups = n_total / group, base = n_groups / world, extra = n_groups % world, gi = i / group;
            const size_t cut = extra * (base + 1);
            const size_t r = gi < cut ? gi / (base + 1) : extra + (base ? (gi - cut) / base : 0);
            f = g_fl[r * width + (i - (r * base + (r < extra ? r : extra)) * group)];
        }
        f = __reduce_or_sync(0xffffffffu, f);
        if ((threadIdx.x & 31) == 0 && f) atomicOr(any_flags, f);
    }
    if (i >= n_total) return;
    const size_t n_groups = n_total / group, base = n_groups / world, extra = n_groups % world;
    const size_t gi = i / group;
    // rank owning group gi: the first `extra` ranks hold base + 1 groups
    const size_t cut = extra * (base + 1);
    const size_t r = gi < cut ? gi / (base + 1) : extra + (base ? (gi - cut) / base : 0);
    const size_t lo = (r * base + (r < extra ? r : extra)) * group;
    out_lnl[i] = g_lnl[r * width + (i - lo)];
    if (out_fl) out_fl[i] = g_fl[r * width + (i - lo)];
}

}  // namespace

struct solarosc_comm_state {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    double *d_send = nullptr, *d_recv = nullptr;
    uint32_t *d_fsend = nullptr, *d_frecv = nullptr;
    size_t width = 0;
};

namespace {

void comm_reserve(solarosc_handle* h, size_t width) {
    solarosc_comm_state* c = h->comm;
    if (width <= c->width) return;
    for (void* p : {(void*)c->d_send, (void*)c->d_recv, (void*)c->d_fsend, (void*)c->d_frecv})
        if (p) CK(cudaFree(p));
    c->d_send = c->d_recv = nullptr; c->d_fsend = c->d_frecv = nullptr; c->width = 0;
    CK(cudaMalloc(reinterpret_cast<void**>(&c->d_send), width * sizeof(double)));
    CK(cudaMalloc(reinterpret_cast<void**>(&c->d_recv), width * c->world * sizeof(double)));
    CK(cudaMalloc(reinterpret_cast<void**>(&c->d_fsend), width * sizeof(uint32_t)));
    CK(cudaMalloc(reinterpret_cast<void**>(&c->d_frecv), width * c->world * sizeof(uint32_t)));
    c->width = width;
}

// device-resident core: this rank's slice is evaluated into the send buffers, gathered, unpadded into d_lnl_all [n_total]
// h_params_local != nullptr: the rows are still on the host; chunk c + 1 is copied into d_params_local on the handle's copy
// stream while the kernels of chunk c run (the caller's buffer is read asynchronously when it is pinned)
void eval_sharded_device(solarosc_handle* h, const double* d_params_local, size_t n_total, size_t group, double* d_lnl_all,
                         uint32_t* d_flags_all, cudaStream_t st, const double* h_params_local = nullptr, uint32_t* d_any = nullptr) {
    solarosc_comm_state* c = h->comm;
    size_t lo, hi;
    shard_bounds(n_total, group, c->rank, c->world, &lo, &hi);
    const size_t n_local = hi - lo;
    size_t w0, w1;
    shard_bounds(n_total, group, 0, c->world, &w0, &w1);           // rank 0 holds the widest slice
    const size_t width = std::max<size_t>(w1 - w0, 1);
    comm_reserve(h, width);
    const size_t step = ((size_t)h->chunk / group) * group;
    h->tev_used = 0;
    h->tev_stream = st;
    for (size_t b0 = 0; b0 < n_local; b0 += step) {
        const int nb = (int)std::min(step, n_local - b0);
        if (h_params_local) {
            CK(cudaMemcpyAsync(const_cast<double*>(d_params_local) + b0 * h->P, h_params_local + b0 * h->P, (size_t)nb * h->P * sizeof(double),
                               cudaMemcpyHostToDevice, h->s_in));
            CK(cudaEventRecord(h->ev_in[0], h->s_in));
            CK(cudaStreamWaitEvent(st, h->ev_in[0], 0));
        }
        RatesOut O{};
        O.lnl = c->d_send + b0; O.flags = c->d_fsend + b0;
        if (h->timing) CK(cudaEventRecord(h->next_tev(), st));
        cudaEvent_t mid = h->timing ? h->next_tev() : nullptr;
        h->last_fused = launch_llh(h, d_params_local + b0 * h->P, nb, (int)group, O, st, mid);
        if (h->timing) CK(cudaEventRecord(h->next_tev(), st));
    }
    NCK(nccl_api().GroupStart());
    NCK(nccl_api().AllGather(c->d_send, c->d_recv, width, ncclDouble, c->comm, st));
    NCK(nccl_api().AllGather(c->d_fsend, c->d_frecv, width, ncclUint32, c->comm, st));
    NCK(nccl_api().GroupEnd());
    if (d_any) CK(cudaMemsetAsync(d_any, 0, sizeof(uint32_t), st));
    unpad_gather_kernel<<<(unsigned)((n_total + 255) / 256), 256, 0, st>>>(c->d_recv, c->d_frecv, width, n_total, group, c->world, d_lnl_all,
                                                                           d_flags_all, d_any);
    CK(cudaGetLastError());
    ++h->launches;
}

}  // namespace

extern "C" {

int solarosc_shard_bounds(size_t n_total, size_t group, int rank, int world, size_t* lo, size_t* hi) {
    if (!lo || !hi || group == 0 || n_total % group || world < 1 || rank < 0 || rank >= world) return SOLAROSC_ERR_ARG;
    shard_bounds(n_total, group, rank, world, lo, hi);
    return SOLAROSC_OK;
}

int solarosc_comm_unique_id(void* id_out) {
    if (!id_out) return SOLAROSC_ERR_ARG;
    if (!nccl_api().ok) return SOLAROSC_ERR_CUDA;
    ncclUniqueId id;
    if (nccl_api().GetUniqueId(&id) != ncclSuccess) return SOLAROSC_ERR_CUDA;
    static_assert(sizeof(ncclUniqueId) == SOLAROSC_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    std::memcpy(id_out, &id, sizeof id);
    return SOLAROSC_OK;
}

int solarosc_comm_init(solarosc_handle* h, const void* id, int rank, int world) {
    return guarded(h, [&]() {
        REQUIRE(id && world >= 1 && rank >= 0 && rank < world, "bad communicator arguments");
        REQUIRE(!h->comm, "the handle already has a communicator");
        if (!nccl_api().ok) throw CudaFail{"libnccl.so.2 could not be loaded (dlopen): multi-GPU sharding needs NCCL"};
        ncclUniqueId uid;
        std::memcpy(&uid, id, sizeof uid);
        auto* c = new solarosc_comm_state();
        c->rank = rank; c->world = world;
        const ncclResult_t r = nccl_api().CommInitRank(&c->comm, world, uid, rank);
        if (r != ncclSuccess) {
            const std::string msg = std::string("ncclCommInitRank failed: ") + nccl_api().GetErrorString(r);
            delete c;
            throw CudaFail{msg};
        }
        h->comm = c;
        return (int)SOLAROSC_OK;
    });
}

}  // extern "C"

// Frees the communicator.  collective = true (solarosc_comm_destroy: every rank calls it, like solarosc_comm_init):
// ncclCommDestroy, which flushes and may wait for the peers.  collective = false (a handle destroyed with its
// communicator still attached, e.g. at interpreter exit of one rank): ncclCommAbort, which never waits for anybody --
// a rank that leaves early must not hang the others or itself.
int comm_release(solarosc_handle* h, bool collective) {
    return guarded(h, [&]() {
        if (!h->comm) return (int)SOLAROSC_OK;
        CK(cudaStreamSynchronize(h->stream));
        solarosc_comm_state* c = h->comm;
        if (c->comm) {
            if (collective || !nccl_api().CommAbort) nccl_api().CommDestroy(c->comm);
            else nccl_api().CommAbort(c->comm);
        }
        for (void* p : {(void*)c->d_send, (void*)c->d_recv, (void*)c->d_fsend, (void*)c->d_frecv})
            if (p) cudaFree(p);
        delete c;
        h->comm = nullptr;
        return (int)SOLAROSC_OK;
    });
}
extern "C" {
int solarosc_comm_destroy(solarosc_handle* h) { return comm_release(h, true); }

int solarosc_eval_llh_sharded_device(solarosc_handle* h, const double* d_params_local, size_t n_total, size_t group, double* d_lnl_all,
                                     uint32_t* d_flags_all, void* stream) {
    return guarded(h, [&]() {
        REQUIRE(h->comm, "no communicator: call solarosc_comm_init first");
        if (group == 0) group = 1;
        REQUIRE(n_total % group == 0 && group <= (size_t)h->chunk, "n_total must be a multiple of group (<= max_batch)");
        if (n_total == 0) return (int)SOLAROSC_OK;
        REQUIRE(d_lnl_all, "d_lnl_all must not be NULL");
        check_data(h, -1);
        h->launches = 0;
        eval_sharded_device(h, d_params_local, n_total, group, d_lnl_all, d_flags_all, stream ? static_cast<cudaStream_t>(stream) : h->stream);
        return (int)SOLAROSC_OK;
    });
}

int solarosc_eval_llh_sharded(solarosc_handle* h, const double* params_local, size_t n_total, size_t group, double* out_lnl,
                              uint32_t* out_flags) {
    return guarded(h, [&]() {
        REQUIRE(h->comm, "no communicator: call solarosc_comm_init first");
        if (group == 0) group = 1;
        REQUIRE(n_total % group == 0 && group <= (size_t)h->chunk, "n_total must be a multiple of group (<= max_batch)");
        if (n_total == 0) return (int)SOLAROSC_OK;
        REQUIRE(out_lnl, "out_lnl must not be NULL");
        check_data(h, -1);
        size_t lo, hi;
        shard_bounds(n_total, group, h->comm->rank, h->comm->world, &lo, &hi);
        REQUIRE(hi == lo || params_local, "params_local must not be NULL");
        // this rank's rows to the device, the gathered vectors back: [params | lnl | flags] in the scratch
        const size_t n_local = hi - lo;
        h->ensure_scratch((n_local * h->P + n_total) * sizeof(double) + (n_total + 2) * sizeof(uint32_t));
        double* d_p = h->d_scratch;
        double* d_l = d_p + n_local * h->P;
        uint32_t* d_f = reinterpret_cast<uint32_t*>(d_l + n_total);
        uint32_t* d_any = d_f + n_total;
        h->launches = 0;
        CK(cudaStreamSynchronize(h->stream));   // (the scratch may just have been reallocated; nothing of an earlier call is in flight)
        eval_sharded_device(h, d_p, n_total, group, d_l, d_f, h->stream, params_local, d_any);
        CK(cudaMemcpyAsync(out_lnl, d_l, n_total * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        // flags straight into the caller's array when it wants them; whether any point carries a negative-rate error comes back
        // as one word (with 8 ranks every rank receives 12 bytes for each of the n_total points: no host pass over them)
        if (out_flags) CK(cudaMemcpyAsync(out_flags, d_f, n_total * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        uint32_t any = 0;
        CK(cudaMemcpyAsync(&any, d_any, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        const bool neg = (any & SOLAROSC_FLAG_NEG_ERROR) != 0;
        if (neg) {
            h->err = "an expected rate was <= -1e-3 for at least one point (lnL = NaN there)";
            return (int)SOLAROSC_ERR_NEGATIVE_RATE;
        }
        return (int)SOLAROSC_OK;
    });
}

}  // extern "C"
