// solarosc_b200.cu -- host side of the C ABI declared in include/solarosc_b200.h.
//
// Owns device copies of the constant tables, a stream, pinned staging buffers and the
// oscillation-probability workspace; every eval call stages the caller's host parameter block in
// chunks, launches osc_kernel then rates_kernel per chunk and copies the results back.
// There is deliberately no CPU fallback: without an sm_100 device create fails.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "solarosc_adjoint.cuh"
#include "solarosc_pipe.cuh"
#include "solarosc_general.cuh"
#include "solarosc_gengrad.cuh"
#include "solarosc_setup.cuh"

using namespace solarosc;

namespace {

thread_local std::string g_create_error;

struct CudaFail {
    std::string msg;
};

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) {                                                                              \
            char buf_[512];                                                                                   \
            snprintf(buf_, sizeof buf_, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            throw CudaFail{buf_};                                                                             \
        }                                                                                                     \
    } while (0)

struct ArgFail {
    std::string msg;
};
#define REQUIRE(cond, text)            \
    do {                               \
        if (!(cond)) throw ArgFail{text}; \
    } while (0)

}  // namespace

struct solarosc_comm_state;
struct solarosc_handle;
int comm_release(solarosc_handle* h, bool collective);   // solarosc_comm.hpp

struct solarosc_handle {
    solarosc_config cfg{};
    solarosc_comm_state* comm = nullptr;   // NCCL communicator of the sharded calls (solarosc_comm.hpp)
    DevTables T{};
    int device = 0, sm_count = 148, chunk = 4096, P = 0;
    int max_ereco = 0;
    int n_es[kMaxDet]{}, n_cc[kMaxDet]{};
    int n_cos[kMaxDet]{};        // cos-scatter bins of the ES arrays (1 unless the detector uses the angular reconstruction)
    bool fold_fine_ok = false;   // every detector has a fine work list for the grouped fold
    bool gen = false;            // some detector runs an inclusive / semi-inclusive / angular branch: gen_kernel serves every call
    cudaStream_t stream = nullptr;       // compute
    cudaStream_t s_in = nullptr, s_out = nullptr;  // H2D / D2H of the host-buffer pipeline
    cudaEvent_t ev_in[2]{}, ev_done[2]{}, ev_out[2]{};
    std::vector<cudaEvent_t> tev;        // timing events, three per chunk (before osc, between, after rates)
    size_t tev_used = 0;
    cudaStream_t tev_stream = nullptr;
    bool timing = false;
    std::vector<void*> dev_allocs;
    // per-detector data buffers (device, [1+n_cosz][n_ereco]) kept so set_data can overwrite them
    double *d_obs_es[kMaxDet]{}, *d_obs_cc[kMaxDet]{}, *d_log_es[kMaxDet]{}, *d_log_cc[kMaxDet]{};
    double2 *d_ol_es[kMaxDet]{}, *d_ol_cc[kMaxDet]{};   // (m, log m) interleaved: one 16-byte load per Poisson term
    double2* d_ol_ang[kMaxDet]{};                        // angular ES data, [1 + n_cosz][n_ereco_es][n_cos]
    bool has_data[kMaxDet]{};
    // workspaces
    double* d_params2[2]{};      // double-buffered parameter chunks
    double* d_params = nullptr;  // == d_params2[0]
    double* d_osc = nullptr;     // stage-separated probabilities [chunk][2][1 + n_cosz][n_etrue]; use osc()
    size_t osc_doubles = 0;
    double* d_lnl2[2]{};
    uint32_t* d_flags2[2]{};
    double* h_lnl2[2]{};         // pinned staging of the results of one chunk, double-buffered
    // small batches (MCMC steps of a few chains) replay one captured CUDA graph: H2D, the kernels and both D2H copies
    // become a single launch.  Keyed by (n, group); inputs go through a pinned staging row block.
    double* h_params_pin = nullptr;
    cudaGraphExec_t small_graph = nullptr;
    size_t small_n = 0, small_group = 0;
    int small_launches = 0;
    bool small_fused = false;
    uint32_t* h_flags2[2]{};
    double* d_scratch = nullptr; // rate / per-bin / p1e outputs, grown on demand
    double* h_dstage[2]{};       // pinned staging of the derived-mode per-sample outputs
    size_t h_dstage_bytes = 0;
    double* d_gradws = nullptr;  // gradient path, two slots: probabilities, d lnL / d P, gradient rows
    size_t gradws_bytes = 0;
    size_t grad_smem = 0, adj_smem = 0, gengrad_smem = 0;
    double* h_grad_pin[2]{};     // pinned staging of the gradient rows of one chunk
    bool grad_ok = false;
    size_t scratch_bytes = 0;
    size_t osc_smem_limit = 200 * 1024;
    size_t fused_smem = 0, seg_bytes = 0;
    int fused_cb = 0, fused_grid = 0;    // coarse energy bins per oscillation chunk, CTAs of the persistent fused kernel
    bool fused_seg = false;
    bool rates_ok = true;
    bool fused_ok = false;
    // binnings whose probability tile exceeds shared memory: row-chunked fold (rates_tiled_kernel)
    int tiled_rows = 0, tiled_chunks = 0;
    size_t tiled_smem = 0;
    double* d_tiled_part = nullptr;
    uint32_t* d_tiled_flags = nullptr;
    std::mutex mu;
    std::string err;
    float ms_osc = 0.f, ms_rates = 0.f;
    int launches = 0;
    bool use_fused = true;
    bool last_fused = false;
    // pipe_kernel (solarosc_pipe.cuh): ticket pattern and fold items per (mode, group, detector), built on first use
    struct PipePlan {
        int mode, group, det_only;
        bool ok;
        PipeLaunch L;
        size_t smem;
    };
    std::vector<PipePlan> pipe_plans;
    bool use_pipe = true, pipe_all = false;
    bool fine_all = false;      // SOLAROSC_FINE_ALL: the fine fold arrangement for independent points too (launch-shape experiments)
    int n_mirror = 0;           // paths that take the mirrored pair walk
    size_t smem_optin = 0;

    template <class Tp>
    Tp* dalloc(size_t n) {
        void* p = nullptr;
        CK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(Tp)));
        dev_allocs.push_back(p);
        return static_cast<Tp*>(p);
    }
    template <class Tp>
    Tp* upload(const Tp* src, size_t n) {
        Tp* d = dalloc<Tp>(n);
        if (n) CK(cudaMemcpy(d, src, n * sizeof(Tp), cudaMemcpyHostToDevice));
        return d;
    }
    double* upload_vec(const std::vector<double>& v) { return upload(v.data(), v.size()); }
    // workspace of the stage-separated path (1 GiB at the default max_batch): allocated when a call first takes that path
    double* osc() {
        if (!d_osc) d_osc = dalloc<double>(osc_doubles);
        return d_osc;
    }
    cudaEvent_t next_tev() {
        if (tev_used == tev.size()) {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            tev.push_back(e);
        }
        return tev[tev_used++];
    }
    void ensure_scratch(size_t bytes) {
        if (bytes <= scratch_bytes) return;
        if (d_scratch) CK(cudaFree(d_scratch));
        d_scratch = nullptr;
        scratch_bytes = 0;
        CK(cudaMalloc(reinterpret_cast<void**>(&d_scratch), bytes));
        scratch_bytes = bytes;
    }
};

namespace {

// [a][b] table given in either layout -> row-major copy (b fastest), zero-padded to na_pad rows
std::vector<double> to_row_major(const double* src, int na, int nb, bool col_major, int na_pad = 0) {
    std::vector<double> out((size_t)std::max(na, na_pad) * nb, 0.0);
    for (int a = 0; a < na; ++a)
        for (int b = 0; b < nb; ++b) out[(size_t)a * nb + b] = col_major ? src[(size_t)b * na + a] : src[(size_t)a * nb + b];
    return out;
}

void upload_data(solarosc_handle* h, int d, const double* es_day, const double* es_night, const double* cc_day,
                 const double* cc_night) {
    const int nZ = h->T.n_cosz;
    const bool cm = h->cfg.col_major != 0;
    auto pack = [&](const double* day, const double* night, int nR, double* d_obs, double* d_log, double2* d_ol) {
        std::vector<double> obs((size_t)(1 + nZ) * nR), lg(obs.size()), ol(2 * obs.size());
        std::copy(day, day + nR, obs.begin());
        std::vector<double> nt = to_row_major(night, nZ, nR, cm);
        std::copy(nt.begin(), nt.end(), obs.begin() + nR);
        for (size_t i = 0; i < obs.size(); ++i) {
            lg[i] = obs[i] > 0.0 ? std::log(obs[i]) : 0.0;
            ol[2 * i] = obs[i]; ol[2 * i + 1] = lg[i];
        }
        CK(cudaMemcpy(d_obs, obs.data(), obs.size() * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_log, lg.data(), lg.size() * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_ol, ol.data(), ol.size() * sizeof(double), cudaMemcpyHostToDevice));
    };
    if (h->T.det[d].ang) {
        // reference shapes (n_cos, n_E) and (n_cos, n_E, n_cosz): [c][j], [c][j][z] row-major, [j][c], [z][j][c] as Julia
        // lays them out; device layout [row][j][c]
        const int nC = h->T.det[d].n_cos, nR = h->n_es[d];
        std::vector<double> ol((size_t)2 * (1 + nZ) * nR * nC);
        for (int row = 0; row <= nZ; ++row)
            for (int j = 0; j < nR; ++j)
                for (int c = 0; c < nC; ++c) {
                    double m;
                    if (row == 0) m = cm ? es_day[(size_t)j * nC + c] : es_day[(size_t)c * nR + j];
                    else m = cm ? es_night[((size_t)(row - 1) * nR + j) * nC + c] : es_night[((size_t)c * nR + j) * nZ + (row - 1)];
                    const size_t o = ((size_t)row * nR + j) * nC + c;
                    ol[2 * o] = m; ol[2 * o + 1] = m > 0.0 ? std::log(m) : 0.0;
                }
        CK(cudaMemcpy(h->d_ol_ang[d], ol.data(), ol.size() * sizeof(double), cudaMemcpyHostToDevice));
        // Poisson terms of the cos-scatter bins whose expectation is identically 0 (likelihood_statistical.jl:16-24 at e = 0)
        DevDetector& D = h->T.det[d];
        std::vector<bool> live(nC, false);
        for (int k = 0; k < D.n_cos_live; ++k) live[D.cos_live[k]] = true;
        const double e0 = 0.5 * (0.0 + 1e-9 + std::sqrt((0.0 - 1e-9) * (0.0 - 1e-9) + 1e-18));
        double acc = 0.0;
        for (int row = 0; row <= nZ; ++row)
            for (int j = D.index_es - 1; j < nR; ++j)
                for (int c = 0; c < nC; ++c) {
                    if (live[c]) continue;
                    const double m = ol[2 * (((size_t)row * nR + j) * nC + c)];
                    acc += m > 0.0 ? e0 - m + m * (std::log(m) - std::log(e0)) : e0;
                }
        D.llh_dead_const = acc;
    } else {
        pack(es_day, es_night, h->n_es[d], h->d_obs_es[d], h->d_log_es[d], h->d_ol_es[d]);
    }
    pack(cc_day, cc_night, h->n_cc[d], h->d_obs_cc[d], h->d_log_cc[d], h->d_ol_cc[d]);
    h->has_data[d] = true;
}

void build(solarosc_handle* h, const solarosc_tables* t, const solarosc_config* c) {
    REQUIRE(t && c, "tables and config must not be NULL");
    REQUIRE(c->abi_version == SOLAROSC_ABI_VERSION, "config.abi_version does not match this library");
    REQUIRE(t->n_etrue > 0 && t->n_cosz > 0 && t->n_cosz_sub > 0, "empty binning");
    REQUIRE(t->n_layers > 0 && t->n_layers <= SOLAROSC_MAX_LAYERS, "n_layers out of range");
    REQUIRE(t->n_det > 0 && t->n_det <= kMaxDet, "n_det out of range");
    REQUIRE(t->e_calc && t->seg_start && t->seg_layer && t->seg_length && t->earth_lookup && t->exposure_weights && t->det,
            "NULL table pointer");
    REQUIRE(c->fast || (t->n_solar > 0 && t->solar_ne && t->solar_prod_8b && t->solar_prod_hep),
            "fast == 0 needs the solar production-region tables");
    h->cfg = *c;
    h->device = c->device;
    h->use_fused = std::getenv("SOLAROSC_NO_FUSED") == nullptr;   // A/B switch for profiling
    int n_dev = 0;
    CK(cudaGetDeviceCount(&n_dev));
    if (c->device < 0 || c->device >= n_dev) throw CudaFail{"no such CUDA device"};
    CK(cudaSetDevice(c->device));
    cudaDeviceProp prop{};
    CK(cudaGetDeviceProperties(&prop, c->device));
    if (prop.major != 10) throw CudaFail{"libsolarosc_b200 holds sm_100a code only; this device is not compute capability 10.x"};
    h->sm_count = prop.multiProcessorCount;
    h->osc_smem_limit = std::min<size_t>(200 * 1024, prop.sharedMemPerBlockOptin);
    h->smem_optin = prop.sharedMemPerBlockOptin;
    h->use_pipe = std::getenv("SOLAROSC_NO_PIPE") == nullptr;   // A/B switches: the lock-step llh_kernel instead of pipe_kernel,
    h->pipe_all = std::getenv("SOLAROSC_PIPE_ALL") != nullptr;  // and pipe_kernel for independent points too
    h->fine_all = std::getenv("SOLAROSC_FINE_ALL") != nullptr;
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CK(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
    }

    DevTables& T = h->T;
    T.n_etrue = t->n_etrue; T.n_e = 2 * t->n_etrue; T.n_cosz = t->n_cosz; T.n_sub = t->n_cosz_sub;
    T.n_paths = t->n_cosz * t->n_cosz_sub; T.n_layers = t->n_layers; T.n_solar = c->fast ? 0 : t->n_solar;
    T.n_det = t->n_det; T.fast = c->fast != 0; T.earth_unc = c->earth_uncertainty != 0; T.night_only = c->night_only != 0;
    const int n_seg = t->seg_start[T.n_paths];
    REQUIRE(t->seg_start[0] == 0 && n_seg > 0, "seg_start must start at 0");
    T.n_seg = n_seg;
    T.e_calc = h->upload(t->e_calc, T.n_e);
    // Applied order of the steps of each path (reference osc.jl:436-446 and :455-456): step m uses
    // layer index[n-1-m] paired with length[m] of the stored detector->surface segment list.
    std::vector<int> app_layer(n_seg);
    std::vector<double> app_len(n_seg);
    std::vector<int> row_cost(T.n_cosz, 0);
    for (int pth = 0; pth < T.n_paths; ++pth) {
        const int s0 = t->seg_start[pth], n = t->seg_start[pth + 1] - s0;
        REQUIRE(n > 0, "a path has no segments");
        for (int m = 0; m < n; ++m) {
            const int lay = t->seg_layer[s0 + n - 1 - m];
            REQUIRE(lay >= 1 && lay <= t->n_layers, "segment layer id out of range");
            app_layer[s0 + m] = lay - 1;
            app_len[s0 + m] = t->seg_length[s0 + m];
        }
        row_cost[pth / T.n_sub] += n;
    }
    T.n_ltbl = 1 + *std::max_element(app_layer.begin(), app_layer.end());
    std::vector<Seg> seg(n_seg);
    for (int i = 0; i < n_seg; ++i) seg[i] = Seg{app_len[i], app_layer[i], 0};
    // Mirrored paths (solarosc_device.cuh): flagged on their first step, with the phase-rotation constants of the
    // interior steps of their first half.  The bound on the phase rates (quarter turns per km) covers every eigenvalue
    // gap of S = diag(0, dm21, dm31) + a u u^T with dm21 <= 1e-3 and earth norms <= 2 (beyond that the likelihood is
    // -Inf anyway): gap <= dm31 + dm21 + a.  SOLAROSC_NO_MIRROR: A/B switch.
    std::vector<double2> mir(n_seg, make_double2(0.0, 0.0));
    h->n_mirror = 0;
    if (std::getenv("SOLAROSC_NO_MIRROR") == nullptr) {
        const double e_min = *std::min_element(t->e_calc, t->e_calc + T.n_e);
        const double rho_max = *std::max_element(t->earth_lookup, t->earth_lookup + T.n_layers);
        const double kq_max = kPhaseQT * ((kDm31 + 1e-3) / e_min + 2.0 * rho_max);
        for (int pth = 0; pth < T.n_paths; ++pth) {
            const int s0 = t->seg_start[pth], s1 = t->seg_start[pth + 1];
            if (!(e_min > 0.0) || !mirror_eligible(app_layer.data(), app_len.data(), s0, s1, kq_max)) continue;
            seg[s0].pad = 1;
            ++h->n_mirror;
            for (int i = 1; i < (s1 - s0) / 2; ++i) {
                const double dr = 1.5707963267948966 * (app_len[s1 - 1 - i] - app_len[s0 + i]);
                mir[s0 + i] = make_double2(dr, -0.5 * dr);
            }
        }
    }
    T.seg = h->upload(seg.data(), n_seg);
    T.seg_mir = h->upload(mir.data(), n_seg);
    std::vector<int> order(T.n_cosz);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return row_cost[a] > row_cost[b]; });
    T.row_order = h->upload(order.data(), T.n_cosz);
    std::vector<int2> se(T.n_paths);
    for (int zr = 0; zr < T.n_cosz; ++zr)
        for (int sub = 0; sub < T.n_sub; ++sub) {
            const int pth = order[zr] * T.n_sub + sub;
            se[zr * T.n_sub + sub] = make_int2(t->seg_start[pth], t->seg_start[pth + 1]);
        }
    T.path_se = h->upload(se.data(), T.n_paths);
    std::vector<unsigned> row_layers(T.n_cosz, 0u);
    for (int zr = 0; zr < T.n_cosz; ++zr)
        for (int sub = 0; sub < T.n_sub; ++sub) {
            const int pth = order[zr] * T.n_sub + sub;
            for (int sg = t->seg_start[pth]; sg < t->seg_start[pth + 1]; ++sg) row_layers[zr] |= 1u << app_layer[sg];
        }
    T.row_layers = h->upload(row_layers.data(), T.n_cosz);
    T.earth_lookup = h->upload(t->earth_lookup, T.n_layers);
    T.wz = h->upload(t->exposure_weights, T.n_cosz);
    T.avg_ne[0] = t->avg_ne_8b; T.avg_ne[1] = t->avg_ne_hep;
    if (!c->fast) {
        T.solar_ne = h->upload(t->solar_ne, t->n_solar);
        T.solar_w[0] = h->upload(t->solar_prod_8b, t->n_solar);
        T.solar_w[1] = h->upload(t->solar_prod_hep, t->n_solar);
        T.solar_wsum[0] = std::accumulate(t->solar_prod_8b, t->solar_prod_8b + t->n_solar, 0.0);
        T.solar_wsum[1] = std::accumulate(t->solar_prod_hep, t->solar_prod_hep + t->n_solar, 0.0);
    }
    int off = SOLAROSC_N_FIXED_PARAMS + (T.earth_unc ? T.n_layers : 0);
    const bool cm = c->col_major != 0;
    for (int d = 0; d < t->n_det; ++d) {
        const solarosc_detector& s = t->det[d];
        DevDetector& D = T.det[d];
        REQUIRE(s.n_ereco_es > 0 && s.n_ereco_cc > 0, "detector without reco bins");
        REQUIRE(s.n_bg_es >= 0 && s.n_bg_es <= SOLAROSC_MAX_BG && s.n_bg_cc >= 0 && s.n_bg_cc <= SOLAROSC_MAX_BG,
                "too many background components");
        REQUIRE(s.index_es >= 1 && s.index_es <= s.n_ereco_es && s.index_cc >= 1 && s.index_cc <= s.n_ereco_cc,
                "threshold index out of range");
        REQUIRE(s.unosc_es_nue_8b && s.unosc_es_nuother_8b && s.unosc_cc_8b && s.unosc_es_nue_hep &&
                    s.unosc_es_nuother_hep && s.unosc_cc_hep && s.log_e_norm && s.resp_es_nue && s.resp_es_nuother &&
                    s.resp_cc && s.eff_es_nue && s.eff_es_nuother && s.eff_cc,
                "NULL detector table pointer");
        REQUIRE((s.n_bg_es == 0 || (s.raw_bg_es && s.bg_es_sides && s.bg_es_par)) && (s.n_bg_cc == 0 || (s.raw_bg_cc && s.bg_cc_par)),
                "NULL background table pointer");
        D.es_mode = s.es_mode != 0; D.cc_mode = s.cc_mode != 0;
        D.n_es = s.n_ereco_es; D.n_cc = s.n_ereco_cc; D.index_es = s.index_es; D.index_cc = s.index_cc;
        D.n_bg_es = s.n_bg_es; D.n_bg_cc = s.n_bg_cc;
        h->n_es[d] = D.n_es; h->n_cc[d] = D.n_cc;
        h->max_ereco = std::max({h->max_ereco, D.n_es, D.n_cc});
        const int nT = T.n_etrue;
        D.log_e_norm = h->upload(s.log_e_norm, nT);
        {   // constant matrices of the fold (derivation at fold_stage in solarosc_kernels.cuh)
            const std::vector<double> re = to_row_major(s.resp_es_nue, nT, D.n_es, cm), ro = to_row_major(s.resp_es_nuother, nT, D.n_es, cm);
            const std::vector<double> rc = to_row_major(s.resp_cc, nT, D.n_cc, cm);
            const int kp = 4 * fold_kt(nT), le = 8 * ((D.n_es + 7) / 8), lc = 8 * ((D.n_cc + 7) / 8);
            std::vector<double> a8((size_t)kp * le, 0.0), ah(a8), c8(le, 0.0), chh(le, 0.0), b8((size_t)kp * lc, 0.0), bh(b8);
            for (int i = 0; i < nT; ++i) {
                for (int j = 0; j < D.n_es; ++j) {
                    const double xe = s.eff_es_nue[j] * re[(size_t)i * D.n_es + j], xo = s.eff_es_nuother[j] * ro[(size_t)i * D.n_es + j];
                    a8[(size_t)i * le + j] = 0.5 * (s.unosc_es_nue_8b[i] * xe - s.unosc_es_nuother_8b[i] * xo);
                    ah[(size_t)i * le + j] = 0.5 * (s.unosc_es_nue_hep[i] * xe - s.unosc_es_nuother_hep[i] * xo);
                    c8[j] += 0.5 * s.unosc_es_nuother_8b[i] * xo;
                    chh[j] += 0.5 * s.unosc_es_nuother_hep[i] * xo;
                }
                for (int j = 0; j < D.n_cc; ++j) {
                    const double xc = 0.5 * s.eff_cc[j] * rc[(size_t)i * D.n_cc + j];
                    b8[(size_t)i * lc + j] = s.unosc_cc_8b[i] * xc;
                    bh[(size_t)i * lc + j] = s.unosc_cc_hep[i] * xc;
                }
            }
            D.cc_out = D.cc_mode; D.f_below = 1.0;
            if (const solarosc_detector_ext* x = s.ext; x && (x->inclusive != 0 || x->angular != 0)) {
                // inclusive / semi-inclusive / angular branches: extra constant matrices (solarosc_general.cuh)
                h->gen = true;
                REQUIRE(x->inclusive >= 0 && x->inclusive <= 2, "ext.inclusive must be 0, 1 or 2");
                D.incl = x->inclusive; D.ang = x->angular != 0;
                if (D.incl) {
                    REQUIRE(x->resp_cc_incl && x->eff_cc_incl, "inclusive analysis needs resp_cc_incl and eff_cc_incl");
                    REQUIRE(D.es_mode || !D.cc_mode, "inclusive analysis with ES_mode off and CC_mode on is not supported");
                    const double f_above = D.incl == 2 ? x->cc_f_above : 1.0;
                    if (D.cc_mode) {   // compute_CC_inclusive_event_rates, propagation_reco.jl:139-170
                        const std::vector<double> ri = to_row_major(x->resp_cc_incl, nT, D.n_es, cm);
                        std::vector<double> x8((size_t)kp * le, 0.0), xh(x8);
                        for (int i = 0; i < nT; ++i)
                            for (int j = 0; j < D.n_es; ++j) {
                                const double xc = f_above * (0.5 * x->eff_cc_incl[j] * ri[(size_t)i * D.n_es + j]);
                                x8[(size_t)i * le + j] = s.unosc_cc_8b[i] * xc;
                                xh[(size_t)i * le + j] = s.unosc_cc_hep[i] * xc;
                            }
                        D.xs8 = h->upload_vec(x8); D.xsh = h->upload_vec(xh);
                    }
                    if (D.incl == 1) D.cc_out = 0;   // CC arrays are zero and use_CC is false (propagation_main.jl:148-150, likelihood_main.jl:12)
                    if (D.incl == 2) {
                        D.f_below = x->cc_f_below;
                        for (auto& v : b8) v *= D.f_below;
                        for (auto& v : bh) v *= D.f_below;
                        if (D.es_mode) {   // compute_ES_misID_CC_event_rates, propagation_reco.jl:187-199
                            REQUIRE(x->resp_cc_misid_nue && x->resp_cc_misid_nuother, "semi-inclusive analysis needs the mis-ID responses");
                            const std::vector<double> mn = to_row_major(x->resp_cc_misid_nue, nT, D.n_cc, cm);
                            const std::vector<double> mo = to_row_major(x->resp_cc_misid_nuother, nT, D.n_cc, cm);
                            std::vector<double> y8((size_t)kp * lc, 0.0), yh(y8), yc8(lc, 0.0), ych(lc, 0.0);
                            for (int i = 0; i < nT; ++i)
                                for (int j = 0; j < D.n_cc; ++j) {
                                    const double xe = s.eff_cc[j] * mn[(size_t)i * D.n_cc + j], xo = s.eff_cc[j] * mo[(size_t)i * D.n_cc + j];
                                    y8[(size_t)i * lc + j] = 0.5 * (s.unosc_es_nue_8b[i] * xe - s.unosc_es_nuother_8b[i] * xo);
                                    yh[(size_t)i * lc + j] = 0.5 * (s.unosc_es_nue_hep[i] * xe - s.unosc_es_nuother_hep[i] * xo);
                                    yc8[j] += 0.5 * s.unosc_es_nuother_8b[i] * xo;
                                    ych[j] += 0.5 * s.unosc_es_nuother_hep[i] * xo;
                                }
                            D.yp8 = h->upload_vec(y8); D.yph = h->upload_vec(yh); D.yc8 = h->upload_vec(yc8); D.ych = h->upload_vec(ych);
                        }
                    }
                }
                if (D.ang) {   // compute_ES_angular_event_rates, propagation_reco.jl:66-108
                    REQUIRE(x->n_cos > 0 && x->resp_es_angular && x->resp_bg_angular && x->angular_mask, "angular analysis needs n_cos and the angular tables");
                    D.n_cos = x->n_cos;
                    const std::vector<double> ra = to_row_major(x->resp_es_angular, D.n_cos, D.n_es, cm);
                    const std::vector<double> rb = to_row_major(x->resp_bg_angular, D.n_cos, D.n_es, cm);
                    std::vector<double> am((size_t)D.n_es * D.n_cos), bm(am.size());
                    // The device applies the cut to the signal table only: the reference's BG.angular is zero in the bins the
                    // cut removes by construction (response.jl:119-123), and the fold relies on that (it multiplies
                    // BG/2 and the unmasked inclusive CC by the same table) -- anything else is refused here.
                    for (int cs = 0; cs < D.n_cos; ++cs)
                        for (int j = 0; j < D.n_es; ++j)
                            REQUIRE(x->angular_mask[cs] != 0.0 || rb[(size_t)cs * D.n_es + j] == 0.0,
                                    "resp_bg_angular must be zero in the cos-scatter bins angular_mask removes (response.jl:119-123)");
                    for (int cs = 0; cs < D.n_cos; ++cs)
                        for (int j = 0; j < D.n_es; ++j) {
                            am[(size_t)j * D.n_cos + cs] = x->angular_mask[cs] * ra[(size_t)cs * D.n_es + j];
                            bm[(size_t)j * D.n_cos + cs] = rb[(size_t)cs * D.n_es + j];
                        }
                    D.angm = h->upload_vec(am); D.bgang = h->upload_vec(bm);
                    std::vector<double> sa(D.n_es, 0.0), sb(D.n_es, 0.0);
                    for (int j = 0; j < D.n_es; ++j)
                        for (int cs = 0; cs < D.n_cos; ++cs) { sa[j] += am[(size_t)j * D.n_cos + cs]; sb[j] += bm[(size_t)j * D.n_cos + cs]; }
                    D.ang_sa = h->upload_vec(sa); D.ang_sb = h->upload_vec(sb);
                    REQUIRE(D.n_cos <= kMaxCos, "too many cos-scatter bins");
                    D.n_cos_live = 0;
                    for (int cs = 0; cs < D.n_cos; ++cs) {
                        bool live = false;
                        for (int j = 0; j < D.n_es; ++j) live |= am[(size_t)j * D.n_cos + cs] != 0.0 || bm[(size_t)j * D.n_cos + cs] != 0.0;
                        if (live) D.cos_live[D.n_cos_live++] = (unsigned char)cs;
                    }
                }
            }
            h->n_cos[d] = D.ang ? D.n_cos : 1;
            D.a8 = h->upload_vec(a8); D.ah = h->upload_vec(ah); D.c8 = h->upload_vec(c8); D.ch = h->upload_vec(chh);
            D.b8 = h->upload_vec(b8); D.bh = h->upload_vec(bh);
            // copies that start at the threshold bin (likelihood / asymmetry modes)
            auto shifted = [&](const std::vector<double>& m, int ld, int nR, int idx0, int& blocks) {
                blocks = (nR - idx0 + 7) / 8;
                std::vector<double> out((size_t)kp * 8 * blocks, 0.0);
                for (int i = 0; i < nT; ++i)
                    for (int j = idx0; j < nR; ++j) out[(size_t)i * 8 * blocks + (j - idx0)] = m[(size_t)i * ld + j];
                return out;
            };
            D.a8_thr = h->upload_vec(shifted(a8, le, D.n_es, D.index_es - 1, D.nte_thr));
            D.ah_thr = h->upload_vec(shifted(ah, le, D.n_es, D.index_es - 1, D.nte_thr));
            D.b8_thr = h->upload_vec(shifted(b8, lc, D.n_cc, D.index_cc - 1, D.ntc_thr));
            D.bh_thr = h->upload_vec(shifted(bh, lc, D.n_cc, D.index_cc - 1, D.ntc_thr));
            // Work list of the likelihood fold (fold_stage): per column block, groups of kMCa row tiles, then the
            // remaining tiles in groups of kMCb; longest item first, each to the warp with the least work so far.
            {
                const int mt = (1 + T.n_cosz + 7) / 8, nb = D.nte_thr + D.ntc_thr;
                const int na = mt / kMCa, nbm = (mt - kMCa * na + kMCb - 1) / kMCb;
                const int ntask = nb * (na + nbm);
                D.fold_ntask = 0;
                if (ntask <= kMaxFoldTasks) {
                    int load[kWarps] = {0};
                    for (int t = 0; t < ntask; ++t) {
                        const bool big = t < nb * na;
                        const int col = (big ? t : t - nb * na) % nb;
                        const bool on = col < D.nte_thr ? D.es_mode != 0 : D.cc_mode != 0;
                        int w = 0;
                        for (int x = 1; x < kWarps; ++x)
                            if (load[x] < load[w]) w = x;
                        D.fold_owner[t] = (unsigned char)w;
                        if (on) load[w] += big ? kMCa : kMCb;
                    }
                    D.fold_ntask = ntask;
                }
            }
        }
        D.raw_bg_es = h->upload(s.raw_bg_es, (size_t)D.n_bg_es * D.n_es);
        D.raw_bg_cc = h->upload(s.raw_bg_cc, (size_t)D.n_bg_cc * D.n_cc);
        D.n_norm_es = 0; D.n_norm_cc = 0;
        for (int i = 0; i < D.n_bg_es; ++i) {
            if (s.bg_es_par[i] != 0) {  // face areas of propagation_bg.jl:6-18
                const int side = s.bg_es_sides[i];
                D.bg_es_area[i] = side == 0 ? 2 * 12e2 * 64e2 : (side == 1 ? 2 * 12e2 * 12e2 : 2 * 12e2 * 64e2 + 2 * 12e2 * 12e2);
                ++D.n_norm_es;
            } else
                D.bg_es_area[i] = 0.0;
        }
        for (int i = 0; i < D.n_bg_cc; ++i) D.bg_cc_norm[i] = s.bg_cc_par[i] != 0 ? D.n_norm_cc++ : -1;
        D.param_off = off;
        off += D.n_norm_es + D.n_norm_cc;
        const size_t ne = (size_t)(1 + T.n_cosz) * D.n_es, nc = (size_t)(1 + T.n_cosz) * D.n_cc;
        h->d_obs_es[d] = h->dalloc<double>(ne); h->d_log_es[d] = h->dalloc<double>(ne);
        h->d_obs_cc[d] = h->dalloc<double>(nc); h->d_log_cc[d] = h->dalloc<double>(nc);
        CK(cudaMemset(h->d_obs_es[d], 0, ne * sizeof(double))); CK(cudaMemset(h->d_log_es[d], 0, ne * sizeof(double)));
        CK(cudaMemset(h->d_obs_cc[d], 0, nc * sizeof(double))); CK(cudaMemset(h->d_log_cc[d], 0, nc * sizeof(double)));
        h->d_ol_es[d] = h->dalloc<double2>(ne); h->d_ol_cc[d] = h->dalloc<double2>(nc);
        CK(cudaMemset(h->d_ol_es[d], 0, ne * sizeof(double2))); CK(cudaMemset(h->d_ol_cc[d], 0, nc * sizeof(double2)));
        D.obs_es = h->d_obs_es[d]; D.logobs_es = h->d_log_es[d]; D.obs_cc = h->d_obs_cc[d]; D.logobs_cc = h->d_log_cc[d];
        D.ol_es = h->d_ol_es[d]; D.ol_cc = h->d_ol_cc[d];
        if (D.ang) {
            h->d_ol_ang[d] = h->dalloc<double2>(ne * D.n_cos);
            CK(cudaMemset(h->d_ol_ang[d], 0, ne * D.n_cos * sizeof(double2)));
            D.ol_es_ang = h->d_ol_ang[d];
        }
        if (s.obs_es_day && s.obs_es_night && s.obs_cc_day && s.obs_cc_night)
            upload_data(h, d, s.obs_es_day, s.obs_es_night, s.obs_cc_day, s.obs_cc_night);
    }
    REQUIRE(!(T.night_only && h->gen), "night_only is offered for the exclusive, non-angular analysis");
    h->P = off;
    T.n_params = off;
    const size_t osc_pt = (size_t)2 * (1 + T.n_cosz) * T.n_etrue;
    // Points per internal chunk.  The oscillation-probability workspace (osc_pt doubles per point)
    // is the only large buffer; the default keeps it near 1 GiB of the 180 GB of HBM.
    h->chunk = c->max_batch > 0 ? c->max_batch : (int)std::max<size_t>(16, std::min<size_t>(32768, ((size_t)1 << 30) / (osc_pt * sizeof(double))));
    for (int i = 0; i < 2; ++i) {
        h->d_params2[i] = h->dalloc<double>((size_t)h->chunk * h->P);
        h->d_lnl2[i] = h->dalloc<double>(h->chunk);
        h->d_flags2[i] = h->dalloc<uint32_t>(h->chunk);
        CK(cudaMallocHost(reinterpret_cast<void**>(&h->h_lnl2[i]), (size_t)h->chunk * sizeof(double)));
        CK(cudaMallocHost(reinterpret_cast<void**>(&h->h_flags2[i]), (size_t)h->chunk * sizeof(uint32_t)));
    }
    h->d_params = h->d_params2[0];
    h->osc_doubles = (size_t)h->chunk * osc_pt;   // allocated on first use (osc()): the fused kernels never need it
    const size_t rs = rates_smem_bytes(T.n_etrue, T.n_cosz, h->max_ereco);
    // A binning whose probability tile does not fit one SM's shared memory (e.g. the 1000 x 1000 zenith x
    // energy convergence sweep) can still use the oscillation-stage entry points; the fold needs the tile.
    h->rates_ok = rs <= prop.sharedMemPerBlockOptin;
    if (h->rates_ok) {
        CK(cudaFuncSetAttribute(rates_kernel<kModeLlh>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs));
        CK(cudaFuncSetAttribute(rates_kernel<kModeRates>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs));
        CK(cudaFuncSetAttribute(rates_kernel<kModePerbin>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs));
        CK(cudaFuncSetAttribute(rates_kernel<kModeAsym>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs));
        CK(cudaFuncSetAttribute(gen_kernel<kModeLlh>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs));
        CK(cudaFuncSetAttribute(gen_kernel<kModeRates>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs));
        CK(cudaFuncSetAttribute(gen_kernel<kModePerbin>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs));
        CK(cudaFuncSetAttribute(gen_kernel<kModeAsym>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs));
    }
    if (!h->rates_ok) {
        // row-chunked fold: the largest multiple of 8 rows (at most 64) whose tile fits
        int rows = 64;
        while (rows > 8 && tiled_smem_bytes(T.n_etrue, rows, h->max_ereco) > prop.sharedMemPerBlockOptin) rows -= 8;
        if (tiled_smem_bytes(T.n_etrue, rows, h->max_ereco) <= prop.sharedMemPerBlockOptin) {
            h->tiled_rows = rows;
            h->tiled_chunks = (1 + T.n_cosz + rows - 1) / rows;
            h->tiled_smem = tiled_smem_bytes(T.n_etrue, rows, h->max_ereco);
            h->d_tiled_part = h->dalloc<double>((size_t)h->chunk * h->tiled_chunks);
            h->d_tiled_flags = h->dalloc<uint32_t>((size_t)h->chunk * h->tiled_chunks);
            CK(cudaFuncSetAttribute(rates_tiled_kernel<kModeLlh>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->tiled_smem));
            CK(cudaFuncSetAttribute(rates_tiled_kernel<kModeRates>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->tiled_smem));
            CK(cudaFuncSetAttribute(rates_tiled_kernel<kModePerbin>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->tiled_smem));
        }
    }
    CK(cudaFuncSetAttribute(osc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->osc_smem_limit));
    h->grad_smem = grad_smem_bytes(T.n_etrue, T.n_cosz, h->max_ereco);
    h->adj_smem = sizeof(double) * (adj_smem_doubles(T.n_ltbl, T.n_e, seg_smem_doubles(T.n_seg, T.n_paths, T.n_cosz)) + 2);
    h->gengrad_smem = gengrad_smem_bytes(T.n_etrue, T.n_cosz, h->max_ereco);
    h->grad_ok = (h->gen ? h->gengrad_smem : h->grad_smem) <= prop.sharedMemPerBlockOptin && h->adj_smem <= prop.sharedMemPerBlockOptin;
    if (h->grad_ok && h->gen) CK(cudaFuncSetAttribute(gen_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->gengrad_smem));
    if (h->grad_ok) {
        CK(cudaFuncSetAttribute(grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->grad_smem));
        CK(cudaFuncSetAttribute(adjoint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->adj_smem));
    }
    // Fused kernel: the energy axis is walked in chunks of fused_cb coarse bins, the smallest eigen table that
    // still gives the propagation loop whole passes of work (n_cosz * cb items over kThreads threads), so that
    // SOLAROSC_LLH_MINBLOCKS CTAs fit the SM's shared memory.  SOLAROSC_LLH_CHUNKS overrides the chunk count.
    h->seg_bytes = sizeof(double) * seg_smem_doubles(T.n_seg, T.n_paths, T.n_cosz);
    {
        const size_t per_sm = prop.sharedMemPerMultiprocessor, per_cta_max = prop.sharedMemPerBlockOptin;
        const size_t reserve = 1024;   // the driver keeps 1 kB per resident CTA
        const size_t budget = std::min(per_cta_max, per_sm / SOLAROSC_LLH_MINBLOCKS - reserve);
        int chunks = 1;
        if (const char* e = std::getenv("SOLAROSC_LLH_CHUNKS")) chunks = std::max(1, atoi(e));
        else
            while (chunks < T.n_etrue &&
                   llh_smem_bytes(T.n_ltbl, (T.n_etrue + chunks - 1) / chunks, T.n_etrue, T.n_cosz, h->max_ereco, h->seg_bytes / sizeof(double)) > budget)
                ++chunks;
        h->fused_cb = (T.n_etrue + chunks - 1) / chunks;
        h->fused_smem = llh_smem_bytes(T.n_ltbl, h->fused_cb, T.n_etrue, T.n_cosz, h->max_ereco, h->seg_bytes / sizeof(double));
        h->fused_seg = true;
        if (const char* e = std::getenv("SOLAROSC_LLH_EXTRA_SMEM")) h->fused_smem += (size_t)atoi(e);   // diagnostics: force fewer CTAs per SM
        h->fused_ok = h->fused_seg && h->fused_smem <= per_cta_max;   // the fused kernel reads the path tables from shared memory
    }
    CK(cudaFuncSetAttribute(pipe_kernel<kModeLlh>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
    CK(cudaFuncSetAttribute(pipe_kernel<kModeAsym>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
    CK(cudaFuncSetAttribute(pipe_kernel<kModeRates>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
    CK(cudaFuncSetAttribute(pipe_kernel<kModePerbin>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
    if (h->fused_ok) {
        CK(cudaFuncSetAttribute(llh_kernel<kModeLlh, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->fused_smem));
        CK(cudaFuncSetAttribute(llh_kernel<kModeLlh, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->fused_smem));
        CK(cudaFuncSetAttribute(llh_kernel<kModeAsym, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->fused_smem));
        // grouped evaluation cuts the fold into finer work items when every detector has a work list (fold_stage)
        h->fold_fine_ok = true;
        for (int d = 0; d < T.n_det; ++d) {
            const DevDetector& D = T.det[d];
            const int mt = (1 + T.n_cosz + 7) / 8, nb = D.nte_thr + D.ntc_thr;
            const int na = mt / kMCa, nbm = (mt - kMCa * na + kMCb - 1) / kMCb;
            h->fold_fine_ok = h->fold_fine_ok && D.fold_ntask > 0 && D.fold_ntask == nb * (na + nbm);
        }
        CK(cudaFuncSetAttribute(gen_llh_kernel<kModeLlh>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->fused_smem));
        CK(cudaFuncSetAttribute(gen_llh_kernel<kModeAsym>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->fused_smem));
        // One CTA per oscillation point by default: measured 8 % faster than a persistent grid of 2 CTAs per SM
        // (starting the second CTA of each SM half a period late made no difference, so it is not a phase effect).
        h->fused_grid = 1 << 30;
        if (const char* e = std::getenv("SOLAROSC_LLH_GRID")) h->fused_grid = std::max(1, atoi(e));
    }
}

// ---- launches ------------------------------------------------------------------------------------
void launch_osc(solarosc_handle* h, const double* d_params, int n_osc, int point_step, double* d_p1e, cudaStream_t st,
                double* d_out = nullptr) {
    const DevTables& T = h->T;
    // coarse bins per CTA: the whole spectrum when the batch alone fills the GPU, otherwise split the
    // energy axis so that about two CTAs per SM exist; never more than shared memory holds
    const size_t per_cb = osc_smem_bytes(T.n_ltbl, 1);
    const int cb_max = (int)std::max<size_t>(1, h->osc_smem_limit / per_cb);
    // (binnings beyond 256 true-energy bins -- the dense convergence grids, a few dozen points per call -- get eight CTAs per SM
    // instead of two: their CTAs are long, and the tail of the last wave costs more than the extra table set-ups)
    static const int ctas_env = std::getenv("SOLAROSC_OSC_CTAS_PER_SM") ? atoi(std::getenv("SOLAROSC_OSC_CTAS_PER_SM")) : 0;   // tuning switch
    const int per_sm = ctas_env > 0 ? ctas_env : (T.n_etrue > 256 ? 8 : 2);
    int want_chunks = (per_sm * h->sm_count + n_osc - 1) / n_osc;
    want_chunks = std::max(1, std::min(want_chunks, std::max(1, T.n_etrue / 4)));
    int cb = (T.n_etrue + want_chunks - 1) / want_chunks;
    cb = std::max(1, std::min(cb, cb_max));
    const int n_chunks = (T.n_etrue + cb - 1) / cb;
    size_t sm = osc_smem_bytes(T.n_ltbl, cb);
    const bool seg = sm + h->seg_bytes <= std::min<size_t>(h->osc_smem_limit, 110 * 1024);
    if (seg) sm += h->seg_bytes;
    OscLaunch L{n_osc, n_chunks, cb, h->P, point_step, seg ? 1 : 0};
    osc_kernel<<<n_osc * n_chunks, kThreads, sm, st>>>(T, L, d_params, d_out ? d_out : h->osc(), d_p1e);
    CK(cudaGetLastError());
    ++h->launches;
}

RatesLaunch rates_launch(solarosc_handle* h, int n, int group, int det_only) {
    const DevTables& T = h->T;
    RatesLaunch L{};
    L.n_points = n; L.group = group; L.param_stride = h->P; L.det_only = det_only; L.max_ereco = h->max_ereco;
    L.seg_smem = h->fused_seg ? 1 : 0;
    L.osc_cb = h->fused_cb;
    const int d = det_only >= 0 ? det_only : 0;
    const bool cm = h->cfg.col_major != 0;
    L.s_row_es = cm ? 1 : h->n_es[d]; L.s_j_es = cm ? T.n_cosz : 1;
    L.s_row_cc = cm ? 1 : h->n_cc[d]; L.s_j_cc = cm ? T.n_cosz : 1;
    return L;
}

void launch_rates(solarosc_handle* h, const double* d_params, int n, int group, int mode, int det_only, RatesOut O,
                  cudaStream_t st) {
    const DevTables& T = h->T;
    const RatesLaunch L = rates_launch(h, n, group, det_only);
    const size_t sm = rates_smem_bytes(T.n_etrue, T.n_cosz, h->max_ereco);
    if (!h->rates_ok) {   // dense binnings: row-chunked fold
        if (h->gen || h->tiled_rows == 0 || mode == kModeAsym)
            throw ArgFail{"binning too large for the fold's shared-memory tile: this call is not available (likelihood, rates and per-bin terms of the exclusive analysis are)"};
        const TiledLaunch TL{h->tiled_rows, h->tiled_chunks, h->d_tiled_part, h->d_tiled_flags};
        const int grid = n * h->tiled_chunks;
        if (mode == kModeLlh) {
            rates_tiled_kernel<kModeLlh><<<grid, kThreads, h->tiled_smem, st>>>(T, L, TL, d_params, h->osc(), O);
            tiled_finish_kernel<<<(n + 127) / 128, 128, 0, st>>>(TL, n, O.lnl, O.flags);
            ++h->launches;
        } else if (mode == kModeRates) rates_tiled_kernel<kModeRates><<<grid, kThreads, h->tiled_smem, st>>>(T, L, TL, d_params, h->osc(), O);
        else rates_tiled_kernel<kModePerbin><<<grid, kThreads, h->tiled_smem, st>>>(T, L, TL, d_params, h->osc(), O);
        CK(cudaGetLastError());
        ++h->launches;
        return;
    }
    if (h->gen) {

        const int d = det_only >= 0 ? det_only : 0;
        const bool cm = h->cfg.col_major != 0;
        const long long nC = h->n_cos[d], nE = h->n_es[d], nZ = T.n_cosz;
        const GenLaunch GL = cm ? GenLaunch{1, nC, 1, nC, nC * nE} : GenLaunch{nE, 1, nE * nZ, nZ, 1};
        if (mode == kModeLlh) gen_kernel<kModeLlh><<<n, kThreads, sm, st>>>(T, L, GL, d_params, h->osc(), O);
        else if (mode == kModeRates) gen_kernel<kModeRates><<<n, kThreads, sm, st>>>(T, L, GL, d_params, h->osc(), O);
        else if (mode == kModePerbin) gen_kernel<kModePerbin><<<n, kThreads, sm, st>>>(T, L, GL, d_params, h->osc(), O);
        else gen_kernel<kModeAsym><<<n, kThreads, sm, st>>>(T, L, GL, d_params, h->osc(), O);
        CK(cudaGetLastError());
        ++h->launches;
        return;
    }
    if (mode == kModeLlh) rates_kernel<kModeLlh><<<n, kThreads, sm, st>>>(T, L, d_params, h->osc(), O);
    else if (mode == kModeAsym) rates_kernel<kModeAsym><<<n, kThreads, sm, st>>>(T, L, d_params, h->osc(), O);
    else if (mode == kModeRates) rates_kernel<kModeRates><<<n, kThreads, sm, st>>>(T, L, d_params, h->osc(), O);
    else rates_kernel<kModePerbin><<<n, kThreads, sm, st>>>(T, L, d_params, h->osc(), O);
    CK(cudaGetLastError());
    ++h->launches;
}

// Ticket pattern and fold items of pipe_kernel for one (mode, group, detector) combination (solarosc_pipe.cuh).
// Within the block of tickets of one step the propagation items P(n) are spread evenly and the eigen/MSW items
// E(n+1) and fold items F(n-1) are spread over the middle of the block: their dependencies (items of the previous
// block) have drained by then, and they are done before the next block's propagation asks for their results.
const solarosc_handle::PipePlan* pipe_plan(solarosc_handle* h, int mode, int group, int det_only) {
    for (const auto& pl : h->pipe_plans)
        if (pl.mode == mode && pl.group == group && pl.det_only == det_only) return pl.ok ? &pl : nullptr;
    const DevTables& T = h->T;
    solarosc_handle::PipePlan pl{};
    pl.mode = mode; pl.group = group; pl.det_only = det_only; pl.ok = false;
    const bool thr = mode == kModeLlh || mode == kModeAsym;
    const int nrow = 1 + T.n_cosz, mt = (nrow + 7) / 8;
    std::vector<PipeFoldItem> items;
    for (int d = 0; d < T.n_det; ++d) {
        if (det_only >= 0 && d != det_only) continue;
        const DevDetector& D = T.det[d];
        const int nb = thr ? D.nte_thr + D.ntc_thr : (D.n_es + 7) / 8 + (D.n_cc + 7) / 8;
        int per = kPipeNT;                             // row tiles per fold item
        if (const char* e = std::getenv("SOLAROSC_PIPE_FTILES")) per = std::max(1, atoi(e));
        for (int col = 0; col < nb; ++col)
            for (int m0 = 0; m0 < mt; m0 += per)
                items.push_back(PipeFoldItem{(unsigned char)d, (unsigned char)col, (unsigned char)m0, (unsigned char)std::min(per, mt - m0)});
    }
    PipeLaunch L{};
    L.group = group; L.param_stride = h->P; L.det_only = det_only; L.max_ereco = h->max_ereco;
    L.n_f = (int)items.size();
    L.n_p = 1 + (T.n_cosz * T.n_etrue + 31) / 32;
    L.n_eig = (2 * T.n_etrue * T.n_ltbl + 32 * kPipeEPL - 1) / (32 * kPipeEPL);
    L.msw_per_item = T.fast ? 0 : 2;
    L.n_e = L.n_eig + (T.fast ? 1 : (2 * T.n_etrue + L.msw_per_item - 1) / L.msw_per_item);
    const int red = mode == kModeAsym ? 3 : 1;
    const PipeSmem S = pipe_smem(T.n_ltbl, T.n_etrue, T.n_cosz, group, L.n_f, red, h->seg_bytes / sizeof(double));
    pl.smem = S.total * sizeof(double);
    const long long n_f_tot = (long long)group * L.n_f;
    const bool fits = !h->gen && pl.smem <= h->smem_optin && group <= kPipeMaxGroup && n_f_tot < (1 << 14) && L.n_p < (1 << 14) &&
                      mt <= 255 && L.n_f > 0 && T.n_sub >= 1;
    if (fits) {
        // Items of one kind are issued in clusters (all E items together, the F items in a few runs), so that the warps
        // that draw them stream the same straight-line code at the same time and share its instruction-cache lines;
        // scattered singly, every item fetched its 30 kB of code on its own (ncu: 33 % of the stalls were
        // no_instructions).  The clusters sit evenly between runs of propagation items.
        struct Tk { double key; unsigned short code; };
        std::vector<Tk> tk;
        int f_clusters = 1, e_clusters = 1;                                 // 0: scattered singly
        if (const char* e = std::getenv("SOLAROSC_PIPE_FCLUSTERS")) f_clusters = std::max(0, atoi(e));
        if (const char* e = std::getenv("SOLAROSC_PIPE_ECLUSTERS")) e_clusters = std::max(0, atoi(e));
        const long long fc = f_clusters == 0 ? n_f_tot : std::min<long long>(f_clusters * (long long)group, n_f_tot);
        const int ec = e_clusters == 0 ? L.n_e : std::min(e_clusters, L.n_e);
        for (int i = 0; i < L.n_p; ++i) tk.push_back({(i + 0.5) / L.n_p, (unsigned short)((kPipeP << 14) | i)});
        // E clusters over the first 45 % of the block (the next step's propagation waits for them longest), F clusters
        // over 50 .. 90 %
        for (int i = 0; i < L.n_e; ++i) {
            const int c = (int)((long long)i * ec / L.n_e);
            tk.push_back({0.05 + 0.40 * (c + 0.5) / ec + 1e-7 * i, (unsigned short)((kPipeE << 14) | i)});
        }
        for (long long i = 0; i < n_f_tot; ++i) {
            const long long c = i * fc / n_f_tot;
            tk.push_back({0.50 + 0.40 * (c + 0.5) / (double)fc + 1e-7 * (double)i, (unsigned short)((kPipeF << 14) | (int)i)});
        }
        std::stable_sort(tk.begin(), tk.end(), [](const Tk& a, const Tk& b) { return a.key < b.key; });
        std::vector<unsigned short> pattern(tk.size());
        for (size_t i = 0; i < tk.size(); ++i) pattern[i] = tk[i].code;
        L.block_len = (int)pattern.size();
        L.pattern = h->upload(pattern.data(), pattern.size());
        L.fitems = h->upload(items.data(), items.size());
        pl.ok = true;
    }
    pl.L = L;
    h->pipe_plans.push_back(pl);
    return pl.ok ? &h->pipe_plans.back() : nullptr;
}

bool launch_pipe(solarosc_handle* h, const double* d_params, int n, int group, int mode, int det_only, RatesOut O, cudaStream_t st,
                 bool row_major = false) {
    if (!h->use_pipe || !h->use_fused || h->gen) return false;
    const int n_osc = n / group;
    if (n_osc < h->sm_count) return false;
    // Independent points (group == 1) stay with the lock-step llh_kernel: measured 7.8 M evaluations/s against 6.7 M for
    // this kernel, whose ticket / dependency bookkeeping costs 23 % more instructions while the SM is co-limited by issue
    // slots (profiles/r02_notes.md).  The pipeline pays where the fold dominates (grouped scan cells: 24.3 M against
    // 21.8 M) and for the array outputs, which otherwise take the probability block through HBM.
    if (mode == kModeLlh && group == 1 && !h->pipe_all) return false;
    const solarosc_handle::PipePlan* pl = pipe_plan(h, mode, group, det_only);
    if (!pl) return false;
    PipeLaunch L = pl->L;
    L.n_osc = n_osc;
    const int d = det_only >= 0 ? det_only : 0;
    const bool cm = h->cfg.col_major != 0 && !row_major;
    L.s_row_es = cm ? 1 : h->n_es[d]; L.s_j_es = cm ? h->T.n_cosz : 1;
    L.s_row_cc = cm ? 1 : h->n_cc[d]; L.s_j_cc = cm ? h->T.n_cosz : 1;
    const int grid = std::min(n_osc, h->sm_count);
    if (mode == kModeLlh) pipe_kernel<kModeLlh><<<grid, kPipeThreads, pl->smem, st>>>(h->T, L, d_params, O);
    else if (mode == kModeAsym) pipe_kernel<kModeAsym><<<grid, kPipeThreads, pl->smem, st>>>(h->T, L, d_params, O);
    else if (mode == kModeRates) pipe_kernel<kModeRates><<<grid, kPipeThreads, pl->smem, st>>>(h->T, L, d_params, O);
    else pipe_kernel<kModePerbin><<<grid, kPipeThreads, pl->smem, st>>>(h->T, L, d_params, O);
    CK(cudaGetLastError());
    ++h->launches;
    return true;
}

// lnL of n points; the fused kernel when the batch fills the GPU, otherwise the energy-split
// oscillation kernel followed by the fold kernel.  Returns true when the fused kernel ran.
bool launch_llh(solarosc_handle* h, const double* d_params, int n, int group, RatesOut O, cudaStream_t st, cudaEvent_t mid,
                int mode = kModeLlh, int det_only = -1) {
    const int n_osc = n / group;
    if (launch_pipe(h, d_params, n, group, mode, det_only, O, st)) {
        if (mid) CK(cudaEventRecord(mid, st));
        return true;
    }
    if (h->fused_ok && h->use_fused && h->gen && n_osc >= h->sm_count) {
        // inclusive / semi-inclusive / angular detectors: the same fusion with the general fold
        const RatesLaunch L = rates_launch(h, n, group, det_only);
        const int grid = std::min(n_osc, h->fused_grid);
        if (mode == kModeLlh) gen_llh_kernel<kModeLlh><<<grid, kThreads, h->fused_smem, st>>>(h->T, L, GenLaunch{}, d_params, O);
        else gen_llh_kernel<kModeAsym><<<grid, kThreads, h->fused_smem, st>>>(h->T, L, GenLaunch{}, d_params, O);
        CK(cudaGetLastError());
        ++h->launches;
        if (mid) CK(cudaEventRecord(mid, st));
        return true;
    }
    if (h->fused_ok && h->use_fused && !h->gen && n_osc >= h->sm_count) {
        const RatesLaunch L = rates_launch(h, n, group, det_only);
        const int grid = std::min(n_osc, h->fused_grid);
        if (mode == kModeLlh && (group > 1 || h->fine_all) && h->fold_fine_ok) llh_kernel<kModeLlh, true><<<grid, kThreads, h->fused_smem, st>>>(h->T, L, d_params, O);
        else if (mode == kModeLlh) llh_kernel<kModeLlh, false><<<grid, kThreads, h->fused_smem, st>>>(h->T, L, d_params, O);
        else llh_kernel<kModeAsym, false><<<grid, kThreads, h->fused_smem, st>>>(h->T, L, d_params, O);
        CK(cudaGetLastError());
        ++h->launches;
        if (mid) CK(cudaEventRecord(mid, st));
        return true;
    }
    launch_osc(h, d_params, n_osc, group, nullptr, st);
    if (mid) CK(cudaEventRecord(mid, st));
    launch_rates(h, d_params, n, group, mode, det_only, O, st);
    return false;
}

void check_data(solarosc_handle* h, int det_only) {
    if (!h->rates_ok && (h->tiled_rows == 0 || h->gen))
        throw ArgFail{"binning too large for the fold's shared-memory tile: only the oscillation-stage calls are available"};
    for (int d = 0; d < h->T.n_det; ++d)
        if ((det_only < 0 || d == det_only) && !h->has_data[d])
            throw ArgFail{"observed data not set for a detector (pass obs_* at create or call solarosc_set_data)"};
}

template <class F>
int guarded(solarosc_handle* h, F&& body) {
    if (!h) return SOLAROSC_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    try {
        CK(cudaSetDevice(h->device));
        return body();
    } catch (const ArgFail& e) {
        h->err = e.msg;
        return SOLAROSC_ERR_ARG;
    } catch (const CudaFail& e) {
        h->err = e.msg;
        return SOLAROSC_ERR_CUDA;
    } catch (const std::bad_alloc&) {
        h->err = "host allocation failed";
        return SOLAROSC_ERR_ALLOC;
    } catch (...) {
        h->err = "unexpected exception";
        return SOLAROSC_ERR_ARG;
    }
}

// Small batches are launch- and latency-bound: the whole step (H2D of the rows, oscillation and fold kernels, D2H of lnL
// and flags) is captured once into a CUDA graph and replayed with one launch per call.
constexpr size_t kSmallBatch = 256;
int eval_llh_small(solarosc_handle* h, const double* params, size_t n, size_t group, double* out, uint32_t* flags) {
    if (!h->h_params_pin) CK(cudaMallocHost(reinterpret_cast<void**>(&h->h_params_pin), kSmallBatch * (size_t)h->P * sizeof(double)));
    if (!h->small_graph || h->small_n != n || h->small_group != group) {
        if (h->small_graph) { cudaGraphExecDestroy(h->small_graph); h->small_graph = nullptr; }
        cudaGraph_t g = nullptr;
        // nothing may allocate inside the capture: the pipe plan uploads its tables, the stage-separated path its workspace
        if (n / group >= (size_t)h->sm_count) pipe_plan(h, kModeLlh, (int)group, -1);
        if (n / group < (size_t)h->sm_count || !h->fused_ok || !h->use_fused) h->osc();
        CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        try {
            CK(cudaMemcpyAsync(h->d_params2[0], h->h_params_pin, n * h->P * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            RatesOut O{};
            O.lnl = h->d_lnl2[0]; O.flags = h->d_flags2[0];
            h->launches = 0;
            h->small_fused = launch_llh(h, h->d_params2[0], (int)n, (int)group, O, h->stream, nullptr);
            h->small_launches = h->launches;
            CK(cudaMemcpyAsync(h->h_lnl2[0], h->d_lnl2[0], n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaMemcpyAsync(h->h_flags2[0], h->d_flags2[0], n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        } catch (...) {
            cudaStreamEndCapture(h->stream, &g);
            if (g) cudaGraphDestroy(g);
            throw;
        }
        CK(cudaStreamEndCapture(h->stream, &g));
        const cudaError_t e = cudaGraphInstantiate(&h->small_graph, g, 0);
        cudaGraphDestroy(g);
        CK(e);
        h->small_n = n; h->small_group = group;
    }
    std::memcpy(h->h_params_pin, params, n * h->P * sizeof(double));
    CK(cudaGraphLaunch(h->small_graph, h->stream));
    h->launches = h->small_launches;
    h->last_fused = h->small_fused;
    CK(cudaStreamSynchronize(h->stream));
    std::memcpy(out, h->h_lnl2[0], n * sizeof(double));
    bool neg = false;
    for (size_t i = 0; i < n; ++i) neg |= (h->h_flags2[0][i] & SOLAROSC_FLAG_NEG_ERROR) != 0;
    if (flags) std::memcpy(flags, h->h_flags2[0], n * sizeof(uint32_t));
    if (neg) {
        h->err = "an expected rate was <= -1e-3 for at least one point (lnL = NaN there)";
        return SOLAROSC_ERR_NEGATIVE_RATE;
    }
    return SOLAROSC_OK;
}

// Host-buffer likelihood evaluation.  Three streams form a pipeline over chunks: H2D of chunk c+1
// (straight from the caller's buffer: asynchronous when it is pinned, staged by the driver when
// pageable) overlaps the two kernels of chunk c and the D2H of chunk c-1 into pinned staging; the host thread
// enqueues two chunks ahead and meanwhile hands the finished chunks to the caller's arrays.
int eval_llh_host(solarosc_handle* h, const double* params, size_t n, size_t group, double* out, uint32_t* flags) {
    REQUIRE(params && out, "params and out_lnl must not be NULL");
    REQUIRE(group >= 1 && n % group == 0, "n_points must be a multiple of group");
    REQUIRE(group <= (size_t)h->chunk, "group larger than max_batch");
    check_data(h, -1);
    h->launches = 0;
    h->tev_used = 0;
    h->tev_stream = h->stream;
    if (n <= kSmallBatch && n <= (size_t)h->chunk && !h->timing) return eval_llh_small(h, params, n, group, out, flags);
    // results come back through two pinned staging slots and are handed to the caller's (possibly pageable) arrays by
    // the host thread two chunks behind the GPU: every copy on the streams stays asynchronous
    const size_t step = ((size_t)h->chunk / group) * group;
    bool neg = false;
    auto drain = [&](size_t c) {
        const size_t b0 = c * step;
        const size_t nb = std::min(step, n - b0);
        const int slot = (int)(c & 1);
        std::memcpy(out + b0, h->h_lnl2[slot], nb * sizeof(double));
        const uint32_t* f = h->h_flags2[slot];
        for (size_t i = 0; i < nb; ++i) neg |= (f[i] & SOLAROSC_FLAG_NEG_ERROR) != 0;
        if (flags) std::memcpy(flags + b0, f, nb * sizeof(uint32_t));
    };
    size_t c = 0;
    for (size_t b0 = 0; b0 < n; b0 += step, ++c) {
        const int nb = (int)std::min(step, n - b0);
        const int slot = (int)(c & 1);
        if (c >= 2) {
            CK(cudaEventSynchronize(h->ev_out[slot]));                           // chunk c-2 has landed in this staging slot
            drain(c - 2);
            CK(cudaStreamWaitEvent(h->s_in, h->ev_done[slot], 0));               // kernels of chunk c-2 have read this parameter slot
        }
        CK(cudaMemcpyAsync(h->d_params2[slot], params + b0 * h->P, (size_t)nb * h->P * sizeof(double), cudaMemcpyHostToDevice, h->s_in));
        CK(cudaEventRecord(h->ev_in[slot], h->s_in));
        CK(cudaStreamWaitEvent(h->stream, h->ev_in[slot], 0));
        if (c >= 2) CK(cudaStreamWaitEvent(h->stream, h->ev_out[slot], 0));  // results of chunk c-2 have left this slot
        RatesOut O{};
        O.lnl = h->d_lnl2[slot]; O.flags = h->d_flags2[slot];
        if (h->timing) CK(cudaEventRecord(h->next_tev(), h->stream));
        cudaEvent_t mid = h->timing ? h->next_tev() : nullptr;
        h->last_fused = launch_llh(h, h->d_params2[slot], nb, (int)group, O, h->stream, mid);
        if (h->timing) CK(cudaEventRecord(h->next_tev(), h->stream));
        CK(cudaEventRecord(h->ev_done[slot], h->stream));
        CK(cudaStreamWaitEvent(h->s_out, h->ev_done[slot], 0));
        CK(cudaMemcpyAsync(h->h_lnl2[slot], h->d_lnl2[slot], nb * sizeof(double), cudaMemcpyDeviceToHost, h->s_out));
        CK(cudaMemcpyAsync(h->h_flags2[slot], h->d_flags2[slot], nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->s_out));
        CK(cudaEventRecord(h->ev_out[slot], h->s_out));
    }
    CK(cudaStreamSynchronize(h->s_out));
    CK(cudaStreamSynchronize(h->stream));
    for (size_t k = c >= 2 ? c - 2 : 0; k < c; ++k) drain(k);
    if (neg) {
        h->err = "an expected rate was <= -1e-3 for at least one point (lnL = NaN there)";
        return SOLAROSC_ERR_NEGATIVE_RATE;
    }
    return SOLAROSC_OK;
}

// Rates / per-bin evaluation of one detector.  Chunks that fill the GPU run the fused pipe_kernel (the probability block
// never goes to HBM); smaller ones osc_kernel + rates_kernel.  Two scratch slots: the kernels of chunk c + 1 are enqueued
// before the host copies chunk c back, so copies into the caller's (usually pageable) arrays overlap the next chunk.
int eval_arrays_host(solarosc_handle* h, int det, int mode, const double* params, size_t n, double* es_day, double* es_night,
                     double* cc_day, double* cc_night, double* bg_es, double* bg_cc) {
    REQUIRE(params, "params must not be NULL");
    REQUIRE(det >= 0 && det < h->T.n_det, "detector index out of range");
    if (mode == kModePerbin) check_data(h, det);
    REQUIRE(h->rates_ok || (h->tiled_rows > 0 && !h->gen), "binning too large for the fold's shared-memory tile");
    const int nZ = h->T.n_cosz, nE = h->n_es[det], nC = h->n_cc[det];
    const size_t nA = (size_t)h->n_cos[det] * nE;   // ES array entries per row (cos-scatter axis in the angular analysis)
    const size_t per = (size_t)(1 + nZ) * (nA + nC) + nE + nC;
    const int step = std::max(1, std::min(h->chunk, nA > (size_t)nE ? 128 : 1024));
    h->ensure_scratch(2 * per * step * sizeof(double));
    h->launches = 0;
    struct Slot { RatesOut O; size_t b0; int nb; };
    Slot sl[2];
    auto back = [&](const Slot& s) {
        CK(cudaEventSynchronize(h->ev_done[&s - sl]));
        auto cp = [&](double* dst, const double* src, size_t per_pt) {
            if (dst) CK(cudaMemcpyAsync(dst + s.b0 * per_pt, src, (size_t)s.nb * per_pt * sizeof(double), cudaMemcpyDeviceToHost, h->s_out));
        };
        cp(es_day, s.O.es_day, nA); cp(es_night, s.O.es_night, (size_t)nZ * nA);
        cp(cc_day, s.O.cc_day, nC); cp(cc_night, s.O.cc_night, (size_t)nZ * nC);
        if (mode == kModeRates) { cp(bg_es, s.O.bg_es, nE); cp(bg_cc, s.O.bg_cc, nC); }
        CK(cudaEventRecord(h->ev_out[&s - sl], h->s_out));
    };
    size_t c = 0;
    for (size_t b0 = 0; b0 < n; b0 += step, ++c) {
        const int nb = (int)std::min<size_t>(step, n - b0), slot = (int)(c & 1);
        if (c >= 2) CK(cudaStreamWaitEvent(h->stream, h->ev_out[slot], 0));   // chunk c - 2 has left this scratch slot
        CK(cudaMemcpyAsync(h->d_params2[slot], params + b0 * h->P, (size_t)nb * h->P * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        double* s = h->d_scratch + (size_t)slot * per * step;
        RatesOut O{};
        O.es_day = s; s += (size_t)nb * nA;
        O.es_night = s; s += (size_t)nb * nZ * nA;
        O.cc_day = s; s += (size_t)nb * nC;
        O.cc_night = s; s += (size_t)nb * nZ * nC;
        O.bg_es = s; s += (size_t)nb * nE;
        O.bg_cc = s;
        if (!launch_pipe(h, h->d_params2[slot], nb, 1, mode, det, O, h->stream)) {
            launch_osc(h, h->d_params2[slot], nb, 1, nullptr, h->stream);
            launch_rates(h, h->d_params2[slot], nb, 1, mode, det, O, h->stream);
        }
        CK(cudaEventRecord(h->ev_done[slot], h->stream));
        sl[slot] = Slot{O, b0, nb};
        if (c >= 1) back(sl[slot ^ 1]);
    }
    if (c >= 1) back(sl[(c - 1) & 1]);
    CK(cudaStreamSynchronize(h->s_out));
    CK(cudaStreamSynchronize(h->stream));
    return SOLAROSC_OK;
}

// lnL and its gradient.  Per chunk: osc_kernel writes the probabilities, grad_kernel folds, differentiates the Poisson sum and
// folds back (flux, cross-section and background columns; d lnL / d P to the workspace), adjoint_kernel carries d lnL / d P
// through the oscillation stage exactly (mixing, dm2_21 and earth-norm columns).  Two workspace slots: the kernels of chunk
// c + 1 are enqueued before the host collects chunk c.
int eval_llh_grad_host(solarosc_handle* h, const double* params, size_t n, double* out_lnl, double* out_grad, uint32_t* flags) {
    REQUIRE(params && out_lnl && out_grad, "params, out_lnl and out_grad must not be NULL");
    REQUIRE(h->grad_ok, "binning too large for the gradient kernels' shared-memory tiles");
    REQUIRE(!h->T.night_only, "the gradient entry does not offer night_only");
    check_data(h, -1);
    const DevTables& T = h->T;
    const size_t osc_pt = (size_t)2 * (1 + T.n_cosz) * T.n_etrue;
    const int step = std::max(1, std::min(h->chunk, 4096));
    const size_t slot_doubles = (size_t)step * (2 * osc_pt + h->P);
    const size_t need = 2 * slot_doubles * sizeof(double);
    if (need > h->gradws_bytes) {
        if (h->d_gradws) CK(cudaFree(h->d_gradws));
        h->d_gradws = nullptr; h->gradws_bytes = 0;
        CK(cudaMalloc(reinterpret_cast<void**>(&h->d_gradws), need));
        h->gradws_bytes = need;
    }
    if (!h->h_grad_pin[0])
        for (int i = 0; i < 2; ++i) CK(cudaMallocHost(reinterpret_cast<void**>(&h->h_grad_pin[i]), (size_t)std::min(h->chunk, 4096) * h->P * sizeof(double)));
    bool neg = false;
    h->launches = 0;
    struct Slot { size_t b0; int nb; };
    Slot sl[2];
    auto collect = [&](int slot) {
        const Slot& s = sl[slot];
        CK(cudaEventSynchronize(h->ev_out[slot]));
        std::memcpy(out_lnl + s.b0, h->h_lnl2[slot], (size_t)s.nb * sizeof(double));
        std::memcpy(out_grad + s.b0 * h->P, h->h_grad_pin[slot], (size_t)s.nb * h->P * sizeof(double));
        const uint32_t* f = h->h_flags2[slot];
        for (int i = 0; i < s.nb; ++i) neg |= (f[i] & SOLAROSC_FLAG_NEG_ERROR) != 0;
        if (flags) std::memcpy(flags + s.b0, f, (size_t)s.nb * sizeof(uint32_t));
    };
    size_t c = 0;
    for (size_t b0 = 0; b0 < n; b0 += step, ++c) {
        const int nb = (int)std::min<size_t>(step, n - b0), slot = (int)(c & 1);
        if (c >= 2) collect(slot);   // chunk c - 2 has left this slot's pinned staging (and its device buffers)
        double* d_osc = h->d_gradws + (size_t)slot * slot_doubles;
        double* d_pbar = d_osc + (size_t)step * osc_pt;
        double* d_grad = d_pbar + (size_t)step * osc_pt;
        CK(cudaMemcpyAsync(h->d_params2[slot], params + b0 * h->P, (size_t)nb * h->P * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        launch_osc(h, h->d_params2[slot], nb, 1, nullptr, h->stream, d_osc);
        GradLaunch G{nb, h->P, h->max_ereco};
        if (h->gen)   // a handle with an inclusive / semi-inclusive / angular detector: the general fold's reverse mode
            gen_grad_kernel<<<nb, kThreads, h->gengrad_smem, h->stream>>>(T, G, h->d_params2[slot], d_osc, d_pbar, h->d_lnl2[slot], d_grad, h->d_flags2[slot]);
        else
            grad_kernel<<<nb, kThreads, h->grad_smem, h->stream>>>(T, G, h->d_params2[slot], d_osc, d_pbar, h->d_lnl2[slot], d_grad, h->d_flags2[slot]);
        AdjLaunch A{nb, h->P};
        adjoint_kernel<<<nb, kThreads, h->adj_smem, h->stream>>>(T, A, h->d_params2[slot], d_pbar, d_grad, h->d_flags2[slot]);
        CK(cudaGetLastError());
        h->launches += 2;
        CK(cudaMemcpyAsync(h->h_lnl2[slot], h->d_lnl2[slot], nb * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(h->h_grad_pin[slot], d_grad, (size_t)nb * h->P * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(h->h_flags2[slot], h->d_flags2[slot], nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaEventRecord(h->ev_out[slot], h->stream));
        sl[slot] = Slot{b0, nb};
    }
    for (size_t k = c >= 2 ? c - 2 : 0; k < c; ++k) collect((int)(k & 1));
    if (neg) {
        h->err = "an expected rate was <= -1e-3 for at least one point (lnL = NaN there)";
        return SOLAROSC_ERR_NEGATIVE_RATE;
    }
    return SOLAROSC_OK;
}

// host transpose of per-point [a][b] blocks into [b][a] (Julia layout) in place via a temporary
void transpose_blocks(double* x, size_t n, int na, int nb) {
    std::vector<double> tmp((size_t)na * nb);
    for (size_t p = 0; p < n; ++p) {
        double* blk = x + p * na * nb;
        for (int a = 0; a < na; ++a)
            for (int b = 0; b < nb; ++b) tmp[(size_t)b * na + a] = blk[(size_t)a * nb + b];
        std::copy(tmp.begin(), tmp.end(), blk);
    }
}

}  // namespace

// ======================================================================================================
extern "C" {

int solarosc_abi_version(void) { return SOLAROSC_ABI_VERSION; }

solarosc_handle* solarosc_create(const solarosc_tables* tables, const solarosc_config* config, int* status) {
    solarosc_handle* h = nullptr;
    int st = SOLAROSC_OK;
    try {
        h = new solarosc_handle();
        build(h, tables, config);
    } catch (const ArgFail& e) {
        g_create_error = e.msg; st = SOLAROSC_ERR_ARG;
    } catch (const CudaFail& e) {
        g_create_error = e.msg; st = SOLAROSC_ERR_CUDA;
    } catch (const std::bad_alloc&) {
        g_create_error = "host allocation failed"; st = SOLAROSC_ERR_ALLOC;
    } catch (...) {
        g_create_error = "unexpected exception"; st = SOLAROSC_ERR_ARG;
    }
    if (st != SOLAROSC_OK && h) {
        solarosc_destroy(h);
        h = nullptr;
    }
    if (status) *status = st;
    return h;
}

void solarosc_destroy(solarosc_handle* h) {
    if (!h) return;
    if (h->comm) comm_release(h, false);
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (void* p : h->dev_allocs) cudaFree(p);
    if (h->d_scratch) cudaFree(h->d_scratch);
    if (h->d_gradws) cudaFree(h->d_gradws);
    if (h->small_graph) cudaGraphExecDestroy(h->small_graph);
    if (h->h_params_pin) cudaFreeHost(h->h_params_pin);
    for (int i = 0; i < 2; ++i)
        if (h->h_grad_pin[i]) cudaFreeHost(h->h_grad_pin[i]);
    for (int i = 0; i < 2; ++i) {
        if (h->h_lnl2[i]) cudaFreeHost(h->h_lnl2[i]);
        if (h->h_dstage[i]) cudaFreeHost(h->h_dstage[i]);
        if (h->h_flags2[i]) cudaFreeHost(h->h_flags2[i]);
    }
    for (int i = 0; i < 2; ++i) {
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
        if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
    }
    for (auto& e : h->tev) cudaEventDestroy(e);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

const char* solarosc_last_error(const solarosc_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int solarosc_n_params(const solarosc_handle* h) { return h ? h->P : -1; }

int solarosc_set_data(solarosc_handle* h, int det, const double* es_day, const double* es_night, const double* cc_day,
                      const double* cc_night) {
    return guarded(h, [&]() {
        REQUIRE(det >= 0 && det < h->T.n_det, "detector index out of range");
        REQUIRE(es_day && es_night && cc_day && cc_night, "data pointers must not be NULL");
        CK(cudaStreamSynchronize(h->stream));
        upload_data(h, det, es_day, es_night, cc_day, cc_night);
        if (h->small_graph) {   // the captured kernel arguments carry data-dependent constants
            cudaGraphExecDestroy(h->small_graph);
            h->small_graph = nullptr;
        }
        return (int)SOLAROSC_OK;
    });
}

int solarosc_eval_llh(solarosc_handle* h, const double* params, size_t n, double* out, uint32_t* flags) {
    return guarded(h, [&]() { return n == 0 ? (int)SOLAROSC_OK : eval_llh_host(h, params, n, 1, out, flags); });
}

int solarosc_eval_llh_grouped(solarosc_handle* h, const double* params, size_t n, size_t group, double* out,
                              uint32_t* flags) {
    return guarded(h, [&]() { return n == 0 ? (int)SOLAROSC_OK : eval_llh_host(h, params, n, group, out, flags); });
}

// One process driving several GPUs: the rows are dealt to the handles in contiguous slices (whole groups), each slice
// is evaluated by its own host thread on its own device, and every slice's lnL lands directly in the caller's array.
int solarosc_eval_llh_multi(solarosc_handle* const* hs, size_t n_handles, const double* params, size_t n, size_t group,
                            double* out, uint32_t* flags) {
    if (!hs || n_handles == 0) return SOLAROSC_ERR_ARG;
    for (size_t k = 0; k < n_handles; ++k)
        if (!hs[k] || hs[k]->P != hs[0]->P) return SOLAROSC_ERR_ARG;
    if (n == 0) return SOLAROSC_OK;
    if (!params || !out || group == 0 || n % group) return SOLAROSC_ERR_ARG;
    const size_t n_groups = n / group, base = n_groups / n_handles, extra = n_groups % n_handles;
    std::vector<int> status(n_handles, SOLAROSC_OK);
    std::vector<std::thread> workers;
    const size_t P = (size_t)hs[0]->P;
    for (size_t k = 0; k < n_handles; ++k) {
        const size_t lo = (k * base + std::min(k, extra)) * group, hi = lo + (base + (k < extra ? 1 : 0)) * group;
        if (hi == lo) continue;
        workers.emplace_back([=, &status]() {
            status[k] = guarded(hs[k], [&]() { return eval_llh_host(hs[k], params + lo * P, hi - lo, group, out + lo, flags ? flags + lo : nullptr); });
        });
    }
    for (auto& w : workers) w.join();
    int worst = SOLAROSC_OK;
    for (int st : status)
        if (st != SOLAROSC_OK && (worst == SOLAROSC_OK || worst == SOLAROSC_ERR_NEGATIVE_RATE)) worst = st;
    return worst;
}

int solarosc_eval_llh_grad(solarosc_handle* h, const double* params, size_t n, double* out_lnl, double* out_grad,
                           uint32_t* flags) {
    return guarded(h, [&]() { return n == 0 ? (int)SOLAROSC_OK : eval_llh_grad_host(h, params, n, out_lnl, out_grad, flags); });
}

int solarosc_eval_rates(solarosc_handle* h, int det, const double* params, size_t n, double* es_day, double* es_night,
                        double* cc_day, double* cc_night, double* bg_es, double* bg_cc) {
    return guarded(h, [&]() {
        return n == 0 ? (int)SOLAROSC_OK
                      : eval_arrays_host(h, det, kModeRates, params, n, es_day, es_night, cc_day, cc_night, bg_es, bg_cc);
    });
}

int solarosc_eval_perbin(solarosc_handle* h, int det, const double* params, size_t n, double* es_day, double* es_night,
                         double* cc_day, double* cc_night) {
    return guarded(h, [&]() {
        return n == 0 ? (int)SOLAROSC_OK
                      : eval_arrays_host(h, det, kModePerbin, params, n, es_day, es_night, cc_day, cc_night, nullptr, nullptr);
    });
}

int solarosc_eval_asymmetry(solarosc_handle* h, int det, const double* params, size_t n, double* asym_es, double* asym_cc) {
    return guarded(h, [&]() {
        if (n == 0) return (int)SOLAROSC_OK;
        REQUIRE(params, "params must not be NULL");
        REQUIRE(det >= 0 && det < h->T.n_det, "detector index out of range");
        REQUIRE(h->rates_ok, "binning too large for the fold's shared-memory tile");
        h->ensure_scratch((size_t)2 * h->chunk * sizeof(double));
        h->launches = 0;
        for (size_t b0 = 0; b0 < n; b0 += h->chunk) {
            const int nb = (int)std::min<size_t>(h->chunk, n - b0);
            CK(cudaMemcpyAsync(h->d_params, params + b0 * h->P, (size_t)nb * h->P * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            RatesOut O{};
            O.asym_es = h->d_scratch; O.asym_cc = h->d_scratch + h->chunk;
            launch_llh(h, h->d_params, nb, 1, O, h->stream, nullptr, kModeAsym, det);
            if (asym_es) CK(cudaMemcpyAsync(asym_es + b0, O.asym_es, nb * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            if (asym_cc) CK(cudaMemcpyAsync(asym_cc + b0, O.asym_cc, nb * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
        }
        return (int)SOLAROSC_OK;
    });
}

// Derived-mode block reduction (derive_variables_from_chain.jl:260-444).  Per chunk: oscillation stage, rates to
// device scratch, per-sample totals (derived_sample_kernel) and the per-bin weighted sums (derived_moment_kernel,
// carried across chunks in the slice partials); the scalar sums are accumulated here, sample by sample, from the
// per-sample totals.
int solarosc_eval_derived(solarosc_handle* h, int det, const double* params, size_t n, const double* weights,
                          const double* lref_es, const double* lref_cc, double* es_day, double* cc_day, double* ltot,
                          double* mom_es, double* mom_cc, double* bg_mom_es, double* bg_mom_cc, double* scalars) {
    return guarded(h, [&]() {
        REQUIRE(params && weights, "params and weights must not be NULL");
        REQUIRE(det >= 0 && det < h->T.n_det, "detector index out of range");
        check_data(h, det);
        const DevTables& T = h->T;
        const DevDetector& D = T.det[det];
        const int nZ = T.n_cosz, nE = h->n_es[det], nC = h->n_cc[det];
        // angular ES reconstruction: the ES arrays carry the cos-scatter axis, [day: nCos][nE] then [night: nCos][nE][nZ]
        const int nCos = D.ang ? D.n_cos : 1;
        const int nel_es = (1 + nZ) * nE * nCos, nel_cc = (1 + nZ) * nC, n_elem = nel_es + nel_cc + nE + nC;
        const int n_slices = 32;
        const size_t per = (size_t)2 * (nel_es + nel_cc) + nE + nC + (D.ang ? nE : 0);   // rates, centred per-bin terms, backgrounds
        const int step = std::max(1, std::min(h->chunk, D.ang ? 2048 : 8192));
        const size_t fixed = (size_t)nel_es + nel_cc + (size_t)(n_slices + 1) * kDerivedQ * n_elem;
        const size_t slot_doubles = per * step + 5 * (size_t)step;
        h->ensure_scratch((2 * slot_doubles + fixed) * sizeof(double));
        double* d_lref_es = h->d_scratch;
        double* d_lref_cc = d_lref_es + nel_es;
        double* d_partial = d_lref_cc + nel_cc;
        double* d_out = d_partial + (size_t)n_slices * kDerivedQ * n_elem;
        double* d_slots = d_out + (size_t)kDerivedQ * n_elem;
        // per-sample outputs come back through two pinned staging slots: [step][4 + nE + nC]
        const size_t stage_bytes = (size_t)step * (4 + nE + nC) * sizeof(double);
        if (stage_bytes > h->h_dstage_bytes) {
            for (int i = 0; i < 2; ++i) {
                if (h->h_dstage[i]) cudaFreeHost(h->h_dstage[i]);
                h->h_dstage[i] = nullptr;
                CK(cudaMallocHost(reinterpret_cast<void**>(&h->h_dstage[i]), stage_bytes));
            }
            h->h_dstage_bytes = stage_bytes;
        }
        if (lref_es) CK(cudaMemcpyAsync(d_lref_es, lref_es, nel_es * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        if (lref_cc) CK(cudaMemcpyAsync(d_lref_cc, lref_cc, nel_cc * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        std::vector<double> acc(21, 0.0);
        h->launches = 0;
        if (n == 0) CK(cudaMemsetAsync(d_out, 0, (size_t)kDerivedQ * n_elem * sizeof(double), h->stream));
        // Chunks are pipelined over two slots: the kernels of chunk c + 1 run while the per-sample outputs of chunk c travel
        // to the pinned staging (s_out) and the host adds chunk c - 1's scalar sums, sample by sample, in sample order.
        struct Slot { size_t b0; int nb; };
        Slot sl[2];
        auto collect = [&](int slot) {
            const Slot& q = sl[slot];
            CK(cudaEventSynchronize(h->ev_out[slot]));
            const double* lt = h->h_dstage[slot];
            const double* ed = lt + (size_t)4 * step;
            const double* cd = ed + (size_t)step * nE;
            if (es_day) std::memcpy(es_day + q.b0 * nE, ed, (size_t)q.nb * nE * sizeof(double));
            if (cc_day) std::memcpy(cc_day + q.b0 * nC, cd, (size_t)q.nb * nC * sizeof(double));
            for (int i = 0; i < q.nb; ++i) {   // derive...:343-348, 388-417, in sample order
                const double w = weights[q.b0 + i], t1 = params[(q.b0 + i) * h->P + 0], t2 = params[(q.b0 + i) * h->P + 2];
                acc[0] += w; acc[1] += w * t1; acc[2] += w * (t1 * t1); acc[3] += w * t2; acc[4] += w * (t2 * t2);
                for (int g = 0; g < 4; ++g) {
                    const double l = lt[(size_t)4 * i + g];
                    acc[5 + 4 * g + 0] += w * l; acc[5 + 4 * g + 1] += w * (l * l);
                    acc[5 + 4 * g + 2] += w * (l * t1); acc[5 + 4 * g + 3] += w * (l * t2);
                }
            }
            if (ltot) std::memcpy(ltot + 4 * q.b0, lt, (size_t)4 * q.nb * sizeof(double));
        };
        size_t c = 0;
        for (size_t b0 = 0; b0 < n; b0 += step, ++c) {
            const int nb = (int)std::min<size_t>(step, n - b0), slot = (int)(c & 1);
            if (c >= 2) {
                collect(slot);                                                    // chunk c - 2 has left this slot's staging
                CK(cudaStreamWaitEvent(h->stream, h->ev_out[slot], 0));           // ... and its device scratch
            }
            double* d_w = d_slots + (size_t)slot * slot_doubles;
            double* d_ltot = d_w + step;
            double* d_rates = d_ltot + 4 * (size_t)step;
            double* d_par = h->d_params2[slot];
            CK(cudaMemcpyAsync(d_par, params + b0 * h->P, (size_t)nb * h->P * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            CK(cudaMemcpyAsync(d_w, weights + b0, nb * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            double* sp = d_rates;
            RatesOut O{};
            O.es_day = sp; sp += (size_t)nb * nE * nCos;
            O.es_night = sp; sp += (size_t)nb * nZ * nE * nCos;
            O.cc_day = sp; sp += (size_t)nb * nC;
            O.cc_night = sp; sp += (size_t)nb * nZ * nC;
            O.bg_es = sp; sp += (size_t)nb * nE;
            O.bg_cc = sp; sp += (size_t)nb * nC;
            double* d_es_day_sum = D.ang ? sp : nullptr;
            if (D.ang) sp += (size_t)nb * nE;
            double* d_lbuf = sp;
            // rates with row-major night arrays whatever the handle's host layout is: the fused pipe_kernel when the chunk
            // fills the GPU (the probability block stays on the SM), else osc_kernel + rates_kernel
            if (!launch_pipe(h, d_par, nb, 1, kModeRates, det, O, h->stream, true)) {
                launch_osc(h, d_par, nb, 1, nullptr, h->stream);
                RatesLaunch L = rates_launch(h, nb, 1, det);
                L.s_row_es = nE; L.s_j_es = 1; L.s_row_cc = nC; L.s_j_cc = 1;
                const size_t sm = rates_smem_bytes(T.n_etrue, T.n_cosz, h->max_ereco);
                const GenLaunch GL{nE, 1, (long long)nE * nZ, nZ, 1};   // row-major [c][j] and [c][j][z]
                if (h->gen) gen_kernel<kModeRates><<<nb, kThreads, sm, h->stream>>>(T, L, GL, d_par, h->osc(), O);
                else rates_kernel<kModeRates><<<nb, kThreads, sm, h->stream>>>(T, L, d_par, h->osc(), O);
                CK(cudaGetLastError());
                ++h->launches;
            }
            DerivedLaunch L{nb, n_slices, h->P, det, nZ, nE, nC, D.index_es - 1, D.index_cc - 1, b0 == 0 ? 1 : 0, T.earth_unc, T.n_layers, nCos};
            DerivedIn I{O.es_day, O.es_night, O.cc_day, O.cc_night, O.bg_es, O.bg_cc, D.obs_es, D.obs_cc,
                        lref_es ? d_lref_es : nullptr, lref_cc ? d_lref_cc : nullptr, d_par, d_w, D.ol_es_ang, d_es_day_sum, d_lbuf};
            derived_sample_kernel<<<nb, 256, 0, h->stream>>>(L, I, d_ltot);
            derived_moment_kernel<<<dim3((n_elem + 255) / 256, n_slices), 256, 0, h->stream>>>(L, I, d_partial);
            CK(cudaGetLastError());
            h->launches += 2;
            CK(cudaEventRecord(h->ev_done[slot], h->stream));
            CK(cudaStreamWaitEvent(h->s_out, h->ev_done[slot], 0));
            double* st = h->h_dstage[slot];
            CK(cudaMemcpyAsync(st, d_ltot, (size_t)4 * nb * sizeof(double), cudaMemcpyDeviceToHost, h->s_out));
            if (es_day) CK(cudaMemcpyAsync(st + (size_t)4 * step, D.ang ? d_es_day_sum : O.es_day, (size_t)nb * nE * sizeof(double), cudaMemcpyDeviceToHost, h->s_out));
            if (cc_day) CK(cudaMemcpyAsync(st + (size_t)4 * step + (size_t)step * nE, O.cc_day, (size_t)nb * nC * sizeof(double), cudaMemcpyDeviceToHost, h->s_out));
            CK(cudaEventRecord(h->ev_out[slot], h->s_out));
            sl[slot] = Slot{b0, nb};
        }
        for (size_t k = c >= 2 ? c - 2 : 0; k < c; ++k) collect((int)(k & 1));
        if (n > 0) {
            derived_finish_kernel<<<(kDerivedQ * n_elem + 255) / 256, 256, 0, h->stream>>>(n_slices, n_elem, d_partial, d_out);
            CK(cudaGetLastError());
            ++h->launches;
        }
        std::vector<double> out((size_t)kDerivedQ * n_elem);
        CK(cudaMemcpyAsync(out.data(), d_out, out.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int q = 0; q < kDerivedQ; ++q) {
            const double* src = out.data() + (size_t)q * n_elem;
            if (mom_es) std::memcpy(mom_es + (size_t)q * nel_es, src, nel_es * sizeof(double));
            if (mom_cc) std::memcpy(mom_cc + (size_t)q * nel_cc, src + nel_es, nel_cc * sizeof(double));
            if (q < 2) {
                if (bg_mom_es) std::memcpy(bg_mom_es + (size_t)q * nE, src + nel_es + nel_cc, nE * sizeof(double));
                if (bg_mom_cc) std::memcpy(bg_mom_cc + (size_t)q * nC, src + nel_es + nel_cc + nE, nC * sizeof(double));
            }
        }
        if (scalars) std::memcpy(scalars, acc.data(), 21 * sizeof(double));
        return (int)SOLAROSC_OK;
    });
}

int solarosc_eval_cc_incl_spectrum(solarosc_handle* h, int det, const double* params, size_t n, double* out) {
    return guarded(h, [&]() {
        REQUIRE(det >= 0 && det < h->T.n_det, "detector index out of range");
        REQUIRE(h->T.det[det].incl != 0, "CC_incl_spectrum exists for the inclusive and semi-inclusive analyses only (the reference returns nothing otherwise)");
        if (n == 0) return (int)SOLAROSC_OK;
        REQUIRE(params && out, "params and out must not be NULL");
        const int nE = h->n_es[det];
        h->ensure_scratch((size_t)h->chunk * nE * sizeof(double));
        h->launches = 0;
        for (size_t b0 = 0; b0 < n; b0 += h->chunk) {
            const int nb = (int)std::min<size_t>(h->chunk, n - b0);
            CK(cudaMemcpyAsync(h->d_params, params + b0 * h->P, (size_t)nb * h->P * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            launch_osc(h, h->d_params, nb, 1, nullptr, h->stream);
            cc_incl_spectrum_kernel<<<nb, kThreads, 2 * h->T.n_etrue * sizeof(double), h->stream>>>(h->T, det, h->P, h->d_params, h->osc(), h->d_scratch);
            CK(cudaGetLastError());
            ++h->launches;
            CK(cudaMemcpyAsync(out + b0 * nE, h->d_scratch, (size_t)nb * nE * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
        }
        return (int)SOLAROSC_OK;
    });
}

int solarosc_eval_oscprobs(solarosc_handle* h, const double* params, size_t n, double* day_8b, double* night_8b,
                           double* day_hep, double* night_hep) {
    return guarded(h, [&]() {
        if (n == 0) return (int)SOLAROSC_OK;
        REQUIRE(params, "params must not be NULL");
        const int nT = h->T.n_etrue, nZ = h->T.n_cosz, nrow = 1 + nZ;
        std::vector<double> host((size_t)h->chunk * 2 * nrow * nT);
        h->launches = 0;
        for (size_t b0 = 0; b0 < n; b0 += h->chunk) {
            const int nb = (int)std::min<size_t>(h->chunk, n - b0);
            CK(cudaMemcpyAsync(h->d_params, params + b0 * h->P, (size_t)nb * h->P * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            launch_osc(h, h->d_params, nb, 1, nullptr, h->stream);
            CK(cudaMemcpyAsync(host.data(), h->osc(), (size_t)nb * 2 * nrow * nT * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            for (int i = 0; i < nb; ++i) {
                const double* s8 = host.data() + (size_t)i * 2 * nrow * nT;
                const double* sh = s8 + (size_t)nrow * nT;
                const size_t b = b0 + i;
                if (day_8b) std::copy(s8, s8 + nT, day_8b + b * nT);
                if (day_hep) std::copy(sh, sh + nT, day_hep + b * nT);
                if (night_8b) std::copy(s8 + nT, s8 + (size_t)nrow * nT, night_8b + b * nZ * nT);
                if (night_hep) std::copy(sh + nT, sh + (size_t)nrow * nT, night_hep + b * nZ * nT);
            }
        }
        if (h->cfg.col_major) {
            if (night_8b) transpose_blocks(night_8b, n, nZ, nT);
            if (night_hep) transpose_blocks(night_hep, n, nZ, nT);
        }
        return (int)SOLAROSC_OK;
    });
}

int solarosc_eval_p1e(solarosc_handle* h, const double* params, size_t n, double* p1e) {
    return guarded(h, [&]() {
        if (n == 0) return (int)SOLAROSC_OK;
        REQUIRE(params && p1e, "params and p1e must not be NULL");
        const size_t per = (size_t)h->T.n_paths * h->T.n_e;
        const int step = std::max(1, std::min(h->chunk, 256));
        h->ensure_scratch(per * step * sizeof(double));
        h->launches = 0;
        for (size_t b0 = 0; b0 < n; b0 += step) {
            const int nb = (int)std::min<size_t>(step, n - b0);
            CK(cudaMemcpyAsync(h->d_params, params + b0 * h->P, (size_t)nb * h->P * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            launch_osc(h, h->d_params, nb, 1, h->d_scratch, h->stream);
            CK(cudaStreamSynchronize(h->stream));
            CK(cudaMemcpy(p1e + b0 * per, h->d_scratch, (size_t)nb * per * sizeof(double), cudaMemcpyDeviceToHost));
        }
        if (h->cfg.col_major) transpose_blocks(p1e, n, h->T.n_paths, h->T.n_e);
        return (int)SOLAROSC_OK;
    });
}

int solarosc_eval_llh_device(solarosc_handle* h, const double* d_params, size_t n, size_t group, double* d_lnl,
                             uint32_t* d_flags, void* stream) {
    return guarded(h, [&]() {
        if (n == 0) return (int)SOLAROSC_OK;
        REQUIRE(d_params && d_lnl, "device pointers must not be NULL");
        if (group == 0) group = 1;
        REQUIRE(n % group == 0 && group <= (size_t)h->chunk, "n_points must be a multiple of group (<= max_batch)");
        check_data(h, -1);
        cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : h->stream;
        const size_t step = ((size_t)h->chunk / group) * group;
        h->launches = 0;
        h->tev_used = 0;
        h->tev_stream = st;
        for (size_t b0 = 0; b0 < n; b0 += step) {
            const int nb = (int)std::min(step, n - b0);
            RatesOut O{};
            O.lnl = d_lnl + b0;
            O.flags = d_flags ? d_flags + b0 : nullptr;
            if (h->timing) CK(cudaEventRecord(h->next_tev(), st));
            cudaEvent_t mid = h->timing ? h->next_tev() : nullptr;
            h->last_fused = launch_llh(h, d_params + b0 * h->P, nb, (int)group, O, st, mid);
            if (h->timing) CK(cudaEventRecord(h->next_tev(), st));
        }
        return (int)SOLAROSC_OK;
    });
}

int solarosc_poisson_llh(solarosc_handle* h, const double* expected, const double* measured, size_t n_vec, size_t len,
                         double* out, uint32_t* out_flags) {
    return guarded(h, [&]() {
        if (n_vec == 0) return (int)SOLAROSC_OK;
        REQUIRE(expected && measured && out, "NULL pointer");
        const size_t tot = n_vec * len;
        // enough CTAs for every SM: long vectors are cut into segments (at least 2048 terms each)
        const int n_split = (int)std::max<size_t>(1, std::min<size_t>(((size_t)4 * h->sm_count + n_vec - 1) / n_vec, (len + 2047) / 2048));
        const size_t n_part = n_vec * n_split;
        h->ensure_scratch((2 * tot + n_vec + n_part) * sizeof(double) + (n_vec + n_part) * sizeof(uint32_t));
        double* de = h->d_scratch;
        double* dm = de + tot;
        double* dout = dm + tot;
        double* dpart = dout + n_vec;
        uint32_t* df = reinterpret_cast<uint32_t*>(dpart + n_part);
        uint32_t* dpf = df + n_vec;
        CK(cudaMemcpyAsync(de, expected, tot * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(dm, measured, tot * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        poisson_kernel<<<(unsigned)n_part, 256, 0, h->stream>>>(de, dm, len, n_split, dpart, dpf);
        poisson_finish_kernel<<<(unsigned)((n_vec + 127) / 128), 128, 0, h->stream>>>(dpart, dpf, n_vec, n_split, dout, df);
        CK(cudaGetLastError());
        h->launches = 2;
        CK(cudaMemcpyAsync(out, dout, n_vec * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        std::vector<uint32_t> fl(n_vec);
        CK(cudaMemcpyAsync(fl.data(), df, n_vec * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        bool neg = false;
        for (size_t i = 0; i < n_vec; ++i) {
            if (out_flags) out_flags[i] = fl[i];
            neg |= (fl[i] & SOLAROSC_FLAG_NEG_ERROR) != 0;
        }
        return neg ? (int)SOLAROSC_ERR_NEGATIVE_RATE : (int)SOLAROSC_OK;
    });
}

int solarosc_set_timing(solarosc_handle* h, int enabled) {
    if (!h) return SOLAROSC_ERR_ARG;
    std::lock_guard<std::mutex> lk(h->mu);
    h->timing = enabled != 0;
    return SOLAROSC_OK;
}

int solarosc_last_timing(solarosc_handle* h, float* ms_osc, float* ms_rates) {
    if (!h) return -1;
    std::lock_guard<std::mutex> lk(h->mu);
    float a = 0.f, c = 0.f;
    if (h->timing && h->tev_used >= 3) {
        cudaSetDevice(h->device);
        if (cudaEventSynchronize(h->tev[h->tev_used - 1]) == cudaSuccess)
            for (size_t i = 0; i + 2 < h->tev_used; i += 3) {
                float x = 0.f, y = 0.f;
                cudaEventElapsedTime(&x, h->tev[i], h->tev[i + 1]);
                cudaEventElapsedTime(&y, h->tev[i + 1], h->tev[i + 2]);
                a += x; c += y;
            }
    }
    if (ms_osc) *ms_osc = a;
    if (ms_rates) *ms_rates = c;
    return h->launches;
}

int solarosc_last_fused(const solarosc_handle* h) { return h ? (h->last_fused ? 1 : 0) : -1; }

#ifdef SOLAROSC_PHASE_TIMING
// diagnostic builds only: clock64 stamps of CTAs kPhaseFirst .. kPhaseFirst + kPhaseCtas - 1 of the last fused launch
int solarosc_debug_read_phases(long long* out, int* n_slots, int* n_ctas) {
    cudaDeviceSynchronize();
    *n_slots = kPhaseSlots; *n_ctas = kPhaseCtas;
    return cudaMemcpyFromSymbol(out, g_phase, sizeof(long long) * kPhaseSlots * kPhaseCtas) == cudaSuccess ? 0 : 2;
}
#endif

#ifdef SOLAROSC_PIPE_STATS
// diagnostic builds only: read and clear the pipe_kernel cycle counters
int solarosc_debug_read_pipe_stats(unsigned long long* out) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out, g_pipe_stats, sizeof(unsigned long long) * 16) != cudaSuccess) return 2;
    unsigned long long z[16] = {0};
    return cudaMemcpyToSymbol(g_pipe_stats, z, sizeof z) == cudaSuccess ? 0 : 2;
}
#endif

int solarosc_measure_fp64_peak(solarosc_handle* h, double* tflops) {
    return guarded(h, [&]() {
        REQUIRE(tflops, "NULL pointer");
        double* d = nullptr;
        CK(cudaMalloc(reinterpret_cast<void**>(&d), sizeof(double) * 1024));
        const int iters = 1 << 14, blocks = h->sm_count * 8, threads = 256;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float best = 1e30f;
        for (int rep = 0; rep < 6; ++rep) {
            CK(cudaEventRecord(e0, h->stream));
            dfma_peak_kernel<<<blocks, threads, 0, h->stream>>>(d, iters, 1.0000001);
            CK(cudaEventRecord(e1, h->stream));
            CK(cudaEventSynchronize(e1));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0) best = std::min(best, ms);
        }
        CK(cudaGetLastError());
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        CK(cudaFree(d));
        const double fl = 2.0 * 16.0 * (double)iters * blocks * threads;   // 16 independent FMA chains per thread
        *tflops = fl / (best * 1e-3) * 1e-12;
        return (int)SOLAROSC_OK;
    });
}

}  // extern "C"

// ---- set-up stage (solarosc_setup.cuh) ----------------------------------------------------------------------------
extern "C" {

int solarosc_setup_bin_index(const double* edges, size_t n_edges, const double* v, size_t n, int32_t* out) {
    if (!edges || n_edges == 0 || (n && (!v || !out))) return SOLAROSC_ERR_ARG;
    for (size_t i = 0; i < n; ++i) out[i] = solarosc_setup::search_last(edges, (int)n_edges, v[i]);
    return SOLAROSC_OK;
}

int solarosc_setup_bin_index_device(solarosc_handle* h, const double* edges, size_t n_edges, const double* v, size_t n, int32_t* out) {
    return guarded(h, [&]() {
        REQUIRE(edges && n_edges > 0 && n_edges <= 4096 && (n == 0 || (v && out)), "bad arguments");
        if (n == 0) return (int)SOLAROSC_OK;
        // edges, values and results live in the handle's scratch (grown on demand, freed with the handle)
        const size_t ne8 = (n_edges + 1) & ~(size_t)1;
        h->ensure_scratch((ne8 + n) * sizeof(double) + n * sizeof(int32_t));
        double* d_e = h->d_scratch;
        double* d_v = d_e + ne8;
        int32_t* d_o = reinterpret_cast<int32_t*>(d_v + n);
        CK(cudaMemcpyAsync(d_e, edges, n_edges * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(d_v, v, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        const int grid = (int)std::min<size_t>((n + 255) / 256, (size_t)h->sm_count * 8);
        solarosc_setup::bin_index_kernel<<<grid, 256, n_edges * sizeof(double), h->stream>>>(d_e, (int)n_edges, d_v, n, d_o);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(out, d_o, n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return (int)SOLAROSC_OK;
    });
}

int solarosc_setup_histogram(const double* v, const double* w, const uint8_t* keep, size_t n, int32_t n_bins, double lo, double hi,
                             const double* edges, int32_t normalise, double* out) {
    if (n_bins <= 0 || !out || (n && !v)) return SOLAROSC_ERR_ARG;
    try {
        solarosc_setup::histogram(v, w, keep, n, solarosc_setup::make_bins(n_bins, lo, hi, edges), normalise != 0, out);
    } catch (...) { return SOLAROSC_ERR_ALLOC; }
    return SOLAROSC_OK;
}

int solarosc_setup_response_matrix(const double* x, const double* y, size_t n, int32_t nbx, double xlo, double xhi,
                                   const double* xedges, int32_t nby, double ylo, double yhi, const double* yedges, double* out,
                                   int32_t* xi, int32_t* yi) {
    if (nbx <= 0 || nby <= 0 || !out || (n && (!x || !y))) return SOLAROSC_ERR_ARG;
    try {
        solarosc_setup::response_matrix(x, y, n, solarosc_setup::make_bins(nbx, xlo, xhi, xedges),
                                        solarosc_setup::make_bins(nby, ylo, yhi, yedges), out, xi, yi);
    } catch (...) { return SOLAROSC_ERR_ALLOC; }
    return SOLAROSC_OK;
}

int solarosc_setup_earth_paths(const double* radius, const double* potential, size_t n_model, const double* cosz, size_t n_paths,
                               int32_t n_layers_expected, int32_t* seg_start, int32_t* seg_layer, double* seg_length,
                               size_t seg_capacity, double* lookup, int32_t* n_layers_found) {
    if (!radius || !potential || n_model < 2 || !cosz || !seg_start || !seg_layer || !seg_length) return SOLAROSC_ERR_ARG;
    try {
        solarosc_setup::PathBuilder pb;
        if (!pb.init(radius, potential, n_model)) return SOLAROSC_ERR_ARG;
        if (n_layers_found) *n_layers_found = (int32_t)pb.layers.size();
        if (n_layers_expected > 0 && (int)pb.layers.size() != n_layers_expected) return SOLAROSC_ERR_ARG;
        std::vector<std::vector<solarosc_setup::Segment>> paths(n_paths);
        size_t total = 0;
        seg_start[0] = 0;
        for (size_t p = 0; p < n_paths; ++p) {
            if (!pb.path(cosz[p], paths[p])) return SOLAROSC_ERR_ARG;
            if (total + paths[p].size() > seg_capacity) return SOLAROSC_ERR_ARG;
            for (const auto& sg : paths[p]) { seg_layer[total] = sg.index; seg_length[total] = sg.length; ++total; }
            seg_start[p + 1] = (int32_t)total;
        }
        if (lookup) {   // earth_lookup padded from the vertical path (setup.jl:52-66)
            std::vector<double> base = solarosc_setup::avg_densities(paths);
            const size_t n_exp = n_layers_expected > 0 ? (size_t)n_layers_expected : base.size();
            if (base.size() < n_exp) {
                std::vector<std::vector<solarosc_setup::Segment>> vert(1);
                if (!pb.path(-1.0, vert[0])) return SOLAROSC_ERR_ARG;
                const std::vector<double> ref = solarosc_setup::avg_densities(vert);
                for (size_t k = base.size(); k < n_exp && k < ref.size(); ++k) base.push_back(ref[k]);
            }
            for (size_t k = 0; k < base.size(); ++k) lookup[k] = base[k];
        }
    } catch (...) { return SOLAROSC_ERR_ALLOC; }
    return SOLAROSC_OK;
}

}  // extern "C"

#include "solarosc_batcher.hpp"
#include "solarosc_comm.hpp"
