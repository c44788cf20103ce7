// solarosc_kernels.cuh -- the batched kernels of the hot path.
//
//   osc_stage   K1 + K2: adiabatic MSW day probability (mean density or production-region integral)
//               and Earth-matter P1e through the layered path table, fused with the IYSW night
//               formula and the (N_COSZ_SUB x 2) block average.
//   fold_stage  K3 + K4 + K5: flux x cross-section x exposure weighting, E_true -> E_reco fold for
//               ES nu_e / ES nu_other / CC on the FP64 tensor pipe (DMMA m8n8k4), background
//               addition and the soft-floored Poisson log-likelihood reduced with warp shuffles.
//
//   llh_kernel    one CTA per oscillation point: osc_stage into shared memory, then fold_stage for
//                 every parameter point that shares it (1 in general, n_sys x 3 for an LLH-scan cell).
//                 The oscillation probabilities never leave the SM.
//   osc_kernel    osc_stage alone, energy axis optionally split over CTAs, probabilities to HBM
//                 (stage-separated API, small batches).
//   rates_kernel  fold_stage alone on probabilities read back from HBM.
//
// Reference code each stage stands in for is cited at the stage.
#pragma once
#include "solarosc_device.cuh"

namespace solarosc {

#ifndef SOLAROSC_OSC_THREADS
#define SOLAROSC_OSC_THREADS 256
#endif
#ifndef SOLAROSC_MIRROR_UNROLL_PAR
#define SOLAROSC_MIRROR_UNROLL_PAR 1   // both fine energies of a coarse bin get their own copy of the pair walk: +1.4 % (profiles/r02_notes.md)
#endif
#ifndef SOLAROSC_EIG_NODUP
#define SOLAROSC_EIG_NODUP 1   // every eigen entry computed once (entries 2t, 2t + 1 per thread): +0.4 %
#endif
#ifndef SOLAROSC_OSC_MINBLOCKS
#define SOLAROSC_OSC_MINBLOCKS 2
#endif
#ifndef SOLAROSC_LLH_MINBLOCKS
#define SOLAROSC_LLH_MINBLOCKS SOLAROSC_OSC_MINBLOCKS
#endif
#ifndef SOLAROSC_RATES_MINBLOCKS
#define SOLAROSC_RATES_MINBLOCKS 2
#endif
constexpr int kThreads = SOLAROSC_OSC_THREADS;   // every kernel here uses this block size
constexpr int kWarps = kThreads / 32;
constexpr int kMC = 7;   // row tiles a fold keeps in flight (51 rows = 7 tiles in the default binning)
#ifndef SOLAROSC_FOLD_MCA
#define SOLAROSC_FOLD_MCA 5
#define SOLAROSC_FOLD_MCB 2
#endif
constexpr int kMCa = SOLAROSC_FOLD_MCA, kMCb = SOLAROSC_FOLD_MCB;   // grouped evaluation: finer work items (fold_stage)   // row tiles per work item of the forward fold: groups of kMCa, remainder in groups of kMCb

struct OscLaunch {
    int n_points;     // oscillation points in this launch
    int n_chunks;     // energy chunks per point
    int cb;           // coarse energy bins per chunk
    int param_stride; // doubles between consecutive points' parameter rows
    int point_step;   // parameter row of oscillation point g is g*point_step (grouped evaluation)
    int seg_smem;     // path tables staged in shared memory behind the eigen table
};

// Phase timestamps of a window of CTAs (diagnostic builds only: -DSOLAROSC_PHASE_TIMING, read back with
// solarosc_debug_read_phases; tools/phase_timing.py prints the averages).
#ifdef SOLAROSC_PHASE_TIMING
constexpr int kPhaseSlots = 16, kPhaseCtas = 4096, kPhaseFirst = 8192;
__device__ long long g_phase[kPhaseSlots * kPhaseCtas];
#define SOLAROSC_PH(i)                                                                                   \
    do {                                                                                                 \
        if (threadIdx.x == 0 && blockIdx.x >= kPhaseFirst && blockIdx.x < kPhaseFirst + kPhaseCtas) {    \
            g_phase[(blockIdx.x - kPhaseFirst) * kPhaseSlots + (i)] = clock64();                         \
            if ((i) == 0) {   /* slot 15: the SM this CTA runs on (pairs co-resident CTAs: tools/phase_pairs.py) */ \
                unsigned smid_;                                                                          \
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_));                                       \
                g_phase[(blockIdx.x - kPhaseFirst) * kPhaseSlots + 15] = smid_;                          \
            }                                                                                            \
        }                                                                                                \
    } while (0)
#else
#define SOLAROSC_PH(i)
#endif

enum RatesMode { kModeLlh = 0, kModeRates = 1, kModePerbin = 2, kModeAsym = 3 };

struct RatesLaunch {
    int n_points;
    int group;         // points per shared oscillation stage
    int param_stride;
    int det_only;      // -1: all detectors (likelihood), else the one detector whose arrays are written
    int max_ereco;     // largest reco-bin count of any detector/channel (sizes the shared-memory tiles)
    int seg_smem;      // fused kernel: path tables staged in shared memory behind the scratch
    int osc_cb;        // fused kernel: coarse energy bins per oscillation chunk
    long long s_row_es, s_j_es, s_row_cc, s_j_cc;  // output strides of the night arrays (layout)
};

struct RatesOut {
    double* lnl; uint32_t* flags;
    double *es_day, *es_night, *cc_day, *cc_night, *bg_es, *bg_cc;
    double *asym_es, *asym_cc;   // kModeAsym: day-night asymmetry per point and channel
};

// ---- shared-memory geometry ------------------------------------------------------------------------
__host__ __device__ inline size_t osc_smem_doubles(int n_layers, int cb) {
    return (size_t)kEntry * n_layers * 2 * cb + (size_t)4 * 2 * cb;
}
__host__ __device__ inline size_t osc_smem_bytes(int n_layers, int cb) { return sizeof(double) * osc_smem_doubles(n_layers, cb); }

// Oscillation probabilities as the A operand of the fold: [proc][rows padded to 8][ldp] doubles.  ldp is
// the K extent rounded up to 4, plus padding to 12 (mod 16): the 8 rows x 4 energies a DMMA A-fragment
// load touches then fall into distinct bank octets.
__host__ __device__ inline int fold_kt(int n_etrue) { return (n_etrue + 3) / 4; }
__host__ __device__ inline int fold_ldp(int n_etrue) {
    int l = 4 * fold_kt(n_etrue);
    while (l % 16 != 12) l += 4;
    return l;
}
__host__ __device__ inline int fold_rows(int nrow) { return 8 * ((nrow + 7) / 8); }
__host__ __device__ inline size_t fold_p_doubles(int n_etrue, int nrow) { return (size_t)2 * fold_rows(nrow) * fold_ldp(n_etrue); }
// scratch of the fold: CC scale vectors [2][4 kt], additive constants and background sums [4][8 nt], then the
// per-warp slots of the closing reduction (kWarps doubles + kWarps flag words)
__host__ __device__ inline size_t fold_vec_doubles(int n_etrue, int max_ereco) {
    return (size_t)2 * 4 * fold_kt(n_etrue) + (size_t)4 * 8 * ((max_ereco + 7) / 8);
}
__host__ __device__ inline size_t fold_scratch_doubles(int n_etrue, int max_ereco) {
    // an even number of doubles: the path tables behind the scratch are read with 16-byte loads
    return (fold_vec_doubles(n_etrue, max_ereco) + kWarps + (kWarps + 1) / 2 + 1) & ~(size_t)1;
}
__host__ __device__ inline size_t rates_smem_bytes(int n_etrue, int n_cosz, int max_ereco) {
    return sizeof(double) * (fold_p_doubles(n_etrue, 1 + n_cosz) + fold_scratch_doubles(n_etrue, max_ereco));
}
// Fused kernel: eigen table + MSW of one energy chunk (cb coarse bins), probability tile, scratch, path tables.
// The probability tile holds exactly 1 + n_cosz rows per process; the last 8-row MMA tile of a process reads past
// them into whatever follows (the other process, the scratch, the path tables) -- harmless, because an output row
// of an MMA depends on the same row of A only and the epilogue drops rows >= 1 + n_cosz -- and `tail` pads the
// allocation in the odd case that what follows is shorter than that overrun.
__host__ __device__ inline size_t llh_smem_bytes(int n_layers, int cb, int n_etrue, int n_cosz, int max_ereco, size_t seg_doubles) {
    const int nrow = 1 + n_cosz, ldp = fold_ldp(n_etrue);
    const size_t over = (size_t)(fold_rows(nrow) - nrow) * ldp, after = fold_scratch_doubles(n_etrue, max_ereco) + seg_doubles;
    const size_t tail = over > after ? over - after : 0;
    return sizeof(double) * (osc_smem_doubles(n_layers, cb) + (size_t)2 * nrow * ldp + fold_scratch_doubles(n_etrue, max_ereco) +
                             seg_doubles + tail);
}

// ------------------------------------------------------------------------------------------
// K1 + K2.  `sink.put(row, i, p8, ph)` receives the block-averaged nu_e survival probabilities of
// coarse energy bin i for row 0 (day) and rows 1 + z (night zenith bin z).  Ends without a barrier.
// ------------------------------------------------------------------------------------------
struct GlobalSink {
    double *out8, *outh;
    int n_etrue;
    __device__ __forceinline__ void put(int row, int i, double p8, double ph) const {
        out8[(size_t)row * n_etrue + i] = p8;
        outh[(size_t)row * n_etrue + i] = ph;
    }
};
struct SmemSink {
    double *p8s, *phs;
    int ldp;
    __device__ __forceinline__ void put(int row, int i, double p8, double ph) const {
        p8s[row * ldp + i] = p8;
        phs[row * ldp + i] = ph;
    }
};

// Path tables as the propagation loop reads them: in shared memory when they fit (every layer step
// needs its layer id and length), else in HBM.  Rows are listed longest path first.
struct SegTables {
    const Seg* seg;         // [n_seg] applied steps of all paths
    const int2* path_se;    // [n_cosz * n_sub] (first, one-past-last) step of the paths of sorted row zr
    const int* row_order;   // [n_cosz] zenith bin of sorted row zr
    const double2* mir;     // [n_seg] (dr, drh) of the interior steps of the first half of a mirrored path (solarosc_device.cuh)
};
// One step's (layer, length).  SMEM: the tables are known to sit in shared memory, so the load is an LDS instead
// of a generic load that has to resolve its address space first -- it feeds the table address of the next step.
template <bool SMEM>
__device__ __forceinline__ Seg load_seg(const SegTables& G, int s) {
    if (!SMEM) return G.seg[s];
    Seg r;
    long long bits;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=d"(r.len), "=l"(bits) : "r"((uint32_t)__cvta_generic_to_shared(G.seg + s)));
    r.layer = (int)bits; r.pad = (int)(bits >> 32);
    return r;
}
template <bool SMEM>
__device__ __forceinline__ double2 load_mir(const SegTables& G, int s) {
    if (!SMEM) return G.mir[s];
    double2 r;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"((uint32_t)__cvta_generic_to_shared(G.mir + s)));
    return r;
}

__host__ __device__ inline size_t seg_smem_doubles(int n_seg, int n_paths, int n_cosz) {
    return (size_t)4 * n_seg + n_paths + ((size_t)n_cosz + 1) / 2 + 1;
}

// Copies the path tables behind `buf` (16-byte aligned; no barrier: the caller's next __syncthreads orders it).
__device__ __forceinline__ SegTables stage_seg_tables(const DevTables& T, double* buf, bool to_smem) {
    if (!to_smem) return SegTables{T.seg, T.path_se, T.row_order, T.seg_mir};
    Seg* seg = reinterpret_cast<Seg*>(buf);
    double2* mir = reinterpret_cast<double2*>(buf + 2 * T.n_seg);
    int2* se = reinterpret_cast<int2*>(buf + 4 * T.n_seg);
    int* order = reinterpret_cast<int*>(buf + 4 * T.n_seg + T.n_paths);
    // first pass with all three loads in flight before any store (one L2 round trip instead of three)
    const int t0 = threadIdx.x;
    Seg sg{};
    double2 mr{};
    int2 pe{};
    int ro = 0;
    if (t0 < T.n_seg) { sg = T.seg[t0]; mr = T.seg_mir[t0]; }
    if (t0 < T.n_paths) pe = T.path_se[t0];
    if (t0 < T.n_cosz) ro = T.row_order[t0];
    if (t0 < T.n_seg) { seg[t0] = sg; mir[t0] = mr; }
    if (t0 < T.n_paths) se[t0] = pe;
    if (t0 < T.n_cosz) order[t0] = ro;
    for (int t = t0 + blockDim.x; t < T.n_seg; t += blockDim.x) { seg[t] = T.seg[t]; mir[t] = T.seg_mir[t]; }
    for (int t = t0 + blockDim.x; t < T.n_paths; t += blockDim.x) se[t] = T.path_se[t];
    for (int t = t0 + blockDim.x; t < T.n_cosz; t += blockDim.x) order[t] = T.row_order[t];
    return SegTables{seg, se, order, mir};
}

// P1e = |<nu_1|X|nu_e>|^2 (osc.jl:436-446) of sorted path (s0, s1) at the two fine energies of a coarse bin (eigen
// entries t0 / t1 at layer 0, ls doubles between layers).  One-layer paths: the closed form; mirrored paths
// (flagged on their first step at create): the pair walk of solarosc_device.cuh, one fine energy after the other
// (psi and chi are the two dependency chains in flight); anything else: the generic walk with both fine energies
// in flight.  The (layer, length) of the next step is fetched while the current one is computed.
template <bool SMEM>
__device__ __forceinline__ void path_p1e_pair(const SegTables& G, const double* t0, const double* t1, int ls, int s0, int s1,
                                              double u1, double u2, double u3, double& pe0, double& pe1) {
    Seg sg = load_seg<SMEM>(G, s0);
    const int k = sg.layer * ls;
    const double len = sg.len;
    if (s1 - s0 == 1) {
        State pa, pb;
        step_only(t0 + k, len, u1, pa);
        step_only(t1 + k, len, u1, pb);
        pe0 = fma(pa.ar, pa.ar, pa.ai * pa.ai);
        pe1 = fma(pb.ar, pb.ar, pb.ai * pb.ai);
    } else if (sg.pad != 0) {
        const double lenx = load_seg<SMEM>(G, s1 - 1).len;
        const int h = (s1 - s0) >> 1;
#if SOLAROSC_MIRROR_UNROLL_PAR
#pragma unroll
#else
#pragma unroll 1
#endif
        for (int par = 0; par < 2; ++par) {
            const double* t = par ? t1 : t0;
            State p, x;
            Seg sn = load_seg<SMEM>(G, s0 + 1);
            double2 mn = load_mir<SMEM>(G, s0 + 1);
            step_first_pair(t + k, len, lenx, u1, u2, u3, p, x);
            for (int i = 1; i < h; ++i) {
                const int km = sn.layer * ls;
                const double lm = sn.len, dr = mn.x, drh = mn.y;
                sn = load_seg<SMEM>(G, s0 + i + 1);
                mn = load_mir<SMEM>(G, s0 + i + 1);
                step_pair(t + km, lm, dr, drh, p, x);
            }
            const double pe = step_middle(t + sn.layer * ls, sn.len, p, x);
            if (par) pe1 = pe; else pe0 = pe;
        }
    } else {
        State pa, pb;
        sg = load_seg<SMEM>(G, s0 + 1);
        step_first(t0 + k, len, u1, u2, u3, pa);
        step_first(t1 + k, len, u1, u2, u3, pb);
        for (int s = s0 + 1; s < s1 - 1; ++s) {
            const int km = sg.layer * ls;
            const double lm = sg.len;
            sg = load_seg<SMEM>(G, s + 1);
            step_mid(t0 + km, lm, pa);
            step_mid(t1 + km, lm, pb);
        }
        const int kl = sg.layer * ls;
        step_last(t0 + kl, sg.len, pa);
        step_last(t1 + kl, sg.len, pb);
        pe0 = fma(pa.ar, pa.ar, pa.ai * pa.ai);
        pe1 = fma(pb.ar, pb.ar, pb.ai * pb.ai);
    }
}

__device__ __forceinline__ double sqrt_pos(double x);
// square root of a number in [0, 1] (exactly 0 at 0), straight-line
__device__ __forceinline__ double sqrt01(double x) { return x > 0.0 ? sqrt_pos(x) : 0.0; }

template <bool SEG_SMEM, class Sink>
__device__ __forceinline__ void osc_stage(const DevTables& T, const SegTables G, const double* __restrict__ p, int i0, int cb,
                                          int cbs, double* tbl, double* msw, double* __restrict__ p1e_out, const Sink sink) {
    const int ne = 2 * cb;   // fine energies handled here
    const int nl = T.n_ltbl;
    // mixing parameters (reference propagation_osc.jl:30-36: theta = asin(sqrt(clamp(sin2, 0, 1))))
    // clamp(x, 0, 1) that lets a NaN through like Julia's does (fmin / fmax would turn it into 0)
    auto clamp01 = [](double x) { return x != x ? x : fmin(fmax(x, 0.0), 1.0); };
    const double s12sq = clamp01(p[0]);
    const double s13sq = clamp01(p[1]);
    const double dm21 = p[2];
    const double c13sq = 1.0 - s13sq, c12sq = 1.0 - s12sq;
    const double c13 = sqrt01(c13sq);
    const double u1 = sqrt01(c12sq) * c13, u2 = sqrt01(s12sq) * c13, u3 = sqrt01(s13sq);
    const double p0 = c12sq * c13sq;                     // P_0 = c12^2 c13^2 (osc.jl:196)
    const double c2 = 1.0 - 2.0 * s12sq;                 // cos(2 theta12)

    SOLAROSC_PH(1);
    // ---- phase 1a: one eigen-system per (fine energy, layer)  (osc.jl:466-470, 370-388) --------
    // two entries per thread and pass: make_entry is branch-free, so the two dependency chains interleave
#if SOLAROSC_EIG_NODUP
    for (int t = 2 * threadIdx.x; t < ne * nl; t += 2 * blockDim.x) {   // entries 2t, 2t + 1: every entry computed once
        const int tb = min(t + 1, ne * nl - 1);
#else
    for (int t = threadIdx.x; t < ne * nl; t += 2 * blockDim.x) {
        const int tb = min(t + (int)blockDim.x, ne * nl - 1);   // a thread without a second entry repeats the last one
#endif
        const int ka = t / ne, ela = t - ka * ne, kb = tb / ne, elb = tb - kb * ne;
        const double ea = T.e_calc[2 * i0 + ela], eb = T.e_calc[2 * i0 + elb];
        const double na = T.earth_unc ? p[SOLAROSC_N_FIXED_PARAMS + ka] : 1.0;
        const double nb = T.earth_unc ? p[SOLAROSC_N_FIXED_PARAMS + kb] : 1.0;
        // lookup = earth_norm .* earth_lookup (propagation_osc.jl:60)
        const double aa = na * T.earth_lookup[ka] * ea, ab = nb * T.earth_lookup[kb] * eb;
        double enta[kEntry], entb[kEntry];
        make_entry(dm21, aa, ea, u1, u2, u3, enta);
        make_entry(dm21, ab, eb, u1, u2, u3, entb);
        double* da = tbl + ((size_t)ka * 2 * cbs + (ela & 1) * cbs + (ela >> 1)) * kEntry;
        double* db = tbl + ((size_t)kb * 2 * cbs + (elb & 1) * cbs + (elb >> 1)) * kEntry;
#pragma unroll
        for (int x = 0; x < kEntry; ++x) { da[x] = enta[x]; db[x] = entb[x]; }   // same values when tb repeats another thread's entry
    }
    // ---- phase 1b: MSW day probability and cos(2 theta_m) per (process, fine energy) ------------
    if (T.fast) {  // osc.jl:185-207 with n_e = production-weighted mean density
        for (int t = threadIdx.x; t < 2 * ne; t += blockDim.x) {
            const int proc = t / ne, el = t - proc * ne;
            const double e = T.e_calc[2 * i0 + el];
            const double c2m = msw_c2m(e, T.avg_ne[proc], dm21, c13sq, c2);
            const int slot = (el & 1) * cbs + (el >> 1);
            msw[(size_t)(proc * 2 + 0) * 2 * cbs + slot] = fma(0.5 * c13sq * c13sq, fma(c2, c2m, 1.0), s13sq * s13sq);
            msw[(size_t)(proc * 2 + 1) * 2 * cbs + slot] = c13sq * c2m;
        }
    } else {  // osc.jl:139-182: the radius loop is linear in (P_day, cos2theta_m), so both collapse to
              // production-weighted means of cos(2 theta_m); one warp per fine energy, lanes over radii.  The matter
              // angle depends on (E, n_e(r)) only, so both processes share it and differ in their weights; the
              // division by dm21 is taken out of the radius loop (x = cos2theta - n_e(r) kE, one FMA).
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
        const double s2sq = 1.0 - c2 * c2;
        const double* __restrict__ sne = T.solar_ne;
        const double* __restrict__ w8 = T.solar_w[0];
        const double* __restrict__ wh = T.solar_w[1];
        for (int el = warp; el < ne; el += nw) {
            const double e = T.e_calc[2 * i0 + el];
            const double kE = 2.0 * 1.4142135623730951 * kGF * e * c13sq / dm21;
            double a8 = 0.0, ah = 0.0;
#pragma unroll 4
            for (int r = lane; r < T.n_solar; r += 32) {
                const double x = fma(-sne[r], kE, c2);
                const double c = x * rsq(fma(x, x, s2sq));
                a8 = fma(w8[r], c, a8);
                ah = fma(wh[r], c, ah);
            }
            a8 = warp_sum(a8);
            ah = warp_sum(ah);
            if (lane == 0) {
                const int slot = (el & 1) * cbs + (el >> 1);
#pragma unroll
                for (int proc = 0; proc < 2; ++proc) {
                    const double c2m = (proc ? ah : a8) / T.solar_wsum[proc];
                    msw[(size_t)(proc * 2 + 0) * 2 * cbs + slot] = fma(0.5 * c13sq * c13sq, fma(c2, c2m, 1.0), s13sq * s13sq);
                    msw[(size_t)(proc * 2 + 1) * 2 * cbs + slot] = c13sq * c2m;
                }
            }
        }
    }
    // work counter of phase 2 (its previous readers all left their loop before the barrier that ended the last call)
    __shared__ int s_next;
    if (threadIdx.x == 0) s_next = 0;
    SOLAROSC_PH(2);
    __syncthreads();
    SOLAROSC_PH(3);

    // ---- day rows: block_average(prob_day, 2)  (propagation_osc.jl:133-134) -----------------------
    for (int j = threadIdx.x; j < cb; j += blockDim.x)
        sink.put(0, i0 + j, 0.5 * (msw[j] + msw[cbs + j]),
                 0.5 * (msw[(size_t)2 * 2 * cbs + j] + msw[(size_t)2 * 2 * cbs + cbs + j]));

    // ---- phase 2: Earth propagation, one item = (coarse zenith row, coarse energy bin) -----------
    // Lanes of a warp walk the same path at neighbouring energies, so the layer loop is warp-uniform.
    // Rows are listed longest path first and every warp draws its next 32 items from a shared counter when it is
    // done with the previous ones (longest-processing-time-first): the warps of a CTA reach the closing barrier
    // within one short path of each other.  The next draw is issued before the current items are computed.
    const double inv_block = 1.0 / (2.0 * T.n_sub);
    const int n_items = T.n_cosz * cb;
    const double inv_cb = 1.0 / cb;
    const int ls = 2 * cbs * kEntry;                            // doubles between consecutive layers
    const int lane = threadIdx.x & 31;
    int drawn = 0;
    if (lane == 0) drawn = atomicAdd(&s_next, 1);
    for (;;) {
        const int wi = __shfl_sync(0xffffffffu, drawn, 0);
        if (wi * 32 >= n_items) break;
        if (lane == 0) drawn = atomicAdd(&s_next, 1);
        const int item = wi * 32 + lane;
        if (item >= n_items) continue;
        const int zr = (int)((item + 0.5) * inv_cb), j = item - zr * cb;   // exact: item < 2^31, cb >= 1
        double acc8 = 0.0, acch = 0.0;
        const double* t0 = tbl + (size_t)j * kEntry;           // even fine energy of this coarse bin, layer 0
        const double* t1 = t0 + (size_t)cbs * kEntry;          // odd fine energy
        for (int sub = 0; sub < T.n_sub; ++sub) {
            const int2 se = G.path_se[zr * T.n_sub + sub];
            const int s0 = se.x, s1 = se.y;
            double pe0, pe1;
            path_p1e_pair<SEG_SMEM>(G, t0, t1, ls, s0, s1, u1, u2, u3, pe0, pe1);
            if (p1e_out) {
                double* o = p1e_out + (size_t)(G.row_order[zr] * T.n_sub + sub) * T.n_e + 2 * (i0 + j);
                o[0] = pe0; o[1] = pe1;
            }
            // IYSW night probability (osc.jl:204): P_day + c13^2 cos2theta_m (P1e - P0)
            const double d0 = pe0 - p0, d1 = pe1 - p0;
            acc8 += fma(msw[(size_t)1 * 2 * cbs + j], d0, msw[j]) + fma(msw[(size_t)1 * 2 * cbs + cbs + j], d1, msw[cbs + j]);
            acch += fma(msw[(size_t)3 * 2 * cbs + j], d0, msw[(size_t)2 * 2 * cbs + j]) +
                    fma(msw[(size_t)3 * 2 * cbs + cbs + j], d1, msw[(size_t)2 * 2 * cbs + cbs + j]);
        }
        sink.put(1 + G.row_order[zr], i0 + j, acc8 * inv_block, acch * inv_block);   // block_average (propagation_osc.jl:136-137)
    }
    SOLAROSC_PH(4);
}

__global__ void __launch_bounds__(kThreads, SOLAROSC_OSC_MINBLOCKS)
osc_kernel(const DevTables T, const OscLaunch L, const double* __restrict__ params,
           double* __restrict__ osc_out /* [n_points][2][1+n_cosz][n_etrue] */,
           double* __restrict__ p1e_out /* [n_points][n_paths][n_e] or nullptr */) {
    extern __shared__ __align__(16) double smem[];
    const int g = blockIdx.x / L.n_chunks;               // oscillation point
    const int chunk = blockIdx.x - g * L.n_chunks;
    const double* p = params + (size_t)g * L.point_step * L.param_stride;
    const int i0 = chunk * L.cb;                         // first coarse energy bin of this CTA
    const int cb = min(L.cb, T.n_etrue - i0);            // coarse bins handled here
    const int nrow = 1 + T.n_cosz;
    double* tbl = smem;                                  // [nl][2 parity][cbs] entries of kEntry doubles
    double* msw = smem + (size_t)kEntry * T.n_ltbl * 2 * L.cb;   // [2 proc][2 what][2 parity][cbs]
    const SegTables G = stage_seg_tables(T, smem + osc_smem_doubles(T.n_ltbl, L.cb), L.seg_smem != 0);
    GlobalSink sink{osc_out + (size_t)g * 2 * nrow * T.n_etrue, osc_out + ((size_t)g * 2 + 1) * nrow * T.n_etrue, T.n_etrue};
    double* p1e = p1e_out ? p1e_out + (size_t)g * T.n_paths * T.n_e : nullptr;
    if (L.seg_smem) osc_stage<true>(T, G, p, i0, cb, L.cb, tbl, msw, p1e, sink);
    else osc_stage<false>(T, G, p, i0, cb, L.cb, tbl, msw, p1e, sink);
}

// ------------------------------------------------------------------------------------------
// K3 + K4 + K5
// ------------------------------------------------------------------------------------------
// Natural logarithm of a positive, finite, normal double (every soft-floored expectation is one): the
// fdlibm algorithm -- x = 2^k m, m in [sqrt(1/2), sqrt(2)), log m = 2s + s R(s^2), s = (m-1)/(m+1) -- with
// the division done as reciprocal seed + Newton steps.  Error below 1 ulp, about a third of the
// instructions of the general-purpose log() with its special-case handling.
__device__ __forceinline__ double log_pos(double x) {
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int k = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;
    if (hi > 0x3ff6a09e) { hi -= 0x00100000; k += 1; }
    const double f = __hiloint2double(hi, lo) - 1.0;
    const double t = 2.0 + f;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(t));
    r = fma(fma(-t, r, 1.0), r, r);
    r = fma(fma(-t, r, 1.0), r, r);
    double sq = f * r;
    sq = fma(fma(-t, sq, f), r, sq);                       // s = f / (2 + f), correctly rounded to ~1 ulp
    const double z = sq * sq, w = z * z;
    const double t1 = w * fma(w, fma(w, 1.531383769920937332e-01, 2.222219843214978396e-01), 3.999999999940941908e-01);
    const double t2 = z * fma(w, fma(w, fma(w, 1.479819860511658591e-01, 1.818357216161805012e-01), 2.857142874366239149e-01),
                              6.666666666666735130e-01);
    const double R = t2 + t1, hfsq = 0.5 * f * f, dk = (double)k;
    return fma(dk, 6.93147180369123816490e-01, -((hfsq - fma(sq, hfsq + R, dk * 1.90821492927058770002e-10)) - f));
}

// Square root of a positive normal double (the soft floor's argument is >= 1e-18): reciprocal-square-root
// seed, one coupled Newton step for sqrt and 1/(2 sqrt), one more for sqrt and a final residual correction.
// Branch-free, so the Poisson terms a lane holds interleave; within 1 ulp.
__device__ __forceinline__ double sqrt_pos(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double g = x * r, h = 0.5 * r;
    double e = fma(-h, g, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    e = fma(-h, g, 0.5);
    g = fma(g, e, g);
    return fma(fma(-g, g, x), h, g);
}

// soft-floored Poisson term (reference likelihood_statistical.jl:16-24), without branches: the logarithm is
// taken of the floored expectation, which is positive whatever raw_e is.
__device__ __forceinline__ double poisson_term(double raw_e, double m, double logm, uint32_t& fl) {
    if (raw_e < 0.0) fl |= (raw_e > -1e-3) ? SOLAROSC_FLAG_NEG_WARN : SOLAROSC_FLAG_NEG_ERROR;
    const double d = raw_e - 1e-9;
    const double e = 0.5 * (raw_e + 1e-9 + sqrt_pos(fma(d, d, 1e-18)));
    const double t = e - m + m * (logm - log_pos(e));
    return (m > 0.0) ? t : e;
}

// per-bin diagnostic term (reference likelihood_statistical.jl:31-57)
__device__ __forceinline__ double perbin_term(double raw_e, double m) {
    if (m > 0.0) {
        const double d = raw_e - 1e-9;
        const double es = 0.5 * (raw_e + 1e-9 + sqrt(fma(d, d, 1e-18)));
        return -(raw_e - m * log(es));
    }
    return -raw_e;
}

// block-wide sum, valid on thread 0 (two barriers; red holds one double per warp)
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kWarps; ++w) s += red[w];
    return s;
}

// D(8x8) += A(8x4) * B(4x8) on the FP64 tensor pipe.  Lane l holds A[l>>2][l&3], B[l&3][l>>2] and
// D[l>>2][2*(l&3) + {0,1}].
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// The response fold is linear in the oscillation probabilities, so everything that does not depend on
// the parameter point is folded into constant matrices at create time (host: build_fold_tables):
//   ES[row][j] = w_row * ( F8 * sum_i P8[row][i] A8[i][j] + Fh * sum_i Ph[row][i] Ah[i][j] + cst_es[j] )
//   CC[row][j] = w_row * ( F8 * sum_i P8[row][i] s_i B8[i][j] + Fh * sum_i Ph[row][i] s_i Bh[i][j] + cst_cc[j] )
// with A = 0.5 (eff_nue u_nue R_nue - eff_nuo u_nuo R_nuo), B = 0.5 eff_cc u_cc R_cc, s_i the CC
// cross-section shape, w_0 = 1 (day) and w_{1+z} the exposure weight, cst_es = F8 C8 + Fh Ch + BG_ES/2,
// cst_cc = BG_CC/2 (reference propagation_osc.jl:159-210, propagation_core.jl:73-79,
// propagation_reco.jl:10-29,117-130 rearranged; same arithmetic up to rounding order).
//
// The products run on the FP64 tensor pipe (DMMA m8n8k4).  One warp owns one (channel, 8 reco bins)
// column block with all (up to kMC) row tiles in flight.  Flux and cross-section shape are applied to the
// B fragment (one multiply per k-step and process instead of one per A element), both processes
// accumulate into the same tile and the accumulators start from the additive constants, so per k-step a
// warp issues 2 global loads (prefetched one step ahead), 2 + 2*kMC shared loads, 2 multiplies and 2*kMC
// DMMAs.  The epilogue consumes the 8x8 output tiles straight from the accumulator registers -- Poisson
// terms, asymmetry sums, or the rate / per-bin arrays -- so the folded rates never touch shared memory and
// every lane carries 2*kMC independent logarithm chains.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// Per-point vectors of the fold of detector D: the B-fragment scale vectors (flux, CC cross-section shape:
// propagation_osc.jl:185-196) and the additive constants with the backgrounds (propagation_bg.jl:36-83; every
// ES component uses norm_1).  No barriers inside.
template <int MODE>
__device__ __forceinline__ void fold_vectors(const DevTables& T, const DevDetector& D, const double* __restrict__ p, size_t b,
                                             double* scr, const RatesOut& O) {
    const int nT = T.n_etrue, kt = fold_kt(nT);
    const int nte = (D.n_es + 7) / 8, ntc = (D.n_cc + 7) / 8;
    const double f8 = p[3], fh = p[4];
    const double* pd = p + D.param_off;
    double* scale = scr;
    double* cst_es = scr + 8 * kt;
    double* cst_cc = cst_es + 8 * nte;
    double* bgs_es = cst_cc + 8 * ntc;
    double* bgs_cc = bgs_es + 8 * nte;
    // ---- per-point vectors: CC cross-section shape (propagation_osc.jl:185-196), backgrounds
    //      (propagation_bg.jl:36-83; every ES component uses norm_1) ----------------------------------
    for (int i = threadIdx.x; i < 4 * kt; i += blockDim.x) {
        double sh = 0.0;
        if (i < nT) {
            const double l = D.log_e_norm[i];
            sh = exp(fma(p[6], l, p[7] * (l * l))) * p[5];
        }
        scale[i] = f8 * sh; scale[4 * kt + i] = fh * sh;
    }
    for (int j = threadIdx.x; j < 8 * (nte + ntc); j += blockDim.x) {
        double acc = 0.0, c = 0.0;
        if (j < 8 * nte) {
            if (j < D.n_es) {
                for (int q = 0; q < D.n_bg_es; ++q) {
                    const double raw = D.raw_bg_es[q * D.n_es + j];
                    acc += (D.bg_es_area[q] != 0.0) ? raw * (D.bg_es_area[q] * pd[0]) : raw;
                }
                c = fma(f8, D.c8[j], fma(fh, D.ch[j], 0.5 * acc));   // F8 C8 + Fh Ch + BG/2
                if (MODE == kModeRates && O.bg_es) O.bg_es[b * D.n_es + j] = acc;
            }
            cst_es[j] = c;
            if (MODE == kModeAsym) bgs_es[j] = acc;
        } else {
            const int jc = j - 8 * nte;
            if (jc < D.n_cc) {
                for (int q = 0; q < D.n_bg_cc; ++q) {
                    const double raw = D.raw_bg_cc[q * D.n_cc + jc];
                    acc += (D.bg_cc_norm[q] >= 0) ? raw * (pd[D.n_norm_es + D.bg_cc_norm[q]] * (1.0 / 2.2e-6)) : raw;
                }
                c = 0.5 * acc;
                if (MODE == kModeRates && O.bg_cc) O.bg_cc[b * D.n_cc + jc] = acc;
            }
            cst_cc[jc] = c;
            if (MODE == kModeAsym) bgs_cc[jc] = acc;
        }
    }
}

// One work item of the fold: column block `col` (ES blocks first, then CC) against NT consecutive row tiles
// starting at tile m0, all NT accumulator tiles in flight, epilogue on the accumulator registers.
template <int MODE>
struct FoldTask {
    const DevTables& T;
    const RatesLaunch& L;
    const DevDetector& D;
    const RatesOut& O;
    const double *p8s, *phs, *scale, *cst_es, *cst_cc;
    double f8, fh;
    size_t b;
    int ldp, nrow, mt, kt, bte, btc;
    bool oob;
    const double* cst_blk = nullptr;   // pipe_kernel: the additive constants of this column block only, [8]
    int row0 = 0;                      // rates_tiled_kernel: the tile holds rows row0 .. row0 + nrow - 1 of the arrays

    template <int NT>
    __device__ __forceinline__ void run(int col, int m0, double& part, uint32_t& fl, double& d_es, double& n_es, double& d_cc,
                                        double& n_cc) const {
        constexpr bool kThr = MODE == kModeLlh || MODE == kModeAsym;
        const int lane = threadIdx.x & 31, fg = lane >> 2, ft = lane & 3;
        const int nZ = T.n_cosz;   // (nrow may be a row chunk: rates_tiled_kernel)
        const bool is_es = col < bte;
        const int n0 = 8 * (is_es ? col : col - bte);
        const int ldb = 8 * (is_es ? bte : btc);
        const int nR = is_es ? D.n_es : D.n_cc;
        const int idx0 = (is_es ? D.index_es : D.index_cc) - 1;
        const int j0 = (kThr ? idx0 : 0) + n0 + 2 * ft;                  // this lane's two reco bins
        const bool v0 = j0 < nR, v1 = j0 + 1 < nR;
        const double* m8 = (kThr ? (is_es ? D.a8_thr : D.b8_thr) : (is_es ? D.a8 : D.b8)) + ft * ldb + n0 + fg;   // B fragment of k-step 0
        const double* mh = (kThr ? (is_es ? D.ah_thr : D.bh_thr) : (is_es ? D.ah : D.bh)) + ft * ldb + n0 + fg;
        const uint32_t sc8 = smem_addr(scale + ft), sch = sc8 + 4u * kt * 8u;
        const double* cb = cst_blk ? cst_blk + 2 * ft : (is_es ? cst_es : cst_cc) + j0;
        const double c0 = v0 ? cb[0] : 0.0, c1 = v1 ? cb[1] : 0.0;
        const double2* ol = is_es ? D.ol_es : D.ol_cc;                   // (m, log m) pairs, [row][nR]
        const double* obs = is_es ? D.obs_es : D.obs_cc;
        double* day = is_es ? O.es_day : O.cc_day;
        double* night = is_es ? O.es_night : O.cc_night;
        const long long sr = is_es ? L.s_row_es : L.s_row_cc, sj = is_es ? L.s_j_es : L.s_j_cc;
        const size_t bstep = (size_t)4 * ldb;
        const uint32_t tile_bytes = 8u * (uint32_t)ldp * 8u;   // one 8-row tile of the probability tile
        double acc[NT][2];
        uint32_t a8[NT];   // shared-memory address of this lane's A element of k-step 0, per row tile
#pragma unroll
        for (int c = 0; c < NT; ++c) {
            acc[c][0] = c0; acc[c][1] = c1;
            // tiles past the end repeat the last one: their results are dropped in the epilogue
            a8[c] = smem_addr(p8s + fg * ldp + ft) + (uint32_t)min(m0 + c, mt - 1) * tile_bytes;
        }
        const uint32_t dproc = smem_addr(phs) - smem_addr(p8s);
        double b8n = __ldg(m8), bhn = __ldg(mh);
        const double *q8 = m8, *qh = mh;
        for (int ks = 0; ks < kt; ++ks) {
            const double b8v = b8n * (is_es ? f8 : lds_f64(sc8 + 32u * ks)), bhv = bhn * (is_es ? fh : lds_f64(sch + 32u * ks));
            if (ks + 1 < kt) { q8 += bstep; qh += bstep; b8n = __ldg(q8); bhn = __ldg(qh); }
#pragma unroll
            for (int c = 0; c < NT; ++c) {
                dmma_8x8x4(acc[c][0], acc[c][1], lds_f64(a8[c] + 32u * ks), b8v);
                dmma_8x8x4(acc[c][0], acc[c][1], lds_f64(a8[c] + dproc + 32u * ks), bhv);
            }
        }
        // ---- epilogue on the accumulator registers ----------------------------------------------
        double r[NT][2];
        bool vr[NT];
#pragma unroll
        for (int c = 0; c < NT; ++c) {
            const int row = 8 * (m0 + c) + fg, grow = row0 + row;
            vr[c] = m0 + c < mt && row < nrow;
            const double w = (grow == 0) ? 1.0 : T.wz[min(grow, T.n_cosz) - 1];
            r[c][0] = w * acc[c][0];
            r[c][1] = w * acc[c][1];
        }
        if (MODE == kModeLlh) {
            double2 o0[NT], o1[NT];
#pragma unroll
            for (int c = 0; c < NT; ++c) {
                const int grow = row0 + 8 * (m0 + c) + fg;
                if (T.night_only && grow == 0) vr[c] = false;   // night-only likelihood: the day row stays out of the sum
                o0[c] = (vr[c] && v0) ? __ldg(ol + grow * nR + j0) : make_double2(0.0, 0.0);
                o1[c] = (vr[c] && v1) ? __ldg(ol + grow * nR + j0 + 1) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int c = 0; c < NT; ++c) {
                const double t0 = poisson_term((vr[c] && v0) ? r[c][0] : 1.0, o0[c].x, o0[c].y, fl);
                const double t1 = poisson_term((vr[c] && v1) ? r[c][1] : 1.0, o1[c].x, o1[c].y, fl);
                part += ((vr[c] && v0) ? t0 : 0.0) + ((vr[c] && v1) ? t1 : 0.0);
            }
        } else if (MODE == kModeAsym) {
#pragma unroll
            for (int c = 0; c < NT; ++c) {
                const double s01 = ((vr[c] && v0) ? r[c][0] : 0.0) + ((vr[c] && v1) ? r[c][1] : 0.0);
                const bool is_day = row0 + 8 * (m0 + c) + fg == 0;
                if (is_es) { if (is_day) d_es += s01; else n_es += s01; }
                else { if (is_day) d_cc += s01; else n_cc += s01; }
            }
        } else {
#pragma unroll
            for (int c = 0; c < NT; ++c) {
                if (!vr[c]) continue;
                const int row = row0 + 8 * (m0 + c) + fg;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int jx = j0 + h;
                    if (jx >= nR) continue;
                    double v = r[c][h];
                    if (MODE == kModePerbin) {
                        v = (jx >= idx0 && jx < nR - 1) ? perbin_term(v, obs[row * nR + jx]) : 0.0;
                        if (oob) v = -INFINITY;
                    }
                    if (row == 0) {
                        if (day) day[b * nR + jx] = v;
                    } else if (night) {
                        night[b * nZ * nR + (size_t)(row - 1) * sr + (size_t)jx * sj] = v;
                    }
                }
            }
        }
    }
};

template <int MODE, bool FINE = false>
__device__ __forceinline__ void fold_stage(const DevTables& T, const RatesLaunch& L, const double* __restrict__ p, size_t b,
                                           const double* p8s, const double* phs, int ldp, double* scr,
                                           const RatesOut& O, bool vectors_ready = false) {
    // vectors_ready: fold_vectors of detector 0 already ran for this point (the fused kernel does it next to the
    // eigen table, where its load and exp latencies hide behind other warps' work)
    const int nT = T.n_etrue, nZ = T.n_cosz, nrow = 1 + nZ;
    const int kt = fold_kt(nT), mt = (nrow + 7) / 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // scale vectors of the CC B fragments, [4*kt] each: 8B (F8 s_i), hep (Fh s_i); ES scales by F8 / Fh alone
    double* scale = scr;
    double* cst_es = scr + 8 * kt;           // per reco bin additive constants
    double* red = scr + fold_vec_doubles(nT, L.max_ereco);
    uint32_t* redf = reinterpret_cast<uint32_t*>(red + kWarps);
    const double f8 = p[3], fh = p[4];
    // earth-norm bounds (reference likelihood_debug.jl:1-4, likelihood_builder.jl:227)
    bool oob = false;
    if (T.earth_unc)
        for (int k = 0; k < T.n_layers; ++k) {
            const double x = p[SOLAROSC_N_FIXED_PARAMS + k];
            oob |= !(x >= 0.0 && x <= 2.0);
        }
    double total = 0.0;
    uint32_t fl = 0;
    for (int di = 0; di < T.n_det; ++di) {
        if (L.det_only >= 0 && di != L.det_only) continue;
        const DevDetector& D = T.det[di];
        const int nte = (D.n_es + 7) / 8, ntc = (D.n_cc + 7) / 8;
        double* cst_cc = cst_es + 8 * nte;
        double* bgs_es = cst_cc + 8 * ntc;       // summed backgrounds per reco bin (asymmetry mode)
        double* bgs_cc = bgs_es + 8 * nte;
        __syncthreads();   // the previous point / detector is done with the scratch; orders the probability tile
        SOLAROSC_PH(5);
        const bool have_vectors = vectors_ready && di == (L.det_only >= 0 ? L.det_only : 0);
        if (!have_vectors) fold_vectors<MODE>(T, D, p, b, scr, O);
        // disabled channels still produce arrays (propagation_reco.jl:12-15,119-122)
        if (MODE == kModeRates || MODE == kModePerbin)
            for (int ch = 0; ch < 2; ++ch) {
                if (ch == 0 ? D.es_mode : D.cc_mode) continue;
                const int nR = ch == 0 ? D.n_es : D.n_cc;
                double* day = ch == 0 ? O.es_day : O.cc_day;
                double* night = ch == 0 ? O.es_night : O.cc_night;
                const double v = (MODE == kModePerbin && oob) ? -INFINITY : 0.0;
                for (int o = threadIdx.x; o < nrow * nR; o += blockDim.x) {
                    if (o < nR) { if (day) day[b * nR + o] = v; }
                    else if (night) night[b * nZ * nR + (o - nR)] = v;
                }
            }
        SOLAROSC_PH(6);
        if (!have_vectors) __syncthreads();
        SOLAROSC_PH(7);
        double part = 0.0;
        double d_es = 0.0, n_es = 0.0, d_cc = 0.0, n_cc = 0.0;   // asymmetry mode: day / night totals above threshold
        // likelihood and asymmetry only look at the bins from the threshold on (likelihood_builder.jl:5-20,129-144):
        // they use the copies of the fold matrices whose columns start at the threshold bin, so no column block is
        // spent on bins that never enter the sum (24 bins = 3 blocks instead of 4 per channel in the default binning)
        constexpr bool kThr = MODE == kModeLlh || MODE == kModeAsym;
        const int bte = kThr ? D.nte_thr : nte, btc = kThr ? D.ntc_thr : ntc;
        FoldTask<MODE> F{T, L, D, O, p8s, phs, scale, cst_es, cst_cc, f8, fh, b, ldp, nrow, mt, kt, bte, btc, oob};
        // Column block (8 reco bins of one channel) x row tiles.  A point that runs its own oscillation stage shares
        // the SM with another CTA that is mostly propagating, and one column block per warp with all row tiles in
        // flight (most independent logarithm chains) is the fastest arrangement even though it leaves warps without
        // a block idle.  When the points of a group share one oscillation stage the kernel is fold-bound: there the
        // work is cut finer -- per column block, groups of kMCa row tiles, then the remainder in groups of kMCb --
        // and dealt by the host's longest-first assignment (D.fold_owner) so that all kWarps warps carry the same
        // number of tiles (6 column blocks x 7 row tiles in the default binning: six warps take 5 tiles, two take 3 x 2).
        const int nb = bte + btc;
        const int na = mt / kMCa, nbm = (mt - kMCa * na + kMCb - 1) / kMCb;
        const int ntask = nb * (na + nbm);
        if (FINE) {   // (a kernel of its own, llh_kernel<MODE, true>: the two arrangements do not share register allocation)
            for (int t = 0; t < ntask; ++t) {
                if ((int)D.fold_owner[t] != warp) continue;
                const bool big = t < nb * na;
                const int tt = big ? t : t - nb * na;
                const int pass = tt / nb, col = tt - pass * nb;
                if (col < bte ? !D.es_mode : !D.cc_mode) continue;
                if (big) F.template run<kMCa>(col, kMCa * pass, part, fl, d_es, n_es, d_cc, n_cc);
                else F.template run<kMCb>(col, kMCa * na + kMCb * pass, part, fl, d_es, n_es, d_cc, n_cc);
            }
        } else {
            for (int col = warp; col < nb; col += kWarps) {
                if (col < bte ? !D.es_mode : !D.cc_mode) continue;
                for (int m0 = 0; m0 < mt; m0 += kMC) F.template run<kMC>(col, m0, part, fl, d_es, n_es, d_cc, n_cc);
            }
        }
        if (MODE == kModeAsym) {
            // derive_variables_from_chain.jl:236-258: half of the background above threshold is taken off
            // the day total and half off the night total; a zero denominator gives 0
            for (int ch = 0; ch < 2; ++ch) {
                const bool is_es = ch == 0;
                double* out = is_es ? O.asym_es : O.asym_cc;
                if (is_es ? !D.es_mode : !D.cc_mode) {
                    if (out && threadIdx.x == 0) out[b] = 0.0;
                    continue;
                }
                const double dtot = block_sum(is_es ? d_es : d_cc, red), ntot = block_sum(is_es ? n_es : n_cc, red);
                if (threadIdx.x == 0) {
                    const double* bgs = is_es ? bgs_es : bgs_cc;
                    const int nR = is_es ? D.n_es : D.n_cc, idx0 = (is_es ? D.index_es : D.index_cc) - 1;
                    double bg = 0.0;
                    for (int j = idx0; j < nR; ++j) bg += bgs[j];
                    const double dd = dtot - 0.5 * bg, nn = ntot - 0.5 * bg, den = dd + nn;
                    if (out) out[b] = (den == 0.0) ? 0.0 : 2.0 * (dd - nn) / den;
                }
            }
        }
        total += part;
    }
    SOLAROSC_PH(8);
    if (MODE != kModeLlh) return;
    // ---- block reduction (fixed order: deterministic) ----------------------------------------------
    total = warp_sum(total);
    for (int o = 16; o > 0; o >>= 1) fl |= __shfl_xor_sync(0xffffffffu, fl, o);
    if (lane == 0) { red[warp] = total; redf[warp] = fl; }
    __syncthreads();
    SOLAROSC_PH(9);
    if (threadIdx.x == 0) {
        double s = 0.0;
        uint32_t f = 0;
        for (int w = 0; w < kWarps; ++w) { s += red[w]; f |= redf[w]; }
        double lnl = -s;
        if (f & SOLAROSC_FLAG_NEG_ERROR) lnl = NAN;
        if (oob) { lnl = -INFINITY; f = SOLAROSC_FLAG_EARTH_OOB; }
        else if (lnl != lnl) f |= (f & SOLAROSC_FLAG_NEG_ERROR) ? 0u : SOLAROSC_FLAG_NONFINITE;
        O.lnl[b] = lnl;
        if (O.flags) O.flags[b] = f;
    }
}

// Zero the K padding of the probability tile (columns n_etrue .. 4*ceil(n_etrue/4) - 1 of both processes): the fold
// multiplies them by the zero rows of the constant matrices, so they must be finite.  Padding ROWS need nothing:
// an output row of an MMA only depends on the same row of A, and the epilogue drops rows >= nrow.
__device__ __forceinline__ void fold_zero_padding(double* ps, int nT, int rows_alloc, int ldp) {
    const int padc = 4 * fold_kt(nT) - nT;   // rows_alloc: rows per process as laid out by the caller
    if (padc == 0) return;
    for (int t = threadIdx.x; t < 2 * rows_alloc * padc; t += blockDim.x) {
        const int r = t / padc, c = t - r * padc;
        ps[(size_t)r * ldp + nT + c] = 0.0;
    }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, SOLAROSC_RATES_MINBLOCKS)
rates_kernel(const DevTables T, const RatesLaunch L, const double* __restrict__ params,
             const double* __restrict__ osc /* [n_points/group][2][1+n_cosz][n_etrue] */, RatesOut O) {
    extern __shared__ __align__(16) double smem[];
    const int nT = T.n_etrue, nrow = 1 + T.n_cosz, ldp = fold_ldp(nT), rows = fold_rows(nrow);
    double* p8s = smem;
    double* phs = p8s + (size_t)rows * ldp;
    double* scr = p8s + fold_p_doubles(nT, nrow);
    const size_t b = blockIdx.x;
    const double* src = osc + (b / L.group) * 2 * nrow * nT;
    fold_zero_padding(p8s, nT, rows, ldp);
    {   // stage the probabilities: (row, i) advance by blockDim.x without a division in the loop
        const int dq = (int)blockDim.x / nT, dr = (int)blockDim.x - dq * nT;
        int row = (int)threadIdx.x / nT, i = (int)threadIdx.x - row * nT;
        while (row < nrow) {
            p8s[row * ldp + i] = src[row * nT + i];
            phs[row * ldp + i] = src[(nrow + row) * nT + i];
            i += dr; row += dq;
            if (i >= nT) { i -= nT; ++row; }
        }
    }
    fold_stage<MODE>(T, L, params + b * L.param_stride, b, p8s, phs, ldp, scr, O);
}

// ------------------------------------------------------------------------------------------
// K3 + K4 + K5 for binnings whose probability tile (2 x (1 + n_cosz) x n_etrue doubles) does not fit one SM's shared
// memory -- the dense zenith x energy grids of utils/checkOscGridConvergence.jl:36-46 (200 x 200, 1000 x 1000).  An
// output row of the fold depends on the same row of the probabilities only, so the rows are cut into chunks of
// TL.rows (a multiple of 8): CTA (point, chunk) stages its rows of both processes with the FULL energy axis and runs
// the same FoldTask on them (row0 = first row of the chunk).  Likelihood partial sums land in part[point][chunk] and
// tiled_finish_kernel adds the chunks in order (deterministic).  The oscillation stage dominates these grids by two
// orders of magnitude, so this kernel is sized for generality, not speed.
// ------------------------------------------------------------------------------------------
struct TiledLaunch {
    int rows;          // rows per chunk
    int n_chunks;      // chunks per point
    double* part;      // [n_points][n_chunks] likelihood partial sums
    uint32_t* part_flags;
};
__host__ __device__ inline size_t tiled_smem_bytes(int n_etrue, int rows, int max_ereco) {
    return sizeof(double) * ((size_t)2 * rows * fold_ldp(n_etrue) + fold_scratch_doubles(n_etrue, max_ereco));
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
rates_tiled_kernel(const DevTables T, const RatesLaunch L, const TiledLaunch TL, const double* __restrict__ params,
                   const double* __restrict__ osc /* [n_points/group][2][1+n_cosz][n_etrue] */, RatesOut O) {
    extern __shared__ __align__(16) double smem[];
    const int nT = T.n_etrue, nZ = T.n_cosz, nrow = 1 + nZ, ldp = fold_ldp(nT), kt = fold_kt(nT);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t b = blockIdx.x / TL.n_chunks;
    const int rc = (int)(blockIdx.x - b * TL.n_chunks), r0 = rc * TL.rows, nr = min(TL.rows, nrow - r0);
    double* p8s = smem;
    double* phs = p8s + (size_t)TL.rows * ldp;
    double* scr = phs + (size_t)TL.rows * ldp;
    double* cst_es = scr + 8 * kt;
    double* red = scr + fold_vec_doubles(nT, L.max_ereco);
    uint32_t* redf = reinterpret_cast<uint32_t*>(red + kWarps);
    const double* p = params + b * L.param_stride;
    const double* src = osc + (b / L.group) * 2 * nrow * nT;
    fold_zero_padding(p8s, nT, TL.rows, ldp);
    for (int t = threadIdx.x; t < nr * nT; t += blockDim.x) {
        const int row = t / nT, i = t - row * nT;
        p8s[row * ldp + i] = src[(size_t)(r0 + row) * nT + i];
        phs[row * ldp + i] = src[(size_t)(nrow + r0 + row) * nT + i];
    }
    bool oob = false;   // earth-norm bounds (likelihood_debug.jl:1-4)
    if (T.earth_unc)
        for (int k = 0; k < T.n_layers; ++k) {
            const double x = p[SOLAROSC_N_FIXED_PARAMS + k];
            oob |= !(x >= 0.0 && x <= 2.0);
        }
    RatesOut Ob = O;
    if (rc != 0) { Ob.bg_es = nullptr; Ob.bg_cc = nullptr; }   // the background vectors are written by the first chunk
    double total = 0.0;
    uint32_t fl = 0;
    for (int di = 0; di < T.n_det; ++di) {
        if (L.det_only >= 0 && di != L.det_only) continue;
        const DevDetector& D = T.det[di];
        const int nte = (D.n_es + 7) / 8, ntc = (D.n_cc + 7) / 8;
        double* cst_cc = cst_es + 8 * nte;
        __syncthreads();
        fold_vectors<MODE>(T, D, p, b, scr, Ob);
        if (MODE == kModeRates || MODE == kModePerbin)   // disabled channels still produce arrays (propagation_reco.jl:12-15,119-122)
            for (int ch = 0; ch < 2; ++ch) {
                if (ch == 0 ? D.es_mode : D.cc_mode) continue;
                const int nR = ch == 0 ? D.n_es : D.n_cc;
                double* day = ch == 0 ? O.es_day : O.cc_day;
                double* night = ch == 0 ? O.es_night : O.cc_night;
                const long long sr = ch == 0 ? L.s_row_es : L.s_row_cc, sj = ch == 0 ? L.s_j_es : L.s_j_cc;
                const double v = (MODE == kModePerbin && oob) ? -INFINITY : 0.0;
                for (int o = threadIdx.x; o < nr * nR; o += blockDim.x) {
                    const int row = r0 + o / nR, j = o % nR;
                    if (row == 0) { if (day) day[b * nR + j] = v; }
                    else if (night) night[b * nZ * nR + (size_t)(row - 1) * sr + (size_t)j * sj] = v;
                }
            }
        __syncthreads();
        constexpr bool kThr = MODE == kModeLlh || MODE == kModeAsym;
        const int bte = kThr ? D.nte_thr : nte, btc = kThr ? D.ntc_thr : ntc;
        FoldTask<MODE> F{T, L, D, O, p8s, phs, scr, cst_es, cst_cc, p[3], p[4], b, ldp, nr, (nr + 7) / 8, kt, bte, btc, oob};
        F.row0 = r0;
        double part = 0.0, d_es = 0.0, n_es = 0.0, d_cc = 0.0, n_cc = 0.0;
        const int nb = bte + btc, mt = (nr + 7) / 8;
        // work items (column block, group of kMCb row tiles) dealt round-robin over the warps
        const int ng = (mt + kMCb - 1) / kMCb;
        for (int it = warp; it < nb * ng; it += kWarps) {
            const int col = it % nb, m0 = kMCb * (it / nb);
            if (col < bte ? !D.es_mode : !D.cc_mode) continue;
            F.template run<kMCb>(col, m0, part, fl, d_es, n_es, d_cc, n_cc);
        }
        total += part;
    }
    if (MODE != kModeLlh) return;
    total = warp_sum(total);
    for (int o = 16; o > 0; o >>= 1) fl |= __shfl_xor_sync(0xffffffffu, fl, o);
    __syncthreads();
    if (lane == 0) { red[warp] = total; redf[warp] = fl; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        uint32_t f = 0;
        for (int w = 0; w < kWarps; ++w) { s += red[w]; f |= redf[w]; }
        TL.part[blockIdx.x] = s;
        TL.part_flags[blockIdx.x] = f | (oob ? 0x80000000u : 0u);
    }
}

__global__ void tiled_finish_kernel(const TiledLaunch TL, int n_points, double* __restrict__ lnl_out, uint32_t* __restrict__ flags) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_points) return;
    double s = 0.0;
    uint32_t f = 0;
    for (int k = 0; k < TL.n_chunks; ++k) { s += TL.part[(size_t)b * TL.n_chunks + k]; f |= TL.part_flags[(size_t)b * TL.n_chunks + k]; }
    const bool oob = (f & 0x80000000u) != 0;
    f &= 0x7fffffffu;
    double lnl = -s;
    if (f & SOLAROSC_FLAG_NEG_ERROR) lnl = NAN;
    if (oob) { lnl = -INFINITY; f = SOLAROSC_FLAG_EARTH_OOB; }
    else if (lnl != lnl) f |= (f & SOLAROSC_FLAG_NEG_ERROR) ? 0u : SOLAROSC_FLAG_NONFINITE;
    lnl_out[b] = lnl;
    if (flags) flags[b] = f;
}

// Persistent fused kernel: every CTA walks oscillation points g = blockIdx.x, blockIdx.x + gridDim.x, ...  Per
// point: K1+K2 into the shared-memory probability tile, one energy chunk (L.osc_cb coarse bins) at a time so
// that the eigen table stays small enough for several CTAs per SM, then K3+K4+K5 for each of the `group`
// parameter points sharing it.  The oscillation probabilities never leave the SM.
template <int MODE, bool FINE>
__global__ void __launch_bounds__(kThreads, SOLAROSC_LLH_MINBLOCKS)
llh_kernel(const DevTables T, const RatesLaunch L, const double* __restrict__ params, RatesOut O) {
    extern __shared__ __align__(16) double smem[];
    const int nT = T.n_etrue, nrow = 1 + T.n_cosz, ldp = fold_ldp(nT);
    SOLAROSC_PH(0);
    const int cbs = L.osc_cb;
    double* tbl = smem;
    double* msw = smem + (size_t)kEntry * T.n_ltbl * 2 * cbs;
    double* p8s = smem + osc_smem_doubles(T.n_ltbl, cbs);   // [nrow][ldp] per process, no padding rows (llh_smem_bytes)
    double* phs = p8s + (size_t)nrow * ldp;
    double* scr = phs + (size_t)nrow * ldp;
    const SegTables G = stage_seg_tables(T, scr + fold_scratch_doubles(nT, L.max_ereco), true);   // the host only launches this kernel when they fit
    SOLAROSC_PH(10);
    fold_zero_padding(p8s, nT, nrow, ldp);
    SOLAROSC_PH(11);
    const int n_osc = L.n_points / L.group;
    for (int g = blockIdx.x; g < n_osc; g += gridDim.x) {
        const double* po = params + (size_t)g * L.group * L.param_stride;
        // point g*group's fold vectors ride along with the first eigen table.  The likelihood fold ends with a barrier
        // behind its last read of the scratch; the asymmetry fold does not (thread 0 still sums the backgrounds), so a
        // persistent grid (SOLAROSC_LLH_GRID) needs one here before the scratch is rewritten.
        if (MODE != kModeLlh && g != (int)blockIdx.x) __syncthreads();
        const int d0 = L.det_only >= 0 ? L.det_only : 0;
        fold_vectors<MODE>(T, T.det[d0], po, (size_t)g * L.group, scr, O);
        SOLAROSC_PH(12);
        for (int i0 = 0; i0 < nT; i0 += cbs) {
            // the table of the previous chunk is still being read by slower warps; the previous point's fold is
            // done with the probability tile once its closing barrier has been passed
            if (i0 > 0) __syncthreads();
            osc_stage<true>(T, G, po, i0, min(cbs, nT - i0), cbs, tbl, msw, nullptr, SmemSink{p8s, phs, ldp});
        }
        for (int q = 0; q < L.group; ++q) {
            const size_t b = (size_t)g * L.group + q;
            // fold_stage begins with a barrier, which also orders the probability tile after osc_stage
            fold_stage<MODE, FINE>(T, L, params + b * L.param_stride, b, p8s, phs, ldp, scr, O, q == 0);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Derived-mode reductions over a block of chain samples (reference derive_variables_from_chain.jl:260-444): the
// replay loop's weighted moments accumulated on the device, so that a block of samples returns O(bins) numbers
// instead of 49 kB of rate arrays per sample.  Input: the rate arrays of the block as rates_kernel<kModeRates>
// wrote them (row-major night arrays) in device scratch.
//   element e of channel c: row = e / nR (0 = day, 1 + z = night zenith bin z), j = e % nR
//   l(e)  = perbin Poisson term (likelihood_statistical.jl:31-57; bins idx..n-1, the overflow bin excluded)
//           minus the caller's reference l_ref(e)                       (derive...:357-361)
//   sums  = [ sum w r, sum w r^2 (bins idx..n: derive...:285-313),  sum w l, sum w l^2, sum w l th1, sum w l th2 (:420-444) ]
//   bg    = [ sum w bg/2, sum w (bg/2)^2 ]                              (derive...:322-339)
// HBM-bound streaming reduction: every rate is read once per kernel.
// ------------------------------------------------------------------------------------------
struct DerivedLaunch {
    int n_samples, n_slices, param_stride, det;
    int nZ, nE, nC, idx_es, idx_cc;      // idx_*: 0-based first bin of the likelihood
    int first;                           // 1: partial sums start from zero, 0: they continue an earlier block
    int earth_unc, n_layers;
    int nCos;                            // cos-scatter bins of the ES arrays (1: no angular reconstruction)
};
struct DerivedIn {
    const double *es_day, *es_night, *cc_day, *cc_night, *bg_es, *bg_cc;   // [n][..] device scratch
    const double *obs_es, *obs_cc;       // [1+nZ][nR]
    const double *lref_es, *lref_cc;     // [1+nZ][nR] or nullptr
    const double *params, *weights;
    const double2* obs_es_ang;           // angular ES data (m, log m), [1 + nZ][nE][nCos]
    double* es_day_sum;                  // angular: [n][nE] day rates summed over the cos-scatter bins (derive...:315-318)
    double* lbuf;                        // [n][nel_es + nel_cc] centred per-bin terms: written by derived_sample_kernel, read by
                                         // derived_moment_kernel (one logarithm per element and sample instead of two)
};
constexpr int kDerivedQ = 6;

__device__ __forceinline__ bool derived_oob(const DerivedLaunch& L, const double* p) {
    bool oob = false;
    if (L.earth_unc)
        for (int k = 0; k < L.n_layers; ++k) {
            const double x = p[SOLAROSC_N_FIXED_PARAMS + k];
            oob |= !(x >= 0.0 && x <= 2.0);
        }
    return oob;
}
// rate and centred per-bin term of element e of a channel for sample b
__device__ __forceinline__ void derived_element(int e, int nR, int nZ, int idx0, size_t b, const double* day, const double* night,
                                                const double* obs, const double* lref, bool oob, double& r, double& l, bool& above,
                                                const double* lpre = nullptr) {
    const int row = e / nR, j = e - row * nR;
    r = row == 0 ? day[b * nR + j] : night[(b * nZ + (row - 1)) * nR + j];
    above = j >= idx0;
    if (lpre) { l = *lpre; return; }
    double v = (j >= idx0 && j < nR - 1) ? perbin_term(r, obs[e]) : 0.0;
    if (oob) v = -INFINITY;              // make_perbin_likelihood, likelihood_builder.jl:271-282
    l = v - (lref ? lref[e] : 0.0);
}

// The same for element e of the angular ES arrays, laid out [day: nCos][nE] then [night: nCos][nE][nZ] (the reference's index
// order, row-major): llh_ES_angle_perbin, likelihood_builder.jl:80-108; moments from the threshold bin on (derive...:288-296)
__device__ __forceinline__ void derived_element_ang(int e, int nE, int nZ, int nCos, int idx0, size_t b, const double* day,
                                                    const double* night, const double2* obs, const double* lref, bool oob,
                                                    double& r, double& l, bool& above, bool& is_day, const double* lpre = nullptr) {
    const int per_d = nCos * nE;
    int c, j, row;
    is_day = e < per_d;
    if (is_day) { c = e / nE; j = e - c * nE; row = 0; r = day[b * per_d + e]; }
    else {
        const int x = e - per_d, cj = x / nZ;
        c = cj / nE; j = cj - c * nE; row = 1 + (x - cj * nZ);
        r = night[b * (size_t)per_d * nZ + x];
    }
    above = j >= idx0;
    if (lpre) { l = *lpre; return; }
    double v = (j >= idx0 && j < nE - 1) ? perbin_term(r, obs[((size_t)row * nE + j) * nCos + c].x) : 0.0;
    if (oob) v = -INFINITY;
    l = v - (lref ? lref[e] : 0.0);
}

// One CTA per sample: totals of the centred per-bin terms per (channel, day/night): out[b][4] = ES night, ES day,
// CC night, CC day (derive...:388-417).
__global__ void __launch_bounds__(256) derived_sample_kernel(const DerivedLaunch L, const DerivedIn I, double* __restrict__ ltot) {
    const size_t b = blockIdx.x;
    const bool oob = derived_oob(L, I.params + b * L.param_stride);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const int nel_es = (1 + L.nZ) * L.nE * L.nCos, nel_cc = (1 + L.nZ) * L.nC;
    for (int t = threadIdx.x; t < nel_es + nel_cc; t += blockDim.x) {
        const bool es = t < nel_es;
        const int e = es ? t : t - nel_es, nR = es ? L.nE : L.nC;
        double r, l;
        bool above, is_day;
        if (es && L.nCos > 1) derived_element_ang(e, L.nE, L.nZ, L.nCos, L.idx_es, b, I.es_day, I.es_night, I.obs_es_ang, I.lref_es, oob, r, l, above, is_day);
        else {
            derived_element(e, nR, L.nZ, es ? L.idx_es : L.idx_cc, b, es ? I.es_day : I.cc_day, es ? I.es_night : I.cc_night,
                            es ? I.obs_es : I.obs_cc, es ? I.lref_es : I.lref_cc, oob, r, l, above);
            is_day = e < nR;
        }
        acc[(es ? 0 : 2) + (is_day ? 1 : 0)] += l;
        if (I.lbuf) I.lbuf[b * (size_t)(nel_es + nel_cc) + t] = l;
    }
    if (L.nCos > 1 && I.es_day_sum)   // the day spectrum per sample: sum over the cos-scatter bins, in bin order
        for (int j = threadIdx.x; j < L.nE; j += blockDim.x) {
            double sj = 0.0;
            for (int c = 0; c < L.nCos; ++c) sj += I.es_day[b * (size_t)L.nCos * L.nE + c * L.nE + j];
            I.es_day_sum[b * L.nE + j] = sj;
        }
    // the four totals in one pass: warp sums to shared memory, one barrier, warps added in order (deterministic)
    __shared__ double red4[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double s = warp_sum(acc[k]);
        if ((threadIdx.x & 31) == 0) red4[k][threadIdx.x >> 5] = s;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red4[threadIdx.x][w];
        ltot[b * 4 + threadIdx.x] = s;
    }
}

// grid (elements / 256, n_slices): thread = one element, slice s walks samples s, s + n_slices, ... in order and adds
// into partial[s][kDerivedQ][n_elem]; derived_finish_kernel adds the slices in order.  Deterministic.
__global__ void __launch_bounds__(256) derived_moment_kernel(const DerivedLaunch L, const DerivedIn I, double* __restrict__ partial) {
    const int nel_es = (1 + L.nZ) * L.nE * L.nCos, nel_cc = (1 + L.nZ) * L.nC, n_elem = nel_es + nel_cc + L.nE + L.nC;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_elem) return;
    double q[kDerivedQ];
    double* out = partial + (size_t)blockIdx.y * kDerivedQ * n_elem + t;
#pragma unroll
    for (int k = 0; k < kDerivedQ; ++k) q[k] = L.first ? 0.0 : out[(size_t)k * n_elem];
    for (size_t b = blockIdx.y; b < (size_t)L.n_samples; b += L.n_slices) {
        const double w = I.weights[b];
        const double* p = I.params + b * L.param_stride;
        if (t < nel_es + nel_cc) {
            const bool es = t < nel_es;
            const int e = es ? t : t - nel_es, nR = es ? L.nE : L.nC;
            double r, l;
            bool above, is_day;
            const double* lpre = I.lbuf ? I.lbuf + b * (size_t)(nel_es + nel_cc) + t : nullptr;
            const bool oob = lpre ? false : derived_oob(L, p);
            if (es && L.nCos > 1) derived_element_ang(e, L.nE, L.nZ, L.nCos, L.idx_es, b, I.es_day, I.es_night, I.obs_es_ang, I.lref_es,
                                                      oob, r, l, above, is_day, lpre);
            else derived_element(e, nR, L.nZ, es ? L.idx_es : L.idx_cc, b, es ? I.es_day : I.cc_day, es ? I.es_night : I.cc_night,
                                 es ? I.obs_es : I.obs_cc, es ? I.lref_es : I.lref_cc, oob, r, l, above, lpre);
            if (above) { q[0] += w * r; q[1] += w * (r * r); }
            q[2] += w * l; q[3] += w * (l * l); q[4] += w * (l * p[0]); q[5] += w * (l * p[2]);
        } else {
            const int e = t - nel_es - nel_cc;
            const double h = 0.5 * (e < L.nE ? I.bg_es[b * L.nE + e] : I.bg_cc[b * L.nC + (e - L.nE)]);
            q[0] += w * h; q[1] += w * (h * h);
        }
    }
#pragma unroll
    for (int k = 0; k < kDerivedQ; ++k) out[(size_t)k * n_elem] = q[k];
}

__global__ void __launch_bounds__(256) derived_finish_kernel(int n_slices, int n_elem, const double* __restrict__ partial,
                                                             double* __restrict__ out /* [kDerivedQ][n_elem] */) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= kDerivedQ * n_elem) return;
    double s = 0.0;
    for (int k = 0; k < n_slices; ++k) s += partial[(size_t)k * kDerivedQ * n_elem + t];
    out[t] = s;
}

// ------------------------------------------------------------------------------------------
// Stand-alone Poisson reduction: one CTA per (expected, measured) vector pair
// (reference poissonLogLikelihood, likelihood_statistical.jl:1-28; the shape timed by
//  tests/benchmark_likelihoods.jl:10-19)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
poisson_kernel(const double* __restrict__ e, const double* __restrict__ m, size_t len, int n_split, double* __restrict__ part,
               uint32_t* __restrict__ part_flags) {
    // CTA (v, s) sums segment s of vector v (a handful of long vectors would otherwise leave most SMs idle);
    // poisson_finish_kernel adds the segments in order
    __shared__ double red[8];
    __shared__ uint32_t redf[8];
    const size_t v = blockIdx.x / n_split, sgm = blockIdx.x - v * n_split;
    const size_t per = (len + n_split - 1) / n_split, i0 = sgm * per, i1 = i0 + per < len ? i0 + per : len;
    const double* ev = e + v * len;
    const double* mv = m + v * len;
    double acc = 0.0;
    uint32_t fl = 0;
    for (size_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const double mm = mv[i];
        acc += poisson_term(ev[i], mm, mm > 0.0 ? log(mm) : 0.0, fl);
    }
    acc = warp_sum(acc);
    for (int o = 16; o > 0; o >>= 1) fl |= __shfl_xor_sync(0xffffffffu, fl, o);
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = acc; redf[threadIdx.x >> 5] = fl; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        uint32_t f = 0;
        for (int w = 0; w < 8; ++w) { t += red[w]; f |= redf[w]; }
        part[blockIdx.x] = t;
        part_flags[blockIdx.x] = f;
    }
}

__global__ void poisson_finish_kernel(const double* __restrict__ part, const uint32_t* __restrict__ part_flags, size_t n_vec,
                                      int n_split, double* __restrict__ out, uint32_t* __restrict__ flags) {
    const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vec) return;
    double t = 0.0;
    uint32_t f = 0;
    for (int k = 0; k < n_split; ++k) { t += part[v * n_split + k]; f |= part_flags[v * n_split + k]; }
    out[v] = (f & SOLAROSC_FLAG_NEG_ERROR) ? NAN : -t;
    if (flags) flags[v] = f;
}

// ------------------------------------------------------------------------------------------
// FP64 FMA peak probe: 16 independent dependent-FMA chains per thread, nothing else in the loop.
// This is the denominator of the FP64 roofline (MEASURED_PEAKS.json has no FP64 figure).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double m) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 1.0 + 1e-3 * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, 1e-9);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 12345.678) out[threadIdx.x] = s;   // never true: keeps the chains alive
}

}  // namespace solarosc
