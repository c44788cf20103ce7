// solarosc_pipe.cuh -- the fused likelihood kernel as a persistent, software-pipelined dataflow.
//
// llh_kernel (solarosc_kernels.cuh) walks the phases of one oscillation point in lock step -- eigen table |
// barrier | propagation | barrier | fold + Poisson | barrier -- and ncu shows where that loses: the propagation
// loop alone keeps the FP64 pipe busy, the eigen table and the Poisson pass are latency-bound chains that cannot,
// and a CTA is in exactly one of them at a time (profiles/r01_llh_kernel_ncu_summary.txt: FP64 pipe 52 % busy).
//
// pipe_kernel runs ONE CTA of 16 warps per SM for the whole launch and cuts the work of every oscillation point
// into warp-sized items of three kinds,
//     E(n)  eigen-table bundles (32 entries each), the MSW item and the point's mixing constants
//     P(n)  propagation bundles (32 (zenith row, coarse energy) items, longest paths first) and the day rows
//     F(n)  fold items: one 8-column block of one channel x a few row tiles, DMMA + Poisson epilogue, for each
//           of the `group` parameter points that share the oscillation stage
// which the warps draw from ONE ticket counter.  Tickets are laid out so that the block of tickets of step n mixes
// P(n) with E(n+1) and F(n-1): at any moment the 16 warps of an SM hold pipe-saturating propagation bundles next to
// latency-bound eigen and Poisson chains of the neighbouring points.  Eigen table, MSW vectors and the probability
// tile are double-buffered in shared memory (~150 kB), so there is no CTA-wide barrier anywhere: an item waits on a
// shared-memory completion counter of the items it depends on,
//     P(n) <- E(n) complete, F(n-2) complete (tile buffer free)
//     E(n) <- P(n-2) complete (table buffer free)          F(n) <- P(n) complete
// Every dependency points at an earlier ticket, and at most 16 tickets are in flight, so the oldest outstanding item
// can always run: no deadlock.  The partial sums of a point's fold items land in fixed slots and the warp that
// completes the point adds them in slot order: lnL is deterministic whatever the interleaving.
//
// Same arithmetic as llh_kernel: the step functions, make_entry and FoldTask are shared.
#pragma once
#include "solarosc_kernels.cuh"

namespace solarosc {

constexpr int kPipeThreads = 512;
constexpr int kPipeWarps = kPipeThreads / 32;
#ifndef SOLAROSC_PIPE_NT
#define SOLAROSC_PIPE_NT 7
#endif
#ifndef SOLAROSC_PIPE_EPL
#define SOLAROSC_PIPE_EPL 2
#endif
constexpr int kPipeNT = SOLAROSC_PIPE_NT;    // row tiles a fold item keeps in flight; it walks its tiles in steps of kPipeNT with ONE copy of the code
constexpr int kPipeEPL = SOLAROSC_PIPE_EPL;  // eigen entries per lane of an E item (interleaved dependency chains)
constexpr int kPipeMaxGroup = 64;

enum PipeKind { kPipeP = 0, kPipeE = 1, kPipeF = 2 };

// one fold item: column block `col` of detector `det` (ES blocks first), row tiles m0 .. m0 + nt - 1
struct PipeFoldItem {
    unsigned char det, col, m0, nt;
};

struct PipeLaunch {
    int n_osc;            // oscillation points of this launch
    int group;            // parameter points per oscillation point
    int param_stride;
    int det_only;
    int block_len;        // tickets per block
    int n_p, n_e, n_f;    // items per oscillation point (n_f: per parameter point)
    int n_eig;            // E items 0 .. n_eig-1 build eigen entries, the others are MSW items
    int msw_per_item;     // fine energies per MSW item (slow mode), 0 in fast mode (one item does all)
    int max_ereco;
    const unsigned short* pattern;   // [block_len] kind << 14 | index
    const PipeFoldItem* fitems;      // [n_f]
    long long s_row_es, s_j_es, s_row_cc, s_j_cc;
};

// ---- shared-memory layout (offsets in doubles) ------------------------------------------------------------
struct PipeSmem {
    size_t tbl, msw, cst, ptile, wscr, red, seg, cnt, total;
    size_t tbl_sz, msw_sz, ptile_sz, wscr_sz, red_sz;
};
constexpr int kPipeCst = 8;        // u1, u2, u3, P0 (+ padding) per buffer
constexpr int kPipeCounters = 8;   // ticket, E[2], P[2], F[2] (+ padding), then 2 * group per-point counters
__host__ __device__ inline PipeSmem pipe_smem(int n_ltbl, int nT, int n_cosz, int group, int n_f, int red_per_item, size_t seg_doubles) {
    PipeSmem s{};
    const int nrow = 1 + n_cosz, ldp = fold_ldp(nT), kt = fold_kt(nT);
    s.tbl_sz = (size_t)kEntry * n_ltbl * 2 * nT;
    s.msw_sz = (size_t)8 * nT;
    s.ptile_sz = (size_t)2 * nrow * ldp;
    s.wscr_sz = (size_t)8 * kt + 8;                                       // scale vectors [2][4 kt] + 8 additive constants
    s.red_sz = (size_t)group * n_f * (red_per_item + 1);                  // partial sums + a flag word each
    size_t o = 0;
    s.tbl = o; o += 2 * s.tbl_sz;
    s.msw = o; o += 2 * s.msw_sz;
    s.cst = o; o += 2 * kPipeCst;
    s.ptile = o; o += 2 * s.ptile_sz;
    // the last 8-row MMA tile of the second process reads up to 7 rows past the tile (results dropped): what follows
    // is at least that long
    const size_t over = (size_t)(fold_rows(nrow) - nrow) * ldp;
    s.wscr = o; o += (size_t)kPipeWarps * s.wscr_sz;
    s.red = o; o += 2 * s.red_sz;
    s.seg = (o + 1) & ~(size_t)1; o = s.seg + seg_doubles;
    s.cnt = o; o += (kPipeCounters + 2 * (size_t)group + 1) / 2 + 1;
    if (o - (s.ptile + 2 * s.ptile_sz) < over) o = s.ptile + 2 * s.ptile_sz + over;
    s.total = o;
    return s;
}

// Diagnostic builds (-DSOLAROSC_PIPE_STATS): cycles every warp spent waiting at each dependency and inside the items of
// each kind, summed over the launch (solarosc_debug_read_pipe_stats; tools/pipe_stats.py).
#ifdef SOLAROSC_PIPE_STATS
__device__ unsigned long long g_pipe_stats[16];
#define SOLAROSC_PS_DECL long long ps_t0 = 0; (void)ps_t0
#define SOLAROSC_PS_BEGIN() ps_t0 = clock64()
#define SOLAROSC_PS_END(i)                                                                         \
    do {                                                                                           \
        if ((threadIdx.x & 31) == 0) atomicAdd(&g_pipe_stats[i], (unsigned long long)(clock64() - ps_t0)); \
    } while (0)
#else
#define SOLAROSC_PS_DECL
#define SOLAROSC_PS_BEGIN()
#define SOLAROSC_PS_END(i)
#endif

// ---- completion counters ------------------------------------------------------------------------------------
// All lanes poll together (one broadcast load, a warp-uniform exit decided by a vote): a poll loop run by lane 0 alone
// left the warp split into two convergence groups for the rest of the item on this architecture -- __syncwarp
// synchronises, it does not reconverge -- and every instruction of the item was then issued twice (ncu: the
// propagation loop ran 1.5 x its trip count).
__device__ __forceinline__ void pipe_wait(const unsigned* cnt, unsigned target) {
    if (target == 0u) return;
    const uint32_t a = smem_addr(cnt);
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
        if (__all_sync(0xffffffffu, v >= target)) break;
        __nanosleep(64);
    }
}
__device__ __forceinline__ void pipe_signal(unsigned* cnt) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) asm volatile("red.release.cta.shared.add.u32 [%0], 1;" ::"r"(smem_addr(cnt)) : "memory");
}

// mixing constants of a point: u = (c12 c13, s12 c13, s13), P0, cos 2 theta12, c13^2, s13^2
struct PipeMix {
    double u1, u2, u3, p0, c2, c13sq, s13sq, dm21;
};
__device__ __forceinline__ PipeMix pipe_mix(const double* __restrict__ p) {
    auto clamp01 = [](double x) { return x != x ? x : fmin(fmax(x, 0.0), 1.0); };   // lets a NaN through like Julia's clamp
    PipeMix m;
    const double s12sq = clamp01(p[0]);
    m.s13sq = clamp01(p[1]);
    m.dm21 = p[2];
    m.c13sq = 1.0 - m.s13sq;
    const double c12sq = 1.0 - s12sq, c13 = sqrt01(m.c13sq);
    m.u1 = sqrt01(c12sq) * c13; m.u2 = sqrt01(s12sq) * c13; m.u3 = sqrt01(m.s13sq);
    m.p0 = c12sq * m.c13sq;
    m.c2 = 1.0 - 2.0 * s12sq;
    return m;
}

// ---- E items ----------------------------------------------------------------------------------------------------
// eigen entries 32 e .. 32 e + 31 of the [layer][fine energy] list, one per lane (osc.jl:466-470, 370-388).  One copy
// of make_entry: the instruction footprint matters more here than the instruction-level parallelism of two
// interleaved entries, because the other warps of the SM are busy with other items (see the header).
__device__ __forceinline__ void pipe_eigen_item(const DevTables& T, const double* __restrict__ p, int e, double* tbl) {
    const int lane = threadIdx.x & 31, nT = T.n_etrue, ne = 2 * nT, n_ent = ne * T.n_ltbl;
    const int t = 32 * kPipeEPL * e + lane;
    if (t >= n_ent) return;
    const PipeMix m = pipe_mix(p);
    double ent[kPipeEPL][kEntry];
    double* dst[kPipeEPL];
#pragma unroll
    for (int x = 0; x < kPipeEPL; ++x) {
        const int tx = min(t + 32 * x, n_ent - 1);   // a lane without a further entry repeats the last one (same values)
        const int k = tx / ne, el = tx - k * ne;
        const double en = T.e_calc[el];
        const double nrm = T.earth_unc ? p[SOLAROSC_N_FIXED_PARAMS + k] : 1.0;
        make_entry(m.dm21, nrm * T.earth_lookup[k] * en, en, m.u1, m.u2, m.u3, ent[x]);   // lookup = earth_norm .* earth_lookup (propagation_osc.jl:60)
        dst[x] = tbl + ((size_t)k * 2 * nT + (el & 1) * nT + (el >> 1)) * kEntry;
    }
#pragma unroll
    for (int x = 0; x < kPipeEPL; ++x)
#pragma unroll
        for (int y = 0; y < kEntry; ++y) dst[x][y] = ent[x][y];
}

// MSW day probability and cos(2 theta_m) per (process, fine energy), and the point's mixing constants.
// msw layout as in osc_stage: [2 proc][2 what][2 parity][nT].  item 0 also writes the constants.
__device__ __forceinline__ void pipe_msw_item(const DevTables& T, const PipeLaunch& L, const double* __restrict__ p, int item, double* msw,
                                              double* cst) {
    const int lane = threadIdx.x & 31, nT = T.n_etrue, ne = 2 * nT;
    const PipeMix m = pipe_mix(p);
    if (item == 0 && lane == 0) { cst[0] = m.u1; cst[1] = m.u2; cst[2] = m.u3; cst[3] = m.p0; }
    if (T.fast) {   // osc.jl:185-207 with n_e = production-weighted mean density
        for (int t = lane; t < 2 * ne; t += 32) {
            const int proc = t / ne, el = t - proc * ne;
            const double c2m = msw_c2m(T.e_calc[el], T.avg_ne[proc], m.dm21, m.c13sq, m.c2);
            const int slot = (el & 1) * nT + (el >> 1);
            msw[(size_t)(proc * 2 + 0) * 2 * nT + slot] = fma(0.5 * m.c13sq * m.c13sq, fma(m.c2, c2m, 1.0), m.s13sq * m.s13sq);
            msw[(size_t)(proc * 2 + 1) * 2 * nT + slot] = m.c13sq * c2m;
        }
    } else {        // osc.jl:139-182 collapsed to production-weighted means of cos(2 theta_m) (see osc_stage)
        const double s2sq = 1.0 - m.c2 * m.c2;
        const double* __restrict__ sne = T.solar_ne;
        const double* __restrict__ w8 = T.solar_w[0];
        const double* __restrict__ wh = T.solar_w[1];
        const int e0 = item * L.msw_per_item, e1 = min(ne, e0 + L.msw_per_item);
        for (int el = e0; el < e1; ++el) {
            const double kE = 2.0 * 1.4142135623730951 * kGF * T.e_calc[el] * m.c13sq / m.dm21;
            double a8 = 0.0, ah = 0.0;
#pragma unroll 4
            for (int r = lane; r < T.n_solar; r += 32) {
                const double x = fma(-sne[r], kE, m.c2);
                const double c = x * rsq(fma(x, x, s2sq));
                a8 = fma(w8[r], c, a8);
                ah = fma(wh[r], c, ah);
            }
            a8 = warp_sum(a8);
            ah = warp_sum(ah);
            if (lane == 0) {
                const int slot = (el & 1) * nT + (el >> 1);
#pragma unroll
                for (int proc = 0; proc < 2; ++proc) {
                    const double c2m = (proc ? ah : a8) / T.solar_wsum[proc];
                    msw[(size_t)(proc * 2 + 0) * 2 * nT + slot] = fma(0.5 * m.c13sq * m.c13sq, fma(m.c2, c2m, 1.0), m.s13sq * m.s13sq);
                    msw[(size_t)(proc * 2 + 1) * 2 * nT + slot] = m.c13sq * c2m;
                }
            }
        }
    }
}

// ---- P items ----------------------------------------------------------------------------------------------------
// item 0: day rows = block_average(prob_day, 2) (propagation_osc.jl:133-134)
__device__ __forceinline__ void pipe_day_item(const DevTables& T, const double* msw, double* p8s, double* phs) {
    const int nT = T.n_etrue;
    for (int j = threadIdx.x & 31; j < nT; j += 32) {
        p8s[j] = 0.5 * (msw[j] + msw[nT + j]);
        phs[j] = 0.5 * (msw[(size_t)4 * nT + j] + msw[(size_t)5 * nT + j]);
    }
}
// bundle of 32 (sorted zenith row, coarse energy bin) items: Earth propagation (osc.jl:436-477), IYSW night formula
// (osc.jl:204) and the (N_COSZ_SUB x 2) block average (propagation_osc.jl:136-137); the body of osc_stage's phase 2
__device__ __forceinline__ void pipe_prop_item(const DevTables& T, const SegTables& G, const double* tbl, const double* msw, const double* cst,
                                               int bundle, double* p8s, double* phs, int ldp) {
    const int nT = T.n_etrue;
    const int item = bundle * 32 + (threadIdx.x & 31);
    if (item >= T.n_cosz * nT) return;
    const double u1 = cst[0], u2 = cst[1], u3 = cst[2], p0 = cst[3];
    const int zr = item / nT, j = item - zr * nT;
    const int ls = 2 * nT * kEntry;                               // doubles between consecutive layers
    const double inv_block = 1.0 / (2.0 * T.n_sub);
    double acc8 = 0.0, acch = 0.0;
    const double* t0 = tbl + (size_t)j * kEntry;                  // even fine energy of this coarse bin, layer 0
    const double* t1 = t0 + (size_t)nT * kEntry;                  // odd fine energy
    for (int sub = 0; sub < T.n_sub; ++sub) {
        const int2 se = G.path_se[zr * T.n_sub + sub];
        const int s0 = se.x, s1 = se.y;
        double pe0, pe1;
        path_p1e_pair<true>(G, t0, t1, ls, s0, s1, u1, u2, u3, pe0, pe1);
        const double d0 = pe0 - p0, d1 = pe1 - p0;
        acc8 += fma(msw[(size_t)2 * nT + j], d0, msw[j]) + fma(msw[(size_t)3 * nT + j], d1, msw[nT + j]);
        acch += fma(msw[(size_t)6 * nT + j], d0, msw[(size_t)4 * nT + j]) + fma(msw[(size_t)7 * nT + j], d1, msw[(size_t)5 * nT + j]);
    }
    const int row = 1 + G.row_order[zr];
    p8s[row * ldp + j] = acc8 * inv_block;
    phs[row * ldp + j] = acch * inv_block;
}

// ---- F items ----------------------------------------------------------------------------------------------------
// Per-item vectors of one column block (the per-warp counterpart of fold_vectors): the CC B-fragment scale vectors
// (flux x cross-section shape, propagation_osc.jl:185-196) and the additive constants of the block's 8 reco bins
// (backgrounds, propagation_bg.jl:36-83; every ES component uses norm_1).  Returns the block's summed backgrounds of
// the bins from the threshold on in lanes 0..7 (asymmetry mode).
template <int MODE>
__device__ __forceinline__ double pipe_item_vectors(const DevTables& T, const DevDetector& D, const double* __restrict__ p, size_t b,
                                                    bool is_es, int jbase, bool first_rows, double* wscr, const RatesOut& O) {
    const int lane = threadIdx.x & 31, nT = T.n_etrue, kt = fold_kt(nT);
    const double f8 = p[3], fh = p[4];
    const double* pd = p + D.param_off;
    double* scale = wscr;
    double* cst = wscr + 8 * kt;
    __syncwarp();   // the warp's previous item is done with the scratch
    if (!is_es)
        for (int i = lane; i < 4 * kt; i += 32) {
            double sh = 0.0;
            if (i < nT) {
                const double l = D.log_e_norm[i];
                sh = exp(fma(p[6], l, p[7] * (l * l))) * p[5];
            }
            scale[i] = f8 * sh; scale[4 * kt + i] = fh * sh;
        }
    double acc = 0.0;
    if (lane < 8) {
        const int j = jbase + lane;
        double c = 0.0;
        if (is_es) {
            if (j < D.n_es) {
                for (int q = 0; q < D.n_bg_es; ++q) {
                    const double raw = D.raw_bg_es[q * D.n_es + j];
                    acc += (D.bg_es_area[q] != 0.0) ? raw * (D.bg_es_area[q] * pd[0]) : raw;
                }
                c = fma(f8, D.c8[j], fma(fh, D.ch[j], 0.5 * acc));   // F8 C8 + Fh Ch + BG/2
                if (MODE == kModeRates && O.bg_es && first_rows) O.bg_es[b * D.n_es + j] = acc;
            }
        } else if (j < D.n_cc) {
            for (int q = 0; q < D.n_bg_cc; ++q) {
                const double raw = D.raw_bg_cc[q * D.n_cc + j];
                acc += (D.bg_cc_norm[q] >= 0) ? raw * (pd[D.n_norm_es + D.bg_cc_norm[q]] * (1.0 / 2.2e-6)) : raw;
            }
            c = 0.5 * acc;
            if (MODE == kModeRates && O.bg_cc && first_rows) O.bg_cc[b * D.n_cc + j] = acc;
        }
        cst[lane] = c;
    }
    __syncwarp();
    return acc;
}

template <int MODE>
__global__ void __launch_bounds__(kPipeThreads, 1)
pipe_kernel(const DevTables T, const PipeLaunch L, const double* __restrict__ params, RatesOut O) {
    extern __shared__ __align__(16) double smem[];
    constexpr bool kThr = MODE == kModeLlh || MODE == kModeAsym;
    constexpr int kRed = MODE == kModeAsym ? 3 : 1;   // partial sums per fold item
    const int nT = T.n_etrue, nrow = 1 + T.n_cosz, ldp = fold_ldp(nT), kt = fold_kt(nT), mt = (nrow + 7) / 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const PipeSmem S = pipe_smem(T.n_ltbl, nT, T.n_cosz, L.group, L.n_f, kRed, seg_smem_doubles(T.n_seg, T.n_paths, T.n_cosz));
    unsigned* cnt = reinterpret_cast<unsigned*>(smem + S.cnt);
    unsigned* c_ticket = cnt;
    unsigned* c_e = cnt + 1;      // [2] completed E items of the points of each parity
    unsigned* c_p = cnt + 3;      // [2]
    unsigned* c_f = cnt + 5;      // [2]
    unsigned* c_pt = cnt + kPipeCounters;   // [2][group] completed F items per parameter point
    double* wscr = smem + S.wscr + (size_t)warp * S.wscr_sz;

    for (int i = threadIdx.x; i < kPipeCounters + 2 * L.group; i += blockDim.x) cnt[i] = 0u;
    const SegTables G = stage_seg_tables(T, smem + S.seg, true);
    for (int buf = 0; buf < 2; ++buf) fold_zero_padding(smem + S.ptile + buf * S.ptile_sz, nT, nrow, ldp);
    __syncthreads();

    const int N = (L.n_osc - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // oscillation points of this CTA
    const unsigned total_tickets = (unsigned)(N + 2) * (unsigned)L.block_len;
    const unsigned n_f_tot = (unsigned)L.group * (unsigned)L.n_f;

    SOLAROSC_PS_DECL;
#ifdef SOLAROSC_PIPE_STATS
    const long long ps_start = clock64();
#endif
    for (;;) {
        unsigned tk = 0;
        if (lane == 0) tk = atomicAdd(c_ticket, 1u);
        tk = __shfl_sync(0xffffffffu, tk, 0);
        if (tk >= total_tickets) break;
        const int blk = (int)(tk / (unsigned)L.block_len), slot = (int)(tk - (unsigned)blk * (unsigned)L.block_len);
        const unsigned short code = L.pattern[slot];
        const int kind = code >> 14, idx = code & 0x3fff;
        if (kind == kPipeP) {
            const int n = blk - 1;
            if (n < 0 || n >= N) continue;
            const int par = n & 1;
            const unsigned k = (unsigned)(n >> 1);
            SOLAROSC_PS_BEGIN();
            pipe_wait(c_e + par, (k + 1u) * (unsigned)L.n_e);
            SOLAROSC_PS_END(0);
            SOLAROSC_PS_BEGIN();
            pipe_wait(c_f + par, k * n_f_tot);                     // F(n-2) is done with this tile buffer
            SOLAROSC_PS_END(1);
            SOLAROSC_PS_BEGIN();
            double* p8s = smem + S.ptile + par * S.ptile_sz;
            double* phs = p8s + (size_t)nrow * ldp;
            const double* msw = smem + S.msw + par * S.msw_sz;
            if (idx == 0) pipe_day_item(T, msw, p8s, phs);
            else pipe_prop_item(T, G, smem + S.tbl + par * S.tbl_sz, msw, smem + S.cst + par * kPipeCst, idx - 1, p8s, phs, ldp);
            pipe_signal(c_p + par);
            SOLAROSC_PS_END(4);
        } else if (kind == kPipeE) {
            const int n = blk;
            if (n >= N) continue;
            const int par = n & 1;
            SOLAROSC_PS_BEGIN();
            pipe_wait(c_p + par, (unsigned)(n >> 1) * (unsigned)L.n_p);   // P(n-2) is done with this table buffer
            SOLAROSC_PS_END(2);
            SOLAROSC_PS_BEGIN();
            const size_t g = (size_t)blockIdx.x + (size_t)n * gridDim.x;
            const double* po = params + g * L.group * L.param_stride;
            if (idx < L.n_eig) pipe_eigen_item(T, po, idx, smem + S.tbl + par * S.tbl_sz);
            else pipe_msw_item(T, L, po, idx - L.n_eig, smem + S.msw + par * S.msw_sz, smem + S.cst + par * kPipeCst);
            pipe_signal(c_e + par);
            SOLAROSC_PS_END(5);
        } else {
            const int n = blk - 2;
            if (n < 0) continue;
            const int par = n & 1;
            const unsigned k = (unsigned)(n >> 1);
            SOLAROSC_PS_BEGIN();
            pipe_wait(c_p + par, (k + 1u) * (unsigned)L.n_p);
            SOLAROSC_PS_END(3);
            SOLAROSC_PS_BEGIN();
            const int q = idx / L.n_f, fi = idx - q * L.n_f;
            const PipeFoldItem it = L.fitems[fi];
            const DevDetector& D = T.det[it.det];
            const size_t g = (size_t)blockIdx.x + (size_t)n * gridDim.x;
            const size_t b = g * L.group + q;
            const double* p = params + b * L.param_stride;
            const double* p8s = smem + S.ptile + par * S.ptile_sz;
            const double* phs = p8s + (size_t)nrow * ldp;
            const int nte = (D.n_es + 7) / 8, ntc = (D.n_cc + 7) / 8;
            const int bte = kThr ? D.nte_thr : nte, btc = kThr ? D.ntc_thr : ntc;
            const bool is_es = it.col < bte;
            const int n0 = 8 * (is_es ? it.col : it.col - bte);
            const int idx0 = (is_es ? D.index_es : D.index_cc) - 1;
            const int nR = is_es ? D.n_es : D.n_cc;
            bool oob = false;   // earth-norm bounds (likelihood_debug.jl:1-4, likelihood_builder.jl:227)
            if (T.earth_unc)
                for (int kk = 0; kk < T.n_layers; ++kk) {
                    const double x = p[SOLAROSC_N_FIXED_PARAMS + kk];
                    oob |= !(x >= 0.0 && x <= 2.0);
                }
            double part = 0.0, d_es = 0.0, n_es = 0.0, d_cc = 0.0, n_cc = 0.0;
            uint32_t fl = 0;
            double bgsum = 0.0;
            const bool on = is_es ? D.es_mode != 0 : D.cc_mode != 0;
            if (on) {
                const double bg = pipe_item_vectors<MODE>(T, D, p, b, is_es, (kThr ? idx0 : 0) + n0, it.m0 == 0, wscr, O);
                if (MODE == kModeAsym && it.m0 == 0) {   // summed background of the block's bins from the threshold on
                    bgsum = (lane < 8 && idx0 + n0 + lane < nR) ? bg : 0.0;
                    bgsum = warp_sum(bgsum);
                }
                RatesLaunch RL{};
                RL.s_row_es = L.s_row_es; RL.s_j_es = L.s_j_es; RL.s_row_cc = L.s_row_cc; RL.s_j_cc = L.s_j_cc;
                // (mt = the item's last tile + 1: tiles past it are another item's, or past the end)
                FoldTask<MODE> F{T, RL, D, O, p8s, phs, wscr, wscr + 8 * kt, wscr + 8 * kt, p[3], p[4], b, ldp, nrow,
                                 min(mt, it.m0 + it.nt), kt, bte, btc, oob};
                F.cst_blk = wscr + 8 * kt;
#pragma unroll 1
                for (int m = it.m0; m < it.m0 + it.nt; m += kPipeNT) F.template run<kPipeNT>(it.col, m, part, fl, d_es, n_es, d_cc, n_cc);
            } else if (MODE == kModeRates || MODE == kModePerbin) {
                // disabled channels still produce arrays (propagation_reco.jl:12-15,119-122): zeros, or -Inf per-bin terms out of bounds
                const double v = (MODE == kModePerbin && oob) ? -INFINITY : 0.0;
                double* day = is_es ? O.es_day : O.cc_day;
                double* night = is_es ? O.es_night : O.cc_night;
                const long long sr = is_es ? L.s_row_es : L.s_row_cc, sj = is_es ? L.s_j_es : L.s_j_cc;
                const int nt = it.nt, nZ = nrow - 1;
                for (int e = lane; e < 8 * nt * 8; e += 32) {
                    const int row = 8 * it.m0 + (e >> 3), jx = n0 + (e & 7);
                    if (row >= nrow || jx >= nR) continue;
                    if (row == 0) { if (day) day[b * nR + jx] = v; }
                    else if (night) night[b * nZ * nR + (size_t)(row - 1) * sr + (size_t)jx * sj] = v;
                }
            }
            if (MODE == kModeLlh || MODE == kModeAsym) {
                // partial sums into the item's slot; the warp that completes the point adds the slots in order
                double* red = smem + S.red + par * S.red_sz + ((size_t)q * L.n_f + fi) * (kRed + 1);
                bool last = false;
                if (MODE == kModeLlh) {
                    part = warp_sum(part);
                    for (int o = 16; o > 0; o >>= 1) fl |= __shfl_xor_sync(0xffffffffu, fl, o);
                } else {
                    part = warp_sum(is_es ? d_es : d_cc);           // day total of this block
                    d_es = warp_sum(is_es ? n_es : n_cc);           // night total
                }
                if (lane == 0) {
                    red[0] = part;
                    if (MODE == kModeAsym) { red[1] = d_es; red[2] = bgsum; }
                    reinterpret_cast<uint32_t*>(red + kRed)[0] = fl;
                    unsigned old;
                    asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_addr(c_pt + par * L.group + q)) : "memory");
                    last = old + 1u == (k + 1u) * (unsigned)L.n_f;
                }
                last = __shfl_sync(0xffffffffu, last ? 1 : 0, 0) != 0;
                if (last && lane == 0) {
                    const double* r0 = smem + S.red + par * S.red_sz + (size_t)q * L.n_f * (kRed + 1);
                    if (MODE == kModeLlh) {
                        double s = 0.0;
                        uint32_t f = 0;
                        for (int x = 0; x < L.n_f; ++x) {
                            s += r0[(size_t)x * (kRed + 1)];
                            f |= reinterpret_cast<const uint32_t*>(r0 + (size_t)x * (kRed + 1) + kRed)[0];
                        }
                        double lnl = -s;
                        if (f & SOLAROSC_FLAG_NEG_ERROR) lnl = NAN;
                        if (oob) { lnl = -INFINITY; f = SOLAROSC_FLAG_EARTH_OOB; }
                        else if (lnl != lnl) f |= (f & SOLAROSC_FLAG_NEG_ERROR) ? 0u : SOLAROSC_FLAG_NONFINITE;
                        O.lnl[b] = lnl;
                        if (O.flags) O.flags[b] = f;
                    } else {
                        // derive_variables_from_chain.jl:236-258: half of the background above threshold is taken off the
                        // day total and half off the night total; a zero denominator gives 0.  (det_only >= 0: one detector)
                        for (int ch = 0; ch < 2; ++ch) {
                            double dt = 0.0, nt = 0.0, bg = 0.0;
                            bool any = false;
                            for (int x = 0; x < L.n_f; ++x) {
                                const PipeFoldItem jt = L.fitems[x];
                                if ((jt.col < T.det[jt.det].nte_thr) != (ch == 0)) continue;
                                any = true;
                                dt += r0[(size_t)x * (kRed + 1)]; nt += r0[(size_t)x * (kRed + 1) + 1]; bg += r0[(size_t)x * (kRed + 1) + 2];
                            }
                            double* out = ch == 0 ? O.asym_es : O.asym_cc;
                            if (!out) continue;
                            const bool ch_on = ch == 0 ? D.es_mode != 0 : D.cc_mode != 0;
                            const double dd = dt - 0.5 * bg, nn = nt - 0.5 * bg, den = dd + nn;
                            out[b] = (!ch_on || !any || den == 0.0) ? 0.0 : 2.0 * (dd - nn) / den;
                        }
                    }
                }
            }
            pipe_signal(c_f + par);
            SOLAROSC_PS_END(6);
        }
    }
#ifdef SOLAROSC_PIPE_STATS
    if (lane == 0) atomicAdd(&g_pipe_stats[7], (unsigned long long)(clock64() - ps_start));
#endif
}

}  // namespace solarosc
