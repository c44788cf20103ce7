// solarosc_adjoint.cuh -- adjoint_kernel: the oscillation columns of the gradient, exactly (derivation and the
// per-step algebra in solarosc_adjoint_math.cuh).
#pragma once
#include "solarosc_adjoint_math.cuh"
#include "solarosc_grad.cuh"

namespace solarosc {

// ------------------------------------------------------------------------------------------------------------------
// adjoint_kernel: one CTA per parameter point.  Input: Pbar[2][1 + n_cosz][n_etrue] of the point (written by
// grad_kernel).  Output: the oscillation columns of the gradient row (sin2_th12, sin2_th13, dm2_21, earth_norm_k)
// are ADDED to out_grad (grad_kernel zeroed the row and wrote the others).
// ------------------------------------------------------------------------------------------------------------------
struct AdjLaunch {
    int n_points, param_stride;
};
constexpr int kAdjG = 3 + SOLAROSC_MAX_LAYERS + 3;   // directions + Re(beta_0)

// Path groups: thread (g, el) walks the paths g, g + G, ... at fine energy el, one after the other, and adds into its own
// accumulators [g][layer][el]; the groups are added in order at the end.  No atomics: the gradient is deterministic, and
// the lanes of a warp (neighbouring energies of one group) walk the same path.
__host__ __device__ inline int adj_groups(int n_e) { return kThreads / n_e > 0 ? kThreads / n_e : 1; }
__host__ __device__ inline size_t adj_smem_doubles(int n_ltbl, int n_e, size_t seg_doubles) {
    // eigen table, (Y, Z) accumulators [G][n_ltbl][n_e][5] and W2 [G][2][n_e], MSW values/derivatives [2][n_e][8], reduction scratch
    const size_t G = (size_t)adj_groups(n_e);
    return (size_t)kEntryA * n_ltbl * n_e + G * ((size_t)5 * n_ltbl * n_e + (size_t)2 * n_e) + (size_t)16 * n_e + (size_t)kWarps * kAdjG + kAdjG + 2 +
           seg_doubles;
}

__global__ void __launch_bounds__(kThreads, 2)
adjoint_kernel(const DevTables T, const AdjLaunch L, const double* __restrict__ params, const double* __restrict__ pbar,
               double* __restrict__ out_grad, const uint32_t* __restrict__ flags) {
    extern __shared__ __align__(16) double smem[];
    const int nT = T.n_etrue, ne = 2 * nT, nl = T.n_ltbl, nrow = 1 + T.n_cosz;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t b = blockIdx.x;
    if (flags && (flags[b] & SOLAROSC_FLAG_EARTH_OOB)) return;   // lnL = -Inf: the gradient row stays zero
    const double* p = params + b * L.param_stride;
    const double* qb8 = pbar + b * 2 * nrow * nT;
    const double* qbh = qb8 + (size_t)nrow * nT;
    const int NG = adj_groups(ne);
    const size_t acc_sz = (size_t)5 * nl * ne;
    double* tbl = smem;                                   // [nl][ne][kEntryA]
    double* acc = tbl + (size_t)kEntryA * nl * ne;        // [NG][nl][ne][5]
    double* w2 = acc + (size_t)NG * acc_sz;               // [NG][2 proc][ne]: sum_p w (P1e - P0)
    double* msw = w2 + (size_t)NG * 2 * ne;               // [2 proc][ne][8]: Pday, K, dPday[3], dK[3]
    double* red = msw + (size_t)16 * ne;                  // [kWarps][kAdjG] + [kAdjG]
    double* segbuf = red + (size_t)kWarps * kAdjG + kAdjG + 2;
    segbuf = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(segbuf) + 15) & ~(uintptr_t)15);
    const SegTables G = stage_seg_tables(T, segbuf, true);

    const double x12 = fmin(fmax(p[0], 0.0), 1.0), y13 = fmin(fmax(p[1], 0.0), 1.0), dm21 = p[2];
    const bool x_live = p[0] > 0.0 && p[0] < 1.0, y_live = p[1] > 0.0 && p[1] < 1.0;
    const double c13sq = 1.0 - y13, c12sq = 1.0 - x12, c13 = sqrt(c13sq);
    const double u1 = sqrt(c12sq) * c13, u2 = sqrt(x12) * c13, u3 = sqrt(y13);
    const double p0 = c12sq * c13sq;

    // ---- phase 1: eigen table (three eigenvectors), MSW values and derivatives, cleared accumulators ------------
    for (int t = threadIdx.x; t < ne * nl; t += blockDim.x) {
        const int k = t / ne, el = t - k * ne;
        const double en = T.e_calc[el];
        const double nrm = T.earth_unc ? p[SOLAROSC_N_FIXED_PARAMS + k] : 1.0;
        double ent[kEntryA];
        make_entry_adj(dm21, nrm * T.earth_lookup[k] * en, en, u1, u2, u3, ent);
        double* d = tbl + (size_t)t * kEntryA;
#pragma unroll
        for (int x = 0; x < kEntryA; ++x) d[x] = ent[x];
    }
    for (int t = threadIdx.x; t < NG * ((int)acc_sz + 2 * ne); t += blockDim.x) acc[t] = 0.0;   // accumulators and W2 (contiguous)
    if (T.fast) {
        for (int t = threadIdx.x; t < 2 * ne; t += blockDim.x) {
            const int proc = t / ne, el = t - proc * ne;
            double v[2], d[6];
            msw_deriv(T.e_calc[el], T.avg_ne[proc], dm21, x12, y13, x_live, y_live, v, d);
            double* m = msw + (size_t)t * 8;
            m[0] = v[0]; m[1] = v[1];
#pragma unroll
            for (int x = 0; x < 6; ++x) m[2 + x] = d[x];
        }
    } else {   // production-region integral (osc.jl:139-182): production-weighted means, derivative of the mean = mean of derivatives
        for (int el = warp; el < ne; el += kWarps) {
            double s8[8], sh[8];
#pragma unroll
            for (int x = 0; x < 8; ++x) s8[x] = sh[x] = 0.0;
            for (int r = lane; r < T.n_solar; r += 32) {
                double v[2], d[6];
                msw_deriv(T.e_calc[el], T.solar_ne[r], dm21, x12, y13, x_live, y_live, v, d);
                const double a8 = T.solar_w[0][r], ah = T.solar_w[1][r];
                s8[0] = fma(a8, v[0], s8[0]); s8[1] = fma(a8, v[1], s8[1]); sh[0] = fma(ah, v[0], sh[0]); sh[1] = fma(ah, v[1], sh[1]);
#pragma unroll
                for (int x = 0; x < 6; ++x) { s8[2 + x] = fma(a8, d[x], s8[2 + x]); sh[2 + x] = fma(ah, d[x], sh[2 + x]); }
            }
#pragma unroll
            for (int x = 0; x < 8; ++x) { s8[x] = warp_sum(s8[x]); sh[x] = warp_sum(sh[x]); }
            if (lane == 0) {
#pragma unroll
                for (int x = 0; x < 8; ++x) {
                    msw[(size_t)el * 8 + x] = s8[x] / T.solar_wsum[0];
                    msw[(size_t)(ne + el) * 8 + x] = sh[x] / T.solar_wsum[1];
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 2: forward + backward sweep; thread (g, el) owns fine energy el of path group g ----------------------
    double b0[3] = {0.0, 0.0, 0.0};   // sum of Re(beta_0): the initial state's share of the mixing-angle derivatives
    const double inv_block = 1.0 / (2.0 * T.n_sub);
    const int n_paths = T.n_cosz * T.n_sub;
    for (int own = threadIdx.x; own < NG * ne; own += blockDim.x) {
        const int g = own / ne, el = own - g * ne, i = el >> 1;
        const double k8 = msw[(size_t)el * 8 + 1], kh = msw[(size_t)(ne + el) * 8 + 1];
        const double* te = tbl + (size_t)el * kEntryA;            // entry of layer 0 at this energy
        const size_t ls = (size_t)ne * kEntryA;
        double* my = acc + (size_t)g * acc_sz + (size_t)el * 5;   // + layer * ne * 5
        double w2a = 0.0, w2b = 0.0;
        for (int pth = g; pth < n_paths; pth += NG) {             // sorted paths (longest first), dealt round-robin to the groups
            const int zr = pth / T.n_sub, row = 1 + G.row_order[zr];
            const double w8 = qb8[(size_t)row * nT + i] * inv_block, wh = qbh[(size_t)row * nT + i] * inv_block;
            const int2 se = G.path_se[pth];
            State phi{u1, 0.0, u2, 0.0, u3, 0.0};
            for (int s = se.x; s < se.y; ++s) {
                const Seg sg = load_seg<true>(G, s);
                adj_step_fwd(te + (size_t)sg.layer * ls, sg.len, phi);
            }
            const double dp = fma(phi.ar, phi.ar, phi.ai * phi.ai) - p0;   // P1e - P0
            w2a = fma(w8, dp, w2a);
            w2b = fma(wh, dp, w2b);
            const double cw = 2.0 * fma(w8, k8, wh * kh);                 // c = 2 w K psi_1
            State beta{cw * phi.ar, cw * phi.ai, 0.0, 0.0, 0.0, 0.0};
            for (int s = se.y - 1; s >= se.x; --s) {
                const Seg sg = load_seg<true>(G, s);
                const AdjAcc o = adj_step_back(te + (size_t)sg.layer * ls, sg.len, phi, beta);
                double* a = my + (size_t)sg.layer * ne * 5;
                a[0] += o.y01; a[1] += o.y02; a[2] += o.y12; a[3] += o.z1; a[4] += o.z2;
            }
            b0[0] += beta.ar; b0[1] += beta.br; b0[2] += beta.cr;
        }
        w2[(size_t)g * 2 * ne + el] = w2a;
        w2[(size_t)g * 2 * ne + ne + el] = w2b;
    }
    __syncthreads();

    // ---- phase 3: directions ------------------------------------------------------------------------------------------
    double g[kAdjG];
#pragma unroll
    for (int x = 0; x < kAdjG; ++x) g[x] = 0.0;
    const MixDeriv md = mix_deriv(p[0], p[1]);
    for (int t = threadIdx.x; t < ne * nl; t += blockDim.x) {
        const int k = t / ne, el = t - k * ne;
        const double en = T.e_calc[el];
        const double nrm = T.earth_unc ? p[SOLAROSC_N_FIXED_PARAMS + k] : 1.0;
        const double aon = T.earth_lookup[k] * en;
        AdjAcc s{0.0, 0.0, 0.0, 0.0, 0.0};
        for (int q = 0; q < NG; ++q) {   // path groups, in order
            const double* a = acc + (size_t)q * acc_sz + (size_t)t * 5;
            s.y01 += a[0]; s.y02 += a[1]; s.y12 += a[2]; s.z1 += a[3]; s.z2 += a[4];
        }
        const double gn = adj_entry_directions(tbl + (size_t)t * kEntryA, s, nrm * aon, aon, en, md, g);
#pragma unroll
        for (int x = 0; x < SOLAROSC_MAX_LAYERS; ++x)
            if (x == k) g[3 + x] += gn;
    }
    // MSW factors: sum_{proc, e} [ dPday Wd + dK W2 ] - dP0 sum K W1n,  Wd = 1/2 colsum(Pbar) over all rows, W1n over the night rows
    for (int t = threadIdx.x; t < 2 * ne; t += blockDim.x) {
        const int proc = t / ne, el = t - proc * ne, i = el >> 1;
        const double* q = proc ? qbh : qb8;
        double night = 0.0;
        for (int row = 1; row < nrow; ++row) night += q[(size_t)row * nT + i];
        const double wd = 0.5 * (night + q[i]), w1n = 0.5 * night;
        const double* m = msw + (size_t)t * 8;
        const double kw = m[1] * w1n;
        double w2t = 0.0;
        for (int gq = 0; gq < NG; ++gq) w2t += w2[(size_t)gq * 2 * ne + t];
        g[0] += fma(m[2], wd, fma(m[5], w2t, x_live ? c13sq * kw : 0.0));     // dP0/dx = -(1 - y)
        g[1] += fma(m[3], wd, fma(m[6], w2t, y_live ? c12sq * kw : 0.0));     // dP0/dy = -(1 - x)
        g[2] += fma(m[4], wd, m[7] * w2t);
    }
    g[3 + SOLAROSC_MAX_LAYERS + 0] = b0[0]; g[3 + SOLAROSC_MAX_LAYERS + 1] = b0[1]; g[3 + SOLAROSC_MAX_LAYERS + 2] = b0[2];
    block_reduce(g, red, red + (size_t)kWarps * kAdjG);
    const double* tot = red + (size_t)kWarps * kAdjG;
    if (threadIdx.x == 0) {
        const int P = T.n_params;
        const double* bt = tot + 3 + SOLAROSC_MAX_LAYERS;
        out_grad[b * P + 0] += tot[0] + fma(bt[0], md.du_x[0], fma(bt[1], md.du_x[1], bt[2] * md.du_x[2]));
        out_grad[b * P + 1] += tot[1] + fma(bt[0], md.du_y[0], fma(bt[1], md.du_y[1], bt[2] * md.du_y[2]));
        out_grad[b * P + 2] += tot[2];
        if (T.earth_unc)
            for (int k = 0; k < T.n_layers; ++k) out_grad[b * P + SOLAROSC_N_FIXED_PARAMS + k] += k < nl ? tot[3 + k] : 0.0;
    }
}
}  // namespace solarosc
