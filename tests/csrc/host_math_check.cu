// host_math_check.cu -- TEST HARNESS: runs the kernels' arithmetic helpers (sincos_phase,
// eig_rank_one, earth_step, msw_c2m from csrc/solarosc_device.cuh) on the CPU so the algebra of the
// CUDA path can be checked against the oracle in the GPU-less container.  Not part of the product.
#include <vector>

#include "../../dune_solaroscfitter_b200/csrc/solarosc_device.cuh"
#include "../../dune_solaroscfitter_b200/csrc/solarosc_adjoint_math.cuh"

using namespace solarosc;

extern "C" {

void hm_sincos(const double* x, int n, double* s, double* c) {
    for (int i = 0; i < n; ++i) sincos_phase(x[i], s[i], c[i]);
}

// sine and cosine of len[i] * kq[i] quarter turns (the layer step's phase factor)
void hm_sincos_qt(const double* len, const double* kq, int n, double* s, double* c) {
    for (int i = 0; i < n; ++i) sincos_qt(len[i], kq[i], s[i], c[i]);
}

void hm_eig(double d21, double d31, double a, double u1, double u2, double u3, double* q, double* w) {
    eig_rank_one(d21, d31, a, u1, u2, u3, q, w[0], w[1]);
}

// P1e[n_paths][n_e] with the kernel's own table builder and step functions (one energy chunk)
void hm_p1e(const double* e_calc, int n_e, const int* path_start, const int* app_layer, const double* app_len,
            int n_paths, const double* lookup, int nl, double s12sq_in, double s13sq_in, double dm21, double* p1e) {
    const int cbs = n_e / 2;
    std::vector<double> tbl((size_t)kEntry * nl * n_e);
    const double s12sq = fmin(fmax(s12sq_in, 0.0), 1.0), s13sq = fmin(fmax(s13sq_in, 0.0), 1.0);
    const double c13sq = 1.0 - s13sq, c12sq = 1.0 - s12sq, c13 = sqrt(c13sq);
    const double u1 = sqrt(c12sq) * c13, u2 = sqrt(s12sq) * c13, u3 = sqrt(s13sq);
    for (int k = 0; k < nl; ++k)
        for (int el = 0; el < n_e; ++el) {
            const int slot = (el & 1) * cbs + (el >> 1);
            make_entry(dm21, lookup[k] * e_calc[el], e_calc[el], u1, u2, u3, tbl.data() + ((size_t)k * 2 * cbs + slot) * kEntry);
        }
    for (int path = 0; path < n_paths; ++path)
        for (int el = 0; el < n_e; ++el) {
            const int slot = (el & 1) * cbs + (el >> 1);
            p1e[(size_t)path * n_e + el] = path_p1e(tbl.data() + (size_t)slot * kEntry, (size_t)2 * cbs * kEntry, app_layer,
                                                    app_len, path_start[path], path_start[path + 1], u1, u2, u3);
        }
}

// the same with the mirrored-pair walk on every eligible path (path_p1e_mirror); returns the number of eligible paths
int hm_p1e_mirror(const double* e_calc, int n_e, const int* path_start, const int* app_layer, const double* app_len,
                  int n_paths, const double* lookup, int nl, double s12sq_in, double s13sq_in, double dm21, double kq_max, double* p1e) {
    const int cbs = n_e / 2;
    std::vector<double> tbl((size_t)kEntry * nl * n_e);
    const double s12sq = fmin(fmax(s12sq_in, 0.0), 1.0), s13sq = fmin(fmax(s13sq_in, 0.0), 1.0);
    const double c13sq = 1.0 - s13sq, c12sq = 1.0 - s12sq, c13 = sqrt(c13sq);
    const double u1 = sqrt(c12sq) * c13, u2 = sqrt(s12sq) * c13, u3 = sqrt(s13sq);
    for (int k = 0; k < nl; ++k)
        for (int el = 0; el < n_e; ++el) {
            const int slot = (el & 1) * cbs + (el >> 1);
            make_entry(dm21, lookup[k] * e_calc[el], e_calc[el], u1, u2, u3, tbl.data() + ((size_t)k * 2 * cbs + slot) * kEntry);
        }
    const int n_seg = path_start[n_paths];
    std::vector<double> mir((size_t)2 * n_seg, 0.0);
    int n_ok = 0;
    for (int path = 0; path < n_paths; ++path) {
        const int s0 = path_start[path], s1 = path_start[path + 1];
        const bool ok = mirror_eligible(app_layer, app_len, s0, s1, kq_max);
        n_ok += ok;
        if (ok)
            for (int i = 1; i < (s1 - s0) / 2; ++i) {
                const double dr = 1.5707963267948966 * (app_len[s1 - 1 - i] - app_len[s0 + i]);
                mir[2 * (s0 + i)] = dr; mir[2 * (s0 + i) + 1] = -0.5 * dr;
            }
        for (int el = 0; el < n_e; ++el) {
            const int slot = (el & 1) * cbs + (el >> 1);
            const double* t = tbl.data() + (size_t)slot * kEntry;
            p1e[(size_t)path * n_e + el] = ok ? path_p1e_mirror(t, (size_t)2 * cbs * kEntry, app_layer, app_len, mir.data(), s0, s1, u1, u2, u3)
                                              : path_p1e(t, (size_t)2 * cbs * kEntry, app_layer, app_len, s0, s1, u1, u2, u3);
        }
    }
    return n_ok;
}

// P1e and its exact derivatives with respect to (sin2_th12, sin2_th13, dm2_21, earth_norm_0..nl-1) for every
// (path, energy), with the adjoint kernel's own step functions: forward sweep, backward sweep with weight
// c = 2 psi_1 (so that Re(beta^dagger x) = d|psi_1|^2), direction contractions per table entry.
void hm_adj_p1e_grad(const double* e_calc, int n_e, const int* path_start, const int* app_layer, const double* app_len,
                     int n_paths, const double* lookup, const double* norms, int nl, double s12sq_in, double s13sq_in, double dm21,
                     double* p1e, double* grad) {
    const double s12sq = fmin(fmax(s12sq_in, 0.0), 1.0), s13sq = fmin(fmax(s13sq_in, 0.0), 1.0);
    const double c13sq = 1.0 - s13sq, c12sq = 1.0 - s12sq, c13 = sqrt(c13sq);
    const double u1 = sqrt(c12sq) * c13, u2 = sqrt(s12sq) * c13, u3 = sqrt(s13sq);
    const MixDeriv md = mix_deriv(s12sq_in, s13sq_in);
    std::vector<double> tbl((size_t)kEntryA * nl * n_e);
    for (int k = 0; k < nl; ++k)
        for (int el = 0; el < n_e; ++el)
            make_entry_adj(dm21, norms[k] * lookup[k] * e_calc[el], e_calc[el], u1, u2, u3, tbl.data() + ((size_t)k * n_e + el) * kEntryA);
    const int nd = 3 + nl;
    for (int path = 0; path < n_paths; ++path)
        for (int el = 0; el < n_e; ++el) {
            const double* te = tbl.data() + (size_t)el * kEntryA;
            const size_t ls = (size_t)n_e * kEntryA;
            State phi{u1, 0.0, u2, 0.0, u3, 0.0};
            for (int s = path_start[path]; s < path_start[path + 1]; ++s) adj_step_fwd(te + app_layer[s] * ls, app_len[s], phi);
            p1e[(size_t)path * n_e + el] = phi.ar * phi.ar + phi.ai * phi.ai;
            State beta{2.0 * phi.ar, 2.0 * phi.ai, 0.0, 0.0, 0.0, 0.0};
            std::vector<AdjAcc> acc(nl, AdjAcc{0, 0, 0, 0, 0});
            for (int s = path_start[path + 1] - 1; s >= path_start[path]; --s) {
                const AdjAcc o = adj_step_back(te + app_layer[s] * ls, app_len[s], phi, beta);
                AdjAcc& a = acc[app_layer[s]];
                a.y01 += o.y01; a.y02 += o.y02; a.y12 += o.y12; a.z1 += o.z1; a.z2 += o.z2;
            }
            double* g = grad + ((size_t)path * n_e + el) * nd;
            for (int d = 0; d < nd; ++d) g[d] = 0.0;
            for (int k = 0; k < nl; ++k) {
                const double aon = lookup[k] * e_calc[el];
                g[3 + k] = adj_entry_directions(te + k * ls, acc[k], norms[k] * aon, aon, e_calc[el], md, g);
            }
            g[0] += beta.ar * md.du_x[0] + beta.br * md.du_x[1] + beta.cr * md.du_x[2];
            g[1] += beta.ar * md.du_y[0] + beta.br * md.du_y[1] + beta.cr * md.du_y[2];
        }
}

// MSW day probability, K = c13^2 cos(2 theta_m) and their derivatives (adjoint kernel's msw_deriv)
void hm_msw_deriv(double e, double ne, double dm21, double x12, double y, double* v, double* d) {
    msw_deriv(e, ne, dm21, x12, y, x12 > 0.0 && x12 < 1.0, y > 0.0 && y < 1.0, v, d);
}

double hm_msw_c2m(double e, double ne, double dm21, double one_m_s13sq, double c2) {
    return msw_c2m(e, ne, dm21, one_m_s13sq, c2);
}
}
