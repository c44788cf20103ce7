// solarosc_setup.cuh -- set-up stage in C++: MC samples -> response / efficiency / background histograms with
// bit-exact integer binning, and the layered Earth path tables.
//
// In the reference these tables are built once by Julia code (response.jl:3-141, backgrounds.jl:37-140,
// histHelpers.jl:11-78, makePaths.jl:48-294, setup.jl:52-66) and then consumed by every likelihood evaluation.
// The functions here restate that stage behind the same C ABI so that a host can build what solarosc_create needs
// without Julia-side histogramming: the integer part (which bin an event falls into) is bit-exact by construction
// -- binary search over the same IEEE edges, left-closed bins, `searchsortedlast` semantics -- and weights are
// summed in file order like the reference's loop.  The bin search of large MC samples also exists as a kernel
// (one thread per event; HBM-bound, 12 bytes per event) behind solarosc_setup_bin_index_device.
// Nothing here runs per evaluation.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <numeric>
#include <vector>

#include "../../include/solarosc_b200.h"

namespace solarosc_setup {

// numpy.linspace(lo, hi, n + 1): lo + i*step in plain double, last point exactly hi.  Julia's
// range(lo, stop=hi, length=n+1) is a twice-precision StepRangeLen whose elements are correctly rounded and can differ
// from these by an ulp: bin indices are bit-exact against this repository's numpy builders (and the shipped CC MC, whose
// values sit far from the edges), parity of the edges with Julia is unverified -- solarosc_setup_bin_index also takes
// explicit edge arrays, which is how a Julia host passes its own.
inline std::vector<double> linspace(double lo, double hi, int n_intervals) {
    std::vector<double> e(n_intervals + 1);
    const double step = (hi - lo) / n_intervals;
    for (int i = 0; i <= n_intervals; ++i) e[i] = lo + i * step;
    e[n_intervals] = hi;
    return e;
}

// searchsortedlast(edges, v): number of edges <= v (1-based bin, 0 below the first edge, n_edges above the last)
__host__ __device__ inline int32_t search_last(const double* edges, int n_edges, double v) {
    int lo = 0, hi = n_edges;              // first index with edges[idx] > v
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (edges[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void bin_index_kernel(const double* __restrict__ edges, int n_edges, const double* __restrict__ v, size_t n,
                                 int32_t* __restrict__ out) {
    extern __shared__ double s_edges[];
    for (int i = threadIdx.x; i < n_edges; i += blockDim.x) s_edges[i] = edges[i];
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = search_last(s_edges, n_edges, v[i]);
}

struct Bins {
    int n;
    double lo, hi;
    std::vector<double> edges;             // n + 1, possibly with the outer two at -Inf / +Inf
};
inline Bins make_bins(int n, double lo, double hi, const double* edges) {
    Bins b{n, lo, hi, {}};
    if (edges) b.edges.assign(edges, edges + n + 1);
    else b.edges = linspace(lo, hi, n);
    return b;
}

// create_histogram (histHelpers.jl:11-78): left-closed bins, events outside are dropped, weights in file order
inline void histogram(const double* v, const double* w, const uint8_t* keep, size_t n, const Bins& b, bool normalise, double* out) {
    std::fill(out, out + b.n, 0.0);
    for (size_t i = 0; i < n; ++i) {
        if (keep && !keep[i]) continue;
        const int k = search_last(b.edges.data(), b.n + 1, v[i]);
        if (k >= 1 && k <= b.n) out[k - 1] += w ? w[i] : 1.0;
    }
    if (normalise) {
        double s = 0.0;
        for (int k = 0; k < b.n; ++k) s += out[k];
        for (int k = 0; k < b.n; ++k) out[k] /= s;
    }
}

// create_response_matrix (response.jl:3-67): exact zeros are nudged to 1e-9, the in-range test uses the
// UN-extended range although the reco bins run to +-Inf (response.jl:47 vs :98-104), indices are clamped to the
// last bin, rows are normalised (empty rows stay zero).
inline void response_matrix(const double* x, const double* y, size_t n, const Bins& bx, const Bins& by, double* out,
                            int32_t* xi, int32_t* yi) {
    std::fill(out, out + (size_t)bx.n * by.n, 0.0);
    for (size_t i = 0; i < n; ++i) {
        const double xv = x[i] == 0.0 ? 1e-9 : x[i], yv = y[i] == 0.0 ? 1e-9 : y[i];
        const bool ok = bx.lo <= xv && xv <= bx.hi && by.lo <= yv && yv <= by.hi;
        const int ti = std::min<int>(search_last(bx.edges.data(), bx.n + 1, xv), bx.n);
        const int ri = std::min<int>(search_last(by.edges.data(), by.n + 1, yv), by.n);
        if (xi) xi[i] = ok ? ti : 0;
        if (yi) yi[i] = ok ? ri : 0;
        if (ok) out[(size_t)(ti - 1) * by.n + (ri - 1)] += 1.0;
    }
    for (int r = 0; r < bx.n; ++r) {
        double s = 0.0;
        for (int c = 0; c < by.n; ++c) s += out[(size_t)r * by.n + c];
        if (s > 0.0)
            for (int c = 0; c < by.n; ++c) out[(size_t)r * by.n + c] /= s;
    }
}

// ---- layered Earth paths (makePaths.jl:48-294, setup.jl:52-66) -----------------------------------------------------
constexpr int kResolution = 12000;         // makePaths.jl:45
constexpr double kEarthRadius = 6371.0;    // int(EARTH_RADIUS_KM)
constexpr double kDetectorDepth = 0.001;   // makePaths.jl:10-11

struct Jump { double x, b, a; };
struct Segment { int index; double length, avg_rho; };

// linear interpolation on sorted knots, flat outside them (np.interp / Interpolations with Flat())
inline double interp(const std::vector<double>& xs, const std::vector<double>& ys, double q) {
    if (q <= xs.front()) return ys.front();
    if (q >= xs.back()) return ys.back();
    const size_t j = std::upper_bound(xs.begin(), xs.end(), q) - xs.begin();   // xs[j-1] <= q < xs[j]
    const double slope = (ys[j] - ys[j - 1]) / (xs[j] - xs[j - 1]);
    return slope * (q - xs[j - 1]) + ys[j - 1];
}

// find_jumps (makePaths.jl:51-107): steps larger than 5 x the median step, consecutive ones merged
inline std::vector<Jump> find_jumps(const std::vector<double>& x, const std::vector<double>& f, double factor = 5.0) {
    const size_t n = x.size();
    std::vector<double> ad(n - 1);
    for (size_t i = 0; i + 1 < n; ++i) ad[i] = std::fabs(f[i + 1] - f[i]);
    std::vector<double> srt(ad);
    std::sort(srt.begin(), srt.end());
    const size_t m = srt.size();
    const double med = (m % 2) ? srt[m / 2] : 0.5 * (srt[m / 2 - 1] + srt[m / 2]);
    const double thr = factor * med;
    std::vector<size_t> idx;
    for (size_t i = 0; i < m; ++i)
        if (ad[i] > thr) idx.push_back(i);
    if (idx.empty()) return {Jump{x.front(), 0.0, f.back()}, Jump{x.back(), f.front(), f.front()}};
    std::vector<Jump> jumps{Jump{x.front(), 0.0, f.back()}};
    size_t s = idx[0], e = idx[0];
    auto close_group = [&]() { jumps.push_back(Jump{(x[s] + x[e + 1]) / 2, f[s], f[e + 1]}); };
    for (size_t k = 1; k < idx.size(); ++k) {
        if (idx[k] == e + 1) e = idx[k];
        else { close_group(); s = e = idx[k]; }
    }
    close_group();
    jumps.push_back(Jump{x.back(), jumps.back().a, jumps.back().a});
    return jumps;
}

struct PathBuilder {
    std::vector<double> rx, ry;            // Earth model knots (radius, potential)
    std::vector<Jump> radial;
    std::vector<std::pair<Jump, Jump>> layers;

    void track(double cosz, std::vector<double>& depth, std::vector<double>& dist) const {
        const double R = kEarthRadius, d = kDetectorDepth;
        const double length = std::sqrt((R - d) * (R - d) * cosz * cosz + 2 * R * d - d * d) - (R - d) * cosz;
        depth.resize(kResolution); dist.resize(kResolution);
        if (cosz != -1) {
            const double sinz = std::sqrt(1 - cosz * cosz), cotz = cosz / sinz;
            const double x_end = length / std::sqrt(1 + cotz * cotz);
            const std::vector<double> xs = linspace(0.0, x_end, kResolution - 1);
            const double k = std::sqrt(1 + cotz * cotz);
            for (int i = 0; i < kResolution; ++i) {
                const double ys = (R - d) + xs[i] * cotz;
                depth[i] = std::sqrt(xs[i] * xs[i] + ys * ys);
                dist[i] = xs[i] * k;
            }
        } else {
            const std::vector<double> ys = linspace(R - d, -R, kResolution - 1);
            for (int i = 0; i < kResolution; ++i) {
                depth[i] = std::sqrt(0.0 + ys[i] * ys[i]);
                dist[i] = R - d - ys[i];
            }
        }
    }

    // potential profile along the chord and the discontinuity crossings on it (makePaths.jl:149-177)
    void potential_along(double cosz, std::vector<double>& dist, std::vector<double>& mp, std::vector<Jump>& out) const {
        std::vector<double> depth;
        track(cosz, depth, dist);
        mp.resize(kResolution);
        for (int i = 0; i < kResolution; ++i) mp[i] = interp(rx, ry, depth[i]);
        out.clear();
        for (size_t j = 0; j + 1 < radial.size(); ++j) {
            const double jd = radial[j].x;
            for (int i = 0; i + 1 < kResolution; ++i) {
                const double lo = depth[i], hi = depth[i + 1];
                const bool outward = lo <= jd && hi >= jd;
                const bool inward = !outward && lo >= jd && hi <= jd;
                if (!(outward || inward)) continue;
                const double t = (jd - depth[i]) / (depth[i + 1] - depth[i]);
                const double xd = dist[i] + t * (dist[i + 1] - dist[i]);
                out.push_back(outward ? Jump{xd, radial[j].a, radial[j].b} : Jump{xd, radial[j].b, radial[j].a});
            }
        }
        out.push_back(Jump{0.0, mp[0], mp[0]});
        out.push_back(Jump{dist[kResolution - 1], 0.0, mp[kResolution - 1]});
        std::stable_sort(out.begin(), out.end(), [](const Jump& p, const Jump& q) { return p.x < q.x; });
    }

    bool init(const double* radius, const double* potential, size_t n) {
        rx.assign(radius, radius + n); ry.assign(potential, potential + n);
        const std::vector<double> xv = linspace(0.0, kEarthRadius, kResolution - 1);
        std::vector<double> fv(kResolution);
        for (int i = 0; i < kResolution; ++i) {
            if (xv[i] < rx.front() || xv[i] > rx.back()) return false;   // the reference's interpolant throws
            fv[i] = interp(rx, ry, xv[i]);
        }
        radial = find_jumps(xv, fv);
        std::vector<double> dist, mp;
        std::vector<Jump> ref;
        potential_along(-1, dist, mp, ref);
        const size_t half = (ref.size() + 1) / 2;
        layers.clear();
        for (size_t i = 0; i + 1 < half; ++i) layers.push_back({ref[i], ref[i + 1]});
        layers.push_back({ref[half - 1], ref[half - 1]});
        return true;
    }

    int layer_of(double b_left, double a_right) const {
        const double c0 = std::min(b_left, a_right), c1 = std::max(b_left, a_right);
        for (size_t li = 0; li < layers.size(); ++li) {
            const Jump &l0 = layers[li].first, &l1 = layers[li].second;
            if (std::fabs(c0 - c1) <= 1e-7) {
                if (std::fabs(c0 - l0.b) <= 1e-7) return (int)li + 1;
            } else {
                const double m0 = std::min(l0.b, l1.a), m1 = std::max(l0.b, l1.a);
                if (std::fabs(c0 - m0) <= 1e-7 && std::fabs(c1 - m1) <= 1e-10) return (int)li + 1;
            }
        }
        return -1;
    }

    // segments of one chord: boundaries at the crossings, mean potential from n_nodes + 2 equally spaced samples
    bool path(double cosz, std::vector<Segment>& segs, int n_nodes = 3) const {
        std::vector<double> dist, mp;
        std::vector<Jump> jumps;
        potential_along(cosz, dist, mp, jumps);
        segs.clear();
        for (size_t i = 0; i + 1 < jumps.size(); ++i) {
            const double a = jumps[i].x, b = jumps[i + 1].x;
            const std::vector<double> nodes = linspace(a, b, n_nodes + 1);
            std::vector<double> vals(n_nodes + 2);
            for (int k = 0; k < n_nodes + 2; ++k) vals[k] = interp(dist, mp, nodes[k]);
            vals.front() = jumps[i].b;
            vals.back() = jumps[i + 1].a;
            const int idx = layer_of(jumps[i].b, jumps[i + 1].a);
            if (idx < 0) return false;
            // numpy mean: pairwise summation degenerates to a plain left-to-right sum for 5 values
            double s = 0.0;
            for (double v : vals) s += v;
            segs.push_back(Segment{idx, b - a, s / vals.size()});
        }
        return true;
    }
};

// mean segment potential per layer index over all paths (get_avg_densities, makePaths.jl:268-294)
inline std::vector<double> avg_densities(const std::vector<std::vector<Segment>>& paths) {
    int max_idx = 0;
    for (auto& p : paths)
        for (auto& s : p) max_idx = std::max(max_idx, s.index);
    std::vector<double> sum(max_idx + 1, 0.0);
    std::vector<int> cnt(max_idx + 1, 0);
    for (auto& p : paths)
        for (auto& s : p) { sum[s.index] += s.avg_rho; ++cnt[s.index]; }
    std::vector<double> out;
    for (int u = 1; u <= max_idx; ++u)
        if (cnt[u]) out.push_back(sum[u] / cnt[u]);
    return out;
}

}  // namespace solarosc_setup
