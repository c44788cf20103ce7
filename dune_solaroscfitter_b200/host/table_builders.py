"""Host-side producers of the constant tables that cross the C-ABI boundary.

In the reference these tables are built once at set-up time by Julia code and
then consumed by every likelihood evaluation (SURVEY.md section 8(a) row a21,
Appendix A.1).  A Julia host keeps doing that and hands the arrays to
``solarosc_create``; this module is the same set-up stage restated in
Python/numpy so that tests, ``bench.py`` and the Python mirror of the
reference interface can run where no Julia exists.  Nothing here runs per
evaluation and nothing here touches the GPU.

Each function names the reference code it follows (paths relative to the
reference checkout).  Setup-time quadratures (``QuadGK`` at its default
``rtol ~ 1.5e-8`` in the reference) are replaced by composite Gauss-Legendre
rules that are exact for the piecewise-linear integrands, so these tables agree
with Julia's only to the accuracy of Julia's own quadrature; parity of the hot
path is therefore always stated *at fixed tables* (hard
part 5).
"""
from __future__ import annotations

import csv
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------
# constants (src/core.jl:22-54, src/earthProfile.jl:5, src/oscillations/makePaths.jl:10-11,45,
#            src/unoscillatedSample.jl:9-11)
# --------------------------------------------------------------------------
EARTH_RADIUS_KM = 6371.0
RESOLUTION = 12000
DETECTOR_DEPTH_KM = 0.001
POTENTIAL_FACTOR = 1.52588e-4
M_E = 5.11e-4
SIN2_W = 0.231
RHO_NC = 1.0081
KAPPA_EW = 1.0329
SIGMA_0 = 88.06e-46
DETECTOR_NE = 10 * 2.7e32
DETECTOR_NAR40 = 10 * 1.45e31
DETECTION_TIME = 3600 * 24 * 365


@dataclass
class BinSpec:
    """``(bin_number, min, max[, bins])`` named tuple of the reference (readConfig.jl:143-149)."""
    bin_number: int
    min: float
    max: float
    bins: Optional[np.ndarray] = None


def calculate_bins(b: BinSpec) -> Tuple[np.ndarray, np.ndarray]:
    """Edges ``min + i*w`` and centres ``min + (i+0.5)*w`` (unoscillatedSample.jl:46-58)."""
    w = (b.max - b.min) / b.bin_number
    edges = np.array([b.min + i * w for i in range(b.bin_number + 1)], dtype=np.float64)
    centres = np.array([b.min + (i + 0.5) * w for i in range(b.bin_number)], dtype=np.float64)
    return edges, centres


# --------------------------------------------------------------------------
# piecewise-linear helper (Interpolations.linear_interpolation stand-in)
# --------------------------------------------------------------------------
class Linear1D:
    """Linear interpolant on sorted knots; ``mode`` = 'throw' | 'flat'."""

    def __init__(self, x, y, mode="throw"):
        self.x = np.asarray(x, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64)
        self.mode = mode

    def __call__(self, q):
        q = np.asarray(q, dtype=np.float64)
        if self.mode == "throw":
            if np.any(q < self.x[0]) or np.any(q > self.x[-1]):
                raise ValueError("interpolation query outside the knot range")
        return np.interp(q, self.x, self.y)  # np.interp is flat outside the knots

    def integral(self, a: float, b: float) -> float:
        """Exact integral of the (flat-extrapolated) interpolant over [a, b]."""
        if b < a:
            return -self.integral(b, a)
        inner = self.x[(self.x > a) & (self.x < b)]
        knots = np.concatenate(([a], inner, [b]))
        vals = self(knots) if self.mode == "flat" else np.interp(knots, self.x, self.y)
        return float(np.sum(0.5 * (vals[1:] + vals[:-1]) * np.diff(knots)))


_GL_X, _GL_W = np.polynomial.legendre.leggauss(12)


def _integrate_piecewise(f, a: float, b: float, breaks: np.ndarray) -> float:
    """Composite 12-point Gauss-Legendre between the knots of a piecewise-smooth integrand."""
    inner = breaks[(breaks > a) & (breaks < b)]
    knots = np.concatenate(([a], inner, [b]))
    total = 0.0
    for lo, hi in zip(knots[:-1], knots[1:]):
        h = 0.5 * (hi - lo)
        xs = lo + h * (_GL_X + 1.0)
        total += h * float(np.dot(_GL_W, f(xs)))
    return total


# --------------------------------------------------------------------------
# Earth model and cos(z) grid  (src/earthProfile.jl)
# --------------------------------------------------------------------------
def load_earth_model(path: str) -> Dict[str, np.ndarray]:
    """PREM table -> radius [km] and matter potential (earthProfile.jl:7-33)."""
    data = np.loadtxt(path)
    radius = data[:, 0] * EARTH_RADIUS_KM * 1.0000000001
    potential = data[:, 1] * data[:, 2] * POTENTIAL_FACTOR
    return {"radius": radius, "potential": potential}


def _boundary_cosz(model) -> Tuple[float, float]:
    """cos z of the two largest potential jumps (earthProfile.jl:74-82)."""
    r, p = model["radius"], model["potential"]
    jumps = np.abs(np.diff(p))
    top2 = np.sort(np.argsort(-jumps, kind="stable")[:2])
    r_b = [(r[i] + r[i + 1]) / 2.0 for i in top2]
    cz = sorted(-math.sqrt(max(0.0, 1.0 - (rb / EARTH_RADIUS_KM) ** 2)) for rb in r_b)
    return cz[0], cz[1]


def _alloc_cosz_bins(n_total: int, seg_lengths: np.ndarray) -> np.ndarray:
    """Largest-remainder allocation with >=1 bin per PREM zone (earthProfile.jl:98-113)."""
    if n_total < len(seg_lengths):
        raise ValueError("not enough cos z bins to cover the PREM zones")
    fracs = seg_lengths / np.sum(seg_lengths)
    counts = np.maximum(1, np.floor(fracs * n_total).astype(int))
    rem = fracs * n_total - counts
    for _ in range(n_total - int(np.sum(counts))):
        i = int(np.argmax(rem))
        counts[i] += 1
        rem[i] -= 1.0
    return counts


def _piecewise_edges(breaks: Sequence[float], n_per: Sequence[int]) -> np.ndarray:
    edges = [breaks[0]]
    for i, n in enumerate(n_per):
        edges.extend(np.linspace(breaks[i], breaks[i + 1], n + 1)[1:])
    return np.array(edges, dtype=np.float64)


@dataclass
class CoszGrid:
    edges: np.ndarray      # COARSE_COSZ_EDGES
    n_sub: int             # N_COSZ_SUB
    cosz_calc: np.ndarray  # fine midpoints fed to the path builder
    cosz: np.ndarray       # coarse centres


def build_cosz_grid(model, n_bins: int, exposure_cosz: Optional[np.ndarray],
                    n_fine: int = 0, fine_edge: float = -0.12,
                    cosz_min: float = -1.0, cosz_max: float = 0.0) -> CoszGrid:
    """Coarse/fine zenith grid (earthProfile.jl:84-178)."""
    ic, cm = _boundary_cosz(model)
    exp_min = cosz_min
    if exposure_cosz is not None:
        night = exposure_cosz[exposure_cosz <= 0.0]
        if night.size:
            exp_min = float(np.min(night))

    def prem_breaks(lo, hi):
        br = [lo]
        for b in (ic, cm):
            if lo < b < hi:
                br.append(b)
        br.append(hi)
        return br

    if n_fine > 0 and exp_min < fine_edge < cosz_max:
        nf = min(n_fine, n_bins - 1)
        nd = n_bins - nf
        deep_breaks = prem_breaks(exp_min, fine_edge)
        deep = _piecewise_edges(deep_breaks, _alloc_cosz_bins(nd, np.diff(deep_breaks)))
        a = abs(fine_edge)
        fine = np.array([-(a * (j / nf) ** 1.0) for j in range(nf, -1, -1)], dtype=np.float64)
        edges = np.concatenate((deep, fine[1:]))
    else:
        br = prem_breaks(exp_min, cosz_max)
        edges = _piecewise_edges(br, _alloc_cosz_bins(n_bins, np.diff(br)))
    if len(edges) != n_bins + 1:
        raise AssertionError("coarse cos z grid has the wrong number of edges")
    n_sub = 2 if n_bins < 40 else 1
    calc = []
    for i in range(n_bins):
        lo, hi = edges[i], edges[i + 1]
        w = hi - lo
        calc.extend(lo + (k - 0.5) * w / n_sub for k in range(1, n_sub + 1))
    cosz = np.array([(edges[i] + edges[i + 1]) / 2 for i in range(n_bins)])
    return CoszGrid(edges=edges, n_sub=n_sub, cosz_calc=np.array(calc), cosz=cosz)


# --------------------------------------------------------------------------
# Earth paths  (src/oscillations/makePaths.jl)
# --------------------------------------------------------------------------
@dataclass
class Jump:
    x: float
    b: float
    a: float


@dataclass
class Segment:
    index: int
    start: float
    finish: float
    length: float
    avg_rho: float


@dataclass
class Path:
    jumps: List[Jump]
    segments: List[Segment]


def find_jumps(x: np.ndarray, f: np.ndarray, threshold_factor: float = 5.0) -> List[Dict[str, float]]:
    """Discontinuities of a sampled profile (makePaths.jl:51-107)."""
    adiff = np.abs(np.diff(f))
    thr = threshold_factor * float(np.median(adiff))
    idx = np.nonzero(adiff > thr)[0]
    if idx.size == 0:
        return [dict(x=x[0], b=0.0, a=f[-1]), dict(x=x[-1], b=f[0], a=f[0])]
    groups, cur = [], [int(idx[0])]
    for i in idx[1:]:
        i = int(i)
        if i == cur[-1] + 1:
            cur.append(i)
        else:
            groups.append(cur)
            cur = [i]
    groups.append(cur)
    jumps = [dict(x=float(x[0]), b=0.0, a=float(f[-1]))]
    for g in groups:
        s, e = g[0], g[-1]
        jumps.append(dict(x=float((x[s] + x[e + 1]) / 2), b=float(f[s]), a=float(f[e + 1])))
    jumps.append(dict(x=float(x[-1]), b=jumps[-1]["a"], a=jumps[-1]["a"]))
    return jumps


def _track(cosz: float) -> Tuple[np.ndarray, np.ndarray]:
    """Radii and distances of RESOLUTION points along the chord (makePaths.jl:109-146)."""
    R, d = float(int(EARTH_RADIUS_KM)), DETECTOR_DEPTH_KM
    length = math.sqrt((R - d) ** 2 * cosz ** 2 + 2 * R * d - d ** 2) - (R - d) * cosz
    if cosz != -1:
        sinz = math.sqrt(1 - cosz ** 2)
        cotz = cosz / sinz
        x_end = length / math.sqrt(1 + cotz ** 2)
        xs = np.linspace(0.0, x_end, RESOLUTION)
        ys = (R - d) + xs * cotz
        dist = xs * math.sqrt(1 + cotz ** 2)
    else:
        xs = np.zeros(RESOLUTION)
        ys = np.linspace(R - d, -R, RESOLUTION)
        dist = R - d - ys
    return np.sqrt(xs ** 2 + ys ** 2), dist


def _potential_along(cosz: float, earth: Linear1D, ref_jumps) -> Tuple[np.ndarray, np.ndarray, List[Jump]]:
    """Potential profile and discontinuity crossings of one chord (makePaths.jl:149-177)."""
    depths, dist = _track(cosz)
    mp = earth(depths)
    out: List[Jump] = []
    lo, hi = depths[:-1], depths[1:]
    for jmp in ref_jumps[:-1]:
        jd = jmp["x"]
        outward = (lo <= jd) & (hi >= jd)
        inward = (~outward) & (lo >= jd) & (hi <= jd)
        for i in np.nonzero(outward | inward)[0]:
            t = (jd - depths[i]) / (depths[i + 1] - depths[i])
            xd = dist[i] + t * (dist[i + 1] - dist[i])
            if outward[i]:
                out.append(Jump(float(xd), jmp["a"], jmp["b"]))
            else:
                out.append(Jump(float(xd), jmp["b"], jmp["a"]))
    out.append(Jump(0.0, float(mp[0]), float(mp[0])))
    out.append(Jump(float(dist[-1]), 0.0, float(mp[RESOLUTION - 1])))
    out.sort(key=lambda j: j.x)  # stable
    return dist, mp, out


class PathBuilder:
    """``make_potential_for_integrand`` with the radial jump scan done once (makePaths.jl:180-266)."""

    def __init__(self, earth: Linear1D, n_layers_expected: Optional[int] = None, n_nodes: int = 3):
        self.earth = earth
        self.n = n_nodes
        xvals = np.linspace(0.0, float(int(EARTH_RADIUS_KM)), RESOLUTION)
        self.radial_jumps = find_jumps(xvals, earth(xvals))
        _, _, ref = _potential_along(-1, earth, self.radial_jumps)
        half = math.ceil(len(ref) / 2)
        col = ref[:half]
        self.layers = [(col[i], col[i + 1]) for i in range(len(col) - 1)] + [(col[-1], col[-1])]
        if n_layers_expected is not None and len(self.layers) != n_layers_expected:
            raise ValueError(f"{len(self.layers)} Earth layers found, {n_layers_expected} expected")

    def _layer_of(self, b_left: float, a_right: float) -> int:
        c = sorted((b_left, a_right))
        for li, (l0, l1) in enumerate(self.layers, start=1):
            if abs(c[0] - c[1]) <= 1e-7:
                if abs(c[0] - l0.b) <= 1e-7:
                    return li
            else:
                lc = sorted((l0.b, l1.a))
                if abs(c[0] - lc[0]) <= 1e-7 and abs(c[1] - lc[1]) <= 1e-10:
                    return li
        raise ValueError(f"no Earth layer matches the segment with boundary potentials {c}")

    def path(self, cosz: float) -> Path:
        dist, mp, jumps = _potential_along(cosz, self.earth, self.radial_jumps)
        f = Linear1D(dist, mp, mode="flat")
        segs = []
        for i in range(len(jumps) - 1):
            a, b = jumps[i].x, jumps[i + 1].x
            nodes = np.linspace(a, b, self.n + 2)
            vals = np.array(f(nodes))
            vals[0] = jumps[i].b
            vals[-1] = jumps[i + 1].a
            idx = self._layer_of(jumps[i].b, jumps[i + 1].a)
            segs.append(Segment(idx, a, b, b - a, float(np.mean(vals))))
        return Path(jumps, segs)


def get_avg_densities(paths: Sequence[Path]) -> np.ndarray:
    """Mean segment potential per layer index (makePaths.jl:268-294)."""
    pairs = [(s.index, s.avg_rho) for p in paths for s in p.segments]
    pairs.sort(key=lambda t: t[0])  # stable, like Julia's default
    idx = np.array([p[0] for p in pairs])
    rho = np.array([p[1] for p in pairs])
    out = []
    for u in sorted(set(idx.tolist())):
        sel = rho[idx == u]
        out.append(float(np.sum(sel) / sel.size))
    return np.array(out)


def build_earth_tables(earth: Linear1D, cosz_calc: np.ndarray, n_layers: Optional[int]):
    """``earth_paths`` + ``earth_lookup`` padded from the vertical path (setup.jl:52-66)."""
    pb = PathBuilder(earth, n_layers_expected=n_layers)
    paths = [pb.path(float(z)) for z in cosz_calc]
    base = get_avg_densities(paths)
    n_exp = len(base) if n_layers is None else n_layers
    if len(base) < n_exp:
        ref = get_avg_densities([pb.path(-1.0)])
        base = np.concatenate((base, ref[len(base):n_exp]))
    return paths, base


def pack_paths(paths: Sequence[Path]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Ragged segment tables -> CSR arrays in path order (detector -> far surface)."""
    start = np.zeros(len(paths) + 1, dtype=np.int32)
    layer, length = [], []
    for i, p in enumerate(paths):
        for s in p.segments:
            layer.append(s.index)
            length.append(s.length)
        start[i + 1] = len(layer)
    return start, np.array(layer, dtype=np.int32), np.array(length, dtype=np.float64)


# --------------------------------------------------------------------------
# exposure, solar model, flux, cross-sections, unoscillated sample
# --------------------------------------------------------------------------
def load_exposure(path: str) -> Tuple[np.ndarray, np.ndarray]:
    raw = np.loadtxt(path, delimiter=",")
    return raw[:, 0], raw[:, 1]


def exposure_weights(cosz_x, cosz_y, edges: np.ndarray, cosz_max: float = 0.0) -> np.ndarray:
    """Fraction of night exposure per coarse zenith bin (exposure.jl:8-43)."""
    f = Linear1D(cosz_x, cosz_y, mode="flat")
    total = f.integral(float(edges[0]), cosz_max)
    return np.array([f.integral(float(lo), float(hi)) for lo, hi in zip(edges[:-1], edges[1:])]) / total


def solar_model(tables: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Production-weighted mean electron densities (solarModel.jl:3-25)."""
    ne = tables["n_e"]
    fb, fh = tables["prodFractionBoron"], tables["prodFractionHep"]
    return dict(radii=tables["radii"], n_e=ne, prodFractionBoron=fb, prodFractionHep=fh,
                avgNeBoron=float(np.sum(fb * ne) / np.sum(fb)),
                avgNeHep=float(np.sum(fh * ne) / np.sum(fh)))


def normalised_fluxes(energies, flux8b, fluxhep, etrue_max: float) -> Tuple[Linear1D, Linear1D]:
    """Unit-integral piecewise-linear spectra (solarFlux.jl:8-34)."""
    xe = np.concatenate(([0.0], energies))
    f8 = np.concatenate(([0.0], flux8b))
    fh = np.concatenate(([0.0], fluxhep))
    if etrue_max > energies[-1]:
        xe = np.concatenate((xe, [etrue_max]))
        f8 = np.concatenate((f8, [0.0]))
        fh = np.concatenate((fh, [0.0]))
    i8 = Linear1D(xe, f8).integral(xe[0], xe[-1])
    ih = Linear1D(xe, fh).integral(xe[0], xe[-1])
    return Linear1D(xe, f8 / i8), Linear1D(xe, fh / ih)


def load_cc_xsec(path: str, scale: float = 1.0) -> Linear1D:
    """MARLEY nu_e-40Ar total cross-section, flat outside the table (xsec.jl:13-32)."""
    raw = np.loadtxt(path, delimiter=",")
    e = np.concatenate(([0.0], raw[:, 0] * 1e-3))
    s = np.concatenate(([0.0], raw[:, 1] * 1e-42)) * scale
    order = np.argsort(e, kind="stable")
    e, s = e[order], s[order]
    ue, first = np.unique(e, return_index=True)
    return Linear1D(ue, s[first], mode="flat")


def _es_xsec(enu: np.ndarray, g1: float, g2: float, emin: float) -> np.ndarray:
    tmax = 2 * enu ** 2 / (M_E + 2 * enu)
    return SIGMA_0 / M_E * ((g1 ** 2 + g2 ** 2) * (tmax - emin)
                            - (g2 ** 2 + g1 * g2 * M_E / (2 * enu)) * (tmax ** 2 - emin ** 2) / enu
                            + 1 / 3 * g2 ** 2 * ((tmax ** 3 - emin ** 3) / enu ** 2))


def es_xsec_nue(enu, emin: float):
    """nu_e e- elastic scattering above recoil threshold, clipped at 0 (xsec.jl:37-62)."""
    g1 = math.sqrt(RHO_NC) * (0.5 + KAPPA_EW * SIN2_W)
    g2 = math.sqrt(RHO_NC) * KAPPA_EW * SIN2_W
    return np.maximum(_es_xsec(np.asarray(enu, dtype=np.float64), g1, g2, emin), 0.0)


def es_xsec_nuother(enu, emin: float):
    """nu_mu,tau e- elastic scattering, zero for E <= threshold only (xsec.jl:64-93)."""
    enu = np.asarray(enu, dtype=np.float64)
    g1 = math.sqrt(RHO_NC) * (-0.5 + KAPPA_EW * SIN2_W)
    g2 = math.sqrt(RHO_NC) * KAPPA_EW * SIN2_W
    with np.errstate(divide="ignore", invalid="ignore"):
        x = _es_xsec(enu, g1, g2, emin)
    return np.where(enu <= emin, 0.0, x)


def build_unoscillated_sample(etrue: BinSpec, flux8b: Linear1D, fluxhep: Linear1D, cc_xsec: Linear1D,
                              es_threshold: float, es_norm: float, cc_norm: float) -> Dict[str, np.ndarray]:
    """Flux x cross-section integrated per true-energy bin (unoscillatedSample.jl:86-116)."""
    edges, centres = calculate_bins(etrue)
    breaks = np.unique(np.concatenate((flux8b.x, cc_xsec.x, [es_threshold])))

    def per_bin(f):
        return np.array([_integrate_piecewise(f, float(a), float(b), breaks)
                         for a, b in zip(edges[:-1], edges[1:])])

    kes = DETECTOR_NE * DETECTION_TIME * es_norm
    kcc = DETECTOR_NAR40 * DETECTION_TIME * cc_norm
    out = {
        "ES_nue_8B": per_bin(lambda e: flux8b(e) * es_xsec_nue(e, es_threshold)) * kes,
        "ES_nuother_8B": per_bin(lambda e: flux8b(e) * es_xsec_nuother(e, es_threshold)) * kes,
        "CC_8B": per_bin(lambda e: flux8b(e) * cc_xsec(e)) * kcc,
        "ES_nue_hep": per_bin(lambda e: fluxhep(e) * es_xsec_nue(e, es_threshold)) * kes,
        "ES_nuother_hep": per_bin(lambda e: fluxhep(e) * es_xsec_nuother(e, es_threshold)) * kes,
        "CC_hep": per_bin(lambda e: fluxhep(e) * cc_xsec(e)) * kcc,
    }
    pivot = np.sum(out["CC_8B"] * centres) / np.sum(out["CC_8B"])
    out["log_E_norm"] = np.log(centres / pivot)
    return out


# --------------------------------------------------------------------------
# histogramming, response matrices, backgrounds
# --------------------------------------------------------------------------
def extended_reco_bins(b: BinSpec) -> BinSpec:
    """Reco edges with the outer two pushed to -Inf/+Inf (response.jl:98-104)."""
    e = np.linspace(b.min, b.max, b.bin_number + 1)
    e[0], e[-1] = -np.inf, np.inf
    return BinSpec(b.bin_number, b.min, b.max, e)


def bin_index_left_closed(edges: np.ndarray, v: np.ndarray) -> np.ndarray:
    """1-based ``searchsortedlast(edges, v)``: number of edges <= v (bit-exact integer result)."""
    return np.searchsorted(edges, v, side="right").astype(np.int64)


def create_histogram(data: np.ndarray, b: BinSpec, weights: Optional[np.ndarray] = None,
                     normalise: bool = True) -> np.ndarray:
    """Left-closed histogram, weights summed in file order (histHelpers.jl:11-78)."""
    edges = b.bins if b.bins is not None else np.linspace(b.min, b.max, b.bin_number + 1)
    k = bin_index_left_closed(edges, np.asarray(data, dtype=np.float64))
    ok = (k >= 1) & (k <= b.bin_number)
    h = np.zeros(b.bin_number, dtype=np.float64)
    if weights is None:
        np.add.at(h, k[ok] - 1, 1.0)
    else:
        np.add.at(h, k[ok] - 1, np.asarray(weights, dtype=np.float64)[ok])
    return h / np.sum(h) if normalise else h


def response_bin_indices(x, y, bx: BinSpec, by: BinSpec) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Integer (true, reco) bin of every MC event and the in-range mask (response.jl:19-54)."""
    x = np.where(np.asarray(x, dtype=np.float64) == 0, 1e-9, x)
    y = np.where(np.asarray(y, dtype=np.float64) == 0, 1e-9, y)
    ex = bx.bins if bx.bins is not None else np.linspace(bx.min, bx.max, bx.bin_number + 1)
    ey = by.bins if by.bins is not None else np.linspace(by.min, by.max, by.bin_number + 1)
    ok = (bx.min <= x) & (x <= bx.max) & (by.min <= y) & (y <= by.max)
    ti = np.minimum(bin_index_left_closed(ex, x), bx.bin_number)
    ri = np.minimum(bin_index_left_closed(ey, y), by.bin_number)
    return ti, ri, ok


def create_response_matrix(x, y, bx: BinSpec, by: BinSpec) -> np.ndarray:
    """Row-normalised P(reco bin | true bin) (response.jl:3-67)."""
    ti, ri, ok = response_bin_indices(x, y, bx, by)
    m = np.zeros((bx.bin_number, by.bin_number), dtype=np.float64)
    np.add.at(m, (ti[ok] - 1, ri[ok] - 1), 1.0)
    rs = m.sum(axis=1, keepdims=True)
    return np.divide(m, rs, out=m.copy(), where=rs > 0)


def selection_efficiency(ereco: np.ndarray, mask: np.ndarray, b_ext: BinSpec) -> np.ndarray:
    """Selected / total per extended reco bin, 0 where empty (response.jl:128-141)."""
    sel = create_histogram(ereco[mask], b_ext, normalise=False)
    tot = create_histogram(ereco, b_ext, normalise=False)
    return np.where(tot == 0, 0.0, sel / np.where(tot == 0, 1.0, tot))


def read_csv_columns(path: str) -> Dict[str, np.ndarray]:
    with open(path, newline="") as fh:
        rd = csv.reader(fh)
        header = next(rd)
        cols = [[] for _ in header]
        for row in rd:
            for c, v in zip(cols, row):
                c.append(v)
    out = {}
    for name, c in zip(header, cols):
        if c and c[0] in ("True", "False", "true", "false"):
            out[name] = np.array([v in ("True", "true") for v in c], dtype=bool)
        else:
            out[name] = np.array(c, dtype=np.float64)
    return out


def cc_background_component(ereco, weights, mask, b_ext: BinSpec, cc_norm: float) -> np.ndarray:
    """Selected weighted CC background histogram x 10 kt x exposure (backgrounds.jl:85-99)."""
    m = np.ones(len(ereco), dtype=bool) if mask is None else mask
    w = None if weights is None else weights[m]
    return create_histogram(ereco[m], b_ext, weights=w, normalise=False) * 10 * cc_norm


def es_background_component(shape_norm, eff, n_total: float, es_norm: float) -> np.ndarray:
    """Raw ES background table from a normalised shape (backgrounds.jl:53-83)."""
    attenuation = n_total / 50e6
    return shape_norm * DETECTION_TIME * eff * es_norm * attenuation


def threshold_index(b_ext: BinSpec, thr: float) -> int:
    """1-based ``findfirst(edge > thr)`` on the extended edge vector (setup.jl:169-178)."""
    hits = np.nonzero(b_ext.bins > thr)[0]
    if hits.size == 0:
        raise ValueError("threshold exceeds all reco bins")
    return int(hits[0]) + 1


# ---------------------------------------------------------------------------------------------
# inclusive / semi-inclusive / angular tables  (response.jl:111-236)
# ---------------------------------------------------------------------------------------------
def hermite_boost(eff: np.ndarray, es_bins: BinSpec) -> np.ndarray:
    """The temporary high-energy efficiency boost of the inclusive modes: smooth-step from 10 to 12 MeV
    towards a 90 % plateau (response.jl:147-160,176-179)."""
    edges = np.linspace(es_bins.min, es_bins.max, es_bins.bin_number + 1)
    centres = 0.5 * (edges[:-1] + edges[1:])
    t = np.clip((centres - 0.010) / (0.012 - 0.010), 0.0, 1.0)
    smooth = t ** 2 * (3 - 2 * t)
    return eff + smooth * (0.9 - eff)


def cc_inclusive_tables(cc_etrue, cc_ereco, cc_mask, etrue: BinSpec, es_bins: BinSpec):
    """CC events histogrammed in the ES reco binning: response and boosted selection efficiency
    (response.jl:167-180)."""
    es_ext = extended_reco_bins(es_bins)
    resp = create_response_matrix(cc_etrue, cc_ereco, etrue, es_ext)
    eff = hermite_boost(selection_efficiency(cc_ereco, cc_mask, es_ext), es_bins)
    return resp, eff


def angular_mask(cos_bins: BinSpec, cos_cut: float) -> np.ndarray:
    """1.0 for cos-scatter bins whose centre is at or above the cut (response.jl:117-120)."""
    edges = np.linspace(cos_bins.min, cos_bins.max, cos_bins.bin_number + 1)
    centres = 0.5 * (edges[:-1] + edges[1:])
    return (centres >= cos_cut).astype(np.float64)


def angular_tables(ang_ereco, ang_cos, es_bins: BinSpec, cos_bins: BinSpec, cos_cut: float):
    """ES.angular (n_cos, n_ereco): P(cos bin | reco bin) from the angular MC; BG.angular: uniform over the
    allowed bins (response.jl:113-123)."""
    es_ext = extended_reco_bins(es_bins)
    r_ang = create_response_matrix(ang_ereco, ang_cos, es_ext, cos_bins).T
    return (r_ang,) + angular_bg_tables(r_ang.shape, cos_bins, cos_cut)


def angular_bg_tables(shape, cos_bins: BinSpec, cos_cut: float):
    m = angular_mask(cos_bins, cos_cut)
    r_bg = np.zeros(shape)
    r_bg[m > 0, :] = 1.0 / np.count_nonzero(m)
    return r_bg, m


def misid_tables(resp_es_nue, resp_es_nuother, r_ang, amask, resp_cc, etrue: BinSpec, es_bins: BinSpec):
    """Semi-inclusive tables: CC split fractions and the two-step ES -> CC mis-identification responses
    (response.jl:182-218).  Uses the extended ES edges, whose outer centres are infinite, exactly as the
    reference does."""
    n_cos = len(amask)
    f_above = np.count_nonzero(amask) / n_cos
    f_below = 1.0 - f_above
    es_ext = extended_reco_bins(es_bins)
    with np.errstate(invalid="ignore"):
        es_centres = 0.5 * (es_ext.bins[:-1] + es_ext.bins[1:])
    etrue_edges = np.linspace(etrue.min, etrue.max, etrue.bin_number + 1)
    as_ereco = np.zeros((es_bins.bin_number, resp_cc.shape[1]))
    for j in range(es_bins.bin_number):
        b = min(int(np.searchsorted(etrue_edges, es_centres[j], side="right")), etrue.bin_number)   # searchsortedlast
        if b >= 1:
            as_ereco[j, :] = resp_cc[b - 1, :]
    backward = r_ang[amask == 0, :].sum(axis=0)
    return (resp_es_nue * backward[None, :]) @ as_ereco, (resp_es_nuother * backward[None, :]) @ as_ereco, f_above, f_below
