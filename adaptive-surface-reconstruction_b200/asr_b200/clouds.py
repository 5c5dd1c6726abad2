"""Seeded synthetic oriented point clouds for the BASELINE.json configurations
(SURVEY.md §8d).  Nothing is stored but the seed; numpy only (host side).

Every generator returns a dict with float32 `points` [N,3], unit `normals` [N,3],
per-point footprint `radii` [N] and the bounding box `bb_min`/`bb_max` the Python
path of the reference uses: min/max -/+ 0.1 (models/v0/datareader.py:225-227).

Radii are the analytic k=24 nearest-neighbour radius of the sampling density
(r = sqrt(k / (pi * rho)) for a surface sampled with rho points per unit area),
i.e. what asr.KDTree.compute_k_radius (nsearch.cpp:30-52) estimates, without
running a 10 M-point CPU KD-tree inside the benchmark.
"""
import numpy as np

K_NEIGHBOURS = 24


def _finish(points, normals, radii):
    points = np.ascontiguousarray(points, np.float32)
    normals = np.ascontiguousarray(normals, np.float32)
    radii = np.ascontiguousarray(radii, np.float32)
    return {
        "points": points,
        "normals": normals,
        "radii": radii,
        "bb_min": (points.min(0) - np.float32(0.1)).astype(np.float32),
        "bb_max": (points.max(0) + np.float32(0.1)).astype(np.float32),
    }


def _unit(v):
    return v / np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-30)


def sphere(n=100_000, seed=0):
    """Config 1: n points uniform on the unit sphere, constant radius sqrt(96/n)."""
    rng = np.random.default_rng(seed)
    p = _unit(rng.standard_normal((n, 3))).astype(np.float32)
    r = np.full(n, np.sqrt(96.0 / n), np.float32)
    return _finish(p, p.copy(), r)


def gaussian_blob(n=1_000_000, seed=1, sigma=0.25):
    """Config 2: p ~ N(0, sigma^2 I) clipped to |p| < 1; normals = p/|p|; radii =
    analytic k-NN radius of the local volumetric density -> strongly varying
    radii -> many octree levels."""
    rng = np.random.default_rng(seed)
    p = np.empty((0, 3))
    while p.shape[0] < n:
        q = rng.standard_normal((int(n * 1.05) + 16, 3)) * sigma
        p = np.concatenate([p, q[np.linalg.norm(q, axis=1) < 1.0]])
    p = p[:n]
    d2 = (p * p).sum(1)
    rho = n * np.exp(-0.5 * d2 / sigma**2) / ((2 * np.pi) ** 1.5 * sigma**3)
    r = (3.0 * K_NEIGHBOURS / (4.0 * np.pi * rho)) ** (1.0 / 3.0)
    return _finish(p, _unit(p), np.minimum(r, 0.5))


# ------------------------------------------------------------------ "Thingi10k-shaped" scenes


def _sample_sphere(rng, n):
    d = _unit(rng.standard_normal((n, 3)))
    return 0.5 * d, d, 4 * np.pi * 0.25


def _sample_torus(rng, n, R=0.36, r=0.14):
    # rejection sampling for area-uniform samples on a torus (diameter 2(R+r) = 1)
    u = np.empty(0)
    while u.shape[0] < n:
        cand = rng.uniform(0, 2 * np.pi, 2 * n)
        acc = rng.uniform(0, R + r, 2 * n) < (R + r * np.cos(cand))
        u = np.concatenate([u, cand[acc]])
    u = u[:n]
    v = rng.uniform(0, 2 * np.pi, n)
    cu, su, cv, sv = np.cos(u), np.sin(u), np.cos(v), np.sin(v)
    p = np.stack([(R + r * cu) * cv, (R + r * cu) * sv, r * su], 1)
    nrm = np.stack([cu * cv, cu * sv, su], 1)
    return p, nrm, 4 * np.pi**2 * R * r


def _sample_rounded_box(rng, n, half=0.26, rad=0.08):
    # surface of the Minkowski sum box(half) + ball(rad): project a sphere sample outward
    d = _unit(rng.standard_normal((n, 3)))
    t = half / np.maximum(np.abs(d).max(1, keepdims=True), 1e-9)
    core = np.clip(d * t * 1.6, -half, half)
    nrm = _unit(d * t * 1.6 - core + 1e-9 * d)
    p = core + rad * nrm
    area = 6 * (2 * half) ** 2 + 4 * np.pi * rad**2 + 12 * (2 * half) * (np.pi / 2) * rad
    return p / 1.2, nrm, area / 1.44


def _sample_gyroid_patch(rng, n, cells=1.5):
    # points on the gyroid level set inside the unit ball, by Newton projection
    p = rng.uniform(-0.5, 0.5, (int(n * 2.2) + 64, 3))
    k = 2 * np.pi * cells
    for _ in range(6):
        x, y, z = (k * p).T
        f = np.sin(x) * np.cos(y) + np.sin(y) * np.cos(z) + np.sin(z) * np.cos(x)
        g = k * np.stack([np.cos(x) * np.cos(y) - np.sin(z) * np.sin(x),
                          -np.sin(x) * np.sin(y) + np.cos(y) * np.cos(z),
                          -np.sin(y) * np.sin(z) + np.cos(z) * np.cos(x)], 1)
        p = p - (f / np.maximum((g * g).sum(1), 1e-9))[:, None] * g
    keep = np.linalg.norm(p, axis=1) < 0.5
    p, g = p[keep][:n], g[keep][:n]
    if p.shape[0] < n:  # extremely unlikely; pad by repetition
        reps = int(np.ceil(n / max(p.shape[0], 1)))
        p, g = np.tile(p, (reps, 1))[:n], np.tile(g, (reps, 1))[:n]
    return p, _unit(g), 3.1 * cells * (4 / 3 * np.pi * 0.125) * 2.0


_SHAPES = (_sample_sphere, _sample_torus, _sample_rounded_box, _sample_gyroid_patch)


def thingi_like(n=10_000_000, seed=2, num_shapes=16, outlier_fraction=0.005, noise=0.0):
    """Config 3: union of `num_shapes` closed analytic shapes (spheres, tori,
    rounded boxes, gyroid patches), each normalised to unit hull diameter and
    packed on a grid without overlap (like datareader.py:345-386), per-shape
    sampling density varying 1-8x, plus uniform outliers."""
    rng = np.random.default_rng(seed)
    n_out = int(n * outlier_fraction)
    n_in = n - n_out
    weight = rng.uniform(1.0, 8.0, num_shapes)
    kinds = [(_SHAPES[i % len(_SHAPES)]) for i in range(num_shapes)]
    # shape areas are needed for the split: evaluate with a tiny sample
    areas = np.array([k(np.random.default_rng(0), 8)[2] for k in kinds])
    share = weight * areas
    counts = np.floor(n_in * share / share.sum()).astype(np.int64)
    counts[0] += n_in - counts.sum()
    side = int(np.ceil(num_shapes ** (1 / 3)))
    pts, nrms, rads = [], [], []
    for i, (kind, c) in enumerate(zip(kinds, counts)):
        p, nr, area = kind(rng, int(c))
        scale = rng.uniform(0.7, 1.0)
        centre = 1.15 * np.array([i % side, (i // side) % side, i // (side * side)], np.float64)
        rho = c / (area * scale**2)
        r = np.sqrt(K_NEIGHBOURS / (np.pi * rho)) * rng.uniform(0.85, 1.15, p.shape[0])
        if noise > 0:
            p = p + rng.laplace(0, noise, p.shape) * nr
        pts.append(p * scale + centre)
        nrms.append(nr)
        rads.append(r)
    P = np.concatenate(pts)
    lo, hi = P.min(0), P.max(0)
    po = rng.uniform(lo, hi, (n_out, 3))
    pts.append(po)
    nrms.append(_unit(rng.standard_normal((n_out, 3))))
    rads.append(np.full(n_out, float(np.median(np.concatenate(rads))) * 4.0))
    P = np.concatenate(pts)
    perm = rng.permutation(P.shape[0])
    return _finish(P[perm], np.concatenate(nrms)[perm], np.concatenate(rads)[perm])


def multi_scan(n=50_000_000, seed=3, scans=5):
    """Config 5: `scans` simulated scans of the config-3 geometry with different
    densities and Laplace noise sigma in {0, .5, 1, 1.5, 2} * 1e-3
    (like datareader.py:441-446,555-557), fused into one cloud."""
    parts = []
    share = np.linspace(1.0, 2.0, scans)
    share = share / share.sum()
    for s in range(scans):
        parts.append(thingi_like(int(n * share[s]) if s else n - sum(int(n * x) for x in share[1:]),
                                 seed=seed * 100 + s, noise=0.5e-3 * s))
    return _finish(np.concatenate([p["points"] for p in parts]), np.concatenate([p["normals"] for p in parts]),
                   np.concatenate([p["radii"] for p in parts]))


def adaptive_blob(n=200_000, seed=0):
    """Small strongly adaptive test cloud (leaf levels spread over ~8 levels),
    the survey's probe distribution: sigma=0.25 blob, radius = 0.002 e^{6d} 2^U(0,2)."""
    rng = np.random.default_rng(seed)
    p = (rng.standard_normal((n, 3)) * 0.25).astype(np.float32)
    d = np.linalg.norm(p, axis=1)
    r = 0.002 * np.exp(6 * d) * 2 ** rng.uniform(0, 2, n)
    return _finish(p, _unit(p.astype(np.float64) + 1e-12), r)


def make(name, n=None, seed=None):
    gens = {"sphere": sphere, "gaussian_blob": gaussian_blob, "thingi_like": thingi_like, "multi_scan": multi_scan,
            "adaptive_blob": adaptive_blob}
    kw = {}
    if n is not None:
        kw["n"] = n
    if seed is not None:
        kw["seed"] = seed
    return gens[name](**kw)
