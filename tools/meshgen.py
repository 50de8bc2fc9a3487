"""Mesh generators standing in for gmsh (absent here) -- tooling, not product code.

The reference ships only the `.geo` recipes for its 3-D meshes
(reference meshes/mesh-3D-{5,10,20,40}.geo, `.gitignore:39`) and `mesh-2D-200.msh` is a
missing blob, so the benchmark meshes are regenerated here with the same geometry,
physical tags and target sizes:

  * `mesh_3d(level)`  box 0.41 x 0.41 x 2.5 minus the x-aligned cylinder r=0.05 at
    (y,z)=(0.2,0.45); size lc_cyl inside the .geo refinement box y in [0.1,0.3],
    z in [0.35,1.05] (all x), graded to lc_global outside
    (mesh-3D-20.geo:16-17,27-32); tags 101 inlet (z=0), 102 outlet (z=L),
    103 cylinder, 104 walls, 201 fluid (mesh-3D-20.geo:41-56).  Built as a graded 2-D
    triangulation of the (z,y) cross-section (DistMesh-style force equilibrium on scipy's
    Delaunay), extruded along x into prisms, each split into 3 tets with a consistent
    diagonal rule.  Mesh nodes sit exactly at the pressure probes (0.205,0.2,0.40) and
    (0.205,0.2,0.50) (reference NavierStokes.cpp:878-879).
  * `refine_2d(mesh)` red refinement of a shipped 2-D mesh with new cylinder nodes
    snapped to r=0.05 -- mesh-2D-200-equivalent from mesh-2D-100 (mesh-2D-200.geo:11-12
    halves both sizes of mesh-2D-100.geo).

Everything is deterministic (fixed seed, no wall-clock dependence).
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import Delaunay

from .msh import Mesh

# geometry of reference meshes/mesh-3D-*.geo
L3, H3, W3, R_CYL, CYL_Z, CYL_Y = 2.5, 0.41, 0.41, 0.05, 0.45, 0.2
LEVELS_3D = {5: (0.02, 0.15), 10: (0.01, 0.1), 20: (0.005, 0.05), 40: (0.0025, 0.025)}


# --------------------------------------------------------------------------- 2-D --
def _drect(p, x1, x2, y1, y2):
    d1, d2, d3, d4 = y1 - p[:, 1], -y2 + p[:, 1], x1 - p[:, 0], -x2 + p[:, 0]
    d5 = np.sqrt(d1 ** 2 + d3 ** 2)
    d6 = np.sqrt(d1 ** 2 + d4 ** 2)
    d7 = np.sqrt(d2 ** 2 + d3 ** 2)
    d8 = np.sqrt(d2 ** 2 + d4 ** 2)
    d = -np.minimum(np.minimum(np.minimum(-d1, -d2), -d3), -d4)
    ix = (d1 > 0) & (d3 > 0); d[ix] = d5[ix]
    ix = (d1 > 0) & (d4 > 0); d[ix] = d6[ix]
    ix = (d2 > 0) & (d3 > 0); d[ix] = d7[ix]
    ix = (d2 > 0) & (d4 > 0); d[ix] = d8[ix]
    return d


def _tri_quality(p, t):
    a = np.linalg.norm(p[t[:, 1]] - p[t[:, 0]], axis=1)
    b = np.linalg.norm(p[t[:, 2]] - p[t[:, 1]], axis=1)
    c = np.linalg.norm(p[t[:, 0]] - p[t[:, 2]], axis=1)
    return (b + c - a) * (c + a - b) * (a + b - c) / (a * b * c)


def cross_section(lc_cyl, lc_global, length=L3, height=H3, cz=CYL_Z, cy=CYL_Y, r=R_CYL,
                  box=(0.35, 1.05, 0.1, 0.3), grade=0.25, max_iter=600, seed=0, verbose=False):
    """Graded triangulation of [0,length] x [0,height] minus the disc, coordinates (z, y).
    Returns (points (n,2), triangles (m,3) ccw, boundary edges (k,2), edge tags (k,) in
    {'inlet','outlet','walls','cylinder'} encoded 0..3)."""
    def fd(p):
        return np.maximum(_drect(p, 0.0, length, 0.0, height), -(np.sqrt((p[:, 0] - cz) ** 2 + (p[:, 1] - cy) ** 2) - r))

    def fh(p):
        dx = np.maximum(np.maximum(box[0] - p[:, 0], p[:, 0] - box[1]), 0.0)
        dy = np.maximum(np.maximum(box[2] - p[:, 1], p[:, 1] - box[3]), 0.0)
        return np.minimum(lc_global, lc_cyl + grade * np.sqrt(dx * dx + dy * dy))

    h0 = lc_cyl
    geps = 1e-3 * h0
    deps = np.sqrt(np.finfo(float).eps) * h0
    rng = np.random.default_rng(seed)
    # initial hexagonal lattice, rejection to match the size function
    xs = np.arange(0.0, length + h0, h0)
    ys = np.arange(0.0, height + h0 * np.sqrt(3) / 2, h0 * np.sqrt(3) / 2)
    X, Y = np.meshgrid(xs, ys)
    X[1::2, :] += h0 / 2
    p = np.stack([X.ravel(), Y.ravel()], axis=1)
    p = p[fd(p) < geps]
    r0 = 1.0 / fh(p) ** 2
    p = p[rng.random(p.shape[0]) < r0 / r0.max()]
    # fixed points: rectangle corners + equally spaced circle nodes (multiple of 4, so the
    # probe points (cz -+ r, cy) are nodes)
    n_c = 4 * max(2, int(np.ceil(2 * np.pi * r / lc_cyl / 4)))
    ang = 2 * np.pi * np.arange(n_c) / n_c
    circ = np.stack([cz + r * np.cos(ang), cy + r * np.sin(ang)], axis=1)
    circ[0] = [cz + r, cy]; circ[n_c // 4] = [cz, cy + r]; circ[n_c // 2] = [cz - r, cy]; circ[3 * n_c // 4] = [cz, cy - r]
    pfix = np.concatenate([np.array([[0, 0], [length, 0], [length, height], [0, height]], dtype=float), circ])
    r_in = r * np.cos(np.pi / n_c) * (1 - 1e-9)          # inscribed radius of the circle polygon

    def in_domain(pm):
        return (_drect(pm, 0.0, length, 0.0, height) < -geps) & \
               (np.sqrt((pm[:, 0] - cz) ** 2 + (pm[:, 1] - cy) ** 2) > r_in)
    # drop generated points that coincide with fixed ones
    keep = np.ones(p.shape[0], bool)
    for f in pfix:
        keep &= np.linalg.norm(p - f, axis=1) > 0.4 * fh(f[None, :])[0]
    p = np.concatenate([pfix, p[keep]], axis=0)
    nfix = pfix.shape[0]
    pold = np.full_like(p, np.inf)
    Fscale, dtf, ttol, dptol = 1.2, 0.2, 0.1, 2e-3
    t = None
    for it in range(max_iter):
        if np.max(np.linalg.norm(p - pold, axis=1) / h0) > ttol:
            # free points must not crowd the fixed circle nodes (they would make slivers)
            rp = np.sqrt((p[:, 0] - cz) ** 2 + (p[:, 1] - cy) ** 2)
            crowd = rp < r + 0.45 * lc_cyl
            crowd[:nfix] = False
            if crowd.any():
                p = p[~crowd]
            pold = p.copy()
            t = Delaunay(p).simplices
            pmid = p[t].mean(axis=1)
            t = t[in_domain(pmid)]
            bars = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]], axis=0)
            bars = np.unique(np.sort(bars, axis=1), axis=0)
        barvec = p[bars[:, 0]] - p[bars[:, 1]]
        Lb = np.linalg.norm(barvec, axis=1)
        hb = fh(0.5 * (p[bars[:, 0]] + p[bars[:, 1]]))
        L0 = hb * Fscale * np.sqrt((Lb ** 2).sum() / (hb ** 2).sum())
        F = np.maximum(L0 - Lb, 0.0)
        Fvec = (F / Lb)[:, None] * barvec
        Ftot = np.zeros_like(p)
        np.add.at(Ftot, bars[:, 0], Fvec)
        np.add.at(Ftot, bars[:, 1], -Fvec)
        Ftot[:nfix] = 0.0
        p = p + dtf * Ftot
        d = fd(p)
        ix = d > 0
        if ix.any():
            dgx = (fd(p[ix] + [deps, 0]) - d[ix]) / deps
            dgy = (fd(p[ix] + [0, deps]) - d[ix]) / deps
            g2 = dgx ** 2 + dgy ** 2
            p[ix] -= np.stack([d[ix] * dgx / g2, d[ix] * dgy / g2], axis=1)
        move = np.max(np.linalg.norm(dtf * Ftot[d < -geps], axis=1) / h0) if (d < -geps).any() else 0.0
        if verbose and it % 50 == 0:
            print("  distmesh it %d  n=%d  move=%.2e" % (it, p.shape[0], move))
        if move < dptol:
            break
    # final triangulation
    rp = np.sqrt((p[:, 0] - cz) ** 2 + (p[:, 1] - cy) ** 2)
    crowd = rp < r + 0.45 * lc_cyl
    crowd[:nfix] = False
    p = p[~crowd]
    t = Delaunay(p).simplices
    t = t[in_domain(p[t].mean(axis=1))]
    # ---- snap boundary nodes exactly, classify boundary edges
    def boundary_edges(tt):
        e = np.concatenate([tt[:, [0, 1]], tt[:, [1, 2]], tt[:, [2, 0]]], axis=0)
        es = np.sort(e, axis=1)
        u, inv, cnt = np.unique(es, axis=0, return_inverse=True, return_counts=True)
        return u[cnt == 1]

    # remove degenerate boundary slivers (DistMesh leaves a few almost-flat triangles on straight edges)
    for _ in range(5):
        q = _tri_quality(p, t)
        be = boundary_edges(t)
        bnode = np.zeros(p.shape[0], bool)
        bnode[be.ravel()] = True
        bad = (q < 0.1) & (bnode[t].sum(axis=1) >= 2)
        if not bad.any():
            break
        t = t[~bad]
    be = boundary_edges(t)
    bn = np.unique(be.ravel())
    # classify boundary nodes by the nearest boundary curve and snap them onto it
    rr = np.sqrt((p[bn, 0] - cz) ** 2 + (p[bn, 1] - cy) ** 2)
    dist = np.stack([np.abs(p[bn, 0]), np.abs(p[bn, 0] - length), np.abs(p[bn, 1]),
                     np.abs(p[bn, 1] - height), np.abs(rr - r)], axis=1)
    which = np.argmin(dist, axis=1)
    assert np.all(dist.min(axis=1) < 0.6 * fh(p[bn])), "boundary node far from every boundary curve"
    tol = 0.3 * h0
    p[bn[which == 0], 0] = 0.0
    p[bn[which == 1], 0] = length
    p[bn[which == 2], 1] = 0.0
    p[bn[which == 3], 1] = height
    c = bn[which == 4]
    p[c] = np.array([cz, cy]) + (p[c] - [cz, cy]) * (r / rr[which == 4])[:, None]
    p[:nfix] = pfix
    # drop unused points, orient ccw
    used = np.unique(t.ravel())
    remap = -np.ones(p.shape[0], dtype=np.int64)
    remap[used] = np.arange(used.shape[0])
    p = p[used]
    t = remap[t]
    be = remap[be]
    a2 = (p[t[:, 1], 0] - p[t[:, 0], 0]) * (p[t[:, 2], 1] - p[t[:, 0], 1]) - \
         (p[t[:, 2], 0] - p[t[:, 0], 0]) * (p[t[:, 1], 1] - p[t[:, 0], 1])
    flip = a2 < 0
    t[flip] = t[flip][:, [0, 2, 1]]
    mid = 0.5 * (p[be[:, 0]] + p[be[:, 1]])
    tag = np.full(be.shape[0], 2)                       # walls
    tag[np.abs(mid[:, 0]) < 1e-9] = 0                   # inlet
    tag[np.abs(mid[:, 0] - length) < 1e-9] = 1          # outlet
    rm = np.sqrt((mid[:, 0] - cz) ** 2 + (mid[:, 1] - cy) ** 2)
    tag[rm < r + tol] = 3                               # cylinder
    return p, t.astype(np.int64), be.astype(np.int64), tag


def _hilbert_key(x, y, order=12):
    """Hilbert curve index of integer grid coordinates (vectorised)."""
    n = 1 << order
    x = x.copy(); y = y.copy()
    d = np.zeros(x.shape[0], dtype=np.int64)
    s = n >> 1
    while s > 0:
        rx = ((x & s) > 0).astype(np.int64)
        ry = ((y & s) > 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        # rotate
        m = ry == 0
        mf = m & (rx == 1)
        x[mf] = s - 1 - x[mf]
        y[mf] = s - 1 - y[mf]
        xm = x[m].copy(); x[m] = y[m]; y[m] = xm
        x &= (s - 1); y &= (s - 1)            # keep the low bits for the next level
        s >>= 1
    return d


def _locality_order_2d(p, t):
    """Sort triangles along a Hilbert curve, renumber vertices by first appearance."""
    c = p[t].mean(axis=1)
    lo = p.min(axis=0); span = (p.max(axis=0) - lo).max()
    g = np.minimum(((c - lo) / span * 4095).astype(np.int64), 4095)
    t = t[np.argsort(_hilbert_key(g[:, 0], g[:, 1]), kind="stable")]
    flat = t.ravel()
    _, first = np.unique(flat, return_index=True)
    order = flat[np.sort(first)]
    remap = np.empty(p.shape[0], dtype=np.int64)
    remap[order] = np.arange(order.shape[0])
    return p[order], remap[t], remap


def cached_cross_section(level):
    """Cross-section triangulation cached by tools/make_cross_sections.py (identical to what cross_section() returns)."""
    import os
    f = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cross_sections", "level%d.npz" % level)
    if not os.path.exists(f):
        return None
    d = np.load(f)
    return d["p2"], d["t2"], d["be2"], d["tag2"]


def mesh_3d(level=None, lc_cyl=None, lc_global=None, nx=None, verbose=False, cross=None) -> Mesh:
    """mesh-3D-<level>-equivalent (level in {5,10,20,40}) or explicit sizes."""
    if level is not None:
        lc_cyl, lc_global = LEVELS_3D[level]
    if nx is None:
        nx = int(np.ceil(W3 / lc_cyl))
        nx += nx % 2                                   # a node layer at x = W/2 = 0.205
    if cross is None and level is not None:
        cross = cached_cross_section(level)
    p2, t2, be2, tag2 = cross if cross is not None else cross_section(lc_cyl, lc_global, verbose=verbose)
    p2, t2, remap = _locality_order_2d(p2, t2)
    be2 = remap[be2]
    n2 = p2.shape[0]
    nl = nx + 1
    xs = np.linspace(0.0, W3, nl)
    xs[nx // 2] = 0.5 * W3
    # node (v, k) -> v*nl + k ; coordinates (x, y, z) = (xs[k], p2[v,1], p2[v,0])
    pts = np.empty((n2 * nl, 3))
    pts[:, 0] = np.tile(xs, n2)
    pts[:, 1] = np.repeat(p2[:, 1], nl)
    pts[:, 2] = np.repeat(p2[:, 0], nl)
    ts = np.sort(t2, axis=1)                           # a < b < c by 2-D id
    a, b, c = ts[:, 0], ts[:, 1], ts[:, 2]
    k = np.arange(nx)
    def nid(v, kk):
        return v[:, None] * nl + kk[None, :]
    a0, b0, c0 = nid(a, k), nid(b, k), nid(c, k)
    a1, b1, c1 = nid(a, k + 1), nid(b, k + 1), nid(c, k + 1)
    T1 = np.stack([a0, b0, c0, a1], axis=-1)
    T2 = np.stack([b0, c0, a1, b1], axis=-1)
    T3 = np.stack([c0, a1, b1, c1], axis=-1)
    cells = np.stack([T1, T2, T3], axis=2).reshape(-1, 4)   # column-major: triangle, layer, tet
    X = pts[cells]
    vol = np.einsum("ij,ij->i", np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), X[:, 3] - X[:, 0])
    neg = vol < 0
    cells[neg] = cells[neg][:, [0, 1, 3, 2]]
    # boundary faces
    tagmap = np.array([101, 102, 104, 103])            # inlet, outlet, walls, cylinder (3-D ids)
    es = np.sort(be2, axis=1)
    ea0, eb0 = nid(es[:, 0], k), nid(es[:, 1], k)
    ea1, eb1 = nid(es[:, 0], k + 1), nid(es[:, 1], k + 1)
    F1 = np.stack([ea0, eb0, ea1], axis=-1).reshape(-1, 3)
    F2 = np.stack([eb0, ea1, eb1], axis=-1).reshape(-1, 3)
    ftag_lat = np.repeat(tagmap[tag2], nx)
    bottom = t2 * nl
    top = t2 * nl + nx
    faces = np.concatenate([F1, F2, bottom, top], axis=0)
    ftag = np.concatenate([ftag_lat, ftag_lat, np.full(bottom.shape[0] + top.shape[0], 104)])
    return Mesh(3, pts, cells.astype(np.int32), np.full(cells.shape[0], 201, np.int32),
                faces.astype(np.int32), ftag.astype(np.int32))


def refine_2d(mesh: Mesh, cx=0.2, cy=0.2, r=R_CYL, cyl_tag=104) -> Mesh:
    """Red refinement (each triangle -> 4); midpoints of cylinder boundary edges are
    projected onto the circle."""
    assert mesh.dim == 2
    V = mesh.n_vertices
    cells = mesh.cells.astype(np.int64)
    e = np.stack([cells[:, [0, 1]], cells[:, [1, 2]], cells[:, [2, 0]]], axis=1)     # (C,3,2)
    es = np.sort(e, axis=2)
    key = es[:, :, 0] * V + es[:, :, 1]
    ukey, inv = np.unique(key.ravel(), return_inverse=True)
    mid_id = V + inv.reshape(-1, 3)
    ev = np.stack([ukey // V, ukey % V], axis=1)
    mids = 0.5 * (mesh.points[ev[:, 0]] + mesh.points[ev[:, 1]])
    # boundary faces -> two halves; snap cylinder midpoints
    fs = np.sort(mesh.faces.astype(np.int64), axis=1)
    fe_ = np.searchsorted(ukey, fs[:, 0] * V + fs[:, 1])
    cyl = fe_[mesh.face_tag == cyl_tag]
    d = mids[cyl] - [cx, cy]
    mids[cyl] = np.array([cx, cy]) + d * (r / np.linalg.norm(d, axis=1))[:, None]
    pts = np.concatenate([mesh.points, mids], axis=0)
    v0, v1, v2 = cells[:, 0], cells[:, 1], cells[:, 2]
    m01, m12, m20 = mid_id[:, 0], mid_id[:, 1], mid_id[:, 2]
    new = np.stack([np.stack([v0, m01, m20], 1), np.stack([m01, v1, m12], 1),
                    np.stack([m20, m12, v2], 1), np.stack([m01, m12, m20], 1)], axis=1).reshape(-1, 3)
    fm = V + fe_
    f = mesh.faces.astype(np.int64)
    nf = np.stack([np.stack([f[:, 0], fm], 1), np.stack([fm, f[:, 1]], 1)], axis=1).reshape(-1, 2)
    nft = np.repeat(mesh.face_tag, 2)
    return Mesh(2, pts, new.astype(np.int32), np.repeat(mesh.cell_tag, 4), nf.astype(np.int32), nft)


def mesh_quality_3d(mesh: Mesh):
    X = mesh.points[mesh.cells]
    vol = np.einsum("ij,ij->i", np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), X[:, 3] - X[:, 0]) / 6.0
    d = X[:, :, None, :] - X[:, None, :, :]
    hmax = np.sqrt((d ** 2).sum(-1)).max(axis=(1, 2))
    return vol, hmax
