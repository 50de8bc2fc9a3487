"""ORACLE (test infrastructure, not product code) -- DoF enumeration, sparsity, constraints.

PARITY UNPINNED by the reference (no tests / fixtures); follows deal.II's documented
behaviour as recorded in SURVEY.md Appendix A.2, A.4, A.5.  Integer pins that do exist
(n_u, n_p, nnz, constraint counts on the three shipped 2-D
meshes) are checked in tests/test_oracle_pins.py.

Restates, for ONE MPI rank:
  dof_handler.distribute_dofs(*fe)                         reference NavierStokes.cpp:83-84
  DoFRenumbering::component_wise(dh, {0,..,0,1})           reference NavierStokes.cpp:87-89
  DoFTools::make_sparsity_pattern(dh, bdsp, empty, true)   reference NavierStokes.cpp:256-268
  VectorTools::interpolate_boundary_values(...) x4         reference NavierStokes.cpp:229-253, 617-639
"""
import numpy as np
from . import fe_tables as fe


class DofMap:
    pass


def enumerate_dofs(mesh) -> DofMap:
    """First-touch enumeration over cells in file order: for each cell its vertices
    (dim+1 DoFs each: u_0..u_{d-1}, p), then its lines (dim DoFs each); then the stable
    component-wise renumbering to [velocity | pressure]."""
    dim = mesh.dim
    nv = dim + 1
    lines = fe.LINES[dim]
    cells = mesh.cells.astype(np.int64)
    C = cells.shape[0]
    V = mesh.n_vertices

    # unique edges, keyed by sorted vertex pair
    ev = np.sort(cells[:, lines], axis=2)                      # (C, L, 2)
    key = ev[:, :, 0] * V + ev[:, :, 1]                        # (C, L)
    ukey, first_pos, edge_of = np.unique(key.ravel(), return_index=True, return_inverse=True)
    edge_of = edge_of.reshape(C, -1)                           # cell-local line -> edge id (sorted-key order)
    E = ukey.shape[0]

    # entity visit sequence: per cell, vertices 0..d then lines 0..L-1
    nl = lines.shape[0]
    ent = np.concatenate([cells, V + edge_of], axis=1)         # (C, nv+nl) entity ids, edges offset by V
    flat = ent.ravel()
    _, first = np.unique(flat, return_index=True)              # first occurrence of each entity
    order = np.sort(first)                                     # visit positions in first-touch order
    ent_in_order = flat[order]                                 # entities in first-touch order
    is_vertex = ent_in_order < V
    ndof_ent = np.where(is_vertex, dim + 1, dim)
    start = np.concatenate([[0], np.cumsum(ndof_ent)[:-1]])    # raw (pre-renumbering) first DoF of entity
    raw_start = np.empty(V + E, dtype=np.int64)
    raw_start[ent_in_order] = start
    n_raw = int(ndof_ent.sum())

    # raw dof -> is pressure?
    is_p = np.zeros(n_raw, dtype=bool)
    vstart = raw_start[:V]
    used_vertex = np.zeros(V, dtype=bool)
    used_vertex[ent_in_order[is_vertex]] = True
    if not used_vertex.all():
        raise ValueError("mesh has vertices that belong to no cell")
    is_p[vstart + dim] = True
    # stable component-wise renumbering
    new_index = np.empty(n_raw, dtype=np.int64)
    n_u = int((~is_p).sum())
    n_p = int(is_p.sum())
    new_index[~is_p] = np.arange(n_u)
    new_index[is_p] = n_u + np.arange(n_p)

    # cell -> dofs in FESystem local order
    dpc = fe.dofs_per_cell(dim)
    cell_dofs = np.empty((C, dpc), dtype=np.int64)
    k = 0
    for v in range(nv):
        for c in range(dim + 1):
            cell_dofs[:, k] = new_index[raw_start[cells[:, v]] + c]
            k += 1
    for l in range(nl):
        for c in range(dim):
            cell_dofs[:, k] = new_index[raw_start[V + edge_of[:, l]] + c]
            k += 1

    # support points (vertices; straight-edge midpoints)
    sp = np.zeros((n_u + n_p, dim))
    pts = mesh.points
    for v in range(nv):
        for c in range(dim + 1):
            sp[cell_dofs[:, v * (dim + 1) + c]] = pts[cells[:, v]]
    base = nv * (dim + 1)
    for l, (i, j) in enumerate(lines):
        mid = 0.5 * (pts[cells[:, i]] + pts[cells[:, j]])
        for c in range(dim):
            sp[cell_dofs[:, base + l * dim + c]] = mid

    dm = DofMap()
    dm.dim = dim
    dm.n_u, dm.n_p, dm.n_dofs = n_u, n_p, n_u + n_p
    dm.cell_dofs = cell_dofs.astype(np.int32)
    dm.support_points = sp
    dm.n_edges = E
    dm.edge_of = edge_of          # (C, L) edge ids (0..E-1)
    dm.edge_vertices = np.stack([ukey // V, ukey % V], axis=1)
    comp = np.full(n_u + n_p, dim, dtype=np.int32)
    node, lcomp = fe.local_dof_layout(dim)
    comp[dm.cell_dofs.ravel()] = np.tile(lcomp, C)
    dm.component = comp
    return dm


def make_sparsity(dm: DofMap):
    """Union over cells of the full dofs_per_cell x dofs_per_cell coupling, rows sorted
    (SURVEY.md A.4).  Returns CSR (rowptr int64[N+1], col int32[nnz])."""
    N = dm.n_dofs
    cd = dm.cell_dofs.astype(np.int64)
    C, K = cd.shape
    chunks = []
    step = max(1, 2_000_000 // (K * K))
    for s in range(0, C, step):
        blk = cd[s:s + step]
        keys = (blk[:, :, None] * N + blk[:, None, :]).ravel()
        chunks.append(np.unique(keys))
        if len(chunks) > 16:
            chunks = [np.unique(np.concatenate(chunks))]
    keys = np.unique(np.concatenate(chunks))
    rows = keys // N
    cols = (keys % N).astype(np.int32)
    rowptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr, cols


def make_sparsity_fast(dm: DofMap):
    """Same pattern as make_sparsity, as the structure of G^T G for the cell-dof incidence matrix G
    (seconds instead of a minute on 5e4 tets; tests/test_oracle_pins.py checks both agree)."""
    import scipy.sparse as sp
    N = dm.n_dofs
    C, K = dm.cell_dofs.shape
    G = sp.csr_matrix((np.ones(C * K, dtype=np.float32), dm.cell_dofs.ravel().astype(np.int64),
                       np.arange(0, C * K + 1, K, dtype=np.int64)), shape=(C, N))
    P = sp.csr_matrix(G.T @ G)
    P.sort_indices()
    return P.indptr.astype(np.int64), P.indices.astype(np.int32)


def boundary_dofs(mesh, dm: DofMap, boundary_id, velocity=True, pressure=False):
    """DoFs located on boundary faces with the given id (vertices and lines of the face),
    restricted to the component mask -- the index set interpolate_boundary_values touches."""
    dim = mesh.dim
    sel = mesh.faces[mesh.face_tag == boundary_id].astype(np.int64)
    if sel.shape[0] == 0:
        return np.zeros(0, dtype=np.int64)
    V = mesh.n_vertices
    cells = mesh.cells.astype(np.int64)
    # vertex -> first DoFs:  recover from cell_dofs
    vdof = np.full((V, dim + 1), -1, dtype=np.int64)
    for v in range(dim + 1):
        vdof[cells[:, v]] = dm.cell_dofs[:, v * (dim + 1):(v + 1) * (dim + 1)]
    out = []
    fv = np.unique(sel.ravel())
    if velocity:
        out.append(vdof[fv, :dim].ravel())
    if pressure:
        out.append(vdof[fv, dim])
    if velocity:
        # lines of the boundary faces
        pairs = []
        for a in range(dim):
            for b in range(a + 1, dim):
                p = np.sort(sel[:, [a, b]], axis=1)
                pairs.append(p[:, 0] * V + p[:, 1])
        pk = np.unique(np.concatenate(pairs))
        ekey = dm.edge_vertices[:, 0] * V + dm.edge_vertices[:, 1]
        eid = np.searchsorted(ekey, pk)
        assert np.all(ekey[eid] == pk), "boundary face edge not found in the mesh"
        # edge -> dofs
        edof = np.full((dm.n_edges, dim), -1, dtype=np.int64)
        base = (dim + 1) * (dim + 1)
        for l in range(fe.LINES[dim].shape[0]):
            edof[dm.edge_of[:, l]] = dm.cell_dofs[:, base + l * dim: base + (l + 1) * dim]
        out.append(edof[eid].ravel())
    return np.unique(np.concatenate(out))


class Constraints:
    """Pure Dirichlet AffineConstraints: x[dof] = value.  First call wins (A.5)."""

    def __init__(self, n_dofs):
        self.is_c = np.zeros(n_dofs, dtype=bool)
        self.val = np.zeros(n_dofs)

    def add(self, dofs, values):
        dofs = np.asarray(dofs, dtype=np.int64)
        values = np.broadcast_to(np.asarray(values, dtype=np.float64), dofs.shape)
        new = ~self.is_c[dofs]
        self.is_c[dofs[new]] = True
        self.val[dofs[new]] = values[new]

    @property
    def dofs(self):
        return np.nonzero(self.is_c)[0]

    def distribute(self, x):
        x = x.copy()
        x[self.is_c] = self.val[self.is_c]
        return x


def build_constraints(mesh, dm, inlet_value_fn, ids, homogeneous=False):
    """inlet (101) -> walls -> cylinder velocity Dirichlet, then pressure = 0 on the outlet
    (reference NavierStokes.cpp:229-253 homogeneous / 617-639 with the inlet profile).
    inlet_value_fn(points (n,dim), comp (n,)) -> values; ids = dict(inlet, outlet, wall, cylinder)."""
    con = Constraints(dm.n_dofs)
    d_in = boundary_dofs(mesh, dm, ids["inlet"])
    if homogeneous:
        con.add(d_in, 0.0)
    else:
        con.add(d_in, inlet_value_fn(dm.support_points[d_in], dm.component[d_in]))
    con.add(boundary_dofs(mesh, dm, ids["wall"]), 0.0)
    con.add(boundary_dofs(mesh, dm, ids["cylinder"]), 0.0)
    con.add(boundary_dofs(mesh, dm, ids["outlet"], velocity=False, pressure=True), 0.0)
    return con
