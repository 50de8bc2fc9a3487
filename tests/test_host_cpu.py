"""CPU: the C-ABI libraries load and export every declared symbol; the C++ host mirror reproduces the
oracle's integer setup bit-exactly (DoF maps, sparsity, constraints); the node-block structure builder
of the CUDA library exports the same scalar pattern, also when the mesh is split over ranks."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import dofs as odofs, postprocess as pp
from tests.conftest import PKG_DIR, ROOT
from tools import msh

LIB = os.path.join(PKG_DIR, "libnsb200.so")
HOSTLIB = os.path.join(PKG_DIR, "libnsbhost.so")


def _declared(header, prefix):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(%s_[a-z0-9_]+)\s*\(" % prefix, txt)))


def test_libraries_export_all_declared_symbols():
    assert os.path.exists(LIB), "libnsb200.so missing: run __graft_entry__.build()"
    lib = C.CDLL(LIB, mode=C.RTLD_GLOBAL)
    names = _declared("nsb200.h", "nsb")
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
    host = C.CDLL(HOSTLIB)
    for n in _declared("nsb200_host.h", "nsh") + _declared("nsb200_host.h", "nshd"):
        assert hasattr(host, n), n


def test_no_cuda_device_fails_loudly(nsb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(nsb.NsbError):
        nsb.Device(2)


def HostSetup(path, dim):
    from tests.conftest import load_nsb
    nsb = load_nsb()
    try:
        return nsb.HostSetup(path, dim)
    except nsb.NsbError as e:
        raise RuntimeError(str(e))


@pytest.mark.parametrize("name", ["mesh-2D", "mesh-2D-40"])
def test_host_dofs_sparsity_constraints_bit_exact(golden_mesh, msh_file, name):
    m = golden_mesh(name)
    hs = HostSetup(msh_file(name), 2)
    dm = odofs.enumerate_dofs(m)
    assert (hs.n_u, hs.n_p, hs.n_cells) == (dm.n_u, dm.n_p, m.n_cells)
    assert np.array_equal(hs.cell_dofs(), dm.cell_dofs.astype(np.uint32))
    p, c = hs.support_points()
    assert np.array_equal(p, dm.support_points) and np.array_equal(c, dm.component.astype(np.uint8))
    rp, col = hs.pattern()
    orp, ocol = odofs.make_sparsity(dm)
    assert np.array_equal(rp, orp) and np.array_equal(col.astype(np.int64), ocol.astype(np.int64))
    ids = pp.boundary_ids(2)
    for case, t, hom in (("2D-2", 1.0, False), ("2D-3", 3.0, False), ("2D-1", 0.5, True)):
        tc = pp.TEST_CASES[case]
        con = odofs.build_constraints(m, dm, pp.inlet_profile(2, tc["U_m"], tc["time_dep"], tc["T_ramp"], t), ids, homogeneous=hom)
        d, v = hs.constraints(case, t, hom)
        assert np.array_equal(d.astype(np.int64), con.dofs)
        assert np.array_equal(v, con.val[con.dofs])


def test_host_3d_setup_matches_oracle(small_3d_mesh, tmp_path):
    m = small_3d_mesh
    path = str(tmp_path / "m3.bin")
    msh.write_bin(path, m)
    hs = HostSetup(path, 3)
    dm = odofs.enumerate_dofs(m)
    assert np.array_equal(hs.cell_dofs(), dm.cell_dofs.astype(np.uint32))
    ids = pp.boundary_ids(3)
    con = odofs.build_constraints(m, dm, pp.inlet_profile(3, 2.25, False, 4.0, 2.0), ids)
    d, v = hs.constraints("3D-2Z", 2.0, False)
    assert np.array_equal(d.astype(np.int64), con.dofs) and np.array_equal(v, con.val[con.dofs])
    # the ASCII writer / reader round trip gives the same mesh
    path2 = str(tmp_path / "m3.msh")
    msh.write_msh(path2, m)
    hs2 = HostSetup(path2, 3)
    p, c = hs2.mesh()
    assert np.array_equal(p, m.points) and np.array_equal(c, m.cells.astype(np.uint32))
    assert np.array_equal(hs2.cell_dofs(), hs.cell_dofs())


def test_msh_reader_prepass(golden_mesh, tmp_path):
    """$ParametricNodes blocks and CRLF line ends are accepted (reference cpp:16-51)."""
    m = golden_mesh("mesh-2D")
    path = str(tmp_path / "a.msh")
    msh.write_msh(path, m)
    txt = open(path).read()
    nodes = txt[txt.index("$Nodes"):txt.index("$EndNodes")]
    lines = nodes.split("\n")
    par = ["$ParametricNodes", lines[1]] + [ln + " 1 7 0.25" for ln in lines[2:] if ln]
    txt2 = txt.replace(nodes + "$EndNodes", "\n".join(par) + "\n$EndParametricNodes").replace("\n", "\r\n")
    path2 = str(tmp_path / "b.msh")
    open(path2, "w", newline="").write(txt2)
    a, b = HostSetup(path, 2), HostSetup(path2, 2)
    assert np.array_equal(a.cell_dofs(), b.cell_dofs())
    assert np.array_equal(a.mesh()[0], b.mesh()[0])
    m2 = msh.read_msh(path2)
    assert np.array_equal(m2.cells, m.cells) and np.allclose(m2.points, m.points, rtol=0, atol=0)
    with pytest.raises(RuntimeError, match="Could not open mesh file"):
        HostSetup(str(tmp_path / "missing.msh"), 2)


def _build_pattern(lib, dim, m, dm, part, rank, nranks):
    pts = np.ascontiguousarray(m.points, np.float64)
    cv = np.ascontiguousarray(m.cells, np.uint32)
    cd = np.ascontiguousarray(dm.cell_dofs, np.uint32)
    pa = None if part is None else np.ascontiguousarray(part, np.int32)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    n, nnz, ng, ns, nc = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    args = (dim, C.c_int64(pts.shape[0]), P(pts, C.c_double), C.c_int64(cv.shape[0]), P(cv, C.c_uint32), P(cd, C.c_uint32),
            C.c_int64(dm.n_u), C.c_int64(dm.n_p), P(pa, C.c_int32) if pa is not None else None, rank, nranks)
    assert lib.nsb_test_build_pattern(*args, C.byref(n), C.byref(nnz), None, None, None, C.byref(ng), C.byref(ns), C.byref(nc)) == 0
    rp = np.empty(n.value + 1, np.int64)
    col = np.empty(nnz.value, np.uint32)
    gid = np.empty(n.value, np.int64)
    assert lib.nsb_test_build_pattern(*args, C.byref(n), C.byref(nnz), P(rp, C.c_int64), P(col, C.c_uint32), P(gid, C.c_int64),
                                      C.byref(ng), C.byref(ns), C.byref(nc)) == 0
    return rp, col, gid, ng.value, ns.value, nc.value


@pytest.mark.parametrize("which", ["2d", "3d"])
def test_device_structure_pattern_bit_exact_and_partition_invariant(golden_mesh, small_3d_mesh, which):
    lib = C.CDLL(LIB, mode=C.RTLD_GLOBAL)
    m = golden_mesh("mesh-2D") if which == "2d" else small_3d_mesh
    dim = m.dim
    dm = odofs.enumerate_dofs(m)
    orp, ocol = odofs.make_sparsity(dm)
    rp, col, gid, ng, ns, nc = _build_pattern(lib, dim, m, dm, None, 0, 1)
    assert np.array_equal(gid, np.arange(dm.n_dofs)) and ng == 0 and ns == 0 and nc == m.n_cells
    assert np.array_equal(rp, orp) and np.array_equal(col.astype(np.int64), ocol.astype(np.int64))
    # split into R contiguous chunks: rows are partitioned, every row keeps exactly its global pattern
    for R in (2, 4):
        part = (np.arange(m.n_cells) * R) // m.n_cells
        seen = np.zeros(dm.n_dofs, int)
        ghosts, sends = 0, 0
        for r in range(R):
            rp, col, gid, ng, ns, nc = _build_pattern(lib, dim, m, dm, part, r, R)
            seen[gid] += 1
            for k in range(0, gid.size, max(1, gid.size // 400)):
                g = gid[k]
                assert np.array_equal(col[rp[k]:rp[k + 1]].astype(np.int64), ocol[orp[g]:orp[g + 1]].astype(np.int64))
            assert np.array_equal(np.diff(rp), np.diff(orp)[gid])
            ghosts += ng
            sends += ns
        assert np.all(seen == 1)
        assert ghosts == sends and ghosts > 0          # every ghost DoF is sent by exactly one owner


def test_hessenberg_eigenvalues():
    lib = C.CDLL(LIB, mode=C.RTLD_GLOBAL)
    rng = np.random.default_rng(0)
    for n in (1, 2, 5, 9, 16, 33):
        A = np.triu(rng.standard_normal((n, n)), -1)
        wr, wi = np.zeros(n), np.zeros(n)
        P = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        assert lib.nsb_test_hessenberg_eigs(n, P(A), P(wr), P(wi)) == 0
        assert np.abs(np.sort_complex(np.linalg.eigvals(A)) - np.sort_complex(wr + 1j * wi)).max() < 1e-11


@pytest.mark.parametrize("which", ["2d", "3d"])
def test_spmv_tile_plan_invariants(golden_mesh, small_3d_mesh, which):
    """The host-built plans the SpMV kernels and the streamed velocity operator consume (tile ranges, unique neighbour
    lists, 16-bit positions, interior / boundary split; per tile the padded block count, the block metadata words and the
    node-aligned split among the consumer warps) pass the independent C++ verifiers on one rank and on every rank of 2-
    and 3-way partitions; a single rank has no boundary tile, partitioned ranks have some."""
    lib = C.CDLL(LIB, mode=C.RTLD_GLOBAL)
    m = golden_mesh("mesh-2D") if which == "2d" else small_3d_mesh
    dm = odofs.enumerate_dofs(m)
    pts = np.ascontiguousarray(m.points, np.float64)
    cv = np.ascontiguousarray(m.cells, np.uint32)
    cd = np.ascontiguousarray(dm.cell_dofs, np.uint32)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    for nranks in (1, 2, 3):
        part = (np.arange(m.n_cells, dtype=np.int64) * nranks // m.n_cells).astype(np.int32)
        for rank in range(nranks):
            out = np.zeros(5, np.int64)
            rc = lib.nsb_test_tile_plan(m.dim, C.c_int64(pts.shape[0]), P(pts, C.c_double), C.c_int64(cv.shape[0]), P(cv, C.c_uint32),
                                        P(cd, C.c_uint32), C.c_int64(dm.n_u), C.c_int64(dm.n_p),
                                        P(part, C.c_int32) if nranks > 1 else None, rank, nranks, P(out, C.c_int64))
            assert rc == 0
            tiles, n_int, n_bnd, padded_blocks, bad = out
            assert bad == 0, (nranks, rank, out)
            assert tiles > 0 and n_int + n_bnd == tiles
            assert (n_bnd == 0) if nranks == 1 else (n_bnd > 0)
            assert padded_blocks % 32 == 0 and padded_blocks > 0      # streamed operator: blocks rounded up to 32 per tile


@pytest.mark.parametrize("which", ["2d", "3d"])
def test_coarse_level_is_the_p1_prolongation(golden_mesh, small_3d_mesh, which):
    """Host structure of the two-level velocity cycle (structure.cpp build_coarse) on 1, 2 and 3 ranks: the end vertices of
    every owned node give exactly the P1 -> P2 prolongation (vertex nodes weight 1, line nodes 1/2 + 1/2), the line-node lists
    of the owned vertices are its transpose pattern, and the coarse neighbour lists are the P1 stencil."""
    import scipy.sparse as sp
    lib = C.CDLL(LIB, mode=C.RTLD_GLOBAL)
    m = golden_mesh("mesh-2D") if which == "2d" else small_3d_mesh
    dim, nv = m.dim, m.dim + 1
    dm = odofs.enumerate_dofs(m)
    pts = np.ascontiguousarray(m.points, np.float64)
    cv = np.ascontiguousarray(m.cells, np.uint32)
    cd = np.ascontiguousarray(dm.cell_dofs, np.uint32)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    # reference prolongation on NODES (global node id = velocity dof // dim, global vertex id = pressure dof - n_u)
    lines = [(0, 1), (1, 2), (2, 0)] if dim == 2 else [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)]
    cdl = dm.cell_dofs.astype(np.int64)
    pid = cdl[:, [v * (dim + 1) + dim for v in range(nv)]] - dm.n_u
    rows, cols, vals = [], [], []
    for v in range(nv):
        rows.append(cdl[:, v * (dim + 1)] // dim); cols.append(pid[:, v]); vals.append(np.ones(len(cdl)))
    for l, (i, j) in enumerate(lines):
        r = cdl[:, nv * (dim + 1) + dim * l] // dim
        rows += [r, r]; cols += [pid[:, i], pid[:, j]]; vals += [np.full(len(cdl), 0.5)] * 2
    rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    nn, npv = dm.n_u // dim, dm.n_p
    _, idx = np.unique(rows * npv + cols, return_index=True)
    Pref = sp.csr_matrix((vals[idx], (rows[idx], cols[idx])), shape=(nn, npv))
    p1 = sp.csr_matrix((np.ones(pid.size * nv), (np.repeat(pid, nv, axis=1).ravel(), np.tile(pid, (1, nv)).ravel())), shape=(npv, npv))
    p1.sum_duplicates(); p1.sort_indices()
    for nranks in (1, 2, 3):
        part = (np.arange(m.n_cells, dtype=np.int64) * nranks // m.n_cells).astype(np.int32)
        seen_nodes, seen_vert = 0, 0
        for rank in range(nranks):
            args = (m.dim, C.c_int64(pts.shape[0]), P(pts, C.c_double), C.c_int64(cv.shape[0]), P(cv, C.c_uint32), P(cd, C.c_uint32),
                    C.c_int64(dm.n_u), C.c_int64(dm.n_p), P(part, C.c_int32) if nranks > 1 else None, rank, nranks)
            sz = np.zeros(4, np.int64)
            assert lib.nsb_test_coarse_level(*args, P(sz, C.c_int64), None, None, None, None, None, None) == 0
            ng, ends, ve = np.empty(sz[0], np.int64), np.empty((sz[0], 2), np.int64), np.empty((sz[1], 2), np.int64)
            vg, cp, cg = np.empty(sz[2], np.int64), np.empty(sz[2] + 1, np.int64), np.empty(sz[3], np.int64)
            assert lib.nsb_test_coarse_level(*args, P(sz, C.c_int64), P(ng, C.c_int64), P(ends, C.c_int64), P(ve, C.c_int64), P(vg, C.c_int64),
                                             P(cp, C.c_int64), P(cg, C.c_int64)) == 0
            Pl = sp.csr_matrix((np.full(2 * len(ng), 0.5), (np.repeat(ng, 2), ends.ravel())), shape=(nn, npv))
            Pl.sum_duplicates()
            sel = sp.csr_matrix((np.ones(len(ng)), (ng, ng)), shape=(nn, nn))
            assert abs(Pl - sel @ Pref).max() == 0
            # line nodes at owned vertices = entries of weight 1/2 in the columns of the owned vertices
            T = Pref.tocsc()[:, vg].tocoo()
            want = {(int(vg[c_]), int(r_)) for r_, c_, w_ in zip(T.row, T.col, T.data) if w_ == 0.5}
            assert want == {(int(a_), int(b_)) for a_, b_ in ve}
            for k, v_ in enumerate(vg):
                assert np.array_equal(np.sort(cg[cp[k]:cp[k + 1]]), p1.indices[p1.indptr[v_]:p1.indptr[v_ + 1]])
            seen_nodes += len(ng); seen_vert += len(vg)
        assert seen_nodes == nn and seen_vert == npv


@pytest.mark.parametrize("which", ["2d", "3d"])
def test_metis_partition_and_structure_on_it(golden_mesh, small_3d_mesh, msh_file, tmp_path, which):
    """partition_cells(method 1) = METIS on the face-dual graph (GridTools::partition_triangulation, reference cpp:56): every
    cell gets a rank, parts are balanced, the face cut beats contiguous chunks of an unordered cell list -- and the device
    structure builder, the SpMV tile plans and the streamed-operator plans hold their invariants on such a (non-contiguous)
    partition: rows partitioned exactly, every row keeps its global pattern, ghosts = sends."""
    from tools import msh as mshmod
    import nsb200 as nsb
    lib = C.CDLL(LIB, mode=C.RTLD_GLOBAL)
    if which == "2d":
        m, path = golden_mesh("mesh-2D"), msh_file("mesh-2D")
    else:
        m, path = small_3d_mesh, str(tmp_path / "m3.bin")
        mshmod.write_bin(path, m)
    hs = nsb.HostSetup(path, m.dim)
    dm = odofs.enumerate_dofs(m)
    orp, ocol = odofs.make_sparsity(dm)
    nv = m.dim + 1
    # faces shared by two cells
    faces = {}
    lv = [[0, 1], [1, 2], [2, 0]] if m.dim == 2 else [[0, 1, 2], [1, 0, 3], [0, 2, 3], [2, 1, 3]]
    for f in lv:
        for c, key in enumerate(map(tuple, np.sort(m.cells[:, f], axis=1))):
            faces.setdefault(key, []).append(c)
    pairs = np.array([v for v in faces.values() if len(v) == 2])

    def cut(part):
        return int(np.count_nonzero(part[pairs[:, 0]] != part[pairs[:, 1]]))

    pts = np.ascontiguousarray(m.points, np.float64)
    cv = np.ascontiguousarray(m.cells, np.uint32)
    cd = np.ascontiguousarray(dm.cell_dofs, np.uint32)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    for R in (2, 3, 8):
        chunks, metis = hs.partition(R, 0), hs.partition(R, 1)
        assert np.array_equal(chunks, (np.arange(m.n_cells, dtype=np.int64) * R // m.n_cells).astype(np.int32))
        assert metis.min() == 0 and metis.max() == R - 1
        sizes = np.bincount(metis, minlength=R)
        assert sizes.max() <= 1.06 * m.n_cells / R + 1, sizes
        assert cut(metis) < cut(np.random.default_rng(0).permutation(chunks))
        if R == 8:
            continue
        seen = np.zeros(dm.n_dofs, int)
        ghosts = sends = 0
        for r in range(R):
            rp, col, gid, ng, ns, nc = _build_pattern(lib, m.dim, m, dm, metis, r, R)
            seen[gid] += 1
            assert np.array_equal(np.diff(rp), np.diff(orp)[gid])
            for k in range(0, gid.size, max(1, gid.size // 200)):
                g = gid[k]
                assert np.array_equal(col[rp[k]:rp[k + 1]].astype(np.int64), ocol[orp[g]:orp[g + 1]].astype(np.int64))
            ghosts += ng; sends += ns
            out = np.zeros(5, np.int64)
            assert lib.nsb_test_tile_plan(m.dim, C.c_int64(pts.shape[0]), P(pts, C.c_double), C.c_int64(cv.shape[0]), P(cv, C.c_uint32),
                                          P(cd, C.c_uint32), C.c_int64(dm.n_u), C.c_int64(dm.n_p), P(np.ascontiguousarray(metis, np.int32), C.c_int32),
                                          r, R, P(out, C.c_int64)) == 0
            assert out[4] == 0 and out[2] > 0, out
        assert np.all(seen == 1) and ghosts == sends > 0
    hs.close()


@pytest.mark.parametrize("which", ["2d", "3d"])
def test_coarse_level_references_are_covered_by_the_vertex_halo(golden_mesh, small_3d_mesh, which):
    """Several ranks: every coarse vertex a rank's two-level cycle touches -- the end vertices of its owned nodes
    (prolongation) and the P1 neighbours of its owned vertices (coarse operator) -- is an owned vertex or one of the
    pressure ghosts the halo plan delivers (the coarse vectors travel with the pressure-halo plan, dim components per
    vertex), for contiguous chunks and for a METIS partition."""
    import nsb200 as nsb
    from tools import msh as mshmod
    import tempfile
    lib = C.CDLL(LIB, mode=C.RTLD_GLOBAL)
    m = golden_mesh("mesh-2D") if which == "2d" else small_3d_mesh
    dm = odofs.enumerate_dofs(m)
    pts = np.ascontiguousarray(m.points, np.float64)
    cv = np.ascontiguousarray(m.cells, np.uint32)
    cd = np.ascontiguousarray(dm.cell_dofs, np.uint32)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "m.bin")
        mshmod.write_bin(path, m)
        hs = nsb.HostSetup(path, m.dim)
        parts = {("chunks", R): hs.partition(R, 0) for R in (2, 3)}
        parts[("metis", 3)] = hs.partition(3, 1)
        hs.close()
    for (kind, R), part in parts.items():
        part = np.ascontiguousarray(part, np.int32)
        for rank in range(R):
            args = (m.dim, C.c_int64(pts.shape[0]), P(pts, C.c_double), C.c_int64(cv.shape[0]), P(cv, C.c_uint32), P(cd, C.c_uint32),
                    C.c_int64(dm.n_u), C.c_int64(dm.n_p), P(part, C.c_int32), rank, R)
            sz = np.zeros(4, np.int64)
            assert lib.nsb_test_coarse_level(*args, P(sz, C.c_int64), None, None, None, None, None, None) == 0
            ng, ends, ve = np.empty(sz[0], np.int64), np.empty((sz[0], 2), np.int64), np.empty((sz[1], 2), np.int64)
            vg, cp, cg = np.empty(sz[2], np.int64), np.empty(sz[2] + 1, np.int64), np.empty(sz[3], np.int64)
            assert lib.nsb_test_coarse_level(*args, P(sz, C.c_int64), P(ng, C.c_int64), P(ends, C.c_int64), P(ve, C.c_int64), P(vg, C.c_int64),
                                             P(cp, C.c_int64), P(cg, C.c_int64)) == 0
            nghost, npeers = C.c_int64(), C.c_int32()
            assert lib.nsb_test_halo_plan(*args, C.byref(nghost), None, C.byref(npeers), None, None, None, None, None, None, None) == 0
            ghost = np.empty(nghost.value, np.int64)
            K = npeers.value
            peers = np.empty(K, np.int32)
            su_ptr, sp_ptr, ru, rpc = np.empty(K + 1, np.int64), np.empty(K + 1, np.int64), np.empty(K, np.int64), np.empty(K, np.int64)
            assert lib.nsb_test_halo_plan(*args, C.byref(nghost), P(ghost, C.c_int64), C.byref(npeers), P(peers, C.c_int32), P(su_ptr, C.c_int64), None,
                                          P(sp_ptr, C.c_int64), None, P(ru, C.c_int64), P(rpc, C.c_int64)) == 0
            ghost_vertices = ghost[ghost >= dm.n_u] - dm.n_u              # pressure ghosts = ghost vertices
            assert ghost_vertices.size == int(rpc.sum())
            have = set(vg.tolist()) | set(ghost_vertices.tolist())
            need = set(ends.ravel().tolist()) | set(cg.tolist())
            assert need <= have, (kind, R, rank, len(need - have))
            # line nodes listed at owned vertices are local nodes (owned or velocity ghosts)
            ghost_nodes = np.unique(ghost[ghost < dm.n_u] // m.dim)
            assert set(ve[:, 1].tolist()) <= set(ng.tolist()) | set(ghost_nodes.tolist())
