"""CPU, world_size 2 over gloo: the multi-rank path's host logic -- ownership (lowest part wins), the halo plan
(grouped sends, receives landing in place in the ghost tail), a distributed SpMV over the exported local rows and
the Krylov-style all-reduce -- driven by the same structure builder the CUDA library uses
(nsb_test_build_pattern / nsb_test_halo_plan).  The GPU halo exchange itself (ncclSend/ncclRecv with the same plan)
is exercised by tools/gpu_multi.py under `gpurun --gpus N`."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

from tests.conftest import GOLDEN, PKG_DIR

LIB = os.path.join(PKG_DIR, "libnsb200.so")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _P(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _worker(rank, world, port, dim_mesh, q):
    import torch
    import torch.distributed as dist
    from oracle import assemble as asm, dofs as odofs
    from tools import meshgen, msh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if dim_mesh == 2:
            m = msh.load_npz(os.path.join(GOLDEN, "mesh-2D.npz"))
        else:
            m = meshgen.mesh_3d(lc_cyl=0.06, lc_global=0.15)
        dim = m.dim
        dm = odofs.enumerate_dofs(m)
        orp, ocol = odofs.make_sparsity(dm)
        N = dm.n_dofs
        lib = C.CDLL(LIB, mode=C.RTLD_GLOBAL)
        pts = np.ascontiguousarray(m.points, np.float64)
        cv = np.ascontiguousarray(m.cells, np.uint32)
        cd = np.ascontiguousarray(dm.cell_dofs, np.uint32)
        part = np.ascontiguousarray((np.arange(m.n_cells, dtype=np.int64) * world) // m.n_cells, np.int32)
        margs = (dim, C.c_int64(pts.shape[0]), _P(pts, C.c_double), C.c_int64(cv.shape[0]), _P(cv, C.c_uint32), _P(cd, C.c_uint32),
                 C.c_int64(dm.n_u), C.c_int64(dm.n_p), _P(part, C.c_int32), rank, world)
        n, nnz, ng, ns, nc = (C.c_int64() for _ in range(5))
        assert lib.nsb_test_build_pattern(*margs, C.byref(n), C.byref(nnz), None, None, None, C.byref(ng), C.byref(ns), C.byref(nc)) == 0
        rp = np.empty(n.value + 1, np.int64)
        col = np.empty(nnz.value, np.uint32)
        gid = np.empty(n.value, np.int64)
        assert lib.nsb_test_build_pattern(*margs, C.byref(n), C.byref(nnz), _P(rp, C.c_int64), _P(col, C.c_uint32), _P(gid, C.c_int64),
                                          C.byref(ng), C.byref(ns), C.byref(nc)) == 0
        nghost, npeers = C.c_int64(), C.c_int32()
        assert lib.nsb_test_halo_plan(*margs, C.byref(nghost), None, C.byref(npeers), None, None, None, None, None, None, None) == 0
        K = npeers.value
        ghost = np.empty(nghost.value, np.int64)
        peers = np.empty(K, np.int32)
        su_ptr, sp_ptr = np.empty(K + 1, np.int64), np.empty(K + 1, np.int64)
        ru, rpc = np.empty(K, np.int64), np.empty(K, np.int64)
        su = np.empty(ns.value, np.int64)
        sp_ = np.empty(ns.value, np.int64)
        assert lib.nsb_test_halo_plan(*margs, C.byref(nghost), _P(ghost, C.c_int64), C.byref(npeers), _P(peers, C.c_int32),
                                      _P(su_ptr, C.c_int64), _P(su, C.c_int64), _P(sp_ptr, C.c_int64), _P(sp_, C.c_int64),
                                      _P(ru, C.c_int64), _P(rpc, C.c_int64)) == 0
        assert K == world - 1 and nghost.value > 0
        # ---- halo exchange of a seeded global vector
        x = np.random.default_rng(99).uniform(-1, 1, N)
        n_own = n.value
        x_loc = np.full(n_own + nghost.value, np.nan)
        x_loc[:n_own] = x[gid]
        n_gu = int(ru.sum())
        reqs, bufs = [], []
        off_u, off_p = 0, n_gu
        for k in range(K):
            peer = int(peers[k])
            for lst, lo, hi in ((su, su_ptr[k], su_ptr[k + 1]), (sp_, sp_ptr[k], sp_ptr[k + 1])):
                t = torch.from_numpy(np.ascontiguousarray(x[lst[lo:hi]]))
                bufs.append(t)
                reqs.append(dist.isend(t, peer))
            tu, tp = torch.empty(int(ru[k]), dtype=torch.float64), torch.empty(int(rpc[k]), dtype=torch.float64)
            reqs.append(dist.irecv(tu, peer))
            reqs.append(dist.irecv(tp, peer))
            bufs.append((tu, off_u, tp, off_p))
            off_u += int(ru[k])
            off_p += int(rpc[k])
        for r in reqs:
            r.wait()
        for b in bufs:
            if isinstance(b, tuple):
                tu, ou, tp, op = b
                x_loc[n_own + ou:n_own + ou + tu.numel()] = tu.numpy()
                x_loc[n_own + op:n_own + op + tp.numel()] = tp.numpy()
        assert np.array_equal(x_loc[n_own:], x[ghost]), "ghost entries did not land in place"
        # ---- distributed SpMV over the local rows (values from the oracle's global matrix)
        vals = np.random.default_rng(5).uniform(-1, 1, ocol.size)
        g2l = np.full(N, -1, np.int64)
        g2l[gid] = np.arange(n_own)
        g2l[ghost] = n_own + np.arange(nghost.value)
        assert np.all(g2l[col.astype(np.int64)] >= 0), "a column of an owned row is neither owned nor ghost"
        y = np.zeros(n_own)
        for k in range(n_own):
            g = gid[k]
            assert np.array_equal(col[rp[k]:rp[k + 1]].astype(np.int64), ocol[orp[g]:orp[g + 1]].astype(np.int64))
            y[k] = vals[orp[g]:orp[g + 1]] @ x_loc[g2l[col[rp[k]:rp[k + 1]].astype(np.int64)]]
        import scipy.sparse as sps
        yref = sps.csr_matrix((vals, ocol, orp), shape=(N, N)) @ x
        assert np.abs(y - yref[gid]).max() < 1e-12
        # ---- all-reduce of an inner product over the owned entries = the global inner product
        t = torch.tensor([float(x_loc[:n_own] @ x_loc[:n_own]), float(n_own)], dtype=torch.float64)
        dist.all_reduce(t)
        assert abs(t[0].item() - float(x @ x)) < 1e-9 and int(t[1].item()) == N
        q.put((rank, "ok"))
    except Exception as e:   # noqa: BLE001
        import traceback
        q.put((rank, "FAIL: " + repr(e) + "\n" + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dim_mesh", [2, 3])
def test_two_ranks_halo_spmv_allreduce(dim_mesh):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dim_mesh, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", "rank %d: %s" % (rank, msg)
