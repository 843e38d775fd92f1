"""GPU parity of the block products (transpose, SpGEMM, row scaling, SpMV) against the oracle's
PETSc-semantics restatement: bit-exact pattern (symbolic zeros kept, sorted columns), values 1e-12."""

import numpy as np
import pytest
import scipy.sparse as sp

from fenicsx_ii_b200 import Circle, DeviceCSR, PointwiseTrace, create_interpolation_matrix
from fenicsx_ii_b200 import synthetic as sy
from fenicsx_ii_b200.assembly import restrict_columns, restrict_rows, restrict_rows_and_columns
from fenicsx_ii_b200.distributed import quadrature_mass_diagonal
from fenicsx_ii_b200.mesh import functionspace
from oracle import pi_oracle as orc

pytestmark = pytest.mark.gpu


def as_host(M):
    M = M.tocsr()
    M.sort_indices()
    return M.indptr.astype(np.int64), M.indices.astype(np.int32), M.data.astype(np.float64)


def check(dev: DeviceCSR, ref, rtol=1e-12):
    ip, ij, iv = dev.to_host()
    np.testing.assert_array_equal(ip, ref[0])
    np.testing.assert_array_equal(ij, ref[1])
    scale = np.abs(ref[2]).max() if ref[2].size else 1.0
    np.testing.assert_allclose(iv, ref[2], rtol=rtol, atol=rtol * scale)


def random_csr(m, n, density, seed, heavy_rows=0, heavy_density=0.5):
    rng = np.random.default_rng(seed)
    M = sp.random(m, n, density=density, random_state=seed, format="lil")
    for r in rng.choice(m, size=heavy_rows, replace=False):
        cols = rng.choice(n, size=int(heavy_density * n), replace=False)
        M[r, cols] = rng.normal(size=cols.size)
    return M.tocsr()


@pytest.mark.parametrize("m,k,n,density,heavy", [(50, 40, 60, 0.1, 0), (300, 200, 5000, 0.02, 6), (40, 3000, 20000, 0.01, 3),
                                                 (1, 1, 1, 1.0, 0), (20, 30, 10, 0.0, 0)])
def test_spgemm_random(cuda, m, k, n, density, heavy):
    """Rows of every size bin: tiny rows, rows of a few thousand products, and rows whose product
    bound exceeds the largest shared-memory table (long-row path)."""
    A = random_csr(m, k, density, 1, heavy_rows=min(heavy, m))
    B = random_csr(k, n, density, 2, heavy_rows=min(heavy, k), heavy_density=0.3)
    a, b = as_host(A), as_host(B)
    dA = DeviceCSR.from_host(*a, A.shape)
    dB = DeviceCSR.from_host(*b, B.shape)
    C, ip = dA.matmult(dB, return_ip=True)
    ref = orc.spgemm(*a, *b, n)
    assert ip == ref[3]
    check(C, ref[:3])
    # row window (the V-dof row shard of the multi-GPU product)
    if m >= 4:
        r0, r1 = m // 4, 3 * m // 4
        Cw = dA.matmult(dB, r0, r1)
        sub = A[r0:r1]
        refw = orc.spgemm(*as_host(sub), *b, n)
        check(Cw, refw[:3])


def sparse_rows(m, n, mean_len, seed, heavy_rows=0, heavy_len=0):
    """CSR with about `mean_len` entries per row (Π-like short rows) and a few rows of `heavy_len` entries."""
    rng = np.random.default_rng(seed)
    lens = np.minimum(rng.poisson(mean_len, size=m), n)
    if heavy_rows:
        lens[rng.choice(m, size=min(heavy_rows, m), replace=False)] = min(heavy_len, n)
    indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    indices = np.concatenate([np.sort(rng.choice(n, size=k, replace=False)) for k in lens] + [np.zeros(0, dtype=np.int64)])
    data = rng.normal(size=indices.shape[0])
    return sp.csr_matrix((data, indices.astype(np.int32), indptr), shape=(m, n))


# (m, k, n, mean |A_i|, mean |B_k|, heavy rows of A, their length, heavy rows of B, their length)
LEAN_CASES = [
    (400, 300, 4000, 6, 2, 0, 0, 0, 0),          # G = 4
    (400, 300, 4000, 10, 7, 4, 200, 0, 0),       # G = 8, rows of a few thousand products
    (500, 800, 6000, 8, 14, 6, 400, 5, 300),     # G = 16 (the Π shape), rows of B longer than G
    (300, 600, 9000, 12, 20, 4, 500, 3, 2000),   # G = 32, numeric rows beyond the 2048-slot bin (block kernels)
    (60, 3000, 20000, 5, 20, 3, 1500, 0, 0),     # symbolic escalation: 1024 -> 8192 -> 32768-slot tables
    (12, 3000, 60000, 5, 20, 2, 3000, 0, 0),     # ... -> global-memory tables (> 16384 distinct columns)
    (700, 50, 64, 20, 12, 0, 0, 0, 0),           # heavy compression: every row of C is (nearly) dense in 64 columns
    (33, 1, 5, 1, 3, 0, 0, 0, 0),                # single row of B
]


@pytest.mark.parametrize("case", LEAN_CASES)
def test_spgemm_short_rows_of_b(cuda, case):
    """The lean warp kernel (rows of B of at most 32 entries on average): every sub-group width, the register sort of
    every bin, rows of B longer than a sub-group, and the symbolic escalation chain — pattern bit-exact, values 1e-12,
    fused row scaling bitwise equal to scaling first."""
    import torch

    m, k, n, la, lb, ha, hla, hb, hlb = case
    A = sparse_rows(m, k, la, 11, ha, hla)
    B = sparse_rows(k, n, lb, 12, hb, hlb)
    a, b = as_host(A), as_host(B)
    dA = DeviceCSR.from_host(*a, A.shape).validate()
    dB = DeviceCSR.from_host(*b, B.shape).validate()
    C, ip = dA.matmult(dB, return_ip=True)
    ref = orc.spgemm(*a, *b, n)
    assert ip == ref[3]
    check(C, ref[:3])
    # A · diag(d) · B: fused == scale_rows first (bitwise), == oracle on the scaled matrix
    d = np.random.default_rng(3).uniform(0.5, 2.0, size=k)
    Cs = dA.matmult(dB, row_scale=torch.from_numpy(d).to("cuda"))
    Cu = dA.matmult(dB.copy().scale_rows(d))
    np.testing.assert_array_equal(Cs.to_host()[1], Cu.to_host()[1])
    assert Cs.to_host()[2].tobytes() == Cu.to_host()[2].tobytes()
    bs = (b[0], b[1], b[2] * np.repeat(d, np.diff(b[0])))
    check(Cs, orc.spgemm(*a, *bs, n)[:3])
    if m >= 4:
        r0, r1 = m // 4, 3 * m // 4
        refw = orc.spgemm(*as_host(A[r0:r1]), *b, n)
        check(dA.matmult(dB, r0, r1), refw[:3])


def test_wrap_device_arrays_and_validate(cuda):
    import torch

    from fenicsx_ii_b200._lib import FxiiError

    A = sparse_rows(50, 40, 5, 1)
    a = as_host(A)
    t = [torch.from_numpy(x).to("cuda") for x in a]
    W = DeviceCSR.wrap_device_arrays(*t, A.shape).validate()
    for x, y in zip(W.to_host(), a):
        np.testing.assert_array_equal(x, y)
    check(W.transpose(), orc.csr_transpose(*a, 40), rtol=0)
    S = W.row_slice(10, 30)
    for x, y in zip(S.to_host(), as_host(A[10:30])):
        np.testing.assert_array_equal(x, y)
    del W, S  # handles go, the torch tensors stay valid
    assert int(t[1].sum().item()) == int(a[1].sum())
    # duplicate / unsorted / out-of-range columns are rejected
    bad = a[1].copy()
    r = int(np.argmax(np.diff(a[0]) >= 2))
    bad[a[0][r] + 1] = bad[a[0][r]]
    with pytest.raises(FxiiError):
        DeviceCSR.from_host(a[0], bad, a[2], A.shape).validate()
    bad = a[1].copy()
    bad[0] = 40
    with pytest.raises(FxiiError):
        DeviceCSR.from_host(a[0], bad, a[2], A.shape).validate()
    # from_scipy sums duplicates instead of passing them on
    D = sp.csr_matrix((np.array([1.0, 2.0, 3.0]), np.array([1, 1, 0]), np.array([0, 3])), shape=(1, 3))
    dd = DeviceCSR.from_scipy(D).validate().to_host()
    np.testing.assert_array_equal(dd[1], [0, 1])
    np.testing.assert_array_equal(dd[2], [3.0, 3.0])


def test_sharded_products_single_rank(cuda):
    """ShardedProducts on one rank: C = Πᵀ(M_ΓΠ) with the mass diagonal fused and M Π — against the oracle / scipy."""
    import torch

    from fenicsx_ii_b200.distributed import ShardedProducts, p1_quadrature_coupling

    mesh = sy.kuhn_box(8, 8, 8)
    V = functionspace(mesh, ("Lagrange", 1))
    gamma = sy.polyline_network(150, 1 / 8, n_lines=4, lo=0.2, hi=0.8)
    W = functionspace(gamma, ("Quadrature", 10))
    Q = functionspace(gamma, ("Lagrange", 1))
    P, _, _ = create_interpolation_matrix(V, W, Circle(gamma, 0.06, degree=9))
    p = (P.indptr, P.indices, P.data)
    diag = quadrature_mass_diagonal(gamma, W)
    M = p1_quadrature_coupling(gamma, Q, W)
    np.testing.assert_allclose(np.asarray(M.sum(axis=0)).reshape(-1), diag, rtol=1e-14)  # P1 is a partition of unity
    prod = ShardedProducts(V.num_dofs, torch.from_numpy(diag).to("cuda"), DeviceCSR.from_scipy(M))
    c, mp, ip, nnz_a, ip_mp = prod(P.device)
    pt = orc.csr_transpose(*p, V.num_dofs)
    mpi = (p[0], p[1], p[2] * np.repeat(diag, np.diff(p[0])))
    ref = orc.spgemm(*pt, *mpi, V.num_dofs)
    assert ip == ref[3] and nnz_a == P.device.nnz
    check(c, ref[:3])
    refm = orc.spgemm(*as_host(M), *p, V.num_dofs)
    assert ip_mp == refm[3]
    check(mp, refm[:3])


@pytest.mark.parametrize("n_sub", [1, 3, 4])
def test_sharded_products_streamed_to_host(cuda, n_sub):
    """ShardedProducts.to_host — MΠ first, C in `n_sub` row pieces whose downloads overlap the next piece's product —
    delivers bit for bit the matrices of the one-shot products into the caller's pinned host arrays."""
    import torch

    from fenicsx_ii_b200.distributed import ShardedProducts, p1_quadrature_coupling

    mesh = sy.kuhn_box(8, 8, 8)
    V = functionspace(mesh, ("Lagrange", 1))
    gamma = sy.polyline_network(150, 1 / 8, n_lines=4, lo=0.2, hi=0.8)
    W = functionspace(gamma, ("Quadrature", 10))
    Q = functionspace(gamma, ("Lagrange", 1))
    P, _, _ = create_interpolation_matrix(V, W, Circle(gamma, 0.06, degree=9))
    diag = quadrature_mass_diagonal(gamma, W)
    M = p1_quadrature_coupling(gamma, Q, W)
    prod = ShardedProducts(V.num_dofs, torch.from_numpy(diag).to("cuda"), DeviceCSR.from_scipy(M))
    c, mp, *_ = prod(P.device)
    ref_c, ref_mp = c.to_host(), mp.to_host()

    def pinned(n, dtype):
        return torch.zeros(n, dtype=dtype).pin_memory().numpy()

    host_c = (pinned(V.num_dofs + 1, torch.int64), pinned(c.nnz + 5, torch.int32), pinned(c.nnz + 5, torch.float64))
    host_mp = (pinned(mp.shape[0] + 1, torch.int64), pinned(mp.nnz, torch.int32), pinned(mp.nnz, torch.float64))
    for _ in range(2):  # the second call reuses the cached piece boundaries
        for a in host_c + host_mp:
            a[:] = 0
        r = prod.to_host(P.device, host_c, host_mp, n_sub=n_sub)
        r["copy_stream"].synchronize()
        assert r["nnz_c"] == c.nnz and r["nnz_mp"] == mp.nnz and r["rows_c"] == V.num_dofs
        np.testing.assert_array_equal(host_c[0], ref_c[0])
        np.testing.assert_array_equal(host_c[1][: c.nnz], ref_c[1])
        np.testing.assert_array_equal(host_c[2][: c.nnz], ref_c[2])
        for got, ref in zip(host_mp, ref_mp):
            np.testing.assert_array_equal(got, ref)
    with pytest.raises(ValueError, match="too small"):
        prod.to_host(P.device, (host_c[0], host_c[1][:10], host_c[2][:10]), host_mp, n_sub=n_sub)
    torch.cuda.synchronize()


def test_spgemm_keeps_cancellation_and_explicit_zeros(cuda):
    A = sp.csr_matrix(np.array([[1.0, -1.0, 0.0], [0.0, 0.0, 2.0]]))
    B = sp.csr_matrix((np.array([3.0, 3.0, 0.0]), np.array([0, 0, 1]), np.array([0, 1, 2, 3])), shape=(3, 2))
    C = DeviceCSR.from_host(*as_host(A), A.shape).matmult(DeviceCSR.from_host(B.indptr, B.indices, B.data, B.shape))
    ip, ij, iv = C.to_host()
    np.testing.assert_array_equal(ip, [0, 1, 2])  # (0,0) = 3-3 = 0 kept; (1,1) = 2*0 = 0 kept
    np.testing.assert_array_equal(ij, [0, 1])
    np.testing.assert_array_equal(iv, [0.0, 0.0])


@pytest.mark.parametrize("m,n,density", [(200, 300, 0.05), (1000, 7, 0.3), (5, 2000, 0.01)])
def test_transpose_random(cuda, m, n, density):
    A = random_csr(m, n, density, 5)
    a = as_host(A)
    T = DeviceCSR.from_host(*a, A.shape).transpose()
    check(T, orc.csr_transpose(*a, n), rtol=0)
    assert T.shape == (n, m)


def test_transpose_long_columns_and_windows(cuda):
    """Very long columns (thousands of entries), short columns and empty column windows: the transpose and its
    windows give MatTranspose's order bit for bit."""
    rng = np.random.default_rng(9)
    A = sp.random(6000, 40, density=0.85, random_state=4, format="csr")  # columns of ~5100 entries
    B = sp.random(3000, 500, density=0.01, random_state=5, format="csr")  # short columns
    for M in (A, B):
        M.sort_indices()
        a = as_host(M)
        d = DeviceCSR.from_host(*a, M.shape)
        ref = orc.csr_transpose(*a, M.shape[1])
        check(d.transpose(), ref, rtol=0)
        for c0, c1 in ((0, M.shape[1] // 3), (M.shape[1] // 3, M.shape[1]), (5, 5)):
            Tw = d.transpose(c0, c1).to_host()
            np.testing.assert_array_equal(Tw[0], ref[0][c0:c1 + 1] - ref[0][c0])
            np.testing.assert_array_equal(Tw[1], ref[1][ref[0][c0]:ref[0][c1]])
            assert Tw[2].tobytes() == ref[2][ref[0][c0]:ref[0][c1]].tobytes()
    assert rng is not None


def test_block_products_of_the_demo(cuda):
    """demos/coupled_poisson.py:107-121,164-177 on the scaled config-1 geometry:
    C = A10 T, Pt = P^T, Z = Pt (M_Γ T), D = Pt A01 — all against the oracle."""
    mesh = sy.kuhn_box(8, 8, 16)
    V = functionspace(mesh, ("Lagrange", 1))
    gamma = sy.straight_line(17)
    W = functionspace(gamma, ("Quadrature", 10))
    Q = functionspace(gamma, ("Lagrange", 1))
    T, _, _ = create_interpolation_matrix(V, W, Circle(gamma, 0.2, degree=5))
    P, _, _ = create_interpolation_matrix(V, W, PointwiseTrace(gamma))
    t = (T.indptr, T.indices, T.data)
    p = (P.indptr, P.indices, P.data)
    nV, nW, nQ = V.num_dofs, W.num_dofs, Q.num_dofs
    # A10 (Q x W): P1 test functions against quadrature-element trial functions: phi_q(x_j) w_j |seg|
    diag = quadrature_mass_diagonal(gamma, W)
    xref = W.element.interpolation_points[:, 0]
    rows, cols, vals = [], [], []
    for c in range(gamma.geometry.dofmap.shape[0]):
        for j, X in enumerate(xref):
            wdof = W.dofmap.list[c, j]
            for q, phi in zip(Q.dofmap.list[c], (1 - X, X)):
                rows.append(q), cols.append(wdof), vals.append(-10.0 * phi * diag[wdof])
    A10 = sp.csr_matrix((vals, (rows, cols)), shape=(nQ, nW))
    A01 = A10.T.tocsr()
    a10, a01 = as_host(A10), as_host(A01)
    # C = A10 T   (case [1])
    Cd = restrict_columns(A10, None, None, None, pi=T.device)
    check(Cd, orc.spgemm(*a10, *t, nV)[:3])
    # D = Pt A01  (case [0])
    Dd = restrict_rows(A01, None, None, None, pi=P.device)
    pt = orc.csr_transpose(*p, nV)
    check(Dd, orc.spgemm(*pt, *a01, nQ)[:3])
    # Z = Pt M T  (case [0,1]) in both associations
    Mg = sp.diags(diag).tocsr()
    mg = as_host(Mg)
    z_left = orc.spgemm(*orc.spgemm(*pt, *mg, nW)[:3], *t, nV)
    Zd = restrict_rows_and_columns(Mg, None, None, None, None, None, None, pi_test=P.device, pi_trial=T.device)
    check(Zd, z_left[:3])
    z_right = orc.spgemm(*pt, *orc.spgemm(*mg, *t, nV)[:3], nV)
    Zd2 = restrict_rows_and_columns(Mg, None, None, None, None, None, None, pi_test=P.device, pi_trial=T.device,
                                    association="north_star")
    check(Zd2, z_right[:3])
    np.testing.assert_array_equal(z_left[1], z_right[1])  # same symbolic pattern either way
    # diagonal mass as a row scaling == SpGEMM with the diagonal matrix
    MT = T.device.copy().scale_rows(diag)
    check(MT, orc.spgemm(*mg, *t, nV)[:3])


def test_spmv_and_transpose_spmv(cuda):
    """assembly.py:67 (K.mult) and vector_assembler.py:177 (Π^T b)."""
    mesh = sy.kuhn_box(6, 6, 6)
    V = functionspace(mesh, ("Lagrange", 1))
    gamma = sy.polyline_network(80, 1 / 6, n_lines=3, lo=0.3, hi=0.7)
    K = functionspace(gamma, ("DG", 1))
    A, _, _ = create_interpolation_matrix(V, K, Circle(gamma, 0.04, degree=7))
    rng = np.random.default_rng(4)
    u = rng.normal(size=V.num_dofs)
    bvec = rng.normal(size=K.num_dofs)
    S = A.to_scipy()
    np.testing.assert_allclose(A.mult(u), S @ u, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(A.device.mult_transpose(bvec), S.T @ bvec, rtol=1e-12, atol=1e-13)
    # average_coefficient(u) of a linear function along the circle axis direction... row sums are one
    np.testing.assert_allclose(A.mult(np.ones(V.num_dofs)), 1.0, atol=1e-13)


def test_transpose_window_and_sharded_product(cuda):
    """The multi-GPU product on one device: every 'rank' transposes only its window of 3D dofs and
    multiplies; the row-concatenation equals the unsharded Πᵀ(MΠ) bit for bit."""
    from fenicsx_ii_b200.distributed import shard_range

    mesh = sy.kuhn_box(8, 8, 8)
    V = functionspace(mesh, ("Lagrange", 1))
    gamma = sy.polyline_network(150, 1 / 8, n_lines=4, lo=0.2, hi=0.8)
    W = functionspace(gamma, ("Quadrature", 10))
    P, _, _ = create_interpolation_matrix(V, W, Circle(gamma, 0.06, degree=9))
    diag = quadrature_mass_diagonal(gamma, W)
    MP = P.device.copy().scale_rows(diag)
    full_t = P.device.transpose()
    ft = full_t.to_host()
    C_full = full_t.matmult(MP).to_host()
    parts = []
    for world in (1, 3):
        ind, idx, val = [np.zeros(1, dtype=np.int64)], [], []
        for rank in range(world):
            r0, r1 = shard_range(V.num_dofs, world, rank)
            Tw = P.device.transpose(r0, r1)
            assert Tw.shape == (r1 - r0, W.num_dofs)
            ti, tj, tv = Tw.to_host()
            np.testing.assert_array_equal(ti, ft[0][r0 : r1 + 1] - ft[0][r0])
            np.testing.assert_array_equal(tj, ft[1][ft[0][r0] : ft[0][r1]])
            assert tv.tobytes() == ft[2][ft[0][r0] : ft[0][r1]].tobytes()
            ci, cj, cv = Tw.matmult(MP).to_host()
            ind.append(ci[1:] + ind[-1][-1])
            idx.append(cj)
            val.append(cv)
        np.testing.assert_array_equal(np.concatenate(ind), C_full[0])
        np.testing.assert_array_equal(np.concatenate(idx), C_full[1])
        assert np.concatenate(val).tobytes() == C_full[2].tobytes()


def test_scale_rows_with_device_diagonal(cuda):
    """fxii_csr_scale_rows takes the diagonal from host memory or, stream-ordered and without a copy, from device memory."""
    import scipy.sparse as sp
    import torch

    from fenicsx_ii_b200.sparse import DeviceCSR

    rng = np.random.default_rng(3)
    M = sp.random(500, 300, density=0.02, random_state=7, format="csr")
    M.sort_indices()
    d = rng.uniform(-2, 2, size=500)
    a = DeviceCSR.from_scipy(M).scale_rows(d).to_host()
    b = DeviceCSR.from_scipy(M).scale_rows(torch.from_numpy(d).to("cuda")).to_host()
    ref = M.data * np.repeat(d, np.diff(M.indptr))  # row-wise product in M's (sorted) entry order
    np.testing.assert_array_equal(a[1], b[1])
    assert a[2].tobytes() == b[2].tobytes()
    np.testing.assert_array_equal(a[2], ref)


@pytest.mark.parametrize("pi_pieces,c_pieces,subset", [(1, 1, False), (2, 4, False), (3, 2, True), (1, 2, True)])
def test_assemble_blocks_to_host(cuda, pi_pieces, c_pieces, subset):
    """assemble_blocks_to_host — Π in blocks of Γ cells, each downloaded while the next is built, concatenated on the
    device for the streamed products — fills the host arrays with exactly the matrices of the one-shot calls."""
    import torch

    from fenicsx_ii_b200.distributed import ShardedProducts, assemble_blocks_to_host, p1_quadrature_coupling

    mesh = sy.kuhn_box(8, 8, 8)
    V = functionspace(mesh, ("Lagrange", 1))
    gamma = sy.polyline_network(150, 1 / 8, n_lines=4, lo=0.2, hi=0.8)
    W = functionspace(gamma, ("Quadrature", 10))
    Q = functionspace(gamma, ("Lagrange", 1))
    nip = W.element.interpolation_points.shape[0]
    seg_r = np.linspace(0.03, 0.07, 150)
    op = Circle(gamma, seg_r, degree=9)
    cells = np.arange(20, 131, dtype=np.int32) if subset else None
    P, _, _ = create_interpolation_matrix(V, W, op, cells_K=cells)
    pd = P.device if cells is None else P.device.row_slice(20 * nip, 131 * nip)
    ref_p = pd.to_host()
    diag = quadrature_mass_diagonal(gamma, W)
    M = p1_quadrature_coupling(gamma, Q, W)
    if subset:  # the products of a sub-block: restrict M_Γ and M to its rows / columns
        diag = diag[20 * nip:131 * nip]
        M = M[:, 20 * nip:131 * nip].tocsr()
    prod = ShardedProducts(V.num_dofs, torch.from_numpy(np.ascontiguousarray(diag)).to("cuda"), DeviceCSR.from_scipy(M))
    c, mp, *_ = prod(pd)
    ref_c, ref_mp = c.to_host(), mp.to_host()

    def pinned(n, dtype):
        return torch.zeros(n, dtype=dtype).pin_memory().numpy()

    def triple(rows, nnz):
        return pinned(rows + 1, torch.int64), pinned(nnz, torch.int32), pinned(nnz, torch.float64)

    host_pi, host_c, host_mp = triple(pd.shape[0], pd.nnz), triple(V.num_dofs, c.nnz), triple(mp.shape[0], mp.nnz)
    r = assemble_blocks_to_host(V, W, op, prod, host_pi, host_c, host_mp, cells_K=cells, pi_pieces=pi_pieces, c_pieces=c_pieces)
    r["copy_stream"].synchronize()
    assert (r["nnz_pi"], r["nnz_c"], r["nnz_mp"], r["rows_pi"]) == (pd.nnz, c.nnz, mp.nnz, pd.shape[0])
    for got, ref in ((host_pi, ref_p), (host_c, ref_c), (host_mp, ref_mp)):
        for g, e in zip(got, ref):
            np.testing.assert_array_equal(g, e)
