"""GPU: the reference's entry points for the path — ``average_coefficients(form)`` (assembly.py:47-68),
``apply_matrix_replacer`` (matrix_assembler.py:130-264, cases [], [0], [1], [0,1] and the axpy sum of the parts),
``apply_vector_replacer`` (vector_assembler.py:120-185) — on duck-typed forms (tests/boundary_stubs.py), with scipy
matrices (MatrixCSR out) and with a recording stand-in for petsc4py (PETSc AIJ + LG maps out)."""

import numpy as np
import pytest
import scipy.sparse as sp

from fenicsx_ii_b200 import (Circle, MatrixCSR, PointwiseTrace, apply_matrix_replacer, apply_vector_replacer,
                             average_coefficients, create_interpolation_matrix)
from fenicsx_ii_b200 import synthetic as sy
from fenicsx_ii_b200.distributed import p1_quadrature_coupling, quadrature_mass_diagonal
from fenicsx_ii_b200.mesh import functionspace
from oracle import pi_oracle as orc
from tests.boundary_stubs import Argument, AveragedArgument, AveragedCoefficient, Form, Function, SplitForm
from tests.test_petsc_handoff import fake_petsc  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def problem():
    mesh = sy.kuhn_box(8, 8, 16)
    V = functionspace(mesh, ("Lagrange", 1))
    gamma = sy.straight_line(17)
    W = functionspace(gamma, ("Quadrature", 10))
    Q = functionspace(gamma, ("Lagrange", 1))
    circle, trace = Circle(gamma, 0.2, degree=5), PointwiseTrace(gamma)
    T = create_interpolation_matrix(V, W, circle)[0].to_scipy()
    P = create_interpolation_matrix(V, W, trace)[0].to_scipy()
    Mg = sp.diags(quadrature_mass_diagonal(gamma, W)).tocsr()
    A10 = (-10.0 * p1_quadrature_coupling(gamma, Q, W)).tocsr()
    return dict(mesh=mesh, V=V, gamma=gamma, W=W, Q=Q, circle=circle, trace=trace, T=T, P=P, Mg=Mg, A10=A10)


def host(M):
    M = M.tocsr()
    M.sort_indices()
    return M.indptr.astype(np.int64), M.indices.astype(np.int32), M.data.astype(np.float64)


def assert_same(got, ref, rtol=1e-12):
    """got: MatrixCSR or a fake PETSc Mat; ref: (indptr, indices, data) with the symbolic pattern."""
    ip, ij, iv = (got.indptr, got.indices, got.data) if isinstance(got, MatrixCSR) else got.getValuesCSR()
    np.testing.assert_array_equal(ip, ref[0])
    np.testing.assert_array_equal(ij, ref[1])
    np.testing.assert_allclose(iv, ref[2], rtol=rtol, atol=rtol * max(np.abs(ref[2]).max(), 1e-300))


def test_average_coefficients_form(cuda, problem):
    V, W, gamma = problem["V"], problem["W"], problem["gamma"]
    x = problem["mesh"].geometry.x
    u = Function(V, 0.3 * x[:, 0] - 1.7 * x[:, 1] + 0.5 * x[:, 2] + 2.0)
    c_avg = AveragedCoefficient(W, problem["circle"], u)
    c_tr = AveragedCoefficient(W, problem["trace"], u)
    plain = Function(W, 1.0)
    average_coefficients(Form(coefficients=[c_avg, plain, c_tr]))
    np.testing.assert_allclose(c_avg.x.array, problem["T"] @ u.x.array, rtol=1e-13, atol=1e-14)
    np.testing.assert_allclose(c_tr.x.array, problem["P"] @ u.x.array, rtol=1e-13, atol=1e-14)
    x0 = orc.physical_points(gamma.geometry.x, gamma.geometry.dofmap, np.arange(16), W.element.interpolation_points)
    np.testing.assert_allclose(c_tr.x.array, 0.3 * x0[:, 0] - 1.7 * x0[:, 1] + 0.5 * x0[:, 2] + 2.0, atol=1e-13)  # trace of a linear function
    np.testing.assert_array_equal(plain.x.array, 1.0)  # untouched
    assert c_avg.x.scatters == 1 and plain.x.scatters == 1 and c_tr.x.scatters == 1  # scatter_forward on every coefficient


def forms_of(problem):
    V, W, Q = problem["V"], problem["W"], problem["Q"]
    f_plain = Form([Argument(Q, 0), Argument(W, 1)], tag="A10")                                      # []
    f_rows = Form([AveragedArgument(W, 0, V, problem["trace"]), Argument(Q, 1)], tag="A01")          # [0]
    f_cols = Form([Argument(Q, 0), AveragedArgument(W, 1, V, problem["circle"])], tag="A10")         # [1]
    f_both = Form([AveragedArgument(W, 0, V, problem["trace"]), AveragedArgument(W, 1, V, problem["circle"])], tag="Mg")  # [0,1]
    return f_plain, f_rows, f_cols, f_both


def references(problem):
    t, p = host(problem["T"]), host(problem["P"])
    nV, nW, nQ = problem["V"].num_dofs, problem["W"].num_dofs, problem["Q"].num_dofs
    a10, a01, mg = host(problem["A10"]), host(problem["A10"].T), host(problem["Mg"])
    pt = orc.csr_transpose(*p, nV)
    return dict(rows=orc.spgemm(*pt, *a01, nQ)[:3], cols=orc.spgemm(*a10, *t, nV)[:3],
                both=orc.spgemm(*orc.spgemm(*pt, *mg, nW)[:3], *t, nV)[:3])


def test_apply_matrix_replacer_cases_scipy(cuda, problem):
    f_plain, f_rows, f_cols, f_both = forms_of(problem)
    mats = {"A10": problem["A10"], "A01": problem["A10"].T.tocsr(), "Mg": problem["Mg"]}
    posted = []
    get = lambda f: mats[f.tag]  # noqa: E731
    ref = references(problem)
    D = apply_matrix_replacer(f_rows, get, posted.append)
    assert isinstance(D, MatrixCSR) and D.shape == (problem["V"].num_dofs, problem["Q"].num_dofs) and len(posted) == 1
    assert_same(D, ref["rows"])
    assert_same(apply_matrix_replacer(f_cols, get), ref["cols"])
    assert_same(apply_matrix_replacer(f_both, get), ref["both"])
    assert_same(apply_matrix_replacer(f_plain, get), host(problem["A10"]), rtol=0)
    with pytest.raises(ValueError):
        apply_matrix_replacer(Form([Argument(problem["Q"], 0)]), get)
    # a form that splits into two parts: the parts are summed with axpy into the union pattern (explicit zeros stay)
    nV = problem["V"].num_dofs
    extra = sp.random(nV, nV, density=0.002, random_state=3, format="csr") + sp.eye(nV, format="csr")
    mats["extra"] = extra
    parts = SplitForm([f_both, Form([Argument(problem["V"], 0), Argument(problem["V"], 1)], tag="extra")])
    S = apply_matrix_replacer(parts, get, apply_replacer=lambda a: a.parts)
    both = sp.csr_matrix((ref["both"][2], ref["both"][1], ref["both"][0]), shape=(nV, nV))
    ones = lambda M: sp.csr_matrix((np.ones(M.nnz), M.indices, M.indptr), shape=M.shape)  # noqa: E731
    pattern = (ones(both) + ones(extra.tocsr())).tocsr()
    pattern.sort_indices()
    np.testing.assert_array_equal(S.indptr, pattern.indptr)
    np.testing.assert_array_equal(S.indices, pattern.indices)
    np.testing.assert_allclose(S.to_scipy().toarray(), (both + extra).toarray(), rtol=1e-12, atol=1e-14)
    # into an existing matrix
    S2 = apply_matrix_replacer(f_both, get, matrix=MatrixCSR(__import__("fenicsx_ii_b200").DeviceCSR.from_scipy(extra)))
    np.testing.assert_allclose(S2.to_scipy().toarray(), (both + extra).toarray(), rtol=1e-12, atol=1e-14)


def test_apply_matrix_replacer_petsc_in_petsc_out(cuda, problem, fake_petsc):  # noqa: F811
    f_plain, f_rows, f_cols, f_both = forms_of(problem)

    def petsc(M):
        ip, ij, iv = host(M)
        return fake_petsc.Mat().createAIJ(size=M.shape, csr=(ip.astype(fake_petsc.IntType), ij, iv))

    mats = {"A10": petsc(problem["A10"]), "A01": petsc(problem["A10"].T.tocsr()), "Mg": petsc(problem["Mg"])}
    ref = references(problem)
    nV, nW, nQ = problem["V"].num_dofs, problem["W"].num_dofs, problem["Q"].num_dofs
    for form, key, size in ((f_rows, "rows", (nV, nQ)), (f_cols, "cols", (nQ, nV)), (f_both, "both", (nV, nV))):
        fake_petsc.CALLS.clear()
        D = apply_matrix_replacer(form, lambda f: mats[f.tag], lambda A: A.assemble())
        assert isinstance(D, fake_petsc.Mat) and D.assembled and D.getSize() == size
        assert_same(D, ref[key])
        names = [c[0] for c in fake_petsc.CALLS]
        assert names == ["Mat.createAIJ", "LGMap.create", "LGMap.create", "Mat.setLGMap"], names
        assert fake_petsc.CALLS[0][1]["size"] == size
        assert fake_petsc.CALLS[3][1] == dict(rows=size[0], cols=size[1], rbs=1, cbs=1)
    # [] passes the assembled matrix through; a second part is added with PETSc's own axpy
    assert apply_matrix_replacer(f_plain, lambda f: mats[f.tag]) is mats["A10"]
    fake_petsc.CALLS.clear()
    tgt = petsc(sp.eye(nV, format="csr"))
    out = apply_matrix_replacer(f_both, lambda f: mats[f.tag], matrix=tgt)
    assert out is tgt and ("Mat.axpy", dict(alpha=1.0)) in fake_petsc.CALLS


def test_apply_vector_replacer(cuda, problem, fake_petsc):  # noqa: F811
    V, W = problem["V"], problem["W"]
    rng = np.random.default_rng(2)
    b = rng.normal(size=W.num_dofs)
    g = rng.normal(size=V.num_dofs)
    f_avg = Form([AveragedArgument(W, 0, V, problem["circle"])], tag="b")
    f_plain = Form([Argument(V, 0)], tag="g")
    vecs = {"b": b, "g": g}
    z = apply_vector_replacer(f_avg, lambda f: vecs[f.tag])
    np.testing.assert_allclose(z, problem["T"].T @ b, rtol=1e-12, atol=1e-13)
    parts = SplitForm([f_avg, f_plain])
    z2 = apply_vector_replacer(parts, lambda f: vecs[f.tag], apply_replacer=lambda a: a.parts)
    np.testing.assert_allclose(z2, problem["T"].T @ b + g, rtol=1e-12, atol=1e-13)
    # PETSc Vec in, PETSc Vec out
    pv = {"b": fake_petsc.Vec().createWithArray(b), "g": fake_petsc.Vec().createWithArray(g)}
    zv = apply_vector_replacer(parts, lambda f: pv[f.tag], lambda v: None, apply_replacer=lambda a: a.parts)
    assert isinstance(zv, fake_petsc.Vec)
    np.testing.assert_allclose(zv.getArray(), problem["T"].T @ b + g, rtol=1e-12, atol=1e-13)
    with pytest.raises(ValueError):
        apply_vector_replacer(Form([Argument(V, 0), Argument(V, 1)]), lambda f: g)


def test_overridden_plugin_contract_is_honoured(cuda, problem):
    """A subclass of a stock operator that overrides compute_quadrature is a generic operator: its own points are used
    (uploaded), not the device-generated stock circle."""
    V, W, gamma = problem["V"], problem["W"], problem["gamma"]

    class HalfRadius(Circle):
        def compute_quadrature(self, cells, reference_points):
            q = super().compute_quadrature(cells, reference_points)
            centre = q.points.mean(axis=1, keepdims=True)
            q.points = centre + 0.5 * (q.points - centre)
            return q

    A = create_interpolation_matrix(V, W, HalfRadius(gamma, 0.2, degree=5))[0]
    stock = create_interpolation_matrix(V, W, Circle(gamma, 0.2, degree=5))[0]
    assert A.stats["n_points"] == stock.stats["n_points"]
    x = problem["mesh"].geometry.x
    r2 = (x[:, 0] - 0.5) ** 2 + (x[:, 1] - 0.5) ** 2
    # the P1 interpolant of r^2 averaged over the smaller circle is clearly smaller than over the stock one
    assert np.all(A.mult(r2) < stock.mult(r2) - 1e-3)


def test_scalar_type_of_the_returned_matrix(cuda, problem):
    """interpolation.py:175-203: ``complex_dtype=True`` returns a complex matrix (Π is real: zero imaginary part), a
    single-precision mesh a float32 one; the device arithmetic is float64 in every case, the cast happens once on the host."""
    from fenicsx_ii_b200.mesh import Mesh

    V, W = problem["V"], problem["W"]
    A = create_interpolation_matrix(V, W, problem["circle"])[0]
    Ac = create_interpolation_matrix(V, W, problem["circle"], complex_dtype=True)[0]
    assert A.data.dtype == np.float64 and Ac.data.dtype == np.complex128
    np.testing.assert_array_equal(Ac.indices, A.indices)
    np.testing.assert_array_equal(Ac.data.real, A.data)
    assert not Ac.data.imag.any() and Ac.to_scipy().dtype == np.complex128
    u = np.random.default_rng(0).normal(size=V.num_dofs) + 1j * np.random.default_rng(1).normal(size=V.num_dofs)
    np.testing.assert_allclose(Ac.mult(u), problem["T"] @ u, rtol=1e-12, atol=1e-13)
    m32 = Mesh(problem["mesh"].geometry.x.astype(np.float32), problem["mesh"].geometry.dofmap, tdim=3)
    m32.geometry.x = m32.geometry.x.astype(np.float32)  # the stand-in stores float64: make the mesh single precision
    V32 = functionspace(m32, ("Lagrange", 1))
    A32 = create_interpolation_matrix(V32, W, problem["circle"])[0]
    assert A32.data.dtype == np.float32
    np.testing.assert_array_equal(A32.indices, A.indices)  # the box mesh is exactly representable in float32
    np.testing.assert_allclose(A32.data, A.data, rtol=0, atol=1e-6)
