"""GPU path against the golden fixtures produced by the reference's own interpolation.py
(tests/golden/pi_*.npz): bit-exact pattern and ownership, values to 1e-12 of the row scale.

What the fixtures pin: everything the reference itself implements (point ordering, operator quadrature, pattern,
weights/scales arithmetic, first-visit-wins, accumulation order).  What they cannot pin: point location and basis
tabulation — the dolfinx shim the reference ran on takes both from oracle/pi_oracle.py (its `determine_point_ownership`
IS the oracle's GridLocator), so the `cell` arrays compared here are GPU-vs-oracle, not GPU-vs-dolfinx."""

import numpy as np
import pytest

from fenicsx_ii_b200 import create_interpolation_matrix
from tests.golden_cases import CASES, load_case
from tests.helpers import assert_csr_parity, make_operator

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("materialize", [False, True])
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_matches_reference_fixture(cuda, name, materialize):
    mesh, V, gamma, K, op_name, radius, degree, cells_K, fx = load_case(name)
    op = make_operator(op_name, gamma, radius, degree)
    A, _, _ = create_interpolation_matrix(V, K, op, cells_K=cells_K, materialize_coo=materialize)
    cell, ncoll = A.rows.diagnostics()
    np.testing.assert_array_equal(cell, fx["cell"])
    np.testing.assert_array_equal(ncoll.astype(np.int64), np.minimum(fx["n_colliding"], 255))
    assert_csr_parity((A.indptr, A.indices, A.data), dict(indptr=fx["indptr"], indices=fx["indices"], data=fx["data"]))
