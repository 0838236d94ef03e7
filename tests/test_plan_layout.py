"""Host logic of the persistent group engine (csrc/pdhg_group.cuh: contiguous row blocks, halo lists, sliced-ELL
with split long rows) checked on the CPU through `asm_plan_check`: every row of the pattern must be rebuilt exactly
from the layout, for ragged, empty, very long and tiny inputs."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from activesetmethods_b200 import capi
from activesetmethods_b200.examples import acopf


def _check(lib, K, G):
    K = sp.csr_matrix(K)
    K.sort_indices()
    rp = np.ascontiguousarray(K.indptr, dtype=np.int64)
    ci = np.ascontiguousarray(K.indices, dtype=np.int32)
    smem, mats, padded = C.c_int64(), C.c_int32(), C.c_int64()
    rc = lib.asm_plan_check(K.shape[1], K.shape[0], rp.ctypes.data_as(capi.c_int64_p),
                            ci.ctypes.data_as(capi.c_int32_p), G, C.byref(smem), C.byref(mats), C.byref(padded))
    assert rc == 0, lib.asm_last_error().decode()
    return smem.value, mats.value, padded.value


@pytest.mark.parametrize("G", [1, 2, 7, 16, 148])
def test_random_ragged_patterns(built_lib, G):
    rng = np.random.default_rng(G)
    K = sp.random(500, 300, density=0.02, random_state=G, format="lil")
    K[3, :] = 1.0                     # one full row: 300 entries -> split over 32 lanes
    K[10:20, :] = 0.0                 # empty rows
    K[:, 5] = 1.0                     # a full column
    K[40, :70] = rng.standard_normal(70)
    smem, mats, padded = _check(built_lib, K, G)
    assert padded >= sp.csr_matrix(K).nnz and smem > 0
    _check(built_lib, sp.csr_matrix(K).T, G)      # the column side is the same builder on the transpose


def test_degenerate_shapes(built_lib):
    _check(built_lib, sp.csr_matrix((0, 3)), 1)                   # m = 0: a pure box LP
    _check(built_lib, sp.csr_matrix(np.ones((1, 1))), 4)          # fewer rows than blocks
    _check(built_lib, sp.csr_matrix((40, 9)), 3)                  # all rows empty
    _check(built_lib, sp.csr_matrix(np.ones((2, 5000))), 2)       # rows far longer than 32 x 8 entries


def test_acopf_patterns_fit_the_machine(built_lib):
    """The Jacobian patterns of the benchmark cases: layout exact, and the planned group sizes fit 227 KB."""
    limit = 227 * 1024 - 4096
    for name, G in (("case118", 2), ("case1354pegase", 16), ("case2869pegase", 74)):
        mdl = acopf.AcopfModel(acopf.synthetic_network(*acopf.PEGASE_SHAPES[name]))
        j = np.asarray(mdl.j_str) - 1
        K = sp.csr_matrix((np.ones(len(j)), (j[:, 0], j[:, 1])), shape=(mdl.m, mdl.n))
        smem, mats, padded = _check(built_lib, K, G)
        assert 0 < smem <= limit, (name, G, smem)
        assert padded <= 1.35 * K.nnz + 32 * 8 * G, (name, padded, K.nnz)     # little sliced-ELL padding
