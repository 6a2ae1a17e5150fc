"""CPU: meshfem_b200/csrc/elem_math.cuh -- the element math the CUDA kernels inline (embedding, the factorised W (x) S
form of Ke, the rotated row slices of the block-owner assembly, the integrated shape-function gradients) -- compiled for
the HOST with g++ (tests/host_harness/elem_math_host.cc) and compared with the oracle's restatement of
EmbeddedElement.hh:170-231 and Element::perElementStiffness (LinearElasticity.hh:165-232) on random, badly shaped
simplices with isotropic, orthotropic and fully anisotropic tensors."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import meshfem_oracle as orc
from util import ORTHO, ROOT


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("harness") / "libelem_math_host.so")
    src = os.path.join(ROOT, "tests", "host_harness", "elem_math_host.cc")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", src, "-o", out])
    lib = ctypes.CDLL(out)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.harness_ke.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, dp]
    lib.harness_ke_rot.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, ctypes.c_int]
    lib.harness_int_grads.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp]
    lib.harness_elem_apply.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, dp]
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _materials(N, rng):
    F = N * (N + 1) // 2
    A = rng.standard_normal((F, F))
    mats = [orc.isotropic_D(N, 200.0, 0.35), A @ A.T + F * np.eye(F)]
    if N == 3:
        mats.append(orc.material_from_json(3, ORTHO))
    return mats


def _simplex(N, rng):
    while True:
        P = rng.standard_normal((N + 1, N)) * rng.uniform(0.2, 3.0, size=N)
        vol, _ = orc.embed_simplices(P[None])
        if vol[0] > 1e-3:
            return P


@pytest.mark.parametrize("N", [2, 3])
@pytest.mark.parametrize("deg", [1, 2])
def test_host_compiled_element_math_matches_oracle(harness, N, deg):
    rng = np.random.default_rng(100 * N + deg)
    nn = orc.num_nodes(N, deg)
    n = N * nn
    for trial in range(6):
        P = _simplex(N, rng)
        vol, G = orc.embed_simplices(P[None])
        for D in _materials(N, rng):
            D = np.ascontiguousarray(D)
            Ke = np.zeros((n, n)); geom = np.zeros(1 + N * (N + 1))
            assert harness.harness_ke(N, deg, _ptr(np.ascontiguousarray(P)), _ptr(D), _ptr(Ke), _ptr(geom)) == 0
            assert abs(geom[0] - vol[0]) <= 1e-14 * abs(vol[0])
            assert np.abs(geom[1:].reshape(N, N + 1) - G[0]).max() <= 1e-13 * np.abs(G[0]).max()
            ref = orc.per_element_stiffness(N, deg, vol, G, D)[0]
            assert np.abs(Ke - ref).max() <= 1e-13 * np.abs(ref).max()
            assert np.abs(Ke - Ke.T).max() <= 1e-13 * np.abs(ref).max()
            # the literal loop nest of the reference writes the upper triangle only
            loops = orc.per_element_stiffness_reference_loops(N, deg, vol[0], G[0], D)
            iu = np.triu_indices(n)
            assert np.abs(Ke[iu] - loops[iu]).max() <= 1e-13 * np.abs(ref).max()
            # rotated row slices (what the block-owner assembly evaluates): same matrix for every rotation
            for rot in range(N + 1):
                Kr = np.zeros((n, n))
                assert harness.harness_ke_rot(N, deg, _ptr(np.ascontiguousarray(P)), _ptr(D), _ptr(Kr), rot) == 0
                assert np.abs(Kr - ref).max() <= 1e-13 * np.abs(ref).max()


@pytest.mark.parametrize("N", [2, 3])
@pytest.mark.parametrize("deg", [1, 2])
def test_integrated_shape_function_gradients(harness, N, deg):
    """int grad phi_i over the element (what constantStrainLoad needs, LinearElasticity.hh:135-162): for degree 2 the
    integral of the linear interpolant = vol / (K+1) * sum over vertices (Functions.hh:247-253)."""
    rng = np.random.default_rng(7 + N + deg)
    nn = orc.num_nodes(N, deg)
    P = _simplex(N, rng)
    vol, G = orc.embed_simplices(P[None])
    out = np.zeros((nn, N))
    assert harness.harness_int_grads(N, deg, _ptr(np.ascontiguousarray(P)), _ptr(out)) == 0
    T = orc.grad_phi_interpolant(N, deg)                      # (node, interp node, barycentric index)
    ref = vol[0] * np.einsum("iva,ra->ir", T, G[0]) / T.shape[1]
    assert np.abs(out - ref).max() <= 1e-13 * np.abs(ref).max()
    assert np.abs(out.sum(axis=0)).max() <= 1e-12 * np.abs(ref).max()        # partition of unity: gradients sum to zero


@pytest.mark.parametrize("N", [2, 3])
@pytest.mark.parametrize("deg", [1, 2])
def test_matrix_free_element_operator_equals_Ke_times_x(harness, N, deg):
    """elem_apply (the PCG's matrix-free operator, csrc/matfree.inl) == perElementStiffness (LinearElasticity.hh:165-232)
    applied to a vector, on badly shaped simplices and isotropic / orthotropic / fully anisotropic tensors."""
    rng = np.random.default_rng(900 + 10 * N + deg)
    nn = orc.num_nodes(N, deg)
    for trial in range(6):
        P = _simplex(N, rng)
        vol, G = orc.embed_simplices(P[None])
        for D in _materials(N, rng):
            D = np.ascontiguousarray(D)
            ref = orc.per_element_stiffness(N, deg, vol, G, D)[0]
            for _ in range(3):
                xe = np.ascontiguousarray(rng.standard_normal(N * nn))
                ye = np.zeros(N * nn)
                assert harness.harness_elem_apply(N, deg, _ptr(np.ascontiguousarray(P)), _ptr(D), _ptr(xe), _ptr(ye)) == 0
                want = ref @ xe
                assert np.abs(ye - want).max() <= 1e-13 * np.abs(ref).max() * np.abs(xe).max() * N * nn
            # rigid translations are in the kernel of Ke
            ye = np.zeros(N * nn)
            xe = np.ascontiguousarray(np.tile(rng.standard_normal(N), nn))
            harness.harness_elem_apply(N, deg, _ptr(np.ascontiguousarray(P)), _ptr(D), _ptr(xe), _ptr(ye))
            assert np.abs(ye).max() <= 1e-12 * np.abs(ref).max()
