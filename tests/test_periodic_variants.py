"""PeriodicCondition variants (BoundaryConditions.hh:457-610, PeriodicBoundaryMatcher.hh:268-372 of the reference):
mismatch-permitting matching, non-periodic axes (ignoreDims), identified pairs from a file; and the
macroscopic-to-microscopic tensors of --m2mstress (PeriodicHomogenization.hh:188-210).  Host C++ against the oracle."""
import re

import numpy as np
import pytest

import meshfem_oracle as orc


@pytest.fixture(scope="module")
def hostlib(lib_built):
    from meshfem_b200 import hostlib as hl
    return hl


def _mismatched_cell(hostlib, N):
    """A voxel cell whose max-x face lost some nodes' partners: remove a corner voxel so that part of the x=max face
    has no geometry (its nodes are gone) while the x=min face keeps them."""
    raw = hostlib.perforated_cell(N, 4, 0)
    V, T = raw.arrays()
    cent = V[T].mean(axis=1)
    keep = ~((cent[:, 0] > 0.75) & (cent[:, 1] > 0.75) & ((cent[:, 2] > 0.75) if N == 3 else True))
    T = T[keep]
    used = np.unique(T)
    remap = -np.ones(V.shape[0], dtype=np.int64); remap[used] = np.arange(used.size)
    return hostlib.from_arrays(N, V[used], remap[T]), V[used], remap[T]


@pytest.mark.parametrize("N,deg", [(2, 1), (2, 2), (3, 1), (3, 2)])
def test_mismatch_permitting_matching(hostlib, N, deg):
    raw, V, T = _mismatched_cell(hostlib, N)
    m = orc.build_mesh(N, deg, V, T)
    with pytest.raises(RuntimeError):
        orc.periodic_condition(m)                        # strict matching fails on this cell ...
    with pytest.raises(RuntimeError):
        raw.periodic_condition(deg)
    dof, nd, pbe = orc.periodic_condition(m, ignore_mismatch=True)      # ... the permissive one identifies what it can
    dof2, nd2, pbe2 = raw.periodic_condition(deg, ignore_mismatch=True)
    assert nd == nd2 and np.array_equal(dof, dof2) and np.array_equal(pbe, pbe2)
    assert nd < m.num_nodes
    # on a cell whose faces DO match, both matchers give the same identification
    full = hostlib.perforated_cell(N, 4, 2)
    Vf, Tf = full.arrays()
    mf = orc.build_mesh(N, deg, Vf, Tf)
    a = orc.periodic_condition(mf)
    b = orc.periodic_condition(mf, ignore_mismatch=True)
    c = full.periodic_condition(deg, ignore_mismatch=True)
    assert a[1] == b[1] == c[1] and np.array_equal(a[0], b[0]) and np.array_equal(a[0], c[0])
    assert np.array_equal(a[2], c[2])


@pytest.mark.parametrize("N,deg,ignore", [(2, 2, (1,)), (3, 1, (2,)), (3, 2, (0, 1))])
def test_ignore_dims(hostlib, N, deg, ignore):
    raw = hostlib.perforated_cell(N, 4, 2)
    V, T = raw.arrays()
    m = orc.build_mesh(N, deg, V, T)
    dof, nd, pbe = orc.periodic_condition(m, ignore_dims=ignore)
    dof2, nd2, pbe2 = raw.periodic_condition(deg, ignore_dims=ignore)
    assert nd == nd2 and np.array_equal(dof, dof2) and np.array_equal(pbe, pbe2)
    full = orc.periodic_condition(m)
    assert nd > full[1]                                   # fewer identifications than the fully periodic cell
    # identified nodes differ only along the periodic axes
    for d in np.unique(dof):
        nodes = np.nonzero(dof == d)[0]
        if nodes.size > 1:
            spread = np.ptp(m.nodes[nodes], axis=0)
            assert all(spread[a] < 1e-12 for a in ignore)


def test_pairs_file(hostlib, tmp_path):
    raw = hostlib.perforated_cell(2, 4, 2)
    V, T = raw.arrays()
    m = orc.build_mesh(2, 2, V, T)
    ref_dof, ref_nd, _ = orc.periodic_condition(m)
    # write the geometric identification as explicit pairs (chains, not stars, to exercise the component search)
    pairs = []
    for d in np.unique(ref_dof):
        nodes = np.nonzero(ref_dof == d)[0]
        pairs += [(int(nodes[k + 1]), int(nodes[k])) for k in range(nodes.size - 1)]
    path = tmp_path / "pairs.txt"
    path.write_text("".join(f"{a} {b}\n" for a, b in pairs))
    dof, nd, pbe = raw.periodic_condition(2, pairs_file=str(path))
    odof, ond, opbe = orc.periodic_condition_from_pairs(m, pairs)
    assert nd == ond == ref_nd and np.array_equal(dof, odof) and np.array_equal(dof, ref_dof)
    assert not pbe.any() and not opbe.any()               # the file constructor marks no periodic boundary elements
    with pytest.raises(RuntimeError, match="Couldn't open"):
        raw.periodic_condition(2, pairs_file=str(tmp_path / "missing.txt"))


@pytest.mark.parametrize("N,deg", [(2, 2), (3, 1)])
def test_macro_to_micro_stress_tensors(hostlib, N, deg):
    raw = hostlib.perforated_cell(N, 4, 2)
    V, T = raw.arrays()
    sim = orc.Simulator(N, deg, V, T)
    D = orc.isotropic_D(N, 200.0, 0.35)
    sim.set_material(D)
    w = orc.solve_cell_problems(sim)
    Eh = orc.homogenized_tensor_displacement_form(sim, w)
    M = orc.macro_to_micro_stress_tensors(sim, w, Eh)
    F = orc.flat_len(N)
    # theory: the cell average of the micro stress under a unit macro stress is that macro stress,
    # sum_e vol_e M_e / |Y| = symmetric identity (flattened: 1 on normal, 1/2 on shear entries)
    avg = np.einsum("e,eij->ij", sim.mesh.vol, M) / np.prod(sim.mesh.bbox_max - sim.mesh.bbox_min)
    ident = np.diag([1.0] * N + [0.5] * (F - N))
    assert np.abs(avg - ident).max() < 1e-10
    # the C++ contraction and its Mathematica-array printout
    dbl = np.ones(F); dbl[N:] = 2.0
    Sh = np.linalg.inv(Eh) / np.outer(dbl, dbl)
    G = orc.macro_to_micro_strain_tensors(sim, w)
    for e in (0, sim.mesh.num_elements // 2, sim.mesh.num_elements - 1):
        out, text = hostlib.m2m_tensor(N, D, G[e], Sh)
        assert np.abs(out - M[e]).max() <= 1e-13 * np.abs(M[e]).max()
        vals = np.array([float(x) for x in re.findall(r"[-+0-9.eE]+", text)]).reshape((N,) * 4)
        assert np.allclose(vals, orc.unflatten_rank4(N, M[e]), rtol=1e-13, atol=1e-15)
        assert text.startswith("{{{{") and text.endswith("}}}}")
