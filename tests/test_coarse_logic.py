"""CPU checks of the logic of the two-level preconditioner kernels (meshfem_b200/csrc/coarse.inl), which cannot run
here: lane-level emulations of the two warp segmented reductions (equal-key variant of the single-GPU kernels,
head-flag variant of the multi-GPU kernels) against a direct per-key sum, and the rigid-mode operators
(coarse_R / coarse_Rt restated line by line) against each other and against the infinitesimal rigid motions."""
import numpy as np

L = np.arange(32)


def _shfl_down(v, o):
    out = v.copy()
    out[:32 - o] = v[o:]
    return out                      # lanes >= 32-o keep their own value (CUDA semantics)


def _shfl_up1(v):
    out = v.copy()
    out[1:] = v[:-1]
    return out


def _equal_keys(keys, vals, active):               # k_coarse_matrix / k_coarse_restrict
    aj = np.where(active, keys, -1 - L)
    C = np.where(active, vals, 0.0)
    for o in (1, 2, 4, 8, 16):
        take = (L + o < 32) & (_shfl_down(aj, o) == aj)
        C = np.where(take, C + _shfl_down(C, o), C)
    head = active & ((L == 0) | (_shfl_up1(aj) != aj))
    out = {}
    for l in L[head]:
        out[aj[l]] = out.get(aj[l], 0.0) + C[l]
    return out


def _head_flags(keys, vals, active):               # k_coarse_matrix_idx / k_coarse_restrict_idx (coarse_same_run)
    aj = np.where(active, keys, -1 - L)
    C = np.where(active, vals, 0.0)
    head = (L == 0) | (_shfl_up1(aj) != aj)
    heads = sum(1 << int(l) for l in L[head])
    for o in (1, 2, 4, 8, 16):
        take = np.array([(l + o < 32) and (((heads >> (l + 1)) & ((1 << o) - 1)) == 0) for l in L])
        C = np.where(take, C + _shfl_down(C, o), C)
    out = {}
    for l in L[active & head]:
        out[aj[l]] = out.get(aj[l], 0.0) + C[l]
    return out


def _direct(keys, vals, active):
    out = {}
    for k, v, a in zip(keys, vals, active):
        if a:
            out[k] = out.get(k, 0.0) + v
    return out


def _same(r, d):
    return r.keys() == d.keys() and all(abs(r[k] - d[k]) < 1e-12 for k in d)


def test_warp_segmented_reductions():
    rng = np.random.default_rng(0)
    for _ in range(500):
        active = L < rng.integers(1, 33)
        vals = rng.random(32)
        keys = np.sort(rng.integers(0, rng.integers(1, 12), size=32))          # contiguous runs (sorted columns)
        d = _direct(keys, vals, active)
        assert _same(_equal_keys(keys, vals, active), d) and _same(_head_flags(keys, vals, active), d)
        keys = rng.integers(0, rng.integers(1, 6), size=32)                    # interleaved foreign aggregates
        assert _same(_head_flags(keys, vals, active), _direct(keys, vals, active))


def _R(N, y, c):                                    # coarse_R
    if N == 3:
        return np.array([c[0] + y[2] * c[4] - y[1] * c[5], c[1] - y[2] * c[3] + y[0] * c[5], c[2] + y[1] * c[3] - y[0] * c[4]])
    return np.array([c[0] - y[1] * c[2], c[1] + y[0] * c[2]])


def _Rt(N, y, v):                                   # coarse_Rt
    if N == 3:
        return np.array([v[0], v[1], v[2], y[1] * v[2] - y[2] * v[1], y[2] * v[0] - y[0] * v[2], y[0] * v[1] - y[1] * v[0]])
    return np.array([v[0], v[1], y[0] * v[1] - y[1] * v[0]])


def test_rigid_mode_operators():
    rng = np.random.default_rng(1)
    for N, M in ((3, 6), (2, 3)):
        for _ in range(20):
            y, c, v = rng.standard_normal(N), rng.standard_normal(M), rng.standard_normal(N)
            assert abs(_R(N, y, c) @ v - c @ _Rt(N, y, v)) < 1e-12                 # transposes of each other
            # translation + infinitesimal rotation omega x y
            if N == 3:
                assert np.allclose(_R(N, y, c), c[:3] + np.cross(c[3:], y))
            else:
                assert np.allclose(_R(N, y, c), c[:2] + c[2] * np.array([-y[1], y[0]]))


def _choose_boxes(L, budget):                       # coarse_choose_boxes
    L = np.asarray(L, float)
    nz = L > 0
    h = (np.prod(L[nz]) / max(budget, 1)) ** (1.0 / nz.sum()) if nz.any() else 1.0
    b = [int(max(1, np.floor(l / h + 0.5))) if l > 0 else 1 for l in L]
    while np.prod(b) > budget:
        k = int(np.argmax(b))
        if b[k] == 1:
            break
        b[k] -= 1
    return b


def test_box_grid_choice():
    """Near-cubic boxes within the budget; flat directions get one layer."""
    assert _choose_boxes((220, 44, 44), 2048) == [37, 7, 7]
    assert _choose_boxes((20, 4, 4), 128) == [14, 3, 3]
    assert _choose_boxes((40, 8), 96) == [22, 4]
    assert _choose_boxes((1, 1, 0), 16) == [4, 4, 1]
    rng = np.random.default_rng(2)
    for _ in range(200):
        L = rng.random(3) * 10 + 0.1
        budget = int(rng.integers(1, 5000))
        b = _choose_boxes(L, budget)
        assert np.prod(b) <= budget and min(b) >= 1
        if budget >= 64 and np.prod(b) > 8:            # box edge lengths within a factor ~2 of each other away from the 1-layer limit
            edges = L / np.array(b)
            multi = np.array(b) > 1
            if multi.sum() >= 2:
                assert edges[multi].max() / edges[multi].min() < 2.5


def test_emulated_multilevel_pcg_needs_far_fewer_iterations():
    """tools/emulate_multilevel.py: the device algorithm (nested box grids, masked rigid modes, level-1 6x6 blocks with
    dropped dead modes, regularised dense E, additive terms with r.z = r.B0^-1 r + c1.y1 + c2.y2, level 1 <-> 2 through
    the shift formulas) restated in numpy, on a small linear-tet cantilever: large boxes alone, then with level 1."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from emulate_multilevel import run
    it0, it2, it3 = run(3, 1, (16, 4, 4), 8, 12, verbose=False)
    assert it2 < 0.6 * it0 and it3 < it2, (it0, it2, it3)


def test_level_transfer_formulas_compose_to_the_large_box_modes():
    """coarse_shift_prolong / coarse_shift_restrict (restated in tools/emulate_multilevel.py): the rigid modes of a large
    box are exact combinations of the modes of its small boxes, R1(y1) P2(d) = R2(y1 + d), and the restriction is the
    transpose; the dropping Cholesky inverse (k_coarse_invert1) inverts the live principal submatrix."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import emulate_multilevel as em
    rng = np.random.default_rng(3)
    for N, M in ((3, 6), (2, 3)):
        y1, d, Yc, v = rng.standard_normal((5, N)), rng.standard_normal((5, N)), rng.standard_normal((5, M)), rng.standard_normal((5, N))
        q = em.shift_prolong(d, Yc)
        R1, R2 = em.rigid(y1), em.rigid(y1 + d)
        assert np.allclose(np.einsum("ncm,nm->nc", R1, q), np.einsum("ncm,nm->nc", R2, Yc))
        c1 = np.einsum("ncm,nc->nm", R1, v)
        assert np.allclose(em.shift_restrict(d, c1), np.einsum("ncm,nc->nm", R2, v))
    G = rng.standard_normal((6, 6)); S = G @ G.T + 0.1 * np.eye(6)
    assert np.abs(em.dropping_cholesky_inverse(S) @ S - np.eye(6)).max() < 1e-12
    S[2, :] = 0; S[:, 2] = 0
    B = em.dropping_cholesky_inverse(S); live = [0, 1, 3, 4, 5]
    assert np.abs(B[np.ix_(live, live)] @ S[np.ix_(live, live)] - np.eye(5)).max() < 1e-12 and np.abs(B[2]).max() == 0.0
