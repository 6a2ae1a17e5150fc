"""CPU emulation (numpy) of exactly what the chunked matrix-free operator does on the device (csrc/matfree.inl k_mf_chunk /
k_mf_gather with the tables of csrc/setup.cu build_mf_chunks), so that its index logic is covered without a GPU:

  tables (per chunk of `chunk` consecutive elements, fixed stride S = chunk * npe, slot = i * chunk + t):
    a stable sort of the chunk's (DoF, slot) pairs in blocked order (thread t holds slots t*npe .. = local nodes of
    element t), padding of the last chunk sorted to the end;
    chunk_dof[u]   the distinct DoFs, ascending          local_idx[slot]  chunk-local DoF index of the slot
    rank[slot]     position of the slot in the sorted order       csr_ptr[u]  first sorted position of DoF u
    chunk_base     exclusive prefix of the distinct counts = index of the chunk's first partial
    inc_ptr2 / inc_list2   per DoF row its partials, ascending chunk order (stable sort of the partials by DoF)
  product:
    stage x[chunk_dof[u]] -> buf[u]; element t reads buf[local_idx[i*chunk+t]], computes ye = Ke xe (here with the
    oracle's per-element stiffness; the device uses elem_apply, checked against it by tests/test_elem_math_host.py),
    writes ye[i] to out[rank[i*chunk+t]]; partial[base+u] = sum of out[csr_ptr[u] .. csr_ptr[u+1]) in that order;
    y[row] = sum of partial[inc_list2[inc_ptr2[row] .. inc_ptr2[row+1])] in that order, masked; dot = x . y.

tests/test_matrix_free_logic.py runs it against the oracle's assembled matrix (plain and periodic DoF maps, chunk sizes
that do and do not divide the element count).

  python tools/emulate_matrix_free.py           # a quadratic-tet example, prints the reduction slots -> partials
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np

import meshfem_oracle as orc

SENT = 0xFFFFFFFF


def build_chunk_tables(elem_dof, n_dofs, chunk):
    """elem_dof: (nElems, npe) int.  Returns the dict of tables described above."""
    ne, npe = elem_dof.shape
    S = chunk * npe
    n_chunks = (ne + chunk - 1) // chunk
    chunk_dof = np.zeros((n_chunks, S), dtype=np.int64)          # 0 beyond the count: a valid address on the device
    local_idx = np.zeros((n_chunks, S), dtype=np.int64)
    rank = np.zeros((n_chunks, S), dtype=np.int64)
    csr_ptr = np.zeros((n_chunks, S + 1), dtype=np.int64)
    n_unique = np.zeros(n_chunks + 1, dtype=np.int64)
    for b in range(n_chunks):
        keys = np.full(S, SENT, dtype=np.int64)
        vals = np.zeros(S, dtype=np.int64)
        for t in range(chunk):
            e = b * chunk + t
            for i in range(npe):
                pos = t * npe + i                                # blocked arrangement of the CTA-wide sort
                vals[pos] = i * chunk + t
                if e < ne:
                    keys[pos] = elem_dof[e, i]
        order = np.argsort(keys, kind="stable")
        keys, vals = keys[order], vals[order]
        valid = keys != SENT
        head = np.ones(S, dtype=bool)
        head[1:] = keys[1:] != keys[:-1]
        flag = head & valid
        idx = np.cumsum(flag) - flag                             # exclusive scan
        u = idx + flag - 1
        for pos in range(S):
            if not valid[pos]:
                continue
            local_idx[b, vals[pos]] = u[pos]
            rank[b, vals[pos]] = pos
            if flag[pos]:
                csr_ptr[b, u[pos]] = pos
                chunk_dof[b, u[pos]] = keys[pos]
        total = int(flag.sum())
        csr_ptr[b, total] = int(valid.sum())
        n_unique[b] = total
    chunk_base = np.concatenate([[0], np.cumsum(n_unique[:-1])])
    n_partials = int(chunk_base[-1])
    partial_dof = np.concatenate([chunk_dof[b, :n_unique[b]] for b in range(n_chunks)]) if n_chunks else np.zeros(0, int)
    order = np.argsort(partial_dof, kind="stable")               # per row: ascending chunk order
    inc_list2 = order
    inc_ptr2 = np.searchsorted(partial_dof[order], np.arange(n_dofs + 1), side="left")
    return dict(chunk=chunk, S=S, n_chunks=n_chunks, chunk_dof=chunk_dof, local_idx=local_idx, rank=rank, csr_ptr=csr_ptr,
                chunk_base=chunk_base, n_partials=n_partials, inc_ptr2=inc_ptr2, inc_list2=inc_list2)


def apply_operator(tab, Ke, x, N, fixed_mask=None):
    """Ke: (nElems, npe*N, npe*N) per-element stiffness; x: (nDofs, N).  Returns (y, x.y) as the two kernels compute them."""
    ne = Ke.shape[0]
    npe = Ke.shape[1] // N
    chunk, S = tab["chunk"], tab["S"]
    partial = np.zeros((tab["n_partials"], N))
    for b in range(tab["n_chunks"]):
        base = tab["chunk_base"][b]
        nu = tab["chunk_base"][b + 1] - base
        buf = np.zeros((S, N))
        buf[:nu] = x[tab["chunk_dof"][b, :nu]]                   # stage the distinct x blocks
        out = np.zeros((S, N))
        for t in range(chunk):
            e = b * chunk + t
            if e >= ne:
                continue
            xe = np.stack([buf[tab["local_idx"][b, i * chunk + t]] for i in range(npe)]).reshape(-1)
            ye = (Ke[e] @ xe).reshape(npe, N)
            for i in range(npe):
                out[tab["rank"][b, i * chunk + t]] = ye[i]
        for u in range(nu):
            acc = np.zeros(N)
            for k in range(tab["csr_ptr"][b, u], tab["csr_ptr"][b, u + 1]):
                acc = acc + out[k]
            partial[base + u] = acc
    y = np.zeros_like(x)
    for row in range(x.shape[0]):
        acc = np.zeros(N)
        for k in range(tab["inc_ptr2"][row], tab["inc_ptr2"][row + 1]):
            acc = acc + partial[tab["inc_list2"][k]]
        y[row] = acc
    if fixed_mask is not None:
        y[fixed_mask.reshape(-1, N)] = 0.0
    return y, float((x * y).sum())


def element_matrices(mesh, D):
    return orc.per_element_stiffness(mesh.N, mesh.deg, mesh.vol, mesh.G, D)


if __name__ == "__main__":
    from util import grid_mesh
    mesh = grid_mesh(3, 2, (6, 3, 2))
    D = orc.isotropic_D(3, 200.0, 0.35)
    Ke = element_matrices(mesh, D)
    rng = np.random.default_rng(0)
    x = rng.normal(size=(mesh.num_nodes, 3))
    yref = (orc.stiffness_matrix(mesh, D) @ x.reshape(-1)).reshape(-1, 3)
    for chunk in (32, 64, 128):
        tab = build_chunk_tables(mesh.elem_nodes, mesh.num_nodes, chunk)
        y, dot = apply_operator(tab, Ke, x, 3)
        print(f"chunk {chunk}: {mesh.num_elements * 10} (element, node) slots -> {tab['n_partials']} partials, "
              f"rel err vs assembled K x = {np.linalg.norm(y - yref) / np.linalg.norm(yref):.2e}")
