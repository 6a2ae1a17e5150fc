/* mfem_b200.h -- C ABI of the B200 linear-elasticity assemble-and-solve path.
 *
 * MeshFEM has no plugin/FFI boundary of its own: its operator surface is C++
 * templates.  This header is the seam a maintainer binds instead of the
 * reference's CPU implementation; each entry point cites the reference
 * interface it replaces (paths relative to src/lib/MeshFEM/ of MeshFEM).
 * The host C++ mirror of the reference classes (include/MeshFEM/ in this repo)
 * and the ctypes binding (meshfem_b200/capi.py) both call ONLY these
 * functions.
 *
 * Conventions
 *   - every function returns 0 on success, a negative mfem_b200_status
 *     otherwise; mfem_b200_last_error(h) describes the last failure.
 *   - all pointers are HOST pointers to caller-owned buffers; device memory
 *     is owned by the handle.  One handle <-> one CUDA device and stream;
 *     a handle is not thread-safe (mirrors the reference's Simulator).
 *   - Real is double everywhere (Types.hh:8).  Per-node / per-DoF vector
 *     fields are flat, index N*node + c  (Fields.hh:46-50).
 *   - there is NO CPU fallback: if no CUDA device is usable, create() fails.
 */
#ifndef MFEM_B200_H
#define MFEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mfem_b200_ctx *mfem_b200_handle;

typedef enum {
    MFEM_B200_OK = 0,
    MFEM_B200_ERR_INVALID = -1,       /* bad argument / call order                     */
    MFEM_B200_ERR_CUDA = -2,          /* CUDA runtime failure                          */
    MFEM_B200_ERR_NEG_VOLUME = -3,    /* "Mesh has negatively oriented elements." LinearElasticity.hh:465-472 */
    MFEM_B200_ERR_ALREADY_FIXED = -4, /* "Variable already fixed." SparseMatrices.hh:2432 */
    MFEM_B200_ERR_BAD_RHS = -5,       /* "Bad RHS" SparseMatrices.hh:2521              */
    MFEM_B200_ERR_NOT_SPD = -6,       /* PCG breakdown p'Ap <= 0 (CHOLMOD: not positive definite, SparseMatrices.hh:2010) */
    MFEM_B200_ERR_NO_CONVERGE = -7,   /* max iterations reached                        */
    MFEM_B200_ERR_NAN = -8,           /* NaN/Inf in the iteration                      */
    MFEM_B200_ERR_COMM = -9           /* NCCL failure                                  */
} mfem_b200_status;

typedef struct {
    int32_t iterations;       /* PCG iterations of the last right-hand side           */
    int32_t converged;        /* 1 if ||r||_2 <= rtol ||b||_2                         */
    double rel_residual;      /* final ||r||_2 / ||b||_2 (recurrence residual)        */
    double seconds;           /* device time of the PCG loop (CUDA events)            */
    double spmv_seconds;      /* device time of one stand-alone SpMV launch measured after the solve (0 if not measured) */
} mfem_b200_solve_info;

/* ---- lifetime ------------------------------------------------------------------- */
/* Replaces constructing LinearElasticity::Simulator's device-side state.             */
int mfem_b200_create(int device, mfem_b200_handle *out);
int mfem_b200_destroy(mfem_b200_handle h);
const char *mfem_b200_last_error(mfem_b200_handle h);
/* number of CUDA devices visible, or a negative status                               */
int mfem_b200_device_count(void);

/* Multi-GPU: element-partitioned execution, one process per GPU.  nccl_unique_id is the
 * 128-byte ncclUniqueId obtained on rank 0 (mfem_b200_comm_unique_id) and broadcast by the
 * caller (torch.distributed / MPI / files).  Must be called before set_mesh; after it,
 * set_mesh receives this rank's LOCAL sub-mesh and set_interface its shared DoFs.       */
int mfem_b200_comm_unique_id(void *out128);
int mfem_b200_comm_init(mfem_b200_handle h, int n_ranks, int rank, const void *nccl_unique_id128);
/* Put h on the communicator of `parent` (same process, same device) instead of creating another one:
 * a long-lived parent handle plays the role of the process group; parent must outlive h.      */
int mfem_b200_comm_share(mfem_b200_handle h, mfem_b200_handle parent);
/* Peer window (one NVSwitch box, <= 8 ranks): after comm_init every rank allocates a window of device memory
 * (comm_window_handle returns its 64-byte CUDA IPC handle), the caller gathers the handles of all ranks in rank order
 * and every rank maps them (comm_window_open).  From then on the small collectives of the PCG iteration -- the interface
 * sum-exchange, the all-reduces of the dot products / coarse residuals, the all-gather of the row-split dense level --
 * are single kernels of this library that store into the peers' windows over NVLink and synchronise through flag words
 * (csrc/comm.cu), not NCCL calls; NCCL keeps the large set-up all-reduce and is the fallback if IPC mapping is refused
 * on any rank (comm_uses_peer_window then returns 0) or option "comm_p2p" is 0.                                  */
int mfem_b200_comm_window_handle(mfem_b200_handle h, void *out64);
int mfem_b200_comm_window_open(mfem_b200_handle h, const void *handles_rank_order);
int mfem_b200_comm_uses_peer_window(mfem_b200_handle h);

/* ---- options ("reorder" before set_mesh; the others at any time between solves) ------- */
/* "reorder": 1 (default) renumbers DoFs along a space-filling curve inside the handle;
 *            every ABI function still speaks the caller's numbering.
 * "assembly": 0 = block-owner (default: every BSR block summed by one thread from its sorted
 *                 contribution list and written exactly once, coalesced),
 *             1 = graph-coloured element scatter (read-modify-write, no atomics),
 *             2 = owner-gather by DoF row (first-generation kernel, kept for A/B).
 * "coarse_aggregates": -1 (default) = multilevel aggregation preconditioner sized from the problem (block-Jacobi alone
 *             below 30 k DoFs or when the coarse matrix is not SPD), 0 = block-Jacobi only, S > 0 = at most S large
 *             aggregates (near-cubic boxes over the bounding box of the DoFs, rigid-body modes, dense level inverted
 *             explicitly); "coarse_fine_nodes": nodes per small (level-1) aggregate, default 64, 0 = none
 *             (csrc/coarse.inl; on several GPUs set the same values on every rank; may be changed between solves).
 * "matrix_free": what the Krylov loop multiplies with.  -1 (default) = for quadratic tetrahedra of a mesh given through
 *             set_mesh the product K*p is evaluated from the mesh (csrc/matfree.inl: what applyStiffnessMatrix,
 *             LinearElasticity.hh:801-823, computes; no atomics, bit-reproducible), the stored block-CSR matrix
 *             otherwise; 0 = always the stored matrix; 1 = the mesh-based operator for every element type.  Matrices from
 *             set_matrix_triplets are always multiplied as stored.  A/B variants of the operator: "mf_chunked" (1 = per-chunk
 *             partial sums, default; 0 = one slot per (element, node)), "mf_chunk_elems" (32 / 64 / 128), "mf_chunk_warps"
 *             (12 / 16 / 20), "mf_gather_lanes" (0 auto / 1 / 4 / 8), "mf_gather_policy" (0..3), "mf_slot_pad", "mf_elem_order".
 * "spmv_kernel", "spmv_lanes", "spmv_prefetch", "spmv_min_blocks": A/B variants of the stored-matrix SpMV (spmv_kernel 6:
 *             mfem_b200_spmv evaluates the mesh-based operator instead -- parity tests); "graph" (CUDA graph of the
 *             iteration, default 1), "batch_rhs" (flatLen(N) right-hand sides as one batched PCG when the levels are
 *             off), "comm_p2p" (N ranks: peer-window collectives, default 1). */
int mfem_b200_set_option(mfem_b200_handle h, const char *name, int64_t value);

/* ---- mesh ------------------------------------------------------------------------ */
/* Replaces FEMMesh(elems, vertices) + Simulator ctor (FEMMesh.inl:11-82,
 * LinearElasticity.hh:460-473).  dim in {2,3}, degree in {1,2}.  nodes: [n_nodes*dim].
 * elem_nodes: [n_elems * nodesPerElem] in the reference local order (Simplex.hh:31-46:
 * vertices, then edge nodes (0,1)(1,2)(2,0)(0,3)(2,3)(1,3)); only the first dim+1 nodes'
 * coordinates define the (straight-sided) element (FEMMesh.hh:228-233).
 * dof_for_node: NULL (identity) or PeriodicCondition::periodicDoFsForNodes()
 * (BoundaryConditions.hh:457-561) with n_dofs distinct values 0..n_dofs-1.
 * Fails with MFEM_B200_ERR_NEG_VOLUME if any element volume is < 0.                  */
int mfem_b200_set_mesh(mfem_b200_handle h, int dim, int degree, int64_t n_nodes, const double *nodes,
                       int64_t n_elems, const int32_t *elem_nodes, const int64_t *dof_for_node,
                       int64_t n_dofs);
/* Multi-GPU only, after set_mesh: the interface of this rank's element partition
 * (include/MeshFEM/Partition.hh).  For neighbour q (n_neighbors of them, ranks ascending) the
 * LOCAL DoFs shared with it are shared_local_dofs[neighbor_offsets[q] .. neighbor_offsets[q+1]),
 * listed in the same order (ascending global id) on both sides; owned[d] = 1 iff this rank is
 * the lowest rank sharing local DoF d (dot products count owned DoFs only).  Right-hand sides
 * given to solve() must be CONSISTENT (the full global value on every sharer); the returned u is
 * consistent too.                                                                        */
int mfem_b200_set_interface(mfem_b200_handle h, int n_neighbors, const int32_t *neighbor_ranks,
                            const int64_t *neighbor_offsets, const int32_t *shared_local_dofs,
                            const uint8_t *owned);

/* ---- material -------------------------------------------------------------------- */
/* D: flattened elasticity tensor, row-major flat x flat (flat = 3 in 2D, 6 in 3D),
 * Voigt order xx,yy,zz,yz,xz,xy (Flattening.hh:47-60, ElasticityTensor.hh:274-289).
 * constant  <-> _HMG static material (LinearElasticity.hh:31-44);
 * per_element <-> ETensorStoreGetter (LinearElasticity.hh:20-29).                     */
int mfem_b200_set_material_constant(mfem_b200_handle h, const double *D);
int mfem_b200_set_material_per_element(mfem_b200_handle h, const double *D_per_elem);

/* ---- assembly -------------------------------------------------------------------- */
/* Replaces Simulator::m_assembleStiffnessMatrix + TripletMatrix::sumRepeated
 * (LinearElasticity.hh:1408-1466, SparseMatrices.hh:280-374): K in DoF space, stored as
 * full block-CSR (dim x dim blocks).  The sparsity pattern is built on first call and
 * cached; later calls (new material / node positions) only recompute values.           */
int mfem_b200_assemble(mfem_b200_handle h);
int mfem_b200_get_bsr_sizes(mfem_b200_handle h, int64_t *n_block_rows, int64_t *nnz_blocks);
/* Export in the CALLER's numbering: rowptr[nb+1], colidx[nnzb] sorted per row,
 * vals[nnzb*dim*dim] row-major blocks.                                                */
int mfem_b200_get_bsr(mfem_b200_handle h, int64_t *rowptr, int32_t *colidx, double *vals);
/* TripletMatrix::dumpBinary layout of the summed upper triangle
 * (SparseMatrices.hh:623-645: uint64 nnz | uint64 rows[] | uint64 cols[] | double vals[]);
 * Simulate_cli --dumpMatrix.                                                           */
int mfem_b200_dump_upper_triplets(mfem_b200_handle h, const char *path);
/* Update node positions (Simulator::updateMeshNodePositions); pattern is kept.         */
int mfem_b200_set_node_positions(mfem_b200_handle h, const double *nodes);

/* ---- a matrix assembled elsewhere -------------------------------------------------- */
/* SPSDSystem(K) / setConstrained(K, C = empty) (SparseMatrices.hh:2321-2348) for a symmetric positive
 * (semi-)definite matrix the caller assembled itself: COO triplets over n_vars scalar variables
 * ordered block_dim*DoF + component (block_dim in {2,3}); repeated entries are summed
 * (TripletMatrix::sumRepeated); upper_triangle_only != 0: only i <= j given (the reference's storage),
 * mirrored here.  Replaces any mesh of the handle; fix_variables / solve / spmv / get_bsr work on it,
 * mesh-based calls (assemble, apply_K, loads, strains) fail.                            */
int mfem_b200_set_matrix_triplets(mfem_b200_handle h, int block_dim, int64_t n_vars, int64_t nnz, const int64_t *rows,
                                  const int64_t *cols, const double *vals, int upper_triangle_only);

/* ---- constraints + solve (SPSDSystem) -------------------------------------------- */
/* SPSDSystem::fixVariables (SparseMatrices.hh:2389-2500): scalar variable indices
 * dim*DoF+c and the values they are fixed to (values may be NULL = 0).  Cumulative;
 * fixing a variable twice fails with MFEM_B200_ERR_ALREADY_FIXED.                      */
int mfem_b200_fix_variables(mfem_b200_handle h, int64_t n, const int64_t *vars, const double *values);
int mfem_b200_clear_fixed_variables(mfem_b200_handle h);
/* SPSDSystem::solve (SparseMatrices.hh:2516-2606) with the CHOLMOD factorisation
 * replaced by block-Jacobi PCG: for each of nrhs right-hand sides f (length dim*n_dofs,
 * stored one after another) returns the full-length u with fixed variables at their
 * values.  rtol on ||r||_2/||b||_2 of the reduced system; info may be NULL, else
 * info[nrhs].                                                                          */
int mfem_b200_solve(mfem_b200_handle h, int nrhs, const double *f, double *u, double rtol,
                    int max_iters, mfem_b200_solve_info *info);

/* ---- operators around the solve -------------------------------------------------- */
/* Simulator::applyStiffnessMatrix (LinearElasticity.hh:801-823): raw K on per-NODE
 * fields (ignores periodic DoFs and Dirichlet conditions).                              */
int mfem_b200_apply_K(mfem_b200_handle h, const double *u_nodes, double *Ku_nodes);
/* y = K x on per-DoF fields with the assembled block-CSR (no masking).                 */
int mfem_b200_spmv(mfem_b200_handle h, const double *x_dofs, double *y_dofs);
/* Simulator::constantStrainLoad (LinearElasticity.hh:551-562, 135-162); eps flattened. */
int mfem_b200_const_strain_load(mfem_b200_handle h, const double *eps_flat, double *f_dofs);
/* Simulator::averageStrainField / averageStressField (LinearElasticity.hh:528-549):
 * per-element flattened averages [n_elems*flat]; either output may be NULL.           */
int mfem_b200_avg_strain_stress(mfem_b200_handle h, const double *u_nodes, double *strain, double *stress);
/* Per-element volumes and negative-volume count (Simulator ctor check).                */
int mfem_b200_get_volumes(mfem_b200_handle h, double *vol);

/* ---- discrete shape derivatives (csrc/shape.cu) ------------------------------------ */
/* Replace Simulator::applyDeltaStiffnessMatrix (LinearElasticity.hh:1301-1330), deltaConstantStrainLoad (:1333-1348)
 * and deltaAverageStrainField (:1365-1375): the change of K u, of constantStrainLoad(eps) and of the element-averaged
 * strain under a perturbation delta_p of the VERTEX positions [n_vertices*dim] (the first n_vertices nodes are the
 * vertices, FEMMesh.inl:17-37).  u_nodes / delta_u_nodes: per node [n_nodes*dim]; out_dofs: per DoF [n_dofs*dim];
 * strain: [n_elems*flat].  One element loop each on the device (FP64 atomics for the per-DoF sums).               */
int mfem_b200_apply_delta_K(mfem_b200_handle h, const double *u_nodes, const double *delta_p, int64_t n_vertices, double *out_dofs);
int mfem_b200_delta_const_strain_load(mfem_b200_handle h, const double *eps_flat, const double *delta_p, int64_t n_vertices,
                                      double *out_dofs);
int mfem_b200_delta_avg_strain(mfem_b200_handle h, const double *u_nodes, const double *delta_u_nodes, const double *delta_p,
                               int64_t n_vertices, double *strain);

/* ---- timers (BENCHMARK_REPORT sections, GlobalBenchmark.hh:8-58) ------------------ */
/* Seconds of device time (CUDA events) accumulated under a section name, e.g.
 * "Pattern", "Assemble System", "Elasticity Solve", "SpMV".  Returns -1.0 if unknown.  */
double mfem_b200_get_timer(mfem_b200_handle h, const char *section);
int mfem_b200_reset_timers(mfem_b200_handle h);
/* Device memory of destroyed handles is kept in a process-wide cache and reused by later handles (cudaMalloc / cudaFree
 * of multi-GB arrays cost hundreds of milliseconds).  This call returns the cached blocks to the driver; the environment
 * variable MFEM_B200_POOL=0 disables caching altogether.                                                            */
int mfem_b200_release_cached_memory(void);
/* number of kernel launches issued by this handle since creation / last reset          */
int64_t mfem_b200_launch_count(mfem_b200_handle h);

/* ---- device-resident benchmarking hooks ------------------------------------------ */
/* Run `iters` PCG iterations' worth of SpMV on the assembled matrix with device-resident
 * vectors and return the mean device seconds per launch (CUDA events on the handle's
 * stream).  Used by bench.py for the roofline of the dominant kernel.                  */
int mfem_b200_time_spmv(mfem_b200_handle h, int iters, double *seconds_per_launch);
/* The same for the product the PCG actually launches per iteration, y = mask(K p) with the fused p.y: the mesh-based
 * (matrix-free) operator of csrc/matfree.inl where option "matrix_free" selects it (*matrix_free = 1; seconds_parts[0..1]
 * = its element kernel and its gather kernel timed alone), else the stored-matrix SpMV (*matrix_free = 0, parts 0).
 * seconds_parts and matrix_free may be NULL.                                                                         */
int mfem_b200_time_operator(mfem_b200_handle h, int iters, double *seconds_per_product, double *seconds_parts,
                            int *matrix_free);
/* Diagnostics: z = M^-1 r and *rz = r.z for the preconditioner the next solve would use (block-Jacobi alone, or with the
 * aggregation levels of "coarse_aggregates" / "coarse_fine_nodes"), r masked on the fixed variables first; per-DoF
 * vectors in the caller's numbering.  Runs the PCG's own start-up kernels, so the parity tests can compare the operator
 * inside the iteration with its numpy restatement (tools/emulate_multilevel.py) term by term.                        */
int mfem_b200_apply_preconditioner(mfem_b200_handle h, const double *r, double *z, double *rz);
/* Diagnostics: a named array of the aggregation levels as doubles ("sizes" = S1, S2, R, n1, aggBase, level1; "agg1",
 * "Y1" in the handle's internal DoF order with "int2ext" the map to the caller's; "shift", "B1inv", "Einv", "y1", "y2").
 * out may be NULL to query the length *n.                                                                           */
int mfem_b200_get_coarse_array(mfem_b200_handle h, const char *name, double *out, int64_t capacity, int64_t *n);

#ifdef __cplusplus
}
#endif
#endif /* MFEM_B200_H */
