// Multi-GPU plumbing: NCCL communicator owned by the handle (one process per GPU).
// Element-partitioned execution with interface sum-exchange is described in DESIGN.md;
// this file currently provides the communicator lifecycle used by the halo exchange.
#include <nccl.h>

#include "core.cuh"

using namespace mfem;

namespace mfem {
void comm_destroy(mfem_b200_ctx *c) {
    if (c->ncclComm) {
        ncclCommDestroy(static_cast<ncclComm_t>(c->ncclComm));
        c->ncclComm = nullptr;
    }
}
}  // namespace mfem

extern "C" {

int mfem_b200_comm_unique_id(void *out128) {
    if (!out128) return MFEM_B200_ERR_INVALID;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return MFEM_B200_ERR_COMM;
    memcpy(out128, &id, sizeof(id));
    return MFEM_B200_OK;
}

int mfem_b200_comm_init(mfem_b200_handle h, int n_ranks, int rank, const void *nccl_unique_id128) {
    if (!h) return MFEM_B200_ERR_INVALID;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks || !nccl_unique_id128) {
        h->err = "comm_init: bad arguments";
        return MFEM_B200_ERR_INVALID;
    }
    if (h->nElems != 0) { h->err = "comm_init must precede set_mesh"; return MFEM_B200_ERR_INVALID; }
    cudaSetDevice(h->device);
    ncclUniqueId id;
    memcpy(&id, nccl_unique_id128, sizeof(id));
    ncclComm_t comm;
    ncclResult_t r = ncclCommInitRank(&comm, n_ranks, id, rank);
    if (r != ncclSuccess) {
        h->err = std::string("ncclCommInitRank: ") + ncclGetErrorString(r);
        return MFEM_B200_ERR_COMM;
    }
    h->ncclComm = comm;
    h->nRanks = n_ranks;
    h->rank = rank;
    return MFEM_B200_OK;
}

int mfem_b200_set_global_dof_ids(mfem_b200_handle h, const int64_t *ids) {
    if (!h) return MFEM_B200_ERR_INVALID;
    (void)ids;
    h->err = "set_global_dof_ids: multi-GPU interface exchange not built yet";
    return MFEM_B200_ERR_INVALID;
}

}  // extern "C"
