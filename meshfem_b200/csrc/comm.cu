// Multi-GPU plumbing (one process per GPU): NCCL communicator owned by the handle, the
// interface ("halo") sum-exchange that completes partial sums on DoFs shared between element
// partitions, and the scalar all-reduce behind the PCG dot products.
//
// New design -- the reference is single-process.  Non-overlapping element
// partition with SHARED interface DoFs: a rank's local K holds partial sums on interface rows, so
// after every local SpMV each pair of neighbouring ranks swaps the values of the DoFs they share
// and adds what it receives (ncclSend/ncclRecv grouped per neighbour, on the solver's stream).
// Dot products count every DoF once through the `owned` mask (owner = lowest sharing rank).
#include <nccl.h>

#include <cstring>

#include <algorithm>

#include "core.cuh"

namespace mfem {

struct Halo {
    std::vector<int> ranks;              // neighbour ranks
    std::vector<int64_t> offsets;        // [nNeighbors+1] into idx
    DevBuf<int32_t> idx;                 // internal DoF ids shared with each neighbour, per-neighbour segments
    DevBuf<uint8_t> owned;               // [nDofs] internal numbering
    DevBuf<uint8_t> shared;              // [nDofs] 1 = DoF appears in some neighbour's list
    // receive side, per DISTINCT shared DoF: the slots of recvBuf that carry its partial sums (a DoF on a partition
    // edge is shared with several neighbours), so that one launch adds everything in a fixed order
    DevBuf<int32_t> uIdx;                // [nUnique] internal DoF id
    DevBuf<int64_t> uPtr;                // [nUnique+1] into uPos
    DevBuf<int32_t> uPos;                // [total] slot in recvBuf (units of `width` doubles)
    int64_t nUnique = 0;
    DevBuf<double> sendBuf, recvBuf;     // total * maxWidth doubles
    int64_t total = 0;
    int maxWidth = 9;
};

#define MFEM_NCCL(call)                                                                              \
    do {                                                                                             \
        ncclResult_t r_ = (call);                                                                    \
        if (r_ != ncclSuccess)                                                                       \
            throw mfem::CudaError(MFEM_B200_ERR_COMM, std::string(#call) + ": " + ncclGetErrorString(r_)); \
    } while (0)

__global__ void k_halo_pack(int64_t n, int width, const int32_t *__restrict__ idx, const double *__restrict__ vec,
                            double *__restrict__ buf) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n * width) return;
    const int64_t k = t / width;
    const int c = (int)(t - k * width);
    buf[t] = vec[(int64_t)idx[k] * width + c];
}
// one launch for all neighbours: a thread owns one component of one distinct shared DoF and adds its received
// partial sums in list order (deterministic; no write conflicts)
__global__ void k_halo_unpack_add(int64_t nUnique, int width, const int32_t *__restrict__ uIdx, const int64_t *__restrict__ uPtr,
                                  const int32_t *__restrict__ uPos, const double *__restrict__ buf, double *__restrict__ vec) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nUnique * width) return;
    const int64_t u = t / width;
    const int c = (int)(t - u * width);
    double s = 0.0;
    for (int64_t k = uPtr[u]; k < uPtr[u + 1]; ++k) s += buf[(int64_t)uPos[k] * width + c];
    vec[(int64_t)uIdx[u] * width + c] += s;
}
__global__ void k_mark_shared(int64_t n, const int32_t *__restrict__ idx, uint8_t *__restrict__ shared) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) shared[idx[t]] = 1;
}
__global__ void k_map_shared(int64_t n, const int32_t *__restrict__ localIdx, const int32_t *__restrict__ ext2int,
                             int32_t *__restrict__ out) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) out[t] = ext2int[localIdx[t]];
}
__global__ void k_map_owned(int64_t n, const uint8_t *__restrict__ ownedExt, const int32_t *__restrict__ int2ext,
                            uint8_t *__restrict__ out) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < n) out[t] = ownedExt[int2ext[t]];
}

void comm_destroy(mfem_b200_ctx *c) {
    delete c->halo;
    c->halo = nullptr;
    if (c->ncclComm && c->ownsComm) ncclCommDestroy(static_cast<ncclComm_t>(c->ncclComm));
    c->ncclComm = nullptr;
}

const uint8_t *halo_owned(mfem_b200_ctx *c) {
    MFEM_REQUIRE(c->halo && c->halo->owned.n == (size_t)c->nDofs, MFEM_B200_ERR_INVALID,
                 "multi-GPU handle without interface description: call mfem_b200_set_interface after set_mesh");
    return c->halo->owned;
}

void halo_exchange_add(mfem_b200_ctx *c, double *vec, int width) {
    if (c->nRanks <= 1) return;
    MFEM_REQUIRE(c->halo, MFEM_B200_ERR_INVALID, "multi-GPU handle without interface description");
    Halo &h = *c->halo;
    if (h.total == 0) return;
    MFEM_REQUIRE(width <= h.maxWidth, MFEM_B200_ERR_INVALID, "halo width too large");
    cudaStream_t s = c->stream;
    ncclComm_t comm = static_cast<ncclComm_t>(c->ncclComm);
    k_halo_pack<<<grid_for(h.total * width, 256), 256, 0, s>>>(h.total, width, h.idx, vec, h.sendBuf);
    c->launches++;
    MFEM_NCCL(ncclGroupStart());
    for (size_t q = 0; q < h.ranks.size(); ++q) {
        const int64_t off = h.offsets[q] * width, cnt = (h.offsets[q + 1] - h.offsets[q]) * width;
        MFEM_NCCL(ncclSend(h.sendBuf.p + off, (size_t)cnt, ncclDouble, h.ranks[q], comm, s));
        MFEM_NCCL(ncclRecv(h.recvBuf.p + off, (size_t)cnt, ncclDouble, h.ranks[q], comm, s));
    }
    MFEM_NCCL(ncclGroupEnd());
    k_halo_unpack_add<<<grid_for(h.nUnique * width, 256), 256, 0, s>>>(h.nUnique, width, h.uIdx, h.uPtr, h.uPos, h.recvBuf, vec);
    c->launches++;
}

const uint8_t *halo_shared(mfem_b200_ctx *c) {
    MFEM_REQUIRE(c->halo && c->halo->shared.n == (size_t)c->nDofs, MFEM_B200_ERR_INVALID,
                 "multi-GPU handle without interface description: call mfem_b200_set_interface after set_mesh");
    return c->halo->shared;
}

void allreduce_sum(mfem_b200_ctx *c, const double *in, double *out, int n) {
    if (c->nRanks <= 1) {
        if (in != out) MFEM_CUDA(cudaMemcpyAsync(out, in, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
        return;
    }
    MFEM_NCCL(ncclAllReduce(in, out, (size_t)n, ncclDouble, ncclSum, static_cast<ncclComm_t>(c->ncclComm), c->stream));
}

// every rank contributes buf[rank*count .. (rank+1)*count) and receives all slices (in place)
void allgather_inplace(mfem_b200_ctx *c, double *buf, int count) {
    if (c->nRanks <= 1) return;
    MFEM_NCCL(ncclAllGather(buf + (size_t)c->rank * count, buf, (size_t)count, ncclDouble, static_cast<ncclComm_t>(c->ncclComm), c->stream));
}

}  // namespace mfem

using namespace mfem;

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
    h->ownsComm = true;
    h->nRanks = n_ranks;
    h->rank = rank;
    return MFEM_B200_OK;
}

// A further handle of the same process on the communicator of `parent` (a long-lived "process group"
// handle): no second ncclCommInitRank.  `parent` must outlive h; the two must not run collectives
// concurrently (one stream at a time per communicator).
int mfem_b200_comm_share(mfem_b200_handle h, mfem_b200_handle parent) {
    if (!h || !parent) return MFEM_B200_ERR_INVALID;
    if (!parent->ncclComm || parent->nRanks < 2) { h->err = "comm_share: parent has no communicator"; return MFEM_B200_ERR_INVALID; }
    if (h->nElems != 0) { h->err = "comm_share must precede set_mesh"; return MFEM_B200_ERR_INVALID; }
    if (h->device != parent->device) { h->err = "comm_share: handles live on different devices"; return MFEM_B200_ERR_INVALID; }
    h->ncclComm = parent->ncclComm;
    h->ownsComm = false;
    h->nRanks = parent->nRanks;
    h->rank = parent->rank;
    return MFEM_B200_OK;
}

int mfem_b200_set_interface(mfem_b200_handle h, int n_neighbors, const int32_t *neighbor_ranks,
                            const int64_t *neighbor_offsets, const int32_t *shared_local_dofs, const uint8_t *owned) {
    if (!h) return MFEM_B200_ERR_INVALID;
    try {
        MFEM_CUDA(cudaSetDevice(h->device));
        MFEM_REQUIRE(h->nElems > 0, MFEM_B200_ERR_INVALID, "set_interface: set the mesh first");
        MFEM_REQUIRE(h->nRanks > 1 && h->ncclComm, MFEM_B200_ERR_INVALID, "set_interface: comm_init was not called");
        MFEM_REQUIRE(n_neighbors >= 0 && owned && (n_neighbors == 0 || (neighbor_ranks && neighbor_offsets && shared_local_dofs)),
                     MFEM_B200_ERR_INVALID, "set_interface: bad arguments");
        delete h->halo;
        h->halo = new Halo();
        Halo &H = *h->halo;
        H.maxWidth = h->N * std::max(h->N, flat_len(h->N));    // diagonal blocks (N*N) or a batch of flatLen(N) vectors
        H.ranks.assign(neighbor_ranks, neighbor_ranks + n_neighbors);
        H.offsets.assign(1, 0);
        if (n_neighbors) H.offsets.assign(neighbor_offsets, neighbor_offsets + n_neighbors + 1);
        H.total = H.offsets.back();
        for (int64_t k = 0; k < H.total; ++k)
            MFEM_REQUIRE(shared_local_dofs[k] >= 0 && shared_local_dofs[k] < h->nDofs, MFEM_B200_ERR_INVALID,
                         "set_interface: shared DoF out of range");
        cudaStream_t s = h->stream;
        H.shared.alloc((size_t)h->nDofs);
        MFEM_CUDA(cudaMemsetAsync(H.shared, 0, H.shared.bytes(), s));
        if (H.total) {
            DevBuf<int32_t> tmp((size_t)H.total);
            MFEM_CUDA(cudaMemcpyAsync(tmp, shared_local_dofs, tmp.bytes(), cudaMemcpyHostToDevice, s));
            H.idx.alloc((size_t)H.total);
            k_map_shared<<<grid_for(H.total, 256), 256, 0, s>>>(H.total, tmp, h->ext2int, H.idx);
            k_mark_shared<<<grid_for(H.total, 256), 256, 0, s>>>(H.total, H.idx, H.shared);
            H.sendBuf.alloc((size_t)H.total * H.maxWidth);
            H.recvBuf.alloc((size_t)H.total * H.maxWidth);
            // distinct shared DoFs (caller's numbering) and the receive slots of each, in slot order
            std::vector<int32_t> order((size_t)H.total);
            for (int64_t k = 0; k < H.total; ++k) order[(size_t)k] = (int32_t)k;
            std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return shared_local_dofs[a] < shared_local_dofs[b]; });
            std::vector<int32_t> uExt;
            std::vector<int64_t> uPtr;
            for (int64_t k = 0; k < H.total; ++k) {
                const int32_t d = shared_local_dofs[order[(size_t)k]];
                if (uExt.empty() || uExt.back() != d) { uExt.push_back(d); uPtr.push_back(k); }
            }
            uPtr.push_back(H.total);
            H.nUnique = (int64_t)uExt.size();
            DevBuf<int32_t> tmpU((size_t)H.nUnique);
            MFEM_CUDA(cudaMemcpyAsync(tmpU, uExt.data(), tmpU.bytes(), cudaMemcpyHostToDevice, s));
            H.uIdx.alloc((size_t)H.nUnique);
            k_map_shared<<<grid_for(H.nUnique, 256), 256, 0, s>>>(H.nUnique, tmpU, h->ext2int, H.uIdx);
            H.uPtr.alloc(uPtr.size());
            H.uPos.alloc((size_t)H.total);
            MFEM_CUDA(cudaMemcpyAsync(H.uPtr, uPtr.data(), H.uPtr.bytes(), cudaMemcpyHostToDevice, s));
            MFEM_CUDA(cudaMemcpyAsync(H.uPos, order.data(), H.uPos.bytes(), cudaMemcpyHostToDevice, s));
            h->launches += 3;
            MFEM_CUDA(cudaStreamSynchronize(s));
        }
        DevBuf<uint8_t> oext((size_t)h->nDofs);
        MFEM_CUDA(cudaMemcpyAsync(oext, owned, oext.bytes(), cudaMemcpyHostToDevice, s));
        H.owned.alloc((size_t)h->nDofs);
        k_map_owned<<<grid_for(h->nDofs, 256), 256, 0, s>>>(h->nDofs, oext, h->int2ext, H.owned);
        h->launches += 2;
        MFEM_CUDA(cudaStreamSynchronize(s));
        MFEM_CUDA(cudaGetLastError());
        h->precondValid = false;
        h->meshVersion++;
        return MFEM_B200_OK;
    } catch (const CudaError &e) {
        h->err = e.what();
        return e.status;
    }
}

}  // extern "C"
