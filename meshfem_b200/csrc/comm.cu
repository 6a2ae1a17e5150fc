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
#include "peer.cuh"

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
    // peer-window path
    DevBuf<int32_t> nbrRankDev;          // [nNeighbors]
    DevBuf<int64_t> offsetsDev;          // [nNeighbors+1]
    DevBuf<int64_t> uWin;                // [total] sender * kHaloSegCap + slot inside the sender's segment, in uPtr order
    bool p2pFits = false;                // every neighbour segment fits its window slot
};

#define MFEM_NCCL(call)                                                                              \
    do {                                                                                             \
        ncclResult_t r_ = (call);                                                                    \
        if (r_ != ncclSuccess)                                                                       \
            throw mfem::CudaError(MFEM_B200_ERR_COMM, std::string(#call) + ": " + ncclGetErrorString(r_)); \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
// Peer window: the small collectives of the PCG iteration done by OUR kernels over NVLink peer memory instead of NCCL.
//
// Every rank owns one device buffer (the "window") that all other ranks of the box map through CUDA IPC
// (mfem_b200_comm_window_handle / _open).  A collective is ONE kernel per rank: it stores this rank's contribution
// straight into the peers' windows (P2P stores over NVLink / NVSwitch), publishes a sequence number in the peers' flag
// words (system-scope fence before it), waits until the flags of the ranks it depends on show the same sequence number,
// and finishes from LOCAL memory -- sums the R slots in rank order (all-reduce: every rank adds the same numbers in the
// same order, so all ranks hold the same bits), copies them (all-gather), or adds the neighbours' interface values
// (halo).  No launch of a communication library, no proxy thread, no ring: two NVLink hops of latency per collective.
// Data slots are double-buffered by the parity of the sequence number: a rank can run at most one collective ahead of a
// peer (it needs that peer's next flag to go further), so a slot is never overwritten while it is still being read.
// A spin that lasts longer than ~2 s sets the window's error word (the host turns it into MFEM_B200_ERR_COMM).
// NCCL stays underneath for everything large (the coarse-matrix all-reduce at set-up) and as the fallback when the
// window is not available (option comm_p2p = 0, IPC refused, a message larger than its region).
struct PeerWinOwner {                             // host side: owns the local allocation and the IPC mappings
    PeerWin w;
    void *local = nullptr;
    void *mapped[kPeerMax] = {};
    bool open = false;
    ~PeerWinOwner() {
        for (int r = 0; r < kPeerMax; ++r) if (mapped[r]) cudaIpcCloseMemHandle(mapped[r]);
        if (local) cudaFree(local);
    }
};

// all-reduce (sum) of n <= kArCap doubles; grid = R blocks: block q ships my vector to rank q, then (after everybody's
// flag arrived) reduces the q-th slice of the result.  in == out allowed.
__global__ void __launch_bounds__(1024)
k_peer_allreduce(PeerWin w, const double *in, double *out, int n, unsigned *ticket, const int *status) {
    if (status && status[1] != 0) return;
    const int q = blockIdx.x;
    const unsigned long long s = *w.seq(SET_AR) + 1;
    const int phase = (int)(s & 1);
    double *dst = w.ar(q, phase, w.rank);
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = in[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        peer_publish(w.flags(q, SET_AR) + w.rank, s);
        atomicAdd(w.arrive(), 1ull);              // this block is done reading `in` (which may alias `out`)
    }
    // everybody's contribution has landed here, and all MY blocks have read `in`: R blocks per call, s calls so far
    if ((int)threadIdx.x < w.R && !peer_wait(w.flags(w.rank, SET_AR) + threadIdx.x, s)) *w.err() = 1;
    if ((int)threadIdx.x == w.R && !peer_wait(w.arrive(), (unsigned long long)w.R * s)) *w.err() = 1;
    __syncthreads();
    const int per = (n + w.R - 1) / w.R, lo = q * per, hi = min(n, lo + per);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        double t = 0.0;
        for (int r = 0; r < w.R; ++r) t += w.ar(w.rank, phase, r)[i];
        out[i] = t;
    }
    peer_finish(w, SET_AR, s, ticket);
}

// in-place all-gather: rank r owns buf[r*count .. (r+1)*count); grid = R blocks
__global__ void __launch_bounds__(1024)
k_peer_allgather(PeerWin w, double *buf, int count, unsigned *ticket, const int *status) {
    if (status && status[1] != 0) return;
    const int q = blockIdx.x;
    const unsigned long long s = *w.seq(SET_AG) + 1;
    const int phase = (int)(s & 1);
    double *dst = w.ag(q, phase) + (size_t)w.rank * count;
    const double *src = buf + (size_t)w.rank * count;
    for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    if (threadIdx.x == 0) peer_publish(w.flags(q, SET_AG) + w.rank, s);
    if ((int)threadIdx.x < w.R && !peer_wait(w.flags(w.rank, SET_AG) + threadIdx.x, s)) *w.err() = 1;
    __syncthreads();
    if (q != w.rank) {
        const double *g = w.ag(w.rank, phase) + (size_t)q * count;
        for (int i = threadIdx.x; i < count; i += blockDim.x) buf[(size_t)q * count + i] = g[i];
    }
    peer_finish(w, SET_AG, s, ticket);
}

// interface exchange, sender side: block b packs the values of the DoFs shared with neighbour b straight into that
// neighbour's window (slot of sender = my rank) and publishes the sequence number there
__global__ void __launch_bounds__(1024)
k_peer_halo_send(PeerWin w, int nNbr, const int32_t *__restrict__ nbrRank, const int64_t *__restrict__ offsets, int width,
                 const int32_t *__restrict__ idx, const double *__restrict__ vec, unsigned *ticket, const int *status) {
    if (status && status[1] != 0) return;
    const int b = blockIdx.x;
    const unsigned long long s = *w.seq(SET_HALO) + 1;
    const int phase = (int)(s & 1);
    const int q = nbrRank[b];
    const int64_t off = offsets[b], cnt = (offsets[b + 1] - off) * width;
    double *dst = w.halo(q, phase, w.rank);
    for (int64_t t = threadIdx.x; t < cnt; t += blockDim.x) {
        const int64_t k = t / width;
        dst[t] = vec[(int64_t)idx[off + k] * width + (t - k * width)];
    }
    __syncthreads();
    if (threadIdx.x == 0) peer_publish(w.flags(q, SET_HALO) + w.rank, s);
    (void)nNbr;
    (void)ticket;
}
// receiver side: wait for every neighbour's flag, then add the received partial sums of every distinct shared DoF in
// list order; the last block advances the sequence number
__global__ void __launch_bounds__(256)
k_peer_halo_recv(PeerWin w, int nNbr, const int32_t *__restrict__ nbrRank, int64_t nUnique, int width,
                 const int32_t *__restrict__ uIdx, const int64_t *__restrict__ uPtr, const int64_t *__restrict__ uWin /* sender * kHaloSegCap + local slot */,
                 double *__restrict__ vec, unsigned *ticket, const int *status) {
    if (status && status[1] != 0) return;
    const unsigned long long s = *w.seq(SET_HALO) + 1;
    const int phase = (int)(s & 1);
    if ((int)threadIdx.x < nNbr && !peer_wait(w.flags(w.rank, SET_HALO) + nbrRank[threadIdx.x], s)) *w.err() = 1;
    __syncthreads();
    const double *base = w.halo(w.rank, phase, 0);
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t < nUnique * width) {
        const int64_t u = t / width;
        const int c = (int)(t - u * width);
        double acc = 0.0;
        for (int64_t k = uPtr[u]; k < uPtr[u + 1]; ++k) {
            const int64_t wpos = uWin[k];
            const int64_t sender = wpos / kHaloSegCap, slot = wpos - sender * kHaloSegCap;
            acc += base[sender * kHaloSegCap + slot * width + c];
        }
        vec[(int64_t)uIdx[u] * width + c] += acc;
    }
    peer_finish(w, SET_HALO, s, ticket);
}

// The multi-rank PCG iteration's first communication step as ONE kernel: the interface sum-exchange of Ap AND the
// all-reduce of the scalar p.Ap.  Blocks [0, nNbr) are the senders (as k_peer_halo_send), block nNbr all-reduces the
// scalar over the window's all-reduce slots (one double per rank, summed in rank order), the remaining blocks are the
// receivers (as k_peer_halo_recv).  Senders never wait on receivers, so residency order cannot deadlock.
__global__ void __launch_bounds__(256)
k_peer_halo_fused(PeerWin w, int nNbr, const int32_t *__restrict__ nbrRank, const int64_t *__restrict__ offsets, int width,
                  const int32_t *__restrict__ idx, int64_t nUnique, const int32_t *__restrict__ uIdx, const int64_t *__restrict__ uPtr,
                  const int64_t *__restrict__ uWin, double *__restrict__ vec, double *scalar, unsigned *ticket) {
    const unsigned long long s = *w.seq(SET_HALO) + 1;
    const int phase = (int)(s & 1);
    if ((int)blockIdx.x < nNbr) {
        const int b = blockIdx.x, q = nbrRank[b];
        const int64_t off = offsets[b], cnt = (offsets[b + 1] - off) * width;
        double *dst = w.halo(q, phase, w.rank);
        for (int64_t t = threadIdx.x; t < cnt; t += blockDim.x) {
            const int64_t k = t / width;
            dst[t] = vec[(int64_t)idx[off + k] * width + (t - k * width)];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            peer_publish(w.flags(q, SET_HALO) + w.rank, s);
            atomicAdd(w.sendDone(), 1ull);        // this block no longer reads vec (the receivers below update it)
        }
    } else if ((int)blockIdx.x == nNbr) {
        const unsigned long long sa = *w.seq(SET_AR) + 1;
        const int pa = (int)(sa & 1);
        __shared__ double part[kPeerMax];
        if ((int)threadIdx.x < w.R) {
            const int q = threadIdx.x;
            w.ar(q, pa, w.rank)[0] = scalar[0];
            peer_publish(w.flags(q, SET_AR) + w.rank, sa);
            if (!peer_wait(w.flags(w.rank, SET_AR) + q, sa)) *w.err() = 1;
            part[q] = w.ar(w.rank, pa, q)[0];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int r = 0; r < w.R; ++r) t += part[r];
            scalar[0] = t;
            atomicAdd(w.arrive(), (unsigned long long)w.R);      // keeps k_peer_allreduce's "R blocks per call" count
            *w.seq(SET_AR) = sa;
        }
    } else {
        // the neighbours' values have landed AND all of MY sender blocks are done reading vec (they share this launch)
        if ((int)threadIdx.x < nNbr && !peer_wait(w.flags(w.rank, SET_HALO) + nbrRank[threadIdx.x], s)) *w.err() = 1;
        if ((int)threadIdx.x == nNbr && !peer_wait(w.sendDone(), *w.fusedCalls() + (unsigned long long)nNbr)) *w.err() = 1;
        __syncthreads();
        const double *base = w.halo(w.rank, phase, 0);
        const int64_t t = (blockIdx.x - nNbr - 1) * (int64_t)blockDim.x + threadIdx.x;
        if (t < nUnique * width) {
            const int64_t u = t / width;
            const int c = (int)(t - u * width);
            double acc = 0.0;
            for (int64_t k = uPtr[u]; k < uPtr[u + 1]; ++k) {
                const int64_t wpos = uWin[k];
                const int64_t sender = wpos / kHaloSegCap, slot = wpos - sender * kHaloSegCap;
                acc += base[sender * kHaloSegCap + slot * width + c];
            }
            vec[(int64_t)uIdx[u] * width + c] += acc;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) { *ticket = 0u; *w.seq(SET_HALO) = s; *w.fusedCalls() = *w.sendDone(); }      // snapshot: all nNbr senders of this call are in
    }
}

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

static PeerWinOwner *peer_of(mfem_b200_ctx *c) { return static_cast<PeerWinOwner *>(c->peerWin); }
static bool peer_active(mfem_b200_ctx *c) { return c->peerWin && peer_of(c)->open && c->opt_comm_p2p; }

void comm_destroy(mfem_b200_ctx *c) {
    delete c->halo;
    c->halo = nullptr;
    if (c->peerWin && c->ownsWin) {
        // nobody may still be spinning on this window: the ranks leave together
        if (c->ncclComm) {
            cudaStreamSynchronize(c->stream);
        }
        delete peer_of(c);
    }
    c->peerWin = nullptr;
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
    if (peer_active(c) && h.p2pFits) {
        // our own kernels over NVLink peer memory: the sender packs straight into the neighbours' windows
        const PeerWin &w = peer_of(c)->w;
        const int nNbr = (int)h.ranks.size();
        k_peer_halo_send<<<nNbr, 1024, 0, s>>>(w, nNbr, h.nbrRankDev, h.offsetsDev, width, h.idx, vec, w.ticket(SET_HALO), nullptr);
        k_peer_halo_recv<<<grid_for(h.nUnique * width, 256), 256, 0, s>>>(w, nNbr, h.nbrRankDev, h.nUnique, width, h.uIdx, h.uPtr, h.uWin, vec,
                                                                           w.ticket(SET_HALO), nullptr);
        c->launches += 2;
        return;
    }
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

// vec += interface partial sums of the neighbours AND *scalar = sum over ranks, in one kernel over the peer window;
// false (nothing done) when the window path is not available for this exchange
bool halo_exchange_add_allreduce1(mfem_b200_ctx *c, double *vec, int width, double *scalar) {
    if (c->nRanks <= 1 || !c->halo || !peer_active(c) || !c->halo->p2pFits || c->halo->total == 0) return false;
    Halo &h = *c->halo;
    const PeerWin &w = peer_of(c)->w;
    const int nNbr = (int)h.ranks.size();
    const int grid = nNbr + 1 + grid_for(h.nUnique * width, 256);
    k_peer_halo_fused<<<grid, 256, 0, c->stream>>>(w, nNbr, h.nbrRankDev, h.offsetsDev, width, h.idx, h.nUnique, h.uIdx, h.uPtr, h.uWin, vec, scalar,
                                                   w.ticket(SET_HALO));
    c->launches++;
    return true;
}
const PeerWin *comm_peer_window(mfem_b200_ctx *c) { return peer_active(c) ? &peer_of(c)->w : nullptr; }

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
    if (peer_active(c) && n <= kArCap) {
        const PeerWin &w = peer_of(c)->w;
        k_peer_allreduce<<<w.R, 1024, 0, c->stream>>>(w, in, out, n, w.ticket(SET_AR), nullptr);
        c->launches++;
        return;
    }
    MFEM_NCCL(ncclAllReduce(in, out, (size_t)n, ncclDouble, ncclSum, static_cast<ncclComm_t>(c->ncclComm), c->stream));
}

int comm_peer_error(mfem_b200_ctx *c) {
    if (!c->peerWin || !peer_of(c)->open) return 0;
    int e = 0;
    cudaMemcpy(&e, peer_of(c)->w.err(), sizeof(int), cudaMemcpyDeviceToHost);
    return e;
}
bool comm_uses_peer_window(mfem_b200_ctx *c) { return peer_active(c); }

// every rank contributes buf[rank*count .. (rank+1)*count) and receives all slices (in place)
void allgather_inplace(mfem_b200_ctx *c, double *buf, int count) {
    if (c->nRanks <= 1) return;
    if (peer_active(c) && (int64_t)count * c->nRanks <= kAgCap) {
        const PeerWin &w = peer_of(c)->w;
        k_peer_allgather<<<w.R, 1024, 0, c->stream>>>(w, buf, count, w.ticket(SET_AG), nullptr);
        c->launches++;
        return;
    }
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
    h->peerWin = parent->peerWin;          // the peer window is process-level state like the communicator
    h->ownsWin = false;
    return MFEM_B200_OK;
}

// Peer window, step 1: allocate this rank's window and return its 64-byte CUDA IPC handle; the caller gathers the handles
// of all ranks (torch.distributed / MPI) and passes them to mfem_b200_comm_window_open.  After comm_init.
int mfem_b200_comm_window_handle(mfem_b200_handle h, void *out64) {
    if (!h || !out64) return MFEM_B200_ERR_INVALID;
    try {
        MFEM_CUDA(cudaSetDevice(h->device));
        MFEM_REQUIRE(h->nRanks > 1 && h->ncclComm && h->ownsComm, MFEM_B200_ERR_INVALID, "comm_window_handle: call comm_init first");
        MFEM_REQUIRE(h->nRanks <= kPeerMax, MFEM_B200_ERR_INVALID, "comm_window: at most 8 ranks (one NVSwitch box)");
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
        if (!h->peerWin) {
            auto *o = new PeerWinOwner();
            h->peerWin = o;
            h->ownsWin = true;
            MFEM_CUDA(cudaMalloc(&o->local, PeerWin::bytes));
            MFEM_CUDA(cudaMemset(o->local, 0, PeerWin::offAr));          // flags, sequence numbers, error word, tickets
        }
        cudaIpcMemHandle_t ih;
        MFEM_CUDA(cudaIpcGetMemHandle(&ih, peer_of(h)->local));
        memcpy(out64, &ih, sizeof(ih));
        return MFEM_B200_OK;
    } catch (const CudaError &e) {
        h->err = e.what();
        return e.status;
    }
}

// step 2: map the windows of all ranks (handles in rank order, 64 bytes each).  The caller must put a barrier between
// this call and the first collective (every rank has to have opened the windows it will write to -- the NCCL
// collectives of set_interface / the first set-up provide it anyway, mfem_b200_comm_window_open ends with an all-reduce).
int mfem_b200_comm_window_open(mfem_b200_handle h, const void *handles) {
    if (!h || !handles) return MFEM_B200_ERR_INVALID;
    try {
        MFEM_CUDA(cudaSetDevice(h->device));
        MFEM_REQUIRE(h->peerWin && h->ownsWin, MFEM_B200_ERR_INVALID, "comm_window_open: call comm_window_handle first");
        PeerWinOwner &o = *peer_of(h);
        o.w.R = h->nRanks; o.w.rank = h->rank;
        bool ok = true;
        std::string why;
        for (int r = 0; r < h->nRanks && ok; ++r) {
            if (r == h->rank) { o.w.peer[r] = static_cast<char *>(o.local); continue; }
            cudaIpcMemHandle_t ih;
            memcpy(&ih, static_cast<const char *>(handles) + 64 * (size_t)r, sizeof(ih));
            const cudaError_t e = cudaIpcOpenMemHandle(&o.mapped[r], ih, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { ok = false; why = cudaGetErrorString(e); cudaGetLastError(); break; }
            o.w.peer[r] = static_cast<char *>(o.mapped[r]);
        }
        // all ranks agree on whether the window is usable (one refusal disables it everywhere) -- and this all-reduce is
        // the barrier that orders "everybody has mapped everything" before the first peer store
        DevBuf<double> flag(1);
        const double mine = ok ? 0.0 : 1.0;
        MFEM_CUDA(cudaMemcpyAsync(flag, &mine, sizeof(double), cudaMemcpyHostToDevice, h->stream));
        MFEM_NCCL(ncclAllReduce(flag.p, flag.p, 1, ncclDouble, ncclSum, static_cast<ncclComm_t>(h->ncclComm), h->stream));
        double bad = 0.0;
        MFEM_CUDA(cudaMemcpyAsync(&bad, flag, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        MFEM_CUDA(cudaStreamSynchronize(h->stream));
        o.open = bad == 0.0;
        if (!o.open) { h->err = "peer window unavailable (" + (why.empty() ? std::string("refused on another rank") : why) + "): NCCL path"; return MFEM_B200_OK; }
        return MFEM_B200_OK;
    } catch (const CudaError &e) {
        h->err = e.what();
        return e.status;
    }
}

// 1 if the small collectives of this handle run over the peer window, 0 if over NCCL
int mfem_b200_comm_uses_peer_window(mfem_b200_handle h) { return (h && comm_uses_peer_window(h)) ? 1 : 0; }

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
            // peer-window path: neighbour table on the device and, per receive slot, (sender rank, slot in its segment)
            std::vector<int32_t> nbr(H.ranks.begin(), H.ranks.end());
            std::vector<int64_t> uWin((size_t)H.total);
            H.p2pFits = true;
            for (int q = 0; q < n_neighbors; ++q)
                if ((H.offsets[q + 1] - H.offsets[q]) * H.maxWidth > kHaloSegCap) H.p2pFits = false;
            for (int64_t k = 0; k < H.total; ++k) {
                const int64_t pos = order[(size_t)k];
                const int q = (int)(std::upper_bound(H.offsets.begin(), H.offsets.end(), pos) - H.offsets.begin()) - 1;
                uWin[(size_t)k] = (int64_t)H.ranks[q] * kHaloSegCap + (pos - H.offsets[q]);
            }
            H.nbrRankDev.alloc(nbr.size()); H.offsetsDev.alloc(H.offsets.size()); H.uWin.alloc(uWin.size());
            MFEM_CUDA(cudaMemcpyAsync(H.nbrRankDev, nbr.data(), H.nbrRankDev.bytes(), cudaMemcpyHostToDevice, s));
            MFEM_CUDA(cudaMemcpyAsync(H.offsetsDev, H.offsets.data(), H.offsetsDev.bytes(), cudaMemcpyHostToDevice, s));
            MFEM_CUDA(cudaMemcpyAsync(H.uWin, uWin.data(), H.uWin.bytes(), cudaMemcpyHostToDevice, s));
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
