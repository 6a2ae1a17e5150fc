// Peer window (csrc/comm.cu): layout of the per-rank device buffer every other rank of the box maps through CUDA IPC, and
// the device-side primitives of the collectives built on it.  Internal header (comm.cu owns the windows; solver.cu's
// dense-level kernels store into them directly).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace mfem {

constexpr int kPeerMax = 8;                       // ranks of one NVSwitch box
constexpr int64_t kArCap = 32768 + 8;             // doubles per rank slot of the all-reduce region (2 + coarse residuals)
constexpr int64_t kAgCap = 32768;                 // doubles of the all-gather region
constexpr int64_t kHaloSegCap = (int64_t)1 << 19; // doubles per SENDER slot of the halo region (4 MiB: 29k DoFs x 18 values)
enum { SET_AR = 0, SET_AG = 1, SET_HALO = 2, SET_COUNT = 4 };

struct PeerWin {
    int R = 0, rank = 0;
    char *peer[kPeerMax] = {};                    // peer[r] = rank r's window as mapped here (peer[rank] = my own)
    // byte offsets inside a window
    static constexpr size_t offFlags = 0;                                        // [SET_COUNT][kPeerMax] uint64
    static constexpr size_t offSeq = offFlags + SET_COUNT * kPeerMax * 8;         // [SET_COUNT] uint64 (local use only)
    static constexpr size_t offErr = offSeq + SET_COUNT * 8;                      // int
    static constexpr size_t offTicket = offErr + 8;                               // [SET_COUNT] unsigned (local use only)
    static constexpr size_t offArrive = offTicket + SET_COUNT * 4;                // uint64: blocks of MY all-reduce kernels that shipped their copy
    static constexpr size_t offSendDone = offArrive + 8;                          // uint64: sender blocks of MY fused halo kernels that finished packing
    static constexpr size_t offFusedCalls = offSendDone + 8;                      // uint64: value of sendDone when the previous fused halo kernel finished
    static constexpr size_t offAr = 4096;                                        // [2][kPeerMax][kArCap] double
    static constexpr size_t offAg = offAr + 2 * kPeerMax * kArCap * 8;            // [2][kAgCap] double
    static constexpr size_t offHalo = offAg + 2 * kAgCap * 8;                     // [2][kPeerMax][kHaloSegCap] double
    static constexpr size_t bytes = offHalo + 2 * kPeerMax * kHaloSegCap * 8;
    __host__ __device__ unsigned long long *flags(int r, int set) const { return reinterpret_cast<unsigned long long *>(peer[r] + offFlags) + set * kPeerMax; }
    __host__ __device__ unsigned long long *seq(int set) const { return reinterpret_cast<unsigned long long *>(peer[rank] + offSeq) + set; }
    __host__ __device__ int *err() const { return reinterpret_cast<int *>(peer[rank] + offErr); }
    __host__ __device__ unsigned *ticket(int set) const { return reinterpret_cast<unsigned *>(peer[rank] + offTicket) + set; }
    __host__ __device__ unsigned long long *arrive() const { return reinterpret_cast<unsigned long long *>(peer[rank] + offArrive); }
    __host__ __device__ unsigned long long *sendDone() const { return reinterpret_cast<unsigned long long *>(peer[rank] + offSendDone); }
    __host__ __device__ unsigned long long *fusedCalls() const { return reinterpret_cast<unsigned long long *>(peer[rank] + offFusedCalls); }
    __host__ __device__ double *ar(int r, int phase, int slot) const { return reinterpret_cast<double *>(peer[r] + offAr) + ((size_t)phase * kPeerMax + slot) * kArCap; }
    __host__ __device__ double *ag(int r, int phase) const { return reinterpret_cast<double *>(peer[r] + offAg) + (size_t)phase * kAgCap; }
    __host__ __device__ double *halo(int r, int phase, int sender) const { return reinterpret_cast<double *>(peer[r] + offHalo) + ((size_t)phase * kPeerMax + sender) * kHaloSegCap; }
};

__device__ __forceinline__ void peer_publish(unsigned long long *flag, unsigned long long s) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long *>(flag) = s;
}
// spin until *flag >= s (flag lives in LOCAL memory, written by a peer); false on timeout
__device__ __forceinline__ bool peer_wait(const unsigned long long *flag, unsigned long long s) {
    const volatile unsigned long long *f = reinterpret_cast<const volatile unsigned long long *>(flag);
    const long long t0 = clock64();
    while (*f < s) {
        if (clock64() - t0 > 4000000000ll) return false;
        __nanosleep(20);
    }
    __threadfence_system();
    return true;
}
// the last block of a collective kernel advances the local sequence number of its set
__device__ __forceinline__ void peer_finish(const PeerWin &w, int set, unsigned long long s, unsigned *ticket) {
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
        if (last) { *ticket = 0u; *w.seq(set) = s; }
    }
}


}  // namespace mfem
