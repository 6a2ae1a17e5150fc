// Mesh upload, DoF renumbering, element geometry, and the symbolic phase
// (block-CSR pattern + DoF->element incidence lists) -- all on the device.
//
// Reference being replaced:
//   FEMMesh ctor / Simulator ctor            FEMMesh.inl:11-82, LinearElasticity.hh:460-473
//   LinearlyEmbeddedSimplex::embed           EmbeddedElement.hh:170-190, 211-231
//   triplet reservation + sumRepeated's      LinearElasticity.hh:1441-1443,
//   bucket/sort/unique (symbolic part)       SparseMatrices.hh:280-374
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <limits>

#include "core.cuh"

namespace mfem {

// ---------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------
__global__ void k_check_elem_nodes(int64_t n, const int32_t *en, int64_t nNodes, int *bad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n && (en[i] < 0 || en[i] >= nNodes)) atomicAdd(bad, 1);
}

__global__ void k_rep_node(int64_t nNodes, const int32_t *dofOfNode, int32_t *repNode) {
    int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n < nNodes) atomicMin(&repNode[dofOfNode[n]], (int32_t)n);
}

__device__ __forceinline__ uint64_t spread3(uint64_t v) {   // 21 bits -> every third bit
    v &= 0x1fffffULL;
    v = (v | (v << 32)) & 0x1f00000000ffffULL;
    v = (v | (v << 16)) & 0x1f0000ff0000ffULL;
    v = (v | (v << 8)) & 0x100f00f00f00f00fULL;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ULL;
    v = (v | (v << 2)) & 0x1249249249249249ULL;
    return v;
}
__device__ __forceinline__ uint64_t spread2(uint64_t v) {   // 31 bits -> every second bit
    v &= 0x7fffffffULL;
    v = (v | (v << 16)) & 0x0000ffff0000ffffULL;
    v = (v | (v << 8)) & 0x00ff00ff00ff00ffULL;
    v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0fULL;
    v = (v | (v << 2)) & 0x3333333333333333ULL;
    v = (v | (v << 1)) & 0x5555555555555555ULL;
    return v;
}

// Morton key of every DoF from the position of its representative node.  One common
// scale for all axes keeps the curve's cells cubic on elongated domains.
__global__ void k_morton_keys(int N, int64_t nDofs, const int32_t *repNode, const double *nodes, double mn0,
                              double mn1, double mn2, double invScale, uint64_t *keys, int32_t *ids) {
    int64_t d = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (d >= nDofs) return;
    const int32_t n = repNode[d];
    uint64_t key;
    if (n == 0x7f7f7f7f) {   // memset marker: no node maps to this DoF
        key = ~0ULL;                 // DoF without node (should not happen): park at the end
    } else if (N == 3) {
        const double s = 2097151.0;  // 2^21 - 1
        uint64_t x = (uint64_t)(fmin(fmax((nodes[3 * (int64_t)n + 0] - mn0) * invScale, 0.0), 1.0) * s);
        uint64_t y = (uint64_t)(fmin(fmax((nodes[3 * (int64_t)n + 1] - mn1) * invScale, 0.0), 1.0) * s);
        uint64_t z = (uint64_t)(fmin(fmax((nodes[3 * (int64_t)n + 2] - mn2) * invScale, 0.0), 1.0) * s);
        key = spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
    } else {
        const double s = 2147483647.0;   // 2^31 - 1
        uint64_t x = (uint64_t)(fmin(fmax((nodes[2 * (int64_t)n + 0] - mn0) * invScale, 0.0), 1.0) * s);
        uint64_t y = (uint64_t)(fmin(fmax((nodes[2 * (int64_t)n + 1] - mn1) * invScale, 0.0), 1.0) * s);
        key = spread2(x) | (spread2(y) << 1);
    }
    keys[d] = key;
    ids[d] = (int32_t)d;
}

__global__ void k_iota(int64_t n, int32_t *a) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) a[i] = (int32_t)i;
}

__global__ void k_invert_perm(int64_t n, const int32_t *int2ext, int32_t *ext2int) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) ext2int[int2ext[i]] = (int32_t)i;
}

__global__ void k_node_dof(int64_t nNodes, const int32_t *dofOfNode, const int32_t *ext2int, int32_t *nodeDof) {
    int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n < nNodes) nodeDof[n] = ext2int[dofOfNode[n]];
}

__global__ void k_elem_dof(int64_t n, const int32_t *elemNodes, const int32_t *nodeDof, int32_t *elemDof) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) elemDof[i] = nodeDof[elemNodes[i]];
}

// K1 elem_geom: one thread per element (EmbeddedElement.hh:170-190, 211-231).
template <int N>
__global__ void k_elem_geom(int64_t nElems, int npe, const int32_t *elemNodes, const double *nodes, double *geom,
                            int *negCount) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nElems) return;
    double p[N + 1][N];
#pragma unroll
    for (int v = 0; v <= N; ++v) {
        const int64_t n = elemNodes[e * npe + v];
#pragma unroll
        for (int r = 0; r < N; ++r) p[v][r] = nodes[n * N + r];
    }
    ElemGeom<N> g;
    embed(p, g);
    constexpr int GS = 1 + N * (N + 1);
    double *o = geom + e * GS;
    o[0] = g.vol;
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int a = 0; a <= N; ++a) o[1 + r * (N + 1) + a] = g.G[r][a];
    if (!(g.vol >= 0.0)) atomicAdd(negCount, 1);
}

// (row,col) block keys of every element: npe*npe per element.
__global__ void k_pair_keys(int64_t nElems, int npe, const int32_t *elemDof, uint64_t *keys) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t per = (int64_t)npe * npe;
    if (t >= nElems * per) return;
    const int64_t e = t / per;
    const int ij = (int)(t - e * per);
    const int i = ij / npe, j = ij - i * npe;
    keys[t] = ((uint64_t)(uint32_t)elemDof[e * npe + i] << 32) | (uint32_t)elemDof[e * npe + j];
}

__global__ void k_low32(int64_t n, const uint64_t *keys, int32_t *out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int32_t)(keys[i] & 0xffffffffULL);
}

// rowptr[r] = first index k with (keys[k] >> 32) >= r   (keys sorted)
__global__ void k_rowptr_from_keys64(int64_t nRows, int64_t nKeys, const uint64_t *keys, int64_t *rowptr) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r > nRows) return;
    int64_t lo = 0, hi = nKeys;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)(keys[mid] >> 32) < r) lo = mid + 1; else hi = mid;
    }
    rowptr[r] = lo;
}

__global__ void k_rowptr_from_keys32(int64_t nRows, int64_t nKeys, const uint32_t *keys, int64_t *rowptr) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r > nRows) return;
    int64_t lo = 0, hi = nKeys;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)keys[mid] < r) lo = mid + 1; else hi = mid;
    }
    rowptr[r] = lo;
}

// jobRow[k] = first row r with incPtr[r] >= k * chunk  (k = 0..nJobs; jobRow[nJobs] = nb + 1 sentinel side)
__global__ void k_job_rows(int64_t nJobs, int64_t nb, int chunk, const int64_t *incPtr, int64_t *jobRow) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k > nJobs) return;
    const int64_t target = k * chunk;
    int64_t lo = 0, hi = nb + 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (incPtr[mid] < target) lo = mid + 1; else hi = mid;
    }
    jobRow[k] = lo;
}

// tileRow[t] = first row r with rowptr[r] >= t * window (t < nTiles); tileRow[nTiles] = nb
__global__ void k_tile_rows(int64_t nTiles, int64_t nb, int window, const int64_t *rowptr, int64_t *tileRow) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t > nTiles) return;
    if (t == nTiles) { tileRow[t] = nb; return; }
    const int64_t target = t * window;
    int64_t lo = 0, hi = nb;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (rowptr[mid] < target) lo = mid + 1; else hi = mid;
    }
    tileRow[t] = lo;
}
__global__ void k_max_row_len(int64_t nb, const int64_t *rowptr, unsigned long long *out) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    unsigned long long len = r < nb ? (unsigned long long)(rowptr[r + 1] - rowptr[r]) : 0ULL;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, len);      // one atomic per warp instead of per row
}

// ---- block-owner assembly plan (assemble.cu k_assemble_blocks) -----------------------------
// Blocks are processed in chunks of kBlkChunk consecutive BSR blocks, one CTA per chunk.  The
// contribution list of every block is cut into SEGMENTS of at most L pair ids (L = 4 unless a
// chunk would then need more than kSegSlots segments), so that a vertex-row diagonal block with
// 48 contributions costs twelve threads four loop trips each instead of one thread 48 trips.
// The segments of a chunk are handed to the thread slots in order of decreasing length (the 32
// lanes of a warp loop the same number of times); the pair ids of a warp-round are stored
// interleaved (entry `it` of lane l at warpBase + 32*it + l, padded with kPlanSentinel) so that
// every loop trip is one coalesced 128-byte load.
__device__ __forceinline__ int plan_segments_of(int cnt, int L) { return (cnt + L - 1) / L; }

__global__ void __launch_bounds__(kBlkChunk)
k_plan_segments(int64_t nnzb, const int32_t *__restrict__ counts, uint16_t *__restrict__ planCnt,
                uint16_t *__restrict__ planSegOff, uint8_t *__restrict__ planNseg,
                uint8_t *__restrict__ chunkL, uint16_t *__restrict__ segOrder, int64_t *__restrict__ warpEntries,
                int *__restrict__ overflow, const int32_t *__restrict__ blockStart, const uint32_t *__restrict__ sortedPairs,
                int pairsPerElem) {
    typedef cub::BlockScan<int, kBlkChunk> Scan;
    typedef cub::BlockReduce<int, kBlkChunk> Reduce;
    typedef cub::BlockRadixSort<uint32_t, kBlkChunk, 2, uint32_t> Sort;
    __shared__ union { typename Scan::TempStorage scan; typename Reduce::TempStorage red; typename Sort::TempStorage sort; } temp;
    __shared__ int sTotal;
    __shared__ uint16_t sLen[kSegSlots], sLenSorted[kSegSlots], sIdxSorted[kSegSlots];
    __shared__ uint32_t sElem[kSegSlots];     // element of the first contribution of every segment
    const int t = threadIdx.x;
    const int64_t chunk = blockIdx.x, k = chunk * kBlkChunk + t;
    int cnt = k < nnzb ? counts[k] : 0;
    if (cnt > 65535) { atomicExch(overflow, 1); cnt = 65535; }
    int L = 4;
    while (true) {
        // (a block with more than 255 segments counts as "too many" so that nseg fits a byte)
        const int mine = plan_segments_of(cnt, L);
        const int tot = Reduce(temp.red).Sum(mine > 255 ? kSegSlots + 1 : mine);
        if (t == 0) sTotal = tot;
        __syncthreads();
        const bool ok = sTotal <= kSegSlots;
        __syncthreads();
        if (ok) break;
        L *= 2;
    }
    const int nseg = plan_segments_of(cnt, L);
    int off;
    Scan(temp.scan).ExclusiveSum(nseg, off);
    for (int j = t; j < kSegSlots; j += kBlkChunk) sLen[j] = 0;
    __syncthreads();
    for (int p = 0; p < nseg; ++p) {
        sLen[off + p] = (uint16_t)min(L, cnt - p * L);
        sElem[off + p] = sortedPairs[(int64_t)blockStart[k] + (int64_t)p * L] / (uint32_t)pairsPerElem;
    }
    __syncthreads();
    // Sort key: length first (the lanes of a warp loop equally often), then the element of the first
    // contribution: segments that read the same element records end up in the same warp, so a geometry
    // gather touches few distinct cache lines.
    uint32_t key[2];
    static_assert(kSegSlots <= 2 * kBlkChunk, "two sort items per thread");
    for (int i = 0; i < 2; ++i) {
        const int sgi = 2 * t + i;
        key[i] = (sgi < kSegSlots && sLen[sgi]) ? (((uint32_t)min((int)sLen[sgi], 255) << 24) | (sElem[sgi] & 0xffffffu)) : 0u;
    }
    uint32_t val[2] = {(uint32_t)(2 * t), (uint32_t)(2 * t + 1)};
    Sort(temp.sort).SortDescending(key, val);        // stable; blocked: thread t holds ranks 2t, 2t+1
    for (int i = 0; i < 2; ++i)
        if (2 * t + i < kSegSlots) {          // ranks beyond kSegSlots hold only empty items
            sLenSorted[2 * t + i] = key[i] ? sLen[val[i]] : (uint16_t)0;
            sIdxSorted[2 * t + i] = (uint16_t)val[i];
        }
    __syncthreads();
    for (int j = t; j < kSegSlots; j += kBlkChunk)
        segOrder[chunk * kSegSlots + j] = sLenSorted[j] ? sIdxSorted[j] : (uint16_t)0xffff;
    if (t < kSegSlots / 32) {          // trips of a warp-round = its longest segment
        int mx = 0;
        for (int j = 0; j < 32; ++j) mx = max(mx, (int)sLenSorted[32 * t + j]);
        warpEntries[chunk * (kSegSlots / 32) + t] = 32 * (int64_t)mx;
    }
    if (k < nnzb) { planCnt[k] = (uint16_t)cnt; planSegOff[k] = (uint16_t)off; planNseg[k] = (uint8_t)nseg; }
    if (t == 0) chunkL[chunk] = (uint8_t)L;
}

// ---- distinct elements of a chunk (TMA staging of the geometry records, assemble.cu) ----------
// The contributions of a chunk are one contiguous range of the sorted pair list.  Their elements
// (a few dozen: the tets around ~9 DoF rows) go into a shared-memory hash set; slot order gives the
// chunk-local element numbering.  Chunks with more than kGeomCap distinct elements are left unstaged.
constexpr int kElemHash = 1024;                  // hash slots (power of two), load factor <= 1/2 enforced
constexpr uint32_t kHashEmpty = 0xffffffffu;

__device__ __forceinline__ int elem_hash_slot(uint32_t e) { return (int)((e * 2654435761u) >> 22); }

// inserts the elements of pairs [a, b) -- returns false (for every thread) if the table overflowed
__device__ __forceinline__ bool chunk_hash_build(uint32_t *sKeys, int *sFlag, int64_t a, int64_t b,
                                                 const uint32_t *__restrict__ sortedPairs, uint32_t pairsPerElem) {
    for (int j = threadIdx.x; j < kElemHash; j += blockDim.x) sKeys[j] = kHashEmpty;
    if (threadIdx.x == 0) *sFlag = 0;
    __syncthreads();
    for (int64_t i = a + threadIdx.x; i < b; i += blockDim.x) {
        const uint32_t e = sortedPairs[i] / pairsPerElem;
        int slot = elem_hash_slot(e);
        for (int probe = 0; probe < kElemHash; ++probe) {
            const uint32_t old = atomicCAS(&sKeys[slot], kHashEmpty, e);
            if (old == kHashEmpty || old == e) break;
            slot = (slot + 1) & (kElemHash - 1);
            if (probe > kElemHash / 2) { *sFlag = 1; break; }
        }
    }
    __syncthreads();
    return *sFlag == 0;
}

__device__ __forceinline__ int chunk_hash_find(const uint32_t *sKeys, uint32_t e) {
    int slot = elem_hash_slot(e);
    while (sKeys[slot] != e) slot = (slot + 1) & (kElemHash - 1);
    return slot;
}

// number of elements to stage per chunk (0 = chunk stays on the global-load path)
__global__ void __launch_bounds__(kBlkChunk)
k_plan_count_elems(int64_t nnzb, int64_t nPairs, const int32_t *__restrict__ blockStart,
                   const uint32_t *__restrict__ sortedPairs, int pairsPerElem, int64_t *__restrict__ chunkNElems) {
    typedef cub::BlockReduce<int, kBlkChunk> Reduce;
    __shared__ typename Reduce::TempStorage temp;
    __shared__ uint32_t sKeys[kElemHash];
    __shared__ int sFlag;
    const int64_t k0 = (int64_t)blockIdx.x * kBlkChunk;
    const int64_t a = blockStart[k0], b = (k0 + kBlkChunk < nnzb) ? blockStart[k0 + kBlkChunk] : nPairs;
    const bool ok = chunk_hash_build(sKeys, &sFlag, a, b, sortedPairs, (uint32_t)pairsPerElem);
    int mine = 0;
    for (int j = threadIdx.x; j < kElemHash; j += blockDim.x) mine += sKeys[j] != kHashEmpty;
    const int total = Reduce(temp).Sum(mine);
    if (threadIdx.x == 0) chunkNElems[blockIdx.x] = (ok && total <= kGeomCap) ? total : 0;
}

__global__ void __launch_bounds__(kBlkChunk)
k_plan_fill(int64_t nnzb, const uint16_t *__restrict__ planCnt, const uint8_t *__restrict__ chunkL,
            const int32_t *__restrict__ blockStart, const uint32_t *__restrict__ sortedPairs,
            const uint16_t *__restrict__ segOrder, const int64_t *__restrict__ warpBase, uint32_t *__restrict__ list,
            int64_t nPairs, int pairsPerElem, const int64_t *__restrict__ chunkElemPtr, uint32_t *__restrict__ chunkElems) {
    typedef cub::BlockScan<int, kBlkChunk> Scan;
    __shared__ typename Scan::TempStorage temp;
    __shared__ int sOff[kBlkChunk + 1], sCnt[kBlkChunk];
    __shared__ uint32_t sKeys[kElemHash];
    __shared__ uint16_t sLocal[kElemHash];
    __shared__ int sFlag;
    const int t = threadIdx.x;
    const int64_t chunk = blockIdx.x, k = chunk * kBlkChunk + t;
    // staged chunk: chunk-local element numbering = order of the occupied hash slots
    const int64_t ePtr = chunkElemPtr[chunk];
    const bool staged = chunkElemPtr[chunk + 1] > ePtr;
    if (staged) {
        const int64_t k0 = chunk * kBlkChunk;
        const int64_t a = blockStart[k0], b = (k0 + kBlkChunk < nnzb) ? blockStart[k0 + kBlkChunk] : nPairs;
        chunk_hash_build(sKeys, &sFlag, a, b, sortedPairs, (uint32_t)pairsPerElem);
        constexpr int PER = kElemHash / kBlkChunk;
        int occ = 0;
        for (int q = 0; q < PER; ++q) occ += sKeys[PER * t + q] != kHashEmpty;
        int first;
        Scan(temp).ExclusiveSum(occ, first);
        for (int q = 0; q < PER; ++q)
            if (sKeys[PER * t + q] != kHashEmpty) {
                sLocal[PER * t + q] = (uint16_t)first;
                chunkElems[ePtr + first] = sKeys[PER * t + q];
                ++first;
            }
        __syncthreads();
    }
    const int cnt = k < nnzb ? planCnt[k] : 0;
    const int L = chunkL[chunk];
    int off, total;
    Scan(temp).ExclusiveSum(plan_segments_of(cnt, L), off, total);
    sOff[t] = off; sCnt[t] = cnt;
    if (t == 0) sOff[kBlkChunk] = total;
    __syncthreads();
    for (int j = t; j < kSegSlots; j += kBlkChunk) {
        const int64_t wr = chunk * (kSegSlots / 32) + (j >> 5);
        const int64_t base = warpBase[wr];
        const int nIt = (int)((warpBase[wr + 1] - base) >> 5);
        const int sg = segOrder[chunk * kSegSlots + j];
        int len = 0;
        int64_t start = 0;
        if (sg != 0xffff) {
            int lo = 0, hi = kBlkChunk - 1;           // last block kb with sOff[kb] <= sg and a segment of its own
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (sOff[mid] <= sg) lo = mid; else hi = mid - 1;
            }
            const int part = sg - sOff[lo];
            len = min(L, sCnt[lo] - part * L);
            start = (int64_t)blockStart[chunk * kBlkChunk + lo] + (int64_t)part * L;
        }
        for (int it = 0; it < nIt; ++it) {
            uint32_t v = kPlanSentinel;
            if (it < len) {
                v = sortedPairs[start + it];
                if (staged) {       // (element, i, j) -> (chunk-local element, i, j)
                    const uint32_t e = v / (uint32_t)pairsPerElem;
                    v = (uint32_t)sLocal[chunk_hash_find(sKeys, e)] * (uint32_t)pairsPerElem + (v - e * (uint32_t)pairsPerElem);
                }
            }
            list[base + 32 * (int64_t)it + (j & 31)] = v;
        }
    }
}

// chunkRow[c] = block row containing block min(c * kBlkChunk, nnzb - 1), c = 0..nChunks
__global__ void k_chunk_rows(int64_t nChunks, int64_t nb, int64_t nnzb, const int64_t *rowptr, int32_t *chunkRow) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c > nChunks) return;
    int64_t k = c * kBlkChunk;
    if (k > nnzb - 1) k = nnzb - 1;
    int64_t lo = 0, hi = nb - 1;                     // last r with rowptr[r] <= k
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (rowptr[mid] <= k) lo = mid; else hi = mid - 1;
    }
    chunkRow[c] = (int32_t)lo;
}

__global__ void k_iota_u32(int64_t n, uint32_t *a) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) a[i] = (uint32_t)i;
}

static int bits_for(int64_t n) {
    int b = 1;
    while ((int64_t(1) << b) < n) ++b;
    return b;
}

// upperStart[r] = first slot of block row r whose column is >= r (the diagonal when it is present)
__global__ void k_upper_start(int64_t nb, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                              int32_t *__restrict__ upperStart) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= nb) return;
    const int64_t b0 = rowptr[r];
    int lo = 0, hi = (int)(rowptr[r + 1] - b0);
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (colidx[b0 + mid] < r) lo = mid + 1; else hi = mid;
    }
    upperStart[r] = lo;
}

// Pattern by-products every SpMV variant needs: tile table of the TMA-ring kernel, the longest row, and
// the start of every row's upper-triangular tail (symmetric SpMV).
void finish_pattern(mfem_b200_ctx *c) {
    cudaStream_t s = c->stream;
    const int64_t nb = c->nDofs;
    c->upperStart.alloc((size_t)nb + 2);
    k_upper_start<<<grid_for(nb, 256), 256, 0, s>>>(nb, c->rowptr, c->colidx, c->upperStart);
    c->launches++;
    const int64_t nTiles = (c->nnzb + kSpmvTileWindow - 1) / kSpmvTileWindow;
    c->tileRow.alloc((size_t)nTiles + 1);
    k_tile_rows<<<grid_for(nTiles + 1, 256), 256, 0, s>>>(nTiles, nb, kSpmvTileWindow, c->rowptr, c->tileRow);
    DevBuf<unsigned long long> mx(1);
    MFEM_CUDA(cudaMemsetAsync(mx, 0, 8, s));
    k_max_row_len<<<grid_for(nb, 256), 256, 0, s>>>(nb, c->rowptr, mx);
    c->launches += 2;
    unsigned long long hmx = 0;
    MFEM_CUDA(cudaMemcpyAsync(&hmx, mx, 8, cudaMemcpyDeviceToHost, s));
    MFEM_CUDA(cudaStreamSynchronize(s));
    c->maxRowLen = (int64_t)hmx;
}

// External matrix (mfem_b200_set_matrix_triplets): block-major values [k][r][c] -> row-plane layout
template <int N>
__global__ void k_blocks_to_planes(int64_t nb, const int64_t *__restrict__ rowptr, const double *__restrict__ blk,
                                   double *__restrict__ vals) {
    const int64_t row = blockIdx.x;
    if (row >= nb) return;
    const int64_t b0 = rowptr[row], n = rowptr[row + 1] - b0;
    for (int64_t t = threadIdx.x; t < n * N * N; t += blockDim.x) {
        const int64_t j = t / (N * N);
        const int rc = (int)(t - j * N * N);
        vals[val_index<N>(b0, n, j, rc / N, rc % N)] = blk[(b0 + j) * N * N + rc];
    }
}

void upload_external_bsr(mfem_b200_ctx *c, int dim, int64_t nb, const std::vector<int64_t> &rowptr,
                         const std::vector<int32_t> &colidx, const std::vector<double> &blocks) {
    cudaStream_t s = c->stream;
    c->N = dim; c->deg = 0; c->npe = 0;
    c->nNodes = nb; c->nElems = 0; c->nDofs = nb;
    c->periodic = false;
    c->externalMatrix = true;
    c->geomValid = c->haveMaterial = false;
    c->precondValid = c->workValid = false;
    c->meshVersion++;
    c->fixedHost.assign((size_t)nb * dim, 0);
    c->nFixed = 0;
    // a re-set system starts without constraints on the device too (ensure_work re-zeroes only on a size change)
    c->fixedMask.free(); c->fixedVals.free();
    c->nnzb = (int64_t)colidx.size();
    c->rowptr.alloc((size_t)nb + 1 + 2);
    c->colidx.alloc((size_t)c->nnzb + 4);
    c->vals.alloc((size_t)c->nnzb * dim * dim + 2);
    DevBuf<double> blk(blocks.size());
    MFEM_CUDA(cudaMemcpyAsync(c->rowptr, rowptr.data(), (nb + 1) * 8, cudaMemcpyHostToDevice, s));
    MFEM_CUDA(cudaMemcpyAsync(c->colidx, colidx.data(), colidx.size() * 4, cudaMemcpyHostToDevice, s));
    MFEM_CUDA(cudaMemcpyAsync(blk, blocks.data(), blocks.size() * 8, cudaMemcpyHostToDevice, s));
    if (dim == 3) k_blocks_to_planes<3><<<(unsigned)nb, 128, 0, s>>>(nb, c->rowptr, blk, c->vals);
    else k_blocks_to_planes<2><<<(unsigned)nb, 128, 0, s>>>(nb, c->rowptr, blk, c->vals);
    // identity numbering: the caller's variable order is kept (no coordinates to order along)
    c->int2ext.alloc((size_t)nb);
    c->ext2int.alloc((size_t)nb);
    k_iota<<<grid_for(nb, 256), 256, 0, s>>>(nb, c->int2ext);
    k_iota<<<grid_for(nb, 256), 256, 0, s>>>(nb, c->ext2int);
    c->launches += 3;
    MFEM_CUDA(cudaStreamSynchronize(s));
    MFEM_CUDA(cudaGetLastError());
    finish_pattern(c);
    c->patternValid = c->valuesValid = true;
}

// ---------------------------------------------------------------------------
void setup_mesh(mfem_b200_ctx *c, int dim, int degree, int64_t nNodes, const double *nodes, int64_t nElems,
                const int32_t *elemNodes, const int64_t *dofForNode, int64_t nDofs) {
    MFEM_REQUIRE(dim == 2 || dim == 3, MFEM_B200_ERR_INVALID, "dim must be 2 or 3");
    MFEM_REQUIRE(degree == 1 || degree == 2, MFEM_B200_ERR_INVALID, "degree must be 1 or 2");
    MFEM_REQUIRE(nNodes > 0 && nElems > 0 && nodes && elemNodes, MFEM_B200_ERR_INVALID, "empty mesh");
    MFEM_REQUIRE(nNodes < INT32_MAX && nElems * 16 < INT32_MAX, MFEM_B200_ERR_INVALID,
                 "mesh too large for 32-bit node/element ids");
    cudaStream_t s = c->stream;
    c->N = dim; c->deg = degree; c->npe = nodes_per_elem(dim, degree);
    c->nNodes = nNodes; c->nElems = nElems;
    c->periodic = dofForNode != nullptr;
    c->nDofs = c->periodic ? nDofs : nNodes;
    MFEM_REQUIRE(c->nDofs > 0 && c->nDofs <= nNodes, MFEM_B200_ERR_INVALID, "bad n_dofs");
    c->patternValid = c->valuesValid = c->geomValid = c->precondValid = c->workValid = false;
    c->mfPlanValid = c->mfGeomValid = c->mfChunksValid = false;
    c->meshVersion++;
    c->externalMatrix = false;
    c->fixedHost.assign((size_t)c->nDofs * dim, 0);
    c->nFixed = 0;
    // a re-set mesh starts without constraints on the device too (ensure_work re-zeroes only on a size change)
    c->fixedMask.free(); c->fixedVals.free();

    const int npe = c->npe;
    c->nodes.alloc((size_t)nNodes * dim);
    c->elemNodes.alloc((size_t)nElems * npe);
    MFEM_CUDA(cudaMemcpyAsync(c->nodes, nodes, c->nodes.bytes(), cudaMemcpyHostToDevice, s));
    MFEM_CUDA(cudaMemcpyAsync(c->elemNodes, elemNodes, c->elemNodes.bytes(), cudaMemcpyHostToDevice, s));

    DevBuf<int> flag(1);
    MFEM_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), s));
    k_check_elem_nodes<<<grid_for(nElems * npe, 256), 256, 0, s>>>(nElems * npe, c->elemNodes, nNodes, flag);
    c->launches++;
    int bad = 0;
    MFEM_CUDA(cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    MFEM_CUDA(cudaStreamSynchronize(s));
    MFEM_REQUIRE(bad == 0, MFEM_B200_ERR_INVALID, "Bad vertex index encountered.");

    // DoF of every node in the caller's numbering
    DevBuf<int32_t> dofOfNode((size_t)nNodes);
    if (c->periodic) {
        std::vector<int32_t> tmp((size_t)nNodes);
        for (int64_t i = 0; i < nNodes; ++i) {
            MFEM_REQUIRE(dofForNode[i] >= 0 && dofForNode[i] < c->nDofs, MFEM_B200_ERR_INVALID,
                         "dof_for_node out of range");
            tmp[(size_t)i] = (int32_t)dofForNode[i];
        }
        MFEM_CUDA(cudaMemcpyAsync(dofOfNode, tmp.data(), dofOfNode.bytes(), cudaMemcpyHostToDevice, s));
        MFEM_CUDA(cudaStreamSynchronize(s));
    } else {
        k_iota<<<grid_for(nNodes, 256), 256, 0, s>>>(nNodes, dofOfNode);
        c->launches++;
    }

    c->int2ext.alloc((size_t)c->nDofs);
    c->ext2int.alloc((size_t)c->nDofs);
    if (c->opt_reorder) {
        // bounding box on the host (one pass over data the caller already holds)
        double mn[3] = {std::numeric_limits<double>::max(), std::numeric_limits<double>::max(),
                        std::numeric_limits<double>::max()};
        double mx[3] = {-mn[0], -mn[0], -mn[0]};
        for (int64_t i = 0; i < nNodes; ++i)
            for (int r = 0; r < dim; ++r) {
                const double v = nodes[i * dim + r];
                mn[r] = std::min(mn[r], v); mx[r] = std::max(mx[r], v);
            }
        double ext = 0.0;
        for (int r = 0; r < dim; ++r) ext = std::max(ext, mx[r] - mn[r]);
        if (dim == 2) { mn[2] = 0.0; }
        const double invScale = ext > 0 ? 1.0 / ext : 0.0;

        DevBuf<int32_t> repNode((size_t)c->nDofs);
        MFEM_CUDA(cudaMemsetAsync(repNode, 0x7f, repNode.bytes(), s));   // 0x7f7f7f7f > any id
        k_rep_node<<<grid_for(nNodes, 256), 256, 0, s>>>(nNodes, dofOfNode, repNode);
        DevBuf<uint64_t> keys((size_t)c->nDofs), keysOut((size_t)c->nDofs);
        DevBuf<int32_t> ids((size_t)c->nDofs);
        k_morton_keys<<<grid_for(c->nDofs, 256), 256, 0, s>>>(dim, c->nDofs, repNode, c->nodes, mn[0], mn[1], mn[2],
                                                              invScale, keys, ids);
        c->launches += 2;
        size_t tmpBytes = 0;
        MFEM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys.p, keysOut.p, ids.p, c->int2ext.p,
                                                  c->nDofs, 0, 64, s));
        DevBuf<uint8_t> tmp(tmpBytes);
        MFEM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, keys.p, keysOut.p, ids.p, c->int2ext.p, c->nDofs,
                                                  0, 64, s));
        MFEM_CUDA(cudaStreamSynchronize(s));
    } else {
        k_iota<<<grid_for(c->nDofs, 256), 256, 0, s>>>(c->nDofs, c->int2ext);
        c->launches++;
    }
    k_invert_perm<<<grid_for(c->nDofs, 256), 256, 0, s>>>(c->nDofs, c->int2ext, c->ext2int);
    c->nodeDof.alloc((size_t)nNodes);
    k_node_dof<<<grid_for(nNodes, 256), 256, 0, s>>>(nNodes, dofOfNode, c->ext2int, c->nodeDof);
    c->elemDof.alloc((size_t)nElems * npe);
    k_elem_dof<<<grid_for(nElems * npe, 256), 256, 0, s>>>(nElems * npe, c->elemNodes, c->nodeDof, c->elemDof);
    c->launches += 3;
    MFEM_CUDA(cudaStreamSynchronize(s));

    compute_geometry(c);
}

void compute_geometry(mfem_b200_ctx *c) {
    cudaStream_t s = c->stream;
    const int GS = 1 + c->N * (c->N + 1);
    if (c->geom.n != (size_t)c->nElems * GS) c->geom.alloc((size_t)c->nElems * GS);
    DevBuf<int> neg(1);
    MFEM_CUDA(cudaMemsetAsync(neg, 0, sizeof(int), s));
    if (c->N == 3)
        k_elem_geom<3><<<grid_for(c->nElems, 256), 256, 0, s>>>(c->nElems, c->npe, c->elemNodes, c->nodes, c->geom, neg);
    else
        k_elem_geom<2><<<grid_for(c->nElems, 256), 256, 0, s>>>(c->nElems, c->npe, c->elemNodes, c->nodes, c->geom, neg);
    c->launches++;
    int nneg = 0;
    MFEM_CUDA(cudaMemcpyAsync(&nneg, neg, sizeof(int), cudaMemcpyDeviceToHost, s));
    MFEM_CUDA(cudaStreamSynchronize(s));
    MFEM_CUDA(cudaGetLastError());
    c->geomValid = true;
    c->geomPValid = false;
    c->mfGeomValid = false;
    c->valuesValid = false;
    if (nneg > 0)
        throw CudaError(MFEM_B200_ERR_NEG_VOLUME,
                        "Found " + std::to_string(nneg) +
                            " elements with negative volume...\nMesh has negatively oriented elements.\n"
                            "Correct with: mesh_convert --reorientNegativeElements.");
}

// ---------------------------------------------------------------------------
// Plan of the matrix-free operator (matfree.inl).  Internal DoF ids follow a Morton curve (setup_mesh) while the
// elements keep the caller's order; the gather kernel sweeps the DoF rows in id order, so with the caller's element
// order the slots of a row lie anywhere in elemY (ncu on the 10.2 M-element grid: 10.3 GB of DRAM reads for 3.3 GB of
// slots, L2 hit rate 10 %).  Ordering the elements by the mean id of their DoFs makes the slots a row needs one moving
// window of elemY.  Private to the operator: copies of elemDof / geomP in that order, and the incidence list rebuilt
// in terms of (position, local node); element-indexed inputs and outputs of the library keep the caller's order.
// ---------------------------------------------------------------------------
__global__ void k_mf_elem_keys(int64_t nElems, int npe, const int32_t *__restrict__ elemDof, uint32_t *keys) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nElems) return;
    int64_t sum = 0;
    for (int i = 0; i < npe; ++i) sum += elemDof[e * npe + i];
    keys[e] = (uint32_t)(sum / npe);
}
__global__ void k_mf_permute_dofs(int64_t n, int npe, const int32_t *__restrict__ perm, const int32_t *__restrict__ elemDof,
                                  int32_t *__restrict__ out) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int64_t k = t / npe;
    const int i = (int)(t - k * npe);
    out[t] = elemDof[(int64_t)perm[k] * npe + i];
}
__global__ void k_mf_permute_geom(int64_t nElems, const int32_t *__restrict__ perm, const double *__restrict__ geomP,
                                  double *__restrict__ out) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // one 32-byte slot per thread
    if (t >= nElems * 4) return;
    const int64_t k = t >> 2;
    const int a = (int)(t & 3);
    const double4 v = *reinterpret_cast<const double4 *>(geomP + (int64_t)perm[k] * 16 + a * 4);
    *reinterpret_cast<double4 *>(out + t * 4) = v;
}

void build_mf_plan(mfem_b200_ctx *c) {
    cudaStream_t s = c->stream;
    const int npe = c->npe;
    const int64_t nE = c->nElems, nInc = nE * npe;
    if (!c->mfPlanValid) {
        MFEM_REQUIRE(c->patternValid && c->totalInc == nInc, MFEM_B200_ERR_INVALID, "matrix-free plan: no pattern");
        ScopedTimer timer(c, "Matrix-free Plan");
        {
            DevBuf<uint32_t> keys((size_t)nE), keysOut((size_t)nE);
            DevBuf<int32_t> ids((size_t)nE);
            c->mfPerm.alloc((size_t)nE);
            k_mf_elem_keys<<<grid_for(nE, 256), 256, 0, s>>>(nE, npe, c->elemDof, keys);
            k_iota<<<grid_for(nE, 256), 256, 0, s>>>(nE, ids);
            size_t tmpBytes = 0;
            const int bits = bits_for(c->nDofs + 1);
            MFEM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys.p, keysOut.p, ids.p, c->mfPerm.p, nE, 0, bits, s));
            DevBuf<uint8_t> tmp(tmpBytes);
            MFEM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, keys.p, keysOut.p, ids.p, c->mfPerm.p, nE, 0, bits, s));
            c->launches += 2;
            MFEM_CUDA(cudaStreamSynchronize(s));
        }
        c->mfElemDof.alloc((size_t)nInc);
        k_mf_permute_dofs<<<grid_for(nInc, 256), 256, 0, s>>>(nInc, npe, c->mfPerm, c->elemDof, c->mfElemDof);
        c->launches++;
        {   // incidence list in terms of (position, local node): a stable sort by DoF keeps (position, local node) order
            DevBuf<uint32_t> keysOut((size_t)nInc);
            DevBuf<int32_t> ids((size_t)nInc);
            c->mfIncList.alloc((size_t)nInc);
            k_iota<<<grid_for(nInc, 256), 256, 0, s>>>(nInc, ids);
            size_t tmpBytes = 0;
            const int dofBits = bits_for(c->nDofs + 1);
            const uint32_t *kin = reinterpret_cast<const uint32_t *>(c->mfElemDof.p);
            MFEM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, kin, keysOut.p, ids.p, c->mfIncList.p, nInc, 0, dofBits, s));
            DevBuf<uint8_t> tmp(tmpBytes);
            MFEM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, kin, keysOut.p, ids.p, c->mfIncList.p, nInc, 0, dofBits, s));
            c->launches += 2;
            MFEM_CUDA(cudaStreamSynchronize(s));
        }
        c->mfPlanValid = true;
        c->mfGeomValid = false;
    }
    if (!c->mfGeomValid) {
        ensure_packed_geometry(c);
        if (c->mfGeomP.n != (size_t)nE * 16) c->mfGeomP.alloc((size_t)nE * 16);
        k_mf_permute_geom<<<grid_for(nE * 4, 256), 256, 0, s>>>(nE, c->mfPerm, c->geomP, c->mfGeomP);
        c->launches++;
        c->mfGeomValid = true;
    }
    MFEM_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// Chunk tables of the matrix-free operator (matfree.inl k_mf_chunk).  A chunk = kMfChunk consecutive elements = one CTA.
// Neighbouring elements share most of their DoFs (a quadratic-tet mesh has 1.4 DoFs per element but 10 (element, node)
// slots), so the CTA reads each distinct x block once, sums its elements' results per distinct DoF in shared memory
// and writes ONE partial per (chunk, DoF).  Per chunk, by one CTA-wide radix sort of its (DoF, slot) pairs:
//   mfChunkDof    the distinct DoFs, ascending, fixed stride        (x staging, and the keys of the rows' partial lists)
//   mfLocalIdx    chunk-local DoF index of every slot               (element threads read x / nothing else)
//   mfCsrPtr/List extents of the chunk-local DoFs in the sorted order / rank of every slot in that order, which keeps
//                 (element, local node) order inside a DoF (fixed-order sums: bit-reproducible)
// and per DoF row the list of its partials (mfIncPtr2 / mfIncList2) for the unchanged gather kernel.
// ---------------------------------------------------------------------------
template <int NPE, int kMfChunk>
__global__ void __launch_bounds__(kMfChunk)
k_mf_chunk_tables(int64_t nElems, const int32_t *__restrict__ elemDof, uint16_t *__restrict__ localIdx,
                  uint16_t *__restrict__ csrPtr, uint16_t *__restrict__ csrList, int32_t *__restrict__ uniqueDof,
                  int32_t *__restrict__ nUnique) {
    constexpr int S = kMfChunk * NPE;
    using Sort = cub::BlockRadixSort<uint32_t, kMfChunk, NPE, uint32_t>;
    using Disc = cub::BlockDiscontinuity<uint32_t, kMfChunk>;
    using Scan = cub::BlockScan<int, kMfChunk>;
    __shared__ union { typename Sort::TempStorage sort; typename Disc::TempStorage disc; typename Scan::TempStorage scan; } tmp;
    const int64_t b = blockIdx.x;
    const int t = threadIdx.x;
    const int64_t e = b * kMfChunk + t;
    uint32_t keys[NPE], vals[NPE];
#pragma unroll
    for (int i = 0; i < NPE; ++i) {
        keys[i] = e < nElems ? (uint32_t)elemDof[e * NPE + i] : 0xffffffffu;     // padding of the last chunk sorts to the end
        vals[i] = (uint32_t)(i * kMfChunk + t);
    }
    Sort(tmp.sort).Sort(keys, vals);            // blocked arrangement: thread t holds sorted positions t*NPE ..
    __syncthreads();
    int heads[NPE];
    Disc(tmp.disc).FlagHeads(heads, keys, cub::Inequality());
    __syncthreads();
    int flag[NPE], idx[NPE], total = 0;
#pragma unroll
    for (int i = 0; i < NPE; ++i) flag[i] = (heads[i] && keys[i] != 0xffffffffu) ? 1 : 0;
    Scan(tmp.scan).ExclusiveSum(flag, idx, total);
    int nValid = 0;
#pragma unroll
    for (int i = 0; i < NPE; ++i) {
        const int pos = t * NPE + i;
        if (keys[i] == 0xffffffffu) continue;
        ++nValid;
        const int u = idx[i] + flag[i] - 1;     // chunk-local DoF index of this sorted position
        localIdx[b * S + vals[i]] = (uint16_t)u;
        csrList[b * S + vals[i]] = (uint16_t)pos;      // rank of the slot: its position in the order sorted by chunk-local DoF
        if (flag[i]) {
            csrPtr[b * (S + 1) + u] = (uint16_t)pos;
            uniqueDof[b * S + u] = (int32_t)keys[i];
        }
    }
    // number of valid sorted positions = end of the last extent
    __shared__ int sValid;
    if (t == 0) sValid = 0;
    __syncthreads();
    if (nValid) atomicAdd(&sValid, nValid);
    __syncthreads();
    if (t == 0) {
        csrPtr[b * (S + 1) + total] = (uint16_t)sValid;
        nUnique[b] = total;
    }
}

__global__ void k_mf_partial_dofs(int64_t nChunks, int S, const int32_t *__restrict__ chunkBase, const int32_t *__restrict__ uniqueDof,
                                  int32_t *__restrict__ partialDof) {
    const int64_t b = blockIdx.x;
    const int base = chunkBase[b], nu = chunkBase[b + 1] - base;
    for (int u = threadIdx.x; u < nu; u += blockDim.x) partialDof[base + u] = uniqueDof[b * S + u];
}

void build_mf_chunks(mfem_b200_ctx *c) {
    if (c->mfChunksValid && c->mfChunkElems == c->opt_mf_chunk_elems) return;
    MFEM_REQUIRE(c->nElems > 0 && c->elemDof.p, MFEM_B200_ERR_INVALID, "matrix-free chunks: no mesh");
    ScopedTimer timer(c, "Matrix-free Plan");
    cudaStream_t s = c->stream;
    const int kMfChunk = c->opt_mf_chunk_elems;
    c->mfChunkElems = kMfChunk;
    const int npe = c->npe, S = kMfChunk * npe;
    const int64_t nChunks = (c->nElems + kMfChunk - 1) / kMfChunk;
    c->mfLocalIdx.alloc((size_t)nChunks * S);
    c->mfCsrPtr.alloc((size_t)nChunks * (S + 1));
    c->mfCsrList.alloc((size_t)nChunks * S);
    c->mfChunkBase.alloc((size_t)nChunks + 1);
    c->mfChunkDof.alloc((size_t)nChunks * S);
    int32_t *uniqueDof = c->mfChunkDof.p;
    DevBuf<int32_t> nUnique((size_t)nChunks + 1);
    MFEM_CUDA(cudaMemsetAsync(nUnique, 0, nUnique.bytes(), s));
    // the element kernel reads the fixed-stride tables before it knows the chunk's count: entries past it must be valid
    MFEM_CUDA(cudaMemsetAsync(c->mfChunkDof, 0, c->mfChunkDof.bytes(), s));
    MFEM_CUDA(cudaMemsetAsync(c->mfCsrPtr, 0, c->mfCsrPtr.bytes(), s));
#define MFEM_TABLES(NPE_, CH_)                                                                                      \
    k_mf_chunk_tables<NPE_, CH_><<<(unsigned)nChunks, CH_, 0, s>>>(c->nElems, c->elemDof, c->mfLocalIdx, c->mfCsrPtr, \
                                                                    c->mfCsrList, uniqueDof, nUnique)
#define MFEM_TABLES_CH(NPE_)                                                                                       \
    do {                                                                                                           \
        if (kMfChunk == 32) MFEM_TABLES(NPE_, 32);                                                                 \
        else if (kMfChunk == 64) MFEM_TABLES(NPE_, 64);                                                            \
        else MFEM_TABLES(NPE_, 128);                                                                               \
    } while (0)
    switch (npe) {
        case 3: MFEM_TABLES_CH(3); break;
        case 4: MFEM_TABLES_CH(4); break;
        case 6: MFEM_TABLES_CH(6); break;
        default: MFEM_TABLES_CH(10); break;
    }
#undef MFEM_TABLES_CH
#undef MFEM_TABLES
    c->launches++;
    {
        size_t tmpBytes = 0;
        MFEM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, nUnique.p, c->mfChunkBase.p, nChunks + 1, s));
        DevBuf<uint8_t> tmp(tmpBytes);
        MFEM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmpBytes, nUnique.p, c->mfChunkBase.p, nChunks + 1, s));
        c->launches++;
        int32_t total = 0;
        MFEM_CUDA(cudaMemcpyAsync(&total, c->mfChunkBase.p + nChunks, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        MFEM_CUDA(cudaStreamSynchronize(s));
        MFEM_REQUIRE(total > 0, MFEM_B200_ERR_INVALID, "matrix-free chunks: empty plan");
        c->mfPartials = total;
        c->timers["Matrix-free Partials"] = (double)total;     // diagnostic, not a time
    }
    const int64_t nP = c->mfPartials;
    DevBuf<int32_t> partialDof((size_t)nP);        // compact copy: the keys of the rows' partial lists
    k_mf_partial_dofs<<<(unsigned)nChunks, 128, 0, s>>>(nChunks, S, c->mfChunkBase, uniqueDof, partialDof);
    c->launches++;
    {   // per DoF row: its partials (stable sort: ascending chunk order inside a row)
        DevBuf<uint32_t> keysOut((size_t)nP);
        DevBuf<int32_t> ids((size_t)nP);
        c->mfIncList2.alloc((size_t)nP);
        k_iota<<<grid_for(nP, 256), 256, 0, s>>>(nP, ids);
        size_t tmpBytes = 0;
        const int dofBits = bits_for(c->nDofs + 1);
        const uint32_t *kin = reinterpret_cast<const uint32_t *>(partialDof.p);
        MFEM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, kin, keysOut.p, ids.p, c->mfIncList2.p, nP, 0, dofBits, s));
        DevBuf<uint8_t> tmp(tmpBytes);
        MFEM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, kin, keysOut.p, ids.p, c->mfIncList2.p, nP, 0, dofBits, s));
        c->mfIncPtr2.alloc((size_t)c->nDofs + 1);
        k_rowptr_from_keys32<<<grid_for(c->nDofs + 1, 256), 256, 0, s>>>(c->nDofs, nP, keysOut, c->mfIncPtr2);
        c->launches += 3;
        MFEM_CUDA(cudaStreamSynchronize(s));
    }
    MFEM_CUDA(cudaGetLastError());
    c->mfChunksValid = true;
}

// Symbolic phase: sorted unique (row, col) block keys -> rowptr / colidx, and the
// DoF -> (element, local node) incidence lists the owner-gather assembly walks.
void build_pattern(mfem_b200_ctx *c) {
    if (c->patternValid) return;
    ScopedTimer timer(c, "Pattern");
    cudaStream_t s = c->stream;
    const int npe = c->npe;
    const int64_t nb = c->nDofs;
    const int64_t nPairs = c->nElems * npe * npe;
    MFEM_REQUIRE(nPairs < INT32_MAX, MFEM_B200_ERR_INVALID,
                 "pattern build: more than 2^31 element block pairs on one device; partition the mesh");
    const int dofBits = bits_for(nb + 1);

    {   // ---- block pattern + the per-block element contribution lists
        // key = (row DoF, col DoF), value = pair id e*npe*npe + i*npe + j; the stable sort keeps the
        // contributions of a block in element order (fixed summation order => reproducible values)
        DevBuf<uint64_t> keys((size_t)nPairs), keysSorted((size_t)nPairs);
        DevBuf<uint32_t> pairs((size_t)nPairs), pairsSorted((size_t)nPairs);
        k_pair_keys<<<grid_for(nPairs, 256), 256, 0, s>>>(c->nElems, npe, c->elemDof, keys);
        k_iota_u32<<<grid_for(nPairs, 256), 256, 0, s>>>(nPairs, pairs);
        c->launches += 2;
        size_t tmpBytes = 0;
        MFEM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys.p, keysSorted.p, pairs.p, pairsSorted.p, nPairs,
                                                  0, 32 + dofBits, s));
        {
            DevBuf<uint8_t> tmp(tmpBytes);
            MFEM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, keys.p, keysSorted.p, pairs.p, pairsSorted.p,
                                                      nPairs, 0, 32 + dofBits, s));
            MFEM_CUDA(cudaStreamSynchronize(s));
        }
        pairs.free();
        // run-length encode: unique keys (into `keys`, reused as output) + contributions per block
        DevBuf<int32_t> counts((size_t)nPairs);   // at most one run per pair
        DevBuf<int> nSel(1);
        tmpBytes = 0;
        MFEM_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, tmpBytes, keysSorted.p, keys.p, counts.p, nSel.p, (int)nPairs, s));
        {
            DevBuf<uint8_t> tmp(tmpBytes);
            MFEM_CUDA(cub::DeviceRunLengthEncode::Encode(tmp.p, tmpBytes, keysSorted.p, keys.p, counts.p, nSel.p,
                                                         (int)nPairs, s));
            int n = 0;
            MFEM_CUDA(cudaMemcpyAsync(&n, nSel, sizeof(int), cudaMemcpyDeviceToHost, s));
            MFEM_CUDA(cudaStreamSynchronize(s));
            c->nnzb = n;
        }
        keysSorted.free();
        c->colidx.alloc((size_t)c->nnzb + 4);     // +16 B: bulk copies round the last tile up to 16 B
        c->rowptr.alloc((size_t)nb + 1 + 2);      // +16 B pad for the same reason
        k_low32<<<grid_for(c->nnzb, 256), 256, 0, s>>>(c->nnzb, keys, c->colidx);
        k_rowptr_from_keys64<<<grid_for(nb + 1, 256), 256, 0, s>>>(nb, c->nnzb, keys, c->rowptr);
        c->launches += 2;
        MFEM_CUDA(cudaStreamSynchronize(s));
        keys.free();

        // ---- plan of the block-owner assembly
        const int64_t nChunks = (c->nnzb + kBlkChunk - 1) / kBlkChunk;
        DevBuf<int32_t> blockStart((size_t)c->nnzb);
        tmpBytes = 0;
        MFEM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, counts.p, blockStart.p, (int)c->nnzb, s));
        {
            DevBuf<uint8_t> tmp(tmpBytes);
            MFEM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmpBytes, counts.p, blockStart.p, (int)c->nnzb, s));
        }
        const int64_t nWarpRounds = nChunks * (kSegSlots / 32);
        c->planCnt.alloc((size_t)c->nnzb);
        c->planSegOff.alloc((size_t)c->nnzb);
        c->planNseg.alloc((size_t)c->nnzb);
        c->planChunkL.alloc((size_t)nChunks);
        c->planSegOrder.alloc((size_t)nChunks * kSegSlots);
        c->planWarpBase.alloc((size_t)nWarpRounds + 1);
        MFEM_CUDA(cudaMemsetAsync(c->planWarpBase, 0, c->planWarpBase.bytes(), s));
        DevBuf<int> overflow(1);
        MFEM_CUDA(cudaMemsetAsync(overflow, 0, sizeof(int), s));
        k_plan_segments<<<(unsigned)nChunks, kBlkChunk, 0, s>>>(c->nnzb, counts, c->planCnt, c->planSegOff, c->planNseg, c->planChunkL, c->planSegOrder,
                                                               c->planWarpBase, overflow, blockStart, pairsSorted, npe * npe);
        c->launches++;
        tmpBytes = 0;
        MFEM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, c->planWarpBase.p, c->planWarpBase.p, (int)(nWarpRounds + 1), s));
        int64_t planEntries = 0;
        {
            DevBuf<uint8_t> tmp(tmpBytes);
            MFEM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmpBytes, c->planWarpBase.p, c->planWarpBase.p,
                                                    (int)(nWarpRounds + 1), s));
            int ovf = 0;
            MFEM_CUDA(cudaMemcpyAsync(&planEntries, c->planWarpBase.p + nWarpRounds, 8, cudaMemcpyDeviceToHost, s));
            MFEM_CUDA(cudaMemcpyAsync(&ovf, overflow, sizeof(int), cudaMemcpyDeviceToHost, s));
            MFEM_CUDA(cudaStreamSynchronize(s));
            MFEM_REQUIRE(ovf == 0, MFEM_B200_ERR_INVALID, "pattern build: a block receives more than 65535 element contributions");
        }
        c->planEntries = planEntries;
        c->timers["Plan Padding Ratio"] = (double)planEntries / (double)nPairs;   // diagnostic, not a time
        c->planList.alloc((size_t)planEntries + 32);
        // elements to stage per chunk (TMA) and their list
        c->planElemPtr.alloc((size_t)nChunks + 1);
        MFEM_CUDA(cudaMemsetAsync(c->planElemPtr, 0, c->planElemPtr.bytes(), s));
        k_plan_count_elems<<<(unsigned)nChunks, kBlkChunk, 0, s>>>(c->nnzb, nPairs, blockStart, pairsSorted, npe * npe, c->planElemPtr);
        tmpBytes = 0;
        MFEM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, c->planElemPtr.p, c->planElemPtr.p, (int)(nChunks + 1), s));
        int64_t nStaged = 0;
        {
            DevBuf<uint8_t> tmp(tmpBytes);
            MFEM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmpBytes, c->planElemPtr.p, c->planElemPtr.p, (int)(nChunks + 1), s));
            MFEM_CUDA(cudaMemcpyAsync(&nStaged, c->planElemPtr.p + nChunks, 8, cudaMemcpyDeviceToHost, s));
            MFEM_CUDA(cudaStreamSynchronize(s));
        }
        c->planElems.alloc((size_t)nStaged + 4);
        c->timers["Plan Staged Elements Per Chunk"] = (double)nStaged / (double)nChunks;   // diagnostic, not a time
        k_plan_fill<<<(unsigned)nChunks, kBlkChunk, 0, s>>>(c->nnzb, c->planCnt, c->planChunkL, blockStart, pairsSorted,
                                                           c->planSegOrder, c->planWarpBase, c->planList, nPairs, npe * npe,
                                                           c->planElemPtr, c->planElems);
        c->launches++;
        c->planChunkRow.alloc((size_t)nChunks + 1);
        k_chunk_rows<<<grid_for(nChunks + 1, 256), 256, 0, s>>>(nChunks, nb, c->nnzb, c->rowptr, c->planChunkRow);
        c->launches += 2;
        MFEM_CUDA(cudaStreamSynchronize(s));
        MFEM_CUDA(cudaGetLastError());
    }
    {   // ---- incidence lists (stable sort keeps (element, local node) order inside a row)
        const int64_t nInc = c->nElems * npe;
        DevBuf<uint32_t> keysOut((size_t)nInc);
        DevBuf<int32_t> ids((size_t)nInc);
        c->incList.alloc((size_t)nInc);
        k_iota<<<grid_for(nInc, 256), 256, 0, s>>>(nInc, ids);
        c->launches++;
        size_t tmpBytes = 0;
        const uint32_t *kin = reinterpret_cast<const uint32_t *>(c->elemDof.p);
        MFEM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, kin, keysOut.p, ids.p, c->incList.p, nInc, 0,
                                                  dofBits, s));
        DevBuf<uint8_t> tmp(tmpBytes);
        MFEM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, kin, keysOut.p, ids.p, c->incList.p, nInc, 0,
                                                  dofBits, s));
        c->incPtr.alloc((size_t)nb + 1);
        k_rowptr_from_keys32<<<grid_for(nb + 1, 256), 256, 0, s>>>(nb, nInc, keysOut, c->incPtr);
        c->launches++;
        c->totalInc = nInc;
        const int64_t nJobs = (nInc + kAsmChunk - 1) / kAsmChunk;
        c->jobRow.alloc((size_t)nJobs + 1);
        k_job_rows<<<grid_for(nJobs + 1, 256), 256, 0, s>>>(nJobs, nb, kAsmChunk, c->incPtr, c->jobRow);
        c->launches++;
        MFEM_CUDA(cudaStreamSynchronize(s));
    }
    finish_pattern(c);
    MFEM_CUDA(cudaGetLastError());
    c->vals.alloc((size_t)c->nnzb * c->N * c->N + 2);
    c->patternValid = true;
    c->valuesValid = false;
    c->precondValid = false;
}

}  // namespace mfem

namespace mfem {

// Greedy element colouring for assembly mode 1 (elements of one colour share no DoF).
// Host-side first-fit over the element list; the colour classes are uploaded once and cached
// with the pattern.
void build_coloring(mfem_b200_ctx *c) {
    if (c->nColors > 0 && c->colorElems.n == (size_t)c->nElems) return;
    ScopedTimer timer(c, "Coloring");
    const int npe = c->npe;
    std::vector<int32_t> ed((size_t)c->nElems * npe);
    MFEM_CUDA(cudaMemcpy(ed.data(), c->elemDof, ed.size() * 4, cudaMemcpyDeviceToHost));
    constexpr int W = 4;                                   // up to 256 colours
    std::vector<uint64_t> used((size_t)c->nDofs * W, 0);
    std::vector<int32_t> color((size_t)c->nElems);
    int nColors = 0;
    for (int64_t e = 0; e < c->nElems; ++e) {
        uint64_t forb[W] = {0, 0, 0, 0};
        for (int j = 0; j < npe; ++j)
            for (int w = 0; w < W; ++w) forb[w] |= used[(size_t)ed[(size_t)e * npe + j] * W + w];
        int col = -1;
        for (int w = 0; w < W && col < 0; ++w)
            if (~forb[w]) col = w * 64 + __builtin_ctzll(~forb[w]);
        MFEM_REQUIRE(col >= 0, MFEM_B200_ERR_INVALID, "colouring needs more than 256 colours");
        color[(size_t)e] = col;
        nColors = std::max(nColors, col + 1);
        for (int j = 0; j < npe; ++j) used[(size_t)ed[(size_t)e * npe + j] * W + col / 64] |= (1ULL << (col % 64));
    }
    c->colorPtr.assign((size_t)nColors + 1, 0);
    for (int64_t e = 0; e < c->nElems; ++e) c->colorPtr[(size_t)color[(size_t)e] + 1]++;
    for (int k = 0; k < nColors; ++k) c->colorPtr[(size_t)k + 1] += c->colorPtr[(size_t)k];
    std::vector<int32_t> sorted((size_t)c->nElems);
    std::vector<int64_t> pos(c->colorPtr.begin(), c->colorPtr.end() - 1);
    for (int64_t e = 0; e < c->nElems; ++e) sorted[(size_t)pos[(size_t)color[(size_t)e]]++] = (int32_t)e;
    c->colorElems.alloc((size_t)c->nElems);
    MFEM_CUDA(cudaMemcpy(c->colorElems, sorted.data(), sorted.size() * 4, cudaMemcpyHostToDevice));
    c->nColors = nColors;
}

}  // namespace mfem
