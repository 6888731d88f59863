// patterns.cu -- site-pattern compression on the device.
//
// Replaces the step that feeds the likelihood path, Data::compressPatterns (reference src/data.hpp:102-161: a
// std::map<std::vector<int>, count> over all site columns, O(sites * taxa * log sites) and one heap vector per
// site), for alignments of the 10^6-site class.  Same result: the distinct site columns in lexicographic order
// (taxon 0 most significant, codes compared as integers) and how often each occurs.
//
//   1. LSD radix sort of the site indices, one stable counting-sort pass per taxon from the last taxon to the
//      first (keys are 1-byte state codes < 32: histogram kernel, one-block scan, stable scatter kernel);
//   2. head flags: a sorted site starts a new pattern iff its column differs from its predecessor's;
//   3. exclusive scan of the flags = pattern index of every sorted site; counts = distances between heads;
//   4. gather of one representative column per pattern.
// All HBM-bound byte/integer work: coalesced streams over the site axis, gathers that stay inside one taxon's
// row (<= a few MB, L2-resident), no tensor cores.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <vector>

#include "../../include/strom_b200.h"

namespace sb200 {
namespace {

constexpr int kKeys = 32;                 // state codes must be < kKeys
constexpr int kSortThreads = 256;
constexpr int kSortRounds = 8;            // a block sorts kSortThreads * kSortRounds consecutive elements
constexpr int kTile = kSortThreads * kSortRounds;

struct Fail {
    cudaError_t err;
};
#define PCK(call)                                   \
    do {                                            \
        cudaError_t e_ = (call);                    \
        if (e_ != cudaSuccess) throw Fail{e_};      \
    } while (0)

// per-block, per-key counts of one taxon's codes in the current site order: hist[key * nblocks + block]
__global__ void __launch_bounds__(kSortThreads)
key_histogram_kernel(const uint8_t* __restrict__ row, const unsigned* __restrict__ perm, int n, unsigned* __restrict__ hist) {
    __shared__ unsigned cnt[kKeys];
    if (threadIdx.x < kKeys) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * kTile;
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        const int i = base + r * kSortThreads + threadIdx.x;
        const bool live = i < n;
        const unsigned liveMask = __ballot_sync(0xffffffffu, live);
        if (live) {
            const unsigned key = row[perm[i]];
            const unsigned peers = __match_any_sync(liveMask, key);
            if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&cnt[key], __popc(peers));
        }
    }
    __syncthreads();
    if (threadIdx.x < kKeys) hist[threadIdx.x * gridDim.x + blockIdx.x] = cnt[threadIdx.x];
}

// exclusive scan of `m` unsigned values by ONE block (m = kKeys * nblocks, a few 10^4), in place
__global__ void __launch_bounds__(1024) single_block_scan_kernel(unsigned* __restrict__ v, int m) {
    __shared__ unsigned warpSums[32];
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < m; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned x = i < m ? v[i] : 0u;
        unsigned s = x;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, s, off);
            if ((threadIdx.x & 31) >= off) s += y;
        }
        if ((threadIdx.x & 31) == 31) warpSums[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned w = warpSums[threadIdx.x];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, w, off);
                if (threadIdx.x >= off) w += y;
            }
            warpSums[threadIdx.x] = w;
        }
        __syncthreads();
        const unsigned before = carry + (threadIdx.x >= 32 ? warpSums[(threadIdx.x >> 5) - 1] : 0u) + s - x;
        if (i < m) v[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + x;
        __syncthreads();
    }
}

// stable scatter of one counting-sort pass: out[offset(key, block) + rank among equal keys of the block] = perm[i]
__global__ void __launch_bounds__(kSortThreads)
key_scatter_kernel(const uint8_t* __restrict__ row, const unsigned* __restrict__ perm, int n, const unsigned* __restrict__ offsets,
                   unsigned* __restrict__ out) {
    __shared__ unsigned running[kKeys];                       // next free position of every key
    __shared__ unsigned warpCount[kSortThreads / 32][kKeys];  // this round: elements of every key in every warp
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < kKeys) running[threadIdx.x] = offsets[threadIdx.x * gridDim.x + blockIdx.x];
    const int base = blockIdx.x * kTile;
    for (int r = 0; r < kSortRounds; ++r) {
        for (int q = threadIdx.x; q < (kSortThreads / 32) * kKeys; q += kSortThreads) (&warpCount[0][0])[q] = 0;
        __syncthreads();
        const int i = base + r * kSortThreads + threadIdx.x;
        unsigned key = 0, site = 0, rankInWarp = 0;
        const bool live = i < n;
        const unsigned liveMask = __ballot_sync(0xffffffffu, live);
        if (live) {
            site = perm[i];
            key = row[site];
            const unsigned peers = __match_any_sync(liveMask, key);
            rankInWarp = __popc(peers & ((1u << lane) - 1u));
            if (lane == __ffs(peers) - 1) warpCount[warp][key] = __popc(peers);
        }
        __syncthreads();
        if (live) {
            unsigned pos = running[key] + rankInWarp;
            for (int w = 0; w < warp; ++w) pos += warpCount[w][key];
            out[pos] = site;
        }
        __syncthreads();
        if (threadIdx.x < kKeys) {
            unsigned total = 0;
#pragma unroll
            for (int w = 0; w < kSortThreads / 32; ++w) total += warpCount[w][threadIdx.x];
            running[threadIdx.x] += total;
        }
        __syncthreads();
    }
}

// *bad = 1 if any code is outside the key range of the counting sort
__global__ void range_check_kernel(const uint8_t* __restrict__ codes, size_t total, unsigned* __restrict__ bad) {
    bool any = false;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        any = any || codes[i] >= kKeys;
    if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0) *bad = 1u;
}

__global__ void iota_kernel(unsigned* __restrict__ v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

// head[i] = 1 iff the column of sorted site i differs from the column of sorted site i-1
__global__ void head_flags_kernel(const uint8_t* __restrict__ codes, const unsigned* __restrict__ perm, int ntaxa, int n,
                                  unsigned* __restrict__ head) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned differs = i == 0 ? 1u : 0u;
    if (i > 0) {
        const unsigned a = perm[i], b = perm[i - 1];
        for (int t = 0; t < ntaxa; ++t) {
            const uint8_t* row = codes + static_cast<size_t>(t) * n;
            if (row[a] != row[b]) { differs = 1u; break; }
        }
    }
    head[i] = differs;
}

// generic exclusive scan over n unsigned values: per-tile sums, scan of the sums (one block), tile scan + offset
__global__ void __launch_bounds__(256) tile_sums_kernel(const unsigned* __restrict__ v, int n, unsigned* __restrict__ sums) {
    __shared__ unsigned part[8];
    const int base = blockIdx.x * 2048;
    unsigned s = 0;
    for (int r = 0; r < 8; ++r) {
        const int i = base + r * 256 + threadIdx.x;
        if (i < n) s += v[i];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < 8; ++w) t += part[w];
        sums[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(256)
tile_scan_kernel(const unsigned* __restrict__ v, int n, const unsigned* __restrict__ tileOffsets, unsigned* __restrict__ out) {
    __shared__ unsigned warpSums[8];
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = tileOffsets[blockIdx.x];
    __syncthreads();
    const int base = blockIdx.x * 2048;
    for (int r = 0; r < 8; ++r) {
        const int i = base + r * 256 + threadIdx.x;
        const unsigned x = i < n ? v[i] : 0u;
        unsigned s = x;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, s, off);
            if ((threadIdx.x & 31) >= off) s += y;
        }
        if ((threadIdx.x & 31) == 31) warpSums[threadIdx.x >> 5] = s;
        __syncthreads();
        unsigned before = carry + s - x;
        for (int w = 0; w < (threadIdx.x >> 5); ++w) before += warpSums[w];
        if (i < n) out[i] = before;
        __syncthreads();
        if (threadIdx.x == 255) carry = before + x;
        __syncthreads();
    }
}

// headPos[pattern] = sorted position of the pattern's first site (+ headPos[npatterns] = n)
__global__ void head_positions_kernel(const unsigned* __restrict__ head, const unsigned* __restrict__ index, int n,
                                      unsigned* __restrict__ headPos, unsigned* __restrict__ npatterns) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (head[i]) headPos[index[i]] = i;
    if (i == n - 1) {
        const unsigned np = index[i] + head[i];
        headPos[np] = n;
        *npatterns = np;
    }
}
__global__ void counts_kernel(const unsigned* __restrict__ headPos, int np, double* __restrict__ counts) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < np) counts[p] = static_cast<double>(headPos[p + 1] - headPos[p]);
}
// out[taxon][pattern] = codes[taxon][first site of the pattern]; output rows have stride `outStride`
__global__ void gather_patterns_kernel(const uint8_t* __restrict__ codes, const unsigned* __restrict__ perm,
                                       const unsigned* __restrict__ headPos, int n, int np, size_t outStride,
                                       uint8_t* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (p < np) out[static_cast<size_t>(t) * outStride + p] = codes[static_cast<size_t>(t) * n + perm[headPos[p]]];
}

template <typename T>
struct DeviceArray {
    T* p = nullptr;
    explicit DeviceArray(size_t count) { PCK(cudaMalloc(&p, sizeof(T) * (count ? count : 1))); }
    ~DeviceArray() { cudaFree(p); }
    DeviceArray(const DeviceArray&) = delete;
    DeviceArray& operator=(const DeviceArray&) = delete;
};

}  // namespace
}  // namespace sb200

extern "C" int sb200CompressPatterns(int device, int ntaxa, int nsites, const unsigned char* codes, unsigned char* patternsOut,
                                     double* countsOut, int* patternCountOut) {
    using namespace sb200;
    if (ntaxa <= 0 || nsites <= 0 || !codes || !patternsOut || !countsOut || !patternCountOut) return BEAGLE_ERROR_OUT_OF_RANGE;
    const size_t total = static_cast<size_t>(ntaxa) * nsites;
    try {
        PCK(cudaSetDevice(device));
        cudaStream_t stream;
        PCK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        struct StreamGuard {
            cudaStream_t s;
            ~StreamGuard() { cudaStreamDestroy(s); }
        } guard{stream};
        const int n = nsites;
        const int nblocks = (n + kTile - 1) / kTile;
        const int ntiles = (n + 2047) / 2048;
        DeviceArray<uint8_t> dCodes(total);
        DeviceArray<unsigned> permA(n), permB(n), hist(static_cast<size_t>(kKeys) * nblocks), head(n), index(n), headPos(n + 1),
            tileSums(ntiles), np(1);
        PCK(cudaMemcpyAsync(dCodes.p, codes, total, cudaMemcpyHostToDevice, stream));
        PCK(cudaMemsetAsync(np.p, 0, sizeof(unsigned), stream));
        range_check_kernel<<<148 * 8, 256, 0, stream>>>(dCodes.p, total, np.p);
        unsigned bad = 0;
        PCK(cudaMemcpyAsync(&bad, np.p, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
        PCK(cudaStreamSynchronize(stream));
        if (bad) return BEAGLE_ERROR_OUT_OF_RANGE;       // a state code the 32-key counting sort cannot hold
        iota_kernel<<<(n + 255) / 256, 256, 0, stream>>>(permA.p, n);
        unsigned* cur = permA.p;
        unsigned* nxt = permB.p;
        for (int t = ntaxa - 1; t >= 0; --t) {
            const uint8_t* row = dCodes.p + static_cast<size_t>(t) * n;
            key_histogram_kernel<<<nblocks, kSortThreads, 0, stream>>>(row, cur, n, hist.p);
            single_block_scan_kernel<<<1, 1024, 0, stream>>>(hist.p, kKeys * nblocks);
            key_scatter_kernel<<<nblocks, kSortThreads, 0, stream>>>(row, cur, n, hist.p, nxt);
            unsigned* tmp = cur;
            cur = nxt;
            nxt = tmp;
        }
        PCK(cudaGetLastError());
        head_flags_kernel<<<(n + 255) / 256, 256, 0, stream>>>(dCodes.p, cur, ntaxa, n, head.p);
        tile_sums_kernel<<<ntiles, 256, 0, stream>>>(head.p, n, tileSums.p);
        single_block_scan_kernel<<<1, 1024, 0, stream>>>(tileSums.p, ntiles);
        tile_scan_kernel<<<ntiles, 256, 0, stream>>>(head.p, n, tileSums.p, index.p);
        head_positions_kernel<<<(n + 255) / 256, 256, 0, stream>>>(head.p, index.p, n, headPos.p, np.p);
        PCK(cudaGetLastError());
        unsigned npHost = 0;
        PCK(cudaMemcpyAsync(&npHost, np.p, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
        PCK(cudaStreamSynchronize(stream));
        const int P = static_cast<int>(npHost);
        DeviceArray<double> dCounts(P);
        DeviceArray<uint8_t> dOut(static_cast<size_t>(ntaxa) * P);
        counts_kernel<<<(P + 255) / 256, 256, 0, stream>>>(headPos.p, P, dCounts.p);
        gather_patterns_kernel<<<dim3((P + 255) / 256, ntaxa), 256, 0, stream>>>(dCodes.p, cur, headPos.p, n, P, static_cast<size_t>(P), dOut.p);
        PCK(cudaGetLastError());
        PCK(cudaMemcpyAsync(countsOut, dCounts.p, sizeof(double) * P, cudaMemcpyDeviceToHost, stream));
        PCK(cudaMemcpyAsync(patternsOut, dOut.p, static_cast<size_t>(ntaxa) * P, cudaMemcpyDeviceToHost, stream));
        PCK(cudaStreamSynchronize(stream));
        *patternCountOut = P;
        return BEAGLE_SUCCESS;
    } catch (const Fail& f) {
        cudaGetLastError();
        std::fprintf(stderr, "strom_b200: CUDA error %d (%s) in sb200CompressPatterns\n", static_cast<int>(f.err), cudaGetErrorString(f.err));
        return f.err == cudaErrorMemoryAllocation ? BEAGLE_ERROR_OUT_OF_MEMORY : BEAGLE_ERROR_GENERAL;
    } catch (...) {
        return BEAGLE_ERROR_UNIDENTIFIED_EXCEPTION;
    }
}
