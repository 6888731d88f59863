// kernels.cuh -- the three hand-written sm_100a kernels of the tree-likelihood hot path.
//
//   (a) transition_matrices_kernel : BeagleLib beagleUpdateTransitionMatrices
//                                     (reference call site src/likelihood.hpp:357-370)
//   (b) prune_chains_kernel         : BeagleLib beagleUpdatePartials incl. per-node rescaling
//                                     (reference call site src/likelihood.hpp:372-388)
//   (c) root_loglik_kernel          : BeagleLib beagleCalculateEdgeLogLikelihoods
//                                     (reference call site src/likelihood.hpp:418-447)
//
// HBM layout (pattern axis padded to Ppad = whole thread-block tiles, so no kernel needs tail predicates):
//   partials [buffer][Ppad][category][state]  one pattern x category = 4 reals = one 256-bit load (FP64)
//   tips     [tip][Ppad]                      1-byte codes 0..4 (4 = missing; padding is 4)
//   matrices [matrix][36*C]                   already in the shared-memory image the pruning kernel wants:
//              16*C reals  row-major P_k[i][j], interleaved over categories as [(i*4+j)/2][k][(i*4+j)%2]
//                          (the C category lanes of a warp read C consecutive 16-byte words: no bank conflict)
//              20*C reals  tip table T[k][i][s] = P_k[i][s], column s=4 all ones (the missing state); state
//                          fastest so that the 16 lanes of a shared-memory pass (4 patterns x 4 categories)
//                          fall into distinct banks for s <= 3
//   log-scalers [slot][Ppad] FP64 (cumulative buffers and per-chain subtree sums)
//
// The work is an HBM-bound 4x4 matrix-vector product (about 0.6 flop/byte): no tensor cores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "schedule.hpp"

namespace sb200 {

constexpr int kMaxCategories = 8;
constexpr int kMatStride = 36;      // reals per (matrix, category)
// tip x tip table of one operation: 25 (stateA, stateB) combinations x C categories x 4 states of
// rescaled partials, followed by the 25 rescaling maxima (padded to 32)
__host__ __device__ constexpr int ttStride(int C) { return 100 * C + 32; }
// Threads per block of the (pattern, category) pruning kernel.  Streaming sizes (R = 2 patterns per thread): 256
// threads, 3 blocks per SM.  Small pattern counts (R = 1): 128 threads, 6 blocks per SM -- a step is bound by the
// latency of one warp's instruction stream and by the block barrier, so smaller blocks and twice as many groups
// per launch shorten the sequential part (rbcl738: 108.6 -> 100.9 us per evaluation).
// Variants (template parameter V of prune_chains_kernel): patterns per thread R, threads per block, blocks per SM.
// (A third variant, 64 threads x 2 patterns for small pattern counts, was measured and dropped: rbcl738 104.4 vs
// 100.6 us.)  TIPS = the launch holds only operations whose operands from memory are tips and that read no subtree
// scalers (level 0 of every full evaluation: 37 % of the bytes of a 1000-taxon evaluation): the partials-operand
// and scaler-operand code is compiled out, which frees registers for one more resident block per SM.
// (tuning aids: resident blocks per SM of the general kernels; 7 / 8 for the small and 4 for the streaming variant were
// measured and are slower, profiles/r1_negative_results.md)
#ifndef SB200_SMALL_BPS
#define SB200_SMALL_BPS 6
#endif
#ifndef SB200_STREAM_BPS
#define SB200_STREAM_BPS 3
#endif
enum { kVariantSmall = 0, kVariantStream = 1 };
__host__ __device__ constexpr int pruneR(int V) { return V == kVariantSmall ? 1 : 2; }
__host__ __device__ constexpr int pruneThreads(int V) { return V == kVariantSmall ? 128 : 256; }
__host__ __device__ constexpr int pruneBlocksPerSm(int V, bool tips = false) {
    return V == kVariantSmall ? (tips ? 8 : SB200_SMALL_BPS) : (tips ? 4 : SB200_STREAM_BPS);
}
// Lane layout of the (pattern, category) kernels: a warp holds 32/C whole patterns, category fastest.  For
// C = 1, 2, 4, 8 that is all 32 lanes; for 3, 5 (+I+G4), 6, 7 the last 32 - (32/C)*C lanes (2, 2, 2, 4) idle: they
// shadow lane 0's pattern so that every address they form is valid, and never store.
__host__ __device__ constexpr int patternsPerWarp(int C) { return 32 / C; }
__host__ __device__ constexpr int prunePatternsPerBlock(int C, int V) { return (pruneThreads(V) / 32) * patternsPerWarp(C) * pruneR(V); }
__host__ __device__ constexpr int rootPatternsPerBlock(int C) { return (256 / 32) * patternsPerWarp(C); }
// Maximum over the C category lanes of a pattern when C is not a power of two: lane k first holds the maximum
// of `cover` lanes k, k+1, ... (cyclically); taking lane k+step's value, step = min(cover, C-cover), extends
// that to cover+step lanes.  catSteps(C) rounds; catStep(C, s) is the step of round s.
__host__ __device__ constexpr int catSteps(int C) { return C <= 1 ? 0 : C <= 2 ? 1 : C <= 4 ? 2 : 3; }
__host__ __device__ constexpr int catStep(int C, int s) {
    const int cover = (1 << s) < C ? (1 << s) : C;
    return cover < C - cover ? cover : C - cover;
}

// Per-eigen-index model block, resident in HBM and refreshed once per evaluation.
struct DevModel {
    double freqs[4];
    double cmat[64];                 // cmat[(i*4+j)*4+x] = evec[i][x] * ivec[x][j]
    double evals[4];
    double weights[kMaxCategories];
};

template <typename real>
struct PruneArgs {
    real* partials;
    const uint8_t* tips;
    const real* mats;
    double* slotAcc;                 // subtree log-scaler sums: log part   [slot][Ppad]
    double* slotProd;                //                          product part [slot][Ppad] (scaler = acc + log(prod))
    double* cum;                     // cumulative scale buffer or nullptr
    const real* ttTables;            // [tipTipCount][ttStride(C)]
    long long* trace;                // optional per-step clock samples of block (0,0) (tuning builds only)
    const DevOp* ops;
    const Group* groups;
    long long tipStride;
    int P;
    int Ppad;
    int nTips;
};

template <typename real>
struct RootArgs {
    const real* partials;
    const uint8_t* tips;
    const real* mats;
    const DevModel* model;           // already offset to the frequencies' index
    const double* catWeights;        // weights of the requested index
    const double* patWeights;
    const double* cum;               // or nullptr
    double* blockSums;
    unsigned int* ticket;
    double* out;
    unsigned long long* doneFlag;    // pinned host word that receives doneSeq after `out` (host memory) was written, or nullptr
    unsigned long long doneSeq;
    long long tipStride;
    int P;
    int Ppad;
    int nTips;
    int parent;                      // buffer index
    int child;                       // buffer index (tip or partials)
    int matrix;
};

// ------------------------------------------------------------------------------------------
// vector load/store helpers: one (pattern, category) = 4 reals
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld4(const double* p, double (&v)[4]) {
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
}
__device__ __forceinline__ void st4(double* p, const double (&v)[4]) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}

// Streaming variants: partials that no later operation re-reads from memory (every node inside a chain)
// are written with an evict-first hint so that, when the working set is near the L2 size (rbcl738: 120 MB),
// L2 keeps the chain-end buffers that WILL be re-read.
__device__ __forceinline__ void st4Stream(double* p, const double (&v)[4]) {
    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
__device__ __forceinline__ void st4Stream(float* p, const float (&v)[4]) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}

template <typename real> __device__ __forceinline__ real rmax(real a, real b) { return a > b ? a : b; }

// ------------------------------------------------------------------------------------------
// matrix image indexing (see the layout note at the top of this file)
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ constexpr int rmIndex(int C, int k, int e) { return (((e >> 1) * C + k) << 1) + (e & 1); }
__host__ __device__ __forceinline__ constexpr int tipIndex(int C, int k, int s, int i) { return 16 * C + k * 20 + i * 5 + s; }

__device__ __forceinline__ void cpAsync16(void* smemDst, const void* gmemSrc) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smemDst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmemSrc) : "memory");
}
// Asks the memory system to bring `bytes` (multiple of 16) at `gmemSrc` into L2; no destination, no wait.
__device__ __forceinline__ void bulkPrefetchL2(const void* gmemSrc, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmemSrc), "r"(bytes) : "memory");
}
// Programmatic dependent launch: a kernel lets its successor in the stream start (blocks scheduled, prologue
// running) as soon as all of its own blocks are under way; the successor must not touch anything an earlier
// kernel of the evaluation produces before pdlWait(), which returns once the predecessor has completed and its
// writes are visible.  Both are no-ops for launches without the attribute.
__device__ __forceinline__ void pdlLaunchDependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdlWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cpAsyncWaitAll() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// (a) transition matrices: one thread per (edge, category, matrix entry)
// ------------------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(256)
transition_matrices_kernel(const DevModel* __restrict__ model, const double* __restrict__ rates,
                           const int* __restrict__ matrixIndex, const double* __restrict__ edgeLength,
                           int count, int C, real* __restrict__ mats) {
    pdlLaunchDependents();
    const int flat = blockIdx.x * blockDim.x + threadIdx.x;
    const int e = flat & 15, uk = flat >> 4;
    if (uk >= count * C) return;
    const int u = uk / C, k = uk - u * C;
    const int i = e >> 2, j = e & 3;
    const double t = edgeLength[u] * rates[k];
    const double* c = model->cmat + e * 4;
    double sum = 0.0;
#pragma unroll
    for (int x = 0; x < 4; ++x) sum += c[x] * exp(model->evals[x] * t);
    const real v = static_cast<real>(sum > 0.0 ? sum : 0.0);
    real* base = mats + static_cast<size_t>(matrixIndex[u]) * C * kMatStride;
    base[rmIndex(C, k, e)] = v;                            // P_k[i][j]
    base[tipIndex(C, k, j, i)] = v;                        // T[k][s=j][i]
    if (j == 0) base[tipIndex(C, k, 4, i)] = real(1);      // missing state
}

// ------------------------------------------------------------------------------------------
// (a') tip x tip tables: for an operation whose two children are tips, the rescaled partials of a
//      pattern depend only on the two tip states.  One block per such operation computes all 25
//      combinations exactly as the per-pattern code would (same products, same maximum, same
//      reciprocal), so the pruning kernel only copies.
// ------------------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(128)
tiptip_tables_kernel(const DevOp* __restrict__ ops, const int* __restrict__ ttOps, const real* __restrict__ mats, int C,
                     real* __restrict__ tables) {
    __shared__ real tab[100 * kMaxCategories];
    __shared__ real mx[32];
    pdlLaunchDependents();
    const DevOp op = ops[ttOps[blockIdx.x]];      // (plan data: uploaded by a copy, not produced by a kernel)
    pdlWait();                                    // the matrices are
    const real* ma = mats + static_cast<size_t>(op.ma) * C * kMatStride;
    const real* mb = mats + static_cast<size_t>(op.mb) * C * kMatStride;
    const int n = 100 * C;
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
        const int combo = q / (4 * C), rem = q - combo * 4 * C, k = rem >> 2, i = rem & 3;
        const int sA = combo / 5, sB = combo - sA * 5;
        tab[q] = ma[tipIndex(C, k, sA, i)] * mb[tipIndex(C, k, sB, i)];
    }
    __syncthreads();
    if (threadIdx.x < 25) {
        real m = real(1);
        if (op.flags & OP_RESCALE) {
            m = real(0);
            for (int e = 0; e < 4 * C; ++e) m = rmax(m, tab[threadIdx.x * 4 * C + e]);
            if (m == real(0)) m = real(1);
        }
        mx[threadIdx.x] = m;
    }
    __syncthreads();
    real* out = tables + static_cast<size_t>(blockIdx.x) * ttStride(C);
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
        const int combo = q / (4 * C);
        real v = tab[q];
        if (op.flags & OP_RESCALE) v *= real(1) / mx[combo];
        out[q] = v;
    }
    if (threadIdx.x < 32) out[n + threadIdx.x] = threadIdx.x < 25 ? mx[threadIdx.x] : real(1);
}

// ------------------------------------------------------------------------------------------
// (b) pruning, chain-fused with per-node rescaling.
//
// A thread block owns one tile of patterns and one *group* of operations of the current launch
// and streams through the group as ONE software-pipelined loop.  Per step a thread
//   - reads its category's two 4x4 matrices from shared memory (image staged by cp.async one
//     step ahead, one __syncthreads per step),
//   - takes operand B from registers (256-bit load or 1-byte tip state requested one step ahead),
//   - takes operand A from the partials it produced in the previous step (still in registers), or,
//     at the start of a chain, from memory,
//   - multiplies, rescales by the per-pattern maximum over categories and states (C-lane shuffle),
//     writes the 4 new partials with one 256-bit store.
// The log of each rescaling maximum is NOT taken per step: the maxima are multiplied into a running
// product that is folded into a log sum only before it could underflow, chains hand (log sum,
// product) pairs to their parents, and only the chain ending at the root takes one log per pattern.
//     C in {1,2,4,8}: one thread per (pattern, category), R patterns per thread.
// ------------------------------------------------------------------------------------------
template <typename real, int C>
__device__ __forceinline__ void factorPartial(const real* __restrict__ sm, int k, const real (&v)[4], real (&F)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        real s = sm[rmIndex(C, k, i * 4 + 0)] * v[0];
        s += sm[rmIndex(C, k, i * 4 + 1)] * v[1];
        s += sm[rmIndex(C, k, i * 4 + 2)] * v[2];
        s += sm[rmIndex(C, k, i * 4 + 3)] * v[3];
        F[i] = s;
    }
}

template <typename real, int C>
__device__ __forceinline__ void factorTip(const real* __restrict__ sm, int k, int state, real (&F)[4]) {
    const real* T = sm + tipIndex(C, k, state, 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) F[i] = T[i * 5];
}

// Branch-free reciprocal for the rescaling (argument is a normal number in (0, ~1]; the caller routes
// anything smaller than 1e-290 to an exact division): hardware seed + Newton steps, < 1 ulp error.
// Branch-free matters: it lets the compiler interleave the R independent rescalings of a thread.
__device__ __forceinline__ double fastRcp(double m) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(m));
    double e = fma(-m, x, 1.0);
    x = fma(x, e, x);
    e = fma(-m, x, 1.0);
    x = fma(x, e, x);
    e = fma(-m, x, 1.0);
    return fma(x, e, x);
}
__device__ __forceinline__ float fastRcp(float m) {
    float x;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(m));
    const float e = fmaf(-m, x, 1.0f);
    return fmaf(x, e, x);
}

// Rare path of the deferred log: fold the running product (and the factor that would have
// underflowed it) into the log sum.  Kept out of line so the hot loop carries no log() code.
__device__ __noinline__ double foldedLog(double prod, double factor) { return log(prod) + log(factor); }
__device__ __forceinline__ void foldProduct(double& acc, double& prod, double factor) {
    acc += foldedLog(prod, factor);
    prod = 1.0;
}
__device__ __noinline__ double scalerTotal(double acc, double prod) { return acc + log(prod); }

// (acc, prod) *= factor, keeping prod >= 1e-200.  `factor` may be denormal (then prod*factor
// rounds to 0 and the fold path takes the logs separately).
__device__ __forceinline__ void accumulateFactor(double& acc, double& prod, double factor) {
    const double t = prod * factor;
    if (t < 1e-200) foldProduct(acc, prod, factor);
    else prod = t;
}

constexpr int kOpChunk = 256;                  // operations of a group held in shared memory at a time
#ifndef SB200_PREFETCH_AHEAD
#define SB200_PREFETCH_AHEAD 3
#endif
constexpr int kPrefetchAhead = SB200_PREFETCH_AHEAD;   // steps between an operand tile's L2 prefetch and its register load
constexpr int kOpWords = sizeof(DevOp) / 16;   // 16-byte words per operation

template <typename real, int C, int V, bool TIPS>
__global__ void __launch_bounds__(pruneThreads(V), pruneBlocksPerSm(V, TIPS))
prune_chains_kernel(const PruneArgs<real> A, const int groupBase, const int cumMode) {
    constexpr int R = pruneR(V);
    constexpr int kPruneThreads = pruneThreads(V);
    constexpr int LPW = patternsPerWarp(C);              // patterns per warp
    constexpr bool kAllLanes = LPW * C == 32;
    constexpr bool kPow2 = (C & (C - 1)) == 0;
    constexpr int G = (kPruneThreads / 32) * LPW;        // patterns per pass
    constexpr int MSZ = kMatStride * C;                  // reals per operand image
    constexpr int CHUNK = 16 / static_cast<int>(sizeof(real));
    constexpr int NCH = MSZ / CHUNK;                     // 16-byte chunks per operand image
    constexpr int TTS = ttStride(C);                     // reals per tip x tip table
    constexpr int NTT = TTS / CHUNK;                     // 16-byte chunks per table
    constexpr int IMG = TTS > 2 * MSZ ? TTS : 2 * MSZ;   // shared image of one step: two matrices OR one table
    constexpr int NSTG = (IMG / CHUNK + kPruneThreads - 1) / kPruneThreads;   // staging chunks per thread
    constexpr unsigned TILE_BYTES = G * R * C * 4 * sizeof(real);
    __shared__ __align__(16) real sm[2][IMG];
    __shared__ int4 sops[kOpChunk * kOpWords];
    __shared__ double sAcc[R][kPruneThreads];            // log part of the running scaler (touched rarely)

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const bool active = kAllLanes || lane < LPW * C;     // (idle lanes shadow the warp's first pattern, category 0)
    const int k = active ? lane % C : 0;
    const int g = (tid >> 5) * LPW + (active ? lane / C : 0);
    int catSrc[3];                                       // source lanes of the category maximum (C not a power of two)
#pragma unroll
    for (int s = 0; s < 3; ++s) catSrc[s] = lane - k + (k + catStep(C, s)) % C;
    pdlLaunchDependents();
    const Group grp = A.groups[groupBase + blockIdx.y];
    const int nOps = grp.opEnd - grp.opStart;
    const int Ppad = A.Ppad;
    const size_t tile0 = static_cast<size_t>(blockIdx.x) * (G * R);   // first pattern of this block's tile
    const size_t bufStride = static_cast<size_t>(Ppad) * C * 4;
    real* const pmine = A.partials + ((tile0 + g) * C + k) * 4;       // r-th pattern of this thread: + r*G*C*4
    const uint8_t* const tmine = A.tips + tile0 + g;
    double* const accMine = A.slotAcc + tile0 + g;
    double* const prodMine = A.slotProd + tile0 + g;

    // ---- helpers ---------------------------------------------------------------------------
    // matrix images (or the tip x tip table) of one operation -> sm[buf], 16 bytes per cp.async
    auto stage = [&](int buf, const int4& w0, const int4& w1, int tt) {
        if (w1.w & OP_TIPTIP) {
            const real* src = A.ttTables + static_cast<size_t>(tt) * TTS;
#pragma unroll
            for (int s = 0; s < NSTG; ++s) {
                const int c = tid + s * kPruneThreads;
                if (c < NTT) cpAsync16(&sm[buf][c * CHUNK], src + c * CHUNK);
            }
        } else {
#pragma unroll
            for (int s = 0; s < NSTG; ++s) {
                const int c = tid + s * kPruneThreads;
                if (c < 2 * NCH) {
                    const int o = c >= NCH ? 1 : 0;
                    cpAsync16(&sm[buf][c * CHUNK], A.mats + static_cast<size_t>(o ? w1.x : w0.z) * MSZ + (c - o * NCH) * CHUNK);
                }
            }
        }
        cpAsyncCommit();
    };
    auto fetchTip = [&](int src, int (&state)[R]) {
        const uint8_t* t = tmine + static_cast<size_t>(src) * A.tipStride;
#pragma unroll
        for (int r = 0; r < R; ++r) state[r] = t[r * G];
    };
    auto fetchPartial = [&](int src, real (&v)[R][4]) {
        const real* base = pmine + static_cast<size_t>(src - A.nTips) * bufStride;
#pragma unroll
        for (int r = 0; r < R; ++r) ld4(base + r * (G * C * 4), v[r]);
    };
    auto loadSlot = [&](int slot, double (&a)[R], double (&pr)[R]) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            a[r] = accMine[static_cast<size_t>(slot) * Ppad + r * G];
            pr[r] = prodMine[static_cast<size_t>(slot) * Ppad + r * G];
        }
    };
    // L2 prefetch of everything an operation will read from HBM (one bulk request per tile, issued by one
    // thread kPrefetchAhead steps early): the register loads one step ahead then hit L2.  (Dealing the requests
    // to lane 0 of different warps was measured and is slower: 104.9 vs 101.3 us on rbcl738.)
    // Bulk requests are 16-byte granular: the 1-byte tip codes of a tile that does not start / end on such a
    // boundary are asked for with the neighbouring bytes; rows are padded to a multiple of 16.
    constexpr bool kTipAligned = (G * R) % 16 == 0;
    const size_t tipLo = kTipAligned ? tile0 : (tile0 & ~static_cast<size_t>(15));
    const unsigned tipBytes = kTipAligned ? G * R : static_cast<unsigned>(((tile0 + G * R + 15) & ~static_cast<size_t>(15)) - tipLo);
    auto prefetchOperands = [&](const int4& w0, const int4& w1) {
        const real* ptile = A.partials + tile0 * C * 4;
        if (TIPS || (w1.w & OP_B_TIP)) bulkPrefetchL2(A.tips + static_cast<size_t>(w0.w) * A.tipStride + tipLo, tipBytes);
        else bulkPrefetchL2(ptile + static_cast<size_t>(w0.w - A.nTips) * bufStride, TILE_BYTES);
        if (!TIPS && w1.z >= 0) {
            bulkPrefetchL2(A.slotAcc + static_cast<size_t>(w1.z) * Ppad + tile0, G * R * 8);
            bulkPrefetchL2(A.slotProd + static_cast<size_t>(w1.z) * Ppad + tile0, G * R * 8);
        }
        if (w1.w & OP_CHAIN_START) {
            if (TIPS || (w1.w & OP_A_TIP)) bulkPrefetchL2(A.tips + static_cast<size_t>(w0.y) * A.tipStride + tipLo, tipBytes);
            else bulkPrefetchL2(ptile + static_cast<size_t>(w0.y - A.nTips) * bufStride, TILE_BYTES);
            if (!TIPS && w1.y >= 0) {
                bulkPrefetchL2(A.slotAcc + static_cast<size_t>(w1.y) * Ppad + tile0, G * R * 8);
                bulkPrefetchL2(A.slotProd + static_cast<size_t>(w1.y) * Ppad + tile0, G * R * 8);
            }
        }
    };

    const bool writer = active && k == 0;               // the lane that stores a pattern's scalers
    real X[R][4], Y[R][4];
    int sX[R], sY[R];
    double prod[R], pendAcc[R], pendProd[R];             // pend*: scaler of the next step's sibling chain, if any
#pragma unroll
    for (int r = 0; r < R; ++r) { sAcc[r][tid] = 0.0; prod[r] = 1.0; pendAcc[r] = 0.0; pendProd[r] = 1.0; sX[r] = 0; sY[r] = 0; }

    // operand A of a chain start, and the scaler of the chain that produced it
    auto startChain = [&](const int4& w0, const int4& w1) {
        if (TIPS || (w1.w & OP_A_TIP)) fetchTip(w0.y, sX);
        else fetchPartial(w0.y, X);
        if (!TIPS && w1.y >= 0) {
            double a[R], pr[R];
            loadSlot(w1.y, a, pr);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                double acc = sAcc[r][tid] + a[r];
                accumulateFactor(acc, prod[r], pr[r]);
                sAcc[r][tid] = acc;
            }
        }
    };
    // operand B (and its chain's scaler) of an operation, one step ahead of its use
    auto fetchB = [&](const int4& w0, const int4& w1) {
        if (TIPS || (w1.w & OP_B_TIP)) fetchTip(w0.w, sY);
        else fetchPartial(w0.w, Y);
        if (!TIPS && w1.z >= 0) loadSlot(w1.z, pendAcc, pendProd);
    };

    // ---- one step: operation (w0,w1) at position j of the chunk; (n0,n1) is the next one (if more) -----
#ifdef SB200_TRACE
#ifndef SB200_TRACE_TID
#define SB200_TRACE_TID 0
#endif
#define SB200_TICK(slot) if (A.trace && tid == SB200_TRACE_TID && blockIdx.x == 0 && blockIdx.y == 0 && j < 60) A.trace[j * 8 + (slot)] = clock64();
#else
#define SB200_TICK(slot)
#endif
    auto step = [&](const int j, const int nc, const int4& w0, const int4& w1, const int4& n0, const int4& n1) {
        SB200_TICK(0)
        cpAsyncWaitAll();
        SB200_TICK(1)
        __syncthreads();          // images of step j visible; everyone is done with step j-1's image
        SB200_TICK(2)
        const bool more = j + 1 < nc;
        const int flags = w1.w;
        if (more) stage((j + 1) & 1, n0, n1, (n1.w & OP_TIPTIP) ? sops[(j + 1) * kOpWords + 2].y : -1);
        if (kPrefetchAhead > 0 && tid == 0 && j + 1 + kPrefetchAhead < nc)
            prefetchOperands(sops[(j + 1 + kPrefetchAhead) * kOpWords], sops[(j + 1 + kPrefetchAhead) * kOpWords + 1]);
        const real* sa = sm[j & 1];
        const real* sb = sa + MSZ;
        real* dest = pmine + static_cast<size_t>(w0.x) * bufStride;
        double mfac[R];                                      // this step's rescaling maxima (1 when not rescaled)
        SB200_TICK(6)

        if (flags & OP_TIPTIP) {
            // both children are tips: copy the precomputed rescaled vector of (stateA, stateB)
            int combo[R];
#pragma unroll
            for (int r = 0; r < R; ++r) combo[r] = sX[r] * 5 + sY[r];
            if (more) fetchB(n0, n1);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const real* T = sa + (combo[r] * C + k) * 4;
#pragma unroll
                for (int i = 0; i < 4; ++i) X[r][i] = T[i];
                mfac[r] = static_cast<double>(sa[100 * C + combo[r]]);
            }
        } else {
            // factor of operand A (from registers, or tip table at a chain start), then operand B
            if (!TIPS && (flags & OP_A_TIP)) {       // (TIPS: a chain start with a tip for A is a tip x tip operation)
#pragma unroll
                for (int r = 0; r < R; ++r) factorTip<real, C>(sa, k, sX[r], X[r]);
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    real F[4];
                    factorPartial<real, C>(sa, k, X[r], F);
#pragma unroll
                    for (int i = 0; i < 4; ++i) X[r][i] = F[i];
                }
            }
            SB200_TICK(7)
            if (TIPS || (flags & OP_B_TIP)) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    real F[4];
                    factorTip<real, C>(sb, k, sY[r], F);
#pragma unroll
                    for (int i = 0; i < 4; ++i) X[r][i] *= F[i];
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    real F[4];
                    factorPartial<real, C>(sb, k, Y[r], F);
#pragma unroll
                    for (int i = 0; i < 4; ++i) X[r][i] *= F[i];
                }
            }
            if (!TIPS && w1.z >= 0) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    double acc = sAcc[r][tid] + pendAcc[r];
                    accumulateFactor(acc, prod[r], pendProd[r]);
                    sAcc[r][tid] = acc;
                }
            }
            SB200_TICK(3)
            // next step's operand B is requested now; the rescaling below hides its latency
            if (more) fetchB(n0, n1);

            if (flags & OP_RESCALE) {
                // stage-major over the R patterns: R independent dependency chains, no branch in between
                real m[R];
#pragma unroll
                for (int r = 0; r < R; ++r) m[r] = rmax(rmax(X[r][0], X[r][1]), rmax(X[r][2], X[r][3]));
                if (kPow2) {
#pragma unroll
                    for (int off = 1; off < C; off <<= 1) {
#pragma unroll
                        for (int r = 0; r < R; ++r) m[r] = rmax(m[r], __shfl_xor_sync(0xffffffffu, m[r], off));
                    }
                } else {
#pragma unroll
                    for (int s = 0; s < catSteps(C); ++s) {
#pragma unroll
                        for (int r = 0; r < R; ++r) m[r] = rmax(m[r], __shfl_sync(0xffffffffu, m[r], catSrc[s]));
                    }
                }
                // below this the hardware reciprocal seed (flush-to-zero) is not usable: exact division instead
                constexpr real kTiny = sizeof(real) == 4 ? real(1e-30) : real(1e-290);
                bool tiny = false;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    m[r] = (m[r] == real(0)) ? real(1) : m[r];
                    tiny |= m[r] < kTiny;
                }
                real inv[R];
                if (!tiny) {
#pragma unroll
                    for (int r = 0; r < R; ++r) inv[r] = fastRcp(m[r]);
                } else {                                     // denormal-range maximum: exact division
#pragma unroll
                    for (int r = 0; r < R; ++r) inv[r] = real(1) / m[r];
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) X[r][i] *= inv[r];
                    mfac[r] = static_cast<double>(m[r]);
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) mfac[r] = 1.0;
            }
        }
        SB200_TICK(4)
        if (active) {
            if (flags & OP_CHAIN_END) {
#pragma unroll
                for (int r = 0; r < R; ++r) st4(dest + r * (G * C * 4), X[r]);
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) st4Stream(dest + r * (G * C * 4), X[r]);
            }
        }
        {
            // deferred log: multiply the maxima, fold into the log sum only before an underflow
            double t[R];
            bool low = false;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                t[r] = prod[r] * mfac[r];
                low |= t[r] < 1e-200;
            }
            if (!low) {
#pragma unroll
                for (int r = 0; r < R; ++r) prod[r] = t[r];
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (t[r] < 1e-200) {
                        sAcc[r][tid] += foldedLog(prod[r], mfac[r]);
                        prod[r] = 1.0;
                    } else {
                        prod[r] = t[r];
                    }
                }
            }
        }
        if (flags & OP_STORE_S) {
            // incremental plans: every node keeps the log-scaler of its subtree (the chain carries on)
            const int sKeep = sops[j * kOpWords + 2].x;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                double a = sAcc[r][tid], pr = prod[r];
                if (pr < 1e-100) {
                    a += foldedLog(pr, 1.0);
                    pr = 1.0;
                }
                if (writer) {
                    accMine[static_cast<size_t>(sKeep) * Ppad + r * G] = a;
                    prodMine[static_cast<size_t>(sKeep) * Ppad + r * G] = pr;
                }
            }
        }
        if (flags & OP_CHAIN_END) {
            const int sOut = sops[j * kOpWords + 2].x;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                double a = sAcc[r][tid];
                if (sOut >= 0) {
                    if (prod[r] < 1e-100) {                 // keep a parent's product of two stored factors normal
                        a += foldedLog(prod[r], 1.0);
                        prod[r] = 1.0;
                    }
                    if (writer) {
                        accMine[static_cast<size_t>(sOut) * Ppad + r * G] = a;
                        prodMine[static_cast<size_t>(sOut) * Ppad + r * G] = prod[r];
                    }
                }
                if ((flags & OP_ADD_TO_CUM) && A.cum != nullptr && writer) {
                    const double total = scalerTotal(a, prod[r]);
                    double* c = A.cum + tile0 + g + r * G;
                    if (cumMode) *c = total;
                    else *c += total;
                }
                sAcc[r][tid] = 0.0;
                prod[r] = 1.0;
            }
        }
        if (more && (n1.w & OP_CHAIN_START)) startChain(n0, n1);
        SB200_TICK(5)
    };

    // ---- the group, a chunk of operations at a time (one chunk unless the group is very long) ---------
    for (int chunk0 = 0; chunk0 < nOps; chunk0 += kOpChunk) {
        const int nc = min(kOpChunk, nOps - chunk0);
        if (chunk0 > 0) __syncthreads();
        {
            const int4* gops = reinterpret_cast<const int4*>(A.ops + grp.opStart + chunk0);
            for (int w = tid; w < nc * kOpWords; w += kPruneThreads) sops[w] = gops[w];
        }
        __syncthreads();
        if (chunk0 == 0) pdlWait();   // the plan is in shared memory; everything below reads what earlier kernels wrote
        int4 a0 = sops[0], a1 = sops[1];
        stage(0, a0, a1, (a1.w & OP_TIPTIP) ? sops[2].y : -1);
        if (a1.w & OP_CHAIN_START) startChain(a0, a1);       // (a later chunk may begin in mid-chain: A is carried)
        fetchB(a0, a1);
        if (kPrefetchAhead > 0 && tid == 0) {
#pragma unroll
            for (int a = 1; a <= kPrefetchAhead; ++a)
                if (a < nc) prefetchOperands(sops[a * kOpWords], sops[a * kOpWords + 1]);
        }
        // two steps per iteration with the operation words ping-ponging between (a0,a1) and (b0,b1): no copies
        int j = 0;
        for (; j + 1 < nc; j += 2) {
            const int4 b0 = sops[(j + 1) * kOpWords], b1 = sops[(j + 1) * kOpWords + 1];
            step(j, nc, a0, a1, b0, b1);
            if (j + 2 < nc) {
                a0 = sops[(j + 2) * kOpWords];
                a1 = sops[(j + 2) * kOpWords + 1];
            }
            step(j + 1, nc, b0, b1, a0, a1);
        }
        if (j < nc) step(j, nc, a0, a1, a0, a1);
    }
}

// Folds the subtree scalers of several forest roots into the cumulative buffer, in list order.
__global__ void fold_roots_kernel(const double* __restrict__ slotAcc, const double* __restrict__ slotProd,
                                  const int* __restrict__ rootSlots, int nRoots, double* __restrict__ cum, int Ppad,
                                  int cumMode) {
    pdlLaunchDependents();
    pdlWait();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Ppad) return;
    double s = cumMode ? 0.0 : cum[p];
    for (int r = 0; r < nRoots; ++r) {
        const size_t i = static_cast<size_t>(rootSlots[r]) * Ppad + p;
        s += slotAcc[i] + log(slotProd[i]);
    }
    cum[p] = s;
}

// ------------------------------------------------------------------------------------------
// (c) root-edge log-likelihood: site likelihoods, pattern weights, deterministic reduction
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warpSum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    return v;
}

// Block-level sum of `v` (fixed order), then the last block to finish adds the per-block sums in
// index order, so the result does not depend on scheduling.
template <int THREADS>
__device__ __forceinline__ void blockReduceAndFinish(double v, double* blockSums, unsigned int* ticket, double* out,
                                                     unsigned long long* doneFlag, unsigned long long doneSeq) {
    __shared__ double warpPart[THREADS / 32];
    __shared__ bool isLast;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    v = warpSum(v);
    if (lane == 0) warpPart[w] = v;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < THREADS / 32; ++i) s += warpPart[i];
        blockSums[blockIdx.x] = s;
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        isLast = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (isLast && w == 0) {
        __threadfence();
        double s = 0.0;
        for (unsigned int i = lane; i < gridDim.x; i += 32) s += reinterpret_cast<volatile double*>(blockSums)[i];
        s = warpSum(s);
        if (lane == 0) {
            *out = s;
            *ticket = 0u;
            if (doneFlag) {                  // the host polls this word instead of waiting for the stream to drain
                __threadfence_system();
                *reinterpret_cast<volatile unsigned long long*>(doneFlag) = doneSeq;
            }
        }
    }
}

// Shared-memory tables of the root edge: W[k][s][i] = pi_i * w_k * P_k[i][s] (tip child),
// M[k][i][j] = P_k[i][j] and fw[k][i] = pi_i * w_k (partials child).
template <typename real, int C, int THREADS>
__device__ __forceinline__ void loadRootTables(const RootArgs<real>& A, double* W, double* M, double* fw) {
    const int tid = threadIdx.x;
    const real* mat = A.mats + static_cast<size_t>(A.matrix) * C * kMatStride;
    for (int q = tid; q < C * 20; q += THREADS) {
        const int kk = q / 20, e = q - kk * 20, s = e >> 2, i = e & 3;
        W[q] = A.model->freqs[i] * A.catWeights[kk] * static_cast<double>(mat[tipIndex(C, kk, s, i)]);
    }
    for (int q = tid; q < C * 16; q += THREADS) M[q] = static_cast<double>(mat[rmIndex(C, q >> 4, q & 15)]);
    for (int q = tid; q < C * 4; q += THREADS) fw[q] = A.model->freqs[q & 3] * A.catWeights[q >> 2];
    __syncthreads();
}

template <typename real, int C>
__global__ void __launch_bounds__(256)
root_loglik_kernel(const RootArgs<real> A) {
    constexpr int LPW = patternsPerWarp(C);
    constexpr int G = rootPatternsPerBlock(C);
    constexpr bool kPow2 = (C & (C - 1)) == 0;
    __shared__ double W[C * 20];
    __shared__ double M[C * 16];
    __shared__ double fw[C * 4];
    const int tid = threadIdx.x, lane = tid & 31;
    const bool active = LPW * C == 32 || lane < LPW * C;  // (idle lanes shadow the warp's first pattern)
    const int k = active ? lane % C : 0, g = (tid >> 5) * LPW + (active ? lane / C : 0);
    const bool childTip = A.child < A.nTips;
    pdlLaunchDependents();
    pdlWait();
    loadRootTables<real, C, 256>(A, W, M, fw);

    const int p = blockIdx.x * G + g;                     // < Ppad
    const size_t bufStride = static_cast<size_t>(A.Ppad) * C * 4;
    real par[4];
    ld4(A.partials + static_cast<size_t>(A.parent - A.nTips) * bufStride + (static_cast<size_t>(p) * C + k) * 4, par);
    double term = 0.0;
    if (childTip) {
        const int s = A.tips[static_cast<size_t>(A.child) * A.tipStride + p];
        const double* w = W + (k * 5 + s) * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) term += static_cast<double>(par[i]) * w[i];
    } else {
        real chv[4];
        ld4(A.partials + static_cast<size_t>(A.child - A.nTips) * bufStride + (static_cast<size_t>(p) * C + k) * 4, chv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double sj = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) sj += M[k * 16 + i * 4 + j] * static_cast<double>(chv[j]);
            term += sj * static_cast<double>(par[i]) * fw[k * 4 + i];
        }
    }
    if (kPow2) {
#pragma unroll
        for (int off = 1; off < C; off <<= 1) term += __shfl_xor_sync(0xffffffffu, term, off);
    } else {                                              // category 0's lane adds the C terms in category order
        const double mine = term;
        term = 0.0;
#pragma unroll
        for (int j = 0; j < C; ++j) term += __shfl_sync(0xffffffffu, mine, lane - k + j);
    }
    double v = 0.0;
    if (active && k == 0 && p < A.P) {
        double ll = log(term);
        if (A.cum) ll += A.cum[p];
        v = ll * A.patWeights[p];
    }
    blockReduceAndFinish<256>(v, A.blockSums, A.ticket, A.out, A.doneFlag, A.doneSeq);
}

}  // namespace sb200
