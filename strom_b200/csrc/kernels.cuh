// kernels.cuh -- the hand-written sm_100a kernels of the tree-likelihood hot path.
//
//   (a) transition_matrices_kernel : BeagleLib beagleUpdateTransitionMatrices
//                                     (reference call site src/likelihood.hpp:357-370)
//   (b) prune_chains_kernel         : BeagleLib beagleUpdatePartials incl. per-node rescaling
//                                     (reference call site src/likelihood.hpp:372-388)
//   (c) root_loglik_kernel          : BeagleLib beagleCalculateEdgeLogLikelihoods
//                                     (reference call site src/likelihood.hpp:418-447)
//
// HBM layout (pattern axis padded to Ppad = whole thread-block tiles, so no kernel needs tail predicates):
//   partials [buffer][category][Ppad][state]  CATEGORY-MAJOR: the 32 lanes of a warp hold 32 consecutive patterns of
//                                             ONE category, one pattern x category = 4 reals = one 256-bit load (FP64),
//                                             a warp-wide load is 1 KB contiguous
//   tips     [tip][Ppad]                      1-byte codes 0..4 (4 = missing; padding is 4)
//   matrices [matrix][category][36]           16 reals row-major P_k[i][j], then the 20 reals of the tip table
//                                             T_k[s][i] = P_k[i][s] (a COLUMN of P per tip state; s = 4, the missing state,
//                                             all ones).  Every lane of a warp reads the same matrix words: shared-memory
//                                             broadcast, one wavefront per 128-bit read
//   log-scalers [slot][Ppad] FP64 (cumulative buffers and per-chain subtree sums)
//
// Every kernel takes up to kMaxBatch independent evaluations (the heated chains of an MC3 run) side by side: the
// evaluation index is a grid dimension, so K trees cost ONE set of launches (sb200EvalBatchAsync).
//
// The work is an HBM-bound 4x4 matrix-vector product (about 0.6 flop/byte): no tensor cores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "schedule.hpp"

namespace sb200 {

constexpr int kMaxCategories = 8;
constexpr int kMatStride = 36;      // reals per (matrix, category)
constexpr int kMaxBatch = 8;        // evaluations per launch set

// tip x tip table of one operation: 25 (stateA, stateB) entries of C categories x 4 states of rescaled partials, each
// entry padded by 16 bytes (entries of different state pairs then start in different shared-memory bank groups: the
// lanes of a warp read entries of different pairs), followed by the 25 rescaling maxima (padded to 32).
__host__ __device__ constexpr int ttEntry(int C, int realBytes) { return 4 * C + 16 / realBytes; }
__host__ __device__ constexpr int ttStride(int C, int realBytes) { return 25 * ttEntry(C, realBytes) + 32; }

// ------------------------------------------------------------------------------------------
// Shapes of the pruning kernel (template parameter V).  A thread owns ONE pattern and CPT of the C categories; a
// warp = 32 consecutive patterns x CPT categories; NCW = ceil(C / CPT) warps cover a pattern sub-tile (a category count
// that does not divide leaves the last warp's spare slots repeating the last category, without storing); a block holds
// TP sub-tiles.
//   kVarStream  streaming sizes: as many categories per thread as 96 registers carry without spilling -- all four of
//               GTR+G4 in FP64 (matrix reads are warp-wide broadcasts, the per-pattern maximum over categories never
//               leaves the registers, ONE reciprocal per pattern), 3+2 for five; 4-warp blocks, 5 per SM;
//   kVarWide    the same thread shape for launches of small instances with so many independent chains that even this
//               shape fills the GPU (the first launches of several evaluations batched);
//   kVarWide2   two categories per thread (two category warps exchange their maxima through shared memory): twice the
//               warps, 80 registers -- launches of small instances with enough chains for THIS shape to fill the GPU
//               (the lower levels of one rbcl738 evaluation);
//   kVarNarrow  one category per thread (a warp per category): C times the warps and a C times shorter dependent
//               instruction stream per step, for launches with few chains, where a step is bound by the latency of
//               one warp's instructions; a HELPER warp per block takes everything that is not arithmetic off the
//               category warps: it stages the matrix images and keeps the patterns' log-scalers (it reads the
//               step's maxima from the category warps' exchange);
//   kVarQuad    one STATE of one category of one pattern per thread (prune_quad_kernel: a warp = 8 patterns x 4 states
//               of a category, a block = C such warps): below one warp per SM (tiny data sets).
// ------------------------------------------------------------------------------------------
enum { kVarStream = 0, kVarWide = 1, kVarNarrow = 2, kVarQuad = 3, kVarWide2 = 4, kVarCount = 5 };
constexpr int kQuadPatterns = 8;    // patterns per block of the quad shape
#ifndef SB200_CPT64
#define SB200_CPT64 4              // (2: the round's first layout; 32.6 against 30.6 ms per 1M-pattern evaluation)
#endif
#ifndef SB200_CPT32
#define SB200_CPT32 4
#endif
__host__ __device__ constexpr int cptMost(int C, int realBytes) {       // stream / wide
    return realBytes == 8 ? (C <= SB200_CPT64 ? C : (C == 5 || C == 6) ? 3 : SB200_CPT64)
                          : (C <= SB200_CPT32 ? C : (C == 5 || C == 6) ? 3 : SB200_CPT32);
}
__host__ __device__ constexpr int cptTwo(int C, int realBytes) {        // wide2: 2, or 3+2 for five categories
    return C <= 2 ? C : (C == 5 && cptMost(C, realBytes) >= 3) ? 3 : 2;
}
__host__ __device__ constexpr int varCPT(int C, int V, int realBytes) {
    return V == kVarNarrow ? 1 : V == kVarWide2 ? cptTwo(C, realBytes) : cptMost(C, realBytes);
}
__host__ __device__ constexpr int varNCW(int C, int V, int realBytes) { return (C + varCPT(C, V, realBytes) - 1) / varCPT(C, V, realBytes); }
#ifndef SB200_STREAM_WARPS
#define SB200_STREAM_WARPS 4
#endif
#ifndef SB200_WIDE_WARPS
#define SB200_WIDE_WARPS 4
#endif
__host__ __device__ constexpr int varTP(int C, int V, int realBytes) {
    return V == kVarNarrow ? 1
         : V == kVarStream ? (SB200_STREAM_WARPS / varNCW(C, V, realBytes) > 0 ? SB200_STREAM_WARPS / varNCW(C, V, realBytes) : 1)
                           : (SB200_WIDE_WARPS / varNCW(C, V, realBytes) > 0 ? SB200_WIDE_WARPS / varNCW(C, V, realBytes) : 1);
}
__host__ __device__ constexpr int varWarps(int C, int V, int realBytes) { return varNCW(C, V, realBytes) * varTP(C, V, realBytes); }   // compute warps
// Narrow shape: one more warp per block that does no arithmetic.  It stages the matrix images of the steps ahead and keeps
// the log-scalers of the block's 32 patterns, so that the category warps -- whose instruction stream IS the duration of a
// step at the top of a small tree -- carry neither (SB200_HELPER_WARP=0: the category warps do both, as in the other shapes).
#ifndef SB200_HELPER_WARP
#define SB200_HELPER_WARP 1
#endif
__host__ __device__ constexpr int varHelperWarps(int V) { return (V == kVarNarrow && SB200_HELPER_WARP) ? 1 : 0; }
__host__ __device__ constexpr int varThreads(int C, int V, int realBytes) { return 32 * (varWarps(C, V, realBytes) + varHelperWarps(V)); }
__host__ __device__ constexpr int varPatterns(int C, int V, int realBytes) { return 32 * varTP(C, V, realBytes); }   // patterns per block
#ifndef SB200_REG_BUDGET
#define SB200_REG_BUDGET 96
#endif
#ifndef SB200_REG_BUDGET_MOST
#define SB200_REG_BUDGET_MOST 96
#endif
#ifndef SB200_REG_BUDGET_TIPS
#define SB200_REG_BUDGET_TIPS 80
#endif
__host__ __device__ constexpr int varBlocksPerSm(int C, int V, bool tips, int realBytes) {
    // registers per thread the kernel is compiled for: 80 for the tips-only variants, 96 otherwise (no spills with four FP64
    // categories per thread; 5 resident blocks of 4 warps).  Measured, rbcl738 single / 8 batched chains / 250k patterns:
    // 64 + 80 registers 85.6 us / 318 us / 7.36 ms, 80 + 96 (kept) 81.2 / 306 / 7.20, 80 + 112 80.2 / 308 / 7.22, 128 80.2 / 317 / 7.35:
    // the launches are bound by latency, not by occupancy, and a wider register window lets ptxas overlap more of a step
    const int regs = tips ? SB200_REG_BUDGET_TIPS : (realBytes == 8 && varCPT(C, V, realBytes) >= 3) ? SB200_REG_BUDGET_MOST : SB200_REG_BUDGET;
    const int byRegs = 65536 / (regs * varThreads(C, V, realBytes));
    return byRegs < 1 ? 1 : byRegs > 16 ? 16 : byRegs;
}
__host__ __device__ constexpr int rootPatternsPerBlock() { return 128; }

// One operation as the pruning kernels read it: built by the engine from the planner's DevOp when a plan is uploaded.
// Operands are byte offsets (from the start of the partials arena, or of the tip-state table), so a thread forms an
// address with one 64-bit add; the first 16 bytes are all that staging an operation's matrix images needs.
struct alignas(16) DevRec {
    int flags;             // OP_* bits
    int ma, mb;            // transition matrices of operand A's and B's edge
    int tt;                // OP_TIPTIP: index of the precomputed (stateA, stateB) table
    long long destOff;     // destination buffer
    long long bOff;        // operand B (partials buffer, or tip row when OP_B_TIP)
    long long aOff;        // operand A at a chain start (partials buffer, or tip row when OP_A_TIP)
    int sa, sb;            // subtree log-scaler slots to add for A / B, or -1
    int sOut;              // OP_CHAIN_END / OP_STORE_S: slot receiving the subtree log-scaler of dest, or -1
    int pad[3];
};
static_assert(sizeof(DevRec) == 64, "DevRec is four 16-byte words");

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
    const DevRec* ops;               // the plan's operations in the form the kernels read
    const Group* groups;
    long long tipStride;
    int Ppad;
    int nTips;
    int cumMode;                     // 1: the root chain overwrites the cumulative buffer (reset fused away), 0: adds
    int storeAll;                    // 0: partials that only travel in registers inside a chain are not written (sb200SetTransientPartials)
    // extents, read only by builds with -DSB200_CHECK (every operand of every operation is checked against them; a
    // violation traps): compute-sanitizer is not available on the measurement pool
    long long arenaBytes, tipBytes;
    int nMat, nTables, nSlots, nOps;
};
template <typename real>
struct PruneBatch {
    PruneArgs<real> a[kMaxBatch];
    int groupBase[kMaxBatch];        // first group of this launch in a[z].groups
    int groupCount[kMaxBatch];       // blocks with blockIdx.y >= groupCount[z] have nothing to do
};

struct MatJob {
    const DevModel* model;
    const double* rates;
    const int* matrixIndex;
    const double* edgeLength;
    void* mats;
    int count;
    int pad;
};
struct MatBatch {
    MatJob j[kMaxBatch];
};

template <typename real>
struct TableJob {
    const DevRec* ops;
    const int* ttOps;
    const real* mats;
    real* tables;
    int count;
    int pad;
};
template <typename real>
struct TableBatch {
    TableJob<real> j[kMaxBatch];
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
    int blocks;                      // blocks of this evaluation (grid.x may be larger in a batch)
    int pad;
};
template <typename real>
struct RootBatch {
    RootArgs<real> a[kMaxBatch];
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
__host__ __device__ __forceinline__ constexpr int pmIndex(int k, int e) { return k * kMatStride + e; }                     // P_k[i][j], e = 4i+j
__host__ __device__ __forceinline__ constexpr int tipIndex(int k, int s, int i) { return k * kMatStride + 16 + s * 4 + i; }   // P_k[i][s]

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
// (a) transition matrices: one thread per (edge, category, matrix entry); blockIdx.y = evaluation of the batch
// ------------------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(256)
transition_matrices_kernel(const __grid_constant__ MatBatch B, const int C) {
    pdlLaunchDependents();
    const MatJob& J = B.j[blockIdx.y];
    const int flat = blockIdx.x * blockDim.x + threadIdx.x;
    const int e = flat & 15, uk = flat >> 4;
    if (uk >= J.count * C) return;
    const int u = uk / C, k = uk - u * C;
    const int i = e >> 2, j = e & 3;
    const double t = J.edgeLength[u] * J.rates[k];
    const double* c = J.model->cmat + e * 4;
    double sum = 0.0;
#pragma unroll
    for (int x = 0; x < 4; ++x) sum += c[x] * exp(J.model->evals[x] * t);
    const real v = static_cast<real>(sum > 0.0 ? sum : 0.0);
    real* base = static_cast<real*>(J.mats) + static_cast<size_t>(J.matrixIndex[u]) * C * kMatStride;
    base[pmIndex(k, e)] = v;                               // P_k[i][j]
    base[tipIndex(k, j, i)] = v;                           // T_k[s=j][i]
    if (j == 0) base[tipIndex(k, 4, i)] = real(1);         // missing state
}

// ------------------------------------------------------------------------------------------
// (a') tip x tip tables: for an operation whose two children are tips, the rescaled partials of a
//      pattern depend only on the two tip states.  One block per such operation computes all 25
//      combinations exactly as the per-pattern code would (same products, same maximum, same
//      reciprocal), so the pruning kernel only copies.
// ------------------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(128)
tiptip_tables_kernel(const __grid_constant__ TableBatch<real> B, const int C) {
    __shared__ real tab[100 * kMaxCategories];
    __shared__ real mx[32];
    pdlLaunchDependents();
    const TableJob<real>& J = B.j[blockIdx.y];
    if (static_cast<int>(blockIdx.x) >= J.count) return;
    const DevRec op = J.ops[J.ttOps[blockIdx.x]];     // (plan data: uploaded by a copy, not produced by a kernel)
    pdlWait();                                        // the matrices are
    const real* ma = J.mats + static_cast<size_t>(op.ma) * C * kMatStride;
    const real* mb = J.mats + static_cast<size_t>(op.mb) * C * kMatStride;
    const int n = 100 * C;
    const int TTE = ttEntry(C, sizeof(real));
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
        const int combo = q / (4 * C), rem = q - combo * 4 * C, k = rem >> 2, i = rem & 3;
        const int sA = combo / 5, sB = combo - sA * 5;
        tab[q] = ma[tipIndex(k, sA, i)] * mb[tipIndex(k, sB, i)];
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
    real* out = J.tables + static_cast<size_t>(blockIdx.x) * ttStride(C, sizeof(real));
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
        const int combo = q / (4 * C), rem = q - combo * 4 * C;
        real v = tab[q];
        if (op.flags & OP_RESCALE) v *= real(1) / mx[combo];
        out[combo * TTE + rem] = v;
    }
    if (threadIdx.x < 32) out[25 * TTE + threadIdx.x] = threadIdx.x < 25 ? mx[threadIdx.x] : real(1);
}

// ------------------------------------------------------------------------------------------
// (b) pruning, chain-fused with per-node rescaling.
//
// A thread block owns one tile of patterns and one *group* of operations of the current launch
// and streams through the group as ONE software-pipelined loop.  Per step a thread
//   - reads the 4x4 matrices of its categories from shared memory (every lane the same words: broadcast;
//     images staged by cp.async two steps ahead),
//   - takes operand B from registers (256-bit loads or a 1-byte tip state requested one step ahead),
//   - takes operand A from the partials it produced in the previous step (still in registers), or,
//     at the start of a chain, from memory,
//   - multiplies, rescales by the per-pattern maximum over categories and states (in registers; across the
//     category warps of the narrow shape through shared memory), writes the new partials with 256-bit stores.
// ONE block barrier per step, placed between the maximum and the rescaling: it publishes the next step's
// matrix image, retires the previous one and carries the maxima of the category warps.
// Why one barrier is enough (compute-sanitizer's racecheck is not available on the measurement pool, so the argument
// is written down; the five thread shapes and repeated runs are tested to give identical bits):
//   * images: step j reads sm[j&1] only BEFORE barrier B_j; image j+1 sits in the other buffer, was staged after
//     B_{j-1} and is waited for (cp.async.wait_group 0, by every thread for its own copies) before B_j, so it is
//     complete and visible after B_j; image j+2 is staged into sm[j&1] only AFTER B_j, when nobody reads image j
//     any more.  The two images of the prologue are waited for and published by a barrier of their own.
//   * maxima: step j writes smax[j&1] before B_j and reads it after B_j; the next write to the same half is step
//     j+2's, after B_{j+1}, which every warp reaches only after its step-j reads.
//   * operation records: written once per chunk between two barriers, read-only afterwards.
//   * sAcc[tid] is private to its thread.
//   * helper warp (narrow shape): it is the only stager, waits for its own copies before B_j and stages image j+2 after
//     B_j like every thread above; it reads smax[j&1] after B_j like the category warps; the scaler slots and the
//     cumulative buffer of a pattern are touched by its lane alone.  It arrives at the same barriers from its own loop
//     (bar.sync 0 with the block's thread count).
// The log of each rescaling maximum is NOT taken per step: the maxima are multiplied into a running
// product that is folded into a log sum only before it could underflow, chains hand (log sum,
// product) pairs to their parents, and only the chain ending at the root takes one log per pattern.
// ------------------------------------------------------------------------------------------
template <typename real>
__device__ __forceinline__ void matVec(const real* __restrict__ P, const real (&v)[4], real (&F)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        real s = P[i * 4 + 0] * v[0];
        s += P[i * 4 + 1] * v[1];
        s += P[i * 4 + 2] * v[2];
        s += P[i * 4 + 3] * v[3];
        F[i] = s;
    }
}

// Branch-free reciprocal for the rescaling (argument is a normal number in (0, ~1]; the caller routes
// anything smaller than 1e-290 to an exact division): hardware seed + Newton steps, < 1 ulp error.
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
__device__ __noinline__ double scalerTotal(double acc, double prod) { return acc + log(prod); }

constexpr int kOpChunk = 128;                  // operations of a group held in shared memory at a time
#ifndef SB200_PREFETCH_AHEAD
#define SB200_PREFETCH_AHEAD 0
#endif
// streaming shape: steps between an operand's L2 prefetch and its register load.  0 = none (the default: with the register loads one
// step ahead and 24 warps per SM the prefetch only costs issue slots -- 7.72 ms against 7.91 per evaluation at 250k patterns)
constexpr int kPrefetchAhead = SB200_PREFETCH_AHEAD;
constexpr int kRecWords = sizeof(DevRec) / 16;

__device__ __forceinline__ void prefetchL2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
template <typename T>
__device__ __forceinline__ T* byteOffset(T* base, long long bytes) {
    return reinterpret_cast<T*>(reinterpret_cast<char*>(const_cast<typename std::remove_const<T>::type*>(base)) + bytes);
}

#ifdef SB200_CHECK
#define SB200_CHECK_REC(r, idx)                                                                                                      \
    do {                                                                                                                             \
        const long long _vec = 4 * static_cast<long long>(sizeof(real));                                                           \
        const long long _span = (static_cast<long long>(C - 1) * A.Ppad + A.Ppad) * _vec;          /* one buffer, all categories */ \
        bool _bad = (idx) < 0 || (idx) >= A.nOps || (r).ma < 0 || (r).ma >= A.nMat || (r).mb < 0 || (r).mb >= A.nMat;            \
        _bad = _bad || (r).destOff < 0 || (r).destOff + _span > A.arenaBytes;                                                     \
        if ((r).flags & OP_B_TIP) _bad = _bad || (r).bOff < 0 || (r).bOff + A.Ppad > A.tipBytes;                                  \
        else _bad = _bad || (r).bOff < 0 || (r).bOff + _span > A.arenaBytes;                                                      \
        if ((r).flags & OP_CHAIN_START) {                                                                                         \
            if ((r).flags & OP_A_TIP) _bad = _bad || (r).aOff < 0 || (r).aOff + A.Ppad > A.tipBytes;                              \
            else _bad = _bad || (r).aOff < 0 || (r).aOff + _span > A.arenaBytes;                                                  \
        }                                                                                                                         \
        if ((r).flags & OP_TIPTIP) _bad = _bad || (r).tt < 0 || (r).tt >= A.nTables;                                              \
        _bad = _bad || (r).sa >= A.nSlots || (r).sb >= A.nSlots || (r).sOut >= A.nSlots;                                          \
        if (_bad) __trap();                                                                                                       \
    } while (0)
#else
#define SB200_CHECK_REC(r, idx) ((void)0)
#endif

template <typename real, int C, int V, bool TIPS>
__global__ void __launch_bounds__(varThreads(C, V, sizeof(real)), varBlocksPerSm(C, V, TIPS, sizeof(real)))
prune_chains_kernel(const __grid_constant__ PruneBatch<real> B) {
    constexpr int CPT = varCPT(C, V, sizeof(real));      // categories per thread
    constexpr int NCW = varNCW(C, V, sizeof(real));      // category warps per pattern sub-tile
    constexpr int TP = varTP(C, V, sizeof(real));        // pattern sub-tiles (32 patterns) per block
    constexpr int THREADS = varThreads(C, V, sizeof(real));
    constexpr int G = 32 * TP;                           // patterns per block
    constexpr int MSZ = kMatStride * C;                  // reals per operand image
    constexpr int CHUNK = 16 / static_cast<int>(sizeof(real));
    constexpr int NCH = MSZ / CHUNK;                     // 16-byte chunks per operand image
    constexpr int TTE = ttEntry(C, sizeof(real));
    constexpr int TTS = ttStride(C, sizeof(real));       // reals per tip x tip table
    constexpr int NTT = TTS / CHUNK;                     // 16-byte chunks per table
    constexpr int IMG = TTS > 2 * MSZ ? TTS : 2 * MSZ;   // shared image of one step: two matrices OR one table
    constexpr bool kHelper = varHelperWarps(V) > 0;      // the block's last warp stages the images and keeps the log-scalers
    constexpr int STAGERS = kHelper ? 32 : THREADS;      // threads that stage
    constexpr int NSTG = (IMG / CHUNK + STAGERS - 1) / STAGERS;   // staging chunks per staging thread
    constexpr bool kSingleWarp = THREADS == 32;
    static_assert(!kHelper || TP == 1, "the helper warp keeps the scalers of ONE pattern sub-tile");
    constexpr bool kPrefetch = V == kVarStream && kPrefetchAhead > 0 && !TIPS;   // (small instances live in L2 anyway)
    __shared__ __align__(16) real sm[2][IMG];
    __shared__ DevRec sops[kOpChunk];
    __shared__ double sAcc[THREADS];                     // log part of the running scaler (touched rarely)
    __shared__ real smax[2][(NCW > 1 || kHelper) ? THREADS : 1];   // per-warp maxima of a step, double-buffered

    const int z = blockIdx.z;
    pdlLaunchDependents();
    if (static_cast<int>(blockIdx.y) >= B.groupCount[z]) return;
    const PruneArgs<real>& A = B.a[z];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const bool helper = kHelper && warp >= NCW * TP;     // (warp-uniform)
    const int kb = helper ? 0 : (warp % NCW) * CPT;      // first category of this thread
    const int sp = helper ? 0 : warp / NCW;              // pattern sub-tile of this warp
    constexpr bool kEven = NCW * CPT == C;               // otherwise the last category warp repeats category C-1 in its spare slots
    const bool scalerWarp = kHelper ? helper : (NCW == 1 || kb == 0);   // the warp that keeps a pattern's log-scalers
    const Group grp = A.groups[B.groupBase[z] + blockIdx.y];
    const int nOps = grp.opEnd - grp.opStart;
    const int Ppad = A.Ppad;
    const size_t pat = static_cast<size_t>(blockIdx.x) * G + sp * 32 + lane;   // this thread's pattern
    const size_t catStride = static_cast<size_t>(Ppad) * 4;         // reals between categories
    auto kcat = [&](int c) { return kEven ? kb + c : min(kb + c, C - 1); };               // category of slot c
    auto catOff = [&](int c) { return static_cast<size_t>(kcat(c) - kb) * catStride; };     // ... relative to pmine
    real* const pmine = A.partials + (static_cast<size_t>(kb) * Ppad + pat) * 4;   // category slot c of this thread: + catOff(c)
    const uint8_t* const tmine = A.tips + pat;
    double* const accMine = A.slotAcc + pat;
    double* const prodMine = A.slotProd + pat;
    const real* const mats = A.mats;

    auto blockSync = [&]() {
        if (kSingleWarp) __syncwarp();
        else if (kHelper) asm volatile("bar.sync 0, %0;" ::"n"(THREADS) : "memory");   // (the helper warp arrives from its own loop)
        else __syncthreads();
    };
    // ---- helpers ---------------------------------------------------------------------------
    // matrix images (or the tip x tip table) of one operation -> sm[buf], 16 bytes per cp.async
    auto stage = [&](int buf, const DevRec& r) {
        const int4 w = *reinterpret_cast<const int4*>(&r);           // flags, ma, mb, tt
        if (w.x & OP_TIPTIP) {
            const real* src = A.ttTables + static_cast<size_t>(w.w) * TTS;
#pragma unroll
            for (int s = 0; s < NSTG; ++s) {
                const int c = (kHelper ? lane : tid) + s * STAGERS;
                if (c < NTT) cpAsync16(&sm[buf][c * CHUNK], src + c * CHUNK);
            }
        } else {
#pragma unroll
            for (int s = 0; s < NSTG; ++s) {
                const int c = (kHelper ? lane : tid) + s * STAGERS;
                if (c < 2 * NCH) {
                    const int o = c >= NCH ? 1 : 0;
                    cpAsync16(&sm[buf][c * CHUNK], mats + static_cast<size_t>(o ? w.z : w.y) * MSZ + (c - o * NCH) * CHUNK);
                }
            }
        }
        cpAsyncCommit();
    };
    auto fetchPartial = [&](long long off, real (&v)[CPT][4]) {
        const real* base = byteOffset(pmine, off);
#pragma unroll
        for (int c = 0; c < CPT; ++c) ld4(base + catOff(c), v[c]);
    };
    // Streaming sizes: every thread asks L2 for the sectors of operand B (and A, at a chain start) it will load
    // kPrefetchAhead steps later, so that the register loads, issued one step ahead, hit L2.
    auto prefetchOperands = [&](const DevRec& r) {
        const int fl = r.flags;
        if (!(fl & OP_B_TIP)) {
            const real* base = byteOffset(pmine, r.bOff);
#pragma unroll
            for (int c = 0; c < CPT; ++c) prefetchL2(base + catOff(c));
        }
        if ((fl & (OP_CHAIN_START | OP_A_TIP)) == OP_CHAIN_START) {
            const real* base = byteOffset(pmine, r.aOff);
#pragma unroll
            for (int c = 0; c < CPT; ++c) prefetchL2(base + catOff(c));
        }
    };

    real X[CPT][4], Y[CPT][4];
    int sX = 0, sY = 0;
    double prod = 1.0, pendAcc = 0.0, pendProd = 1.0;    // pend*: scaler of the next step's sibling chain, if any
    sAcc[tid] = 0.0;

    // (acc, prod) *= factor, keeping prod >= 1e-200.  `factor` may be denormal (then prod*factor rounds to 0 and
    // the fold path takes the logs separately).
    auto accumulate = [&](double addAcc, double factor) {
        const double t = prod * factor;
        if (addAcc != 0.0) sAcc[tid] += addAcc;
        if (t < 1e-200) {
            sAcc[tid] += foldedLog(prod, factor);
            prod = 1.0;
        } else {
            prod = t;
        }
    };
    // operand A of a chain start, and the scaler of the chain that produced it.  A TIP for A is asked for earlier, with
    // operand B (fetchTipA): it needs no register the current step still uses, and at the end of the step its latency
    // would sit on the critical path of the next one (every other step of a tips-only launch starts a chain).
    // (each in two parts: what the arithmetic warps do, and what the warp that keeps the scalers does -- the same warp
    // unless the shape has a helper warp)
    auto tipAOperand = [&](const DevRec& r) {
        if ((r.flags & OP_CHAIN_START) && (TIPS || (r.flags & OP_A_TIP))) sX = *byteOffset(tmine, r.aOff);
    };
    auto startChainOperand = [&](const DevRec& r) {
        if (!TIPS && !(r.flags & OP_A_TIP)) fetchPartial(r.aOff, X);
    };
    auto startChainScaler = [&](const DevRec& r) {
        if (!TIPS) {
            const int sa = r.sa;
            if (sa >= 0) {
                const double a = accMine[static_cast<size_t>(sa) * Ppad];
                const double pr = prodMine[static_cast<size_t>(sa) * Ppad];
                accumulate(a, pr);
            }
        }
    };
    // operand B (and its chain's scaler) of an operation, one step ahead of its use
    auto fetchBOperand = [&](const DevRec& r) {
        if (TIPS || (r.flags & OP_B_TIP)) sY = *byteOffset(tmine, r.bOff);
        else fetchPartial(r.bOff, Y);
    };
    auto fetchBScaler = [&](const DevRec& r) {
        if (!TIPS) {
            const int sb = r.sb;
            if (sb >= 0) {
                pendAcc = accMine[static_cast<size_t>(sb) * Ppad];
                pendProd = prodMine[static_cast<size_t>(sb) * Ppad];
            }
        }
    };
    const bool inlineScalers = !kHelper && scalerWarp;   // the arithmetic warps keep the scalers themselves
    // the forms the arithmetic warps use inside a step
    auto fetchTipA = [&](const DevRec& r) { tipAOperand(r); };
    auto startChain = [&](const DevRec& r) {
        startChainOperand(r);
        if (inlineScalers) startChainScaler(r);
    };
    auto fetchB = [&](const DevRec& r) {
        fetchBOperand(r);
        if (inlineScalers) fetchBScaler(r);
    };

#ifdef SB200_TRACE
#ifndef SB200_TRACE_TID
#define SB200_TRACE_TID 0
#endif
#define SB200_TICK(slot) if (A.trace && tid == SB200_TRACE_TID && blockIdx.x == 0 && blockIdx.y == 0 && z == 0 && j < 60) A.trace[j * 8 + (slot)] = clock64();
#else
#define SB200_TICK(slot)
#endif
    // the running scaler after operation j: times this step's maximum; a chain end hands it on
    auto scalerTail = [&](const int flags, const int j, const double mfac) {
        // deferred log: multiply the maxima, fold into the log sum only before an underflow
        {
            const double t = prod * mfac;
            if (t < 1e-200) {
                sAcc[tid] += foldedLog(prod, mfac);
                prod = 1.0;
            } else {
                prod = t;
            }
        }
        if (flags & (OP_STORE_S | OP_CHAIN_END)) {
            // a chain end hands its subtree log-scaler to the parent chain (and resets); incremental plans keep
            // every node's (OP_STORE_S: store and carry on)
            const int sOut = sops[j].sOut;
            double a = sAcc[tid];
            if (sOut >= 0) {
                if (prod < 1e-100) {                     // keep a parent's product of two stored factors normal
                    a += foldedLog(prod, 1.0);
                    prod = 1.0;
                    if (!(flags & OP_CHAIN_END)) sAcc[tid] = a;
                }
                accMine[static_cast<size_t>(sOut) * Ppad] = a;
                prodMine[static_cast<size_t>(sOut) * Ppad] = prod;
            }
            if (flags & OP_CHAIN_END) {
                if ((flags & OP_ADD_TO_CUM) && A.cum != nullptr) {
                    const double total = scalerTotal(a, prod);
                    double* c = A.cum + pat;
                    if (A.cumMode) *c = total;
                    else *c += total;
                }
                sAcc[tid] = 0.0;
                prod = 1.0;
            }
        }
    };
    // ---- the helper warp's step (narrow shape): image j+2 on its way, then the scalers of operation j -------------
    // Same barrier, same order of scaler operations as the inline form below: the pending scaler of operand B, the
    // request for the next one, this step's maximum (taken from the category warps' exchange), the chain end, the
    // scaler of a chain start.
    auto helperStep = [&](const int j, const int nc) {
        const bool more = j + 1 < nc;
        const int flags = sops[j].flags;
        const int ib = j & 1;
        if (more) cpAsyncWaitAll();   // image j+1 has landed
        blockSync();
        if (j + 2 < nc) stage(j & 1, sops[j + 2]);
        if (!TIPS && !(flags & OP_TIPTIP) && sops[j].sb >= 0) accumulate(pendAcc, pendProd);
        if (more) fetchBScaler(sops[j + 1]);
        real m = smax[ib][lane];
#pragma unroll
        for (int w = 1; w < NCW; ++w) m = rmax(m, smax[ib][w * 32 + lane]);
        if (!(flags & OP_TIPTIP) && (flags & OP_RESCALE)) m = (m == real(0)) ? real(1) : m;
        scalerTail(flags, j, static_cast<double>(m));
        if (more && (sops[j + 1].flags & OP_CHAIN_START)) startChainScaler(sops[j + 1]);
    };

    // ---- one step: operation j of the chunk ------------------------------------------------------------
    // On entry image j is visible to the whole block and image j+1 is on its way.
    auto step = [&](const int j, const int nc) {
        SB200_TICK(0)
        const bool more = j + 1 < nc;
        const int flags = sops[j].flags;
        if (kPrefetch && j + 1 + kPrefetchAhead < nc) prefetchOperands(sops[j + 1 + kPrefetchAhead]);
        const int ib = j & 1;                              // image buffer of this step: A's matrices at 0, B's at MSZ (or the table)
        double mfac;                                         // this step's rescaling maximum (1 when not rescaled)
        real m = real(1);

        if (flags & OP_TIPTIP) {
            // both children are tips: copy the precomputed rescaled vector of (stateA, stateB)
            const int combo = sX * 5 + sY;
            if (more) {
                fetchB(sops[j + 1]);
                fetchTipA(sops[j + 1]);
            }
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
#pragma unroll
                for (int i = 0; i < 4; ++i) X[c][i] = sm[ib][combo * TTE + kcat(c) * 4 + i];
            }
            mfac = static_cast<double>(sm[ib][25 * TTE + combo]);
            if (kHelper) smax[ib][tid] = sm[ib][25 * TTE + combo];
        } else {
            // factor of operand A (from registers, or a column of P at a chain start with a tip), then operand B
            if (flags & OP_A_TIP) {                  // (a chain start at a tip; with tables and TIPS that is a tip x tip operation)
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) X[c][i] = sm[ib][tipIndex(kcat(c), sX, i)];
                }
            } else {
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    real F[4];
                    const int o = pmIndex(kcat(c), 0);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        real a = sm[ib][o + i * 4] * X[c][0];
                        a += sm[ib][o + i * 4 + 1] * X[c][1];
                        a += sm[ib][o + i * 4 + 2] * X[c][2];
                        a += sm[ib][o + i * 4 + 3] * X[c][3];
                        F[i] = a;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) X[c][i] = F[i];
                }
            }
            SB200_TICK(1)
            if (TIPS || (flags & OP_B_TIP)) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) X[c][i] *= sm[ib][MSZ + tipIndex(kcat(c), sY, i)];
                }
            } else {
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const int o = MSZ + pmIndex(kcat(c), 0);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        real a = sm[ib][o + i * 4] * Y[c][0];
                        a += sm[ib][o + i * 4 + 1] * Y[c][1];
                        a += sm[ib][o + i * 4 + 2] * Y[c][2];
                        a += sm[ib][o + i * 4 + 3] * Y[c][3];
                        X[c][i] *= a;
                    }
                }
            }
            if (!TIPS && inlineScalers && sops[j].sb >= 0) accumulate(pendAcc, pendProd);
            SB200_TICK(2)
            // next step's operand B (and a tip for A) is requested now; the rescaling below hides its latency
            if (more) {
                fetchB(sops[j + 1]);
                fetchTipA(sops[j + 1]);
            }
            if (flags & OP_RESCALE) {
                m = rmax(rmax(X[0][0], X[0][1]), rmax(X[0][2], X[0][3]));
#pragma unroll
                for (int c = 1; c < CPT; ++c) m = rmax(m, rmax(rmax(X[c][0], X[c][1]), rmax(X[c][2], X[c][3])));
                if (NCW > 1 || kHelper) smax[ib][tid] = m;
            } else if (kHelper) {
                smax[ib][tid] = real(1);
            }
            mfac = 1.0;
        }
        SB200_TICK(3)
        if (!kHelper && more) cpAsyncWaitAll();   // (my part of) image j+1 has landed
        blockSync();                  // image j+1 and this step's maxima visible; everyone is done with image j
        SB200_TICK(4)
        if (!kHelper && j + 2 < nc) stage(j & 1, sops[j + 2]);
        if (!(flags & OP_TIPTIP) && (flags & OP_RESCALE)) {
            if (NCW > 1) {
                m = smax[ib][(sp * NCW) * 32 + lane];
#pragma unroll
                for (int w = 1; w < NCW; ++w) m = rmax(m, smax[ib][(sp * NCW + w) * 32 + lane]);
            }
            // below kTiny the hardware reciprocal seed (flush-to-zero) is not usable: exact division instead
            constexpr real kTiny = sizeof(real) == 4 ? real(1e-30) : real(1e-290);
            m = (m == real(0)) ? real(1) : m;
            const real inv = (m < kTiny) ? real(1) / m : fastRcp(m);
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
#pragma unroll
                for (int i = 0; i < 4; ++i) X[c][i] *= inv;
            }
            mfac = static_cast<double>(m);
        }
        SB200_TICK(5)
        real* dest = byteOffset(pmine, sops[j].destOff);
        if (flags & OP_CHAIN_END) {
#pragma unroll
            for (int c = 0; c < CPT; ++c)
                if (kEven || kb + c < C) st4(dest + catOff(c), X[c]);
        } else if (A.storeAll) {
#pragma unroll
            for (int c = 0; c < CPT; ++c)
                if (kEven || kb + c < C) st4Stream(dest + catOff(c), X[c]);
        }
        if (inlineScalers) scalerTail(flags, j, mfac);
        if (more && (sops[j + 1].flags & OP_CHAIN_START)) startChain(sops[j + 1]);
        SB200_TICK(6)
    };

    // ---- the group, a chunk of operations at a time (one chunk unless the group is very long) ---------
    for (int chunk0 = 0; chunk0 < nOps; chunk0 += kOpChunk) {
        const int nc = min(kOpChunk, nOps - chunk0);
        if (chunk0 > 0) blockSync();
        {
            const int4* gops = reinterpret_cast<const int4*>(A.ops + grp.opStart + chunk0);
            int4* dst = reinterpret_cast<int4*>(sops);
            for (int w = tid; w < nc * kRecWords; w += THREADS) dst[w] = gops[w];
        }
        blockSync();
        if (chunk0 == 0) pdlWait();   // the plan is in shared memory; everything below reads what earlier kernels wrote
        for (int w = tid; w < nc; w += THREADS) SB200_CHECK_REC(sops[w], grp.opStart + chunk0 + w);
        if (!kHelper || helper) {
            stage(0, sops[0]);
            if (nc > 1) stage(1, sops[1]);
        }
        if (kHelper && helper) {
            if (sops[0].flags & OP_CHAIN_START) startChainScaler(sops[0]);
            fetchBScaler(sops[0]);
        } else {
            fetchTipA(sops[0]);
            if (sops[0].flags & OP_CHAIN_START) startChain(sops[0]);       // (a later chunk may begin in mid-chain: A is carried)
            fetchB(sops[0]);
        }
        if (kPrefetch) {
#pragma unroll
            for (int a = 1; a <= kPrefetchAhead; ++a)
                if (a < nc) prefetchOperands(sops[a]);
        }
        cpAsyncWaitAll();
        blockSync();                  // images 0 and 1 visible
        if (kHelper && helper) {
#pragma unroll 1
            for (int j = 0; j < nc; ++j) helperStep(j, nc);
        } else {
#pragma unroll 1
            for (int j = 0; j < nc; ++j) step(j, nc);
        }
    }
}

// ------------------------------------------------------------------------------------------
// (b') the same step for launches that are bound by the latency of ONE warp's instruction stream (the top of a small
//      tree: one to a few chains, a few dozen warps in the whole launch): thread = (pattern, category, STATE).
//      A warp holds 8 patterns x 4 states of one category, a block the C categories of its 8 patterns.  A thread keeps
//      the whole 4-vector of its (pattern, category) but computes ONE row of each matrix-vector product; the new vector is
//      put together again by four shuffles inside the quad; the per-pattern maximum takes two shuffles inside the quad
//      and one exchange between the category warps through shared memory.  Same arithmetic, same order of additions
//      and the same reciprocal as prune_chains_kernel: the two produce the same bits.
// ------------------------------------------------------------------------------------------
template <typename real>
__device__ __forceinline__ real quadGather(real v, int lane, int j) { return __shfl_sync(0xffffffffu, v, (lane & ~3) + j); }

template <typename real, int C>
__global__ void __launch_bounds__(32 * C)
prune_quad_kernel(const __grid_constant__ PruneBatch<real> B) {
    constexpr int THREADS = 32 * C;
    constexpr int G = kQuadPatterns;
    constexpr int MSZ = kMatStride * C;
    constexpr int CHUNK = 16 / static_cast<int>(sizeof(real));
    constexpr int NCH = MSZ / CHUNK;
    constexpr int TTE = ttEntry(C, sizeof(real));
    constexpr int TTS = ttStride(C, sizeof(real));
    constexpr int NTT = TTS / CHUNK;
    constexpr int IMG = TTS > 2 * MSZ ? TTS : 2 * MSZ;
    constexpr int NSTG = (IMG / CHUNK + THREADS - 1) / THREADS;
    constexpr bool kSingleWarp = C == 1;
    __shared__ __align__(16) real sm[2][IMG];
    __shared__ DevRec sops[kOpChunk];
    __shared__ double sAcc[THREADS];
    __shared__ real smax[2][C * G];

    const int z = blockIdx.z;
    pdlLaunchDependents();
    if (static_cast<int>(blockIdx.y) >= B.groupCount[z]) return;
    const PruneArgs<real>& A = B.a[z];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int k = tid >> 5;                              // category = warp
    const int i = lane & 3;                              // state (row of the matrices) of this thread
    const int q = lane >> 2;                             // pattern of the tile
    const bool scalerThread = k == 0 && i == 0;          // keeps and stores the pattern's log-scalers
    const Group grp = A.groups[B.groupBase[z] + blockIdx.y];
    const int nOps = grp.opEnd - grp.opStart;
    const int Ppad = A.Ppad;
    const size_t pat = static_cast<size_t>(blockIdx.x) * G + q;
    real* const pmine = A.partials + (static_cast<size_t>(k) * Ppad + pat) * 4;   // this thread's (pattern, category) vector
    const uint8_t* const tmine = A.tips + pat;
    double* const accMine = A.slotAcc + pat;
    double* const prodMine = A.slotProd + pat;
    const real* const mats = A.mats;

    auto blockSync = [&]() {
        if (kSingleWarp) __syncwarp();
        else __syncthreads();
    };
    auto stage = [&](int buf, const DevRec& r) {
        const int4 w = *reinterpret_cast<const int4*>(&r);           // flags, ma, mb, tt
        if (w.x & OP_TIPTIP) {
            const real* src = A.ttTables + static_cast<size_t>(w.w) * TTS;
#pragma unroll
            for (int s = 0; s < NSTG; ++s) {
                const int c = tid + s * THREADS;
                if (c < NTT) cpAsync16(&sm[buf][c * CHUNK], src + c * CHUNK);
            }
        } else {
#pragma unroll
            for (int s = 0; s < NSTG; ++s) {
                const int c = tid + s * THREADS;
                if (c < 2 * NCH) {
                    const int o = c >= NCH ? 1 : 0;
                    cpAsync16(&sm[buf][c * CHUNK], mats + static_cast<size_t>(o ? w.z : w.y) * MSZ + (c - o * NCH) * CHUNK);
                }
            }
        }
        cpAsyncCommit();
    };

    real X[4], Y[4];
    int sX = 0, sY = 0;
    double prod = 1.0, pendAcc = 0.0, pendProd = 1.0;
    sAcc[tid] = 0.0;
    auto accumulate = [&](double addAcc, double factor) {
        const double t = prod * factor;
        if (addAcc != 0.0) sAcc[tid] += addAcc;
        if (t < 1e-200) {
            sAcc[tid] += foldedLog(prod, factor);
            prod = 1.0;
        } else {
            prod = t;
        }
    };
    auto startChain = [&](const DevRec& r) {
        if (r.flags & OP_A_TIP) sX = *byteOffset(tmine, r.aOff);
        else ld4(byteOffset(pmine, r.aOff), X);
        if (scalerThread) {
            const int sa = r.sa;
            if (sa >= 0) {
                const double a = accMine[static_cast<size_t>(sa) * Ppad];
                const double pr = prodMine[static_cast<size_t>(sa) * Ppad];
                accumulate(a, pr);
            }
        }
    };
    auto fetchB = [&](const DevRec& r) {
        if (r.flags & OP_B_TIP) sY = *byteOffset(tmine, r.bOff);
        else ld4(byteOffset(pmine, r.bOff), Y);
        if (scalerThread) {
            const int sb = r.sb;
            if (sb >= 0) {
                pendAcc = accMine[static_cast<size_t>(sb) * Ppad];
                pendProd = prodMine[static_cast<size_t>(sb) * Ppad];
            }
        }
    };

    auto step = [&](const int j, const int nc) {
        SB200_TICK(0)
        const bool more = j + 1 < nc;
        const DevRec& rec = sops[j];
        const int flags = rec.flags;
        const real* sa = sm[j & 1];
        const real* sb = sa + MSZ;
        double mfac = 1.0;
        real m = real(1);
        real v;                                              // state i of the new vector
        if (flags & OP_TIPTIP) {
            const int combo = sX * 5 + sY;
            if (more) fetchB(sops[j + 1]);
            v = sa[combo * TTE + k * 4 + i];
            mfac = static_cast<double>(sa[25 * TTE + combo]);
        } else {
            real fa, fb;
            if (flags & OP_A_TIP) {
                fa = sa[tipIndex(k, sX, i)];
            } else {
                const real* row = sa + pmIndex(k, i * 4);
                fa = row[0] * X[0];
                fa += row[1] * X[1];
                fa += row[2] * X[2];
                fa += row[3] * X[3];
            }
            if (flags & OP_B_TIP) {
                fb = sb[tipIndex(k, sY, i)];
            } else {
                const real* row = sb + pmIndex(k, i * 4);
                fb = row[0] * Y[0];
                fb += row[1] * Y[1];
                fb += row[2] * Y[2];
                fb += row[3] * Y[3];
            }
            v = fa * fb;
            SB200_TICK(1)
            if (scalerThread && rec.sb >= 0) accumulate(pendAcc, pendProd);
            if (more) fetchB(sops[j + 1]);
            SB200_TICK(2)
            if (flags & OP_RESCALE) {
                m = rmax(v, __shfl_xor_sync(0xffffffffu, v, 1));
                m = rmax(m, __shfl_xor_sync(0xffffffffu, m, 2));
                if (C > 1 && i == 0) smax[j & 1][k * G + q] = m;
            }
        }
        SB200_TICK(3)
        if (more) cpAsyncWaitAll();
        blockSync();                  // image j+1 and this step's maxima visible; everyone is done with image j
        SB200_TICK(4)
        if (j + 2 < nc) stage(j & 1, sops[j + 2]);
        SB200_TICK(5)
        if (!(flags & OP_TIPTIP) && (flags & OP_RESCALE)) {
            if (C > 1) {
                m = smax[j & 1][q];
#pragma unroll
                for (int w = 1; w < C; ++w) m = rmax(m, smax[j & 1][w * G + q]);
            }
            constexpr real kTiny = sizeof(real) == 4 ? real(1e-30) : real(1e-290);
            m = (m == real(0)) ? real(1) : m;
            const real inv = (m < kTiny) ? real(1) / m : fastRcp(m);
            v *= inv;
            mfac = static_cast<double>(m);
        }
        SB200_TICK(6)
        // one state per lane: 8 bytes x 32 lanes = 256 contiguous bytes per warp
        real* dest = byteOffset(pmine, rec.destOff) + i;
        if (flags & OP_CHAIN_END) *dest = v;
        else if (A.storeAll) __stcs(dest, v);
        // the whole vector again in every lane of the quad: operand A of the next step
#pragma unroll
        for (int s = 0; s < 4; ++s) X[s] = quadGather(v, lane, s);
        if (scalerThread) {
            {
                const double t = prod * mfac;
                if (t < 1e-200) {
                    sAcc[tid] += foldedLog(prod, mfac);
                    prod = 1.0;
                } else {
                    prod = t;
                }
            }
            if (flags & (OP_STORE_S | OP_CHAIN_END)) {
                const int sOut = rec.sOut;
                double a = sAcc[tid];
                if (sOut >= 0) {
                    if (prod < 1e-100) {
                        a += foldedLog(prod, 1.0);
                        prod = 1.0;
                        if (!(flags & OP_CHAIN_END)) sAcc[tid] = a;
                    }
                    accMine[static_cast<size_t>(sOut) * Ppad] = a;
                    prodMine[static_cast<size_t>(sOut) * Ppad] = prod;
                }
                if (flags & OP_CHAIN_END) {
                    if ((flags & OP_ADD_TO_CUM) && A.cum != nullptr) {
                        const double total = scalerTotal(a, prod);
                        double* c = A.cum + pat;
                        if (A.cumMode) *c = total;
                        else *c += total;
                    }
                    sAcc[tid] = 0.0;
                    prod = 1.0;
                }
            }
        }
        if (more && (sops[j + 1].flags & OP_CHAIN_START)) startChain(sops[j + 1]);
        SB200_TICK(7)
    };

    for (int chunk0 = 0; chunk0 < nOps; chunk0 += kOpChunk) {
        const int nc = min(kOpChunk, nOps - chunk0);
        if (chunk0 > 0) blockSync();
        {
            const int4* gops = reinterpret_cast<const int4*>(A.ops + grp.opStart + chunk0);
            int4* dst = reinterpret_cast<int4*>(sops);
            for (int w = tid; w < nc * kRecWords; w += THREADS) dst[w] = gops[w];
        }
        blockSync();
        if (chunk0 == 0) pdlWait();
        for (int w = tid; w < nc; w += THREADS) SB200_CHECK_REC(sops[w], grp.opStart + chunk0 + w);
        stage(0, sops[0]);
        if (nc > 1) stage(1, sops[1]);
        if (sops[0].flags & OP_CHAIN_START) startChain(sops[0]);
        fetchB(sops[0]);
        cpAsyncWaitAll();
        blockSync();
#pragma unroll 1
        for (int j = 0; j < nc; ++j) step(j, nc);
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
__device__ __forceinline__ void blockReduceAndFinish(double v, int nBlocks, double* blockSums, unsigned int* ticket, double* out,
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
        isLast = (t == static_cast<unsigned int>(nBlocks) - 1u);
    }
    __syncthreads();
    if (isLast && w == 0) {
        __threadfence();
        double s = 0.0;
        for (int i = lane; i < nBlocks; i += 32) s += reinterpret_cast<volatile double*>(blockSums)[i];
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

// thread = pattern; the categories are added in category order inside the thread
template <typename real, int C>
__global__ void __launch_bounds__(rootPatternsPerBlock())
root_loglik_kernel(const __grid_constant__ RootBatch<real> B) {
    constexpr int THREADS = rootPatternsPerBlock();
    __shared__ double W[C * 20];             // W[k][s][i] = pi_i * w_k * P_k[i][s]   (tip child)
    __shared__ double M[C * 16];             // M[k][i][j] = P_k[i][j]                 (partials child)
    __shared__ double fw[C * 4];             // fw[k][i] = pi_i * w_k
    pdlLaunchDependents();
    const RootArgs<real>& A = B.a[blockIdx.y];
    if (static_cast<int>(blockIdx.x) >= A.blocks) return;
    const int tid = threadIdx.x;
    const bool childTip = A.child < A.nTips;
    pdlWait();
    {
        const real* mat = A.mats + static_cast<size_t>(A.matrix) * C * kMatStride;
        for (int q = tid; q < C * 20; q += THREADS) {
            const int kk = q / 20, e = q - kk * 20, s = e >> 2, i = e & 3;
            W[q] = A.model->freqs[i] * A.catWeights[kk] * static_cast<double>(mat[tipIndex(kk, s, i)]);
        }
        for (int q = tid; q < C * 16; q += THREADS) M[q] = static_cast<double>(mat[pmIndex(q >> 4, q & 15)]);
        for (int q = tid; q < C * 4; q += THREADS) fw[q] = A.model->freqs[q & 3] * A.catWeights[q >> 2];
        __syncthreads();
    }
    const int p = blockIdx.x * THREADS + tid;             // < Ppad
    const size_t catStride = static_cast<size_t>(A.Ppad) * 4;
    const size_t bufStride = catStride * C;
    const real* par = A.partials + static_cast<size_t>(A.parent - A.nTips) * bufStride + static_cast<size_t>(p) * 4;
    double site = 0.0;
    if (childTip) {
        const int s = A.tips[static_cast<size_t>(A.child) * A.tipStride + p];
#pragma unroll
        for (int k = 0; k < C; ++k) {
            real pv[4];
            ld4(par + k * catStride, pv);
            const double* w = W + (k * 5 + s) * 4;
            double term = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) term += static_cast<double>(pv[i]) * w[i];
            site += term;
        }
    } else {
        const real* chp = A.partials + static_cast<size_t>(A.child - A.nTips) * bufStride + static_cast<size_t>(p) * 4;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            real pv[4], chv[4];
            ld4(par + k * catStride, pv);
            ld4(chp + k * catStride, chv);
            double term = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double sj = 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) sj += M[k * 16 + i * 4 + j] * static_cast<double>(chv[j]);
                term += sj * static_cast<double>(pv[i]) * fw[k * 4 + i];
            }
            site += term;
        }
    }
    double v = 0.0;
    if (p < A.P) {
        double ll = log(site);
        if (A.cum) ll += A.cum[p];
        v = ll * A.patWeights[p];
    }
    blockReduceAndFinish<THREADS>(v, A.blocks, A.blockSums, A.ticket, A.out, A.doneFlag, A.doneSeq);
}

}  // namespace sb200
