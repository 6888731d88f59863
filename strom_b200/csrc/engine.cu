// engine.cu -- the CUDA engine behind the boundary of include/strom_beagle.h and
// include/strom_b200.h.  One instance = one GPU, one stream, all buffers resident in HBM:
// byte-coded tips, every partials buffer, every transition matrix, the cumulative and subtree
// log-scalers.  There is no CPU fallback: every entry point that computes does so with the
// kernels of kernels.cuh or fails with a BeagleLib error code.
//
// Replaces what strom reaches through Likelihood (reference src/likelihood.hpp:165-448) and
// Model::setBeagle* (reference src/model.hpp:315-354).
//
// Every kernel takes a *set* of up to kMaxBatch instances (same device, stream, category count and precision): the
// single-instance entry points launch a set of one, sb200EvalBatchAsync launches the heated chains of an MC3 run as
// ONE set of launches (reference loop: src/strom.hpp:316-324 steps the chains one after the other).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/strom_b200.h"
#include "kernels.cuh"
#include "schedule.hpp"

#if defined(SB200_NVTX)
#include <nvtx3/nvToolsExt.h>
#define SB200_RANGE_PUSH(name) nvtxRangePushA(name)
#define SB200_RANGE_POP() nvtxRangePop()
#else
#define SB200_RANGE_PUSH(name) ((void)0)
#define SB200_RANGE_POP() ((void)0)
#endif

namespace sb200 {

struct CudaFail {
    cudaError_t err;
};
#define CK(x)                                        \
    do {                                             \
        cudaError_t _e = (x);                        \
        if (_e != cudaSuccess) throw CudaFail{_e};   \
    } while (0)

// Device new instances are created on when beagleCreateInstance gets no resource list: the calling thread's choice
// (sb200SetDevice), else the last choice any thread made.
static std::atomic<int> g_defaultDevice{0};
static thread_local int t_device = -1;

// Makes `device` current for the lifetime of the guard and restores the caller's device afterwards (a torch process, or
// a host thread driving several shards, must not find its current device changed by a library call).
struct DeviceGuard {
    int previous = -1;
    bool switched = false;
    cudaError_t enter(int device) {
        cudaError_t e = cudaGetDevice(&previous);
        if (e != cudaSuccess) return e;
        if (previous != device) {
            e = cudaSetDevice(device);
            if (e != cudaSuccess) return e;
            switched = true;
        }
        return cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(previous);
    }
};

static int envInt(const char* name, int fallback) {
    const char* f = std::getenv(name);
    return f ? std::atoi(f) : fallback;
}

// Launch on `stream`; `dependent`: the kernel follows another kernel of the same evaluation and may be scheduled
// while that one drains (programmatic dependent launch, see pdlWait in kernels.cuh).
template <typename... KArgs, typename... Args>
static void launchK(cudaStream_t stream, bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
}

// Host image of the per-call parameter block; mirrored 1:1 in HBM and refreshed by ONE copy.
struct ParamBlockHeader {
    double rates[kMaxCategories];
};

// The parameter blocks of the instances of a launch set, back to back in ONE pinned host buffer and ONE device buffer, so
// that a set's parameters go up in one copy instead of one per instance (8 copies cost 30 us before the first kernel of a
// 320 us set could start).  Shared by the instances that were put into it; freed with the last of them.
struct ParamArena {
    int device = 0;
    unsigned char* h = nullptr;
    unsigned char* d = nullptr;
    size_t slotBytes = 0;
    int slots = 0;
    cudaEvent_t copied = nullptr;
    ~ParamArena() {
        DeviceGuard guard;
        guard.enter(device);
        if (h) cudaFreeHost(h);
        if (d) cudaFree(d);
        if (copied) cudaEventDestroy(copied);
    }
};

struct Engine {
    int device = 0;
    int smCount = 148;
    bool fp32 = false;
    int nTips = 0, nPartials = 0, nCompact = 0, P = 0, nEigen = 0, nMat = 0, C = 0, nScale = 0;
    int Ppad = 0;                         // pattern axis padded to whole thread-block tiles in every buffer
    bool streaming = false;               // P x C large enough for the streaming thread shape on every launch (bigTiles)
    long long tipStride = 0;
    cudaStream_t ownStream = nullptr, stream = nullptr;
    int batchHint = 1;                    // evaluations launched side by side with this one (shapes the plan's launches)

    // HBM
    uint8_t* dTips = nullptr;
    void* dPartials = nullptr;
    void* dMats = nullptr;
    double* dPatWeights = nullptr;
    std::vector<double*> dScale;          // cumulative scale buffers, allocated on first use
    std::vector<char> scaleZeroPending;   // reset requested, memset deferred (may be fused away)
    double* dSlots = nullptr;
    size_t slotCapacity = 0;
    // parameter block: [ParamBlockHeader][DevModel x nEigen][int midx[edgeCap]][double elen[edgeCap]]
    unsigned char* hBlock = nullptr;      // pinned
    unsigned char* dBlock = nullptr;
    size_t blockBytes = 0, modelOff = 0, midxOff = 0, elenOff = 0;
    int edgeCap = 0, edgeCount = 0;
    cudaEvent_t blockCopied = nullptr;
    std::shared_ptr<ParamArena> arena;    // set: hBlock / dBlock are slice `arenaSlot` of a launch set's arena (not owned)
    int arenaSlot = -1;
    bool blockInFlight = false;
    bool blockDirty = true;               // host image newer than the device image
    // schedule
    Schedule sched;
    bool schedAccumulate = false;
    bool schedValid = false;
    int schedHint = 1;                    // batchHint the active plan's launches were shaped for
    std::vector<BeagleOperation> schedOps;
    DevRec* dOps = nullptr;
    Group* dGroups = nullptr;
    int* dRootSlots = nullptr;
    bool keepSubtreeScalers = false;      // incremental evaluation: permanent per-buffer subtree scaler slots (sb200SetSubtreeScalers)
    bool transientPartials = false;       // partials inside a chain are not written to HBM (sb200SetTransientPartials)
    std::vector<char> slotValid;          // ... and which of them hold the scaler of the partials now in the buffer
    long long* dTrace = nullptr;          // tuning builds: per-step clock samples (sb200DebugTrace)
    int* dTTOps = nullptr;                // op index of every tip x tip table
    void* dTT = nullptr;                  // tip x tip tables
    // the four plan arrays above live in ONE allocation, uploaded by one copy: [ops][groups][root slots][table ops]
    unsigned char* dPlan = nullptr;
    size_t planCap = 0, ttCap = 0;
    // Plans that are not the active one but still sit in HBM, most recently used last.  An incremental MCMC
    // alternates between the full list (every change of the model) and short path lists: without these the
    // 700-operation plan would be rebuilt and uploaded again after every tree move.
    struct ParkedPlan {
        Schedule sched;
        bool accumulate = false, valid = false;
        int hint = 1;
        std::vector<BeagleOperation> key;
        DevRec* dOps = nullptr;
        Group* dGroups = nullptr;
        int* dRootSlots = nullptr;
        int* dTTOps = nullptr;
        void* dTT = nullptr;
        unsigned char* dPlan = nullptr;
        size_t planCap = 0, ttCap = 0;
    };
    static constexpr size_t kParkedPlans = 3;
    std::vector<ParkedPlan> parked;
    void* hSchedStage = nullptr;          // pinned staging for schedule uploads
    size_t schedStageCap = 0;
    cudaEvent_t schedCopied = nullptr;
    bool schedInFlight = false;
    int lastCumIndex = BEAGLE_OP_NONE;
    // root reduction
    double* dBlockSums = nullptr;
    unsigned int* dTicket = nullptr;
    double* dOut = nullptr;
    double* hOut = nullptr;               // pinned, written by the root kernel directly: [0] the scalar, [1] a sequence word
    unsigned long long rootSeq = 0;       // sequence number of the last root launch that targets hOut
    unsigned long long awaitedSeq = 0;    // ... that readScalar has to see in hOut[1] (0: none outstanding)
    bool pollResult = true;               // SB200_POLL=0 (tuning aid): wait with cudaStreamSynchronize only
    int rootBlocks = 0;
    // last root-edge call (for replay)
    int lastParent = -1, lastChild = -1, lastMatrix = -1, lastWIdx = 0, lastFIdx = 0, lastRootCum = BEAGLE_OP_NONE;
    std::vector<char> tipSet;
    std::vector<unsigned char> tipStage;  // host staging of one tip row (int32 -> byte codes)
    long launches = 0;
    double algBytes = 0, schedBytes = 0;
    // optional CUDA-event timing of the pruning launches (bench.py's roofline leg)
    bool profiling = false;
    cudaEvent_t evP0 = nullptr, evP1 = nullptr;
    double partialsMsSum = 0;
    long partialsCalls = 0;
    bool evPending = false;
    bool usePdl = true;
    bool useTipsKernel = true;            // SB200_TIPS_KERNEL=0 (tuning aid): tips-only launches run the general kernel
    bool useTipTipTables = true;          // precomputed tables for operations on two tips (streaming sizes; SB200_TABLES)
    int forceVariant = -1;                // SB200_VARIANT (tuning aid): every launch of a small instance runs this thread shape
    int wideMinWarps = 0;                 // small instances: a level runs a wide thread shape from this many warps (of that shape) up,
    int wide2MinWarps = 0;                // ... the two-categories-per-thread shape from this many (of ITS warps) up,
    int narrowMinWarps = 0;               // ... the narrow one from this many up, the quad shape below

    size_t realSize() const { return fp32 ? 4 : 8; }
    ParamBlockHeader* hHeader() { return reinterpret_cast<ParamBlockHeader*>(hBlock); }
    DevModel* hModel(int i) { return reinterpret_cast<DevModel*>(hBlock + modelOff) + i; }
    int* hMidx() { return reinterpret_cast<int*>(hBlock + midxOff); }
    double* hElen() { return reinterpret_cast<double*>(hBlock + elenOff); }
    const double* dRates() const { return reinterpret_cast<const double*>(dBlock); }
    const DevModel* dModel(int i) const { return reinterpret_cast<const DevModel*>(dBlock + modelOff) + i; }
    const int* dMidx() const { return reinterpret_cast<const int*>(dBlock + midxOff); }
    const double* dElen() const { return reinterpret_cast<const double*>(dBlock + elenOff); }

    ~Engine() { destroy(); }

    void destroy() {
        DeviceGuard guard;
        guard.enter(device);
        if (stream) cudaStreamSynchronize(stream);
        if (ownStream && ownStream != stream) cudaStreamSynchronize(ownStream);
        cudaFree(dTips); cudaFree(dPartials); cudaFree(dMats); cudaFree(dPatWeights);
        for (double* p : dScale) cudaFree(p);
        dScale.clear();
        cudaFree(dSlots); cudaFree(dPlan); cudaFree(dTT);
        if (!arena) cudaFree(dBlock);
        freeParkedPlans();
        cudaFree(dBlockSums); cudaFree(dTicket); cudaFree(dOut); cudaFree(dTrace);
        if (hBlock && !arena) cudaFreeHost(hBlock);
        arena.reset();
        arenaSlot = -1;
        if (hSchedStage) cudaFreeHost(hSchedStage);
        if (hOut) cudaFreeHost(hOut);
        if (blockCopied) cudaEventDestroy(blockCopied);
        if (schedCopied) cudaEventDestroy(schedCopied);
        if (evP0) cudaEventDestroy(evP0);
        if (evP1) cudaEventDestroy(evP1);
        evP0 = evP1 = nullptr;
        if (ownStream) cudaStreamDestroy(ownStream);
        dTips = nullptr; dPartials = nullptr; dMats = nullptr; dPatWeights = nullptr; dSlots = nullptr; dTrace = nullptr;
        dBlock = nullptr; dPlan = nullptr; dOps = nullptr; dGroups = nullptr; dRootSlots = nullptr; dTTOps = nullptr; dTT = nullptr; dBlockSums = nullptr;
        dTicket = nullptr; dOut = nullptr; hBlock = nullptr; hSchedStage = nullptr; hOut = nullptr;
        blockCopied = nullptr; schedCopied = nullptr; ownStream = nullptr; stream = nullptr;
    }

    void layoutBlock(int cap) {
        edgeCap = cap;
        modelOff = sizeof(ParamBlockHeader);
        midxOff = modelOff + sizeof(DevModel) * static_cast<size_t>(nEigen);
        midxOff = (midxOff + 15) & ~size_t(15);
        elenOff = midxOff + sizeof(int) * static_cast<size_t>(cap);
        elenOff = (elenOff + 15) & ~size_t(15);
        blockBytes = elenOff + sizeof(double) * static_cast<size_t>(cap);
    }

    void init() {
        CK(cudaSetDevice(device));
        CK(cudaDeviceGetAttribute(&smCount, cudaDevAttrMultiProcessorCount, device));
        CK(cudaStreamCreateWithFlags(&ownStream, cudaStreamNonBlocking));
        stream = ownStream;
        CK(cudaEventCreateWithFlags(&blockCopied, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&schedCopied, cudaEventDisableTiming));
        // pattern axis padded to a whole number of thread-block tiles of every kernel that walks it: the pruning shapes
        // hold 32 .. 256 patterns per block, the root kernel 128
        streaming = static_cast<long long>(P) * C >= (1LL << 18);
        if (const char* f = std::getenv("SB200_FORCE_BIG")) streaming = (f[0] == '1');   // tuning aid
        {
            const int unit = bigTiles() ? 256 : 128;
            Ppad = ((P + unit - 1) / unit) * unit;
        }
        tipStride = Ppad;
        if (nTips > 0) {
            CK(cudaMalloc(&dTips, static_cast<size_t>(nTips) * tipStride));
            CK(cudaMemsetAsync(dTips, 4, static_cast<size_t>(nTips) * tipStride, stream));
        }
        tipSet.assign(nTips, 0);
        const size_t bufElems = static_cast<size_t>(Ppad) * C * 4;
        if (nPartials > 0) CK(cudaMalloc(&dPartials, bufElems * nPartials * realSize()));
        if (nMat > 0) {
            CK(cudaMalloc(&dMats, static_cast<size_t>(nMat) * C * kMatStride * realSize()));
            CK(cudaMemsetAsync(dMats, 0, static_cast<size_t>(nMat) * C * kMatStride * realSize(), stream));
        }
        CK(cudaMalloc(&dPatWeights, sizeof(double) * Ppad));
        CK(cudaMemsetAsync(dPatWeights, 0, sizeof(double) * Ppad, stream));
        dScale.assign(nScale, nullptr);
        scaleZeroPending.assign(nScale, 0);
        layoutBlock(nMat > 0 ? nMat : 1);
        CK(cudaMallocHost(&hBlock, blockBytes));
        std::memset(hBlock, 0, blockBytes);
        CK(cudaMalloc(&dBlock, blockBytes));
        for (int k = 0; k < kMaxCategories; ++k) hHeader()->rates[k] = 1.0;
        rootBlocks = Ppad / rootPatternsPerBlock();
        CK(cudaMalloc(&dBlockSums, sizeof(double) * rootBlocks));
        CK(cudaMalloc(&dTicket, sizeof(unsigned int)));
        CK(cudaMemsetAsync(dTicket, 0, sizeof(unsigned int), stream));
        CK(cudaMalloc(&dOut, sizeof(double)));
        CK(cudaMallocHost(&hOut, 2 * sizeof(double)));
        std::memset(hOut, 0, 2 * sizeof(double));
        pollResult = envInt("SB200_POLL", 1) != 0;
        usePdl = envInt("SB200_PDL", 1) != 0;                 // tuning aid: plain stream order
        useTipsKernel = envInt("SB200_TIPS_KERNEL", 1) != 0;
        forceVariant = envInt("SB200_VARIANT", -1);
        useTipTipTables = envInt("SB200_TABLES", bigTiles() ? 1 : 0) != 0;
        wideMinWarps = envInt("SB200_WIDE_MIN_WARPS", 64 * smCount);
        wide2MinWarps = envInt("SB200_WIDE2_MIN_WARPS", 16 * smCount);
        narrowMinWarps = envInt("SB200_NARROW_MIN_WARPS", smCount);
        CK(cudaStreamSynchronize(stream));
    }

    // The host image of the parameter block is about to be modified: wait until a copy that is still reading it has
    // finished, and remember that the device image is behind (the next kernel that reads it brings it up to date).
    void waitBlockFree() {
        blockDirty = true;
        if (blockInFlight) {
            CK(cudaEventSynchronize(arena ? arena->copied : blockCopied));
            blockInFlight = false;
        }
    }
    // the block goes back to buffers of its own (its content is kept)
    void leaveArena() {
        if (!arena) return;
        CK(cudaStreamSynchronize(stream));
        unsigned char* nh = nullptr;
        unsigned char* nd = nullptr;
        CK(cudaMallocHost(&nh, blockBytes));
        std::memcpy(nh, hBlock, blockBytes);
        const cudaError_t e = cudaMalloc(&nd, blockBytes);
        if (e != cudaSuccess) {
            cudaFreeHost(nh);
            throw CudaFail{e};
        }
        hBlock = nh;
        dBlock = nd;
        arena.reset();
        arenaSlot = -1;
        blockInFlight = false;
        blockDirty = true;
    }

    double* scaleBuffer(int idx) {
        if (!dScale[idx]) {
            CK(cudaMalloc(&dScale[idx], sizeof(double) * Ppad));
            CK(cudaMemsetAsync(dScale[idx], 0, sizeof(double) * Ppad, stream));
            scaleZeroPending[idx] = 0;
        }
        return dScale[idx];
    }

    void materializeZero(int idx) {
        double* b = scaleBuffer(idx);
        if (scaleZeroPending[idx]) {
            CK(cudaMemsetAsync(b, 0, sizeof(double) * Ppad, stream));
            scaleZeroPending[idx] = 0;
        }
    }

    // ---- (a) ---------------------------------------------------------------------------
    void uploadBlock() {
        CK(cudaMemcpyAsync(dBlock, hBlock, blockBytes, cudaMemcpyHostToDevice, stream));
        CK(cudaEventRecord(arena ? arena->copied : blockCopied, stream));
        blockInFlight = true;
        blockDirty = false;
    }

    MatJob matJob() const {
        MatJob j;
        j.model = dModel(0);
        j.rates = dRates();
        j.matrixIndex = dMidx();
        j.edgeLength = dElen();
        j.mats = dMats;
        j.count = edgeCount;
        j.pad = 0;
        return j;
    }

    int checkEdges(const int* idx, int count) const {
        for (int u = 0; u < count; ++u)
            if (idx[u] < 0 || idx[u] >= nMat) return BEAGLE_ERROR_OUT_OF_RANGE;
        return BEAGLE_SUCCESS;
    }

    int setEdges(const int* idx, const double* len, int count) {
        const int rc = checkEdges(idx, count);
        if (rc) return rc;
        waitBlockFree();
        if (count > edgeCap) {
            // grow the block, preserving the model part
            leaveArena();
            std::vector<unsigned char> keep(hBlock, hBlock + midxOff);
            CK(cudaStreamSynchronize(stream));
            cudaFreeHost(hBlock); cudaFree(dBlock);
            hBlock = nullptr; dBlock = nullptr;
            layoutBlock(count);
            CK(cudaMallocHost(&hBlock, blockBytes));
            std::memset(hBlock, 0, blockBytes);
            std::memcpy(hBlock, keep.data(), keep.size());
            CK(cudaMalloc(&dBlock, blockBytes));
        }
        std::memcpy(hMidx(), idx, sizeof(int) * count);
        std::memcpy(hElen(), len, sizeof(double) * count);
        edgeCount = count;
        return BEAGLE_SUCCESS;
    }

    // ---- (b) ---------------------------------------------------------------------------
    // Streaming sizes run one thread shape (kVarStream) on every launch; small instances choose per launch between the
    // wide and the narrow shape (kernels.cuh).  Fixed at creation: it sets the pattern padding.
    bool bigTiles() const { return streaming; }
    int blocksPerSm(int V, bool tips) const { return varBlocksPerSm(C, V, tips && useTipsKernel, static_cast<int>(realSize())); }

    // Shape and group count of a launch.  Every block pays a fixed set-up (operation words, first images and
    // operands: dependent trips to L2/HBM), so as few groups as possible -- but:
    //   * few pattern tiles (latency-bound sizes such as rbcl738): TWO waves of resident blocks, measured best
    //     between 1 and 4 waves: more groups shorten the longest sequential run of a launch, fewer groups amortise
    //     the set-up; a launch with enough independent chains to give every warp scheduler several warps runs the
    //     wide shape (all categories in one thread: a quarter of the instructions per pattern x category), the
    //     others, at the top of the tree, the narrow shape (a warp per category: a shorter step);
    //   * many tiles (streaming sizes): enough blocks for >= 8 waves, so the last partial wave costs little.
    LevelChoice chooseLevel(int chainCount, bool tipsCandidate) const {
        LevelChoice c;
        if (chainCount == 0) {                           // the planner's question about tip x tip tables (schedule.hpp)
            c.variant = 0;
            c.groups = useTipTipTables ? 1 : kNoTipTipTables;
            return c;
        }
        if (bigTiles()) {
            c.variant = kVarStream;
            const int tiles = Ppad / varPatterns(C, kVarStream, static_cast<int>(realSize()));
            int slots = smCount * blocksPerSm(kVarStream, tipsCandidate);
            slots = envInt("SB200_SLOTS", slots);
            if (tiles * 2 <= slots) c.groups = std::max(1, slots / tiles);
            else c.groups = std::max(1, (8 * slots + tiles - 1) / tiles);
            return c;
        }
        const int K = std::max(1, batchHint);
        const int rb = static_cast<int>(realSize());
        const long perChain = static_cast<long>(K) * chainCount * (Ppad / 32);                  // pattern sub-tiles x chains
        const long wideWarps = perChain * varNCW(C, kVarWide, rb), wide2Warps = perChain * varNCW(C, kVarWide2, rb);
        const long narrowWarps = perChain * C;
        c.variant = wideWarps >= wideMinWarps ? kVarWide : wide2Warps >= wide2MinWarps ? kVarWide2
                  : narrowWarps >= narrowMinWarps ? kVarNarrow : kVarQuad;
        if (forceVariant >= kVarWide && forceVariant < kVarCount) c.variant = forceVariant;
        if (c.variant == kVarQuad) {                   // every chain its own group: the launch has blocks to spare
            c.groups = std::max(1, chainCount);
            return c;
        }
        const int tiles = Ppad / varPatterns(C, c.variant, static_cast<int>(realSize()));
        const int waves = tipsCandidate && useTipsKernel ? 1 : 2;
        int slots = smCount * blocksPerSm(c.variant, tipsCandidate) * waves;
        slots = envInt("SB200_SLOTS", slots);
        if (tipsCandidate) slots = envInt("SB200_TIPS_SLOTS", slots);
        c.groups = std::max(1, slots / (tiles * K));
        return c;
    }

    // active plan <-> parked plan (pointer and vector swaps only)
    void swapPlan(ParkedPlan& p) {
        std::swap(sched, p.sched);
        std::swap(schedAccumulate, p.accumulate);
        std::swap(schedValid, p.valid);
        std::swap(schedHint, p.hint);
        std::swap(schedOps, p.key);
        std::swap(dOps, p.dOps);
        std::swap(dGroups, p.dGroups);
        std::swap(dRootSlots, p.dRootSlots);
        std::swap(dTTOps, p.dTTOps);
        std::swap(dTT, p.dTT);
        std::swap(dPlan, p.dPlan);
        std::swap(planCap, p.planCap);
        std::swap(ttCap, p.ttCap);
    }
    void forgetPlans() {
        schedValid = false;
        for (ParkedPlan& p : parked) p.valid = false;
    }
    void freeParkedPlans() {
        for (ParkedPlan& p : parked) {
            cudaFree(p.dPlan); cudaFree(p.dTT);
        }
        parked.clear();
    }
    void setBatchHint(int k) { batchHint = k < 1 ? 1 : k; }   // (plans are keyed by it: the launch shapes of small instances depend on it)

    // Incremental plans (keepSubtreeScalers): an internal operand that the list does not produce brings the subtree
    // scaler an earlier call left in its permanent slot -- which must exist.
    std::vector<char> producedScratch;
    int checkSlots(const Schedule& s) {
        if (slotValid.size() != static_cast<size_t>(nPartials)) slotValid.assign(nPartials, 0);
        producedScratch.assign(nPartials, 0);
        for (const DevOp& d : s.ops) producedScratch[d.dest] = 1;
        for (const DevOp& d : s.ops) {
            if (d.sa >= 0 && d.a >= nTips && !producedScratch[d.a - nTips] && !slotValid[d.a - nTips]) return BEAGLE_ERROR_OUT_OF_RANGE;
            if (d.sb >= 0 && d.b >= nTips && !producedScratch[d.b - nTips] && !slotValid[d.b - nTips]) return BEAGLE_ERROR_OUT_OF_RANGE;
        }
        return BEAGLE_SUCCESS;
    }

    int prepareSchedule(const BeagleOperation* ops, int count, bool accumulate) {
        // same operation list as last call (the common case: 4 of strom's 5 updaters do not touch the
        // topology): reuse the plan that is already in HBM.  One memcmp of 28 bytes per operation.
        const int hint = bigTiles() ? 1 : batchHint;
        auto same = [&](bool valid, bool acc, int h, const std::vector<BeagleOperation>& key) {
            return valid && acc == accumulate && h == hint && static_cast<int>(key.size()) == count &&
                   (count == 0 || std::memcmp(key.data(), ops, sizeof(BeagleOperation) * count) == 0);
        };
        if (same(schedValid, schedAccumulate, schedHint, schedOps)) return BEAGLE_SUCCESS;
        for (size_t i = parked.size(); i-- > 0;)
            if (same(parked[i].valid, parked[i].accumulate, parked[i].hint, parked[i].key)) {
                swapPlan(parked[i]);                                   // the plan that was active takes its place ...
                std::rotate(parked.begin() + i, parked.begin() + i + 1, parked.end());   // ... as the most recently used
                return BEAGLE_SUCCESS;
            }
        Schedule s;
        const LevelPolicy policy = [this](int chains, bool tips) { return chooseLevel(chains, tips); };
        int rc = buildSchedule(ops, count, nTips, nPartials, nMat, nScale, accumulate, policy, s, keepSubtreeScalers);
        if (rc != BEAGLE_SUCCESS) return rc;
        for (const DevOp& d : s.ops) {
            if (d.a >= 0 && d.a < nTips && !tipSet[d.a]) return BEAGLE_ERROR_OUT_OF_RANGE;
            if (d.b < nTips && !tipSet[d.b]) return BEAGLE_ERROR_OUT_OF_RANGE;
        }
        // miss: park the active plan; the new one is built over the least recently used set of device arrays
        if (schedValid) {
            if (parked.size() < kParkedPlans) parked.insert(parked.begin(), ParkedPlan());
            swapPlan(parked.front());
            std::rotate(parked.begin(), parked.begin() + 1, parked.end());
        }
        schedValid = false;
        sched = std::move(s);
        schedOps.assign(ops, ops + count);
        schedAccumulate = accumulate;
        schedHint = hint;
        // stage + upload
        const size_t opsB = sizeof(DevRec) * sched.ops.size();
        const size_t chB = sizeof(Group) * sched.groups.size();
        const size_t rtB = sizeof(int) * sched.rootSlots.size();
        std::vector<int> ttOps(sched.tipTipCount, 0);
        for (size_t i = 0; i < sched.ops.size(); ++i)
            if (sched.ops[i].flags & OP_TIPTIP) ttOps[sched.ops[i].tt] = static_cast<int>(i);
        const size_t ttB = sizeof(int) * ttOps.size();
        const size_t total = opsB + chB + rtB + ttB;
        if (schedInFlight) {
            CK(cudaEventSynchronize(schedCopied));
            schedInFlight = false;
        }
        if (total > schedStageCap) {
            if (hSchedStage) cudaFreeHost(hSchedStage);
            hSchedStage = nullptr;
            schedStageCap = total + total / 2 + 256;
            CK(cudaMallocHost(&hSchedStage, schedStageCap));
        }
        // device arrays may be in use by kernels still queued on the stream: growing them frees
        // memory, so drain first (rare: only when the operation count grows)
        const size_t ttNeed = static_cast<size_t>(sched.tipTipCount) * ttStride(C, static_cast<int>(realSize())) * realSize();
        if (total > planCap || ttNeed > ttCap) CK(cudaStreamSynchronize(stream));
        if (total > planCap) {
            cudaFree(dPlan);
            dPlan = nullptr;
            planCap = total + total / 2 + 256;
            CK(cudaMalloc(&dPlan, planCap));
        }
        // (64-byte operation records first: every array stays aligned to its element size)
        dOps = reinterpret_cast<DevRec*>(dPlan);
        dGroups = reinterpret_cast<Group*>(dPlan + opsB);
        dRootSlots = reinterpret_cast<int*>(dPlan + opsB + chB);
        dTTOps = reinterpret_cast<int*>(dPlan + opsB + chB + rtB);
        if (ttNeed > ttCap) {
            cudaFree(dTT);
            dTT = nullptr;
            ttCap = ttNeed + ttNeed / 2 + 1024;
            CK(cudaMalloc(&dTT, ttCap));
        }
        unsigned char* st = static_cast<unsigned char*>(hSchedStage);
        {
            // the planner's operations with every operand resolved to a byte offset (kernels.cuh: DevRec)
            DevRec* rec = reinterpret_cast<DevRec*>(st);
            const long long bufBytes = static_cast<long long>(Ppad) * C * 4 * static_cast<long long>(realSize());
            for (size_t i = 0; i < sched.ops.size(); ++i) {
                const DevOp& d = sched.ops[i];
                DevRec& r = rec[i];
                r.flags = d.flags; r.ma = d.ma; r.mb = d.mb; r.tt = d.tt;
                r.destOff = d.dest * bufBytes;
                r.bOff = (d.flags & OP_B_TIP) ? d.b * tipStride : (d.b - nTips) * bufBytes;
                r.aOff = !(d.flags & OP_CHAIN_START) ? 0 : (d.flags & OP_A_TIP) ? d.a * tipStride : (d.a - nTips) * bufBytes;
                r.sa = d.sa; r.sb = d.sb; r.sOut = d.sOut;
                r.pad[0] = r.pad[1] = r.pad[2] = 0;
            }
        }
        if (chB) std::memcpy(st + opsB, sched.groups.data(), chB);
        if (rtB) std::memcpy(st + opsB + chB, sched.rootSlots.data(), rtB);
        if (ttB) std::memcpy(st + opsB + chB + rtB, ttOps.data(), ttB);
        if (total) CK(cudaMemcpyAsync(dPlan, st, total, cudaMemcpyHostToDevice, stream));      // one copy for the whole plan
        CK(cudaEventRecord(schedCopied, stream));
        schedInFlight = true;
        const size_t needSlots = 2 * static_cast<size_t>(sched.slotCount) * Ppad;   // (log sum, product) planes
        if (needSlots > slotCapacity) {
            CK(cudaStreamSynchronize(stream));
            cudaFree(dSlots);
            dSlots = nullptr;
            slotCapacity = needSlots + needSlots / 4;
            CK(cudaMalloc(&dSlots, sizeof(double) * slotCapacity));
            slotValid.assign(slotValid.size(), 0);
        }
        // traffic bookkeeping (bytes), SURVEY.md section 8(d)
        const double b = static_cast<double>(realSize());
        const double PC = static_cast<double>(P) * C, Pd = static_cast<double>(P);
        algBytes = sched.nPartialPartial * (3 * 4 * b * PC) + sched.nTipPartial * (2 * 4 * b * PC + Pd) +
                   sched.nTipTip * (4 * b * PC + 2 * Pd);
        if (accumulate) algBytes += 16.0 * Pd * count;
        schedBytes = (sched.partialReads + sched.partialWrites) * 4 * b * PC + sched.tipReads * Pd + sched.scalerRows * 8.0 * Pd;
        schedValid = true;
        return BEAGLE_SUCCESS;
    }

    // Arguments of the pruning launches of the active plan (and the bookkeeping of the cumulative buffer's reset).
    template <typename real>
    PruneArgs<real> pruneArgs(int cumIdx) {
        double* cum = nullptr;
        int cumMode = 0;
        if (cumIdx != BEAGLE_OP_NONE) {
            cum = scaleBuffer(cumIdx);
            const bool singleRootAdd = !sched.sequential && sched.rootSlots.empty() && !sched.chains.empty();
            if (scaleZeroPending[cumIdx]) {
                if (singleRootAdd) {
                    cumMode = 1;                       // the root chain overwrites: reset fused away
                    scaleZeroPending[cumIdx] = 0;
                } else {
                    materializeZero(cumIdx);
                }
            }
        }
        PruneArgs<real> a;
        a.partials = static_cast<real*>(dPartials);
        a.tips = dTips;
        a.mats = static_cast<const real*>(dMats);
        a.slotAcc = dSlots;
        a.slotProd = dSlots + static_cast<size_t>(sched.slotCount) * Ppad;
        a.cum = cum;
        a.ttTables = static_cast<const real*>(dTT);
        a.trace = dTrace;
        a.ops = dOps;
        a.groups = dGroups;
        a.tipStride = tipStride;
        a.Ppad = Ppad;
        a.nTips = nTips;
        a.cumMode = cumMode;
        // (incremental lists read buffers of earlier calls, and lists that are not forests run one operation per launch:
        // both need every buffer in memory)
        a.storeAll = (transientPartials && !keepSubtreeScalers && !sched.sequential) ? 0 : 1;
        a.arenaBytes = static_cast<long long>(Ppad) * C * 4 * static_cast<long long>(realSize()) * nPartials;
        a.tipBytes = static_cast<long long>(nTips) * tipStride;
        a.nMat = nMat;
        a.nTables = sched.tipTipCount;
        a.nSlots = sched.slotCount;
        a.nOps = static_cast<int>(sched.ops.size());
        return a;
    }

    // permanent subtree-scaler slots written / invalidated by a launch of the active plan
    void noteSlotsAfterPartials(bool accumulate) {
        if (!keepSubtreeScalers) return;
        if (slotValid.size() != static_cast<size_t>(nPartials)) slotValid.assign(nPartials, 0);
        for (const DevOp& d : sched.ops) slotValid[d.dest] = (accumulate && !sched.sequential) ? 1 : 0;
    }

    void collectProfile() {
        if (!evPending) return;
        float ms = 0.f;
        CK(cudaEventSynchronize(evP1));
        CK(cudaEventElapsedTime(&ms, evP0, evP1));
        partialsMsSum += ms;
        ++partialsCalls;
        evPending = false;
    }

    // ---- (c) ---------------------------------------------------------------------------
    int checkRoot(int parent, int child, int matrix, int wIdx, int fIdx, int cumIdx) const {
        const int nbuf = nTips + nPartials;
        if (parent < nTips || parent >= nbuf || child < 0 || child >= nbuf || matrix < 0 || matrix >= nMat)
            return BEAGLE_ERROR_OUT_OF_RANGE;
        if (child < nTips && !tipSet[child]) return BEAGLE_ERROR_OUT_OF_RANGE;
        if (wIdx < 0 || wIdx >= nEigen || fIdx < 0 || fIdx >= nEigen) return BEAGLE_ERROR_OUT_OF_RANGE;
        if (cumIdx != BEAGLE_OP_NONE && (cumIdx < 0 || cumIdx >= nScale)) return BEAGLE_ERROR_OUT_OF_RANGE;
        return BEAGLE_SUCCESS;
    }

    template <typename real>
    RootArgs<real> rootArgs(int parent, int child, int matrix, int wIdx, int fIdx, int cumIdx, double* out) {
        lastParent = parent; lastChild = child; lastMatrix = matrix; lastWIdx = wIdx; lastFIdx = fIdx; lastRootCum = cumIdx;
        RootArgs<real> a;
        a.partials = static_cast<const real*>(dPartials);
        a.tips = dTips;
        a.mats = static_cast<const real*>(dMats);
        a.model = dModel(fIdx);
        a.catWeights = dModel(wIdx)->weights;
        a.patWeights = dPatWeights;
        a.cum = nullptr;
        if (cumIdx != BEAGLE_OP_NONE) {
            materializeZero(cumIdx);
            a.cum = dScale[cumIdx];
        }
        a.blockSums = dBlockSums;
        a.ticket = dTicket;
        a.out = out;
        a.doneFlag = nullptr;
        a.doneSeq = 0;
        awaitedSeq = 0;
        if (out == hOut) {
            a.doneFlag = reinterpret_cast<unsigned long long*>(hOut + 1);
            a.doneSeq = awaitedSeq = ++rootSeq;
        }
        a.tipStride = tipStride;
        a.P = P;
        a.Ppad = Ppad;
        a.nTips = nTips;
        a.parent = parent;
        a.child = child;
        a.matrix = matrix;
        a.blocks = rootBlocks;
        a.pad = 0;
        return a;
    }

    // The root kernel wrote the scalar straight into pinned host memory (hOut is device-accessible through
    // UVA): the 8-byte device->host transfer is part of the kernel, only the stream wait remains.
    int readScalar(double* out) {
        // The last block of the root kernel stores the scalar and then a sequence word in pinned host memory: polling
        // that word returns a few microseconds before the stream is reported idle (kernel completion and the
        // driver's own bookkeeping come after the store).  Everything enqueued earlier is complete by stream order.
        bool seen = false;
        if (pollResult && awaitedSeq != 0) {
            const volatile unsigned long long* flag = reinterpret_cast<const volatile unsigned long long*>(hOut + 1);
            for (unsigned spins = 1;; ++spins) {
                if (*flag == awaitedSeq) { seen = true; break; }
                if ((spins & 0x3fff) == 0) {                       // a failed launch or kernel never writes: ask the driver now and then
                    const cudaError_t q = cudaStreamQuery(stream);
                    if (q != cudaErrorNotReady) break;             // idle (flag is there) or an error (reported below)
                }
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            std::atomic_thread_fence(std::memory_order_acquire);
        }
        if (!seen) CK(cudaStreamSynchronize(stream));
        awaitedSeq = 0;
        blockInFlight = false;
        schedInFlight = false;
        *out = *reinterpret_cast<const volatile double*>(hOut);
        return (*out != *out) ? BEAGLE_ERROR_FLOATING_POINT : BEAGLE_SUCCESS;
    }
};

// ---------------------------------------------------------------------------------------------
// Launch sets: K instances side by side (blockIdx.z / .y = instance).  All instances of a set live on one device and
// run on the first one's stream; they share the category count and the precision.
// ---------------------------------------------------------------------------------------------
static void launchMatricesSet(Engine* const* es, int K) {
    Engine& L = *es[0];
    MatBatch mb;
    int maxCount = 0;
    for (int z = 0; z < kMaxBatch; ++z) {
        mb.j[z] = es[z < K ? z : 0]->matJob();
        if (z >= K) mb.j[z].count = 0;
        maxCount = std::max(maxCount, mb.j[z].count);
    }
    if (maxCount == 0) return;
    const long long threads = static_cast<long long>(maxCount) * L.C * 16;
    const dim3 grid(static_cast<unsigned>((threads + 255) / 256), K);
    SB200_RANGE_PUSH("sb200 A: transition matrices");
    if (L.fp32) launchK(L.stream, false, transition_matrices_kernel<float>, grid, dim3(256), mb, L.C);
    else launchK(L.stream, false, transition_matrices_kernel<double>, grid, dim3(256), mb, L.C);
    SB200_RANGE_POP();
    CK(cudaGetLastError());
    ++L.launches;
}

template <typename real, int CC, int V>
static void launchLevelV(Engine& L, const PruneBatch<real>& pb, dim3 grid, bool tips) {
    if (tips) launchK(L.stream, L.usePdl, prune_chains_kernel<real, CC, V, true>, grid, dim3(varThreads(CC, V, sizeof(real))), pb);
    else launchK(L.stream, L.usePdl, prune_chains_kernel<real, CC, V, false>, grid, dim3(varThreads(CC, V, sizeof(real))), pb);
}
template <typename real, int CC>
static void launchLevelC(Engine& L, const PruneBatch<real>& pb, int tilesOf[kVarCount], int maxGroups, int K, int V, bool tips) {
    const dim3 grid(tilesOf[V], maxGroups, K);
    if (V == kVarWide2) launchLevelV<real, CC, kVarWide2>(L, pb, grid, tips);
    else if (V == kVarQuad) launchK(L.stream, L.usePdl, prune_quad_kernel<real, CC>, grid, dim3(32 * CC), pb);
    else if (V == kVarStream) launchLevelV<real, CC, kVarStream>(L, pb, grid, tips);
    else if (V == kVarWide) launchLevelV<real, CC, kVarWide>(L, pb, grid, tips);
    else launchLevelV<real, CC, kVarNarrow>(L, pb, grid, tips);
}

template <typename real>
static void launchPartialsSetT(Engine* const* es, int K, const int* cumIdx) {
    Engine& L = *es[0];
    const int C = L.C;
    PruneArgs<real> args[kMaxBatch];
    TableBatch<real> tb;
    int maxTT = 0;
    size_t maxLevels = 0;
    for (int z = 0; z < K; ++z) {
        Engine& E = *es[z];
        args[z] = E.template pruneArgs<real>(cumIdx[z]);
        maxLevels = std::max(maxLevels, E.sched.levels.size());
    }
    for (int z = 0; z < kMaxBatch; ++z) {
        Engine& E = *es[z < K ? z : 0];
        tb.j[z].ops = E.dOps;
        tb.j[z].ttOps = E.dTTOps;
        tb.j[z].mats = static_cast<const real*>(E.dMats);
        tb.j[z].tables = static_cast<real*>(E.dTT);
        tb.j[z].count = z < K ? E.sched.tipTipCount : 0;
        tb.j[z].pad = 0;
        maxTT = std::max(maxTT, tb.j[z].count);
    }
    SB200_RANGE_PUSH("sb200 B: partials");
    if (maxTT > 0) {
        launchK(L.stream, L.usePdl, tiptip_tables_kernel<real>, dim3(maxTT, K), dim3(128), tb, C);
        CK(cudaGetLastError());
        ++L.launches;
    }
    int tilesOf[kVarCount];
    for (int v = 0; v < kVarCount; ++v) tilesOf[v] = L.Ppad / varPatterns(C, v, sizeof(real));
    tilesOf[kVarQuad] = L.Ppad / kQuadPatterns;
    static const int traceLevel = envInt("SB200_TRACE_LEVEL", -1);   // tuning aid (needs -DSB200_TRACE)
    for (size_t l = 0; l < maxLevels; ++l) {
        // the instances' launches number l, grouped by (thread shape, tips-only): normally ONE launch
        bool done[kMaxBatch] = {false};
        for (int z0 = 0; z0 < K; ++z0) {
            if (done[z0]) continue;
            if (l >= es[z0]->sched.levels.size() || es[z0]->sched.levels[l].groupCount == 0) { done[z0] = true; continue; }
            const Level& L0 = es[z0]->sched.levels[l];
            const int V = L0.variant;
            const bool tips = L0.tipsOnly != 0 && L.useTipsKernel;
            PruneBatch<real> pb;
            int maxGroups = 0, n = 0;
            for (int z = z0; z < K; ++z) {
                if (done[z] || l >= es[z]->sched.levels.size()) continue;
                const Level& Lz = es[z]->sched.levels[l];
                if (Lz.groupCount == 0) { done[z] = true; continue; }
                if (Lz.variant != V || (V != kVarQuad && (Lz.tipsOnly != 0 && L.useTipsKernel) != tips)) continue;
                pb.a[n] = args[z];
                pb.a[n].trace = (traceLevel < 0 || traceLevel == static_cast<int>(l)) ? args[z].trace : nullptr;
                pb.groupBase[n] = Lz.groupStart;
                pb.groupCount[n] = Lz.groupCount;
                maxGroups = std::max(maxGroups, Lz.groupCount);
                done[z] = true;
                ++n;
            }
            for (int q = n; q < kMaxBatch; ++q) { pb.a[q] = pb.a[0]; pb.groupBase[q] = 0; pb.groupCount[q] = 0; }
            switch (C) {
                case 1: launchLevelC<real, 1>(L, pb, tilesOf, maxGroups, n, V, tips); break;
                case 2: launchLevelC<real, 2>(L, pb, tilesOf, maxGroups, n, V, tips); break;
                case 3: launchLevelC<real, 3>(L, pb, tilesOf, maxGroups, n, V, tips); break;
                case 4: launchLevelC<real, 4>(L, pb, tilesOf, maxGroups, n, V, tips); break;
                case 5: launchLevelC<real, 5>(L, pb, tilesOf, maxGroups, n, V, tips); break;
                case 6: launchLevelC<real, 6>(L, pb, tilesOf, maxGroups, n, V, tips); break;
                case 7: launchLevelC<real, 7>(L, pb, tilesOf, maxGroups, n, V, tips); break;
                case 8: launchLevelC<real, 8>(L, pb, tilesOf, maxGroups, n, V, tips); break;
            }
            CK(cudaGetLastError());
            ++L.launches;
        }
    }
    for (int z = 0; z < K; ++z) {
        Engine& E = *es[z];
        if (args[z].cum && !E.sched.rootSlots.empty()) {
            const int nr = static_cast<int>(E.sched.rootSlots.size());
            int mode = 0;
            if (E.scaleZeroPending[cumIdx[z]]) { mode = 1; E.scaleZeroPending[cumIdx[z]] = 0; }
            launchK(L.stream, L.usePdl, fold_roots_kernel, dim3((E.Ppad + 255) / 256), dim3(256), static_cast<const double*>(args[z].slotAcc),
                    static_cast<const double*>(args[z].slotProd), static_cast<const int*>(E.dRootSlots), nr, args[z].cum, E.Ppad, mode);
            CK(cudaGetLastError());
            ++L.launches;
        }
        E.noteSlotsAfterPartials(E.schedAccumulate);
    }
    SB200_RANGE_POP();
}

static void launchPartialsSet(Engine* const* es, int K, const int* cumIdx) {
    Engine& L = *es[0];
    for (int z = 0; z < K; ++z) es[z]->lastCumIndex = cumIdx[z];
    if (L.profiling) {
        L.collectProfile();
        if (!L.evP0) {
            CK(cudaEventCreate(&L.evP0));
            CK(cudaEventCreate(&L.evP1));
        }
        CK(cudaEventRecord(L.evP0, L.stream));
    }
    if (L.fp32) launchPartialsSetT<float>(es, K, cumIdx);
    else launchPartialsSetT<double>(es, K, cumIdx);
    if (L.profiling) {
        CK(cudaEventRecord(L.evP1, L.stream));
        L.evPending = true;
    }
}

struct RootCall {
    int parent, child, matrix, wIdx, fIdx, cumIdx;
    double* out;
};

template <typename real>
static void launchRootSetT(Engine* const* es, int K, const RootCall* rc) {
    Engine& L = *es[0];
    RootBatch<real> rb;
    int maxBlocks = 0;
    for (int z = 0; z < K; ++z) {
        rb.a[z] = es[z]->template rootArgs<real>(rc[z].parent, rc[z].child, rc[z].matrix, rc[z].wIdx, rc[z].fIdx, rc[z].cumIdx, rc[z].out);
        maxBlocks = std::max(maxBlocks, rb.a[z].blocks);
    }
    for (int z = K; z < kMaxBatch; ++z) { rb.a[z] = rb.a[0]; rb.a[z].blocks = 0; }
    const dim3 grid(maxBlocks, K), block(rootPatternsPerBlock());
    SB200_RANGE_PUSH("sb200 C: root edge log-likelihood");
    switch (L.C) {
        case 1: launchK(L.stream, L.usePdl, root_loglik_kernel<real, 1>, grid, block, rb); break;
        case 2: launchK(L.stream, L.usePdl, root_loglik_kernel<real, 2>, grid, block, rb); break;
        case 3: launchK(L.stream, L.usePdl, root_loglik_kernel<real, 3>, grid, block, rb); break;
        case 4: launchK(L.stream, L.usePdl, root_loglik_kernel<real, 4>, grid, block, rb); break;
        case 5: launchK(L.stream, L.usePdl, root_loglik_kernel<real, 5>, grid, block, rb); break;
        case 6: launchK(L.stream, L.usePdl, root_loglik_kernel<real, 6>, grid, block, rb); break;
        case 7: launchK(L.stream, L.usePdl, root_loglik_kernel<real, 7>, grid, block, rb); break;
        case 8: launchK(L.stream, L.usePdl, root_loglik_kernel<real, 8>, grid, block, rb); break;
    }
    SB200_RANGE_POP();
    CK(cudaGetLastError());
    ++L.launches;
}

// Validates, then launches.  Frequencies / category weights set since the last matrix update take effect now.
static int launchRootSet(Engine* const* es, int K, const RootCall* rc) {
    for (int z = 0; z < K; ++z) {
        const int code = es[z]->checkRoot(rc[z].parent, rc[z].child, rc[z].matrix, rc[z].wIdx, rc[z].fIdx, rc[z].cumIdx);
        if (code) return code;
    }
    for (int z = 0; z < K; ++z)
        if (es[z]->blockDirty) es[z]->uploadBlock();
    if (es[0]->fp32) launchRootSetT<float>(es, K, rc);
    else launchRootSetT<double>(es, K, rc);
    return BEAGLE_SUCCESS;
}

// ---------------------------------------------------------------------------------------------
// instance table
// ---------------------------------------------------------------------------------------------
static std::mutex g_mu;
static std::vector<std::unique_ptr<Engine>> g_inst;

static Engine* getEngine(int id) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (id < 0 || id >= static_cast<int>(g_inst.size()) || !g_inst[id]) return nullptr;
    return g_inst[id].get();
}

static int mapCuda(cudaError_t e) {
    if (e == cudaErrorMemoryAllocation) return BEAGLE_ERROR_OUT_OF_MEMORY;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice)
        return BEAGLE_ERROR_NO_RESOURCE;
    std::fprintf(stderr, "strom_b200: CUDA error %d (%s)\n", static_cast<int>(e), cudaGetErrorString(e));
    return BEAGLE_ERROR_GENERAL;
}

// Runs `body` with `device` current (the caller's current device is restored afterwards); converts every failure into a code.
template <typename F>
static int onDevice(int device, F&& body) {
    try {
        DeviceGuard guard;
        cudaError_t e = guard.enter(device);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return mapCuda(e);
        }
        return body();
    } catch (const CudaFail& f) {
        cudaGetLastError();
        return mapCuda(f.err);
    } catch (const std::bad_alloc&) {
        return BEAGLE_ERROR_OUT_OF_MEMORY;
    } catch (...) {
        return BEAGLE_ERROR_UNIDENTIFIED_EXCEPTION;
    }
}
template <typename F>
static int withEngine(int id, F&& body) {
    Engine* E = getEngine(id);
    if (!E) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
    return onDevice(E->device, [&]() { return body(*E); });
}

static void fillCmat(const double* evec, const double* ivec, double* cmat) {
    int l = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            for (int x = 0; x < 4; ++x) cmat[l++] = evec[i * 4 + x] * ivec[x * 4 + j];
}

// Puts the parameter blocks of the set's instances into slots 0..K-1 of one arena (the first instance's), unless they
// already are.  Done once per set (it drains the stream); false if the blocks cannot share an arena.
static bool shareArena(Engine* const* es, int K, const StromEvalArgs* args) {
    Engine& L = *es[0];
    for (int z = 0; z < K; ++z)
        if (es[z]->blockBytes != L.blockBytes || args[z].edgeCount > es[z]->edgeCap) return false;
    bool ok = L.arena && L.arenaSlot == 0 && L.arena->slots >= K;
    for (int z = 1; z < K && ok; ++z) ok = es[z]->arena == L.arena && es[z]->arenaSlot == z;
    if (ok) return true;
    CK(cudaStreamSynchronize(L.stream));
    std::shared_ptr<ParamArena> A(new ParamArena());
    A->device = L.device;
    A->slotBytes = (L.blockBytes + 255) & ~size_t(255);
    A->slots = K;
    CK(cudaMallocHost(&A->h, A->slotBytes * K));
    CK(cudaMalloc(&A->d, A->slotBytes * K));
    CK(cudaEventCreateWithFlags(&A->copied, cudaEventDisableTiming));
    for (int z = 0; z < K; ++z) {
        Engine& E = *es[z];
        unsigned char* nh = A->h + A->slotBytes * z;
        std::memcpy(nh, E.hBlock, E.blockBytes);
        if (!E.arena) {
            cudaFreeHost(E.hBlock);
            cudaFree(E.dBlock);
        }
        E.hBlock = nh;
        E.dBlock = A->d + A->slotBytes * z;
        E.arena = A;
        E.arenaSlot = z;
        E.blockInFlight = false;
        E.blockDirty = true;
    }
    return true;
}

// One whole evaluation of every instance of the set: parameter blocks up, then ONE set of launches.  Everything is
// validated before any instance's parameter block is touched or anything is enqueued.
static int evalSet(Engine* const* es, int K, const StromEvalArgs* args, double* const* deviceOut) {
    Engine& L = *es[0];
    for (int z = 0; z < K; ++z) {
        Engine& E = *es[z];
        const StromEvalArgs& a = args[z];
        if (a.edgeCount < 0 || a.operationCount < 0) return BEAGLE_ERROR_OUT_OF_RANGE;
        if (!a.stateFrequencies || !a.eigenVectors || !a.inverseEigenVectors || !a.eigenValues || !a.categoryRates || !a.categoryWeights)
            return BEAGLE_ERROR_OUT_OF_RANGE;
        if ((a.edgeCount > 0 && (!a.matrixIndices || !a.edgeLengths)) || (a.operationCount > 0 && !a.operations)) return BEAGLE_ERROR_OUT_OF_RANGE;
        if (a.categoryCount != 0 && a.categoryCount != E.C) return BEAGLE_ERROR_OUT_OF_RANGE;
        if (E.nScale < 1) return BEAGLE_ERROR_OUT_OF_RANGE;
        if (E.device != L.device || E.C != L.C || E.fp32 != L.fp32) return BEAGLE_ERROR_OUT_OF_RANGE;
        int rc = E.checkEdges(a.matrixIndices, a.edgeCount);
        if (rc) return rc;
        rc = E.checkRoot(a.rootParentBuffer, a.rootChildBuffer, a.rootMatrix, 0, 0, 0);
        if (rc) return rc;
    }
    // the set runs on the first instance's stream: an instance that joins drains its own stream once and stays
    for (int z = 1; z < K; ++z) {
        Engine& E = *es[z];
        if (E.stream != L.stream) {
            CK(cudaStreamSynchronize(E.stream));
            E.stream = L.stream;
        }
    }
    for (int z = 0; z < K; ++z) es[z]->setBatchHint(K);
    // plans first (host work; a new topology also enqueues its plan upload), so that a rejected list leaves no trace
    for (int z = 0; z < K; ++z) {
        const int rc = es[z]->prepareSchedule(args[z].operations, args[z].operationCount, true);
        if (rc) return rc;
        if (es[z]->keepSubtreeScalers) {
            const int sc = es[z]->checkSlots(es[z]->sched);
            if (sc) return sc;
        }
    }
    static const bool useArena = envInt("SB200_ARENA", 1) != 0;          // tuning aid: one parameter copy per instance
    const bool oneCopy = K > 1 && useArena && shareArena(es, K, args);
    for (int z = 0; z < K; ++z) {
        Engine& E = *es[z];
        const StromEvalArgs& a = args[z];
        const int rc = E.setEdges(a.matrixIndices, a.edgeLengths, a.edgeCount);
        if (rc) return rc;
        DevModel* m = E.hModel(0);
        std::memcpy(m->freqs, a.stateFrequencies, sizeof(double) * 4);
        fillCmat(a.eigenVectors, a.inverseEigenVectors, m->cmat);
        std::memcpy(m->evals, a.eigenValues, sizeof(double) * 4);
        std::memcpy(m->weights, a.categoryWeights, sizeof(double) * E.C);
        std::memcpy(E.hHeader()->rates, a.categoryRates, sizeof(double) * E.C);
        if (!oneCopy) E.uploadBlock();
    }
    if (oneCopy) {
        // the K parameter blocks sit back to back in the set's arena: one copy for all of them
        ParamArena& A = *L.arena;
        CK(cudaMemcpyAsync(A.d, A.h, A.slotBytes * K, cudaMemcpyHostToDevice, L.stream));
        CK(cudaEventRecord(A.copied, L.stream));
        for (int z = 0; z < K; ++z) {
            es[z]->blockInFlight = true;
            es[z]->blockDirty = false;
        }
    }
    launchMatricesSet(es, K);
    int cumIdx[kMaxBatch];
    RootCall rc[kMaxBatch];
    for (int z = 0; z < K; ++z) {
        es[z]->scaleBuffer(0);
        es[z]->scaleZeroPending[0] = 1;
        cumIdx[z] = 0;
        rc[z] = RootCall{args[z].rootParentBuffer, args[z].rootChildBuffer, args[z].rootMatrix, 0, 0, 0,
                         (deviceOut && deviceOut[z]) ? deviceOut[z] : es[z]->hOut};
    }
    launchPartialsSet(es, K, cumIdx);
    return launchRootSet(es, K, rc);
}

}  // namespace sb200

using namespace sb200;

// =============================================================================================
// The 13 BeagleLib-compatible entry points
// =============================================================================================
static char g_resName[256] = "CUDA device";
static char g_resDesc[256] = "strom_b200 sm_100a engine";
static BeagleResource g_res[16];
static BeagleResourceList g_resList = {g_res, 0};
static char g_resNames[16][256];

extern "C" BeagleResourceList* beagleGetResourceList(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    if (n > 16) n = 16;
    for (int i = 0; i < n; ++i) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, i) == cudaSuccess)
            std::snprintf(g_resNames[i], sizeof(g_resNames[i]), "%s", prop.name);
        else
            std::snprintf(g_resNames[i], sizeof(g_resNames[i]), "%.200s %d", g_resName, i);
        g_res[i].name = g_resNames[i];
        g_res[i].description = g_resDesc;
        g_res[i].supportFlags = BEAGLE_FLAG_PRECISION_DOUBLE | BEAGLE_FLAG_PRECISION_SINGLE | BEAGLE_FLAG_SCALING_MANUAL |
                                BEAGLE_FLAG_PROCESSOR_GPU | BEAGLE_FLAG_SCALERS_LOG | BEAGLE_FLAG_EIGEN_REAL;
        g_res[i].requiredFlags = 0;
    }
    g_resList.length = n;
    return &g_resList;
}

extern "C" int beagleCreateInstance(int tipCount, int partialsBufferCount, int compactBufferCount, int stateCount,
                                    int patternCount, int eigenBufferCount, int matrixBufferCount, int categoryCount,
                                    int scaleBufferCount, int* resourceList, int resourceCount, long preferenceFlags,
                                    long requirementFlags, BeagleInstanceDetails* returnInfo) {
    (void)preferenceFlags;   // the reference *prefers* PROCESSOR_CPU (src/likelihood.hpp:193); a preference only
    if (stateCount != 4) return BEAGLE_ERROR_NO_IMPLEMENTATION;
    if (requirementFlags & (BEAGLE_FLAG_PROCESSOR_CPU | BEAGLE_FLAG_SCALING_AUTO | BEAGLE_FLAG_SCALING_ALWAYS |
                            BEAGLE_FLAG_EIGEN_COMPLEX | BEAGLE_FLAG_SCALERS_RAW | BEAGLE_FLAG_VECTOR_SSE |
                            BEAGLE_FLAG_THREADING_OPENMP))
        return BEAGLE_ERROR_NO_IMPLEMENTATION;
    if ((requirementFlags & BEAGLE_FLAG_PRECISION_SINGLE) && (requirementFlags & BEAGLE_FLAG_PRECISION_DOUBLE))
        return BEAGLE_ERROR_NO_IMPLEMENTATION;
    if (tipCount < 0 || partialsBufferCount < 0 || compactBufferCount < 0 || patternCount < 1 || eigenBufferCount < 1 ||
        matrixBufferCount < 0 || categoryCount < 1 || scaleBufferCount < 0)
        return BEAGLE_ERROR_OUT_OF_RANGE;
    if (categoryCount > kMaxCategories) return BEAGLE_ERROR_NO_IMPLEMENTATION;
    if (compactBufferCount < tipCount) return BEAGLE_ERROR_NO_IMPLEMENTATION;   // tips are compact states only
    int device = t_device >= 0 ? t_device : g_defaultDevice.load();
    if (resourceList && resourceCount > 0) device = resourceList[0];
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return BEAGLE_ERROR_NO_RESOURCE;
    }
    if (device < 0 || device >= ndev) return BEAGLE_ERROR_NO_RESOURCE;
    std::unique_ptr<Engine> E(new (std::nothrow) Engine());
    if (!E) return BEAGLE_ERROR_OUT_OF_MEMORY;
    E->device = device;
    E->fp32 = (requirementFlags & BEAGLE_FLAG_PRECISION_SINGLE) != 0;
    E->nTips = tipCount; E->nPartials = partialsBufferCount; E->nCompact = compactBufferCount; E->P = patternCount;
    E->nEigen = eigenBufferCount; E->nMat = matrixBufferCount; E->C = categoryCount; E->nScale = scaleBufferCount;
    {
        Engine* raw = E.get();
        const int rc = onDevice(device, [&]() {
            raw->init();
            return static_cast<int>(BEAGLE_SUCCESS);
        });
        if (rc != BEAGLE_SUCCESS) return rc;
    }
    if (returnInfo) {
        returnInfo->resourceNumber = device;
        returnInfo->resourceName = g_resName;
        returnInfo->implName = const_cast<char*>(E->fp32 ? "strom_b200-CUDA-sm100a-4State-Single" : "strom_b200-CUDA-sm100a-4State-Double");
        returnInfo->implDescription = g_resDesc;
        returnInfo->flags = (E->fp32 ? BEAGLE_FLAG_PRECISION_SINGLE : BEAGLE_FLAG_PRECISION_DOUBLE) | BEAGLE_FLAG_SCALING_MANUAL |
                            BEAGLE_FLAG_PROCESSOR_GPU | BEAGLE_FLAG_SCALERS_LOG | BEAGLE_FLAG_EIGEN_REAL;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    for (size_t i = 0; i < g_inst.size(); ++i)
        if (!g_inst[i]) {
            g_inst[i] = std::move(E);
            return static_cast<int>(i);
        }
    g_inst.push_back(std::move(E));
    return static_cast<int>(g_inst.size()) - 1;
}

extern "C" int beagleFinalizeInstance(int instance) {
    std::unique_ptr<Engine> E;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (instance < 0 || instance >= static_cast<int>(g_inst.size()) || !g_inst[instance])
            return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
        E = std::move(g_inst[instance]);
        for (auto& other : g_inst)                        // instances that joined this one's stream (launch sets) go back to their own
            if (other && other->stream == E->ownStream) {
                cudaStreamSynchronize(other->stream);
                other->stream = other->ownStream;
            }
    }
    E.reset();
    return BEAGLE_SUCCESS;
}

static int setTipsCommon(Engine& E, int tipIndex, const int* s32, const unsigned char* s8) {
    if (tipIndex < 0 || tipIndex >= E.nTips || tipIndex >= E.nCompact) return BEAGLE_ERROR_OUT_OF_RANGE;
    // (pageable source: cudaMemcpyAsync returns once the bytes sit in the driver's staging memory, so the row buffer can
    // be reused at once and no stream wait is needed per tip; later work on the stream is ordered behind the copy)
    std::vector<unsigned char>& codes = E.tipStage;
    codes.resize(static_cast<size_t>(E.P));
    if (s32)
        for (int j = 0; j < E.P; ++j) codes[j] = (s32[j] >= 0 && s32[j] < 4) ? static_cast<unsigned char>(s32[j]) : 4;
    else
        for (int j = 0; j < E.P; ++j) codes[j] = s8[j] < 4 ? s8[j] : 4;
    CK(cudaMemcpyAsync(E.dTips + static_cast<size_t>(tipIndex) * E.tipStride, codes.data(), static_cast<size_t>(E.P),
                       cudaMemcpyHostToDevice, E.stream));
    E.tipSet[tipIndex] = 1;
    E.forgetPlans();
    return BEAGLE_SUCCESS;
}

extern "C" int beagleSetTipStates(int instance, int tipIndex, const int* inStates) {
    if (!inStates) return BEAGLE_ERROR_OUT_OF_RANGE;
    return withEngine(instance, [&](Engine& E) { return setTipsCommon(E, tipIndex, inStates, nullptr); });
}

extern "C" int sb200SetTipStatesU8(int instance, int tipIndex, const unsigned char* inStates) {
    if (!inStates) return BEAGLE_ERROR_OUT_OF_RANGE;
    return withEngine(instance, [&](Engine& E) { return setTipsCommon(E, tipIndex, nullptr, inStates); });
}

// All tips in one call: codes[tip * patternCount + pattern], one strided copy (row pitch = the padded pattern axis).
extern "C" int sb200SetAllTipStatesU8(int instance, const unsigned char* inStates) {
    if (!inStates) return BEAGLE_ERROR_OUT_OF_RANGE;
    return withEngine(instance, [&](Engine& E) {
        if (E.nTips == 0) return static_cast<int>(BEAGLE_SUCCESS);
        if (E.nCompact < E.nTips) return static_cast<int>(BEAGLE_ERROR_OUT_OF_RANGE);
        std::vector<unsigned char>& codes = E.tipStage;
        const size_t n = static_cast<size_t>(E.nTips) * E.P;
        codes.resize(n);
        for (size_t q = 0; q < n; ++q) codes[q] = inStates[q] < 4 ? inStates[q] : 4;
        CK(cudaMemcpy2DAsync(E.dTips, static_cast<size_t>(E.tipStride), codes.data(), static_cast<size_t>(E.P), static_cast<size_t>(E.P),
                             static_cast<size_t>(E.nTips), cudaMemcpyHostToDevice, E.stream));
        CK(cudaStreamSynchronize(E.stream));
        std::vector<unsigned char>().swap(codes);
        E.tipSet.assign(E.nTips, 1);
        E.forgetPlans();
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

extern "C" int beagleSetPatternWeights(int instance, const double* w) {
    if (!w) return BEAGLE_ERROR_OUT_OF_RANGE;
    return withEngine(instance, [&](Engine& E) {
        CK(cudaMemcpyAsync(E.dPatWeights, w, sizeof(double) * E.P, cudaMemcpyHostToDevice, E.stream));
        CK(cudaStreamSynchronize(E.stream));
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

extern "C" int beagleSetStateFrequencies(int instance, int idx, const double* f) {
    if (!f) return BEAGLE_ERROR_OUT_OF_RANGE;
    return withEngine(instance, [&](Engine& E) {
        if (idx < 0 || idx >= E.nEigen) return static_cast<int>(BEAGLE_ERROR_OUT_OF_RANGE);
        E.waitBlockFree();
        std::memcpy(E.hModel(idx)->freqs, f, sizeof(double) * 4);
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

extern "C" int beagleSetEigenDecomposition(int instance, int idx, const double* evec, const double* ivec,
                                           const double* eval) {
    if (!evec || !ivec || !eval) return BEAGLE_ERROR_OUT_OF_RANGE;
    return withEngine(instance, [&](Engine& E) {
        if (idx < 0 || idx >= E.nEigen) return static_cast<int>(BEAGLE_ERROR_OUT_OF_RANGE);
        E.waitBlockFree();
        fillCmat(evec, ivec, E.hModel(idx)->cmat);
        std::memcpy(E.hModel(idx)->evals, eval, sizeof(double) * 4);
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

extern "C" int beagleSetCategoryRates(int instance, const double* r) {
    if (!r) return BEAGLE_ERROR_OUT_OF_RANGE;
    return withEngine(instance, [&](Engine& E) {
        E.waitBlockFree();
        std::memcpy(E.hHeader()->rates, r, sizeof(double) * E.C);
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

extern "C" int beagleSetCategoryWeights(int instance, int idx, const double* w) {
    if (!w) return BEAGLE_ERROR_OUT_OF_RANGE;
    return withEngine(instance, [&](Engine& E) {
        if (idx < 0 || idx >= E.nEigen) return static_cast<int>(BEAGLE_ERROR_OUT_OF_RANGE);
        E.waitBlockFree();
        std::memcpy(E.hModel(idx)->weights, w, sizeof(double) * E.C);
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

extern "C" int beagleUpdateTransitionMatrices(int instance, int eigenIndex, const int* probabilityIndices,
                                              const int* d1, const int* d2, const double* edgeLengths, int count) {
    if (d1 || d2) return BEAGLE_ERROR_NO_IMPLEMENTATION;
    if (count < 0 || (count > 0 && (!probabilityIndices || !edgeLengths))) return BEAGLE_ERROR_OUT_OF_RANGE;
    return withEngine(instance, [&](Engine& E) {
        if (eigenIndex != 0 && (eigenIndex < 0 || eigenIndex >= E.nEigen)) return static_cast<int>(BEAGLE_ERROR_OUT_OF_RANGE);
        if (eigenIndex != 0) return static_cast<int>(BEAGLE_ERROR_NO_IMPLEMENTATION);
        int rc = E.setEdges(probabilityIndices, edgeLengths, count);
        if (rc) return rc;
        E.uploadBlock();
        Engine* es[1] = {&E};
        launchMatricesSet(es, 1);
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

extern "C" int beagleResetScaleFactors(int instance, int cumulativeScaleIndex) {
    return withEngine(instance, [&](Engine& E) {
        if (cumulativeScaleIndex < 0 || cumulativeScaleIndex >= E.nScale) return static_cast<int>(BEAGLE_ERROR_OUT_OF_RANGE);
        E.scaleBuffer(cumulativeScaleIndex);
        E.scaleZeroPending[cumulativeScaleIndex] = 1;   // folded into the next partials update when possible
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

extern "C" int beagleUpdatePartials(int instance, const BeagleOperation* operations, int operationCount,
                                    int cumulativeScaleIndex) {
    if (operationCount < 0 || (operationCount > 0 && !operations)) return BEAGLE_ERROR_OUT_OF_RANGE;
    return withEngine(instance, [&](Engine& E) {
        if (cumulativeScaleIndex != BEAGLE_OP_NONE && (cumulativeScaleIndex < 0 || cumulativeScaleIndex >= E.nScale))
            return static_cast<int>(BEAGLE_ERROR_OUT_OF_RANGE);
        E.setBatchHint(1);
        int rc = E.prepareSchedule(operations, operationCount, cumulativeScaleIndex != BEAGLE_OP_NONE);
        if (rc) return rc;
        if (E.keepSubtreeScalers && cumulativeScaleIndex != BEAGLE_OP_NONE) {
            rc = E.checkSlots(E.sched);
            if (rc) return rc;
        }
        Engine* es[1] = {&E};
        launchPartialsSet(es, 1, &cumulativeScaleIndex);
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

extern "C" int beagleCalculateEdgeLogLikelihoods(int instance, const int* parentBufferIndices, const int* childBufferIndices,
                                                 const int* probabilityIndices, const int* d1, const int* d2,
                                                 const int* categoryWeightsIndices, const int* stateFrequenciesIndices,
                                                 const int* cumulativeScaleIndices, int count, double* outSumLogLikelihood,
                                                 double* outD1, double* outD2) {
    if (d1 || d2 || outD1 || outD2 || count != 1) return BEAGLE_ERROR_NO_IMPLEMENTATION;
    if (!parentBufferIndices || !childBufferIndices || !probabilityIndices || !categoryWeightsIndices ||
        !stateFrequenciesIndices || !cumulativeScaleIndices || !outSumLogLikelihood)
        return BEAGLE_ERROR_OUT_OF_RANGE;
    return withEngine(instance, [&](Engine& E) {
        Engine* es[1] = {&E};
        const RootCall rc1{parentBufferIndices[0], childBufferIndices[0], probabilityIndices[0], categoryWeightsIndices[0],
                           stateFrequenciesIndices[0], cumulativeScaleIndices[0], E.hOut};
        int rc = launchRootSet(es, 1, &rc1);
        if (rc) return rc;
        return E.readScalar(outSumLogLikelihood);
    });
}

// =============================================================================================
// Engine-level entry points (include/strom_b200.h)
// =============================================================================================
extern "C" const char* sb200Version(void) { return "strom_b200 0.1.0 (sm_100a)"; }

extern "C" int sb200SetDevice(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return BEAGLE_ERROR_NO_RESOURCE;
    }
    if (device < 0 || device >= n) return BEAGLE_ERROR_NO_RESOURCE;
    t_device = device;
    g_defaultDevice.store(device);
    return BEAGLE_SUCCESS;
}

extern "C" int sb200SetStream(int instance, void* cudaStream) {
    return withEngine(instance, [&](Engine& E) {
        CK(cudaStreamSynchronize(E.stream));
        E.stream = cudaStream ? static_cast<cudaStream_t>(cudaStream) : E.ownStream;
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

extern "C" int sb200Eval(int instance, const StromEvalArgs* args, double* outLogLikelihood) {
    if (!outLogLikelihood || !args) return BEAGLE_ERROR_OUT_OF_RANGE;
    return withEngine(instance, [&](Engine& E) {
        Engine* es[1] = {&E};
        int rc = evalSet(es, 1, args, nullptr);
        if (rc) return rc;
        return E.readScalar(outLogLikelihood);
    });
}

extern "C" int sb200EvalAsync(int instance, const StromEvalArgs* args, double* deviceOut) {
    if (!args) return BEAGLE_ERROR_OUT_OF_RANGE;
    return withEngine(instance, [&](Engine& E) {
        Engine* es[1] = {&E};
        double* outs[1] = {deviceOut};
        return evalSet(es, 1, args, outs);
    });
}

// Up to kMaxBatch instances per set of launches; longer lists are cut into sets.
template <typename F>
static int forSets(int count, const int* instances, F&& body) {
    if (count < 1 || !instances) return BEAGLE_ERROR_OUT_OF_RANGE;
    std::vector<Engine*> es(static_cast<size_t>(count));
    for (int i = 0; i < count; ++i) {
        es[i] = getEngine(instances[i]);
        if (!es[i]) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
        for (int j = 0; j < i; ++j)
            if (es[j] == es[i]) return BEAGLE_ERROR_OUT_OF_RANGE;            // an instance holds ONE evaluation at a time
    }
    return onDevice(es[0]->device, [&]() {
        for (int first = 0; first < count; first += kMaxBatch) {
            const int K = std::min(kMaxBatch, count - first);
            const int rc = body(es.data() + first, K, first);
            if (rc) return rc;
        }
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

extern "C" int sb200EvalBatchAsync(int count, const int* instances, const StromEvalArgs* args) {
    if (!args) return BEAGLE_ERROR_OUT_OF_RANGE;
    return forSets(count, instances, [&](Engine* const* es, int K, int first) { return evalSet(es, K, args + first, nullptr); });
}

extern "C" int sb200Finish(int instance, double* outLogLikelihood) {
    if (!outLogLikelihood) return BEAGLE_ERROR_OUT_OF_RANGE;
    return withEngine(instance, [&](Engine& E) { return E.readScalar(outLogLikelihood); });
}

static int replaySet(Engine* const* es, int K, double* const* deviceOut) {
    int cumIdx[kMaxBatch];
    RootCall rc[kMaxBatch];
    for (int z = 0; z < K; ++z) {
        Engine& E = *es[z];
        if (!E.schedValid || E.lastParent < 0) return static_cast<int>(BEAGLE_ERROR_GENERAL);
        if (E.device != es[0]->device || E.C != es[0]->C || E.fp32 != es[0]->fp32 || E.stream != es[0]->stream)
            return static_cast<int>(BEAGLE_ERROR_OUT_OF_RANGE);
        cumIdx[z] = E.lastCumIndex;
        if (E.lastCumIndex != BEAGLE_OP_NONE) E.scaleZeroPending[E.lastCumIndex] = 1;
        rc[z] = RootCall{E.lastParent, E.lastChild, E.lastMatrix, E.lastWIdx, E.lastFIdx, E.lastRootCum,
                         (deviceOut && deviceOut[z]) ? deviceOut[z] : E.hOut};
    }
    launchMatricesSet(es, K);
    launchPartialsSet(es, K, cumIdx);
    return launchRootSet(es, K, rc);
}

extern "C" int sb200ReplayAsync(int instance, double* deviceOut) {
    return withEngine(instance, [&](Engine& E) {
        Engine* es[1] = {&E};
        double* outs[1] = {deviceOut};
        return replaySet(es, 1, outs);
    });
}

extern "C" int sb200ReplayBatchAsync(int count, const int* instances) {
    return forSets(count, instances, [&](Engine* const* es, int K, int) { return replaySet(es, K, nullptr); });
}

extern "C" int sb200Synchronize(int instance) {
    return withEngine(instance, [&](Engine& E) {
        CK(cudaStreamSynchronize(E.stream));
        E.blockInFlight = false;
        E.schedInFlight = false;
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

extern "C" int sb200GetPartials(int instance, int bufferIndex, double* out) {
    if (!out) return BEAGLE_ERROR_OUT_OF_RANGE;
    return withEngine(instance, [&](Engine& E) {
        if (bufferIndex < E.nTips || bufferIndex >= E.nTips + E.nPartials) return static_cast<int>(BEAGLE_ERROR_OUT_OF_RANGE);
        // HBM holds [category][Ppad][4]; the caller gets BeagleLib's [pattern][category][4] ... as documented in the header
        const size_t stride = static_cast<size_t>(E.Ppad) * E.C * 4;
        std::vector<double> full(stride);
        CK(cudaStreamSynchronize(E.stream));
        if (E.fp32) {
            std::vector<float> tmp(stride);
            CK(cudaMemcpy(tmp.data(), static_cast<float*>(E.dPartials) + stride * (bufferIndex - E.nTips), stride * 4, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < stride; ++i) full[i] = tmp[i];
        } else {
            CK(cudaMemcpy(full.data(), static_cast<double*>(E.dPartials) + stride * (bufferIndex - E.nTips), stride * 8, cudaMemcpyDeviceToHost));
        }
        for (int p = 0; p < E.P; ++p)
            for (int k = 0; k < E.C; ++k)
                for (int i = 0; i < 4; ++i)
                    out[(static_cast<size_t>(p) * E.C + k) * 4 + i] = full[(static_cast<size_t>(k) * E.Ppad + p) * 4 + i];
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

extern "C" int sb200GetScaleFactors(int instance, int scaleIndex, double* out) {
    if (!out) return BEAGLE_ERROR_OUT_OF_RANGE;
    return withEngine(instance, [&](Engine& E) {
        if (scaleIndex < 0 || scaleIndex >= E.nScale) return static_cast<int>(BEAGLE_ERROR_OUT_OF_RANGE);
        E.materializeZero(scaleIndex);
        CK(cudaStreamSynchronize(E.stream));
        CK(cudaMemcpy(out, E.dScale[scaleIndex], sizeof(double) * E.P, cudaMemcpyDeviceToHost));
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

extern "C" int sb200GetTransitionMatrix(int instance, int matrixIndex, double* out) {
    if (!out) return BEAGLE_ERROR_OUT_OF_RANGE;
    return withEngine(instance, [&](Engine& E) {
        if (matrixIndex < 0 || matrixIndex >= E.nMat) return static_cast<int>(BEAGLE_ERROR_OUT_OF_RANGE);
        const size_t n = static_cast<size_t>(E.C) * kMatStride;
        std::vector<double> full(n);
        CK(cudaStreamSynchronize(E.stream));
        if (E.fp32) {
            std::vector<float> tmp(n);
            CK(cudaMemcpy(tmp.data(), static_cast<float*>(E.dMats) + n * matrixIndex, n * 4, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < n; ++i) full[i] = tmp[i];
        } else {
            CK(cudaMemcpy(full.data(), static_cast<double*>(E.dMats) + n * matrixIndex, n * 8, cudaMemcpyDeviceToHost));
        }
        for (int k = 0; k < E.C; ++k)
            for (int e = 0; e < 16; ++e) out[k * 16 + e] = full[pmIndex(k, e)];
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

extern "C" int sb200SetProfiling(int instance, int enabled) {
    return withEngine(instance, [&](Engine& E) {
        E.collectProfile();
        E.profiling = enabled != 0;
        E.partialsMsSum = 0;
        E.partialsCalls = 0;
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

extern "C" int sb200PartialsTime(int instance, double* totalMs, long* calls) {
    return withEngine(instance, [&](Engine& E) {
        E.collectProfile();
        if (totalMs) *totalMs = E.partialsMsSum;
        if (calls) *calls = E.partialsCalls;
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

// Tuning aid (kernels built with -DSB200_TRACE): clock64 samples of block (0,0), 8 per step, 60 steps,
// of the LAST pruning launch.  Allocates the buffer on first use; pass NULL to only enable.
extern "C" __attribute__((visibility("default"))) int sb200DebugTrace(int instance, long long* out) {
    return withEngine(instance, [&](Engine& E) {
        if (!E.dTrace) {
            CK(cudaMalloc(&E.dTrace, sizeof(long long) * 480));
            CK(cudaMemset(E.dTrace, 0, sizeof(long long) * 480));
            E.forgetPlans();
        }
        if (out) {
            CK(cudaStreamSynchronize(E.stream));
            CK(cudaMemcpy(out, E.dTrace, sizeof(long long) * 480, cudaMemcpyDeviceToHost));
        }
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

extern "C" int sb200SetSubtreeScalers(int instance, int enable) {
    return withEngine(instance, [&](Engine& E) {
        const bool on = enable != 0;
        if (on != E.keepSubtreeScalers) {
            E.keepSubtreeScalers = on;
            E.slotValid.assign(E.nPartials, 0);
            E.forgetPlans();                           // every plan depends on it
        }
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

extern "C" int sb200SetTransientPartials(int instance, int enable) {
    return withEngine(instance, [&](Engine& E) {
        E.transientPartials = enable != 0;
        return static_cast<int>(BEAGLE_SUCCESS);
    });
}

extern "C" long sb200LaunchCount(int instance) {
    Engine* E = getEngine(instance);
    return E ? E->launches : -1;
}

extern "C" int sb200LastTraffic(int instance, double* algorithmicBytes, double* scheduledBytes) {
    Engine* E = getEngine(instance);
    if (!E) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
    if (algorithmicBytes) *algorithmicBytes = E->algBytes;
    if (scheduledBytes) {
        *scheduledBytes = E->schedBytes;
        if (E->transientPartials && !E->keepSubtreeScalers && !E->sched.sequential)      // only chain ends are written
            *scheduledBytes -= static_cast<double>(E->sched.partialWrites - static_cast<long>(E->sched.chains.size())) * 4.0 *
                               static_cast<double>(E->realSize()) * static_cast<double>(E->P) * E->C;
    }
    return BEAGLE_SUCCESS;
}

extern "C" int sb200DebugSchedule(const BeagleOperation* operations, int operationCount, int tipCount,
                                  int partialsBufferCount, int* levelOut, int* chainOut, int* chainCountOut) {
    try {
        Schedule s;
        int rc = buildSchedule(operations, operationCount, tipCount, partialsBufferCount, 1 << 30, 1 << 30, true, 1, s);
        if (rc) return rc;
        for (int o = 0; o < operationCount; ++o) {
            if (levelOut) levelOut[o] = s.opLevel[o];
            if (chainOut) chainOut[o] = s.opChain[o];
        }
        if (chainCountOut) *chainCountOut = static_cast<int>(s.chains.size());
        return static_cast<int>(s.levels.size());
    } catch (...) {
        return BEAGLE_ERROR_UNIDENTIFIED_EXCEPTION;
    }
}

// Host-only inspection of the launches of a plan (see include/strom_b200.h).
extern "C" int sb200DebugPlanLaunches(const BeagleOperation* operations, int operationCount, int tipCount, int partialsBufferCount,
                                      int keepSubtreeScalers, int groupsPerLaunch, int groupsPerTipsLaunch, int maxLaunches,
                                      int* launchesOut) {
    try {
        Schedule s;
        int rc = buildSchedule(operations, operationCount, tipCount, partialsBufferCount, 1 << 30, 1 << 30, true, groupsPerLaunch, s,
                               keepSubtreeScalers != 0, groupsPerTipsLaunch);
        if (rc) return rc;
        const int n = static_cast<int>(s.levels.size());
        if (launchesOut)
            for (int l = 0; l < n && l < maxLaunches; ++l) {
                const Level& L = s.levels[l];
                int ops = 0, longest = 0;
                for (int g = L.groupStart; g < L.groupStart + L.groupCount; ++g) {
                    ops += s.groups[g].opEnd - s.groups[g].opStart;
                    longest = std::max(longest, s.groups[g].opEnd - s.groups[g].opStart);
                }
                int* r = launchesOut + 5 * l;
                r[0] = L.chainCount; r[1] = L.groupCount; r[2] = L.tipsOnly; r[3] = ops; r[4] = longest;
            }
        return n;
    } catch (...) {
        return BEAGLE_ERROR_UNIDENTIFIED_EXCEPTION;
    }
}

// Host-only inspection of a plan (see include/strom_b200.h).
extern "C" int sb200DebugPlan(const BeagleOperation* operations, int operationCount, int tipCount, int partialsBufferCount,
                              int keepSubtreeScalers, int* planOut) {
    try {
        Schedule s;
        int rc = buildSchedule(operations, operationCount, tipCount, partialsBufferCount, 1 << 30, 1 << 30, true, 4, s,
                               keepSubtreeScalers != 0);
        if (rc) return rc;
        if (planOut) {
            for (size_t l = 0; l < s.levels.size(); ++l)
                for (int g = s.levels[l].groupStart; g < s.levels[l].groupStart + s.levels[l].groupCount; ++g)
                    for (int o = s.groups[g].opStart; o < s.groups[g].opEnd; ++o) {
                        const DevOp& d = s.ops[o];
                        int* r = planOut + 8 * static_cast<size_t>(o);
                        r[0] = d.dest + tipCount; r[1] = d.a; r[2] = d.b; r[3] = d.sa; r[4] = d.sb; r[5] = d.sOut; r[6] = d.flags;
                        r[7] = static_cast<int>(l);
                    }
        }
        return static_cast<int>(s.ops.size());
    } catch (...) {
        return BEAGLE_ERROR_UNIDENTIFIED_EXCEPTION;
    }
}
