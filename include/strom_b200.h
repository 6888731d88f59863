/*
 * strom_b200.h -- engine-level C ABI of libstrom_b200.so (beyond the 13 BeagleLib-compatible
 * calls of strom_beagle.h).  These are the "thin C-ABI layer over CUDA streams" the host
 * (strom_b200/host/likelihood.hpp, bench.py, tests) uses for:
 *   - device / stream selection (one instance per GPU, caller-provided stream),
 *   - byte-coded tip upload (reference path: Likelihood::setTipStates, src/likelihood.hpp:228-247,
 *     uploads 4-byte ints; at 1000 taxa x 1M patterns the engine takes the 1-byte codes directly),
 *   - one fused evaluation call replacing the per-call sequence of
 *     Likelihood::calcLogLikelihood (src/likelihood.hpp:409-447),
 *   - a non-synchronising variant that leaves the scalar on the device for the cross-GPU sum,
 *   - buffer read-back for per-buffer parity tests,
 *   - schedule inspection (host logic, callable without a GPU).
 * All functions return a BeagleReturnCodes value unless stated otherwise.
 */
#ifndef STROM_B200_H
#define STROM_B200_H

#include "strom_beagle.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Model + tree description of one evaluation; every pointer is caller-owned host memory. */
typedef struct {
    const double* stateFrequencies;      /* [4]            src/model.hpp:327-335 */
    const double* eigenVectors;          /* [16] row-major src/model.hpp:315-325 */
    const double* inverseEigenVectors;   /* [16] row-major */
    const double* eigenValues;           /* [4] */
    const double* categoryRates;         /* [categoryCount] src/model.hpp:337-344 */
    const double* categoryWeights;       /* [categoryCount] src/model.hpp:346-354 */
    const int*    matrixIndices;         /* [edgeCount]     src/likelihood.hpp:308-315, :353-354 */
    const double* edgeLengths;           /* [edgeCount] */
    int           edgeCount;
    const BeagleOperation* operations;   /* [operationCount] src/likelihood.hpp:321-347 */
    int           operationCount;
    int           rootParentBuffer;      /* preorder[0] number, src/likelihood.hpp:427 */
    int           rootChildBuffer;       /* root tip number,   src/likelihood.hpp:424 */
    int           rootMatrix;            /* = rootChildBuffer, src/likelihood.hpp:433 */
    int           categoryCount;         /* entries of categoryRates / categoryWeights; must equal the instance's
                                            category count (0 = not checked) */
} StromEvalArgs;

/* Library version string (static storage). */
STROM_API const char* sb200Version(void);

/* CUDA device new instances are created on (default: device 0). */
STROM_API int sb200SetDevice(int device);

/* Run the instance on a caller-owned CUDA stream (cudaStream_t passed as void*); NULL restores
 * the instance's own stream. */
STROM_API int sb200SetStream(int instance, void* cudaStream);

/* Tip upload from 1-byte codes (0..3, anything else = missing). */
STROM_API int sb200SetTipStatesU8(int instance, int tipIndex, const unsigned char* inStates);

/* All tips in one call: inStates[tip * patternCount + pattern], 1-byte codes as above (one strided copy instead of a copy
 * per taxon; reference path: the per-taxon loop of Likelihood::setTipStates, src/likelihood.hpp:228-247). */
STROM_API int sb200SetAllTipStatesU8(int instance, const unsigned char* inStates);

/* One whole Likelihood::calcLogLikelihood: parameter upload, all transition matrices, all
 * partials with rescaling, root-edge reduction, scalar back on the host.  Synchronises. */
STROM_API int sb200Eval(int instance, const StromEvalArgs* args, double* outLogLikelihood);

/* Same work, no host synchronisation: the pattern-weighted log-likelihood of this instance's
 * pattern slice is left in *deviceOut (a device pointer to one double) in stream order, or, when
 * deviceOut is NULL, in the instance's own result slot, to be collected with sb200Finish. */
STROM_API int sb200EvalAsync(int instance, const StromEvalArgs* args, double* deviceOut);

/* The evaluations of `count` instances (the heated chains of a Metropolis-coupled run: same data, one tree and model
 * each; the reference steps them one after the other, src/strom.hpp:316-324) as ONE set of kernel launches: the
 * instance is a grid dimension of every kernel, so eight 738-taxon trees cost the launches -- and fill the GPU far
 * better than -- one.  args[i] belongs to instances[i]; all instances must live on one device and share the category
 * count and the precision; they run on the first instance's stream from this call on.  Lists longer than 8 are cut
 * into sets of 8.  Every scalar is collected with sb200Finish(instances[i]).  Nothing is enqueued if any argument is
 * rejected.  The parameter blocks of a set's instances share one pinned / device arena and go up in one copy.  The
 * instances of a set must be driven from ONE host thread at a time (as every single instance must). */
STROM_API int sb200EvalBatchAsync(int count, const int* instances, const StromEvalArgs* args);

/* sb200ReplayAsync for a set of instances that were last evaluated together by sb200EvalBatchAsync. */
STROM_API int sb200ReplayBatchAsync(int count, const int* instances);

/* Waits for the instance's stream and returns the scalar of the last evaluation that used the
 * instance's own result slot (BEAGLE_ERROR_FLOATING_POINT if it is NaN).  With one instance per
 * GPU, a host enqueues sb200EvalAsync on every shard and then collects with sb200Finish. */
STROM_API int sb200Finish(int instance, double* outLogLikelihood);

/* Re-run the last evaluation's device work (matrices, partials, root reduction) with the
 * parameters already resident in HBM; no host->device copy, no synchronisation.  Used by
 * bench.py for the device-only `value` leg. */
STROM_API int sb200ReplayAsync(int instance, double* deviceOut);

/* Blocks until the instance's stream is idle. */
STROM_API int sb200Synchronize(int instance);

/* Read-back for parity tests: partials as [pattern][category][4] doubles (FP32 instances are
 * widened), cumulative scale buffer as [pattern] log factors, matrix as [category][4][4]. */
STROM_API int sb200GetPartials(int instance, int bufferIndex, double* out);
STROM_API int sb200GetScaleFactors(int instance, int scaleIndex, double* out);
STROM_API int sb200GetTransitionMatrix(int instance, int matrixIndex, double* out);

/* CUDA-event timing of the pruning launches, on the stream they are launched on: enable, run
 * evaluations, then read the accumulated milliseconds and the number of timed calls. */
STROM_API int sb200SetProfiling(int instance, int enabled);
STROM_API int sb200PartialsTime(int instance, double* totalMs, long* calls);

/* Number of kernels this instance has launched since creation (bench.py's gpu_launches). */
STROM_API long sb200LaunchCount(int instance);

/* Algorithmic HBM bytes of the last beagleUpdatePartials call per SURVEY.md section 8(d)
 * (unfused per-operation traffic), and the bytes the chain-fused schedule actually moves. */
STROM_API int sb200LastTraffic(int instance, double* algorithmicBytes, double* scheduledBytes);

/* Incremental evaluation (the reference recomputes every node on every call,
 * src/likelihood.hpp:409-413).  With enable != 0 the instance keeps, for every partials buffer, the total
 * log-scaler of the subtree below it (16 bytes per pattern and buffer of extra HBM), and an operation list given
 * to beagleUpdatePartials / sb200Eval may then cover only the nodes whose partials changed since the previous
 * call -- children before parents, ending at the root's partials -- together with only the transition matrices
 * that changed: operands that the list does not produce are taken, with their subtree scalers, as the previous
 * calls left them, and the cumulative scale buffer still receives the whole tree's log-scaler.  The first call
 * after enabling must list every node.  Lists must be forests (no buffer written twice or read before written). */
STROM_API int sb200SetSubtreeScalers(int instance, int enable);

/* Transient partials.  The planner fuses the operations of a list into chains: inside a chain the partials of a node
 * go from one operation to the next in registers, and the copy written to the node's buffer is read by nobody -- the
 * reference recomputes every node on every call (src/likelihood.hpp:409-413) and never reads a partials buffer back.
 * With enable != 0 only the buffers a later launch (or the root edge) reads are written: the last node of every chain.
 * Same operations, same arithmetic, same log-likelihood bit for bit; 40 % less HBM traffic on a 1000-taxon tree.
 * OFF by default: as in BeagleLib, every destination buffer then holds its node's partials after
 * beagleUpdatePartials (sb200GetPartials of a buffer inside a chain is meaningless while this is on).  Ignored
 * while sb200SetSubtreeScalers is on and for lists that are not forests (both read buffers of earlier operations
 * from memory). */
STROM_API int sb200SetTransientPartials(int instance, int enable);

/* Site-pattern compression on the device: what Data::compressPatterns does with a
 * std::map over all site columns (reference src/data.hpp:102-161).  codes[taxon * nsites + site] are state codes
 * < 32; on return patternsOut[taxon * (*patternCountOut) + pattern] holds the distinct site columns in
 * lexicographic order (taxon 0 most significant, codes compared as integers) and countsOut[pattern] how many sites
 * show each.  patternsOut needs room for ntaxa * nsites bytes, countsOut for nsites doubles.  Runs on CUDA device
 * `device`; returns BEAGLE_ERROR_OUT_OF_RANGE (and writes nothing) if a code is >= 32. */
STROM_API int sb200CompressPatterns(int device, int ntaxa, int nsites, const unsigned char* codes, unsigned char* patternsOut,
                                    double* countsOut, int* patternCountOut);

/* Host-only schedule inspection (no GPU needed): builds the chain/level schedule for an
 * operation list and reports, per operation, its level and chain id.  Returns the number of
 * levels (>= 0) or a negative error code.  chainOut/levelOut may be NULL. */
STROM_API int sb200DebugSchedule(const BeagleOperation* operations, int operationCount, int tipCount,
                                 int partialsBufferCount, int* levelOut, int* chainOut, int* chainCountOut);

/* Host-only inspection of the plan built for an operation list (with or without the permanent subtree scalers
 * of sb200SetSubtreeScalers): per planned operation o, in launch and group order, planOut[8o..8o+7] = {destination
 * buffer, operand A buffer or -1 (carried in registers), operand B buffer, scaler slot read with A or -1, scaler
 * slot read with B or -1, scaler slot written or -1, flag bits (1 rescale, 2 chain start, 4 chain end, 8 adds to the
 * cumulative buffer, 16 tip x tip, 32 A is a tip, 64 B is a tip, 128 stores its subtree scaler and carries on),
 * launch index}.  Returns the number of planned operations (>= 0) or a negative error code. */
STROM_API int sb200DebugPlan(const BeagleOperation* operations, int operationCount, int tipCount, int partialsBufferCount,
                             int keepSubtreeScalers, int* planOut);

/* Host-only, needs no GPU: the launches the planner makes of an operation list.  groupsPerLaunch / groupsPerTipsLaunch: the
 * most thread-block groups a launch (a tips-only launch; 0 = the same) is split into.  launchesOut receives, per launch (at
 * most maxLaunches), 5 ints: {chains, groups, tips-only (every operand read from memory is a tip and no subtree scaler is
 * read: the launch may run the kernel variant without partials-operand code), operations, operations of the longest group}.
 * Returns the number of launches (>= 0) or a negative error code. */
STROM_API int sb200DebugPlanLaunches(const BeagleOperation* operations, int operationCount, int tipCount, int partialsBufferCount,
                                     int keepSubtreeScalers, int groupsPerLaunch, int groupsPerTipsLaunch, int maxLaunches,
                                     int* launchesOut);

#ifdef __cplusplus
}
#endif
#endif /* STROM_B200_H */
