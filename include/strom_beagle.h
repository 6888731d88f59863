/*
 * strom_beagle.h -- the drop-in boundary of the strom tree-likelihood hot path.
 *
 * strom delegates every floating-point operation of Likelihood::calcLogLikelihood
 * to 13 BeagleLib C calls (reference src/likelihood.hpp:8, src/model.hpp:5;
 * `#include "libhmsbeagle/beagle.h"`, linked with -lhmsbeagle, src/Makefile.am:29).
 * This header declares exactly those 13 entry points, with the argument meaning and
 * error behaviour the reference's call sites rely on.  libstrom_b200.so exports them
 * backed by hand-written sm_100a CUDA kernels (no CPU fallback); the test oracle
 * (oracle/beagle_cpu.c) exports the same symbols backed by a CPU restatement.
 *
 * Plain C ABI: ints, longs, pointers to caller-owned arrays.  Every input array is
 * copied before the call returns; the library owns all instance buffers between
 * beagleCreateInstance and beagleFinalizeInstance.  No exception crosses this line.
 */
#ifndef STROM_BEAGLE_H
#define STROM_BEAGLE_H

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define STROM_API __attribute__((visibility("default")))
#else
#define STROM_API
#endif

/* Return codes; the reference's message table is src/likelihood.hpp:79-87. */
enum BeagleReturnCodes {
    BEAGLE_SUCCESS                      =  0,
    BEAGLE_ERROR_GENERAL                = -1,
    BEAGLE_ERROR_OUT_OF_MEMORY          = -2,
    BEAGLE_ERROR_UNIDENTIFIED_EXCEPTION = -3,
    BEAGLE_ERROR_UNINITIALIZED_INSTANCE = -4,
    BEAGLE_ERROR_OUT_OF_RANGE           = -5,
    BEAGLE_ERROR_NO_RESOURCE            = -6,
    BEAGLE_ERROR_NO_IMPLEMENTATION      = -7,
    BEAGLE_ERROR_FLOATING_POINT         = -8
};

/* Flags used by the reference at src/likelihood.hpp:186-195 (bit values follow the
 * public BEAGLE API so a caller compiled against either header passes the same bits). */
enum BeagleFlags {
    BEAGLE_FLAG_PRECISION_SINGLE = 1L << 0,
    BEAGLE_FLAG_PRECISION_DOUBLE = 1L << 1,
    BEAGLE_FLAG_COMPUTATION_SYNCH = 1L << 2,
    BEAGLE_FLAG_COMPUTATION_ASYNCH = 1L << 3,
    BEAGLE_FLAG_EIGEN_REAL       = 1L << 4,
    BEAGLE_FLAG_EIGEN_COMPLEX    = 1L << 5,
    BEAGLE_FLAG_SCALING_MANUAL   = 1L << 6,
    BEAGLE_FLAG_SCALING_AUTO     = 1L << 7,
    BEAGLE_FLAG_SCALING_ALWAYS   = 1L << 8,
    BEAGLE_FLAG_SCALERS_RAW      = 1L << 9,
    BEAGLE_FLAG_SCALERS_LOG      = 1L << 10,
    BEAGLE_FLAG_VECTOR_SSE       = 1L << 11,
    BEAGLE_FLAG_VECTOR_NONE      = 1L << 12,
    BEAGLE_FLAG_THREADING_OPENMP = 1L << 13,
    BEAGLE_FLAG_THREADING_NONE   = 1L << 14,
    BEAGLE_FLAG_PROCESSOR_CPU    = 1L << 15,
    BEAGLE_FLAG_PROCESSOR_GPU    = 1L << 16
};

/* "no buffer" marker in operation lists (src/likelihood.hpp:329). */
enum BeagleOpCodes { BEAGLE_OP_COUNT = 7, BEAGLE_OP_NONE = -1 };

typedef struct {
    int   resourceNumber;
    char* resourceName;
    char* implName;
    char* implDescription;
    long  flags;
} BeagleInstanceDetails;

typedef struct {
    char* name;
    char* description;
    long  supportFlags;
    long  requiredFlags;
} BeagleResource;

typedef struct {
    BeagleResource* list;
    int             length;
} BeagleResourceList;

/* Seven consecutive ints; the reference casts an int vector to this
 * (src/likelihood.hpp:321-347, :382). */
typedef struct {
    int destinationPartials;
    int destinationScaleWrite;
    int destinationScaleRead;
    int child1Partials;
    int child1TransitionMatrix;
    int child2Partials;
    int child2TransitionMatrix;
} BeagleOperation;

/* src/likelihood.hpp:105 -- library-owned static storage. */
STROM_API BeagleResourceList* beagleGetResourceList(void);

/* src/likelihood.hpp:198-212 -- returns instance id >= 0 or a negative error code.
 * stateCount must be 4; requirementFlags must not ask for anything but
 * PRECISION_DOUBLE|SCALING_MANUAL (PRECISION_SINGLE selects the FP32 variant). */
STROM_API int beagleCreateInstance(int tipCount, int partialsBufferCount, int compactBufferCount,
                                   int stateCount, int patternCount, int eigenBufferCount,
                                   int matrixBufferCount, int categoryCount, int scaleBufferCount,
                                   int* resourceList, int resourceCount,
                                   long preferenceFlags, long requirementFlags,
                                   BeagleInstanceDetails* returnInfo);

/* src/likelihood.hpp:96, :131, :146 */
STROM_API int beagleFinalizeInstance(int instance);

/* src/likelihood.hpp:238-241 -- patternCount ints; values outside 0..3 mean missing. */
STROM_API int beagleSetTipStates(int instance, int tipIndex, const int* inStates);

/* src/likelihood.hpp:258-260 -- patternCount doubles. */
STROM_API int beagleSetPatternWeights(int instance, const double* inPatternWeights);

/* src/model.hpp:329-332 -- 4 doubles. */
STROM_API int beagleSetStateFrequencies(int instance, int stateFrequenciesIndex,
                                        const double* inStateFrequencies);

/* src/model.hpp:317-322 -- row-major 4x4 eigenvectors, inverse eigenvectors, 4 eigenvalues. */
STROM_API int beagleSetEigenDecomposition(int instance, int eigenIndex,
                                          const double* inEigenVectors,
                                          const double* inInverseEigenVectors,
                                          const double* inEigenValues);

/* src/model.hpp:339-341 -- categoryCount doubles. */
STROM_API int beagleSetCategoryRates(int instance, const double* inCategoryRates);

/* src/model.hpp:348-351 -- categoryCount doubles. */
STROM_API int beagleSetCategoryWeights(int instance, int categoryWeightsIndex,
                                       const double* inCategoryWeights);

/* src/likelihood.hpp:359-366 -- derivative index arrays must be NULL. */
STROM_API int beagleUpdateTransitionMatrices(int instance, int eigenIndex,
                                             const int* probabilityIndices,
                                             const int* firstDerivativeIndices,
                                             const int* secondDerivativeIndices,
                                             const double* edgeLengths, int count);

/* src/likelihood.hpp:374 */
STROM_API int beagleResetScaleFactors(int instance, int cumulativeScaleIndex);

/* src/likelihood.hpp:380-384 -- operations in dependency order (children first). */
STROM_API int beagleUpdatePartials(int instance, const BeagleOperation* operations,
                                   int operationCount, int cumulativeScaleIndex);

/* src/likelihood.hpp:429-442 -- derivative outputs must be NULL; count must be 1.
 * Returns BEAGLE_ERROR_FLOATING_POINT when the sum is NaN. */
STROM_API int beagleCalculateEdgeLogLikelihoods(int instance, const int* parentBufferIndices,
                                                const int* childBufferIndices,
                                                const int* probabilityIndices,
                                                const int* firstDerivativeIndices,
                                                const int* secondDerivativeIndices,
                                                const int* categoryWeightsIndices,
                                                const int* stateFrequenciesIndices,
                                                const int* cumulativeScaleIndices, int count,
                                                double* outSumLogLikelihood,
                                                double* outSumFirstDerivative,
                                                double* outSumSecondDerivative);

#ifdef __cplusplus
}
#endif
#endif /* STROM_BEAGLE_H */
