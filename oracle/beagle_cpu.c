/*
 * oracle/beagle_cpu.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the arithmetic that strom's hot path delegates to BeagleLib
 * (reference src/likelihood.hpp:198-442, src/model.hpp:315-354).  The arithmetic itself
 * lives in the third-party library BeagleLib (libhmsbeagle, API-v1 series 2.x/3.x,
 * `-lhmsbeagle` at src/Makefile.am:29; NOT vendored under /root/reference and not
 * version-pinned by any manifest).  This file restates BeagleLib's published CPU,
 * double-precision, 4-state, SCALING_MANUAL (raw scalers) algorithm:
 *
 *   - transition matrices: EigenDecompositionCube::updateTransitionMatrices
 *     (cMatrix[i][j][k] = evec[i][k]*ivec[k][j]; P = sum_k c*exp(eval[k]*(t*rate)); clamp <0 to 0;
 *      padded 5th column = 1 for the missing state),
 *   - partials: BeagleCPU4StateImpl calcStatesStates / calcStatesPartials / calcPartialsPartials,
 *     buffer layout [category][pattern][state],
 *   - rescalePartials: per-pattern max over categories and states, max==0 -> 1,
 *     multiply by 1/max, scale buffer holds raw max, cumulative buffer += log(max),
 *   - calcEdgeLogLikelihoods with a compact-state child, NaN -> BEAGLE_ERROR_FLOATING_POINT.
 *
 * Parity anchor: the reference's only two golden values for this path
 * (paup.nex:12 lnL -144730.8 for rbcl738/GTR+G4; rbcLjc.tre:57 -lnL 286.9238 for rbcL/JC69);
 * tests/test_oracle_golden.py pins this file against both.  Intermediate buffers are not
 * pinned by any reference test.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  It exports the same 13 symbols as include/strom_beagle.h, plus
 * oracleSetThreads / oracleGetPartials / oracleGetScaleFactors / oracleGetTransitionMatrix
 * for per-buffer parity checks.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "../include/strom_beagle.h"

#define S 4          /* states */
#define SP 5         /* padded matrix row: 4 states + missing column */
#define MAX_INST 64

typedef struct {
    int used;
    int tipCount, partialsCount, compactCount, patternCount, eigenCount, matrixCount, categoryCount,
        scaleCount;
    int**    tipStates;     /* [tipCount] -> patternCount ints (0..4) or NULL */
    double** partials;      /* [tipCount+partialsCount] -> C*P*4 (NULL for tips) */
    double** matrices;      /* [matrixCount] -> C*4*5 */
    double** scale;         /* [scaleCount] -> P */
    double*  cmat;          /* [eigenCount][64] */
    double*  evals;         /* [eigenCount][4] */
    double*  freqs;         /* [eigenCount][4] */
    double*  catWeights;    /* [eigenCount][C] */
    double*  catRates;      /* [C] */
    double*  patWeights;    /* [P] */
    double*  tmpInt;        /* [P*4] */
    int threads;
} Inst;

static Inst g_inst[MAX_INST];
static int g_threads = 1;

static BeagleResource g_res = {(char*)"CPU", (char*)"oracle: CPU restatement of BeagleLib double/manual-scaling path",
                               BEAGLE_FLAG_PRECISION_DOUBLE | BEAGLE_FLAG_SCALING_MANUAL | BEAGLE_FLAG_PROCESSOR_CPU, 0};
static BeagleResourceList g_reslist = {&g_res, 1};

STROM_API void oracleSetThreads(int n) { g_threads = n < 1 ? 1 : n; }

static Inst* get(int id) {
    if (id < 0 || id >= MAX_INST || !g_inst[id].used) return NULL;
    return &g_inst[id];
}

BeagleResourceList* beagleGetResourceList(void) { return &g_reslist; }

int beagleCreateInstance(int tipCount, int partialsBufferCount, int compactBufferCount, int stateCount,
                         int patternCount, int eigenBufferCount, int matrixBufferCount, int categoryCount,
                         int scaleBufferCount, int* resourceList, int resourceCount, long preferenceFlags,
                         long requirementFlags, BeagleInstanceDetails* returnInfo) {
    (void)resourceList; (void)resourceCount; (void)preferenceFlags;
    if (stateCount != S) return BEAGLE_ERROR_NO_IMPLEMENTATION;
    if (requirementFlags & (BEAGLE_FLAG_PRECISION_SINGLE | BEAGLE_FLAG_PROCESSOR_GPU | BEAGLE_FLAG_SCALING_AUTO |
                            BEAGLE_FLAG_SCALING_ALWAYS | BEAGLE_FLAG_EIGEN_COMPLEX))
        return BEAGLE_ERROR_NO_IMPLEMENTATION;
    if (tipCount < 0 || partialsBufferCount < 0 || patternCount < 1 || categoryCount < 1 || eigenBufferCount < 1 ||
        matrixBufferCount < 0 || scaleBufferCount < 0 || compactBufferCount < 0)
        return BEAGLE_ERROR_OUT_OF_RANGE;
    int id = -1;
    for (int i = 0; i < MAX_INST; ++i) if (!g_inst[i].used) { id = i; break; }
    if (id < 0) return BEAGLE_ERROR_OUT_OF_MEMORY;
    Inst* I = &g_inst[id];
    memset(I, 0, sizeof(*I));
    I->tipCount = tipCount; I->partialsCount = partialsBufferCount; I->compactCount = compactBufferCount;
    I->patternCount = patternCount; I->eigenCount = eigenBufferCount; I->matrixCount = matrixBufferCount;
    I->categoryCount = categoryCount; I->scaleCount = scaleBufferCount;
    size_t P = (size_t)patternCount, C = (size_t)categoryCount;
    I->tipStates = (int**)calloc((size_t)tipCount + 1, sizeof(int*));
    I->partials = (double**)calloc((size_t)(tipCount + partialsBufferCount) + 1, sizeof(double*));
    I->matrices = (double**)calloc((size_t)matrixBufferCount + 1, sizeof(double*));
    I->scale = (double**)calloc((size_t)scaleBufferCount + 1, sizeof(double*));
    I->cmat = (double*)calloc((size_t)eigenBufferCount * 64, sizeof(double));
    I->evals = (double*)calloc((size_t)eigenBufferCount * 4, sizeof(double));
    I->freqs = (double*)calloc((size_t)eigenBufferCount * 4, sizeof(double));
    I->catWeights = (double*)calloc((size_t)eigenBufferCount * C, sizeof(double));
    I->catRates = (double*)calloc(C, sizeof(double));
    I->patWeights = (double*)calloc(P, sizeof(double));
    I->tmpInt = (double*)calloc(P * S, sizeof(double));
    int ok = I->tipStates && I->partials && I->matrices && I->scale && I->cmat && I->evals && I->freqs &&
             I->catWeights && I->catRates && I->patWeights && I->tmpInt;
    for (int i = 0; ok && i < partialsBufferCount; ++i) {
        I->partials[tipCount + i] = (double*)calloc(C * P * S, sizeof(double));
        ok = I->partials[tipCount + i] != NULL;
    }
    for (int i = 0; ok && i < matrixBufferCount; ++i) {
        I->matrices[i] = (double*)calloc(C * S * SP, sizeof(double));
        ok = I->matrices[i] != NULL;
    }
    for (int i = 0; ok && i < scaleBufferCount; ++i) {
        I->scale[i] = (double*)calloc(P, sizeof(double));
        ok = I->scale[i] != NULL;
    }
    for (size_t k = 0; k < C; ++k) I->catRates[k] = 1.0;
    I->used = 1;
    I->threads = g_threads;
    if (!ok) { beagleFinalizeInstance(id); return BEAGLE_ERROR_OUT_OF_MEMORY; }
    if (returnInfo) {
        returnInfo->resourceNumber = 0;
        returnInfo->resourceName = g_res.name;
        returnInfo->implName = (char*)"oracle-CPU-4State-Double";
        returnInfo->implDescription = g_res.description;
        returnInfo->flags = g_res.supportFlags | BEAGLE_FLAG_SCALERS_RAW | BEAGLE_FLAG_THREADING_NONE;
    }
    return id;
}

int beagleFinalizeInstance(int instance) {
    Inst* I = get(instance);
    if (!I) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
    for (int i = 0; i < I->tipCount; ++i) free(I->tipStates[i]);
    for (int i = 0; i < I->tipCount + I->partialsCount; ++i) free(I->partials[i]);
    for (int i = 0; i < I->matrixCount; ++i) free(I->matrices[i]);
    for (int i = 0; i < I->scaleCount; ++i) free(I->scale[i]);
    free(I->tipStates); free(I->partials); free(I->matrices); free(I->scale);
    free(I->cmat); free(I->evals); free(I->freqs); free(I->catWeights); free(I->catRates);
    free(I->patWeights); free(I->tmpInt);
    memset(I, 0, sizeof(*I));
    return BEAGLE_SUCCESS;
}

int beagleSetTipStates(int instance, int tipIndex, const int* inStates) {
    Inst* I = get(instance);
    if (!I) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
    if (tipIndex < 0 || tipIndex >= I->tipCount || tipIndex >= I->compactCount) return BEAGLE_ERROR_OUT_OF_RANGE;
    if (!I->tipStates[tipIndex]) I->tipStates[tipIndex] = (int*)malloc((size_t)I->patternCount * sizeof(int));
    if (!I->tipStates[tipIndex]) return BEAGLE_ERROR_OUT_OF_MEMORY;
    for (int j = 0; j < I->patternCount; ++j) {
        int s = inStates[j];
        I->tipStates[tipIndex][j] = (s >= 0 && s < S) ? s : S;   /* anything else is the missing state */
    }
    return BEAGLE_SUCCESS;
}

int beagleSetPatternWeights(int instance, const double* w) {
    Inst* I = get(instance);
    if (!I) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
    memcpy(I->patWeights, w, (size_t)I->patternCount * sizeof(double));
    return BEAGLE_SUCCESS;
}

int beagleSetStateFrequencies(int instance, int idx, const double* f) {
    Inst* I = get(instance);
    if (!I) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
    if (idx < 0 || idx >= I->eigenCount) return BEAGLE_ERROR_OUT_OF_RANGE;
    memcpy(I->freqs + 4 * idx, f, 4 * sizeof(double));
    return BEAGLE_SUCCESS;
}

int beagleSetEigenDecomposition(int instance, int idx, const double* evec, const double* ivec, const double* eval) {
    Inst* I = get(instance);
    if (!I) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
    if (idx < 0 || idx >= I->eigenCount) return BEAGLE_ERROR_OUT_OF_RANGE;
    double* c = I->cmat + 64 * idx;
    int l = 0;
    for (int i = 0; i < S; ++i)
        for (int j = 0; j < S; ++j)
            for (int k = 0; k < S; ++k)
                c[l++] = evec[i * S + k] * ivec[k * S + j];
    memcpy(I->evals + 4 * idx, eval, 4 * sizeof(double));
    return BEAGLE_SUCCESS;
}

int beagleSetCategoryRates(int instance, const double* r) {
    Inst* I = get(instance);
    if (!I) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
    memcpy(I->catRates, r, (size_t)I->categoryCount * sizeof(double));
    return BEAGLE_SUCCESS;
}

int beagleSetCategoryWeights(int instance, int idx, const double* w) {
    Inst* I = get(instance);
    if (!I) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
    if (idx < 0 || idx >= I->eigenCount) return BEAGLE_ERROR_OUT_OF_RANGE;
    memcpy(I->catWeights + (size_t)idx * I->categoryCount, w, (size_t)I->categoryCount * sizeof(double));
    return BEAGLE_SUCCESS;
}

int beagleUpdateTransitionMatrices(int instance, int eigenIndex, const int* probIdx, const int* d1, const int* d2,
                                   const double* edgeLengths, int count) {
    Inst* I = get(instance);
    if (!I) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
    if (d1 || d2) return BEAGLE_ERROR_NO_IMPLEMENTATION;
    if (eigenIndex < 0 || eigenIndex >= I->eigenCount) return BEAGLE_ERROR_OUT_OF_RANGE;
    const double* c = I->cmat + 64 * eigenIndex;
    const double* ev = I->evals + 4 * eigenIndex;
    for (int u = 0; u < count; ++u) {
        if (probIdx[u] < 0 || probIdx[u] >= I->matrixCount) return BEAGLE_ERROR_OUT_OF_RANGE;
        double* tm = I->matrices[probIdx[u]];
        int n = 0;
        for (int l = 0; l < I->categoryCount; ++l) {
            double tmp[S];
            for (int i = 0; i < S; ++i) tmp[i] = exp(ev[i] * (edgeLengths[u] * I->catRates[l]));
            int m = 0;
            for (int i = 0; i < S; ++i) {
                for (int j = 0; j < S; ++j) {
                    double sum = 0.0;
                    for (int k = 0; k < S; ++k) sum += c[m++] * tmp[k];
                    tm[n++] = sum > 0 ? sum : 0;
                }
                tm[n++] = 1.0;   /* missing-state column */
            }
        }
    }
    return BEAGLE_SUCCESS;
}

int beagleResetScaleFactors(int instance, int cumIdx) {
    Inst* I = get(instance);
    if (!I) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
    if (cumIdx < 0 || cumIdx >= I->scaleCount) return BEAGLE_ERROR_OUT_OF_RANGE;
    memset(I->scale[cumIdx], 0, (size_t)I->patternCount * sizeof(double));
    return BEAGLE_SUCCESS;
}

/* One operation over the pattern range [k0,k1): child factors, product, optional rescale. */
static void do_op_range(Inst* I, const BeagleOperation* op, double* cum, int k0, int k1) {
    const int P = I->patternCount, C = I->categoryCount;
    double* dest = I->partials[op->destinationPartials];
    const double* m1 = I->matrices[op->child1TransitionMatrix];
    const double* m2 = I->matrices[op->child2TransitionMatrix];
    const int* s1 = op->child1Partials < I->tipCount ? I->tipStates[op->child1Partials] : NULL;
    const int* s2 = op->child2Partials < I->tipCount ? I->tipStates[op->child2Partials] : NULL;
    const double* p1 = I->partials[op->child1Partials];
    const double* p2 = I->partials[op->child2Partials];
    for (int l = 0; l < C; ++l) {
        const double* a = m1 + l * S * SP;
        const double* b = m2 + l * S * SP;
        for (int k = k0; k < k1; ++k) {
            size_t v = ((size_t)l * P + k) * S;
            for (int i = 0; i < S; ++i) {
                double f1, f2;
                if (s1) f1 = a[i * SP + s1[k]];
                else {
                    f1 = a[i * SP + 0] * p1[v + 0];
                    f1 += a[i * SP + 1] * p1[v + 1];
                    f1 += a[i * SP + 2] * p1[v + 2];
                    f1 += a[i * SP + 3] * p1[v + 3];
                }
                if (s2) f2 = b[i * SP + s2[k]];
                else {
                    f2 = b[i * SP + 0] * p2[v + 0];
                    f2 += b[i * SP + 1] * p2[v + 1];
                    f2 += b[i * SP + 2] * p2[v + 2];
                    f2 += b[i * SP + 3] * p2[v + 3];
                }
                dest[v + i] = f1 * f2;
            }
        }
    }
    if (op->destinationScaleWrite >= 0) {
        double* sf = I->scale[op->destinationScaleWrite];
        for (int k = k0; k < k1; ++k) {
            double mx = 0.0;
            for (int l = 0; l < C; ++l) {
                size_t v = ((size_t)l * P + k) * S;
                for (int i = 0; i < S; ++i) if (dest[v + i] > mx) mx = dest[v + i];
            }
            if (mx == 0) mx = 1.0;
            double inv = 1.0 / mx;
            for (int l = 0; l < C; ++l) {
                size_t v = ((size_t)l * P + k) * S;
                for (int i = 0; i < S; ++i) dest[v + i] *= inv;
            }
            sf[k] = mx;                       /* raw scalers (no SCALERS_LOG flag) */
            if (cum) cum[k] += log(mx);
        }
    } else if (op->destinationScaleRead >= 0) {
        const double* sf = I->scale[op->destinationScaleRead];
        for (int k = k0; k < k1; ++k)
            for (int l = 0; l < C; ++l) {
                size_t v = ((size_t)l * P + k) * S;
                for (int i = 0; i < S; ++i) dest[v + i] /= sf[k];
            }
    }
}

#define MAX_THREADS 256
typedef struct { Inst* I; const BeagleOperation* ops; int opCount; double* cum; int k0, k1; } Slice;
static void* slice_main(void* arg) {
    Slice* s = (Slice*)arg;
    for (int o = 0; o < s->opCount; ++o) do_op_range(s->I, &s->ops[o], s->cum, s->k0, s->k1);
    return NULL;
}

int beagleUpdatePartials(int instance, const BeagleOperation* ops, int opCount, int cumIdx) {
    Inst* I = get(instance);
    if (!I) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
    if (cumIdx != BEAGLE_OP_NONE && (cumIdx < 0 || cumIdx >= I->scaleCount)) return BEAGLE_ERROR_OUT_OF_RANGE;
    const int nbuf = I->tipCount + I->partialsCount;
    for (int o = 0; o < opCount; ++o) {
        const BeagleOperation* op = &ops[o];
        if (op->destinationPartials < I->tipCount || op->destinationPartials >= nbuf) return BEAGLE_ERROR_OUT_OF_RANGE;
        if (op->child1Partials < 0 || op->child1Partials >= nbuf || op->child2Partials < 0 || op->child2Partials >= nbuf)
            return BEAGLE_ERROR_OUT_OF_RANGE;
        if (op->child1Partials < I->tipCount && !I->tipStates[op->child1Partials]) return BEAGLE_ERROR_OUT_OF_RANGE;
        if (op->child2Partials < I->tipCount && !I->tipStates[op->child2Partials]) return BEAGLE_ERROR_OUT_OF_RANGE;
        if (op->child1TransitionMatrix < 0 || op->child1TransitionMatrix >= I->matrixCount ||
            op->child2TransitionMatrix < 0 || op->child2TransitionMatrix >= I->matrixCount)
            return BEAGLE_ERROR_OUT_OF_RANGE;
        if (op->destinationScaleWrite >= I->scaleCount || op->destinationScaleRead >= I->scaleCount)
            return BEAGLE_ERROR_OUT_OF_RANGE;
    }
    double* cum = cumIdx == BEAGLE_OP_NONE ? NULL : I->scale[cumIdx];
    const int P = I->patternCount;
    int T = I->threads;
    if (T > P) T = P;
    if (T > MAX_THREADS) T = MAX_THREADS;
    if (T <= 1) {
        for (int o = 0; o < opCount; ++o) do_op_range(I, &ops[o], cum, 0, P);
    } else {
        /* patterns are independent: each thread runs the whole operation list on its slice */
        pthread_t th[MAX_THREADS];
        Slice sl[MAX_THREADS];
        for (int t = 0; t < T; ++t) {
            sl[t].I = I; sl[t].ops = ops; sl[t].opCount = opCount; sl[t].cum = cum;
            sl[t].k0 = (int)((long)P * t / T); sl[t].k1 = (int)((long)P * (t + 1) / T);
            if (pthread_create(&th[t], NULL, slice_main, &sl[t]) != 0) {
                slice_main(&sl[t]);
                th[t] = 0;
                sl[t].opCount = -1;
            }
        }
        for (int t = 0; t < T; ++t) if (sl[t].opCount >= 0) pthread_join(th[t], NULL);
    }
    return BEAGLE_SUCCESS;
}

int beagleCalculateEdgeLogLikelihoods(int instance, const int* parentIdx, const int* childIdx, const int* probIdx,
                                      const int* d1, const int* d2, const int* catWIdx, const int* freqIdx,
                                      const int* cumIdx, int count, double* outSum, double* outD1, double* outD2) {
    Inst* I = get(instance);
    if (!I) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
    if (d1 || d2 || outD1 || outD2 || count != 1) return BEAGLE_ERROR_NO_IMPLEMENTATION;
    const int P = I->patternCount, C = I->categoryCount, nbuf = I->tipCount + I->partialsCount;
    int par = parentIdx[0], ch = childIdx[0], pm = probIdx[0];
    if (par < I->tipCount || par >= nbuf || ch < 0 || ch >= nbuf || pm < 0 || pm >= I->matrixCount)
        return BEAGLE_ERROR_OUT_OF_RANGE;
    if (catWIdx[0] < 0 || catWIdx[0] >= I->eigenCount || freqIdx[0] < 0 || freqIdx[0] >= I->eigenCount)
        return BEAGLE_ERROR_OUT_OF_RANGE;
    if (cumIdx[0] != BEAGLE_OP_NONE && (cumIdx[0] < 0 || cumIdx[0] >= I->scaleCount)) return BEAGLE_ERROR_OUT_OF_RANGE;
    const double* pp = I->partials[par];
    const double* tm = I->matrices[pm];
    const double* wt = I->catWeights + (size_t)catWIdx[0] * C;
    const double* fr = I->freqs + 4 * freqIdx[0];
    double* tmp = I->tmpInt;
    memset(tmp, 0, (size_t)P * S * sizeof(double));
    if (ch < I->tipCount) {
        const int* st = I->tipStates[ch];
        if (!st) return BEAGLE_ERROR_OUT_OF_RANGE;
        for (int l = 0; l < C; ++l)
            for (int k = 0; k < P; ++k) {
                size_t v = ((size_t)l * P + k) * S;
                for (int i = 0; i < S; ++i)
                    tmp[(size_t)k * S + i] += tm[l * S * SP + i * SP + st[k]] * pp[v + i] * wt[l];
            }
    } else {
        const double* pc = I->partials[ch];
        for (int l = 0; l < C; ++l)
            for (int k = 0; k < P; ++k) {
                size_t v = ((size_t)l * P + k) * S;
                for (int i = 0; i < S; ++i) {
                    double sj = 0.0;
                    for (int j = 0; j < S; ++j) sj += tm[l * S * SP + i * SP + j] * pc[v + j];
                    tmp[(size_t)k * S + i] += sj * pp[v + i] * wt[l];
                }
            }
    }
    const double* cum = cumIdx[0] == BEAGLE_OP_NONE ? NULL : I->scale[cumIdx[0]];
    double total = 0.0;
    for (int k = 0; k < P; ++k) {
        double sumI = 0.0;
        for (int i = 0; i < S; ++i) sumI += fr[i] * tmp[(size_t)k * S + i];
        double ll = log(sumI);
        if (cum) ll += cum[k];
        total += ll * I->patWeights[k];
    }
    *outSum = total;
    if (total != total) return BEAGLE_ERROR_FLOATING_POINT;
    return BEAGLE_SUCCESS;
}

/* ---- inspection hooks for per-buffer parity tests (oracle only) ---- */

/* Copies buffer `idx` as [pattern][category][state] (the product's HBM layout). */
STROM_API int oracleGetPartials(int instance, int idx, double* out) {
    Inst* I = get(instance);
    if (!I) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
    if (idx < I->tipCount || idx >= I->tipCount + I->partialsCount) return BEAGLE_ERROR_OUT_OF_RANGE;
    const int P = I->patternCount, C = I->categoryCount;
    for (int l = 0; l < C; ++l)
        for (int k = 0; k < P; ++k)
            for (int i = 0; i < S; ++i)
                out[((size_t)k * C + l) * S + i] = I->partials[idx][((size_t)l * P + k) * S + i];
    return BEAGLE_SUCCESS;
}

/* Copies scale buffer `idx` (raw per-node maxima, or cumulative log sums) */
STROM_API int oracleGetScaleFactors(int instance, int idx, double* out) {
    Inst* I = get(instance);
    if (!I) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
    if (idx < 0 || idx >= I->scaleCount) return BEAGLE_ERROR_OUT_OF_RANGE;
    memcpy(out, I->scale[idx], (size_t)I->patternCount * sizeof(double));
    return BEAGLE_SUCCESS;
}

/* Copies matrix `idx` as [category][4][4] (padding column dropped). */
STROM_API int oracleGetTransitionMatrix(int instance, int idx, double* out) {
    Inst* I = get(instance);
    if (!I) return BEAGLE_ERROR_UNINITIALIZED_INSTANCE;
    if (idx < 0 || idx >= I->matrixCount) return BEAGLE_ERROR_OUT_OF_RANGE;
    for (int l = 0; l < I->categoryCount; ++l)
        for (int i = 0; i < S; ++i)
            for (int j = 0; j < S; ++j)
                out[(l * S + i) * S + j] = I->matrices[idx][l * S * SP + i * SP + j];
    return BEAGLE_SUCCESS;
}
