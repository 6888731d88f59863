// capi.cpp -- plain C entry points over the host classes, for Python (ctypes) tests and bench.py.
// Built twice: libstrom_host.so (-DSTROM_B200_ENGINE, linked to libstrom_b200.so: the product) and,
// for the test-suite only, a variant linked to the CPU checker (same code, 13-call path).
#include <chrono>
#include <cstring>
#include <memory>
#include <mutex>

#include "likelihood.hpp"
#include "mcmc.hpp"
#include "tree_manip.hpp"

namespace strom {
struct LikelihoodInspector {
    static void operations(Likelihood& L, Tree::SharedPtr t, unsigned ntaxa, std::vector<int>& ops, std::vector<int>& pidx,
                           std::vector<double>& elen) {
        L._ntaxa = ntaxa;
        L.defineOperations(t);
        ops = L._operations;
        pidx = L._pmatrix_index;
        elen = L._edge_lengths;
    }
};
struct ModelInspector {
    static void eigen(const Model& m, double* evec, double* ivec, double* evals) {
        std::memcpy(evec, m._eigenvectors, sizeof(double) * 16);
        std::memcpy(ivec, m._inverse_eigenvectors, sizeof(double) * 16);
        std::memcpy(evals, m._eigenvalues, sizeof(double) * 4);
    }
};
}  // namespace strom

using namespace strom;

#define HOST_API extern "C" __attribute__((visibility("default")))

namespace {
std::mutex g_mu;
std::vector<Data::SharedPtr> g_data;

int fail(const std::exception& e, char* err, int errlen) {
    if (err && errlen > 0) {
        std::strncpy(err, e.what(), static_cast<size_t>(errlen) - 1);
        err[errlen - 1] = 0;
    }
    return -1;
}
Data::SharedPtr dataOf(int h) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (h < 0 || h >= static_cast<int>(g_data.size()) || !g_data[h]) throw XStrom("bad data handle");
    return g_data[h];
}
int keep(Data::SharedPtr d) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_data.push_back(d);
    return static_cast<int>(g_data.size()) - 1;
}
Model::SharedPtr makeModel(const double* freqs, const double* rmat, double shape, int ncat, double pinvar) {
    Model::SharedPtr m(new Model());
    m->setExchangeabilitiesAndStateFreqs(std::vector<double>(rmat, rmat + 6), std::vector<double>(freqs, freqs + 4));
    m->setGammaShape(shape);
    m->setGammaNCateg(static_cast<unsigned>(ncat));
    if (pinvar > 0.0) m->setPinvar(pinvar);
    return m;
}
}  // namespace

HOST_API int strom_data_from_file(const char* path, char* err, int errlen) {
    try {
        Data::SharedPtr d(new Data());
        d->getDataFromFile(path);
        return keep(d);
    } catch (const std::exception& e) { return fail(e, err, errlen); }
}

// codes: [ntaxa][nsites] ints (0..3, >= 4 missing / ambiguity)
HOST_API int strom_data_from_codes(int ntaxa, int nsites, const int* codes, int compress, char* err, int errlen) {
    try {
        Data::taxon_names_t names;
        Data::data_matrix_t m(static_cast<size_t>(ntaxa));
        for (int t = 0; t < ntaxa; ++t) {
            names.push_back("taxon" + std::to_string(t + 1));
            m[t].assign(codes + static_cast<size_t>(t) * nsites, codes + static_cast<size_t>(t + 1) * nsites);
        }
        Data::SharedPtr d(new Data());
        if (compress == 2) d->setDeviceCompression(0);        // always on the GPU (engine builds)
        else if (compress == 3) d->setDeviceCompression(~0u);  // never
        d->setDataMatrix(names, m, compress != 0);
        return keep(d);
    } catch (const std::exception& e) { return fail(e, err, errlen); }
}

HOST_API int strom_data_dims(int h, int* ntaxa, int* npatterns, int* nsites) {
    try {
        Data::SharedPtr d = dataOf(h);
        *ntaxa = static_cast<int>(d->getNumTaxa());
        *npatterns = static_cast<int>(d->getNumPatterns());
        *nsites = static_cast<int>(d->getSeqLen());
        return 0;
    } catch (...) { return -1; }
}

HOST_API int strom_data_copy(int h, int* matrix, double* counts) {
    try {
        Data::SharedPtr d = dataOf(h);
        const unsigned P = d->getNumPatterns();
        for (unsigned t = 0; t < d->getNumTaxa(); ++t)
            std::memcpy(matrix + static_cast<size_t>(t) * P, d->getDataMatrix()[t].data(), sizeof(int) * P);
        std::memcpy(counts, d->getPatternCounts().data(), sizeof(double) * P);
        return 0;
    } catch (...) { return -1; }
}

HOST_API int strom_data_free(int h) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (h < 0 || h >= static_cast<int>(g_data.size())) return -1;
    g_data[h].reset();
    return 0;
}

// Operation list, matrix indices and edge lengths Likelihood::defineOperations builds for a newick
// (ntaxa = number of leaves).  Arrays must hold 7*(n-2), 2n-3 and 2n-3 entries.  newick_out (may be NULL)
// receives TreeManip::makeNewick(6) of the re-rooted tree.
HOST_API int strom_tree_ops(const char* newick, int* ops, int* pidx, double* elen, int* nleaves, int* parent, char* newick_out,
                            int newick_len, char* err, int errlen) {
    try {
        TreeManip tm;
        tm.buildFromNewick(newick, false, false);
        Tree::SharedPtr t = tm.getTree();
        Likelihood L;
        std::vector<int> o, p;
        std::vector<double> e;
        LikelihoodInspector::operations(L, t, t->numLeaves(), o, p, e);
        if (ops) std::memcpy(ops, o.data(), sizeof(int) * o.size());
        if (pidx) std::memcpy(pidx, p.data(), sizeof(int) * p.size());
        if (elen) std::memcpy(elen, e.data(), sizeof(double) * e.size());
        *nleaves = static_cast<int>(t->numLeaves());
        *parent = 2 * static_cast<int>(t->numLeaves()) - 3;
        if (newick_out && newick_len > 0) {
            const std::string s = tm.makeNewick(6);
            std::strncpy(newick_out, s.c_str(), static_cast<size_t>(newick_len) - 1);
            newick_out[newick_len - 1] = 0;
        }
        return 0;
    } catch (const std::exception& e) { return fail(e, err, errlen); }
}

// Model numbers exactly as the facade uploads them.  rates/probs need ncat (+1 with pinvar > 0) entries.
HOST_API int strom_model_numbers(const double* freqs, const double* rmat, double shape, int ncat, double pinvar, double* evec, double* ivec,
                                 double* evals, double* rates, double* probs, int* ncat_total, char* err, int errlen) {
    try {
        Model::SharedPtr m = makeModel(freqs, rmat, shape, ncat, pinvar);
        const std::vector<double> r = m->getDiscreteGammaRelRates(), w = m->getDiscreteGammaRateProbs();
        *ncat_total = static_cast<int>(r.size());
        std::memcpy(rates, r.data(), sizeof(double) * r.size());
        std::memcpy(probs, w.data(), sizeof(double) * w.size());
        ModelInspector::eigen(*m, evec, ivec, evals);
        return 0;
    } catch (const std::exception& e) { return fail(e, err, errlen); }
}

// Builds Data/Tree/Model/Likelihood and evaluates `repeats` times (parameters re-uploaded every time,
// scalar back on the host every time).  devices: GPUs to shard the patterns over.
HOST_API int strom_loglik(int data_handle, const char* newick, const double* freqs, const double* rmat, double shape, int ncat,
                          double pinvar, int ndevices, const int* devices, int single, int repeats, double* logl,
                          double* seconds_per_eval, char* err, int errlen) {
    try {
        Data::SharedPtr d = dataOf(data_handle);
        TreeManip tm;
        tm.buildFromNewick(newick, false, false);
        Likelihood::SharedPtr L(new Likelihood());
        L->setQuiet(true);
        L->setData(d);
        L->setModel(makeModel(freqs, rmat, shape, ncat, pinvar));
        if (ndevices > 0) L->setDevices(std::vector<int>(devices, devices + ndevices));
        if (single & 1) L->setSinglePrecision(true);
        L->setTransientPartials((single & 4) == 0);              // 4 = every node's buffer written, as the engine-level default
        double v = L->calcLogLikelihood(tm.getTree());
        if (repeats > 1) {
            const auto t0 = std::chrono::steady_clock::now();
            for (int i = 1; i < repeats; ++i) v = L->calcLogLikelihood(tm.getTree());
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (seconds_per_eval) *seconds_per_eval = dt / (repeats - 1);
        } else if (seconds_per_eval) {
            *seconds_per_eval = 0.0;
        }
        *logl = v;
        return 0;
    } catch (const std::exception& e) { return fail(e, err, errlen); }
}

// A random walk of LOCAL-style tree moves (three edges rescaled, half of them with an NNI swap; half of the moves
// taken back afterwards, as a rejected proposal would be; every 16th step a new gamma shape) evaluated twice per
// step: by a Likelihood with incremental evaluation and by one that recomputes everything.  Fills, per step, both
// log-likelihoods and the number of operations the incremental one listed; returns the number of steps.
HOST_API int strom_incremental_walk(int data_handle, const char* newick, const double* freqs, const double* rmat, double shape, int ncat,
                                    double pinvar, int ndevices, const int* devices, unsigned seed, int nsteps, double* incremental_logl,
                                    double* full_logl,
                                    int* incremental_ops, double* seconds_incremental, double* seconds_full, char* err, int errlen) {
    try {
        Data::SharedPtr d = dataOf(data_handle);
        TreeManip tm;
        tm.buildFromNewick(newick, false, false);
        Likelihood::SharedPtr inc(new Likelihood()), full(new Likelihood());
        Model::SharedPtr model = makeModel(freqs, rmat, shape, ncat, pinvar);
        for (Likelihood::SharedPtr L : {inc, full}) {
            L->setQuiet(true);
            if (ndevices > 0) L->setDevices(std::vector<int>(devices, devices + ndevices));   // pattern shards
            L->setData(d);
            L->setModel(model);
        }
        inc->setIncremental(true);
        Lot lot(seed);
        double ti = 0.0, tf = 0.0;
        int step = 0;
        auto evaluate = [&]() {
            auto t0 = std::chrono::steady_clock::now();
            incremental_logl[step] = inc->calcLogLikelihood(tm.getTree());
            auto t1 = std::chrono::steady_clock::now();
            full_logl[step] = full->calcLogLikelihood(tm.getTree());
            auto t2 = std::chrono::steady_clock::now();
            if (step > 0) {
                ti += std::chrono::duration<double>(t1 - t0).count();
                tf += std::chrono::duration<double>(t2 - t1).count();
            }
            incremental_ops[step] = static_cast<int>(inc->lastOperationCount());
            ++step;
        };
        evaluate();
        while (step < nsteps) {
            if (step % 16 == 15) {
                model->setGammaShape(shape * std::exp(lot.uniform() - 0.5));
                evaluate();
                continue;
            }
            Node* x = tm.randomInternalEdge(lot.uniform());
            Node* a = x->getLeftChild();
            Node* b = a->getRightSib();
            Node* y = x->getParent();
            Node* dsib = (x == y->getLeftChild()) ? x->getRightSib() : y->getLeftChild();
            Node* top = lot.uniform() < 0.5 ? a : b;
            const double lt = top->getEdgeLength(), lm = x->getEdgeLength(), lb = dsib->getEdgeLength();
            const double m = std::exp(0.4 * (lot.uniform() - 0.5));
            top->setEdgeLength(m * lt);
            x->setEdgeLength(m * lm);
            dsib->setEdgeLength(m * lb);
            const bool swapped = lot.uniform() < 0.5;
            Node* moved = (top == a) ? b : a;
            if (swapped) tm.nniNodeSwap(moved, dsib);
            evaluate();
            if (step < nsteps && lot.uniform() < 0.5) {           // take the move back, as a rejection does
                if (swapped) tm.nniNodeSwap(dsib, moved);
                top->setEdgeLength(lt);
                x->setEdgeLength(lm);
                dsib->setEdgeLength(lb);
                evaluate();
            }
        }
        if (seconds_incremental) *seconds_incremental = ti / std::max(1, step - 1);
        if (seconds_full) *seconds_full = tf / std::max(1, step - 1);
        return step;
    } catch (const std::exception& e) { return fail(e, err, errlen); }
}

// TEST HOOK: n variates of the host's Lot from `seed`.  kind 0 = uniform, 1 = randint(lo = a, hi = b), 2 = gamma(shape a, scale b),
// 3 = logUniform; boost != 0: the Boost 1.66 restatement (mcmc.hpp), else the std stand-in.
HOST_API int strom_lot_draws(unsigned seed, int boost, int kind, double a, double b, int n, double* out) {
    Lot lot(seed);
    lot.setBoostCompatible(boost != 0);
    for (int i = 0; i < n; ++i)
        out[i] = kind == 0 ? lot.uniform() : kind == 1 ? lot.randint(static_cast<int>(a), static_cast<int>(b)) : kind == 2 ? lot.gamma(a, b) : lot.logUniform();
    return n;
}

// TEST HOOK: an incremental Likelihood is shown a sequence of trees, each built anew by
// TreeManip::buildFromNewick -- a new Tree object every round, usually at the address of the one just destroyed -- after an
// NNI swap and a change of three edge lengths; a second Likelihood recomputes everything.  worst_out receives the largest
// |incremental - full| / |full|.
HOST_API int strom_incremental_rebuild_check(int data_handle, const char* newick, const double* freqs, const double* rmat, double shape,
                                             int ncat, int rounds, unsigned seed, double* worst_out, char* err, int errlen) {
    try {
        Data::SharedPtr d = dataOf(data_handle);
        Likelihood::SharedPtr inc(new Likelihood()), full(new Likelihood());
        Model::SharedPtr model = makeModel(freqs, rmat, shape, ncat, 0.0);
        for (Likelihood::SharedPtr L : {inc, full}) {
            L->setQuiet(true);
            L->setData(d);
            L->setModel(model);
        }
        inc->setIncremental(true);
        Lot lot(seed);
        std::string current = newick;
        double worst = 0.0;
        TreeManip tm;
        for (int r = 0; r < rounds; ++r) {
            tm.buildFromNewick(current, false, false);           // the previous Tree dies here
            const double a = inc->calcLogLikelihood(tm.getTree());
            const double b = full->calcLogLikelihood(tm.getTree());
            worst = std::max(worst, std::fabs(a - b) / std::fabs(b));
            Node* x = tm.randomInternalEdge(lot.uniform());
            Node* child = x->getLeftChild();
            Node* y = x->getParent();
            Node* sib = (x == y->getLeftChild()) ? x->getRightSib() : y->getLeftChild();
            child->setEdgeLength(child->getEdgeLength() * std::exp(lot.uniform() - 0.5));
            x->setEdgeLength(x->getEdgeLength() * std::exp(lot.uniform() - 0.5));
            sib->setEdgeLength(sib->getEdgeLength() * std::exp(lot.uniform() - 0.5));
            tm.nniNodeSwap(child, sib);
            current = tm.makeNewick(12);
        }
        if (worst_out) *worst_out = worst;
        return 0;
    } catch (const std::exception& e) { return fail(e, err, errlen); }
}

// TEST HOOK for the LOCAL move (mcmc.hpp::TreeUpdater): `nsteps` proposals on the tree; each is first made and taken back
// (the tree must then be exactly the one before: same shape, same 17-digit edge lengths) and then made again and kept.  Returns the number of take-backs that did
// NOT restore the tree; counts topology-changing proposals, and the largest |focal path after / (m * before) - 1| and
// |log Hastings - (3 log m - E log(Pac m + 1 - Pac))| with m recovered from the path lengths.
HOST_API int strom_local_move_check(const char* newick, unsigned seed, double lambda, int nsteps, int* topology_changes,
                                    double* worst_scale_error, char* final_newick, int final_len, char* err, int errlen) {
    try {
        TreeManip::SharedPtr tm(new TreeManip);
        tm->buildFromNewick(newick, false, false);
        TreeUpdater up;
        up.setTreeManip(tm);
        up.setLot(Lot::SharedPtr(new Lot(seed)));
        up.setLambda(lambda);
        int bad = 0, changes = 0;
        double worst = 0.0;
        for (int i = 0; i < nsteps; ++i) {
            const std::string before = up.canonicalForTest();
            const Lot probe(seed * 7919u + static_cast<unsigned>(i));   // (copied twice: the same six deviates for both proposals)
            up.setLot(Lot::SharedPtr(new Lot(probe)));
            up.proposeForTest(true);
            if (up.canonicalForTest() != before) ++bad;
            up.setLot(Lot::SharedPtr(new Lot(probe)));          // the same deviates again: the same proposal, kept this time
            const double TL = tm->calcTreeLength();
            const double E = static_cast<double>(tm->getTree()->numNodes() - 1);
            double f0 = 0.0, f1 = 0.0, lh = 0.0;
            if (up.proposeForTest(false, &f0, &f1, &lh)) ++changes;
            const double m = f1 / f0, Pac = f0 / TL;
            worst = std::max(worst, std::fabs(lh - (3.0 * std::log(m) - E * std::log(Pac * m + 1.0 - Pac))));
            worst = std::max(worst, std::fabs(tm->calcTreeLength() - (TL - f0 + f1)) / TL);
        }
        if (topology_changes) *topology_changes = changes;
        if (worst_scale_error) *worst_scale_error = worst;
        if (final_newick && final_len > 0) std::snprintf(final_newick, static_cast<size_t>(final_len), "%s", tm->makeNewick(6).c_str());
        return bad;
    } catch (const std::exception& e) { return fail(e, err, errlen); }
}

// Strom::run's MCMC branch on the host classes.  Fills `rows` (max_rows x 15 doubles: iteration, lnL,
// lnPrior, TL, 6 exchangeabilities, 4 frequencies, alpha) with the cold chain's samples; returns the row count.
// `flags`: 1 = every chain draws from its own random stream, 2 = chains are stepped side by side on host threads
// (one heated chain per GPU; implies 1), 4 = incremental likelihood (only the changed path is re-evaluated).
// 8 = lockstep: one host thread begins every update on all chains before finishing it on any (implies 1).
// 32 = transient partials OFF (Likelihood::setTransientPartials(false): every node's buffer is written).
// 16 = lockstep WITHOUT handing the chains' evaluations to the engine as one set of launches (for comparison).
// 0 = the reference's single shared stream, chains in turn, every node recomputed on every call.
// pinvar > 0: GTR+I+G with that fixed proportion of invariable sites.
HOST_API int strom_mcmc_ex2(int data_handle, const char* newick, unsigned seed, unsigned niter, unsigned burnin, unsigned samplefreq,
                            unsigned nchains, double heatfactor, unsigned ncateg, double gammashape, double pinvar, const double* freqs,
                            const double* rmat, int ndevices, const int* devices, unsigned flags, double* rows, int max_rows, double* seconds,
                            char* err, int errlen);

HOST_API int strom_mcmc_ex(int data_handle, const char* newick, unsigned seed, unsigned niter, unsigned burnin, unsigned samplefreq,
                           unsigned nchains, double heatfactor, unsigned ncateg, double gammashape, const double* freqs, const double* rmat,
                           int ndevices, const int* devices, unsigned flags, double* rows, int max_rows, double* seconds, char* err,
                           int errlen) {
    return strom_mcmc_ex2(data_handle, newick, seed, niter, burnin, samplefreq, nchains, heatfactor, ncateg, gammashape, 0.0, freqs, rmat,
                          ndevices, devices, flags, rows, max_rows, seconds, err, errlen);
}

HOST_API int strom_mcmc_ex2(int data_handle, const char* newick, unsigned seed, unsigned niter, unsigned burnin, unsigned samplefreq,
                            unsigned nchains, double heatfactor, unsigned ncateg, double gammashape, double pinvar, const double* freqs,
                            const double* rmat, int ndevices, const int* devices, unsigned flags, double* rows, int max_rows, double* seconds,
                            char* err, int errlen) {
    try {
        McmcOptions opt;
        opt.pinvar = pinvar;
        opt.batchEvaluations = (flags & 16u) == 0;
        opt.transientPartials = (flags & 32u) == 0;                   // 32 = every node's buffer written
        Lot::defaultBoostCompatible() = (flags & 64u) == 0;         // 64 = the std:: stand-in instead of the Boost 1.66 restatement
        opt.seed = seed; opt.niter = niter; opt.burnin = burnin; opt.samplefreq = samplefreq ? samplefreq : 1; opt.nchains = nchains;
        opt.heatfactor = heatfactor; opt.ncateg = ncateg; opt.gammashape = gammashape;
        opt.statefreq.assign(freqs, freqs + 4);
        opt.rmatrix.assign(rmat, rmat + 6);
        if (ndevices > 0) opt.devices.assign(devices, devices + ndevices);
        opt.perChainStreams = (flags & 1u) != 0;
        opt.parallelChains = (flags & 2u) != 0;
        opt.incremental = (flags & 4u) != 0;
        opt.lockstepChains = (flags & 8u) != 0;
        const auto t0 = std::chrono::steady_clock::now();
        const std::vector<TraceRow> tr = runMCMC(dataOf(data_handle), newick, opt);
        // (flag 128: the iterations alone, without setting up chains, engine instances and data)
        if (seconds) *seconds = (flags & 128u) ? opt.loopSeconds : std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const int n = std::min<int>(static_cast<int>(tr.size()), max_rows);
        for (int i = 0; i < n; ++i) {
            double* r = rows + static_cast<size_t>(i) * 15;
            r[0] = tr[i].iteration; r[1] = tr[i].lnL; r[2] = tr[i].lnPrior; r[3] = tr[i].TL;
            for (int k = 0; k < 11; ++k) r[4 + k] = tr[i].params[k];
        }
        return n;
    } catch (const std::exception& e) { return fail(e, err, errlen); }
}

HOST_API int strom_mcmc(int data_handle, const char* newick, unsigned seed, unsigned niter, unsigned burnin, unsigned samplefreq,
                        unsigned nchains, double heatfactor, unsigned ncateg, double gammashape, const double* freqs, const double* rmat,
                        int ndevices, const int* devices, double* rows, int max_rows, double* seconds, char* err, int errlen) {
    return strom_mcmc_ex(data_handle, newick, seed, niter, burnin, samplefreq, nchains, heatfactor, ncateg, gammashape, freqs, rmat,
                         ndevices, devices, 0u, rows, max_rows, seconds, err, errlen);
}



// ---------------------------------------------------------------------------------------------------------------------
// One heated chain of a Metropolis-coupled run per PROCESS (BASELINE.json configs[4]: one chain per GPU, swap proposals
// between processes; reference loop src/strom.hpp:316-402).  The process steps its chain and, at a swap proposal it takes
// part in, exchanges one small vector with the partner (strom_b200/mc3.py does that with torch.distributed: NCCL over
// NVLink, or gloo).  Same chain set-up, random streams, swap schedule and decisions as runMCMC's side-by-side driver, so the
// sampled trace is the same, bit for bit.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
struct OneChain {
    Chain chain;
    McmcOptions opt;
};
std::vector<std::unique_ptr<OneChain>> g_chains;
OneChain& chainOf(int h) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (h < 0 || h >= static_cast<int>(g_chains.size()) || !g_chains[h]) throw XStrom("bad chain handle");
    return *g_chains[h];
}
}  // namespace

HOST_API int strom_mc3_chain_create(int data_handle, const char* newick, unsigned seed, unsigned chain_index, double heatfactor, unsigned ncateg,
                                    double gammashape, double pinvar, const double* freqs, const double* rmat, int device, unsigned flags,
                                    char* err, int errlen) {
    try {
        std::unique_ptr<OneChain> oc(new OneChain());
        McmcOptions& opt = oc->opt;
        opt.seed = seed; opt.heatfactor = heatfactor; opt.ncateg = ncateg; opt.gammashape = gammashape; opt.pinvar = pinvar;
        opt.statefreq.assign(freqs, freqs + 4);
        opt.rmatrix.assign(rmat, rmat + 6);
        opt.incremental = (flags & 4u) != 0;
        opt.transientPartials = (flags & 32u) == 0;
        Lot::defaultBoostCompatible() = (flags & 64u) == 0;
        setUpChain(oc->chain, dataOf(data_handle), newick, opt, chain_index, device, Lot::SharedPtr(new Lot(seed + 1000003u * (chain_index + 1))));
        std::lock_guard<std::mutex> lk(g_mu);
        g_chains.push_back(std::move(oc));
        return static_cast<int>(g_chains.size()) - 1;
    } catch (const std::exception& e) { return fail(e, err, errlen); }
}

HOST_API int strom_mc3_chain_step(int h, int iteration, char* err, int errlen) {
    try {
        chainOf(h).chain.nextStep(iteration);
        return 0;
    } catch (const std::exception& e) { return fail(e, err, errlen); }
}

HOST_API int strom_mc3_chain_stop_tuning(int h) {
    try {
        chainOf(h).chain.stopTuning();
        return 0;
    } catch (...) { return -1; }
}

// what a chain shows its swap partner: out[0] = log kernel (lnL + lnPrior), [1] = heating power, [2] = chain index,
// [3] = number of tuning parameters, [4..] = the tuning parameters.  Returns the number of doubles written (<= cap).
HOST_API int strom_mc3_chain_offer(int h, double* out, int cap, char* err, int errlen) {
    try {
        Chain& ch = chainOf(h).chain;
        const std::vector<double> lam = ch.getLambdas();
        if (cap < 4 + static_cast<int>(lam.size())) throw XStrom("offer buffer too small");
        out[0] = ch.calcLogLikelihood() + ch.calcLogJointPrior();
        out[1] = ch.getHeatingPower();
        out[2] = ch.getChainIndex();
        out[3] = static_cast<double>(lam.size());
        for (size_t q = 0; q < lam.size(); ++q) out[4 + q] = lam[q];
        return 4 + static_cast<int>(lam.size());
    } catch (const std::exception& e) { return fail(e, err, errlen); }
}

// an accepted swap: this chain takes the partner's heating power, index and tuning parameters (src/strom.hpp:389-399)
HOST_API int strom_mc3_chain_take(int h, const double* other) {
    try {
        Chain& ch = chainOf(h).chain;
        ch.setHeatingPower(other[1]);
        ch.setChainIndex(static_cast<unsigned>(other[2]));
        ch.setLambdas(std::vector<double>(other + 4, other + 4 + static_cast<int>(other[3])));
        return 0;
    } catch (...) { return -1; }
}

HOST_API double strom_mc3_chain_heat(int h) {
    try {
        return chainOf(h).chain.getHeatingPower();
    } catch (...) { return -1.0; }
}

// the sample row of this chain (15 doubles as in strom_mcmc_ex); evaluates the likelihood once more, as Strom::sample does
HOST_API int strom_mc3_chain_row(int h, unsigned iteration, double* r, char* err, int errlen) {
    try {
        const TraceRow t = traceRowOf(chainOf(h).chain, iteration);
        r[0] = t.iteration; r[1] = t.lnL; r[2] = t.lnPrior; r[3] = t.TL;
        for (int k = 0; k < 11; ++k) r[4 + k] = t.params[k];
        return 0;
    } catch (const std::exception& e) { return fail(e, err, errlen); }
}

HOST_API int strom_mc3_chain_free(int h) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (h < 0 || h >= static_cast<int>(g_chains.size())) return -1;
    g_chains[h].reset();
    return 0;
}

// the swap schedule every process computes for itself (mcmc.hpp::swapSchedule)
HOST_API int strom_mc3_schedule(unsigned seed, unsigned nchains, unsigned total, int* ci, int* cj, double* logu, unsigned flags) {
    Lot::defaultBoostCompatible() = (flags & 64u) == 0;
    std::vector<unsigned> a, b;
    std::vector<double> u;
    swapSchedule(seed, nchains, total, a, b, u);
    for (unsigned t = 0; t < total; ++t) { ci[t] = static_cast<int>(a[t]); cj[t] = static_cast<int>(b[t]); logu[t] = u[t]; }
    return 0;
}
