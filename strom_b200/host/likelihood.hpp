// likelihood.hpp -- the Likelihood facade Chain and every Updater drive (same class surface as
// reference src/likelihood.hpp:15-63): setData / setModel / useStoredData / availableResources /
// calcLogLikelihood(Tree).  Behind it sits any library exporting the 13 BeagleLib-compatible calls
// of include/strom_beagle.h; the product links libstrom_b200.so (CUDA, no CPU fallback).
//
// What changes relative to the reference (BASELINE north-star):
//   * the instance lives on a GPU: partials, matrices, scalers and byte-coded tips stay in HBM;
//   * with STROM_B200_ENGINE defined, one evaluation is ONE fused engine call (sb200EvalAsync) per
//     pattern shard instead of 9 BeagleLib calls, and site patterns can be sharded over several
//     GPUs (setDevices): every shard holds all partials of its slice and the only exchange is the
//     sum of one double per shard, added on the host in rank order (deterministic);
//   * without that macro the facade issues exactly the reference's call sequence
//     (src/likelihood.hpp:409-447), which is how the tests run it against the CPU checker.
#pragma once
#include <algorithm>
#include <cassert>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/strom_beagle.h"
#ifdef STROM_B200_ENGINE
#include "../../include/strom_b200.h"
#endif
#include "data.hpp"
#include "model.hpp"
#include "tree.hpp"
#include "xstrom.hpp"

namespace strom {

#ifdef STROM_B200_ENGINE
// Collects the evaluations several Likelihood objects begin (the heated chains of an MC3 run that share a GPU; the
// reference steps them one after the other, src/strom.hpp:316-324) and hands them to the engine as ONE set of launches
// (sb200EvalBatchAsync) when the first of them is finished.  Single-threaded use, like the lockstep driver.
class EvalBatcher {
  public:
    typedef std::shared_ptr<EvalBatcher> SharedPtr;
    void add(int instance, const StromEvalArgs& args) {
        _instances.push_back(instance);
        _args.push_back(args);
    }
    // enqueues what has been added since the last flush; returns that call's code to every evaluation of the set
    int flush() {
        if (_instances.empty()) return _code;
        _code = sb200EvalBatchAsync(static_cast<int>(_instances.size()), &_instances[0], &_args[0]);
        _instances.clear();
        _args.clear();
        return _code;
    }

  private:
    std::vector<int> _instances;
    std::vector<StromEvalArgs> _args;
    int _code = 0;
};
#endif

class Likelihood {
    friend struct LikelihoodInspector;   // test hook: reads the operation list (capi.cpp)

  public:
    Likelihood();
    ~Likelihood();

    void useStoredData(bool using_data) { _using_data = using_data; }
    std::string availableResources();
    double calcLogLikelihood(Tree::SharedPtr t);
    // The same evaluation in two halves: begin enqueues it on the GPU(s) and returns at once, finish waits for the
    // scalar.  Between the two the caller may begin evaluations on OTHER Likelihood objects (the heated chains of an
    // MC3 run sharing a GPU), which then overlap on the device.  calcLogLikelihood = begin + finish.
    void beginLogLikelihood(Tree::SharedPtr t);
    double finishLogLikelihood();

    void setData(Data::SharedPtr d);
    Data::SharedPtr getData() { return _data; }
    void setModel(Model::SharedPtr model);
    Model::SharedPtr getModel() { return _model; }

    // ---- B200 additions -------------------------------------------------------------------
    // GPUs to shard the site patterns over (default: device 0 only).  Call before the first evaluation.
    void setDevices(const std::vector<int>& devices);
    void setSinglePrecision(bool on);         // FP32 partials / matrices variant (scalers stay FP64)
    void setQuiet(bool quiet) { _quiet = quiet; }
    unsigned numShards() const { return static_cast<unsigned>(_instances.size()); }
    // Incremental evaluation (the reference recomputes everything, src/likelihood.hpp:409-413):
    // calcLogLikelihood compares the tree and model with what the engine holds from the previous call and lists only
    // the nodes whose partials changed (the path from every touched edge to the root) and only the transition
    // matrices that changed.  A new model or a different Tree object falls back to the full list.  No updater has to
    // cooperate: a rejected move is simply the next difference.  Has no effect on the CPU call path.
    void setIncremental(bool on);
    // Engine builds: partials that only travel in registers inside a fused chain of operations are not written to HBM
    // (sb200SetTransientPartials).  ON by default behind this facade: calcLogLikelihood recomputes every node on every call
    // (as the reference does, src/likelihood.hpp:409-413) and nothing behind the facade ever reads a node's buffer back, so
    // the copies would be written for nobody.  Same log-likelihood bit for bit.  The engine itself (13 BeagleLib calls,
    // sb200Eval) keeps BeagleLib's behaviour -- every destination buffer written -- unless asked.  No effect with incremental
    // evaluation (which reads earlier calls' buffers) or on the CPU call path.
    void setTransientPartials(bool on);
#ifdef STROM_B200_ENGINE
    // Evaluations begun on this object are enqueued together with those of the other Likelihoods that share `b`
    // (one set of kernel launches for all of them) when the first one is finished.  One pattern shard only.
    void setBatcher(EvalBatcher::SharedPtr b) { _batcher = b; }
#endif
    unsigned lastOperationCount() const { return static_cast<unsigned>(_operations.size() / 7); }

    typedef std::shared_ptr<Likelihood> SharedPtr;

  private:
    void initBeagleLib();
    void finalizeInstances();
    void setTipStates();
    void setPatternWeights();
    void setDiscreteGammaShape();
    void setModelRateMatrix();
    void defineOperations(Tree::SharedPtr t);
    bool defineChangedOperations(Tree::SharedPtr t);
    void snapshotState(Tree::SharedPtr t);
    void commitState(double log_likelihood);
    void modelNumbers(std::vector<double>& v) const;
    void updateTransitionMatrices();
    void calculatePartials();
    std::string errorText(int code) const;

    std::vector<int> _instances;       // one engine instance per pattern shard
    std::vector<int> _devices;
    std::map<int, std::string> _beagle_error;
    std::vector<int> _operations;
    std::vector<int> _pmatrix_index;
    std::vector<double> _edge_lengths;

    Data::SharedPtr _data;
    Model::SharedPtr _model;
    unsigned _ntaxa;
    unsigned _nstates;
    unsigned _npatterns;
    bool _rooted;
    bool _prefer_gpu;
    bool _single;
    bool _quiet;
    bool _using_data;
    bool _pending = false, _pending_on_device = false;   // an evaluation was begun / its scalar is still on the device
    double _pending_value = 0.0;
    unsigned _instance_categ = 0;        // category count the instances were created with
#ifdef STROM_B200_ENGINE
    StromEvalArgs _args;                 // of the evaluation in flight (its arrays are members of this object and the model)
    EvalBatcher::SharedPtr _batcher;
    bool _pending_in_batch = false;
#endif

    // what the engine holds from the previous call (incremental evaluation)
    bool _incremental;
    bool _transient = true;
    bool _remembered;
    unsigned long long _held_tree_uid;   // Tree::uid() of the tree the device state belongs to (0: none)
    std::vector<int> _buffer_of;         // by node slot (index in Tree::_nodes): buffer = matrix index, frozen at the first call
    std::vector<int> _held_left, _held_right;   // by buffer: buffers of the two children the partials were computed from
    std::vector<double> _held_length;    // by matrix index: edge length its transition matrices were computed from
    std::vector<double> _held_model;
    double _held_log_likelihood;
    std::vector<char> _changed;          // scratch, by buffer
    // the same four as the evaluation in flight will leave them: taken when it is begun (the caller may edit the tree or
    // the model before finishing), committed when it has succeeded
    std::vector<int> _next_left, _next_right;
    std::vector<double> _next_length, _next_model;
    unsigned long long _next_tree_uid = 0;
};

inline Likelihood::Likelihood()
    : _ntaxa(0), _nstates(0), _npatterns(0), _rooted(false), _prefer_gpu(true), _single(false), _quiet(false), _using_data(true),
      _incremental(false), _remembered(false), _held_tree_uid(0), _held_log_likelihood(0.0) {
    _model = Model::SharedPtr(new Model());
    _devices.assign(1, 0);
    // BeagleLib error codes, so that useful messages reach the user (reference src/likelihood.hpp:79-87)
    _beagle_error[0] = "success";
    _beagle_error[-1] = "unspecified error";
    _beagle_error[-2] = "not enough memory could be allocated";
    _beagle_error[-3] = "unspecified exception";
    _beagle_error[-4] = "the instance index is out of range, or the instance has not been created";
    _beagle_error[-5] = "one of the indices specified exceeded the range of the array";
    _beagle_error[-6] = "no resource matches requirements";
    _beagle_error[-7] = "no implementation matches requirements";
    _beagle_error[-8] = "floating-point range exceeded";
}

inline std::string Likelihood::errorText(int code) const {
    std::map<int, std::string>::const_iterator it = _beagle_error.find(code);
    return "BeagleLib error code was " + std::to_string(code) + " (" + (it == _beagle_error.end() ? std::string("?") : it->second) + ")";
}

inline void Likelihood::finalizeInstances() {
    _remembered = false;
    for (int inst : _instances) {
        const int code = beagleFinalizeInstance(inst);
        if (code != 0) throw XStrom("Likelihood failed to finalize BeagleLib instance. " + errorText(code));
    }
    _instances.clear();
}

inline Likelihood::~Likelihood() {
    for (int inst : _instances) {
        const int code = beagleFinalizeInstance(inst);
        if (code != 0) std::cerr << "Likelihood destructor failed to finalize BeagleLib instance. " << errorText(code) << std::endl;
    }
}

inline std::string Likelihood::availableResources() {
    BeagleResourceList* rsrcList = beagleGetResourceList();
    std::string s;
    for (int i = 0; i < rsrcList->length; ++i) {
        std::string desc = rsrcList->list[i].description ? rsrcList->list[i].description : "";
        while (!desc.empty() && desc.back() == ' ') desc.pop_back();
        s += std::to_string(i) + ": " + rsrcList->list[i].name;
        if (!desc.empty()) s += " (" + desc + ")";
        s += "\n";
    }
    return s;
}

inline void Likelihood::setData(Data::SharedPtr data) {
    _data = data;
    if (!_instances.empty()) {       // instance exists: re-create it for the new data (reference :124-138)
        finalizeInstances();
        initBeagleLib();
    }
}

inline void Likelihood::setModel(Model::SharedPtr model) {
    _model = model;
    if (!_instances.empty()) {       // category count is fixed at creation (reference :140-153, :206)
        finalizeInstances();
        initBeagleLib();
    }
}

inline void Likelihood::setDevices(const std::vector<int>& devices) {
    if (devices.empty()) throw XStrom("setDevices needs at least one device");
    _devices = devices;
    if (!_instances.empty()) {
        finalizeInstances();
        initBeagleLib();
    }
}

inline void Likelihood::setSinglePrecision(bool on) {
    _single = on;
    if (!_instances.empty()) {
        finalizeInstances();
        initBeagleLib();
    }
}

inline void Likelihood::initBeagleLib() {
    if (!_instances.empty()) return;     // no-op once created
    assert(_data);
    _ntaxa = _data->getNumTaxa();
    _npatterns = _data->getNumPatterns();
    _nstates = 4;
    _rooted = false;
    if (!_quiet) {
        std::cout << "Sequence length:    " << _data->getSeqLen() << std::endl;
        std::cout << "Number of taxa:     " << _ntaxa << std::endl;
        std::cout << "Number of patterns: " << _npatterns << std::endl;
    }
    const unsigned num_internals = _rooted ? (_ntaxa - 1) : (_ntaxa - 2);
    const unsigned num_transition_probs = _rooted ? (2 * _ntaxa - 2) : (2 * _ntaxa - 3);

    long requirementFlags = (_single ? BEAGLE_FLAG_PRECISION_SINGLE : BEAGLE_FLAG_PRECISION_DOUBLE) | BEAGLE_FLAG_SCALING_MANUAL;
    long preferenceFlags = _prefer_gpu ? BEAGLE_FLAG_PROCESSOR_GPU : BEAGLE_FLAG_PROCESSOR_CPU;

    _instance_categ = static_cast<unsigned>(_model->_relative_rates.size());
    const unsigned nshards = std::min<unsigned>(static_cast<unsigned>(_devices.size()), std::max(1u, _npatterns));
    for (unsigned r = 0; r < nshards; ++r) {
        unsigned lo, hi;
        Data::sliceBounds(_npatterns, r, nshards, lo, hi);
        int resource = _devices[r];
        BeagleInstanceDetails details;
        const int inst = beagleCreateInstance(_ntaxa,                 // tips
                                              num_internals,          // partials
                                              _ntaxa,                 // sequences
                                              _nstates,               // states
                                              hi - lo,                // patterns (this shard's slice)
                                              1,                      // models
                                              num_transition_probs,   // transition matrices
                                              static_cast<int>(_instance_categ),   // rate categories (+1 with invariable sites)
                                              num_internals + 1,      // scale buffers
                                              &resource, 1,           // device of this shard
                                              preferenceFlags, requirementFlags, &details);
        if (inst < 0) {
            for (int i : _instances) beagleFinalizeInstance(i);
            _instances.clear();
            throw XStrom("Likelihood init function failed to create Likelihood instance (" + errorText(inst) + ")");
        }
        _instances.push_back(inst);
#ifdef STROM_B200_ENGINE
        if (_incremental) {
            const int code = sb200SetSubtreeScalers(inst, 1);
            if (code != 0) throw XStrom("failed to switch incremental evaluation. " + errorText(code));
        }
        if (_transient) sb200SetTransientPartials(inst, 1);
#endif
    }
    setTipStates();
    setPatternWeights();
}

inline void Likelihood::setTipStates() {
    assert(_data);
    const unsigned nshards = numShards();
    for (unsigned r = 0; r < nshards; ++r) {
        unsigned lo, hi;
        Data::sliceBounds(_npatterns, r, nshards, lo, hi);
#ifdef STROM_B200_ENGINE
        // every taxon's 1-byte codes of this shard's slice in one strided upload (the reference uploads 4-byte ints
        // taxon by taxon, src/likelihood.hpp:228-247)
        std::vector<unsigned char> all;
        all.reserve(static_cast<size_t>(_ntaxa) * (hi - lo));
        for (unsigned t = 0; t < _ntaxa; ++t) {
            const std::vector<unsigned char> codes = _data->packTipCodes(t, lo, hi);
            all.insert(all.end(), codes.begin(), codes.end());
        }
        const int code = sb200SetAllTipStatesU8(_instances[r], all.data());
        if (code != 0) throw XStrom("failed to set tip states (" + errorText(code) + ")");
#else
        for (unsigned t = 0; t < _ntaxa; ++t) {
            const Data::pattern_t& row = _data->getDataMatrix()[t];
            const int code = beagleSetTipStates(_instances[r], static_cast<int>(t), &row[lo]);
            if (code != 0)
                throw XStrom("failed to set tip state for taxon " + std::to_string(t + 1) + " (\"" + _data->getTaxonNames()[t] + "\"; " +
                             errorText(code) + ")");
        }
#endif
    }
}

inline void Likelihood::setPatternWeights() {
    assert(_data);
    const Data::pattern_counts_t& v = _data->getPatternCounts();
    if (v.empty()) throw XStrom("failed to set pattern weights because data matrix has empty pattern count vector");
    const unsigned nshards = numShards();
    for (unsigned r = 0; r < nshards; ++r) {
        unsigned lo, hi;
        Data::sliceBounds(_npatterns, r, nshards, lo, hi);
        const int code = beagleSetPatternWeights(_instances[r], &v[lo]);
        if (code != 0) throw XStrom("failed to set pattern weights. " + errorText(code));
    }
}

inline void Likelihood::setDiscreteGammaShape() {
    for (int inst : _instances) {
        int code = _model->setBeagleAmongSiteRateVariationRates(inst);
        if (code != 0) throw XStrom("failed to set category rates. " + errorText(code));
        code = _model->setBeagleAmongSiteRateVariationProbs(inst);
        if (code != 0) throw XStrom("failed to set category probabilities. " + errorText(code));
    }
}

inline void Likelihood::setModelRateMatrix() {
    for (int inst : _instances) {
        int code = _model->setBeagleStateFrequencies(inst);
        if (code != 0) throw XStrom("failed to set state frequencies. " + errorText(code));
        code = _model->setBeagleEigenDecomposition(inst);
        if (code != 0) throw XStrom("failed to set eigen decomposition. " + errorText(code));
        code = _model->setBeagleAmongSiteRateVariationRates(inst);
        if (code != 0) throw XStrom("failed to set among-site rate variation rates. " + errorText(code));
        code = _model->setBeagleAmongSiteRateVariationProbs(inst);
        if (code != 0) throw XStrom("failed to set among-site rate variation weights. " + errorText(code));
    }
}

// Children before parents (reverse level order), one 7-int operation per internal node; the matrix of
// every node's edge is indexed by the node's number, except that the last entry -- preorder[0], whose
// number 2n-3 is out of range -- uses the root tip's number (reference src/likelihood.hpp:296-355).
inline void Likelihood::defineOperations(Tree::SharedPtr t) {
    // (filled through raw pointers into vectors sized once: this runs on every call, before the GPU has anything
    // to do, and 9000 push_backs cost more than the engine's whole enqueue on a 738-taxon tree)
    const size_t nn = t->_levelorder.size();
    _pmatrix_index.resize(nn);
    _edge_lengths.resize(nn);
    _operations.resize(7 * nn);
    int* pm = nn ? &_pmatrix_index[0] : nullptr;
    double* el = nn ? &_edge_lengths[0] : nullptr;
    int* op = nn ? &_operations[0] : nullptr;
    const int ntaxa = static_cast<int>(_ntaxa);
    for (size_t i = nn; i-- > 0;) {
        const Node* nd = t->_levelorder[i];
        assert(nd->_number >= 0);
        *pm++ = nd->_number;
        *el++ = nd->_edge_length;
        if (nd->_left_child) {
            assert(nd->_left_child->_right_sib);
            const int l = nd->_left_child->_number, r = nd->_left_child->_right_sib->_number;
            op[0] = nd->_number;                 // destination partials
            op[1] = nd->_number - ntaxa + 1;     // scaling buffer to write
            op[2] = BEAGLE_OP_NONE;              // scaling buffer to read
            op[3] = l;                           // left child partials
            op[4] = l;                           // left child transition matrix
            op[5] = r;                           // right child partials (binary tree)
            op[6] = r;                           // right child transition matrix
            op += 7;
        }
    }
    _operations.resize(nn ? static_cast<size_t>(op - &_operations[0]) : 0);
    if (!t->_is_rooted && nn) _pmatrix_index[nn - 1] = t->_root->_number;
}

inline void Likelihood::setTransientPartials(bool on) {
    _transient = on;
#ifdef STROM_B200_ENGINE
    for (int inst : _instances) {
        const int code = sb200SetTransientPartials(inst, on ? 1 : 0);
        if (code != 0) throw XStrom("failed to switch transient partials. " + errorText(code));
    }
#endif
}

inline void Likelihood::setIncremental(bool on) {
#ifdef STROM_B200_ENGINE
    if (on == _incremental) return;
    _incremental = on;
    _remembered = false;
    for (int inst : _instances) {
        const int code = sb200SetSubtreeScalers(inst, on ? 1 : 0);
        if (code != 0) throw XStrom("failed to switch incremental evaluation. " + errorText(code));
    }
#else
    (void)on;
#endif
}

// the numbers the engine turns into transition matrices and root frequencies
inline void Likelihood::modelNumbers(std::vector<double>& v) const {
    v.clear();
    v.insert(v.end(), _model->_eigenvectors, _model->_eigenvectors + 16);
    v.insert(v.end(), _model->_inverse_eigenvectors, _model->_inverse_eigenvectors + 16);
    v.insert(v.end(), _model->_eigenvalues, _model->_eigenvalues + 4);
    v.insert(v.end(), _model->_state_freqs.begin(), _model->_state_freqs.end());
    v.insert(v.end(), _model->_relative_rates.begin(), _model->_relative_rates.end());
    v.insert(v.end(), _model->_rate_probs.begin(), _model->_rate_probs.end());
}

// Like defineOperations, but lists only what differs from the remembered state.  Buffers are tied to node OBJECTS
// (refreshPreorder renumbers the internal nodes after every topology change, src/tree_manip.hpp, which would make
// every number-keyed buffer look stale).  Returns false if nothing at all changed.
inline bool Likelihood::defineChangedOperations(Tree::SharedPtr t) {
    _operations.clear();
    _pmatrix_index.clear();
    _edge_lengths.clear();
    const Node* base = &t->_nodes[0];
    std::vector<double> model;
    modelNumbers(model);
    const bool fresh = !_remembered || _held_tree_uid != t->uid() || model != _held_model || _buffer_of.size() != t->_nodes.size();
    if (fresh) {
        _buffer_of.assign(t->_nodes.size(), -1);
        for (Node* nd : t->_levelorder) {
            // tips keep their taxon number, internal nodes get the partials buffers ntaxa .. 2 ntaxa - 3
            const bool tip = nd->_left_child == nullptr;
            if (nd->_number < 0 || (tip ? nd->_number >= static_cast<int>(_ntaxa)
                                        : (nd->_number < static_cast<int>(_ntaxa) || nd->_number > 2 * static_cast<int>(_ntaxa) - 3)))
                throw XStrom("tree node numbers do not match the data (tips 0..n-1, internal nodes n..2n-3)");
            _buffer_of[nd - base] = nd->_number;
        }
        const size_t nbuf = 2 * static_cast<size_t>(_ntaxa);
        _held_left.assign(nbuf, -1);
        _held_right.assign(nbuf, -1);
        _held_length.assign(nbuf, -1.0);
    }
    _changed.assign(2 * static_cast<size_t>(_ntaxa), 0);      // 1: edge above changed, 2: partials must be recomputed
    Node* top = t->_preorder[0];
    const int root_matrix = t->_root->_number;
    for (size_t i = t->_levelorder.size(); i-- > 0;) {
        Node* nd = t->_levelorder[i];
        const int buf = _buffer_of[nd - base];
        const int mat = (nd == top) ? root_matrix : buf;
        if (fresh || nd->_edge_length != _held_length[mat]) {
            _pmatrix_index.push_back(mat);
            _edge_lengths.push_back(nd->_edge_length);
            _changed[buf] |= 1;
        }
        if (nd->_left_child) {
            const int l = _buffer_of[nd->_left_child - base], r = _buffer_of[nd->_left_child->_right_sib - base];
            if (fresh || l != _held_left[buf] || r != _held_right[buf] || _changed[l] || _changed[r]) _changed[buf] |= 2;
        }
    }
    if (_pmatrix_index.empty()) {
        bool any = false;
        for (char c : _changed) any = any || (c & 2);
        if (!any) return false;
    }
    _changed[_buffer_of[top - base]] |= 2;                       // the root's partials are always listed (they carry the scaler)
    for (size_t i = t->_levelorder.size(); i-- > 0;) {
        Node* nd = t->_levelorder[i];
        const int buf = _buffer_of[nd - base];
        if (!nd->_left_child || !(_changed[buf] & 2)) continue;
        const int l = _buffer_of[nd->_left_child - base], r = _buffer_of[nd->_left_child->_right_sib - base];
        _operations.push_back(buf);
        _operations.push_back(buf - static_cast<int>(_ntaxa) + 1);
        _operations.push_back(BEAGLE_OP_NONE);
        _operations.push_back(l);
        _operations.push_back(l);
        _operations.push_back(r);
        _operations.push_back(r);
    }
    return true;
}

// The state the evaluation being begun will leave on the device (read from the tree and model as they are NOW).
inline void Likelihood::snapshotState(Tree::SharedPtr t) {
    const Node* base = &t->_nodes[0];
    Node* top = t->_preorder[0];
    _next_left = _held_left;
    _next_right = _held_right;
    _next_length = _held_length;
    for (Node* nd : t->_levelorder) {
        const int buf = _buffer_of[nd - base];
        _next_length[(nd == top) ? t->_root->_number : buf] = nd->_edge_length;
        if (nd->_left_child) {
            _next_left[buf] = _buffer_of[nd->_left_child - base];
            _next_right[buf] = _buffer_of[nd->_left_child->_right_sib - base];
        }
    }
    modelNumbers(_next_model);
    _next_tree_uid = t->uid();
}

inline void Likelihood::commitState(double log_likelihood) {
    _held_left.swap(_next_left);
    _held_right.swap(_next_right);
    _held_length.swap(_next_length);
    _held_model.swap(_next_model);
    _held_tree_uid = _next_tree_uid;
    _held_log_likelihood = log_likelihood;
    _remembered = true;
}

inline void Likelihood::updateTransitionMatrices() {
    for (int inst : _instances) {
        const int code = beagleUpdateTransitionMatrices(inst, 0, &_pmatrix_index[0], NULL, NULL, &_edge_lengths[0],
                                                        static_cast<int>(_pmatrix_index.size()));
        if (code != 0) throw XStrom("failed to update transition matrices. " + errorText(code));
    }
}

inline void Likelihood::calculatePartials() {
    const int totalOperations = static_cast<int>(_operations.size() / 7);
    for (int inst : _instances) {
        int code = beagleResetScaleFactors(inst, 0);
        if (code != 0) throw XStrom("failed to reset scale factors in calculatePartials. " + errorText(code));
        code = beagleUpdatePartials(inst, reinterpret_cast<BeagleOperation*>(&_operations[0]), totalOperations, 0);
        if (code != 0) throw XStrom("failed to update partials. " + errorText(code));
    }
}

inline double Likelihood::calcLogLikelihood(Tree::SharedPtr t) {
    beginLogLikelihood(t);
    return finishLogLikelihood();
}

inline void Likelihood::beginLogLikelihood(Tree::SharedPtr t) {
    if (_pending) throw XStrom("beginLogLikelihood called twice without finishLogLikelihood");
    // (_pending is set on every normal way out, never when an exception leaves this function)
    _pending_on_device = false;
    _pending_value = 0.0;
    if (!_using_data) {
        _pending = true;
        return;
    }
    if (t->_is_rooted) throw XStrom("can only compute likelihoods for unrooted trees currently");
    if (!_data) throw XStrom("must call setData before calcLogLikelihood");

    if (!_instances.empty() && _instance_categ != _model->_relative_rates.size()) {
        finalizeInstances();   // the (shared) model's category count changed since the instances were created
    }
    initBeagleLib();   // no-op if valid instances already exist

    // "root" is leaf 0 and has exactly one child
    assert(t->_root->_number == 0 && t->_root->_left_child == t->_preorder[0] && !t->_preorder[0]->_right_sib);

    if (_incremental) {
        if (!defineChangedOperations(t)) {
            _pending_value = _held_log_likelihood;
            _pending = true;
            return;
        }
        _remembered = false;                             // until this call has succeeded
        snapshotState(t);
    } else {
        defineOperations(t);
    }
    assert(_pmatrix_index.size() == _edge_lengths.size());
    int index_focal_child = t->_root->_number;           // the root tip
    int index_focal_parent = _incremental ? _buffer_of[t->_preorder[0] - &t->_nodes[0]] : t->_preorder[0]->_number;   // its only child

#ifdef STROM_B200_ENGINE
    // one fused call per shard (asynchronous); finishLogLikelihood adds the scalars in shard order
    StromEvalArgs& args = _args;
    args.stateFrequencies = &_model->_state_freqs[0];
    args.eigenVectors = _model->_eigenvectors;
    args.inverseEigenVectors = _model->_inverse_eigenvectors;
    args.eigenValues = _model->_eigenvalues;
    args.categoryRates = &_model->_relative_rates[0];
    args.categoryWeights = &_model->_rate_probs[0];
    args.categoryCount = static_cast<int>(_model->_relative_rates.size());
    args.matrixIndices = _pmatrix_index.empty() ? nullptr : &_pmatrix_index[0];
    args.edgeLengths = _edge_lengths.empty() ? nullptr : &_edge_lengths[0];
    args.edgeCount = static_cast<int>(_pmatrix_index.size());
    args.operations = _operations.empty() ? nullptr : reinterpret_cast<const BeagleOperation*>(&_operations[0]);
    args.operationCount = static_cast<int>(_operations.size() / 7);
    args.rootParentBuffer = index_focal_parent;
    args.rootChildBuffer = index_focal_child;
    args.rootMatrix = index_focal_child;
    _pending_in_batch = false;
    if (_batcher && _instances.size() == 1) {
        _batcher->add(_instances[0], args);              // enqueued with the other chains' when the first one is finished
        _pending_in_batch = true;
    } else {
        for (int inst : _instances) {
            const int code = sb200EvalAsync(inst, &args, NULL);
            if (code != 0) throw XStrom("failed to evaluate the likelihood on the engine. " + errorText(code));
        }
    }
    _pending_on_device = true;
    _pending = true;
#else
    // the reference's sequence of BeagleLib calls; the last one is synchronous, so the value is known here
    double log_likelihood = 0.0;
    setModelRateMatrix();
    setDiscreteGammaShape();
    updateTransitionMatrices();
    calculatePartials();
    int stateFrequencyIndex = 0, categoryWeightsIndex = 0, cumulativeScalingIndex = 0;
    for (int inst : _instances) {
        double part = 0.0;
        const int code = beagleCalculateEdgeLogLikelihoods(inst, &index_focal_parent, &index_focal_child, &index_focal_child, NULL, NULL,
                                                           &categoryWeightsIndex, &stateFrequencyIndex, &cumulativeScalingIndex, 1, &part,
                                                           NULL, NULL);
        if (code != 0) throw XStrom("failed to calculate edge logLikelihoods in CalcLogLikelihood. " + errorText(code));
        log_likelihood += part;
    }
    _pending_value = log_likelihood;
    _pending = true;
#endif
}

inline double Likelihood::finishLogLikelihood() {
    if (!_pending) throw XStrom("finishLogLikelihood called without beginLogLikelihood");
    _pending = false;
#ifdef STROM_B200_ENGINE
    if (_pending_on_device) {
        _pending_on_device = false;
        if (_pending_in_batch) {
            _pending_in_batch = false;
            const int code = _batcher->flush();
            if (code != 0) throw XStrom("failed to evaluate the likelihood on the engine. " + errorText(code));
        }
        double log_likelihood = 0.0;
        int failed = 0;
        for (int inst : _instances) {                    // (every shard is drained even if one reports an error)
            double part = 0.0;
            const int code = sb200Finish(inst, &part);
            if (code != 0 && failed == 0) failed = code;
            log_likelihood += part;
        }
        if (failed != 0) throw XStrom("failed to calculate edge logLikelihoods in CalcLogLikelihood. " + errorText(failed));
        if (_incremental) commitState(log_likelihood);
        _pending_value = log_likelihood;
    }
#endif
    return _pending_value;
}

}  // namespace strom
