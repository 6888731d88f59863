// mcmc.hpp -- the callers of the hot path: a compact Metropolis-coupled MCMC driver over the Likelihood
// facade, used to show that the same host code produces the same trace on the CUDA engine and on the
// reference's CPU call path (north-star: "with the same seed the MCMC trace must be identical up to
// tolerance").  It follows the reference's control flow and acceptance rule:
//   Updater::update  pull -> prior -> propose -> push -> calcLogLikelihood -> prior -> heated MH test
//                    -> accept | revert+push -> tune -> reset            (reference src/updater.hpp:184-225)
//   Updater::tune    adaptive lambda, gamma_n = 10/(100+n), cap 1000     (src/updater.hpp:136-151)
//   Chain::nextStep  gamma shape (if C > 1), state freqs, exchangeabilities, tree, tree length
//                                                                        (src/chain.hpp:285-294)
//   swapChains       logR = (b_i - b_j)(k_j - k_i); heating powers, chain indices and lambdas swap
//                                                                        (src/strom.hpp:326-402)
//   heating powers   1 / (1 + heatfactor * i)                            (src/strom.hpp:202-215)
// Deliberate differences (these classes are host control flow, outside the rebuilt hot path):
//   * Lot draws from std::mt19937 / std::gamma_distribution (Boost's generators are not available), so the
//     streams differ from a real strom build; parity is defined between the two back-ends of THIS code;
//   * TreeUpdater is the reference's LOCAL move (same proposal distribution, Hastings ratio and order of the six
//     uniform deviates), written over the focal path instead of as eight cases.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <exception>
#include <functional>
#include <memory>
#include <mutex>
#include <numeric>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "likelihood.hpp"
#include "tree_manip.hpp"

namespace strom {

// The reference's Lot (src/lot.hpp) is boost::mt19937 behind Boost.Random 1.66 distributions.  boost::mt19937 and
// std::mt19937 are the same generator (MT19937, same seeding), so what has to be restated is how Boost turns its
// 32-bit words into variates.  As published in Boost 1.66:
//   * uniform_real<>(0,1) / uniform_01: ONE word, x / 2^32 (may be exactly 0; never 1);
//   * randint (src/lot.hpp:79-106): the reference's own loop over cumulative 1/n sums -- restated with the same
//     floating-point accumulation;
//   * gamma_distribution<>(alpha, beta), alpha > 1 (the only case the Dirichlet updaters produce: alpha = 1 + x/lambda,
//     src/dirichlet_updater.hpp:92): rejection from a Cauchy envelope, y = tan(pi u1), x = sqrt(2 alpha - 1) y + alpha - 1,
//     accepted when u2 <= (1 + y^2) exp((alpha - 1) log(x / (alpha - 1)) - sqrt(2 alpha - 1) y), two words per attempt;
//     alpha <= 1 needs Boost's exponential_distribution, a table-driven ziggurat since 1.56 whose tables are not available
//     here: those cases draw the exponential by inversion (same distribution, other variates).
// UNPINNED: there is no Boost in this environment to check a single variate against; setBoostCompatible(false) gives the
// round-1 stand-in ((x + 0.5) / 2^32 and std::gamma_distribution).  Both consume the stream in a fixed order, so traces
// are reproducible from the seed either way, and engine-vs-CPU trace parity does not depend on the choice.
class Lot {
  public:
    explicit Lot(unsigned seed = 1) : _gen(seed) {}
    void setSeed(unsigned seed) { _gen.seed(seed); }
    void setBoostCompatible(bool on) { _boost = on; }
    bool boostCompatible() const { return _boost; }
    static bool& defaultBoostCompatible() {
        static bool on = true;
        return on;
    }
    double uniform() {
        if (_boost) return static_cast<double>(_gen()) / 4294967296.0;      // [0,1): one 32-bit word
        return (static_cast<double>(_gen()) + 0.5) / 4294967296.0;           // (0,1)
    }
    double logUniform() {
        double u = uniform();
        while (u <= 0.0) u = uniform();         // (the reference asserts u > 0: probability 2^-32 per draw)
        return std::log(u);
    }
    double gamma(double shape, double scale) {
        if (!_boost) return std::gamma_distribution<double>(shape, scale)(_gen);
        if (shape == 1.0) return exponential() * scale;
        if (shape > 1.0) {
            const double pi = 3.14159265358979323846;
            for (;;) {
                const double y = std::tan(pi * uniform());
                const double x = std::sqrt(2.0 * shape - 1.0) * y + shape - 1.0;
                if (x <= 0.0) continue;
                if (uniform() > (1.0 + y * y) * std::exp((shape - 1.0) * std::log(x / (shape - 1.0)) - std::sqrt(2.0 * shape - 1.0) * y)) continue;
                return x * scale;
            }
        }
        const double p = std::exp(1.0) / (shape + std::exp(1.0));
        for (;;) {
            const double u = uniform();
            const double y = exponential();
            double x, q;
            if (u < p) {
                x = std::exp(-y / shape);
                q = p * std::exp(-x);
            } else {
                x = 1.0 + y;
                q = p + (1.0 - p) * std::pow(x, shape - 1.0);
            }
            if (u >= q) continue;
            return x * scale;
        }
    }
    int randint(int low, int high) {        // inclusive, from one uniform (src/lot.hpp:79-106)
        const double u = uniform();
        if (!_boost) {
            const int n = high - low + 1;
            int k = static_cast<int>(u * n);
            if (k >= n) k = n - 1;
            return low + k;
        }
        const double n = static_cast<double>(high - low + 1);
        int retval = static_cast<int>(n);
        double cumprob = 0.0;
        for (unsigned i = 0; i < n; ++i) {
            cumprob += 1.0 / n;
            if (u < cumprob) {
                retval = static_cast<int>(i);
                break;
            }
        }
        return retval + low;
    }
    typedef std::shared_ptr<Lot> SharedPtr;

  private:
    double exponential() { return -std::log(1.0 - uniform()); }     // (Boost: ziggurat; see the note above)
    std::mt19937 _gen;
    bool _boost = defaultBoostCompatible();
};

class Updater {
  public:
    virtual ~Updater() {}
    void setLikelihood(Likelihood::SharedPtr l) { _likelihood = l; }
    void setTreeManip(TreeManip::SharedPtr t) { _tree_manipulator = t; }
    void setLot(Lot::SharedPtr lot) { _lot = lot; }
    void setLambda(double l) { _lambda = l; }
    void setHeatingPower(double p) { _heating_power = p; }
    void setTuning(bool on) { _tuning = on; _naccepts = 0; _nattempts = 0; }
    void setTargetAcceptanceRate(double t) { _target_acceptance = t; }
    void setPriorParameters(const std::vector<double>& c) { _prior_parameters = c; }
    double getLambda() const { return _lambda; }
    double getAcceptPct() const { return _nattempts == 0 ? 0.0 : 100.0 * _naccepts / _nattempts; }
    const std::string& getUpdaterName() const { return _name; }

    virtual double calcLogPrior() const = 0;
    double calcLogLikelihood() const { return _likelihood->calcLogLikelihood(_tree_manipulator->getTree()); }

    double update(double prev_lnL) {
        beginUpdate(prev_lnL);
        return finishUpdate();
    }
    // The two halves of update: everything up to and including the ENQUEUE of the proposal's likelihood, and everything
    // after its value is known.  A driver may begin the same update on every chain before finishing any, so that the
    // chains' evaluations overlap on the GPU; each chain still consumes its random stream in the order update() does.
    void beginUpdate(double prev_lnL) {
        pullCurrentStateFromModel();
        _prev_lnL = prev_lnL;
        _prev_log_prior = calcLogPrior();
        proposeNewState();
        pushCurrentStateToModel();
        _likelihood->beginLogLikelihood(_tree_manipulator->getTree());
    }
    double finishUpdate() {
        double log_likelihood = _likelihood->finishLogLikelihood();
        const double log_prior = calcLogPrior();
        bool accept = true;
        if (log_prior > logMinusInfinity()) {
            const double log_diff = _log_hastings_ratio + _heating_power * ((log_likelihood + log_prior) - (_prev_lnL + _prev_log_prior));
            if (_lot->logUniform() > log_diff) accept = false;
        } else {
            accept = false;
        }
        if (accept) ++_naccepts;
        else {
            revert();
            pushCurrentStateToModel();
            log_likelihood = _prev_lnL;     // the reference does not recompute after a rejection (:216-218)
        }
        tune(accept);
        _log_hastings_ratio = 0.0;
        return log_likelihood;
    }

    // Gamma(a, b) prior on tree length, symmetric Dirichlet(c) on edge proportions (src/updater.hpp:227-268)
    double calcEdgeLengthPrior() const {
        Tree::SharedPtr tree = _tree_manipulator->getTree();
        const double TL = _tree_manipulator->calcTreeLength();
        const double n = tree->numLeaves();
        const double num_edges = 2.0 * n - (tree->isRooted() ? 2.0 : 3.0);
        const double a = _prior_parameters[0], b = _prior_parameters[1], c = _prior_parameters[2];
        double lp = (a - 1.0) * std::log(TL) - TL / b - a * std::log(b) - std::lgamma(a);
        lp += std::lgamma(num_edges * c);
        if (c != 1.0) {
            for (double len : edgeLengths()) lp += (c - 1.0) * std::log(len / TL);
            lp -= std::lgamma(c) * num_edges;
        }
        return lp;
    }

    virtual void pullCurrentStateFromModel() = 0;

  protected:
    static double logMinusInfinity() { return -1.0e300; }
    virtual void tune(bool accepted) {
        ++_nattempts;
        if (!_tuning) return;
        const double gamma_n = 10.0 / (100.0 + static_cast<double>(_nattempts));
        if (accepted) _lambda *= 1.0 + gamma_n * (1.0 - _target_acceptance) / (2.0 * _target_acceptance);
        else _lambda *= 1.0 - gamma_n * 0.5;
        if (_lambda > 1000.0) _lambda = 1000.0;
    }
    virtual void revert() = 0;
    virtual void proposeNewState() = 0;
    virtual void pushCurrentStateToModel() const = 0;
    std::vector<double> edgeLengths() const;     // defined after TreeAccess

    Lot::SharedPtr _lot;
    Likelihood::SharedPtr _likelihood;
    TreeManip::SharedPtr _tree_manipulator;
    std::string _name = "updater";
    double _lambda = 0.0001;
    double _log_hastings_ratio = 0.0;
    double _target_acceptance = 0.3;
    unsigned _naccepts = 0, _nattempts = 0;
    bool _tuning = true;
    std::vector<double> _prior_parameters;
    double _heating_power = 1.0;
    double _prev_lnL = 0.0, _prev_log_prior = 0.0;      // between beginUpdate and finishUpdate
};

// Updater is a friend of Tree and Node (like in the reference); derived classes reach them through it.
inline std::vector<double> Updater::edgeLengths() const {
    std::vector<double> v;
    for (Node* nd : _tree_manipulator->getTree()->_preorder) v.push_back(nd->_edge_length);
    return v;
}

class GammaShapeUpdater : public Updater {
  public:
    GammaShapeUpdater() { _name = "Gamma Rate Variance"; }
    double calcLogPrior() const override {
        const double a = _prior_parameters[0], b = _prior_parameters[1];
        return (a - 1.0) * std::log(_curr_point) - _curr_point / b - a * std::log(b) - std::lgamma(a);
    }
    void pullCurrentStateFromModel() override { _curr_point = _likelihood->getModel()->getGammaShape(); }

  private:
    void pushCurrentStateToModel() const override { _likelihood->getModel()->setGammaShape(_curr_point); }
    void proposeNewState() override {
        _prev_point = _curr_point;
        const double m = std::exp(_lambda * (_lot->uniform() - 0.5));
        _curr_point = m * _prev_point;
        _log_hastings_ratio = std::log(m);
    }
    void revert() override { _curr_point = _prev_point; }
    double _prev_point = 0.0, _curr_point = 0.0;
};

class DirichletUpdater : public Updater {
  public:
    double calcLogPrior() const override {
        bool flat = true, bad = false;
        double lp = 0.0, sum = 0.0;
        for (size_t i = 0; i < _curr_point.size(); ++i) {
            if (_prior_parameters[i] != 1.0) flat = false;
            if (_curr_point[i] == 0.0) bad = true;
            lp += (_prior_parameters[i] - 1.0) * std::log(_curr_point[i]) - std::lgamma(_prior_parameters[i]);
            sum += _prior_parameters[i];
        }
        if (flat) return std::lgamma(sum);
        if (bad) return logMinusInfinity();
        return lp + std::lgamma(sum);
    }

  protected:
    void proposeNewState() override {      // Dirichlet(1 + x/lambda) forward, same reverse (src/dirichlet_updater.hpp:78-142)
        const size_t dim = _curr_point.size();
        _prev_point = _curr_point;
        std::vector<double> fwd(dim), rev(dim);
        double sum_dev = 0.0;
        for (size_t i = 0; i < dim; ++i) {
            fwd[i] = std::max(1.0 + _prev_point[i] / _lambda, 1.e-12);
            _curr_point[i] = _lot->gamma(fwd[i], 1.0);
            sum_dev += _curr_point[i];
        }
        for (double& x : _curr_point) x /= sum_dev;
        double lf = std::lgamma(std::accumulate(fwd.begin(), fwd.end(), 0.0)), lr = 0.0, sr = 0.0;
        for (size_t i = 0; i < dim; ++i) {
            lf += (fwd[i] - 1.0) * std::log(_prev_point[i]) - std::lgamma(fwd[i]);
            rev[i] = 1.0 + _curr_point[i] / _lambda;
            sr += rev[i];
            lr += (rev[i] - 1.0) * std::log(_curr_point[i]) - std::lgamma(rev[i]);
        }
        _log_hastings_ratio = lr + std::lgamma(sr) - lf;
    }
    void revert() override { _curr_point = _prev_point; }
    std::vector<double> _curr_point, _prev_point;
};

class StateFreqUpdater : public DirichletUpdater {
  public:
    StateFreqUpdater() { _name = "State Frequencies"; }
    void pullCurrentStateFromModel() override { _curr_point = _likelihood->getModel()->getStateFreqs(); }

  private:
    void pushCurrentStateToModel() const override { _likelihood->getModel()->setStateFreqs(_curr_point); }
};

class ExchangeabilityUpdater : public DirichletUpdater {
  public:
    ExchangeabilityUpdater() { _name = "Exchangeabilities"; }
    void pullCurrentStateFromModel() override { _curr_point = _likelihood->getModel()->getExchangeabilities(); }

  private:
    void pushCurrentStateToModel() const override { _likelihood->getModel()->setExchangeabilities(_curr_point); }
};

class TreeLengthUpdater : public Updater {
  public:
    TreeLengthUpdater() { _name = "Tree Length"; }
    double calcLogPrior() const override { return calcEdgeLengthPrior(); }
    void pullCurrentStateFromModel() override { _curr_point = _tree_manipulator->calcTreeLength(); }

  private:
    void pushCurrentStateToModel() const override { _tree_manipulator->scaleAllEdgeLengths(_curr_point / _prev_point); }
    void proposeNewState() override {
        _prev_point = _curr_point;
        const double m = std::exp(_lambda * (_lot->uniform() - 0.5));
        _curr_point = m * _prev_point;
        _log_hastings_ratio = std::log(m);
    }
    void revert() override { std::swap(_curr_point, _prev_point); }     // so the push scales back (src/tree_length_updater.hpp:82-89)
    double _prev_point = 0.0, _curr_point = 0.0;
};

// The LOCAL move of Larget & Simon with the reference's conventions (src/tree_updater.hpp:110-437), restated over the
// three-edge focal path instead of as eight drawn cases:
//
//        a     b        x = a random internal edge's upper node, a and b its children, y its parent, d its sibling,
//         \   /         c = y's parent.  The focal path is  T - x - y - (c or d):  its TOP edge belongs to T = a or b
//          x     d      (1/2 each), its MIDDLE edge to x, its BOTTOM edge to y (towards c) or to d (1/2 each).
//           \   /       All three are multiplied by m = exp(lambda (u - 1/2)); then one of the two nodes hanging off the
//            y          path -- the other child of x (1/2) or the other neighbour of y (1/2) -- is cut off and re-attached
//            |          at a uniform point p of the rescaled path.  If it passes the far junction the topology changes
//            c          (an NNI: nniNodeSwap with d), otherwise only the three lengths change.
//
// Six uniform deviates per proposal, in the reference's order: edge, T, bottom, m, p, which node.
class TreeUpdater : public Updater {
  public:
    TreeUpdater() { _name = "Tree Topol. and Edge Lengths"; }
    double calcLogPrior() const override {
        const double n = _tree_manipulator->getTree()->numLeaves();
        const double log_num_topologies = std::lgamma(2.0 * n - 5.0 + 1.0) - (n - 3.0) * std::log(2.0) - std::lgamma(n - 3.0 + 1.0);
        return -log_num_topologies + calcEdgeLengthPrior();
    }
    void pullCurrentStateFromModel() override {}
    // (test hook) one proposal outside the MH step, taken back at once if asked; returns true if it changed the topology
    bool proposeForTest(bool take_back, double* focal_before = nullptr, double* focal_after = nullptr, double* log_hastings = nullptr) {
        proposeNewState();
        if (focal_before) *focal_before = _len_top + _len_mid + _len_bot;
        if (focal_after) *focal_after = _top->getEdgeLength() + _mid->getEdgeLength() + _bot->getEdgeLength();
        if (log_hastings) *log_hastings = _log_hastings_ratio;
        const bool changed = _swapped;
        if (take_back) revert();
        _log_hastings_ratio = 0.0;
        return changed;
    }

    // (test hook) the tree with the children of every node in a fixed order and 17-digit edge lengths: an NNI that is
    // swapped back leaves the same tree but, like the reference's nniNodeSwap, not always the same order of siblings
    std::string canonicalForTest() const { return canonical(_tree_manipulator->getTree()->_root); }

  private:
    static std::string canonical(const Node* nd) {
        char len[40];
        // (a zero length read from a tree file becomes the 1e-12 floor once a node setter touches it -- reference
        // src/node.hpp:76-79, also on the way back -- so lengths are compared at the floor)
        std::snprintf(len, sizeof(len), ":%.17g", std::max(nd->_edge_length, Node::smallestEdgeLength()));
        if (!nd->_left_child) return std::to_string(nd->_number) + len;
        std::vector<std::string> kids;
        for (const Node* c = nd->_left_child; c; c = c->_right_sib) kids.push_back(canonical(c));
        std::sort(kids.begin(), kids.end());
        std::string out = "(";
        for (size_t i = 0; i < kids.size(); ++i) out += (i ? "," : "") + kids[i];
        return out + ")" + len;      // (internal node numbers are renumbered by every refreshPreorder: left out)
    }
    void pushCurrentStateToModel() const override {}
    void proposeNewState() override {
        Node* x = _tree_manipulator->randomInternalEdge(_lot->uniform());
        Node* a = x->getLeftChild();
        Node* b = a->getRightSib();
        Node* y = x->getParent();
        _d = (x == y->getLeftChild()) ? x->getRightSib() : y->getLeftChild();
        const bool a_on_path = _lot->uniform() < 0.5;
        const bool c_on_path = _lot->uniform() < 0.5;
        _top = a_on_path ? a : b;                         // owners of the three focal edges
        _mid = x;
        _bot = c_on_path ? y : _d;
        Node* other_top = a_on_path ? b : a;
        // an NNI exchanges d with the child of x that is NOT on the path when the path runs on to c, and with the one
        // that IS when it ends at d (c cannot move: it is towards the root)
        _partner = c_on_path ? other_top : _top;
        _len_top = _top->getEdgeLength();
        _len_mid = _mid->getEdgeLength();
        _len_bot = _bot->getEdgeLength();

        const double m = std::exp(_lambda * (_lot->uniform() - 0.5));
        const double num_edges = static_cast<double>(_tree_manipulator->getTree()->numNodes() - 1);
        const double TL = _tree_manipulator->calcTreeLength();
        const double Pac = (_len_top + _len_mid + _len_bot) / TL;
        _log_hastings_ratio = 3.0 * std::log(m) - num_edges * std::log(Pac * m + 1.0 - Pac);   // GammaDir parameterisation (:160-164)

        const double t = m * _len_top, mid = m * _len_mid, bt = m * _len_bot, L = t + mid + bt;
        double p = _lot->uniform() * L;                   // new attachment point, measured from the top end
        const double eps = Node::smallestEdgeLength();
        if (p <= eps) p = eps;
        else if (L - p <= eps) p = L - eps;

        const bool move_upper = _lot->uniform() < 0.5;    // the node hanging at the top/middle junction, else the middle/bottom one
        double nt, nm, nb;
        if (move_upper) {
            _swapped = p > t + mid;                       // it slid past the lower junction
            if (_swapped) { nt = t + mid; nm = p - nt; nb = L - p; }
            else { nt = p; nm = t + mid - p; nb = bt; }
        } else {
            _swapped = p < t;                             // it slid past the upper junction
            if (_swapped) { nt = p; nm = t - p; nb = mid + bt; }
            else { nt = t; nm = p - t; nb = L - p; }
        }
        if (_swapped) _tree_manipulator->nniNodeSwap(_partner, _d);
        _top->setEdgeLength(nt);
        _mid->setEdgeLength(nm);
        _bot->setEdgeLength(nb);
    }
    void revert() override {
        if (_swapped) _tree_manipulator->nniNodeSwap(_d, _partner);   // d now hangs from x, the partner from y: swap back
        _top->setEdgeLength(_len_top);
        _mid->setEdgeLength(_len_mid);
        _bot->setEdgeLength(_len_bot);
    }
    Node *_d = nullptr, *_top = nullptr, *_mid = nullptr, *_bot = nullptr, *_partner = nullptr;
    double _len_top = 0, _len_mid = 0, _len_bot = 0;
    bool _swapped = false;
};

class Chain {
  public:
    Chain() {
        _shape_updater.reset(new GammaShapeUpdater);
        _shape_updater->setLambda(1.0);
        _shape_updater->setPriorParameters({1.0, 1.0});
        _statefreq_updater.reset(new StateFreqUpdater);
        _statefreq_updater->setLambda(0.001);
        _statefreq_updater->setPriorParameters({1.0, 1.0, 1.0, 1.0});
        _exchangeability_updater.reset(new ExchangeabilityUpdater);
        _exchangeability_updater->setLambda(0.001);
        _exchangeability_updater->setPriorParameters({1.0, 1.0, 1.0, 1.0, 1.0, 1.0});
        _tree_updater.reset(new TreeUpdater);
        _tree_updater->setLambda(0.2);
        _tree_updater->setPriorParameters({1.0, 10.0, 1.0});
        _tree_length_updater.reset(new TreeLengthUpdater);
        _tree_length_updater->setLambda(0.2);
        _tree_length_updater->setPriorParameters({1.0, 10.0, 1.0});
        for (Updater* u : all()) u->setTargetAcceptanceRate(0.3);
    }
    void setTreeFromNewick(const std::string& newick) {
        _tree_manipulator.reset(new TreeManip);
        _tree_manipulator->buildFromNewick(newick, false, false);
        for (Updater* u : all()) u->setTreeManip(_tree_manipulator);
    }
    void setLikelihood(Likelihood::SharedPtr l) { _likelihood = l; for (Updater* u : all()) u->setLikelihood(l); }
    void setLot(Lot::SharedPtr lot) { for (Updater* u : all()) u->setLot(lot); }
    void setHeatingPower(double p) { _heating_power = p; for (Updater* u : all()) u->setHeatingPower(p); }
    double getHeatingPower() const { return _heating_power; }
    void setChainIndex(unsigned i) { _chain_index = i; }
    unsigned getChainIndex() const { return _chain_index; }
    void startTuning() { for (Updater* u : all()) u->setTuning(true); }
    void stopTuning() { for (Updater* u : all()) u->setTuning(false); }
    std::vector<double> getLambdas() const {
        return {_shape_updater->getLambda(), _statefreq_updater->getLambda(), _exchangeability_updater->getLambda(),
                _tree_updater->getLambda(), _tree_length_updater->getLambda()};
    }
    void setLambdas(const std::vector<double>& v) {     // tree length gets the tree updater's lambda (src/chain.hpp:163-164)
        _shape_updater->setLambda(v[0]);
        _statefreq_updater->setLambda(v[1]);
        _exchangeability_updater->setLambda(v[2]);
        _tree_updater->setLambda(v[3]);
        _tree_length_updater->setLambda(v[3]);
    }
    std::vector<double> getAcceptPercentages() const {
        std::vector<double> v;
        for (Updater* u : all()) v.push_back(u->getAcceptPct());
        return v;
    }
    Model::SharedPtr getModel() { return _likelihood->getModel(); }
    TreeManip::SharedPtr getTreeManip() { return _tree_manipulator; }
    double getLogLikelihood() const { return _log_likelihood; }
    double calcLogLikelihood() const { return _shape_updater->calcLogLikelihood(); }
    double calcLogJointPrior() const {                   // the tree-length updater's prior is left out (src/chain.hpp:274-283)
        return _shape_updater->calcLogPrior() + _statefreq_updater->calcLogPrior() + _exchangeability_updater->calcLogPrior() +
               _tree_updater->calcLogPrior();
    }
    void start() {
        for (Updater* u : all()) u->pullCurrentStateFromModel();
        _log_likelihood = calcLogLikelihood();
    }
    void nextStep(int) {
        for (Updater* u : stepOrder()) _log_likelihood = u->update(_log_likelihood);
    }
    // nextStep one updater at a time, each in two halves (see Updater::beginUpdate): k = 0 .. numUpdates()-1
    size_t numUpdates() const { return stepOrder().size(); }
    void beginUpdate(size_t k) { stepOrder()[k]->beginUpdate(_log_likelihood); }
    void finishUpdate(size_t k) { _log_likelihood = stepOrder()[k]->finishUpdate(); }

  private:
    std::vector<Updater*> all() const {
        return {_shape_updater.get(), _statefreq_updater.get(), _exchangeability_updater.get(), _tree_updater.get(),
                _tree_length_updater.get()};
    }
    std::vector<Updater*> stepOrder() const {     // gamma shape (if C > 1), state freqs, exchangeabilities, tree, tree length
        std::vector<Updater*> v = all();
        if (_likelihood->getModel()->getGammaNCateg() <= 1) v.erase(v.begin());
        return v;
    }
    Likelihood::SharedPtr _likelihood;
    TreeManip::SharedPtr _tree_manipulator;
    std::shared_ptr<GammaShapeUpdater> _shape_updater;
    std::shared_ptr<StateFreqUpdater> _statefreq_updater;
    std::shared_ptr<ExchangeabilityUpdater> _exchangeability_updater;
    std::shared_ptr<TreeUpdater> _tree_updater;
    std::shared_ptr<TreeLengthUpdater> _tree_length_updater;
    unsigned _chain_index = 0;
    double _heating_power = 1.0;
    double _log_likelihood = 0.0;
};

// One sampled row of the cold chain: iteration, lnL, lnPrior, TL, 6 exchangeabilities, 4 frequencies, alpha
// (the columns of the reference's params.txt, src/output_manager.hpp:102-114, src/model.hpp:356-368).
struct TraceRow {
    unsigned iteration;
    double lnL, lnPrior, TL;
    double params[11];
};

struct McmcOptions {
    unsigned seed = 13579, niter = 100, burnin = 100, samplefreq = 1, nchains = 1, ncateg = 4;
    double heatfactor = 0.5, gammashape = 0.5;
    double pinvar = 0.0;                     // GTR+I+G (north-star extension: the reference has no such parameter, src/model.hpp:60-76):
                                             // a FIXED proportion of invariable sites, one more rate category; no updater moves it
    std::vector<double> statefreq = {0.25, 0.25, 0.25, 0.25};
    std::vector<double> rmatrix = {1.0 / 6, 1.0 / 6, 1.0 / 6, 1.0 / 6, 1.0 / 6, 1.0 / 6};   // on the simplex, like strom's default
    std::vector<int> devices = {0};          // chain c runs on devices[c % size]: one heated chain per GPU
    // Reference behaviour (both false): ONE random stream shared by all chains, chains stepped one after the other
    // (src/strom.hpp:316-402).  That order dependence is what forbids running the chains side by side, so:
    bool perChainStreams = false;            // chain c draws from its own stream, seeded from (seed, c); the swap
                                             // proposals keep the master stream.  Changes the trace (documented deviation).
    bool incremental = false;                // Likelihood::setIncremental: tree moves re-evaluate only the changed path (engine builds)
    bool lockstepChains = false;             // ONE host thread: every update is begun (proposed, likelihood enqueued) on all
                                             // chains before it is finished on any, so the chains' evaluations overlap on the
                                             // GPU(s) they share (implies perChainStreams; same trace as perChainStreams alone)
    bool transientPartials = true;           // engine builds: Likelihood::setTransientPartials (the facade's default; no effect with `incremental`)
    mutable double loopSeconds = 0.0;        // out: wall time of the iterations alone (chains, instances and data already set up)
    bool batchEvaluations = true;            // lockstep, engine builds: the evaluations the chains of one GPU begin together go
                                             // to the engine as ONE set of kernel launches (sb200EvalBatchAsync); same values
    bool parallelChains = false;             // every chain runs on its own host thread (implies perChainStreams); only the two
                                             // chains of a swap proposal wait for each other.  Same trace as perChainStreams
                                             // alone, whatever the thread timing
};

// A swap proposal between two chains that run on their own host threads: both post what the other needs and
// wait for each other; nobody else waits.
struct SwapMeeting {
    struct Offer {
        double k = 0.0, heat = 1.0;       // log kernel (lnL + lnPrior) and heating power of the chain
        unsigned index = 0;
        std::vector<double> lambdas;
    };
    std::mutex mu;
    std::condition_variable cv;
    int posted = 0, left = 0;
    Offer offer[2];                        // [0] = chain i of the pair, [1] = chain j
};

// One chain of a Metropolis-coupled run as Strom::initChains sets it up (src/strom.hpp:236-285): its own tree, model, Likelihood
// (engine instance on `device`) and -- unlike the reference -- its own random stream, seeded from (seed, chain).
inline void setUpChain(Chain& ch, Data::SharedPtr data, const std::string& newick, const McmcOptions& opt, unsigned c, int device,
                       Lot::SharedPtr lot) {
    ch.setTreeFromNewick(newick);
    ch.setLot(lot);
    Model::SharedPtr model(new Model());
    model->setExchangeabilitiesAndStateFreqs(opt.rmatrix, opt.statefreq);
    model->setGammaShape(opt.gammashape);
    model->setGammaNCateg(opt.ncateg);
    if (opt.pinvar > 0.0) model->setPinvar(opt.pinvar);
    Likelihood::SharedPtr lik(new Likelihood());
    lik->setQuiet(true);
    lik->setIncremental(opt.incremental);
    lik->setTransientPartials(opt.transientPartials);
    lik->setDevices(std::vector<int>(1, device));
    lik->setData(data);
    lik->setModel(model);
    ch.setLikelihood(lik);
    ch.startTuning();
    ch.setChainIndex(c);
    ch.setHeatingPower(1.0 / (1.0 + opt.heatfactor * c));
    ch.start();
}

inline TraceRow traceRowOf(Chain& ch, unsigned iteration) {
    TraceRow r;
    r.iteration = iteration;
    r.lnL = ch.calcLogLikelihood();                 // one more full evaluation per sample, as in Strom::sample (:455)
    r.lnPrior = ch.calcLogJointPrior();
    r.TL = ch.getTreeManip()->calcTreeLength();
    const std::vector<double> x = ch.getModel()->getExchangeabilities(), f = ch.getModel()->getStateFreqs();
    for (int i = 0; i < 6; ++i) r.params[i] = x[i];
    for (int i = 0; i < 4; ++i) r.params[6 + i] = f[i];
    r.params[10] = ch.getModel()->getGammaShape();
    return r;
}

// The swap schedule of a run whose chains have their own random streams: which two chains, and the log-uniform deviate of the
// acceptance test, per iteration -- drawn from the master stream in the order the in-turn loop draws them (src/strom.hpp:344-345,
// :387).  It consumes no chain state, so every process of a one-chain-per-process run computes the same schedule from the seed.
inline void swapSchedule(unsigned seed, unsigned nchains, unsigned total, std::vector<unsigned>& ci, std::vector<unsigned>& cj,
                         std::vector<double>& logu) {
    Lot lot(seed);
    ci.assign(total, 0); cj.assign(total, 0); logu.assign(total, 0.0);
    if (nchains < 2) return;
    for (unsigned t = 0; t < total; ++t) {
        ci[t] = static_cast<unsigned>(lot.randint(0, static_cast<int>(nchains) - 1));
        cj[t] = (ci[t] + 1 + static_cast<unsigned>(lot.randint(0, static_cast<int>(nchains) - 2))) % nchains;
        logu[t] = lot.logUniform();
    }
}

// Strom::run's MCMC branch (src/strom.hpp:486-523): initChains, burn-in with tuning, sampling with swaps.
inline std::vector<TraceRow> runMCMC(Data::SharedPtr data, const std::string& newick, const McmcOptions& opt) {
    Lot::SharedPtr lot(new Lot(opt.seed));
    const bool ownStreams = opt.perChainStreams || opt.parallelChains || opt.lockstepChains;
    std::vector<Chain> chains(opt.nchains);
#ifdef STROM_B200_ENGINE
    // one per GPU and half: the chains are stepped as two half-sets, so that the host prepares one half's proposals while
    // the GPU evaluates the other half's (see stepAll)
    std::map<std::pair<int, int>, EvalBatcher::SharedPtr> batchers;
#endif
    const unsigned halfSize = (opt.nchains + 1) / 2;
    for (unsigned c = 0; c < opt.nchains; ++c) {
        Chain& ch = chains[c];
        ch.setTreeFromNewick(newick);
        ch.setLot(ownStreams ? Lot::SharedPtr(new Lot(opt.seed + 1000003u * (c + 1))) : lot);
        Model::SharedPtr model(new Model());
        model->setExchangeabilitiesAndStateFreqs(opt.rmatrix, opt.statefreq);
        model->setGammaShape(opt.gammashape);
        model->setGammaNCateg(opt.ncateg);
        if (opt.pinvar > 0.0) model->setPinvar(opt.pinvar);
        Likelihood::SharedPtr lik(new Likelihood());
        lik->setQuiet(true);
        lik->setIncremental(opt.incremental);
        lik->setTransientPartials(opt.transientPartials);
        lik->setDevices(std::vector<int>(1, opt.devices[c % opt.devices.size()]));
#ifdef STROM_B200_ENGINE
        if (opt.lockstepChains && opt.batchEvaluations && !opt.parallelChains && opt.nchains > 1) {
            EvalBatcher::SharedPtr& b = batchers[std::make_pair(opt.devices[c % opt.devices.size()], c < halfSize ? 0 : 1)];
            if (!b) b.reset(new EvalBatcher());
            lik->setBatcher(b);
        }
#endif
        lik->setData(data);
        lik->setModel(model);
        if (std::getenv("STROM_HOST_NODATA")) lik->useStoredData(false);   // tuning aid: host cost of an iteration without likelihoods
        ch.setLikelihood(lik);
        ch.startTuning();
        ch.setChainIndex(c);
        ch.setHeatingPower(1.0 / (1.0 + opt.heatfactor * c));
        ch.start();
    }
    std::vector<TraceRow> trace;
    auto sample = [&](unsigned iteration) {
        for (Chain& ch : chains)
            if (ch.getHeatingPower() == 1.0) trace.push_back(traceRowOf(ch, iteration));
    };
    auto swapChains = [&]() {
        if (opt.nchains == 1) return;
        const unsigned i = static_cast<unsigned>(lot->randint(0, static_cast<int>(opt.nchains) - 1));
        const unsigned j = (i + 1 + static_cast<unsigned>(lot->randint(0, static_cast<int>(opt.nchains) - 2))) % opt.nchains;
        const double heat_i = chains[i].getHeatingPower(), heat_j = chains[j].getHeatingPower();
        const double k_i = chains[i].calcLogLikelihood() + chains[i].calcLogJointPrior();
        const double k_j = chains[j].calcLogLikelihood() + chains[j].calcLogJointPrior();
        const double logR = (heat_i - heat_j) * (k_j - k_i);
        if (lot->logUniform() < logR) {
            const unsigned idx_i = chains[i].getChainIndex(), idx_j = chains[j].getChainIndex();
            chains[j].setHeatingPower(heat_i);
            chains[i].setHeatingPower(heat_j);
            chains[j].setChainIndex(idx_i);
            chains[i].setChainIndex(idx_j);
            const std::vector<double> li = chains[i].getLambdas(), lj = chains[j].getLambdas();
            chains[i].setLambdas(lj);
            chains[j].setLambdas(li);
        }
    };
    sample(0);
    const auto loopStart = std::chrono::steady_clock::now();
    auto stopClock = [&]() { opt.loopSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - loopStart).count(); };
    if (opt.parallelChains && opt.nchains > 1) {
        // Side by side.  The swap schedule (which two chains, and the uniform deviate of the acceptance test) only
        // consumes the master stream, never a chain's state, so it is drawn up front in the order the in-turn loop
        // would draw it; a chain then only ever waits for the partner of its own swap proposals.
        struct Proposal {
            unsigned i, j;
            double logu;
        };
        const unsigned total = opt.burnin + opt.niter;
        std::vector<Proposal> proposals(total);
        for (Proposal& p : proposals) {
            p.i = static_cast<unsigned>(lot->randint(0, static_cast<int>(opt.nchains) - 1));
            p.j = (p.i + 1 + static_cast<unsigned>(lot->randint(0, static_cast<int>(opt.nchains) - 2))) % opt.nchains;
            p.logu = lot->logUniform();
        }
        std::mutex tableMutex;                                    // meetings exist from the first arrival to the second departure
        std::map<unsigned, std::shared_ptr<SwapMeeting>> meetings;
        std::atomic<bool> failed(false);
        std::vector<std::exception_ptr> errors(opt.nchains);
        std::vector<std::vector<TraceRow>> rows(opt.nchains);
        auto propose = [&](unsigned c, unsigned t) {
            const Proposal& p = proposals[t];
            if (c != p.i && c != p.j) return;
            Chain& ch = chains[c];
            const int me = c == p.i ? 0 : 1;
            std::shared_ptr<SwapMeeting> meeting;
            {
                std::lock_guard<std::mutex> lk(tableMutex);
                std::shared_ptr<SwapMeeting>& slot = meetings[t];
                if (!slot) slot.reset(new SwapMeeting());
                meeting = slot;
            }
            SwapMeeting& m = *meeting;
            SwapMeeting::Offer mine;
            mine.k = ch.calcLogLikelihood() + ch.calcLogJointPrior();
            mine.heat = ch.getHeatingPower();
            mine.index = ch.getChainIndex();
            mine.lambdas = ch.getLambdas();
            std::unique_lock<std::mutex> lk(m.mu);
            m.offer[me] = mine;
            ++m.posted;
            m.cv.notify_all();
            while (m.posted < 2 && !failed.load()) m.cv.wait_for(lk, std::chrono::milliseconds(50));
            if (m.posted < 2) throw XStrom("a chain failed while its swap partner was waiting");
            const SwapMeeting::Offer& i = m.offer[0];
            const SwapMeeting::Offer& j = m.offer[1];
            const double logR = (i.heat - j.heat) * (j.k - i.k);
            if (p.logu < logR) {
                const SwapMeeting::Offer& other = m.offer[1 - me];
                ch.setHeatingPower(other.heat);
                ch.setChainIndex(other.index);
                ch.setLambdas(other.lambdas);
            }
            const bool last = ++m.left == 2;
            lk.unlock();
            if (last) {
                std::lock_guard<std::mutex> tl(tableMutex);
                meetings.erase(t);
            }
        };
        auto life = [&](unsigned c) {
            try {
                Chain& ch = chains[c];
                for (unsigned it = 1; it <= opt.burnin; ++it) {
                    ch.nextStep(static_cast<int>(it));
                    propose(c, it - 1);
                }
                ch.stopTuning();
                for (unsigned it = 1; it <= opt.niter; ++it) {
                    ch.nextStep(static_cast<int>(it));
                    if (it % opt.samplefreq == 0 && ch.getHeatingPower() == 1.0) rows[c].push_back(traceRowOf(ch, it));
                    propose(c, opt.burnin + it - 1);
                }
            } catch (...) {
                errors[c] = std::current_exception();
                failed.store(true);                              // (waiting partners poll this flag)
            }
        };
        std::vector<std::thread> threads;
        for (unsigned c = 0; c < opt.nchains; ++c) threads.emplace_back(life, c);
        for (std::thread& t : threads) t.join();
        for (std::exception_ptr& e : errors)
            if (e) std::rethrow_exception(e);
        for (const std::vector<TraceRow>& r : rows) trace.insert(trace.end(), r.begin(), r.end());
        std::stable_sort(trace.begin() + 1, trace.end(), [](const TraceRow& x, const TraceRow& y) { return x.iteration < y.iteration; });
        stopClock();
        return trace;
    }
    // all chains one iteration on: in turn, or (lockstep) update by update with the likelihoods of all chains in flight
    auto stepAll = [&](unsigned it) {
        if (!opt.lockstepChains || opt.nchains == 1) {
            for (Chain& ch : chains) ch.nextStep(static_cast<int>(it));
            return;
        }
        const size_t nu = chains[0].numUpdates();          // (same model dimensions in every chain)
        // Two half-sets A and B, software-pipelined: while the GPU evaluates one half's proposals the host finishes and
        // proposes the other half's.  Every chain still sees its own updates in order and draws from its own stream, so
        // the trace is that of the chains stepped in turn.
        auto beginHalf = [&](int half, size_t k) {
            for (unsigned c = half ? halfSize : 0; c < (half ? opt.nchains : halfSize); ++c) chains[c].beginUpdate(k);
#ifdef STROM_B200_ENGINE
            for (auto& b : batchers)
                if (b.first.second == half) b.second->flush();       // this half's evaluations: one set of launches per GPU
#endif
        };
        auto finishHalf = [&](int half, size_t k) {
            for (unsigned c = half ? halfSize : 0; c < (half ? opt.nchains : halfSize); ++c) chains[c].finishUpdate(k);
        };
        beginHalf(0, 0);
        for (size_t k = 0; k < nu; ++k) {
            beginHalf(1, k);
            finishHalf(0, k);
            if (k + 1 < nu) beginHalf(0, k + 1);
            finishHalf(1, k);
        }
    };
    for (unsigned it = 1; it <= opt.burnin; ++it) {
        stepAll(it);
        swapChains();
    }
    for (Chain& ch : chains) ch.stopTuning();
    for (unsigned it = 1; it <= opt.niter; ++it) {
        stepAll(it);
        if (it % opt.samplefreq == 0) sample(it);
        swapChains();
    }
    stopClock();
    return trace;
}

}  // namespace strom
