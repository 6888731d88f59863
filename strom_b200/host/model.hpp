// model.hpp -- GTR + discrete-gamma (+ optional invariable-sites category) model and the upload
// shims the Likelihood facade calls every evaluation (mirrors reference src/model.hpp):
//   recalcRateMatrix   :212-265  Q normalised to mean rate 1, symmetrised, eigen-decomposed, ascending
//                                eigenvalues, V = Pi^-1/2 U, V^-1 = U^T Pi^1/2, uploaded row-major
//   recalcGammaRates   :267-313  equal-probability discrete gamma(alpha, 1/alpha), mean rate per category
//   setBeagle*         :315-354  one BeagleLib-compatible call each
// Eigen and Boost.Math (the reference's dependencies for these two steps) are replaced by a cyclic
// Jacobi solver for the 4x4 symmetric matrix and by a regularised incomplete gamma + quantile below.
// The 4x4 eigen solve and the gamma quantiles stay on the host: they are microseconds of work.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "../../include/strom_beagle.h"
#include "xstrom.hpp"

namespace strom {

namespace numerics {

// Regularised lower incomplete gamma P(a, x): series for x < a+1, continued fraction otherwise.
inline double gammaP(double a, double x) {
    if (x <= 0.0) return 0.0;
    const double lg = std::lgamma(a);
    if (x < a + 1.0) {
        double term = 1.0 / a, sum = term, ap = a;
        for (int n = 0; n < 100000; ++n) {
            ap += 1.0;
            term *= x / ap;
            sum += term;
            if (std::fabs(term) < std::fabs(sum) * 1e-17) break;
        }
        return sum * std::exp(-x + a * std::log(x) - lg);
    }
    // modified Lentz for Q(a, x)
    const double tiny = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    for (int i = 1; i < 100000; ++i) {
        const double an = -static_cast<double>(i) * (static_cast<double>(i) - a);
        b += 2.0;
        d = an * d + b;
        if (std::fabs(d) < tiny) d = tiny;
        c = b + an / c;
        if (std::fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (std::fabs(del - 1.0) < 1e-16) break;
    }
    return 1.0 - std::exp(-x + a * std::log(x) - lg) * h;
}

// x with P(a, x) = p: bracket, then safeguarded Newton (density is available in closed form).
inline double gammaPinv(double a, double p) {
    if (p <= 0.0) return 0.0;
    if (p >= 1.0) return INFINITY;
    double lo = 0.0, hi = std::max(1.0, a);
    while (gammaP(a, hi) < p) hi *= 2.0;
    double x = 0.5 * (lo + hi);
    const double lg = std::lgamma(a);
    for (int it = 0; it < 200; ++it) {
        const double f = gammaP(a, x) - p;
        if (f > 0) hi = x; else lo = x;
        const double dens = std::exp(-x + (a - 1.0) * std::log(x) - lg);
        double xn = dens > 0 ? x - f / dens : 0.5 * (lo + hi);
        if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
        if (std::fabs(xn - x) <= 1e-16 * std::fabs(x)) { x = xn; break; }
        x = xn;
    }
    return x;
}

// Eigen-decomposition of a symmetric 4x4 matrix (cyclic Jacobi); eigenvalues ascending,
// eigenvectors in the columns of U (row-major storage u[i*4+j]).
inline void symmetricEigen4(const double (&a_in)[16], double (&evals)[4], double (&u)[16]) {
    double a[16];
    for (int i = 0; i < 16; ++i) { a[i] = a_in[i]; u[i] = (i % 5 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 4; ++q) off += a[p * 4 + q] * a[p * 4 + q];
        if (off < 1e-300) break;
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 4; ++q) {
                const double apq = a[p * 4 + q];
                if (std::fabs(apq) < 1e-300) continue;
                const double theta = (a[q * 4 + q] - a[p * 4 + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 4; ++k) {     // columns p, q of A
                    const double akp = a[k * 4 + p], akq = a[k * 4 + q];
                    a[k * 4 + p] = c * akp - s * akq;
                    a[k * 4 + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 4; ++k) {     // rows p, q of A
                    const double apk = a[p * 4 + k], aqk = a[q * 4 + k];
                    a[p * 4 + k] = c * apk - s * aqk;
                    a[q * 4 + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 4; ++k) {     // accumulate eigenvectors
                    const double ukp = u[k * 4 + p], ukq = u[k * 4 + q];
                    u[k * 4 + p] = c * ukp - s * ukq;
                    u[k * 4 + q] = s * ukp + c * ukq;
                }
            }
    }
    int order[4] = {0, 1, 2, 3};
    std::sort(order, order + 4, [&](int x, int y) { return a[x * 4 + x] < a[y * 4 + y]; });
    double us[16];
    for (int j = 0; j < 4; ++j) {
        evals[j] = a[order[j] * 4 + order[j]];
        for (int i = 0; i < 4; ++i) us[i * 4 + j] = u[i * 4 + order[j]];
    }
    for (int i = 0; i < 16; ++i) u[i] = us[i];
}

}  // namespace numerics

class Model {
    friend class Likelihood;
    friend struct ModelInspector;        // test hook: reads the eigen system (capi.cpp)

  public:
    Model() { clear(); }

    void useStoredData(bool using_data) { _using_data = using_data; }
    std::string describeModel() const;

    double getGammaShape() const { return _gamma_shape; }
    unsigned getGammaNCateg() const { return _num_categ; }
    double getPinvar() const { return _pinvar; }
    std::vector<double> getExchangeabilities() const { return _exchangeabilities; }
    std::vector<double> getStateFreqs() const { return _state_freqs; }
    std::vector<double> getDiscreteGammaRelRates() const { return _relative_rates; }
    std::vector<double> getDiscreteGammaCategBoundaries() const { return _categ_boundaries; }
    std::vector<double> getDiscreteGammaRateProbs() const { return _rate_probs; }

    void setGammaShape(double shape) {
        if (shape <= 0.0) throw XStrom("gamma shape must be greater than zero but the value " + std::to_string(shape) + " was supplied");
        _gamma_shape = shape;
        recalcGammaRates();
    }
    void setGammaNCateg(unsigned ncateg) {
        if (ncateg < 1)
            throw XStrom("number of categories used for among-site rate variation must be greater than zero but the value " +
                         std::to_string(ncateg) + " was supplied");
        _num_gamma_categ = ncateg;
        recalcGammaRates();
    }
    // North-star extension (the reference has no +I): a proportion `pinvar` of invariable sites is one
    // more category of rate 0; the gamma rates are divided by (1 - pinvar) so the mean rate stays 1.
    // 0 switches it off.  Changes the category count, so call before Likelihood::setModel.
    void setPinvar(double pinvar) {
        if (pinvar < 0.0 || pinvar >= 1.0) throw XStrom("proportion of invariable sites must be in [0,1)");
        _pinvar = pinvar;
        recalcGammaRates();
    }
    void setExchangeabilities(const std::vector<double>& exchangeabilities) {
        _exchangeabilities = exchangeabilities;
        recalcRateMatrix();
    }
    void setStateFreqs(const std::vector<double>& state_frequencies) {
        _state_freqs = state_frequencies;
        recalcRateMatrix();
    }
    void setExchangeabilitiesAndStateFreqs(const std::vector<double>& exchangeabilities, const std::vector<double>& state_frequencies) {
        _exchangeabilities = exchangeabilities;
        _state_freqs = state_frequencies;
        recalcRateMatrix();
    }

    std::string paramNamesAsString(const std::string& sep) const {
        return "r(A<->C)" + sep + "r(A<->G)" + sep + "r(A<->T)" + sep + "r(C<->G)" + sep + "r(C<->T)" + sep + "r(G<->T)" + sep +
               "pi(A)" + sep + "pi(C)" + sep + "pi(G)" + sep + "pi(T)" + sep + "alpha";
    }
    std::string paramValuesAsString(const std::string& sep) const {
        std::string s;
        char buf[64];
        auto add = [&](double v, bool last) { std::snprintf(buf, sizeof(buf), "%.5f", v); s += buf; if (!last) s += sep; };
        for (double r : _exchangeabilities) add(r, false);
        for (double f : _state_freqs) add(f, false);
        add(_gamma_shape, true);
        return s;
    }

    // ---- upload shims: same four calls as the reference, any BeagleLib-compatible library behind them
    int setBeagleEigenDecomposition(int beagle_instance) {
        return beagleSetEigenDecomposition(beagle_instance, 0, _eigenvectors, _inverse_eigenvectors, _eigenvalues);
    }
    int setBeagleStateFrequencies(int beagle_instance) { return beagleSetStateFrequencies(beagle_instance, 0, &_state_freqs[0]); }
    int setBeagleAmongSiteRateVariationRates(int beagle_instance) { return beagleSetCategoryRates(beagle_instance, &_relative_rates[0]); }
    int setBeagleAmongSiteRateVariationProbs(int beagle_instance) { return beagleSetCategoryWeights(beagle_instance, 0, &_rate_probs[0]); }

    typedef std::shared_ptr<Model> SharedPtr;

  private:
    void clear() {
        _using_data = true;
        _num_categ = 1;
        _num_gamma_categ = 1;
        _gamma_shape = 0.5;
        _pinvar = 0.0;
        _relative_rates.assign(1, 1.0);
        _categ_boundaries.assign(1, 0.0);
        _rate_probs.assign(1, 1.0);
        setExchangeabilitiesAndStateFreqs({1.0, 1.0, 1.0, 1.0, 1.0, 1.0}, {0.25, 0.25, 0.25, 0.25});   // JC69
    }
    void recalcRateMatrix();
    void recalcGammaRates();

    std::vector<double> _state_freqs;
    std::vector<double> _exchangeabilities;
    double _eigenvectors[16];            // row-major, as uploaded
    double _inverse_eigenvectors[16];
    double _eigenvalues[4];

    unsigned _num_categ;                 // categories the engine sees (gamma categories + 1 if pinvar > 0)
    unsigned _num_gamma_categ;
    double _gamma_shape;
    double _pinvar;
    std::vector<double> _relative_rates;
    std::vector<double> _categ_boundaries;
    std::vector<double> _rate_probs;
    bool _using_data;
};

inline void Model::recalcRateMatrix() {
    if (!_using_data) return;
    if (_state_freqs.size() != 4 || _exchangeabilities.size() != 6) throw XStrom("GTR needs 4 state frequencies and 6 exchangeabilities");
    const double* pi = &_state_freqs[0];
    // symmetric exchangeability matrix, order AC AG AT CG CT GT
    double R[4][4] = {{0}};
    int idx = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = i + 1; j < 4; ++j) R[i][j] = R[j][i] = _exchangeabilities[idx++];
    double q[4][4], rowsum[4], mean = 0.0;
    for (int i = 0; i < 4; ++i) {
        rowsum[i] = 0.0;
        for (int j = 0; j < 4; ++j) {
            q[i][j] = (i == j) ? 0.0 : R[i][j] * pi[j];
            rowsum[i] += q[i][j];
        }
        mean += pi[i] * rowsum[i];
    }
    if (!(mean > 0.0)) throw XStrom("Error in the calculation of eigenvectors and eigenvalues of the GTR rate matrix");
    const double scaling = 1.0 / mean;
    double sq[4];
    for (int i = 0; i < 4; ++i) sq[i] = std::sqrt(pi[i]);
    double S[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            const double qij = (i == j) ? -scaling * rowsum[i] : scaling * q[i][j];
            S[i * 4 + j] = sq[i] * qij / sq[j];
        }
    for (int i = 0; i < 4; ++i)
        for (int j = i + 1; j < 4; ++j) S[i * 4 + j] = S[j * 4 + i] = 0.5 * (S[i * 4 + j] + S[j * 4 + i]);
    double U[16];
    numerics::symmetricEigen4(S, _eigenvalues, U);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            _eigenvectors[i * 4 + j] = U[i * 4 + j] / sq[i];
            _inverse_eigenvectors[i * 4 + j] = U[j * 4 + i] * sq[j];
        }
    for (int k = 0; k < 4; ++k)
        if (!std::isfinite(_eigenvalues[k])) throw XStrom("Error in the calculation of eigenvectors and eigenvalues of the GTR rate matrix");
}

inline void Model::recalcGammaRates() {
    const unsigned G = _num_gamma_categ;
    std::vector<double> rates(G, 1.0), bounds(G, 0.0), probs(G, 1.0 / G);
    if (G > 1) {
        const double alpha = _gamma_shape, beta = 1.0 / _gamma_shape;
        double cum_upper = 0.0, cum_upper_plus = 0.0, upper = 0.0, cum_prob = 0.0;
        for (unsigned i = 1; i <= G; ++i) {
            const double lower = upper, cum_lower_plus = cum_upper_plus, cum_lower = cum_upper;
            cum_prob += probs[i - 1];
            if (i < G) {
                upper = numerics::gammaPinv(alpha, cum_prob) * beta;            // quantile of Gamma(alpha, scale beta)
                cum_upper_plus = numerics::gammaP(alpha + 1.0, upper / beta);   // cdf of Gamma(alpha+1, beta)
                cum_upper = numerics::gammaP(alpha, upper / beta);
            } else {
                cum_upper_plus = 1.0;
                cum_upper = 1.0;
            }
            const double numer = cum_upper_plus - cum_lower_plus, denom = cum_upper - cum_lower;
            rates[i - 1] = denom > 0.0 ? alpha * beta * numer / denom : 0.0;
            bounds[i - 1] = lower;
        }
    }
    if (_pinvar > 0.0) {
        _relative_rates.assign(1, 0.0);
        _rate_probs.assign(1, _pinvar);
        _categ_boundaries.assign(1, 0.0);
        for (unsigned i = 0; i < G; ++i) {
            _relative_rates.push_back(rates[i] / (1.0 - _pinvar));
            _rate_probs.push_back(probs[i] * (1.0 - _pinvar));
            _categ_boundaries.push_back(bounds[i]);
        }
    } else {
        _relative_rates = rates;
        _rate_probs = probs;
        _categ_boundaries = bounds;
    }
    _num_categ = static_cast<unsigned>(_relative_rates.size());
}

inline std::string Model::describeModel() const {
    char buf[256];
    std::string s = "\n----------------- Model Info ------------------\n";
    std::snprintf(buf, sizeof(buf), "State frequencies:   \n  piA = %g\n  piC = %g\n  piG = %g\n  piT = %g", _state_freqs[0], _state_freqs[1],
                  _state_freqs[2], _state_freqs[3]);
    s += buf;
    std::snprintf(buf, sizeof(buf), "\nRelative rates:    \n  rAC = %g\n  rAG = %g\n  rAT = %g\n  rCG = %g\n  rCT = %g\n  rGT = %g", _exchangeabilities[0],
                  _exchangeabilities[1], _exchangeabilities[2], _exchangeabilities[3], _exchangeabilities[4], _exchangeabilities[5]);
    s += buf;
    std::snprintf(buf, sizeof(buf), "\nRate categories:   \n  %u\nGamma shape:       \n  %g\nProportion invariable:\n  %g", _num_categ, _gamma_shape, _pinvar);
    s += buf;
    for (unsigned i = 0; i < _num_categ; ++i) {
        std::snprintf(buf, sizeof(buf), "\n%12u %12.5f %12.5f %12.5f", i + 1, _categ_boundaries[i], _relative_rates[i], _rate_probs[i]);
        s += buf;
    }
    s += "\n---------------------------------------------------\n";
    return s;
}

}  // namespace strom
