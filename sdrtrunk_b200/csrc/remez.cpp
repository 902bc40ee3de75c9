// Host-side, run-once: the equiripple (Parks-McClellan) low-pass designer the reference's decoders use for their
// baseband filters, so that a GPU-backed decoder runs the SAME taps as the Java one instead of somebody else's Remez.
// Product code, independent of oracle/.  Follows
//   FIRFilterSpecification.lowPassBuilder()...build()   J/dsp/filter/fir/FIRFilterSpecification.java:381-428, 205-296, 909-937, 991-1135
//   Grid                                                J/dsp/filter/fir/remez/Grid.java:25-78
//   RemezFIRFilterDesigner                              J/dsp/filter/fir/remez/RemezFIRFilterDesigner.java:62-98, 146-232, 250-608
//   FilterFactory.getTaps                               J/dsp/filter/FilterFactory.java:671-681
// The designer is an iteration on doubles whose result is cast to float: every expression keeps the Java's operand
// order so that the taps agree to the last float bit.
#include <cmath>
#include <string>
#include <vector>

#include "../../include/sdrgpu.h"

namespace sdrgpu {
extern thread_local std::string g_last_error;
}

namespace {

constexpr double kPi = 3.14159265358979323846;

struct Band {
    double lo, hi, level, ripple_db;
    int points = 0;
    double width() const { return hi - lo; }
    // FrequencyBand.getRippleAmplitude
    double ripple() const { return (std::pow(10.0, (ripple_db / 20)) - 1) / (std::pow(10.0, (ripple_db / 20)) + 1); }
};

class LowPassRemez {
public:
    LowPassRemez(double fs, double pass_end, double stop_start, double pass_db, double stop_db, int order, int odd, int density)
        : density_(density)
    {
        // LowPassBuilder.build: estimate a missing order, force its parity if a length parity was requested
        if (order < 6) order = sdrgpu_design_remez_estimate_order(fs, pass_end, stop_start, pass_db, stop_db);
        if (odd > 0) {
            symmetric_odd_length_ = true;
            order += order % 2;
        } else if (odd == 0) {
            symmetric_odd_length_ = false;
            order += (order % 2 == 0 ? 1 : 0);
        } else {
            symmetric_odd_length_ = order % 2 == 0;
        }
        order_ = order;
        bands_.push_back(Band{0 / fs, pass_end / fs, 1.0, pass_db});
        bands_.push_back(Band{stop_start / fs, (double)(int)(fs / 2) / fs, 0.0, stop_db});
        // FIRFilterSpecification.updateGridSize: the dense grid is shared out by band width, rounded up per band
        const int nominal = (extrema() - 1) * density_ + 1;
        double total = 0.0;
        for (const Band &b : bands_) total += b.width();
        total_width_ = total;
        for (Band &b : bands_) {
            const int g = (int)std::ceil((double)nominal * (b.width() / total));
            b.points = g > 1 ? g : 1;
        }
    }

    int length() const { return order_ + 1; }

    // RemezFIRFilterDesigner.design + getImpulseResponse; false where FilterFactory.getTaps returns null
    bool run(float *taps)
    {
        build_grid();
        const int want = extrema();
        picks_.clear();
        for (int i = 0; i < want; i++) picks_.push_back(i * density_);
        bool converged = false;
        for (int iteration = 0; iteration < 40 && !converged; iteration++) {
            fit();
            for (size_t i = 0; i < cosine_.size(); i++) error_[i] = weight_[i] * (target_[i] - response_[i]);
            if (!exchange()) return false;
            double worst = std::fabs(error_[picks_[0]]);
            for (size_t i = 1; i < picks_.size(); i++) {
                const double current = std::fabs(error_[picks_[i]]);
                if (current > worst) worst = current;
            }
            converged = worst - std::fabs(delta_) < 0.0001;
        }
        if (!converged) return false;
        fit();
        // resample(): the polynomial at ceil(L / 2) points, L = the length made odd
        int odd_len = length();
        if (odd_len % 2 == 0) odd_len--;
        const double half = (double)odd_len / 2.0;
        std::vector<double> fr((size_t)std::ceil(half));
        for (size_t x = 0; x < fr.size(); x++) fr[x] = evaluate(std::cos(kPi * (double)x / half));
        // getImpulseResponseDoubles: inverse cosine series, then (float)
        const int n_taps = length();
        const double two_pi = 2.0 * kPi;
        if (symmetric_odd_length_) {
            const double M = ((double)n_taps - 1.0) / 2.0;
            for (int n = 0; n < n_taps; n++) {
                double acc = fr[0];
                const double frequency = two_pi * (n - M) / n_taps;
                for (int k = 1; k <= M; k++) acc += 2.0 * fr[k] * std::cos(frequency * (double)k);
                taps[n] = (float)(acc / (double)n_taps);
            }
        } else {
            const double offset = (double)(n_taps - 1) / 2.0;
            for (int n = 0; n < n_taps; n++) {
                double acc = fr[0];
                const double frequency = two_pi * ((double)n - offset) / (double)n_taps;
                for (size_t k = 1; k < fr.size(); k++) acc += 2.0 * fr[k] * std::cos(frequency * (double)k);
                taps[n] = (float)(acc / (double)n_taps);
            }
        }
        return true;
    }

private:
    int extrema() const { return (symmetric_odd_length_ ? order_ / 2 : (order_ - 1) / 2) + 2; }

    // Grid.create: per band a linear ramp of frequencies from its lower edge in steps of the grid interval
    void build_grid()
    {
        int size = 0;
        for (const Band &b : bands_) size += b.points;
        const double step = total_width_ / (double)(size - (int)bands_.size());
        double strongest = 0.0;
        for (const Band &b : bands_)
            if (b.ripple() > strongest) strongest = b.ripple();
        cosine_.clear();
        target_.clear();
        weight_.clear();
        for (const Band &b : bands_) {
            double f = b.lo;
            for (int i = 0; i < b.points; i++) {
                target_.push_back(b.level);
                weight_.push_back(1.0 / (b.ripple() / strongest));
                cosine_.push_back(std::cos(2.0 * kPi * f));   // (Grid pins the band's last FREQUENCY to its edge, not this cosine)
                f += step;
            }
        }
        response_.assign(cosine_.size(), 0.0);
        error_.assign(cosine_.size(), 0.0);
    }

    // barycentric Lagrange form through the first L + 1 extremal points (Oppenheim / Schafer eq. 116a)
    double evaluate(double c) const
    {
        double num = 0.0, den = 0.0;
        for (size_t k = 0; k + 1 < picks_.size(); k++) {
            const double gap = c - cosine_[picks_[k]];
            if (std::fabs(gap) < 1.0e-7) return ordinate_[k];
            const double q = d_[k] / gap;
            num += q * ordinate_[k];
            den += q;
        }
        return num / den;
    }

    // calculateB / calculateDelta / calculateC / calculateD / updateGridFrequencyResponse
    void fit()
    {
        const size_t n = picks_.size();
        std::vector<double> b(n);
        for (size_t k = 0; k < n; k++) {
            b[k] = 1.0;
            const double xk = cosine_[picks_[k]];
            for (size_t i = 0; i < n; i++) {
                if (i == k) continue;
                double gap = xk - cosine_[picks_[i]];
                if (std::fabs(gap) < 0.00001) gap = 0.00001;
                b[k] *= 1.0 / gap;
            }
        }
        double num = 0.0, den = 0.0, sign = 1.0;
        for (size_t k = 0; k < n; k++) {
            num += (b[k] * target_[picks_[k]]);
            den += b[k] * sign / weight_[picks_[k]];
            sign = -sign;
        }
        delta_ = num / den;
        ordinate_.assign((size_t)extrema() - 1, 0.0);
        sign = 1.0;
        for (size_t k = 0; k < ordinate_.size() && k < n; k++) {
            ordinate_[k] = target_[picks_[k]] - (sign * delta_ / weight_[picks_[k]]);
            sign = -sign;
        }
        d_.assign(n - 1, 0.0);
        for (size_t k = 0; k + 1 < n; k++) d_[k] = b[k] * (cosine_[picks_[k]] - cosine_[picks_[n - 1]]);
        for (size_t i = 0; i < cosine_.size(); i++) response_[i] = evaluate(cosine_[i]);
    }

    bool reaches_delta(double v) const { return std::fabs(v) - std::fabs(delta_) > -1.0e-5; }

    // findExtremalIndices: local extrema of the weighted error that reach |delta|, thinned to one per excursion
    bool exchange()
    {
        const std::vector<double> &e = error_;
        const int n = (int)e.size(), want = extrema();
        std::vector<int> found;
        if (((e[0] > 0.0 && e[0] > e[1]) || (e[0] < 0.0 && e[0] < e[1])) && reaches_delta(e[0])) found.push_back(0);
        for (int x = 1; x < n - 1; x++) {
            const bool peak = e[x] > 0.0 && (e[x - 1] <= e[x] && e[x] > e[x + 1]);
            const bool trough = e[x] < 0.0 && (e[x - 1] >= e[x] && e[x] < e[x + 1]);
            if ((peak || trough) && reaches_delta(e[x])) found.push_back(x);
        }
        const int last = n - 1;
        if (((e[last] > 0.0 && (e[last] > e[last - 1])) || (e[last] < 0.0 && (e[last] < e[last - 1]))) && reaches_delta(e[last]))
            found.push_back(last);
        if ((int)found.size() < want) return false;
        // walk the list keeping, of neighbours on the same side of zero, the larger one (ties keep the earlier)
        std::vector<int> kept;
        int champion = found[0];
        bool positive = e[champion] > 0.0;
        for (size_t j = 1; j < found.size(); j++) {
            const int candidate = found[j];
            if ((e[candidate] > 0.0) == positive) {
                if (std::fabs(e[candidate]) > std::fabs(e[champion])) champion = candidate;   // the old one is dropped
            } else {
                kept.push_back(champion);
                champion = candidate;
                positive = !positive;
            }
        }
        kept.push_back(champion);
        if ((int)kept.size() > want) kept.resize((size_t)want);
        picks_ = kept;
        return (int)picks_.size() >= want;
    }

    int order_ = 0, density_ = 16;
    bool symmetric_odd_length_ = true;
    double total_width_ = 0.0, delta_ = 0.0;
    std::vector<Band> bands_;
    std::vector<double> cosine_, target_, weight_, response_, error_, d_, ordinate_;
    std::vector<int> picks_;
};

}  // namespace

extern "C" {

int sdrgpu_design_remez_estimate_order(double sample_rate, double frequency1, double frequency2, double pass_ripple_db,
                                       double stop_ripple_db)
{
    // Herrmann / Rabiner / Chan length estimate as FIRFilterSpecification.estimateFilterOrder writes it
    const double df = std::fabs(frequency2 - frequency1) / sample_rate;
    const double ddp = std::log10(std::fmax(stop_ripple_db, pass_ripple_db));
    const double dds = std::log10(std::fmin(stop_ripple_db, pass_ripple_db));
    const double a1 = 5.309e-3, a2 = 7.114e-2, a3 = -4.761e-1, a4 = -2.66e-3, a5 = -5.941e-1, a6 = -4.278e-1;
    const double b1 = 11.01217, b2 = 0.5124401;
    const double t1 = a1 * ddp * ddp, t2 = a2 * ddp, t3 = a4 * ddp * ddp, t4 = a5 * ddp;
    const double dinf = ((t1 + t2 + a3) * dds) + (t3 + t4 + a6);
    const double ff = b1 + b2 * (ddp - dds);
    const double n = dinf / df - ff * df + 1.0;
    return (int)std::ceil(n);
}

sdrgpu_status sdrgpu_design_remez_low_pass(double sample_rate, double pass_band_end, double stop_band_start,
                                           double pass_ripple_db, double stop_ripple_db, int order, int odd_length,
                                           int grid_density, float *out, int capacity, int *n_taps)
{
    if (!out || !n_taps) {
        sdrgpu::g_last_error = "NULL argument";
        return SDRGPU_ERR_INVALID_ARG;
    }
    if (!(sample_rate > 0.0) || !(pass_band_end > 0.0) || !(stop_band_start > pass_band_end) || stop_band_start > sample_rate / 2 ||
        !(pass_ripple_db > 0.0) || !(stop_ripple_db > 0.0) || grid_density < 1) {
        sdrgpu::g_last_error = "low-pass specification: 0 < pass band end < stop band start <= sample rate / 2, ripples > 0";
        return SDRGPU_ERR_INVALID_ARG;
    }
    LowPassRemez designer(sample_rate, pass_band_end, stop_band_start, pass_ripple_db, stop_ripple_db, order, odd_length,
                          grid_density);
    *n_taps = designer.length();
    if (designer.length() > capacity) {
        sdrgpu::g_last_error = "filter of " + std::to_string(designer.length()) + " taps exceeds the capacity " + std::to_string(capacity);
        return SDRGPU_ERR_INVALID_ARG;
    }
    if (!designer.run(out)) {
        // FilterFactory.getTaps returns null / the decoders log "Couldn't design ... filter"
        sdrgpu::g_last_error = "Can't create filter from specification - failed to converge";
        return SDRGPU_ERR_DESIGN;
    }
    return SDRGPU_OK;
}

}  // extern "C"
