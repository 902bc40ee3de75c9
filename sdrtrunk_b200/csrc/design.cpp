// Host-side, run-once helpers of libsdrgpu: prototype filter designers and the channel index calculator.
// Product code (independent of oracle/): follows J/dsp/filter/FilterFactory.java, J/dsp/filter/Window.java and
// J/dsp/filter/channelizer/ChannelCalculator.java of the reference.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sdrgpu.h"

namespace sdrgpu {
extern thread_local std::string g_last_error;
}

namespace {

constexpr double kPi = 3.14159265358979323846;

sdrgpu_status fail(sdrgpu_status code, const std::string &msg)
{
    sdrgpu::g_last_error = msg;
    return code;
}

// Window.java:386-401
double bessel_i0(double x)
{
    double factorial = 1, quarter_sq = x * x * 0.25, power = quarter_sq, sum = 1 + quarter_sq;
    for (int i = 2; i < 100; i++) {
        factorial *= i;
        power *= quarter_sq;
        double term = power / (factorial * factorial);
        sum += term;
        if (term < 1e-20) break;
    }
    return sum;
}

// Window.java:343-378 (attenuation > 50 dB branch and the two others)
std::vector<double> kaiser_window(int length, double attenuation)
{
    double beta = 0.0;
    if (attenuation > 50.0) beta = 0.1102 * (attenuation - 8.7);
    else if (attenuation >= 21.0) beta = (0.5842 * std::pow(attenuation - 21.0, 0.4)) + (0.07886 * (attenuation - 21.0));
    const double norm = bessel_i0(beta);
    std::vector<double> w(length);
    for (int x = 0; x < length; x++) {
        double arg = beta * std::sqrt(1.0 - std::pow(2.0 * x / (length - 1) - 1.0, 2));
        w[x] = bessel_i0(arg) / norm;
    }
    return w;
}

// FilterFactory.java:970-998
bool kaiser_sinc(int length, double cutoff, double attenuation, std::vector<float> &taps)
{
    if (length % 2 == 0) return false;
    taps.assign(length, 0.0f);
    const int half = length / 2;
    std::vector<double> w = kaiser_window(length, attenuation);
    const double scalor = 2.0 * cutoff, pi_scalor = kPi * scalor;
    taps[half] = (float)(1.0 * scalor * w[half]);
    for (int x = 1; x <= half; x++) {
        double a = pi_scalor * x;
        double c = scalor * std::sin(a) / a;
        c *= w[half + x];
        taps[half + x] = (float)c;
        taps[half - x] = (float)c;
    }
    return true;
}

// FilterFactory.java:690-714: magnitude response in dB, narrowed to float like decibel()
double response_db(const std::vector<float> &taps, double frequency)
{
    double re = 0.0, im = 0.0;
    for (size_t x = 0; x < taps.size(); x++) {
        re += taps[x] * std::cos(kPi * frequency * (double)x);
        im += taps[x] * std::sin(kPi * frequency * (double)x);
    }
    return (float)(10.0 * std::log10(std::pow(re, 2.0) + std::pow(im, 2.0)));
}

bool meets_objective(double db) { return std::fabs(db - (-6.020599842071533)) <= 0.0003; }

// ---------------------------------------------------------------- ChannelCalculator
struct Calculator {
    double fs, center;
    int count;
    double bw() const { return fs / (double)count; }
    double half() const { return bw() / 2.0; }
    int wrap() const { return count / 2; }
    int norm(int i) const
    {
        while (i < 0) i += count;
        while (i >= count) i -= count;
        return i;
    }
    // ChannelCalculator.java:397-426 ; positive = IndexBoundaryPolicy.ADJUST_POSITIVE
    double index_center(int i, bool positive) const
    {
        if (i == wrap()) return positive ? center + (i * bw()) : center - (i * bw());
        if (i < wrap()) return center + (i * bw());
        return center - ((count - i) * bw());
    }
    // :437-467
    double index_min(int i, bool positive) const
    {
        if (i == wrap()) return positive ? center + ((double)i * bw()) - half() : index_center(i, positive);
        if (i <= wrap()) return center + ((double)i * bw()) - half();
        return center - ((double)(count - i) * bw()) - half();
    }
    // :478-508
    double index_max(int i, bool positive) const
    {
        if (i == wrap()) return positive ? index_center(i, positive) : center - ((double)i * bw() - half());
        if (i <= wrap()) return center + ((double)i * bw()) + half();
        return center - ((double)(count - i) * bw()) + half();
    }
    // :343-370
    bool overlap(long long f, int a, int b) const
    {
        a = norm(a);
        b = norm(b);
        if (norm(b - a) != 1) return false;
        long long a_max = (long long)index_max(a, true);
        long long b_min = (long long)index_min(b, false);
        if (a == wrap()) a_max = (long long)index_max(a, false);
        if (b == wrap()) b_min = (long long)index_min(b, true);
        return f == a_max && f == b_min;
    }
    // :293-330
    int index_for(long long f, bool positive) const
    {
        double offset = f - center;
        if (std::fabs(offset) < half()) return 0;
        offset += (offset > 0) ? half() : -half();
        int idx = (int)(offset / bw());
        if (idx < 0) idx += count;
        if (positive && overlap(f, idx, idx + 1)) idx = norm(idx + 1);
        else if (!positive && overlap(f, idx - 1, idx)) idx = norm(idx - 1);
        return idx;
    }
};

}  // namespace

extern "C" {

sdrgpu_status sdrgpu_design_sinc_m2_channelizer(double channel_bandwidth, int channels, int taps_per_channel,
                                                float *out, int capacity, int *n_taps)
{
    if (!out || !n_taps || channels <= 0 || taps_per_channel <= 0)
        return fail(SDRGPU_ERR_INVALID_ARG, "bad channelizer design arguments");
    // FilterFactory.java:808-920
    int tpc = taps_per_channel;
    int length = channels * tpc - 1;
    const double fs = channel_bandwidth * channels;
    const double band_edge = channel_bandwidth / fs;
    double cutoff = band_edge / 2.0;
    double step = cutoff * 0.1;
    const double min_step = 1.0 / fs;
    std::vector<float> taps, trial;
    kaiser_sinc(length, cutoff, 80.0, taps);
    double db = response_db(taps, band_edge);
    while (step > min_step) {
        const bool ok = meets_objective(db);
        if (ok && (cutoff + step <= band_edge)) {
            kaiser_sinc(length, cutoff + step, 80.0, trial);
            double trial_db = response_db(trial, band_edge);
            if (meets_objective(trial_db)) {
                cutoff += step;
                taps.swap(trial);
                db = trial_db;
            } else {
                step /= 2.0;
            }
        } else if (ok) {
            step /= 2.0;
        } else {
            cutoff -= step;
            if (cutoff <= 0) {
                if (++tpc > taps_per_channel + 10)
                    return fail(SDRGPU_ERR_DESIGN, "Couldn't design filter with taps per channel count in the range of " +
                                                       std::to_string(taps_per_channel) + " - " +
                                                       std::to_string(taps_per_channel + 10));
                length = channels * tpc - 1;
                cutoff = channel_bandwidth / fs;
                step = cutoff * 0.1;
            }
            kaiser_sinc(length, cutoff, 80.0, taps);
            db = response_db(taps, band_edge);
        }
    }
    if (!meets_objective(db)) return fail(SDRGPU_ERR_DESIGN, "Cannot design filter to specifications");
    if (length + 1 > capacity) return fail(SDRGPU_ERR_INVALID_ARG, "output capacity too small");
    out[0] = 0.0f;  // odd-length filter, one zero coefficient prepended
    std::memcpy(out + 1, taps.data(), sizeof(float) * (size_t)length);
    *n_taps = length + 1;
    return SDRGPU_OK;
}

sdrgpu_status sdrgpu_design_sinc_m2_synthesizer(double channel_sample_rate, double channel_bandwidth, int channels,
                                                int taps_per_channel, float *out, int capacity, int *n_taps)
{
    // FilterFactory.java:755-770
    if (!out || !n_taps) return fail(SDRGPU_ERR_INVALID_ARG, "NULL output");
    int length = channels * taps_per_channel - 1;
    if (length + 1 > capacity) return fail(SDRGPU_ERR_INVALID_ARG, "output capacity too small");
    double cutoff = (channel_bandwidth * 1.10) / (channel_sample_rate * (double)channels);
    std::vector<float> taps;
    if (!kaiser_sinc(length, cutoff, 80.0, taps)) return fail(SDRGPU_ERR_DESIGN, "Sinc filters must be odd-length");
    out[0] = 0.0f;
    std::memcpy(out + 1, taps.data(), sizeof(float) * (size_t)length);
    *n_taps = length + 1;
    return SDRGPU_OK;
}

sdrgpu_status sdrgpu_design_half_band(int length, int window, float *out)
{
    // FilterFactory.java:1007-1036 ; Window.java:133-150 (Blackman), :277-287 (Hamming)
    if (!out || length < 3 || (length - 3) % 4 != 0)
        return fail(SDRGPU_ERR_INVALID_ARG, "Half Band filter length (N) must be an odd length N=4m+3");
    std::vector<double> w(length);
    if (window == SDRGPU_WINDOW_BLACKMAN) {
        const double d = length - 1, a0 = 0.426590713672, a1 = 0.496560619089, a2 = 0.0768486672399;
        for (int x = 0; x < length; x++)
            w[x] = a0 - (a1 * std::cos((2.0 * kPi * (double)x) / d)) + (a2 * std::cos((4.0 * kPi * (double)x) / d));
    } else if (window == SDRGPU_WINDOW_HAMMING) {
        for (int x = 0; x < length; x++) w[x] = 0.54 - (0.46 * std::cos((2.0 * kPi * x) / (length - 1)));
    } else {
        return fail(SDRGPU_ERR_INVALID_ARG, "unsupported window type");
    }
    const int half = length / 2;
    for (int x = 0; x < length; x++) {
        int offset = x - half;
        if (offset == 0) out[x] = 0.5f;
        else if ((x % 2) == 0) out[x] = (float)((std::sin(offset * kPi / 2) / (offset * kPi)) * w[x]);
        else out[x] = 0.0f;
    }
    return SDRGPU_OK;
}

int sdrgpu_channel_count_for_rate(double sample_rate)
{
    // ComplexPolyphaseChannelizerM2.java:148-161 (25 kHz minimum channel bandwidth)
    int channels = (int)(sample_rate / 25000);
    if (channels % 2 != 0) channels--;
    return channels;
}

sdrgpu_status sdrgpu_channel_indexes(double sample_rate, int channel_count, double center_frequency,
                                     long long channel_frequency, int channel_bandwidth, int *indexes, int capacity,
                                     int *n_indexes)
{
    // ChannelCalculator.java:223-281 ; TunerChannel.java:52-60
    if (!indexes || !n_indexes || channel_count <= 0) return fail(SDRGPU_ERR_INVALID_ARG, "bad arguments");
    Calculator c{sample_rate, center_frequency, channel_count};
    const long long lo = channel_frequency - (channel_bandwidth / 2), hi = channel_frequency + (channel_bandwidth / 2);
    const long long min_f = (long long)(center_frequency - (sample_rate / 2.0));
    const long long max_f = (long long)(center_frequency + (sample_rate / 2.0));
    if (lo < min_f || hi > max_f)
        return fail(SDRGPU_ERR_INVALID_ARG, "Requested channel cannot be provided by this channelizer");
    const int first = c.index_for(lo, true), last = c.index_for(hi, false);
    if (first == c.wrap() && last == c.wrap())
        return fail(SDRGPU_ERR_INVALID_ARG,
                    "Requested tuner channel cannot be provided.  Requested bandwidth is within two channel "
                    "bandwidths of the sample rate.");
    std::vector<int> list{first};
    if (first != last) {
        if (first < 0 || first >= channel_count || last < 0 || last >= channel_count)
            return fail(SDRGPU_ERR_INVALID_ARG, "Something went wrong while calculating the polyphase channel indexes");
        for (int p = c.norm(first + 1); p != last; p = c.norm(p + 1)) list.push_back(p);
        list.push_back(last);
    }
    if ((int)list.size() > capacity) return fail(SDRGPU_ERR_INVALID_ARG, "index capacity too small");
    for (size_t i = 0; i < list.size(); i++) indexes[i] = list[i];
    *n_indexes = (int)list.size();
    return SDRGPU_OK;
}

sdrgpu_status sdrgpu_center_frequency_for_indexes(double sample_rate, int channel_count, double center_frequency,
                                                  const int *indexes, int n_indexes, long long *frequency)
{
    // ChannelCalculator.java:515-541
    if (!indexes || n_indexes <= 0 || !frequency) return fail(SDRGPU_ERR_INVALID_ARG, "Indexes cannot be empty");
    Calculator c{sample_rate, center_frequency, channel_count};
    const int index = indexes[(n_indexes - 1) / 2];
    if (n_indexes % 2 == 0) *frequency = (long long)c.index_max(index, false);
    else if (index == c.wrap()) *frequency = (long long)c.index_center(index, false);
    else *frequency = (long long)c.index_center(index, true);
    return SDRGPU_OK;
}

int sdrgpu_pack_dibits(const uint8_t *dibits, int n, uint8_t *out)
{
    // DibitToByteBufferAssembler.java:58-93
    int bytes = 0;
    for (int k = 0; k + 3 < n; k += 4)
        out[bytes++] = (uint8_t)(((dibits[k] & 3) << 6) | ((dibits[k + 1] & 3) << 4) | ((dibits[k + 2] & 3) << 2) |
                                 (dibits[k + 3] & 3));
    return bytes;
}

}  // extern "C"
