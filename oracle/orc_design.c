/* CPU ORACLE -- test infrastructure only (see sdr_oracle.h).
 * Closed-form filter designers.  Follows J/dsp/filter/FilterFactory.java and J/dsp/filter/Window.java. */
#include "sdr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static const double ORC_PI = 3.14159265358979323846;

/* Window.java:133-150 (Blackman, 3-term "exact" coefficients) and :277-287 (Hamming) */
int orc_window(int type, int length, double *out)
{
    if (type == ORC_WIN_BLACKMAN) {
        double denominator = length - 1;
        double a0 = 0.426590713672, a1 = 0.496560619089, a2 = 0.0768486672399;
        for (int x = 0; x < length; x++) {
            out[x] = a0 - (a1 * cos((2.0 * ORC_PI * (double)x) / denominator)) +
                     (a2 * cos((4.0 * ORC_PI * (double)x) / denominator));
        }
        return 0;
    }
    if (type == ORC_WIN_HAMMING) {
        for (int x = 0; x < length; x++) {
            out[x] = 0.54 - (0.46 * cos((2.0 * ORC_PI * x) / (length - 1)));
        }
        return 0;
    }
    return -1;
}

/* Window.java:386-401 zeroth-order modified Bessel function, series form */
static double bessel_i0(double x)
{
    double f = 1;
    const double x2 = x * x * 0.25;
    double xc = x2;
    double v = 1 + x2;
    for (int i = 2; i < 100; i++) {
        f *= i;
        xc *= x2;
        const double a = xc / (f * f);
        v += a;
        if (a < 1e-20) break;
    }
    return v;
}

/* Window.java:343-355 */
static double kaiser_beta(double attenuation)
{
    if (attenuation > 50.0) return 0.1102 * (attenuation - 8.7);
    if (attenuation >= 21.0) return (0.5842 * pow(attenuation - 21.0, 0.4)) + (0.07886 * (attenuation - 21.0));
    return 0.0;
}

/* Window.java:364-378 */
int orc_kaiser(int length, double attenuation, double *out)
{
    double beta = kaiser_beta(attenuation);
    double denom = bessel_i0(beta);
    for (int x = 0; x < length; x++) {
        double r = 2.0 * x / (length - 1) - 1.0;
        double temp = beta * sqrt(1.0 - pow(r, 2));
        out[x] = bessel_i0(temp) / denom;
    }
    return 0;
}

/* FilterFactory.java:970-998 */
int orc_kaiser_sinc(int length, double cutoff, double attenuation, float *out)
{
    if (length % 2 == 0) return -1;
    int half = length / 2;
    double *window = (double *)malloc(sizeof(double) * (size_t)length);
    orc_kaiser(length, attenuation, window);
    double scalor = 2.0 * cutoff;
    double pi_scalor = ORC_PI * scalor;
    out[half] = (float)(1.0 * scalor * window[half]);
    for (int x = 1; x <= half; x++) {
        double a = pi_scalor * x;
        double coefficient = scalor * sin(a) / a;
        coefficient *= window[half + x];
        out[half + x] = (float)coefficient;
        out[half - x] = (float)coefficient;
    }
    free(window);
    return 0;
}

/* FilterFactory.java:690-714 -- note decibel() narrows to float before returning a double */
double orc_evaluate(const float *filter, int length, double frequency)
{
    double real = 0.0, imag = 0.0;
    for (int x = 0; x < length; x++) {
        real += filter[x] * cos(ORC_PI * frequency * (double)x);
        imag += filter[x] * sin(ORC_PI * frequency * (double)x);
    }
    return (float)(10.0 * log10(pow(real, 2.0) + pow(imag, 2.0)));
}

static int matches_objective(double a)
{
    /* FilterFactory.java:40-41,1045-1048 */
    return fabs(a - (-6.020599842071533)) <= 0.0003;
}

/* FilterFactory.java:808-920 */
int orc_sinc_m2_channelizer(double channel_bandwidth, int channels, int taps_per_channel, float *out, int out_capacity)
{
    int current_tpc = taps_per_channel;
    int filter_length = (channels * current_tpc) - 1;
    double sample_rate = channel_bandwidth * channels;
    double band_edge = channel_bandwidth / sample_rate;
    double cutoff = band_edge / 2.0;
    double increment = cutoff * 0.1;
    int cap = channels * (taps_per_channel + 10);
    float *taps = (float *)malloc(sizeof(float) * (size_t)cap);
    float *higher = (float *)malloc(sizeof(float) * (size_t)cap);
    int rc = 0;

    orc_kaiser_sinc(filter_length, cutoff, 80.0, taps);
    double response = orc_evaluate(taps, filter_length, band_edge);
    double threshold = 1.0 / sample_rate;

    while (increment > threshold) {
        if (matches_objective(response) && (cutoff + increment <= band_edge)) {
            orc_kaiser_sinc(filter_length, cutoff + increment, 80.0, higher);
            double higher_response = orc_evaluate(higher, filter_length, band_edge);
            if (matches_objective(higher_response)) {
                cutoff += increment;
                memcpy(taps, higher, sizeof(float) * (size_t)filter_length);
                response = higher_response;
            } else {
                increment /= 2.0;
            }
        } else if (matches_objective(response)) {
            increment /= 2.0;
        } else {
            cutoff -= increment;
            if (cutoff <= 0) {
                current_tpc++;
                if (current_tpc > (taps_per_channel + 10)) {
                    rc = -2;
                    goto done;
                }
                filter_length = channels * current_tpc - 1;
                cutoff = channel_bandwidth / sample_rate;
                increment = cutoff * 0.1;
            }
            orc_kaiser_sinc(filter_length, cutoff, 80.0, taps);
            response = orc_evaluate(taps, filter_length, band_edge);
        }
    }
    if (!matches_objective(response)) {
        rc = -3;
        goto done;
    }
    if (filter_length + 1 > out_capacity) {
        rc = -4;
        goto done;
    }
    out[0] = 0.0f; /* odd-length filter pre-padded with one zero coefficient */
    memcpy(out + 1, taps, sizeof(float) * (size_t)filter_length);
    rc = filter_length + 1;
done:
    free(taps);
    free(higher);
    return rc;
}

/* FilterFactory.java:755-770 */
int orc_sinc_m2_synthesizer(double channel_sample_rate, double channel_bandwidth, int channels,
                            int taps_per_channel, float *out)
{
    int filter_length = (channels * taps_per_channel) - 1;
    double cutoff = (channel_bandwidth * 1.10) / (channel_sample_rate * (double)channels);
    out[0] = 0.0f;
    if (orc_kaiser_sinc(filter_length, cutoff, 80.0, out + 1) != 0) return -1;
    return filter_length + 1;
}

/* FilterFactory.java:1007-1036 */
int orc_half_band(int length, int window_type, float *out)
{
    if ((length - 3) % 4 != 0) return -1;
    double *window = (double *)malloc(sizeof(double) * (size_t)length);
    if (orc_window(window_type, length, window) != 0) {
        free(window);
        return -1;
    }
    int half_length = length / 2;
    for (int x = 0; x < length; x++) {
        int offset = x - half_length;
        out[x] = 0.0f;
        if (offset == 0) {
            out[x] = 0.5f;
        } else if ((x % 2) == 0) {
            out[x] = (float)((sin(offset * ORC_PI / 2) / (offset * ORC_PI)) * window[x]);
        }
    }
    free(window);
    return length;
}
