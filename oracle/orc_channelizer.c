/* CPU ORACLE -- test infrastructure only (see sdr_oracle.h).
 * Polyphase channelizer (filter bank + per-block inverse DFT), channel calculator.
 * Follows J/dsp/filter/channelizer/ComplexPolyphaseChannelizerM2.java and ChannelCalculator.java. */
#include "sdr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static const double ORC_PI = 3.14159265358979323846;

/* ------------------------------------------------------------------------------------------------
 * Inverse FFT.  The reference calls JTransforms 3.1 FloatFFT_1D(M).complexInverse(a, true)
 * (ComplexPolyphaseChannelizerM2.java:392,423) -- a float32 FFTPACK-style mixed-radix transform that is
 * NOT vendored under /root/reference.  Restated here as: a float32 decimation-in-frequency mixed-radix
 * Stockham transform (radix 4/2/3/5 + generic odd radix, twiddles computed in double and rounded to
 * float, like FFTPACK's cffti) followed by the 1/n float scale.  orc_idft_f64 is the order-independent
 * check; the parity tolerance (1e-4 relative RMS) absorbs the summation-order difference.
 * ---------------------------------------------------------------------------------------------- */
struct orc_fft {
    int n;
    int n_factors;
    int factors[32];
    float *tw; /* e^{+j 2 pi k / n}, k = 0..n-1, interleaved */
    float *scratch;
    float *vr, *vi; /* butterfly scratch, n floats each */
};

orc_fft *orc_fft_create(int n)
{
    orc_fft *f = (orc_fft *)calloc(1, sizeof(orc_fft));
    f->n = n;
    int rem = n;
    static const int tryh[4] = {4, 2, 3, 5};
    for (int t = 0; t < 4; t++) {
        while (rem % tryh[t] == 0) {
            f->factors[f->n_factors++] = tryh[t];
            rem /= tryh[t];
        }
    }
    for (int p = 7; rem > 1; p += 2) {
        while (rem % p == 0) {
            f->factors[f->n_factors++] = p;
            rem /= p;
        }
    }
    f->tw = (float *)malloc(sizeof(float) * 2 * (size_t)n);
    f->scratch = (float *)malloc(sizeof(float) * 2 * (size_t)n);
    f->vr = (float *)malloc(sizeof(float) * (size_t)n);
    f->vi = (float *)malloc(sizeof(float) * (size_t)n);
    for (int k = 0; k < n; k++) {
        double a = 2.0 * ORC_PI * (double)k / (double)n;
        f->tw[2 * k] = (float)cos(a);
        f->tw[2 * k + 1] = (float)sin(a);
    }
    return f;
}

void orc_fft_destroy(orc_fft *f)
{
    if (!f) return;
    free(f->tw);
    free(f->scratch);
    free(f->vr);
    free(f->vi);
    free(f);
}

/* One decimation-in-frequency Stockham (autosort) pass of radix r.  s = product of the radices already
 * applied (stride), cur = n / s the current sub-transform length, m = cur / r.
 *   y[q + s*(r*p + k)] = ( sum_j x[q + s*(p + m*j)] * W_r^{jk} ) * w_cur^{p*k}
 * with W_r = e^{+j 2 pi / r}, w_cur = e^{+j 2 pi / cur}; all twiddles come from the n-th root table. */
static void stockham_pass(const orc_fft *f, int r, int s, const float *x, float *y)
{
    const int n = f->n;
    const int m = n / (r * s);
    float *vr = f->vr, *vi = f->vi;
    for (int p = 0; p < m; p++) {
        for (int q = 0; q < s; q++) {
            for (int j = 0; j < r; j++) {
                int idx = q + s * (p + m * j);
                vr[j] = x[2 * idx];
                vi[j] = x[2 * idx + 1];
            }
            for (int k = 0; k < r; k++) {
                float sr = 0.0f, si = 0.0f;
                for (int j = 0; j < r; j++) {
                    int t = (int)(((long long)j * k * (n / r)) % n);
                    float wr = f->tw[2 * t], wi = f->tw[2 * t + 1];
                    sr += vr[j] * wr - vi[j] * wi;
                    si += vr[j] * wi + vi[j] * wr;
                }
                int t2 = (int)(((long long)s * p * k) % n);
                float wr = f->tw[2 * t2], wi = f->tw[2 * t2 + 1];
                int oidx = q + s * (r * p + k);
                y[2 * oidx] = sr * wr - si * wi;
                y[2 * oidx + 1] = sr * wi + si * wr;
            }
        }
    }
}

void orc_ifft_f32(orc_fft *f, float *a)
{
    const int n = f->n;
    float *x = a, *y = f->scratch;
    int s = 1;
    for (int i = 0; i < f->n_factors; i++) {
        int r = f->factors[i];
        stockham_pass(f, r, s, x, y);
        s *= r;
        float *t = x;
        x = y;
        y = t;
    }
    if (x != a) memcpy(a, x, sizeof(float) * 2 * (size_t)n);
    /* JTransforms scale(): norm = 1.0f / n applied as a float multiply */
    float norm = 1.0f / (float)n;
    for (int i = 0; i < 2 * n; i++) a[i] *= norm;
}

void orc_idft_f64(int n, const float *in, float *out)
{
    for (int k = 0; k < n; k++) {
        double sr = 0.0, si = 0.0;
        for (int j = 0; j < n; j++) {
            long long t = ((long long)j * k) % n;
            double ang = 2.0 * ORC_PI * (double)t / (double)n;
            double c = cos(ang), s = sin(ang);
            sr += (double)in[2 * j] * c - (double)in[2 * j + 1] * s;
            si += (double)in[2 * j] * s + (double)in[2 * j + 1] * c;
        }
        out[2 * k] = (float)(sr / n);
        out[2 * k + 1] = (float)(si / n);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Channelizer
 * ---------------------------------------------------------------------------------------------- */
struct orc_channelizer {
    int m;              /* channel count */
    int taps_per_channel;
    int sub;            /* sub-channel count = 2*m floats */
    int buffer_length;  /* sub * taps_per_channel */
    float *inline_samples;
    float *inline_filter;
    float *interim;
    float *accum;
    int *top_map, *middle_map;
    int top_block;
    int pointer;
    orc_fft *fft;
    float *tmp;
};

/* ComplexPolyphaseChannelizerM2.java:301-330 getAlignedFilter */
static void aligned_filter(const float *coefficients, int n_coefficients, int channel_count, int tpc, float *filter)
{
    int len = channel_count * tpc * 2;
    memset(filter, 0, sizeof(float) * (size_t)len);
    int fp = 0;
    for (int cp = 0; cp < n_coefficients; cp++) {
        filter[fp++] = coefficients[cp];
        filter[fp++] = coefficients[cp];
    }
    int block = channel_count;
    for (int x = 0; x < len; x += block) {
        for (int y = 0; y < block / 2; y++) {
            int i1 = x + y, i2 = x + (block - y - 1);
            float t = filter[i2];
            filter[i2] = filter[i1];
            filter[i1] = t;
        }
    }
}

/* ComplexPolyphaseChannelizerM2.java:244-293 */
static void block_maps(int channel_count, int *top, int *middle)
{
    int block = channel_count / 2;
    int offset = 2 * block;
    for (int channel = 0; channel < block; channel++) {
        int new_index = 2 * channel;
        int original = 2 * (block - channel - 1);
        top[original] = new_index;
        top[original + 1] = new_index + 1;
        top[offset + original] = offset + new_index;
        top[offset + original + 1] = offset + new_index + 1;
        middle[offset + original] = new_index;
        middle[offset + original + 1] = new_index + 1;
        middle[original] = offset + new_index;
        middle[original + 1] = offset + new_index + 1;
    }
}

/* ComplexPolyphaseChannelizerM2.java:93-106,390-400 */
orc_channelizer *orc_chan_create(const float *taps, int n_taps, int channel_count)
{
    if (channel_count % 2 != 0 || channel_count <= 0) return NULL;
    orc_channelizer *c = (orc_channelizer *)calloc(1, sizeof(orc_channelizer));
    c->m = channel_count;
    c->taps_per_channel = (int)ceil((double)n_taps / (double)channel_count);
    c->sub = 2 * channel_count;
    c->buffer_length = c->sub * c->taps_per_channel;
    c->inline_samples = (float *)calloc((size_t)c->buffer_length, sizeof(float));
    c->inline_filter = (float *)calloc((size_t)c->buffer_length, sizeof(float));
    c->interim = (float *)calloc((size_t)c->buffer_length, sizeof(float));
    c->accum = (float *)calloc((size_t)c->sub, sizeof(float));
    c->top_map = (int *)calloc((size_t)c->sub, sizeof(int));
    c->middle_map = (int *)calloc((size_t)c->sub, sizeof(int));
    c->tmp = (float *)calloc((size_t)c->sub, sizeof(float));
    aligned_filter(taps, n_taps, channel_count, c->taps_per_channel, c->inline_filter);
    block_maps(channel_count, c->top_map, c->middle_map);
    c->top_block = 1;
    c->pointer = 0;
    c->fft = orc_fft_create(channel_count);
    return c;
}

void orc_chan_destroy(orc_channelizer *c)
{
    if (!c) return;
    free(c->inline_samples);
    free(c->inline_filter);
    free(c->interim);
    free(c->accum);
    free(c->top_map);
    free(c->middle_map);
    free(c->tmp);
    orc_fft_destroy(c->fft);
    free(c);
}

/* ComplexPolyphaseChannelizerM2.java:337-383 process() */
static void chan_process(orc_channelizer *c, float *processed)
{
    const int sub = c->sub;
    for (int x = 0; x < c->buffer_length; x++) {
        c->interim[x] = c->inline_samples[x] * c->inline_filter[x];
    }
    for (int ch = 0; ch < sub; ch++) c->accum[ch] = 0.0f;
    for (int tap = 0; tap < c->taps_per_channel; tap++) {
        int tap_offset = tap * sub;
        for (int ch = 0; ch < sub; ch++) {
            c->accum[ch] += c->interim[tap_offset + ch];
        }
    }
    const int *map = c->top_block ? c->top_map : c->middle_map;
    for (int x = 0; x < sub; x++) processed[x] = c->accum[map[x]];
    c->top_block = !c->top_block;
}

/* ComplexPolyphaseChannelizerM2.java:190-235 receive(); mode 0 = raw accumulators, 1 = f32 FFT, 2 = f64 DFT */
static int chan_receive(orc_channelizer *c, const float *samples, int n_floats, float *out, int mode)
{
    const int per_block = c->m; /* mSamplesPerBlock = channelCount floats */
    int sp = 0, blocks = 0;
    while (sp < n_floats) {
        if (c->pointer < per_block) {
            int to_copy = per_block - c->pointer;
            int diff = n_floats - sp;
            if (diff < to_copy) to_copy = diff;
            memcpy(c->inline_samples + c->pointer, samples + sp, sizeof(float) * (size_t)to_copy);
            c->pointer += to_copy;
            sp += to_copy;
        }
        if (c->pointer >= per_block) {
            float *dst = out + (size_t)blocks * (size_t)c->sub;
            chan_process(c, dst);
            if (mode == 1) {
                orc_ifft_f32(c->fft, dst);
            } else if (mode == 2) {
                memcpy(c->tmp, dst, sizeof(float) * (size_t)c->sub);
                orc_idft_f64(c->m, c->tmp, dst);
            }
            blocks++;
            memmove(c->inline_samples + per_block, c->inline_samples,
                    sizeof(float) * (size_t)(c->buffer_length - per_block));
            c->pointer = 0;
        }
    }
    return blocks;
}

int orc_chan_receive(orc_channelizer *c, const float *samples, int n_floats, float *out, int use_f64_dft)
{
    return chan_receive(c, samples, n_floats, out, use_f64_dft ? 2 : 1);
}

int orc_chan_receive_raw(orc_channelizer *c, const float *samples, int n_floats, float *out)
{
    return chan_receive(c, samples, n_floats, out, 0);
}

/* ------------------------------------------------------------------------------------------------
 * ChannelCalculator.java:223-541
 * ---------------------------------------------------------------------------------------------- */
enum { POLICY_POSITIVE = 0, POLICY_NEGATIVE = 1 };

static double cc_bw(const orc_calc *c) { return (double)c->sample_rate / (double)c->channel_count; }
static double cc_half_bw(const orc_calc *c) { return cc_bw(c) / 2.0; }
static int cc_wrap(const orc_calc *c) { return c->channel_count / 2; }
static int cc_normalize(const orc_calc *c, int index)
{
    while (index < 0) index += c->channel_count;
    while (index >= c->channel_count) index -= c->channel_count;
    return index;
}

/* ChannelCalculator.java:397-426 */
static double cc_index_center(const orc_calc *c, int index, int policy)
{
    int wrap = cc_wrap(c);
    if (index == wrap) {
        if (policy == POLICY_POSITIVE) return c->center_frequency + (index * cc_bw(c));
        return c->center_frequency - (index * cc_bw(c));
    } else if (index < wrap) {
        return c->center_frequency + (index * cc_bw(c));
    }
    return c->center_frequency - ((c->channel_count - index) * cc_bw(c));
}

/* ChannelCalculator.java:437-467 */
static double cc_index_min(const orc_calc *c, int index, int policy)
{
    int wrap = cc_wrap(c);
    if (index == wrap) {
        if (policy == POLICY_POSITIVE) return c->center_frequency + ((double)index * cc_bw(c)) - cc_half_bw(c);
        return cc_index_center(c, index, policy);
    } else if (index <= wrap) {
        return c->center_frequency + ((double)index * cc_bw(c)) - cc_half_bw(c);
    }
    return c->center_frequency - ((double)(c->channel_count - index) * cc_bw(c)) - cc_half_bw(c);
}

/* ChannelCalculator.java:478-508 */
static double cc_index_max(const orc_calc *c, int index, int policy)
{
    int wrap = cc_wrap(c);
    if (index == wrap) {
        if (policy == POLICY_POSITIVE) return cc_index_center(c, index, policy);
        return c->center_frequency - ((double)index * cc_bw(c) - cc_half_bw(c));
    } else if (index <= wrap) {
        return c->center_frequency + ((double)index * cc_bw(c)) + cc_half_bw(c);
    }
    return c->center_frequency - ((double)(c->channel_count - index) * cc_bw(c)) + cc_half_bw(c);
}

/* ChannelCalculator.java:343-370 */
static int cc_is_overlap(const orc_calc *c, long long frequency, int index1, int index2)
{
    index1 = cc_normalize(c, index1);
    index2 = cc_normalize(c, index2);
    int delta = cc_normalize(c, index2 - index1);
    if (delta != 1) return 0;
    long long max1 = (long long)cc_index_max(c, index1, POLICY_POSITIVE);
    long long min2 = (long long)cc_index_min(c, index2, POLICY_NEGATIVE);
    if (index1 == cc_wrap(c)) max1 = (long long)cc_index_max(c, index1, POLICY_NEGATIVE);
    if (index2 == cc_wrap(c)) min2 = (long long)cc_index_min(c, index2, POLICY_POSITIVE);
    return frequency == max1 && frequency == min2;
}

/* ChannelCalculator.java:293-330 */
static int cc_index_for_frequency(const orc_calc *c, long long frequency, int policy)
{
    double offset = frequency - c->center_frequency;
    if (fabs(offset) < cc_half_bw(c)) return 0;
    if (offset > 0) offset += cc_half_bw(c);
    else offset -= cc_half_bw(c);
    int index_offset = (int)(offset / cc_bw(c));
    if (index_offset < 0) index_offset += c->channel_count;
    if (policy == POLICY_POSITIVE && cc_is_overlap(c, frequency, index_offset, index_offset + 1)) {
        index_offset = cc_normalize(c, index_offset + 1);
    } else if (policy == POLICY_NEGATIVE && cc_is_overlap(c, frequency, index_offset - 1, index_offset)) {
        index_offset = cc_normalize(c, index_offset - 1);
    }
    return index_offset;
}

/* ChannelCalculator.java:223-281; TunerChannel.java:52-60 (min/max = freq -/+ bw/2, integer division) */
int orc_calc_channel_indexes(const orc_calc *c, long long frequency, int bandwidth, int *indexes, int cap)
{
    long long ch_min = frequency - (bandwidth / 2);
    long long ch_max = frequency + (bandwidth / 2);
    long long min_f = (long long)(c->center_frequency - (c->sample_rate / 2.0));
    long long max_f = (long long)(c->center_frequency + (c->sample_rate / 2.0));
    if (ch_min < min_f || ch_max > max_f) return -1;
    int min_index = cc_index_for_frequency(c, ch_min, POLICY_POSITIVE);
    int max_index = cc_index_for_frequency(c, ch_max, POLICY_NEGATIVE);
    if (min_index == cc_wrap(c) && max_index == cc_wrap(c)) return -2;
    int n = 0;
    if (n < cap) indexes[n] = min_index;
    n++;
    if (min_index != max_index) {
        if (min_index < 0 || min_index >= c->channel_count || max_index < 0 || max_index >= c->channel_count) return -3;
        int pointer = min_index + 1;
        if (pointer >= c->channel_count) pointer -= c->channel_count;
        while (pointer != max_index) {
            if (n < cap) indexes[n] = pointer;
            n++;
            pointer++;
            if (pointer >= c->channel_count) pointer -= c->channel_count;
        }
        if (n < cap) indexes[n] = max_index;
        n++;
    }
    return n;
}

/* ChannelCalculator.java:515-541 */
long long orc_calc_center_frequency_for_indexes(const orc_calc *c, const int *indexes, int n)
{
    int center_index = (n - 1) / 2;
    int index = indexes[center_index];
    if (n % 2 == 0) return (long long)cc_index_max(c, index, POLICY_NEGATIVE);
    if (index == cc_wrap(c)) return (long long)cc_index_center(c, index, POLICY_NEGATIVE);
    return (long long)cc_index_center(c, index, POLICY_POSITIVE);
}
