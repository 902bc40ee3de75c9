/* CPU ORACLE -- test infrastructure only (see sdr_oracle.h).
 * Half-band decimators and cascades, streaming FIR, FM discriminator, power squelch, block AGC.
 * Follows J/dsp/filter/halfband/, J/dsp/filter/decimate/, J/dsp/filter/fir/, J/dsp/fm/, J/dsp/squelch/,
 * J/dsp/gain/ComplexFeedForwardGainControl.java. */
#include "sdr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- half-band decimate-by-2
 * ComplexHalfBandDecimationFilter.java:47-123 / RealHalfBandDecimationFilter.java:49-111.
 * The Java keeps [residual | new] in one array and re-seats the residual from the tail of the previous
 * array whichever length the new buffer has, i.e. a plain streaming filter with L-1 samples of history. */
struct orc_halfband {
    int length;          /* L */
    float *coefficients; /* L */
    float *history_c;    /* 2L-2 floats (complex) */
    float *history_r;    /* L-1 floats (real) */
    float *buffer;
    int buffer_cap;
};

orc_halfband *orc_halfband_create(const float *coefficients, int length)
{
    if ((length + 1) % 4 != 0) return NULL;
    orc_halfband *h = (orc_halfband *)calloc(1, sizeof(*h));
    h->length = length;
    h->coefficients = (float *)malloc(sizeof(float) * (size_t)length);
    memcpy(h->coefficients, coefficients, sizeof(float) * (size_t)length);
    h->history_c = (float *)calloc((size_t)(2 * length - 2), sizeof(float));
    h->history_r = (float *)calloc((size_t)(length - 1), sizeof(float));
    return h;
}

void orc_halfband_destroy(orc_halfband *h)
{
    if (!h) return;
    free(h->coefficients);
    free(h->history_c);
    free(h->history_r);
    free(h->buffer);
    free(h);
}

static float *hb_buffer(orc_halfband *h, int n)
{
    if (h->buffer_cap < n) {
        free(h->buffer);
        h->buffer = (float *)malloc(sizeof(float) * (size_t)n);
        h->buffer_cap = n;
    }
    return h->buffer;
}

int orc_halfband_decimate_complex(orc_halfband *h, const float *samples, int n_floats, float *out)
{
    if (n_floats % 4 != 0) return -1;
    const int L = h->length;
    const int lm2 = 2 * L - 2; /* mCoefficientsLengthMinus2 of the I/Q-duplicated coefficient array */
    const int half = L - 1;    /* mHalf = (2L)/2 - 1 */
    float *buf = hb_buffer(h, n_floats + lm2);
    memcpy(buf, h->history_c, sizeof(float) * (size_t)lm2);
    memcpy(buf + lm2, samples, sizeof(float) * (size_t)n_floats);
    for (int bp = 0; bp < n_floats; bp += 4) {
        float acc_i = 0.0f, acc_q = 0.0f;
        for (int cp = 0; cp < half; cp += 4) {
            float c = h->coefficients[cp / 2];
            acc_i += c * (buf[bp + cp] + buf[bp + (lm2 - cp)]);
            acc_q += c * (buf[bp + cp + 1] + buf[bp + (lm2 - cp) + 1]);
        }
        acc_i += buf[bp + half] * 0.5f;
        acc_q += buf[bp + half + 1] * 0.5f;
        out[bp / 2] = acc_i;
        out[bp / 2 + 1] = acc_q;
    }
    memcpy(h->history_c, buf + n_floats, sizeof(float) * (size_t)lm2);
    return n_floats / 2;
}

int orc_halfband_decimate_real(orc_halfband *h, const float *samples, int n_floats, float *out)
{
    if (n_floats % 2 != 0) return -1;
    const int L = h->length;
    const int lm1 = L - 1;
    const int half = lm1 / 2;
    float *buf = hb_buffer(h, n_floats + lm1);
    memcpy(buf, h->history_r, sizeof(float) * (size_t)lm1);
    memcpy(buf + lm1, samples, sizeof(float) * (size_t)n_floats);
    for (int bp = 0; bp < n_floats; bp += 2) {
        float acc = 0.0f;
        for (int cp = 0; cp < half; cp += 2) {
            acc += h->coefficients[cp] * (buf[bp + cp] + buf[bp + (lm1 - cp)]);
        }
        acc += buf[bp + half] * 0.5f;
        out[bp / 2] = acc;
    }
    memcpy(h->history_r, buf + n_floats, sizeof(float) * (size_t)lm1);
    return n_floats / 2;
}

/* ---------------------------------------------------------------- cascades
 * DecimationFilterFactory.java:36-104 and the stage constants of {Complex,Real}DecimateX{2..1024}Filter
 * (e.g. ComplexDecimateX2Filter.java:31-32, X4:33-34, X8:32-33, X16:32-33, X32..X1024:32-33): the
 * highest-rate stage runs first; stage (rate r -> r/2) uses: r>=32: 11-tap Blackman, r=16: 15 Blackman,
 * r=8: 15 Blackman, r=4: 23 Blackman, r=2: 63 Hamming. */
struct orc_decimator {
    int rate;
    int n_stages;
    orc_halfband *stages[10];
    float *tmp_a, *tmp_b;
    int tmp_cap;
};

orc_decimator *orc_decimator_create(int rate)
{
    int ok = (rate == 0);
    for (int r = 2; r <= 1024; r *= 2) ok |= (rate == r);
    if (!ok) return NULL;
    orc_decimator *d = (orc_decimator *)calloc(1, sizeof(*d));
    d->rate = rate;
    for (int r = rate; r >= 2; r /= 2) {
        int len, win;
        if (r >= 32) { len = 11; win = ORC_WIN_BLACKMAN; }
        else if (r == 16) { len = 15; win = ORC_WIN_BLACKMAN; }
        else if (r == 8) { len = 15; win = ORC_WIN_BLACKMAN; }
        else if (r == 4) { len = 23; win = ORC_WIN_BLACKMAN; }
        else { len = 63; win = ORC_WIN_HAMMING; }
        float taps[64];
        orc_half_band(len, win, taps);
        d->stages[d->n_stages++] = orc_halfband_create(taps, len);
    }
    return d;
}

void orc_decimator_destroy(orc_decimator *d)
{
    if (!d) return;
    for (int i = 0; i < d->n_stages; i++) orc_halfband_destroy(d->stages[i]);
    free(d->tmp_a);
    free(d->tmp_b);
    free(d);
}

static int decimator_run(orc_decimator *d, const float *samples, int n_floats, float *out, int is_complex)
{
    if (d->rate == 0) {
        memcpy(out, samples, sizeof(float) * (size_t)n_floats);
        return n_floats;
    }
    int multiple = d->rate * (is_complex ? 2 : 1); /* VALIDATION_LENGTH (e.g. ComplexDecimateX4Filter.java:33,49) */
    if (n_floats % multiple != 0) return -1;
    if (d->tmp_cap < n_floats) {
        free(d->tmp_a);
        free(d->tmp_b);
        d->tmp_a = (float *)malloc(sizeof(float) * (size_t)n_floats);
        d->tmp_b = (float *)malloc(sizeof(float) * (size_t)n_floats);
        d->tmp_cap = n_floats;
    }
    const float *src = samples;
    int n = n_floats;
    for (int i = 0; i < d->n_stages; i++) {
        float *dst = (i == d->n_stages - 1) ? out : ((i & 1) ? d->tmp_b : d->tmp_a);
        n = is_complex ? orc_halfband_decimate_complex(d->stages[i], src, n, dst)
                       : orc_halfband_decimate_real(d->stages[i], src, n, dst);
        src = dst;
    }
    return n;
}

int orc_decimator_complex(orc_decimator *d, const float *s, int n, float *out) { return decimator_run(d, s, n, out, 1); }
int orc_decimator_real(orc_decimator *d, const float *s, int n, float *out) { return decimator_run(d, s, n, out, 0); }

/* ---------------------------------------------------------------- FIR
 * RealFIRFilter2.java:77-95: shift the delay line by one (newest at index 0), then
 * acc = Math.fma(data[x], coefficient[x], acc) for x ascending, then acc *= gain. */
struct orc_fir {
    int n;
    float gain;
    float *coefficients;
    float *data;
};

orc_fir *orc_fir_create(const float *taps, int n, float gain)
{
    orc_fir *f = (orc_fir *)calloc(1, sizeof(*f));
    f->n = n;
    f->gain = gain;
    f->coefficients = (float *)malloc(sizeof(float) * (size_t)n);
    memcpy(f->coefficients, taps, sizeof(float) * (size_t)n);
    f->data = (float *)calloc((size_t)n, sizeof(float));
    return f;
}

void orc_fir_destroy(orc_fir *f)
{
    if (!f) return;
    free(f->coefficients);
    free(f->data);
    free(f);
}

float orc_fir_filter(orc_fir *f, float sample)
{
    memmove(f->data + 1, f->data, sizeof(float) * (size_t)(f->n - 1));
    f->data[0] = sample;
    float acc = 0.0f;
    for (int x = 0; x < f->n; x++) acc = fmaf(f->data[x], f->coefficients[x], acc);
    acc *= f->gain;
    return acc;
}

void orc_fir_filter_real(orc_fir *f, const float *in, int n, float *out)
{
    for (int x = 0; x < n; x++) out[x] = orc_fir_filter(f, in[x]);
}

/* ComplexFIRFilter2.java:37-41,112-129 */
orc_cfir *orc_cfir_create(const float *taps, int n, float gain)
{
    orc_cfir *f = (orc_cfir *)calloc(1, sizeof(*f));
    f->i = orc_fir_create(taps, n, gain);
    f->q = orc_fir_create(taps, n, gain);
    return f;
}

void orc_cfir_destroy(orc_cfir *f)
{
    if (!f) return;
    orc_fir_destroy(f->i);
    orc_fir_destroy(f->q);
    free(f);
}

void orc_cfir_filter(orc_cfir *f, const float *in, int n_floats, float *out)
{
    for (int x = 0; x < n_floats; x += 2) {
        out[x] = orc_fir_filter(f->i, in[x]);
        out[x + 1] = orc_fir_filter(f->q, in[x + 1]);
    }
}

/* ---------------------------------------------------------------- FM discriminator, FMDemodulator.java:62-96 */
void orc_fm_init(orc_fm *f, float gain)
{
    f->prev_i = 0.0f;
    f->prev_q = 0.0f;
    f->gain = gain;
}

float orc_fm_demodulate(orc_fm *f, float ci, float cq)
{
    /* products and sums are float arithmetic, widened afterwards */
    double inphase = (double)((ci * f->prev_i) - (cq * -f->prev_q));
    double quadrature = (double)((cq * f->prev_i) + (ci * -f->prev_q));
    double angle = 0.0;
    if (inphase != 0) {
        double denominator = 1.0 / inphase;
        angle = atan(quadrature * denominator);
    }
    f->prev_i = ci;
    f->prev_q = cq;
    return (float)(angle * (double)f->gain);
}

void orc_fm_demodulate_buffer(orc_fm *f, const float *iq, int n_floats, float *out)
{
    for (int x = 0; x < n_floats; x += 2) out[x / 2] = orc_fm_demodulate(f, iq[x], iq[x + 1]);
}

/* PowerSquelch.java:56-62,88-159 ; SinglePoleIirFilter.java */
enum { SQ_ATTACK = 0, SQ_DECAY = 1, SQ_MUTE = 2, SQ_UNMUTE = 3 };

void orc_squelch_init(orc_squelch *s, double alpha, double threshold_db, int ramp)
{
    memset(s, 0, sizeof(*s));
    s->alpha = alpha;
    s->one_minus_alpha = 1.0 - alpha;
    s->threshold = pow(10.0, threshold_db / 10.0);
    s->ramp_threshold = ramp;
    s->state = SQ_MUTE;
}

void orc_squelch_process(orc_squelch *s, double inphase, double quadrature)
{
    s->output = (s->output * s->one_minus_alpha) + (s->alpha * (inphase * inphase + quadrature * quadrature));
    s->power = s->output;
    int mute = s->power < s->threshold;
    int change = 0;
    switch (s->state) {
        case SQ_MUTE:
            if (!mute) {
                if (s->ramp_threshold > 0) {
                    s->state = SQ_ATTACK;
                    s->ramp_count++;
                } else {
                    s->state = SQ_UNMUTE;
                    change = 1;
                }
            }
            break;
        case SQ_ATTACK:
            if (s->ramp_count >= s->ramp_threshold) {
                s->state = SQ_UNMUTE;
                change = 1;
            } else {
                s->ramp_count++;
            }
            break;
        case SQ_DECAY:
            if (s->ramp_count <= 0) {
                s->state = SQ_MUTE;
                change = 1;
            } else {
                s->ramp_count--;
            }
            break;
        case SQ_UNMUTE:
            if (mute) {
                if (s->ramp_threshold > 0) {
                    s->state = SQ_DECAY;
                    s->ramp_count--;
                } else {
                    s->state = SQ_MUTE;
                    change = 1;
                }
            }
            break;
    }
    s->squelch_changed = change;
}

/* SquelchingFMDemodulator.java:56-101 (default gain 1.0f from FMDemodulator()) */
void orc_sqfm_init(orc_sqfm *s, double alpha, double threshold_db, int ramp)
{
    orc_fm_init(&s->fm, 1.0f);
    orc_squelch_init(&s->sq, alpha, threshold_db, ramp);
    s->squelch_changed = 0;
}

void orc_sqfm_demodulate_buffer(orc_sqfm *s, const float *iq, int n_floats, float *out)
{
    s->squelch_changed = 0;
    for (int x = 0; x < n_floats; x += 2) {
        float i = iq[x], q = iq[x + 1];
        orc_squelch_process(&s->sq, (double)i, (double)q);
        if (s->sq.state == SQ_UNMUTE || s->sq.state == SQ_DECAY) {
            out[x / 2] = orc_fm_demodulate(&s->fm, i, q);
        } else {
            out[x / 2] = 0.0f;
        }
        if (s->sq.squelch_changed) s->squelch_changed = 1;
    }
}

/* ---------------------------------------------------------------- block AGC
 * ComplexFeedForwardGainControl.java:29-30,103-107,147-181 ; Complex.envelope (Complex.java:457-470) */
void orc_agc_block(const float *in, int n_floats, float *out)
{
    float max_envelope = 0.0001f;
    for (int x = 0; x < n_floats; x += 2) {
        float ia = fabsf(in[x]), qa = fabsf(in[x + 1]);
        float env = (ia > qa) ? ia + (0.4f * qa) : qa + (0.4f * ia);
        if (env > max_envelope) max_envelope = env;
    }
    float gain = 1.0f / max_envelope;
    for (int x = 0; x < n_floats; x += 2) {
        out[x] = in[x] * gain;
        out[x + 1] = in[x + 1] * gain;
    }
}
