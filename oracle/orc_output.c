/* CPU ORACLE -- test infrastructure only (see sdr_oracle.h).
 * Per-channel extraction from the channelizer results: gather, oscillator mix, gain, two-channel
 * synthesizer.  Follows the J/dsp/filter/channelizer/output/ classes, TwoChannelSynthesizerM2.java,
 * J/dsp/mixer/{Oscillator,AbstractOscillator,FS4DownConverter}.java. */
#include "sdr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static const double ORC_PI = 3.14159265358979323846;

/* Complex.java:118-129 */
static inline float mul_i(float ia, float qa, float ib, float qb) { return (ia * ib) - (qa * qb); }
static inline float mul_q(float ia, float qa, float ib, float qb) { return (qa * ib) + (ia * qb); }

/* Oscillator.java:24,42-46,64-68; Complex.fromAngle (Complex.java:374-377) */
void orc_osc_set_frequency(orc_oscillator *o, double frequency, double sample_rate)
{
    float angle_per_sample = (float)(2.0 * ORC_PI * frequency / sample_rate);
    o->angle_i = (float)cos((double)angle_per_sample);
    o->angle_q = (float)sin((double)angle_per_sample);
}

void orc_osc_init(orc_oscillator *o, double frequency, double sample_rate)
{
    o->cur_i = 0.0f;
    o->cur_q = -1.0f;
    orc_osc_set_frequency(o, frequency, sample_rate);
}

/* AbstractOscillator.java:102-116 mixComplex; rotate() = multiply then fastNormalize (Complex.java:233-236) */
void orc_osc_mix(orc_oscillator *o, float *samples, int n_floats)
{
    for (int x = 0; x < n_floats; x += 2) {
        float i = mul_i(samples[x], samples[x + 1], o->cur_i, o->cur_q);
        float q = mul_q(samples[x], samples[x + 1], o->cur_i, o->cur_q);
        samples[x] = i;
        samples[x + 1] = q;
        float ni = mul_i(o->cur_i, o->cur_q, o->angle_i, o->angle_q);
        float nq = mul_q(o->cur_i, o->cur_q, o->angle_i, o->angle_q);
        float norm = (float)((ni * ni) + (nq * nq));
        float scalor = (float)(1.9999f - norm);
        o->cur_i = ni * scalor;
        o->cur_q = nq * scalor;
    }
}

/* ReusableChannelResultsBuffer.java:112-153 */
void orc_get_channel(const float *results, int n_blocks, int m, int bin, float *out)
{
    int i_index = 2 * bin;
    for (int b = 0; b < n_blocks; b++) {
        const float *r = results + (size_t)b * 2 * (size_t)m;
        out[2 * b] = r[i_index];
        out[2 * b + 1] = r[i_index + 1];
    }
}

/* ReusableComplexBuffer.java:63-71: samples[x] *= gain with a double gain */
void orc_apply_gain(float *samples, int n_floats, double gain)
{
    for (int x = 0; x < n_floats; x++) samples[x] = (float)((double)samples[x] * gain);
}

/* ---------------------------------------------------------------- OneChannelOutputProcessor.java:81-106 */
struct orc_one_channel {
    int bin;
    double gain, sample_rate;
    orc_oscillator osc;
    int correction_enabled;
};

orc_one_channel *orc_one_channel_create(double sample_rate, int bin, double gain)
{
    orc_one_channel *p = (orc_one_channel *)calloc(1, sizeof(*p));
    p->bin = bin;
    p->gain = gain;
    p->sample_rate = sample_rate;
    orc_osc_init(&p->osc, 0, sample_rate);
    return p;
}

void orc_one_channel_destroy(orc_one_channel *p) { free(p); }

/* ChannelOutputProcessor.java:91-95 */
void orc_one_channel_set_frequency_offset(orc_one_channel *p, long long offset)
{
    orc_osc_set_frequency(&p->osc, (double)offset, p->sample_rate);
    p->correction_enabled = (offset != 0);
}

void orc_one_channel_process(orc_one_channel *p, const float *results, int n_blocks, int m, float *out)
{
    orc_get_channel(results, n_blocks, m, p->bin, out);
    if (p->correction_enabled) orc_osc_mix(&p->osc, out, 2 * n_blocks);
    orc_apply_gain(out, 2 * n_blocks, p->gain);
}

/* ---------------------------------------------------------------- TwoChannelSynthesizerM2.java:74-191 */
struct orc_two_channel {
    int bin1, bin2;
    double gain, sample_rate;
    orc_oscillator osc;
    float *serpentine, *filter, *product;
    int len;
    int top_block;
    int fs4_pointer;
    float *c1, *c2;
    int cap;
};

orc_two_channel *orc_two_channel_create(double sample_rate, int bin1, int bin2, const float *filter,
                                        int filter_len, double gain)
{
    orc_two_channel *p = (orc_two_channel *)calloc(1, sizeof(*p));
    p->bin1 = bin1;
    p->bin2 = bin2;
    p->gain = gain;
    p->sample_rate = sample_rate;
    orc_osc_init(&p->osc, 0, sample_rate);
    /* init(): tapsPerChannel = (int)ceil(filter.length / 2) with INTEGER division (:76) */
    int taps_per_channel = (int)ceil((double)(filter_len / 2));
    p->len = 2 * taps_per_channel * 2;
    p->filter = (float *)calloc((size_t)p->len, sizeof(float));
    p->serpentine = (float *)calloc((size_t)p->len, sizeof(float));
    p->product = (float *)calloc((size_t)p->len, sizeof(float));
    int fp = 0;
    for (int cp = 0; cp < filter_len && fp + 1 < p->len; cp++) {
        p->filter[fp++] = filter[cp];
        p->filter[fp++] = filter[cp];
    }
    p->top_block = 1;
    p->fs4_pointer = 0;
    return p;
}

void orc_two_channel_destroy(orc_two_channel *p)
{
    if (!p) return;
    free(p->filter);
    free(p->serpentine);
    free(p->product);
    free(p->c1);
    free(p->c2);
    free(p);
}

void orc_two_channel_set_frequency_offset(orc_two_channel *p, long long offset)
{
    orc_osc_set_frequency(&p->osc, (double)offset, p->sample_rate);
}

/* TwoChannelOutputProcessor.java:98-121 */
void orc_two_channel_process(orc_two_channel *p, const float *results, int n_blocks, int m, float *out)
{
    if (p->cap < 2 * n_blocks) {
        free(p->c1);
        free(p->c2);
        p->cap = 2 * n_blocks;
        p->c1 = (float *)malloc(sizeof(float) * (size_t)p->cap);
        p->c2 = (float *)malloc(sizeof(float) * (size_t)p->cap);
    }
    orc_get_channel(results, n_blocks, m, p->bin1, p->c1);
    orc_get_channel(results, n_blocks, m, p->bin2, p->c2);

    /* TwoChannelSynthesizerM2.process :90-158 */
    float ifft[4];
    for (int x = 0; x < 2 * n_blocks; x += 2) {
        float a_i = p->c1[x], a_q = p->c1[x + 1], b_i = p->c2[x], b_q = p->c2[x + 1];
        /* FloatFFT_1D(2).complexInverse(buf, true): 2-point butterfly then scale by 1/2 (exact) */
        ifft[0] = (a_i + b_i) * 0.5f;
        ifft[1] = (a_q + b_q) * 0.5f;
        ifft[2] = (a_i - b_i) * 0.5f;
        ifft[3] = (a_q - b_q) * 0.5f;
        memmove(p->serpentine + 4, p->serpentine, sizeof(float) * (size_t)(p->len - 4));
        if (p->top_block) {
            memcpy(p->serpentine, ifft, sizeof(float) * 4);
        } else {
            p->serpentine[2] = ifft[0];
            p->serpentine[3] = ifft[1];
            p->serpentine[0] = ifft[2];
            p->serpentine[1] = ifft[3];
        }
        for (int y = 0; y < p->len; y++) p->product[y] = p->serpentine[y] * p->filter[y];
        float acc_i = 0.0f, acc_q = 0.0f;
        for (int y = 0; y < p->len; y += 2) {
            acc_i += p->product[y];
            acc_q += p->product[y + 1];
        }
        out[x] = acc_i;
        out[x + 1] = acc_q;
        p->top_block = !p->top_block;
    }

    /* FS4DownConverter.java:32-68 */
    for (int x = 0; x < 2 * n_blocks; x += 2) {
        float real;
        switch (p->fs4_pointer) {
            case 1:
                real = out[x];
                out[x] = out[x + 1];
                out[x + 1] = -real;
                break;
            case 2:
                out[x] = -out[x];
                out[x + 1] = -out[x + 1];
                break;
            case 3:
                real = out[x];
                out[x] = -out[x + 1];
                out[x + 1] = real;
                break;
            default:
                break;
        }
        p->fs4_pointer++;
        if (p->fs4_pointer >= 4) p->fs4_pointer = 0;
    }

    /* frequency-correction mixer is applied unconditionally here (TwoChannelOutputProcessor.java:113) */
    orc_osc_mix(&p->osc, out, 2 * n_blocks);
    orc_apply_gain(out, 2 * n_blocks, p->gain);
}

/* ---------------------------------------------------------------- tuner sample converters
 * ByteSampleConverter.java:21-35 (LOOKUP_VALUES[x] = (float)(x - 127) / 128.0f),
 * SignedByteSampleConverter.java:21-35 ((float)((byte)x) / 128.0f),
 * ConversionUtils.java:22-34 (little-endian short / (float)Short.MAX_VALUE) */
void orc_convert_u8(const uint8_t *in, int n, float *out)
{
    for (int x = 0; x < n; x++) out[x] = (float)((int)in[x] - 127) / 128.0f;
}

void orc_convert_s8(const int8_t *in, int n, float *out)
{
    for (int x = 0; x < n; x++) out[x] = (float)in[x] / 128.0f;
}

void orc_convert_s16le(const uint8_t *in, int n, float *out)
{
    for (int x = 0; x < n; x++) {
        short v = (short)((unsigned)in[2 * x] | ((unsigned)in[2 * x + 1] << 8));
        out[x] = (float)v / (float)32767;
    }
}
