/* CPU ORACLE -- test infrastructure only (see sdr_oracle.h).
 * DQPSK demodulators: Costas loop, interpolating sample buffer, decision-directed and Gardner timing
 * recovery, symbol evaluators, dibit packing, and the P25 decoder front-end chains.
 * Follows the J/dsp/psk/ classes, J/dsp/psk/pll/CostasLoop.java, J/dsp/filter/interpolator/RealInterpolator.java,
 * J/dsp/symbol/{Dibit,DibitToByteBufferAssembler}.java, J/module/decode/p25/phase{1,2}/P25P*Decoder*.java. */
#include "sdr_oracle.h"
#include "../include/sdr_mmse_taps.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static const double ORC_PI = 3.14159265358979323846;
#define ORC_TWO_PI (2.0 * ORC_PI)

/* Dibit.java:27-30 getValue() */
enum { D00_PLUS_1 = 0, D01_PLUS_3 = 1, D10_MINUS_1 = 2, D11_MINUS_3 = 3 };

typedef struct { float i, q; } cpx;

static inline float mul_i(float ia, float qa, float ib, float qb) { return (ia * ib) - (qa * qb); }
static inline float mul_q(float ia, float qa, float ib, float qb) { return (qa * ib) + (ia * qb); }

/* Complex.java:215-253 normalize(): magnitude = (float)sqrt((double)(i*i + q*q)); multiply by 1.0f/magnitude */
static inline void normalize(cpx *c)
{
    float norm = (c->i * c->i) + (c->q * c->q);
    float magnitude = (float)sqrt((double)norm);
    if (magnitude != 0) {
        float s = (float)(1.0f / magnitude);
        c->i *= s;
        c->q *= s;
    }
}

static inline float clipf(float v, float max)
{
    if (v > max) return max;
    if (v < -max) return -max;
    return v;
}

static inline float normalize_error(float e, float max)
{
    if (isnan(e)) return 0.0f;
    return clipf(e, max);
}

struct orc_psk {
    int kind;
    /* CostasLoop.java:48-70 */
    double loop_phase, loop_frequency, max_loop_frequency, alpha_gain, beta_gain;
    /* InterpolatingSampleBuffer.java:36-70 */
    float *delay_i, *delay_q;
    int pointer, twice_sps;
    float sampling_point, sample_counter_gain, detected_sps, detected_sps_gain, max_sps, min_sps;
    /* demodulator state */
    cpx prev_a, prev_b; /* DD: previousPreceding/previousCurrent; Gardner: previousMiddle/previousCurrent */
    cpx gardner_previous_symbol;
    cpx rot[4]; /* rotate-from +45, +135, -45, -135 */
    orc_sync *sync; /* optional dibit listener (not owned) */
    orc_p2_framer *p2_framer;
};

orc_psk *orc_psk_create(int kind, double sample_rate, double symbol_rate, double pll_bandwidth, float sample_counter_gain)
{
    orc_psk *p = (orc_psk *)calloc(1, sizeof(*p));
    p->kind = kind;
    /* CostasLoop.java:64-70,109-115 ; PLLBandwidth.java:34-37 */
    p->max_loop_frequency = ORC_TWO_PI * (symbol_rate / 2.0) / sample_rate;
    double damping = sqrt(2.0) / 2.0;
    double bandwidth = ORC_TWO_PI / pll_bandwidth;
    p->alpha_gain = (4.0 * damping * bandwidth) / (1.0 + (2.0 * damping * bandwidth) + (bandwidth * bandwidth));
    p->beta_gain = (4.0 * bandwidth * bandwidth) / (1.0 + (2.0 * damping * bandwidth) + (bandwidth * bandwidth));
    /* InterpolatingSampleBuffer.java:58-70 ; P25P1Decoder.java:135-138 samplesPerSymbol = (float)(fs/symbolRate) */
    float sps = (float)(sample_rate / symbol_rate);
    p->sampling_point = sps;
    p->detected_sps = sps;
    p->max_sps = sps * (1.0f + 0.02f);
    p->min_sps = sps * (1.0f - 0.02f);
    p->twice_sps = (int)floor(2.0 * sps);
    p->delay_i = (float *)calloc((size_t)(2 * p->twice_sps), sizeof(float));
    p->delay_q = (float *)calloc((size_t)(2 * p->twice_sps), sizeof(float));
    p->sample_counter_gain = sample_counter_gain;
    p->detected_sps_gain = 0.1f * sample_counter_gain * sample_counter_gain;
    /* DQPSK*SymbolEvaluator.java:24-27 Complex.fromAngle */
    double a[4] = {-1.0 * ORC_PI / 4.0, -3.0 * ORC_PI / 4.0, 1.0 * ORC_PI / 4.0, 3.0 * ORC_PI / 4.0};
    for (int k = 0; k < 4; k++) {
        p->rot[k].i = (float)cos(a[k]);
        p->rot[k].q = (float)sin(a[k]);
    }
    return p;
}

void orc_psk_destroy(orc_psk *p)
{
    if (!p) return;
    free(p->delay_i);
    free(p->delay_q);
    free(p);
}

/* CostasLoop.java:91-104 */
void orc_psk_correct_inversion(orc_psk *p, double correction)
{
    p->loop_frequency += correction;
    while (p->loop_frequency > p->max_loop_frequency) p->loop_frequency -= 2.0 * p->max_loop_frequency;
    while (p->loop_frequency < -p->max_loop_frequency) p->loop_frequency += 2.0 * p->max_loop_frequency;
}

void orc_psk_attach_sync(orc_psk *p, orc_sync *s) { p->sync = s; }
void orc_psk_attach_p2_framer(orc_psk *p, orc_p2_framer *f) { p->p2_framer = f; }

/* CostasLoop.java:224-229 */
void orc_psk_reset_pll(orc_psk *p)
{
    p->loop_phase = 0.0;
    p->loop_frequency = 0.0;
}

void orc_psk_get_state(const orc_psk *p, double *phase, double *freq, float *sampling_point, float *detected_sps)
{
    if (phase) *phase = p->loop_phase;
    if (freq) *freq = p->loop_frequency;
    if (sampling_point) *sampling_point = p->sampling_point;
    if (detected_sps) *detected_sps = p->detected_sps;
}

/* CostasLoop.java:178-203 adjust() (the once-a-second frequency-error broadcast has no effect on samples) */
static void pll_adjust(orc_psk *p, double phase_error)
{
    p->loop_frequency += (p->beta_gain * phase_error);
    p->loop_phase += p->loop_frequency + (p->alpha_gain * phase_error);
    if (p->loop_phase > ORC_TWO_PI) p->loop_phase -= ORC_TWO_PI;
    if (p->loop_phase < -ORC_TWO_PI) p->loop_phase += ORC_TWO_PI;
    if (p->loop_frequency > p->max_loop_frequency) p->loop_frequency = p->max_loop_frequency;
    if (p->loop_frequency < -p->max_loop_frequency) p->loop_frequency = -p->max_loop_frequency;
}

/* RealInterpolator.java:41-59 (gain 1.0f) */
static float interpolate(const float *samples, int offset, float mu)
{
    int index = (int)(SDR_MMSE_NSTEPS * mu);
    const float *t = SDR_MMSE_TAPS + 8 * index;
    float acc = (t[7] * samples[offset]);
    acc += (t[6] * samples[offset + 1]);
    acc += (t[5] * samples[offset + 2]);
    acc += (t[4] * samples[offset + 3]);
    acc += (t[3] * samples[offset + 4]);
    acc += (t[2] * samples[offset + 5]);
    acc += (t[1] * samples[offset + 6]);
    acc += (t[0] * samples[offset + 7]);
    return acc * 1.0f;
}

/* InterpolatingSampleBuffer.java:185-214 getInphase/getQuadrature */
static float interp_at(const orc_psk *p, const float *line, float interpolation)
{
    if (interpolation < 1.0f) return interpolate(line, p->pointer, interpolation);
    int offset = (int)floor((double)interpolation);
    return interpolate(line, p->pointer + offset, interpolation - offset);
}

/* InterpolatingSampleBuffer.java:106-124 resetAndAdjust */
static void reset_and_adjust(orc_psk *p, float timing_error)
{
    p->detected_sps = p->detected_sps + (timing_error * p->detected_sps_gain);
    if (p->detected_sps > p->max_sps) p->detected_sps = p->max_sps;
    if (p->detected_sps < p->min_sps) p->detected_sps = p->min_sps;
    p->sampling_point += (p->detected_sps + (timing_error * p->sample_counter_gain));
}

/* quadrant slicer shared by both evaluators (DQPSKDecisionDirectedSymbolEvaluator.java:61-95,
 * DQPSKGardnerSymbolEvaluator.java:71-99); returns rotated quadrature and sets dibit and use_less_than (the DD
 * evaluator compares preceding.q '<' current.q for the +/-135 symbols and '>' for +/-45) */
static float slice(const orc_psk *p, cpx cur, int *dibit, int *use_less_than)
{
    int r;
    if (cur.q > 0.0f) {
        if (cur.i > 0.0f) { *dibit = D00_PLUS_1; r = 0; *use_less_than = 0; }
        else { *dibit = D01_PLUS_3; r = 1; *use_less_than = 1; }
    } else {
        if (cur.i > 0.0f) { *dibit = D10_MINUS_1; r = 2; *use_less_than = 0; }
        else { *dibit = D11_MINUS_3; r = 3; *use_less_than = 1; }
    }
    /* only the quadrature of the rotated evaluation symbol is used afterwards */
    return mul_q(cur.i, cur.q, p->rot[r].i, p->rot[r].q);
}

/* DQPSKDecisionDirectedDemodulator.java:50-89 */
static int calculate_symbol_dd(orc_psk *p, double *taps)
{
    /* InterpolatingSampleBuffer.java:148-165 */
    cpx preceding = {p->delay_i[p->pointer + 3], p->delay_q[p->pointer + 3]};
    cpx current = {interp_at(p, p->delay_i, p->sampling_point), interp_at(p, p->delay_q, p->sampling_point)};

    cpx prec_sym = {mul_i(preceding.i, preceding.q, p->prev_a.i, -p->prev_a.q),
                    mul_q(preceding.i, preceding.q, p->prev_a.i, -p->prev_a.q)};
    cpx cur_sym = {mul_i(current.i, current.q, p->prev_b.i, -p->prev_b.q),
                   mul_q(current.i, current.q, p->prev_b.i, -p->prev_b.q)};
    normalize(&prec_sym);
    normalize(&cur_sym);

    int dibit, lt;
    float rotated_q = slice(p, cur_sym, &dibit, &lt);
    float polarity = lt ? (prec_sym.q < cur_sym.q ? 1.0f : -1.0f) : (prec_sym.q > cur_sym.q ? 1.0f : -1.0f);
    float error_normalized = normalize_error(rotated_q, 0.3f);
    float phase_error = -error_normalized;
    float timing_error = error_normalized * polarity;

    reset_and_adjust(p, timing_error);
    pll_adjust(p, (double)clipf(phase_error, 0.5f));
    p->prev_a = preceding;
    p->prev_b = current;
    if (taps) { /* the tap points of DQPSK*DemodulatorInstrumented.java:74-108, taken at the end of calculateSymbol */
        taps[0] = cur_sym.i;
        taps[1] = cur_sym.q;
        taps[2] = p->detected_sps;
        taps[3] = p->loop_frequency;
        taps[4] = p->sampling_point;
        taps[5] = phase_error;
    }
    return dibit;
}

/* DQPSKGardnerDemodulator.java:48-89 ; DQPSKGardnerSymbolEvaluator.java:61-105 */
static int calculate_symbol_gardner(orc_psk *p, double *taps)
{
    /* the two roles are flip-flopped on purpose (DQPSKGardnerDemodulator.java:50-57) */
    cpx middle = {interp_at(p, p->delay_i, p->sampling_point), interp_at(p, p->delay_q, p->sampling_point)};
    float half_sps = p->detected_sps / 2.0f; /* InterpolatingSampleBuffer.java:171-179 */
    cpx current = {interp_at(p, p->delay_i, half_sps), interp_at(p, p->delay_q, half_sps)};

    cpx mid_sym = {mul_i(middle.i, middle.q, p->prev_a.i, -p->prev_a.q),
                   mul_q(middle.i, middle.q, p->prev_a.i, -p->prev_a.q)};
    cpx cur_sym = {mul_i(current.i, current.q, p->prev_b.i, -p->prev_b.q),
                   mul_q(current.i, current.q, p->prev_b.i, -p->prev_b.q)};
    normalize(&mid_sym);
    normalize(&cur_sym);

    float error_i = (p->gardner_previous_symbol.i - cur_sym.i) * mid_sym.i;
    float error_q = (p->gardner_previous_symbol.q - cur_sym.q) * mid_sym.q;
    float timing_error = normalize_error(error_i + error_q, .3f);
    p->gardner_previous_symbol = cur_sym;

    int dibit, lt;
    float rotated_q = slice(p, cur_sym, &dibit, &lt);
    float phase_error = normalize_error(-rotated_q, 0.3f);

    reset_and_adjust(p, timing_error);
    pll_adjust(p, (double)phase_error);
    p->prev_a = middle;
    p->prev_b = current;
    if (taps) { /* the tap points of DQPSK*DemodulatorInstrumented.java:74-108, taken at the end of calculateSymbol */
        taps[0] = cur_sym.i;
        taps[1] = cur_sym.q;
        taps[2] = p->detected_sps;
        taps[3] = p->loop_frequency;
        taps[4] = p->sampling_point;
        taps[5] = phase_error;
    }
    return dibit;
}

/* PSKDemodulator.java:83-117 */
int orc_psk_receive(orc_psk *p, const float *iq, int n_floats, uint8_t *dibits, double *taps)
{
    int n_symbols = 0;
    for (int x = 0; x < n_floats; x += 2) {
        /* CostasLoop.java:135-166 increment + getCurrentVector (Complex.setAngle, Complex.java:383-387) */
        p->loop_phase += p->loop_frequency;
        if (p->loop_phase > ORC_TWO_PI) p->loop_phase -= ORC_TWO_PI;
        if (p->loop_phase < -ORC_TWO_PI) p->loop_phase += ORC_TWO_PI;
        float vi = (float)cos(p->loop_phase);
        float vq = (float)sin(p->loop_phase);
        float si = mul_i(iq[x], iq[x + 1], vi, vq);
        float sq = mul_q(iq[x], iq[x + 1], vi, vq);

        /* InterpolatingSampleBuffer.java:76-89 receive */
        p->sampling_point--;
        p->delay_i[p->pointer] = si;
        p->delay_i[p->pointer + p->twice_sps] = si;
        p->delay_q[p->pointer] = sq;
        p->delay_q[p->pointer + p->twice_sps] = sq;
        p->pointer++;
        p->pointer = p->pointer % p->twice_sps;

        if (p->sampling_point < 1.0f) {
            double *t = taps ? taps + 6 * (size_t)n_symbols : NULL;
            int d = (p->kind == ORC_PSK_GARDNER) ? calculate_symbol_gardner(p, t) : calculate_symbol_dd(p, t);
            /* calculateSymbol ends with broadcast(dibit): the framer's sync detector runs synchronously and may call
             * correctInversion before the next sample is processed (P25P1SyncDetector.java:150-154) */
            if (p->sync) {
                double correction = 0.0;
                int event = orc_sync_receive(p->sync, d, &correction);
                if ((event & 7) >= ORC_SYNC_EVENT_INVERSION_90_CW && (event & 7) <= ORC_SYNC_EVENT_INVERSION_180)
                    orc_psk_correct_inversion(p, correction);
                d |= event << 2;
            }
            if (p->p2_framer) { /* P25P2MessageFramer.receive -> P25P2SuperFrameDetector.receive */
                double correction = 0.0;
                int event = orc_p2_framer_receive(p->p2_framer, d, &correction);
                if (event & ORC_P2_EVENT_INVERSION) orc_psk_correct_inversion(p, correction);
                d |= event << 2;
            }
            dibits[n_symbols++] = (uint8_t)d;
        }
    }
    return n_symbols;
}

/* DibitToByteBufferAssembler.java:58-93 */
int orc_pack_dibits(const uint8_t *dibits, int n, uint8_t *out)
{
    uint8_t current = 0;
    int count = 0, bytes = 0;
    for (int k = 0; k < n; k++) {
        current = (uint8_t)(current << 2);
        current |= (uint8_t)(dibits[k] & 3);
        if (++count >= 4) {
            out[bytes++] = current;
            current = 0;
            count = 0;
        }
    }
    return bytes;
}

/* ---------------------------------------------------------------- decoder front-ends (a18)
 * P25P1DecoderC4FM.java:62-116, P25P1DecoderLSM.java:67-138, P25P2DecoderHDQPSK.java:62-145:
 * filter (C4FM / HDQPSK only) -> ComplexFeedForwardGainControl per 1024-sample buffer -> demodulator. */
struct orc_p25_chain {
    int kind;
    orc_cfir *fir;
    orc_psk *psk;
    float *tmp_a, *tmp_b;
    orc_sync *sync;
    orc_p2_framer *p2_framer;
};

orc_p25_chain *orc_p25_chain_create(int kind, double sample_rate, const float *fir_taps, int n_taps)
{
    orc_p25_chain *c = (orc_p25_chain *)calloc(1, sizeof(*c));
    c->kind = kind;
    if (kind == 0) {
        c->fir = orc_cfir_create(fir_taps, n_taps, 1.0f);
        c->psk = orc_psk_create(ORC_PSK_DECISION_DIRECTED, sample_rate, 4800.0, 300.0, 0.3f);
    } else if (kind == 1) {
        c->psk = orc_psk_create(ORC_PSK_GARDNER, sample_rate, 4800.0, 200.0, 0.3f);
    } else if (kind == 3) { /* DMRDecoder.java:58-131: FIR + AGC + DD, BW_300, gain .4, 4800 */
        c->fir = orc_cfir_create(fir_taps, n_taps, 1.0f);
        c->psk = orc_psk_create(ORC_PSK_DECISION_DIRECTED, sample_rate, 4800.0, 300.0, 0.4f);
    } else {
        c->fir = orc_cfir_create(fir_taps, n_taps, 1.0f);
        c->psk = orc_psk_create(ORC_PSK_GARDNER, sample_rate, 6000.0, 300.0, 0.1f);
    }
    c->tmp_a = (float *)malloc(sizeof(float) * 2048);
    c->tmp_b = (float *)malloc(sizeof(float) * 2048);
    return c;
}

int orc_p25_chain_attach_sync(orc_p25_chain *c, int sync_kind, double sample_rate)
{
    orc_sync_destroy(c->sync);
    orc_p2_framer_destroy(c->p2_framer);
    c->sync = NULL;
    c->p2_framer = NULL;
    if (sync_kind == ORC_SYNC_P25_PHASE2_FRAMED) {
        c->p2_framer = orc_p2_framer_create(sample_rate);
        orc_psk_attach_sync(c->psk, NULL);
        orc_psk_attach_p2_framer(c->psk, c->p2_framer);
        return 0;
    }
    c->sync = orc_sync_create(sync_kind, sample_rate);
    orc_psk_attach_p2_framer(c->psk, NULL);
    orc_psk_attach_sync(c->psk, c->sync);
    return c->sync ? 0 : -1;
}

void orc_p25_chain_destroy(orc_p25_chain *c)
{
    if (!c) return;
    orc_cfir_destroy(c->fir);
    orc_psk_destroy(c->psk);
    orc_sync_destroy(c->sync);
    orc_p2_framer_destroy(c->p2_framer);
    free(c->tmp_a);
    free(c->tmp_b);
    free(c);
}

int orc_p25_chain_receive(orc_p25_chain *c, const float *iq, int n_floats, uint8_t *dibits, float *agc_out)
{
    if (n_floats % 2048 != 0) return -1;
    int n_symbols = 0;
    for (int off = 0; off < n_floats; off += 2048) {
        const float *src = iq + off;
        if (c->fir) {
            orc_cfir_filter(c->fir, src, 2048, c->tmp_a);
            src = c->tmp_a;
        }
        orc_agc_block(src, 2048, c->tmp_b);
        if (agc_out) memcpy(agc_out + off, c->tmp_b, sizeof(float) * 2048);
        n_symbols += orc_psk_receive(c->psk, c->tmp_b, 2048, dibits + n_symbols, NULL);
    }
    return n_symbols;
}
