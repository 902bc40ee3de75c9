/*
 * sdr_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, never shipped, never on the product path).
 *
 * A statement-level C restatement of the arithmetic of sdrtrunk's (smyers119/sdrtrunk, Java) DSP hot
 * path: polyphase channelizer -> per-channel decimation/FIR -> FM discriminator / DQPSK symbol timing
 * recovery.  Every function cites the reference file:line it follows ("J/" =
 * src/main/java/io/github/dsheirer/).  Same float/double types and operation order as the Java; fmaf
 * only where Java calls Math.fma; compiled with -O2 -ffp-contract=off and no fast-math.
 *
 * PARITY UNPINNED: the reference ships no golden vectors / known-answer tests for this path
 *  and no JVM exists in the build container, so fidelity to the Java is by
 * inspection only.  Third-party arithmetic is replaced as follows: JTransforms FloatFFT_1D ->
 * orc_ifft_* (own float32 mixed-radix FFT + a float64 direct DFT to check it); commons-math3 FastMath
 * sin/cos/atan/sqrt -> glibc libm in double, rounded to float where the Java casts.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
 * this library, and only as the checker / baseline.
 */
#ifndef SDR_ORACLE_H
#define SDR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- filter design (a1, a9) */
enum { ORC_WIN_HAMMING = 0, ORC_WIN_BLACKMAN = 1 };

int orc_window(int type, int length, double *out);
int orc_kaiser(int length, double attenuation, double *out);
int orc_kaiser_sinc(int length, double cutoff, double attenuation, float *out);
double orc_evaluate(const float *filter, int length, double frequency);
/* returns number of taps written (channels * actual taps-per-channel) or <0 on design failure */
int orc_sinc_m2_channelizer(double channel_bandwidth, int channels, int taps_per_channel, float *out,
                            int out_capacity);
int orc_sinc_m2_synthesizer(double channel_sample_rate, double channel_bandwidth, int channels,
                            int taps_per_channel, float *out);
int orc_half_band(int length, int window_type, float *out);
/* Remez / Parks-McClellan low-pass designer (orc_remez.c): FIRFilterSpecification.lowPassBuilder()...build() +
 * FilterFactory.getTaps.  order < 6: estimated (estimateFilterOrder); odd_length -1 unset / 0 / 1.  Returns the number
 * of taps, -1 if the design does not converge (getTaps returns null), -2 if capacity is too small. */
int orc_remez_estimate_order(double sampleRate, double frequency1, double frequency2, double passBandRipple,
                             double stopBandRipple);
int orc_remez_low_pass(double sampleRate, double passBandEnd, double stopBandStart, double passBandRipple,
                       double stopBandRipple, int order, int odd_length, int gridDensity, float *out, int capacity);

/* ---------------------------------------------------------------- inverse FFT (a4) */
typedef struct orc_fft orc_fft;
orc_fft *orc_fft_create(int n);
void orc_fft_destroy(orc_fft *f);
/* in place, interleaved complex float32, e^{+j...}, scaled by 1/n (JTransforms complexInverse(a,true)) */
void orc_ifft_f32(orc_fft *f, float *a);
/* direct O(n^2) DFT in float64 rounded to float32: the arithmetic-independent check */
void orc_idft_f64(int n, const float *in, float *out);

/* ---------------------------------------------------------------- channelizer (a2, a3, a4) */
typedef struct orc_channelizer orc_channelizer;
orc_channelizer *orc_chan_create(const float *taps, int n_taps, int channel_count);
void orc_chan_destroy(orc_channelizer *c);
/* feed n_floats interleaved I/Q; appends whole blocks (each 2*M floats, IFFT applied) to out; returns
 * number of blocks produced.  out must hold (pending + n_floats)/M blocks. use_f64_dft selects the
 * direct DFT instead of the float32 FFT. */
int orc_chan_receive(orc_channelizer *c, const float *samples, int n_floats, float *out, int use_f64_dft);
/* filter-bank stage only (no IFFT): the accumulators after the top/middle permutation */
int orc_chan_receive_raw(orc_channelizer *c, const float *samples, int n_floats, float *out);

/* ---------------------------------------------------------------- channel calculator (a5) */
typedef struct {
    double sample_rate;
    int channel_count;
    double center_frequency;
    double oversampling;
} orc_calc;
/* returns count (1..), or <0: -1 out of range, -2 wrap-around, -3 internal */
int orc_calc_channel_indexes(const orc_calc *c, long long frequency, int bandwidth, int *indexes, int cap);
long long orc_calc_center_frequency_for_indexes(const orc_calc *c, const int *indexes, int n);

/* ---------------------------------------------------------------- output processors (a6, a7, a8) */
typedef struct {
    float angle_i, angle_q; /* per-sample rotation */
    float cur_i, cur_q;
} orc_oscillator;
void orc_osc_init(orc_oscillator *o, double frequency, double sample_rate);
void orc_osc_set_frequency(orc_oscillator *o, double frequency, double sample_rate);
void orc_osc_mix(orc_oscillator *o, float *samples, int n_floats);

/* gather bin `bin` from n_blocks result arrays of 2*M floats each into out[2*n_blocks] */
void orc_get_channel(const float *results, int n_blocks, int m, int bin, float *out);
void orc_apply_gain(float *samples, int n_floats, double gain);

typedef struct orc_one_channel orc_one_channel;
orc_one_channel *orc_one_channel_create(double sample_rate, int bin, double gain);
void orc_one_channel_destroy(orc_one_channel *p);
void orc_one_channel_set_frequency_offset(orc_one_channel *p, long long offset);
void orc_one_channel_process(orc_one_channel *p, const float *results, int n_blocks, int m, float *out);

typedef struct orc_two_channel orc_two_channel;
orc_two_channel *orc_two_channel_create(double sample_rate, int bin1, int bin2, const float *filter,
                                        int filter_len, double gain);
void orc_two_channel_destroy(orc_two_channel *p);
void orc_two_channel_set_frequency_offset(orc_two_channel *p, long long offset);
void orc_two_channel_process(orc_two_channel *p, const float *results, int n_blocks, int m, float *out);

/* ---------------------------------------------------------------- tuner sample converters (section 8f #1) */
void orc_convert_u8(const uint8_t *in, int n, float *out);
void orc_convert_s8(const int8_t *in, int n, float *out);
void orc_convert_s16le(const uint8_t *in, int n, float *out);

/* Airspy native buffers (12-bit real samples at twice the complex rate): unpack -> DCRemovalFilter(0.01f) ->
 * HilbertTransform.  n_bytes: 2 per sample (unpacked) or 3 per 2 samples (packed); returns the number of floats written
 * to out (one per real sample = interleaved I/Q of half as many complex samples).  State carries over between calls. */
typedef struct orc_airspy orc_airspy;
orc_airspy *orc_airspy_create(void);
void orc_airspy_destroy(orc_airspy *a);
int orc_airspy_convert(orc_airspy *a, const uint8_t *bytes, int n_bytes, int packed, float *out);

/* ---------------------------------------------------------------- decimation (a9, a10) */
typedef struct orc_halfband orc_halfband;
orc_halfband *orc_halfband_create(const float *coefficients, int length);
void orc_halfband_destroy(orc_halfband *h);
/* complex: n_floats multiple of 4, out n_floats/2; real: multiple of 2. returns out floats or <0 */
int orc_halfband_decimate_complex(orc_halfband *h, const float *samples, int n_floats, float *out);
int orc_halfband_decimate_real(orc_halfband *h, const float *samples, int n_floats, float *out);

typedef struct orc_decimator orc_decimator;
orc_decimator *orc_decimator_create(int rate); /* 0,2,4,...,1024; NULL if unsupported */
void orc_decimator_destroy(orc_decimator *d);
int orc_decimator_complex(orc_decimator *d, const float *samples, int n_floats, float *out);
int orc_decimator_real(orc_decimator *d, const float *samples, int n_floats, float *out);

/* ---------------------------------------------------------------- FIR (a11) */
typedef struct orc_fir orc_fir;
orc_fir *orc_fir_create(const float *taps, int n, float gain);
void orc_fir_destroy(orc_fir *f);
float orc_fir_filter(orc_fir *f, float sample);
void orc_fir_filter_real(orc_fir *f, const float *in, int n, float *out);
typedef struct {
    orc_fir *i, *q;
} orc_cfir;
orc_cfir *orc_cfir_create(const float *taps, int n, float gain);
void orc_cfir_destroy(orc_cfir *f);
void orc_cfir_filter(orc_cfir *f, const float *in, int n_floats, float *out);

/* ---------------------------------------------------------------- FM (a12) */
typedef struct {
    float prev_i, prev_q, gain;
} orc_fm;
void orc_fm_init(orc_fm *f, float gain);
float orc_fm_demodulate(orc_fm *f, float i, float q);
void orc_fm_demodulate_buffer(orc_fm *f, const float *iq, int n_floats, float *out);

typedef struct {
    double alpha, one_minus_alpha, output; /* SinglePoleIirFilter */
    double power, threshold;
    int state, ramp_threshold, ramp_count;
    int squelch_changed;
} orc_squelch;
void orc_squelch_init(orc_squelch *s, double alpha, double threshold_db, int ramp);
void orc_squelch_process(orc_squelch *s, double inphase, double quadrature);
typedef struct {
    orc_fm fm;
    orc_squelch sq;
    int squelch_changed;
} orc_sqfm;
void orc_sqfm_init(orc_sqfm *s, double alpha, double threshold_db, int ramp);
void orc_sqfm_demodulate_buffer(orc_sqfm *s, const float *iq, int n_floats, float *out);

/* ---------------------------------------------------------------- AGC (a13) */
void orc_agc_block(const float *in, int n_floats, float *out);

/* ---------------------------------------------------------------- PSK (a14-a17) */
enum { ORC_PSK_DECISION_DIRECTED = 0, ORC_PSK_GARDNER = 1 };
typedef struct orc_psk orc_psk;
orc_psk *orc_psk_create(int kind, double sample_rate, double symbol_rate, double pll_bandwidth,
                        float sample_counter_gain);
void orc_psk_destroy(orc_psk *p);
/* feeds n_floats interleaved I/Q; writes one byte per symbol (Dibit.getValue 0..3); returns symbol count.
 * taps (optional, may be NULL): per symbol 6 doubles {soft_i, soft_q, detected_sps, pll_frequency (radians per sample),
 * sampling_point, pll_error}: the listener tap points of DQPSKDecisionDirectedDemodulatorInstrumented.java:74-108 */
int orc_psk_receive(orc_psk *p, const float *iq, int n_floats, uint8_t *dibits, double *taps);
void orc_psk_correct_inversion(orc_psk *p, double correction);
void orc_psk_reset_pll(orc_psk *p);
void orc_psk_get_state(const orc_psk *p, double *phase, double *freq, float *sampling_point, float *detected_sps);

/* 4 dibits per byte, MSB first (DibitToByteBufferAssembler.java:58-93); returns whole bytes written */
int orc_pack_dibits(const uint8_t *dibits, int n, uint8_t *out);

/* ---------------------------------------------------------------- sync detection + PLL inversion feedback (8f #3) */
enum { ORC_SYNC_P25_PHASE1 = 1, ORC_SYNC_P25_PHASE2 = 2, ORC_SYNC_P25_PHASE2_FRAMED = 3 };
/* event of one dibit: bits 0-2 one of these, bits 3-5 the primary detector's bit errors (SYNC only) */
enum {
    ORC_SYNC_EVENT_NONE = 0,
    ORC_SYNC_EVENT_SYNC = 1,
    ORC_SYNC_EVENT_INVERSION_90_CW = 2,
    ORC_SYNC_EVENT_INVERSION_90_CCW = 3,
    ORC_SYNC_EVENT_INVERSION_180 = 4,
    ORC_SYNC_EVENT_LOST = 5
};
typedef struct orc_sync orc_sync;
orc_sync *orc_sync_create(int kind, double sample_rate);
void orc_sync_destroy(orc_sync *s);
int orc_sync_delay(const orc_sync *s);
int orc_sync_receive(orc_sync *s, int dibit, double *correction);
/* APCO25 Phase 2 super-frame fragment detector (P25P2SuperFrameDetector + its P25P2SyncDetector): the reference's
 * complete Phase 2 framing and PLL inversion feedback.  Event bits of one dibit: */
enum {
    ORC_P2_EVENT_FRAGMENT = 1,      /* broadcastFragment: the last 720 dibits are a super-frame fragment */
    ORC_P2_EVENT_SYNC_LOSS = 2,     /* broadcastSyncLoss was called */
    ORC_P2_EVENT_INVERSION = 4,     /* correctInversion was called; bits 3-4: 1 = 90 CW, 2 = 90 CCW, 3 = 180 */
    ORC_P2_EVENT_SYNCHRONIZED = 32  /* mSynchronized after this dibit */
};
typedef struct orc_p2_framer orc_p2_framer;
orc_p2_framer *orc_p2_framer_create(double sample_rate);
void orc_p2_framer_destroy(orc_p2_framer *f);
int orc_p2_framer_receive(orc_p2_framer *f, int dibit, double *correction);
void orc_psk_attach_p2_framer(orc_psk *p, orc_p2_framer *f);

/* the demodulator then feeds every dibit to s (as the framer's listener does, synchronously after the PLL update of
 * that symbol), applies requested corrections with correctInversion and reports dibit | event << 2 per symbol */
void orc_psk_attach_sync(orc_psk *p, orc_sync *s);

/* ---------------------------------------------------------------- whole chains (a18) used as CPU baseline */
typedef struct orc_p25_chain orc_p25_chain;
/* kind: 0 = C4FM (FIR + AGC + DD, BW300, gain .3, 4800), 1 = LSM (no FIR, AGC, Gardner BW200 .3, 4800),
 *       2 = HDQPSK (FIR + AGC + Gardner BW300 .1, 6000), 3 = DMR (FIR + AGC + DD, BW300 .4, 4800) */
orc_p25_chain *orc_p25_chain_create(int kind, double sample_rate, const float *fir_taps, int n_taps);
void orc_p25_chain_destroy(orc_p25_chain *c);
/* consumes whole 1024-complex-sample buffers only (the assembler framing); n_floats multiple of 2048 */
int orc_p25_chain_receive(orc_p25_chain *c, const float *iq, int n_floats, uint8_t *dibits, float *agc_out);
/* attaches a sync detector (owned by the chain) to the chain's demodulator; sync_kind ORC_SYNC_P25_PHASE2_FRAMED
 * attaches the Phase 2 super-frame fragment detector instead */
int orc_p25_chain_attach_sync(orc_p25_chain *c, int sync_kind, double sample_rate);

#ifdef __cplusplus
}
#endif
#endif
