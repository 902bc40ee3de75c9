/* CPU ORACLE -- test infrastructure only (see sdr_oracle.h).
 * Sync-pattern detection on the dibit stream and the Costas-loop phase-inversion feedback it drives
 * (SURVEY.md section 8f #3).  Follows, statement by statement:
 *   J/bits/MultiSyncPatternMatcher.java:53-108   receive(bit1, bit2): shift register, processors, sync-loss counter
 *   J/bits/SoftSyncDetector.java:46-67           primary pattern, Hamming distance <= threshold
 *   J/bits/SyncDetector.java:41-55               rotated patterns, exact match
 *   J/dsp/symbol/FrameSync.java:27-35            the eight patterns
 *   J/dsp/symbol/DibitDelayBuffer.java:50-57,215-234   delay line preloaded with D00, getAndPut
 *   J/module/decode/p25/phase1/P25P1SyncDetector.java:37-83,141-168   48-bit matcher, threshold 4, corrections
 *   J/module/decode/p25/phase1/P25P1DataUnitDetector.java:37-41,93-110   33-dibit delay in front of the detector
 *   J/module/decode/p25/phase2/P25P2SyncDetector.java:40-84   40-bit matcher, threshold 4, corrections
 *   J/module/decode/p25/phase2/P25P2SuperFrameDetector.java:66,141-160   160-dibit delay in front of the detector
 * Scope: the detector as it runs while the framer is searching for sync, i.e. fed every dibit.  (The reference's
 * framers stop feeding it while they assemble a message / hold fragment sync; that gating is host-side framing
 * logic and is not restated.) */
#include "sdr_oracle.h"

#include <stdlib.h>

static const double ORC_PI = 3.14159265358979323846;

struct orc_sync {
    /* MultiSyncPatternMatcher */
    uint64_t bits, mask;
    int sync_loss_threshold, bit_count;
    /* processors in the order they are added: primary (soft), 90 CW, 90 CCW, 180 (exact) */
    uint64_t pattern[4];
    int threshold;
    double pll_correction[3];
    /* DibitDelayBuffer */
    uint8_t *delay;
    int delay_len, pointer;
};

orc_sync *orc_sync_create(int kind, double sample_rate)
{
    if (kind != ORC_SYNC_P25_PHASE1 && kind != ORC_SYNC_P25_PHASE2) return NULL;
    orc_sync *s = (orc_sync *)calloc(1, sizeof(*s));
    double symbol_rate;
    int sync_size;
    if (kind == ORC_SYNC_P25_PHASE1) {
        s->pattern[0] = 0x5575F5FF77FFull; /* P25_PHASE1_NORMAL */
        s->pattern[1] = 0x001050551155ull; /* P25_PHASE1_ERROR_90_CW */
        s->pattern[2] = 0xFFEFAFAAEEAAull; /* P25_PHASE1_ERROR_90_CCW */
        s->pattern[3] = 0xAA8A0A008800ull; /* P25_PHASE1_ERROR_180 */
        sync_size = 48;
        s->sync_loss_threshold = 1568; /* P25P1DataUnitID.LOGICAL_LINK_DATA_UNIT_1.getMessageLength() */
        s->delay_len = 57 - 24;        /* DATA_UNIT_DIBIT_LENGTH - SYNC_DIBIT_LENGTH */
        symbol_rate = 4800.0;
    } else {
        s->pattern[0] = 0x575D57F7FFull; /* P25_PHASE2_NORMAL */
        s->pattern[1] = 0x0104015155ull; /* P25_PHASE2_ERROR_90_CW */
        s->pattern[2] = 0xFEFBFEAEAAull; /* P25_PHASE2_ERROR_90_CCW */
        s->pattern[3] = 0xA8A2A80800ull; /* P25_PHASE2_ERROR_180 */
        sync_size = 40;
        s->sync_loss_threshold = 1440;
        s->delay_len = 160;
        symbol_rate = 6000.0;
    }
    s->threshold = 4; /* SYNC_MATCH_THRESHOLD */
    s->mask = (1ull << sync_size) - 1ull; /* (long)(pow(2, syncSize) - 1) */
    /* mPllCorrection = 2.0 * PI * mFrequencyCorrection / mSampleRate with +rate/4, -rate/4, +rate/2 */
    const double correction[3] = {symbol_rate / 4.0, -(symbol_rate / 4.0), symbol_rate / 2.0};
    for (int k = 0; k < 3; k++) s->pll_correction[k] = 2.0 * ORC_PI * correction[k] / sample_rate;
    s->delay = (uint8_t *)calloc((size_t)s->delay_len, 1); /* D00_PLUS_1 = value 0 */
    return s;
}

void orc_sync_destroy(orc_sync *s)
{
    if (!s) return;
    free(s->delay);
    free(s);
}

int orc_sync_delay(const orc_sync *s) { return s->delay_len; }

/* One dibit through the delay buffer into the matcher.  Returns the event of this call (ORC_SYNC_EVENT_*, with the
 * primary detector's bit-error count in bits 3-5) and the PLL correction a PLLPhaseInversionDetector requested. */
int orc_sync_receive(orc_sync *s, int dibit, double *correction)
{
    if (correction) *correction = 0.0;
    /* getAndPut */
    int delayed = s->delay[s->pointer];
    s->delay[s->pointer++] = (uint8_t)(dibit & 3);
    if (s->pointer >= s->delay_len) s->pointer = 0;

    /* MultiSyncPatternMatcher.receive(bit1, bit2); Dibit(bit1, bit2, value): value = 2 * bit1 + bit2.
     * rotateLeft followed by the mask equals a plain shift for masks shorter than 64 bits */
    s->bits = (s->bits << 1) & s->mask;
    if (delayed & 2) s->bits += 1;
    s->bits = (s->bits << 1) & s->mask;
    if (delayed & 1) s->bits += 1;
    s->bit_count += 2;

    int event = ORC_SYNC_EVENT_NONE;
    /* SoftSyncDetector.checkSync */
    uint64_t difference = s->bits ^ s->pattern[0];
    int errors = __builtin_popcountll(difference);
    if (difference == 0 || errors <= s->threshold) {
        event = ORC_SYNC_EVENT_SYNC | (errors << 3);
        s->bit_count = 0;
    }
    /* SyncDetector.checkSync of the three inversion detectors -> correctInversion(mPllCorrection) */
    for (int k = 0; k < 3; k++) {
        if (s->bits == s->pattern[1 + k]) {
            event = ORC_SYNC_EVENT_INVERSION_90_CW + k;
            if (correction) *correction = s->pll_correction[k];
            s->bit_count = 0;
        }
    }
    if (s->bit_count > s->sync_loss_threshold) {
        event = ORC_SYNC_EVENT_LOST;
        s->bit_count = 0;
    }
    return event;
}

/* ---------------------------------------------------------------- APCO25 Phase 2 super-frame fragment detector
 * J/module/decode/p25/phase2/P25P2SuperFrameDetector.java:53-70 (constants, buffers), 105-122 (syncDetected /
 * syncLost), 132-168 (receive), 176-189 (broadcastFragment), 236-302 (checkFragmentSync), with its P25P2SyncDetector
 * (fed only while not synchronized, through the 160-dibit delay buffer that keeps filling either way) and
 * J/module/decode/p25/phase2/P25P2SyncPattern.java:25-57 (bit errors of 20 dibits against the sync pattern).
 * P25P2MessageFramer.receive(Dibit) (:171-174) hands every dibit to this detector, so nothing is gated from outside:
 * this is the reference's complete Phase 2 framing + PLL inversion feedback, statement for statement.  Not restated:
 * what is done with a broadcast fragment (message parsing, scrambling-sequence updates) -- it does not feed back. */
#define P2_FRAGMENT 720
#define P2_DELAY 160

struct orc_p2_framer {
    orc_sync *detector;                 /* P25P2SyncDetector; its delay line is not used here */
    uint8_t fragment[P2_FRAGMENT];      /* mFragmentBuffer */
    int fragment_pointer;
    uint8_t delay[P2_DELAY];            /* mSyncDetectionDelayBuffer */
    int delay_pointer;
    int dibits_processed;
    int synchronized;
    int event;                          /* events raised while the current dibit is processed */
};

orc_p2_framer *orc_p2_framer_create(double sample_rate)
{
    orc_p2_framer *f = (orc_p2_framer *)calloc(1, sizeof(*f));
    f->detector = orc_sync_create(ORC_SYNC_P25_PHASE2, sample_rate);
    return f;
}

void orc_p2_framer_destroy(orc_p2_framer *f)
{
    if (!f) return;
    orc_sync_destroy(f->detector);
    free(f);
}

/* DibitDelayBuffer.getBuffer(start, 20) + P25P2SyncPattern.getBitErrorCount */
static int p2_sync_errors(const orc_p2_framer *f, int start)
{
    int pointer = (f->fragment_pointer + start) % P2_FRAGMENT;
    uint64_t value = 0;
    for (int x = 0; x < 20; x++) {
        value = (value << 2) | f->fragment[pointer++];
        if (pointer >= P2_FRAGMENT) pointer = 0;
    }
    /* per dibit: both bits wrong = 2 errors, one bit wrong = 1: the Hamming distance */
    return __builtin_popcountll(value ^ 0x575D57F7FFull);
}

static void p2_broadcast_fragment(orc_p2_framer *f)
{
    if (f->dibits_processed > P2_FRAGMENT) f->event |= ORC_P2_EVENT_SYNC_LOSS;
    f->dibits_processed = 0;
    f->event |= ORC_P2_EVENT_FRAGMENT;
}

static void p2_check_fragment_sync(orc_p2_framer *f)
{
    if (f->dibits_processed <= 0) return;
    if (f->synchronized) {
        if (p2_sync_errors(f, 360) <= 10 && p2_sync_errors(f, 540) <= 10) {
            p2_broadcast_fragment(f);
            /* syncDetected(...) re-enters here with mDibitsProcessed == 0: nothing happens */
        } else {
            f->synchronized = 0;
        }
        return;
    }
    if (p2_sync_errors(f, 360) <= 4) {
        f->synchronized = 1;
        p2_broadcast_fragment(f);
    } else {
        f->synchronized = 1;
        if (f->dibits_processed > P2_FRAGMENT - 180) f->event |= ORC_P2_EVENT_SYNC_LOSS;
        f->dibits_processed = P2_FRAGMENT - 180;
    }
}

int orc_p2_framer_receive(orc_p2_framer *f, int dibit, double *correction)
{
    if (correction) *correction = 0.0;
    f->event = 0;
    f->dibits_processed++;
    f->fragment[f->fragment_pointer++] = (uint8_t)(dibit & 3);
    if (f->fragment_pointer >= P2_FRAGMENT) f->fragment_pointer = 0;
    if (f->synchronized) {
        f->delay[f->delay_pointer++] = (uint8_t)(dibit & 3);
        if (f->delay_pointer >= P2_DELAY) f->delay_pointer = 0;
        if (f->dibits_processed >= P2_FRAGMENT) p2_check_fragment_sync(f);
    } else {
        int delayed = f->delay[f->delay_pointer];
        f->delay[f->delay_pointer++] = (uint8_t)(dibit & 3);
        if (f->delay_pointer >= P2_DELAY) f->delay_pointer = 0;
        /* P25P2SyncDetector.receive -> MultiSyncPatternMatcher.receive: the matcher of orc_sync without its delay line */
        orc_sync *s = f->detector;
        s->bits = (s->bits << 1) & s->mask;
        if (delayed & 2) s->bits += 1;
        s->bits = (s->bits << 1) & s->mask;
        if (delayed & 1) s->bits += 1;
        s->bit_count += 2;
        uint64_t difference = s->bits ^ s->pattern[0];
        if (difference == 0 || __builtin_popcountll(difference) <= s->threshold) {
            p2_check_fragment_sync(f); /* SoftSyncDetector -> P25P2SuperFrameDetector.syncDetected */
            s->bit_count = 0;
        }
        for (int k = 0; k < 3; k++) {
            if (s->bits == s->pattern[1 + k]) {
                f->event |= ORC_P2_EVENT_INVERSION | ((k + 1) << 3);
                if (correction) *correction = s->pll_correction[k];
                s->bit_count = 0;
            }
        }
        if (s->bit_count > s->sync_loss_threshold) s->bit_count = 0; /* syncLost: rebroadcast only */
    }
    if (f->dibits_processed > 3720) {
        f->dibits_processed -= 3000;
        f->event |= ORC_P2_EVENT_SYNC_LOSS;
    }
    if (f->synchronized) f->event |= ORC_P2_EVENT_SYNCHRONIZED;
    return f->event;
}
