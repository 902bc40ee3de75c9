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
