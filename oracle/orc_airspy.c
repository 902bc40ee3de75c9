/* CPU ORACLE -- test infrastructure only (see sdr_oracle.h).
 * Airspy native-buffer conversion: 12-bit real ADC samples at twice the complex rate ->
 * DC removal -> Hilbert transform (real -> complex, fs/4 + fs/2 translation) -> interleaved I/Q floats.
 * Follows, statement by statement:
 *   J/source/tuner/airspy/AirspySampleConverter.java:27-31,70-84 (convert), 92-110 (convertUnpacked),
 *       118-149 (convertPacked), 155-158 (scale)
 *   J/dsp/filter/dc/DCRemovalFilter.java:52-67 (filter(float), filter(float[]))
 *   J/dsp/filter/hilbert/HilbertTransform.java:56-67 (constructor), 88-132 (filter), 137-142 (insert),
 *       160-196 (generateIndexMap), 223-244 (convertHalfBandToHilbert)
 *   J/dsp/filter/Filters.java:1708-1722 (HALF_BAND_FILTER_47T: data, reproduced verbatim) */
#include "sdr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define HB_LEN 47

/* Filters.HALF_BAND_FILTER_47T */
static const float HALF_BAND_47T[HB_LEN] = {
    -0.000998606272947510f, 0.0f, 0.001695637278417295f, 0.0f, -0.003054430179754289f, 0.0f, 0.005055504379767936f, 0.0f,
    -0.007901319195893647f, 0.0f, 0.011873357051047719f, 0.0f, -0.017411159379930066f, 0.0f, 0.025304817427568772f, 0.0f,
    -0.037225225204559217f, 0.0f, 0.057533286997004301f, 0.0f, -0.102327462004259350f, 0.0f, 0.317034472508947400f, 0.5f,
    0.317034472508947400f, 0.0f, -0.102327462004259350f, 0.0f, 0.057533286997004301f, 0.0f, -0.037225225204559217f, 0.0f,
    0.025304817427568772f, 0.0f, -0.017411159379930066f, 0.0f, 0.011873357051047719f, 0.0f, -0.007901319195893647f, 0.0f,
    0.005055504379767936f, 0.0f, -0.003054430179754289f, 0.0f, 0.001695637278417295f, 0.0f, -0.000998606272947510f};

struct orc_airspy {
    /* DCRemovalFilter(0.01f) */
    float average, ratio;
    /* HilbertTransform */
    float hilbert_filter[HB_LEN];
    float buffer[HB_LEN + 1];
    int buffer_size, buffer_pointer;
    int index_map[HB_LEN / 2 + 1][HB_LEN / 2 + 2];
    int map_height, center_tap_index;
    int invert_flag;
};

orc_airspy *orc_airspy_create(void)
{
    orc_airspy *a = (orc_airspy *)calloc(1, sizeof(*a));
    a->ratio = 0.01f;
    /* convertHalfBandToHilbert */
    int middle = HB_LEN / 2;
    for (int x = 0; x < HB_LEN; x++) {
        if (x < middle) a->hilbert_filter[x] = 2.0f * -fabsf(HALF_BAND_47T[x]);
        else if (x > middle) a->hilbert_filter[x] = 2.0f * fabsf(HALF_BAND_47T[x]);
        else a->hilbert_filter[x] = 2.0f * HALF_BAND_47T[x];
    }
    a->buffer_size = HB_LEN + 1;
    /* generateIndexMap(size = 47) */
    int size = HB_LEN;
    a->map_height = size / 2 + 1;
    int map_width = a->map_height + 1;
    for (int x = 0; x < map_width - 1; x += 2) {
        a->index_map[0][x] = size - 1 - x;
        a->index_map[0][x + 1] = x;
    }
    a->center_tap_index = map_width - 1;
    a->index_map[0][a->center_tap_index] = size / 2;
    for (int x = 1; x < a->map_height; x++) {
        for (int y = 0; y < map_width; y++) {
            a->index_map[x][y] = a->index_map[x - 1][y] + 2;
            if (a->index_map[x][y] >= size) {
                a->index_map[x][y] -= size + 1;
                if (y == a->center_tap_index && a->index_map[x][y] < 0) a->index_map[x][y] = size;
            }
        }
    }
    return a;
}

void orc_airspy_destroy(orc_airspy *a) { free(a); }

/* AirspySampleConverter.scale */
static inline float scale(int value) { return (float)((value & 0xFFF) - 2048) * (1.0f / 2048.0f); }

static inline void insert(orc_airspy *a, float sample)
{
    a->buffer[a->buffer_pointer++] = sample;
    a->buffer_pointer = a->buffer_pointer % a->buffer_size;
}

int orc_airspy_convert(orc_airspy *a, const uint8_t *bytes, int n_bytes, int packed, float *out)
{
    int n = 0;
    if (packed) { /* convertPacked: two samples per 3 bytes; Java bytes are signed, the masks make that irrelevant */
        for (int p = 0; p + 3 <= n_bytes; p += 3) {
            int b1 = (int8_t)bytes[p], b2 = (int8_t)bytes[p + 1], b3 = (int8_t)bytes[p + 2];
            int first = ((b1 << 4) & 0xFF0) | ((b2 >> 4) & 0xF);
            out[n++] = scale(first);
            int second = ((b2 << 8) & 0xF00) | (b3 & 0xFF);
            out[n++] = scale(second);
        }
    } else { /* convertUnpacked */
        for (int p = 0; p + 2 <= n_bytes; p += 2) {
            int lsb = (int8_t)bytes[p], msb = (int8_t)bytes[p + 1];
            out[n++] = scale((lsb & 0xFF) | (msb << 8));
        }
    }
    /* mDCFilter.filter(float[]) */
    for (int x = 0; x < n; x++) {
        float filtered = out[x] - a->average;
        a->average += a->ratio * filtered;
        out[x] = filtered;
    }
    /* mHilbertTransform.filter(float[]) -- in place, pairs of real samples become (I, Q) */
    for (int y = 0; y + 1 < n; y += 2) {
        insert(a, out[y]);
        insert(a, out[y + 1]);
        float accumulator = 0.0f;
        int index = a->buffer_pointer / 2;
        for (int x = 0; x < HB_LEN / 2; x += 2)
            accumulator += a->hilbert_filter[x] * (a->buffer[a->index_map[index][x + 1]] - a->buffer[a->index_map[index][x]]);
        if (a->invert_flag) {
            out[y] = -(a->buffer[a->index_map[index][a->center_tap_index]]);
            out[y + 1] = -accumulator;
        } else {
            out[y] = a->buffer[a->index_map[index][a->center_tap_index]];
            out[y + 1] = accumulator;
        }
        a->invert_flag = !a->invert_flag;
    }
    return n;
}
