/* CPU ORACLE -- test infrastructure only (see sdr_oracle.h).
 * Parks-McClellan / Remez exchange low-pass designer, statement for statement after
 *   J/dsp/filter/fir/FIRFilterSpecification.java:381-428 (LowPassBuilder.build), :205-296 (extrema count, grid sizes,
 *     grid interval), :909-937 (estimateFilterOrder), :991-1135 (FrequencyBand)
 *   J/dsp/filter/fir/remez/Grid.java:25-78
 *   J/dsp/filter/fir/remez/RemezFIRFilterDesigner.java:62-98 (design loop), :146-193 (impulse response),
 *     :207-232 (Lagrange evaluation), :250-262 (initial extrema), :268-397 (b, delta, C, D), :405-417 (grid error),
 *     :424-525 (extremal search), :537-563 (convergence), :586-608 (resample)
 *   J/dsp/filter/FilterFactory.java:671-681 (getTaps)
 * as the decoders call it: P25P1DecoderC4FM.java:136-148, P25P2DecoderHDQPSK.java:155-166, NBFMDecoder.java:306-341.
 * FastMath.cos / pow / log10 / ceil -> libm in double (commons-math3 documents < 1 ulp; see sdr_oracle.h). */
#include "sdr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static const double REMEZ_PI = 3.14159265358979323846;

/* FIRFilterSpecification.estimateFilterOrder (:909-937), Herrmann et al. 1973 */
int orc_remez_estimate_order(double sampleRate, double frequency1, double frequency2, double passBandRipple,
                             double stopBandRipple)
{
    double df = fabs(frequency2 - frequency1) / sampleRate;
    double ddp = log10(fmax(stopBandRipple, passBandRipple));
    double dds = log10(fmin(stopBandRipple, passBandRipple));
    double a1 = 5.309e-3, a2 = 7.114e-2, a3 = -4.761e-1, a4 = -2.66e-3, a5 = -5.941e-1, a6 = -4.278e-1;
    double b1 = 11.01217, b2 = 0.5124401;
    double t1 = a1 * ddp * ddp;
    double t2 = a2 * ddp;
    double t3 = a4 * ddp * ddp;
    double t4 = a5 * ddp;
    double dinf = ((t1 + t2 + a3) * dds) + (t3 + t4 + a6);
    double ff = b1 + b2 * (ddp - dds);
    double n = dinf / df - ff * df + 1.0;
    return (int)ceil(n);
}

typedef struct {
    int gridSize;
    double start, end, amplitude, rippleDB;
} Band;

/* FrequencyBand.getRippleAmplitude (:1103-1107) */
static double band_ripple_amplitude(const Band *b)
{
    return (pow(10.0, (b->rippleDB / 20)) - 1) / (pow(10.0, (b->rippleDB / 20)) + 1);
}

typedef struct {
    int type; /* 1 = TYPE_1 odd length, 2 = TYPE_2 even length */
    int order, gridDensity;
    Band bands[2];
    int nBands;
} Spec;

static int spec_extrema_count(const Spec *s)
{
    /* getHalfFilterOrder (:217-233) + 2 */
    return (s->type == 1 ? s->order / 2 : (s->order - 1) / 2) + 2;
}

static double spec_total_bandwidth(const Spec *s)
{
    double bandwidth = 0.0;
    for (int i = 0; i < s->nBands; i++) bandwidth += s->bands[i].end - s->bands[i].start;
    return bandwidth;
}

/* updateGridSize (:273-286) + FrequencyBand.setGridSize (:1041-1044) */
static void spec_update_grid_size(Spec *s)
{
    int gridSize = (spec_extrema_count(s) - 1) * s->gridDensity + 1;
    double totalBandwidth = spec_total_bandwidth(s);
    for (int i = 0; i < s->nBands; i++) {
        Band *b = &s->bands[i];
        int g = (int)ceil((double)gridSize * ((b->end - b->start) / totalBandwidth));
        b->gridSize = g > 1 ? g : 1;
    }
}

static int spec_grid_size(const Spec *s)
{
    int n = 0;
    for (int i = 0; i < s->nBands; i++) n += s->bands[i].gridSize;
    return n;
}

typedef struct {
    const Spec *spec;
    int gridSize;
    double *cosGrid, *desired, *weight; /* Grid */
    int *extremal;
    int nExtremal;
    double *d, *ideal, *gridResponse, *gridErrors;
    int nD, nIdeal;
    double delta;
    int converged;
} Designer;

/* RemezFIRFilterDesigner.getFrequencyResponse(double) (:207-232) */
static double response_at(const Designer *z, double cosineOfFrequency)
{
    double numerator = 0.0, denominator = 0.0;
    for (int k = 0; k < z->nExtremal - 1; k++) {
        double cosineDelta = cosineOfFrequency - z->cosGrid[z->extremal[k]];
        if (fabs(cosineDelta) < 1.0e-7) return z->ideal[k];
        double dkOverCosineDelta = z->d[k] / cosineDelta;
        numerator += dkOverCosineDelta * z->ideal[k];
        denominator += dkOverCosineDelta;
    }
    return numerator / denominator;
}

/* calculateGridFrequencyResponse (:268-279): calculateB, calculateDelta, calculateC, calculateD, update */
static void calculate_grid_frequency_response(Designer *z, double *b)
{
    const int length = z->nExtremal;
    for (int k = 0; k < length; k++) { /* calculateB (:301-330) */
        b[k] = 1.0;
        double xk = z->cosGrid[z->extremal[k]];
        for (int i = 0; i < length; i++) {
            if (i != k) {
                double xi = z->cosGrid[z->extremal[i]];
                double denominator = xk - xi;
                if (fabs(denominator) < 0.00001) denominator = 0.00001;
                b[k] *= 1.0 / denominator;
            }
        }
    }
    { /* calculateDelta (:340-364) */
        double numerator = 0.0, denominator = 0.0, sign = 1.0;
        for (int k = 0; k < length; k++) {
            int extremalIndex = z->extremal[k];
            numerator += (b[k] * z->desired[extremalIndex]);
            denominator += b[k] * sign / z->weight[extremalIndex];
            sign = -sign;
        }
        z->delta = numerator / denominator;
    }
    { /* calculateC (:374-395) */
        int n = spec_extrema_count(z->spec) - 1;
        z->nIdeal = n;
        double sign = 1.0;
        for (int k = 0; k < n; k++) {
            z->ideal[k] = 0.0;
            if (k < z->nExtremal) {
                int index = z->extremal[k];
                z->ideal[k] = z->desired[index] - (sign * z->delta / z->weight[index]);
                sign = -sign;
            }
        }
    }
    { /* calculateD (:402-415) */
        int n = z->nExtremal - 1;
        z->nD = n;
        for (int k = 0; k < n; k++) z->d[k] = b[k] * (z->cosGrid[z->extremal[k]] - z->cosGrid[z->extremal[n]]);
    }
    for (int i = 0; i < z->gridSize; i++) z->gridResponse[i] = response_at(z, z->cosGrid[i]); /* (:284-293) */
}

/* isGTEDelta (:537-540) */
static int is_gte_delta(const Designer *z, double value) { return fabs(value) - fabs(z->delta) > -1.0e-5; }

/* findExtremalIndices (:424-525); returns -1 where the Java throws FilterDesignException */
static int find_extremal_indices(Designer *z)
{
    const double *e = z->gridErrors;
    const int n = z->gridSize, want = spec_extrema_count(z->spec);
    int *list = z->extremal, size = 0;
    if (((e[0] > 0.0 && e[0] > e[1]) || (e[0] < 0.0 && e[0] < e[1])) && is_gte_delta(z, e[0])) list[size++] = 0;
    for (int x = 1; x < n - 1; x++) {
        if (((e[x] > 0.0 && (e[x - 1] <= e[x] && e[x] > e[x + 1])) || (e[x] < 0.0 && (e[x - 1] >= e[x] && e[x] < e[x + 1]))) &&
            is_gte_delta(z, e[x]))
            list[size++] = x;
    }
    int last = n - 1;
    if (((e[last] > 0.0 && (e[last] > e[last - 1])) || (e[last] < 0.0 && (e[last] < e[last - 1]))) && is_gte_delta(z, e[last]))
        list[size++] = last;
    if (size < want) return -1;

    /* alternation: one extremal per excursion; `keep` marks what survives it.remove() / removeAll(indicesToRemove) */
    char *removed = (char *)calloc((size_t)size, 1);
    int current = 0; /* position in list of `current` */
    int positiveAxis = e[list[current]] > 0.0;
    for (int j = 1; j < size; j++) {
        int next = j;
        if (!(positiveAxis ^ (e[list[next]] > 0.0))) {
            if (fabs(e[list[next]]) <= fabs(e[list[current]])) {
                removed[next] = 1; /* it.remove() */
                next = current;
            } else {
                removed[current] = 1; /* indicesToRemove.add(current) */
            }
        } else {
            positiveAxis = !positiveAxis;
        }
        current = next;
    }
    int kept = 0;
    for (int j = 0; j < size; j++)
        if (!removed[j]) list[kept++] = list[j];
    free(removed);
    size = kept;
    while (size > want) size--; /* remove excess trailing indices */
    if (size > want) {          /* (unreachable after the loop above, kept for the Java's shape) */
        if (fabs(e[list[0]]) > fabs(e[list[size - 1]])) size--;
        else {
            memmove(list, list + 1, sizeof(int) * (size_t)(size - 1));
            size--;
        }
    }
    z->nExtremal = size;
    if (size < want) return -1;
    return 0;
}

/* checkConvergence (:548-563) */
static void check_convergence(Designer *z)
{
    double maximum = fabs(z->gridErrors[z->extremal[0]]);
    for (int i = 1; i < z->nExtremal; i++) {
        double current = fabs(z->gridErrors[z->extremal[i]]);
        if (current > maximum) maximum = current;
    }
    double convergence = maximum - fabs(z->delta);
    z->converged = convergence < 0.0001;
}

/* LowPassBuilder.build + Grid + RemezFIRFilterDesigner + getImpulseResponse.  order < 6: estimated; odd_length: -1
 * unset, 0 / 1 as oddLength(false / true).  Returns the number of taps, -1 when the design does not converge
 * (FilterFactory.getTaps returns null), -2 when `capacity` is too small. */
int orc_remez_low_pass(double sampleRate, double passBandEnd, double stopBandStart, double passBandRipple,
                       double stopBandRipple, int order, int odd_length, int gridDensity, float *out, int capacity)
{
    Spec spec;
    memset(&spec, 0, sizeof(spec));
    if (order < 6) order = orc_remez_estimate_order(sampleRate, passBandEnd, stopBandStart, passBandRipple, stopBandRipple);
    if (odd_length >= 0) {
        if (odd_length) {
            spec.type = 1;
            order += order % 2;
        } else {
            spec.type = 2;
            order += (order % 2 == 0 ? 1 : 0);
        }
    } else {
        spec.type = (order % 2 == 0) ? 1 : 2;
    }
    spec.order = order;
    spec.gridDensity = gridDensity;
    /* FrequencyBand(sampleRate, start, end, amplitude, ripple): edges normalised to the sample rate */
    spec.bands[0] = (Band){0, 0 / sampleRate, passBandEnd / sampleRate, 1.0, passBandRipple};
    spec.nBands = 1;
    spec_update_grid_size(&spec);
    spec.bands[1] = (Band){0, stopBandStart / sampleRate, (double)(int)(sampleRate / 2) / sampleRate, 0.0, stopBandRipple};
    spec.nBands = 2;
    spec_update_grid_size(&spec);
    const int length = spec.order + 1;
    if (length > capacity) return -2;

    Designer z;
    memset(&z, 0, sizeof(z));
    z.spec = &spec;
    z.gridSize = spec_grid_size(&spec);
    const int count = spec_extrema_count(&spec);
    z.cosGrid = (double *)calloc((size_t)z.gridSize, sizeof(double));
    z.desired = (double *)calloc((size_t)z.gridSize, sizeof(double));
    z.weight = (double *)calloc((size_t)z.gridSize, sizeof(double));
    z.gridResponse = (double *)calloc((size_t)z.gridSize, sizeof(double));
    z.gridErrors = (double *)calloc((size_t)z.gridSize, sizeof(double));
    z.extremal = (int *)calloc((size_t)z.gridSize + 2, sizeof(int));
    z.d = (double *)calloc((size_t)z.gridSize + 2, sizeof(double));
    z.ideal = (double *)calloc((size_t)z.gridSize + 2, sizeof(double));
    double *b = (double *)calloc((size_t)z.gridSize + 2, sizeof(double));

    { /* Grid.create (:25-78) */
        double gridFrequencyInterval = spec_total_bandwidth(&spec) / (double)(z.gridSize - spec.nBands);
        double grid0 = spec.bands[0].start; /* symmetric types */
        double maxRipple = 0.0;
        for (int x = 0; x < spec.nBands; x++) {
            double r = band_ripple_amplitude(&spec.bands[x]);
            if (r > maxRipple) maxRipple = r;
        }
        int j = 0;
        for (int x = 0; x < spec.nBands; x++) {
            const Band *band = &spec.bands[x];
            double lowFrequency = (x == 0 ? grid0 : band->start);
            for (int i = 0; i < band->gridSize; i++) {
                z.desired[j] = band->amplitude;
                z.weight[j] = 1.0 / (band_ripple_amplitude(band) / maxRipple);
                /* (the Java then overwrites the LAST frequency of the band with the band edge, but not its cosine) */
                z.cosGrid[j] = cos(2.0 * REMEZ_PI * lowFrequency);
                lowFrequency += gridFrequencyInterval;
                j++;
            }
        }
    }
    /* getInitialExtremalIndices (:250-262) */
    z.nExtremal = count;
    for (int i = 0; i < count; i++) z.extremal[i] = i * spec.gridDensity;

    int iterationCount = 0;
    do { /* design (:62-98) */
        calculate_grid_frequency_response(&z, b);
        for (int i = 0; i < z.gridSize; i++) z.gridErrors[i] = z.weight[i] * (z.desired[i] - z.gridResponse[i]);
        if (find_extremal_indices(&z) != 0) {
            z.converged = 0;
            break;
        }
        check_convergence(&z);
        iterationCount++;
    } while (!z.converged && iterationCount < 40);

    int result = -1;
    if (z.converged) {
        calculate_grid_frequency_response(&z, b);
        /* resample (:586-608) */
        int rl = length;
        if (rl % 2 == 0) rl--;
        double half = (double)rl / 2.0;
        int nfr = (int)ceil(half);
        double *fr = (double *)calloc((size_t)nfr, sizeof(double));
        for (int x = 0; x < nfr; x++) fr[x] = response_at(&z, cos(REMEZ_PI * (double)x / half));
        /* getImpulseResponseDoubles (:146-193), then (float) */
        const double TWO_PI = 2.0 * REMEZ_PI;
        if (spec.type == 1) {
            double M = ((double)length - 1.0) / 2.0;
            for (int n = 0; n < length; n++) {
                double accumulator = fr[0];
                double frequency = TWO_PI * (n - M) / length;
                for (int k = 1; k <= M; k++) accumulator += 2.0 * fr[k] * cos(frequency * (double)k);
                out[n] = (float)(accumulator / (double)length);
            }
        } else {
            double offset = (double)(length - 1) / 2.0;
            for (int n = 0; n < length; n++) {
                double accumulator = fr[0];
                double frequency = TWO_PI * ((double)n - offset) / (double)length;
                for (int k = 1; k < nfr; k++) accumulator += 2.0 * fr[k] * cos(frequency * (double)k);
                out[n] = (float)(accumulator / (double)length);
            }
        }
        free(fr);
        result = length;
    }
    free(z.cosGrid);
    free(z.desired);
    free(z.weight);
    free(z.gridResponse);
    free(z.gridErrors);
    free(z.extremal);
    free(z.d);
    free(z.ideal);
    free(b);
    return result;
}
