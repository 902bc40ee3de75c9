// Airspy native-buffer conversion on the device: AirspySampleConverter.convert
// (J/source/tuner/airspy/AirspySampleConverter.java:70-158) = unpack 12-bit real samples -> DCRemovalFilter(0.01f)
// (J/dsp/filter/dc/DCRemovalFilter.java:52-67) -> HilbertTransform.filter (J/dsp/filter/hilbert/HilbertTransform.java:
// 88-132).  The tuner delivers real samples at twice the complex rate; two of them make one complex sample, so the
// sample-value count of a buffer equals the float count of the I/Q stream it becomes.
//
// DC removal is a float recursion over the whole stream (average += 0.01f * (x - average)): sequential by
// definition, 12 cycles per sample on one thread = 160 MS/s.  It is contractive, though, and in float arithmetic two
// runs started from nearby averages become bit-identical after a few hundred samples (and stay so).  So the stream is
// cut into 512-sample segments, one thread each.  airspy_guess_kernel first computes, in exact-ish (double) arithmetic,
// what the linear recursion makes of every segment's samples; a thread combines the four sums behind its warm-up into a
// guess that is a few ulps from the true average, runs the exact recursion for one segment of warm-up and then its own
// (round 2; before: 2 048-sample segments, 4 096 samples of warm-up from a guessed 0).  A check then compares the value
// every segment arrived at on its first sample with the value its predecessor ended on.  If all agree the outputs are
// exactly the sequential ones by induction from segment 0, which starts from the true carried state.  Runs of segments
// that do not agree (~1 % on noise; whole stretches of slow noiseless inputs) are redone by airspy_walk_kernel from the
// true value in front of each run, the runs in parallel; only if the check still fails -- noiseless constant inputs
// never merge -- a single thread redoes the call's samples from the carried state (correct, slow, not seen with ADC
// noise present).
//
// The Hilbert transform is a plain FIR once the Java's circular buffer + index map are unrolled (checked against
// the literal restatement in oracle/orc_airspy.c): with n the second sample of pair k,
//     I[k] = s * f[n - 24],   Q[k] = s * sum_{x = 0, 2, .., 22} h[x] * (f[n - 47 + x] - f[n - x - 1]),   s = (-1)^k
// accumulated in x order with separately rounded products (no FMA), h[x] = 2 * -|halfband47[x]|.  Q only touches the
// first sample of each pair, I only the second, so the DC stage writes the filtered stream as two planes (even- and
// odd-indexed samples): e[k] = f[2k], o[k] = f[2k+1], and  I[k] = s * o[k - 12],
// Q[k] = s * sum_j h[2j] * (e[k - 23 + j] - e[k - j]) -- every load of the Hilbert kernel is contiguous across lanes.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "common.cuh"

using namespace sdrgpu;

namespace {

constexpr int kSegment = 512;    // samples per thread of the DC stage (more, shorter segments: more warps to hide latency)
constexpr int kWarmup = 512;     // samples a thread runs ahead of its segment from the guessed state (= one segment)
constexpr int kGuessSegments = 4; // segments behind the warm-up whose samples enter the guess ((1 - ratio)^(4 * 512) = 1e-9)
constexpr int kPlaneHistory = 24; // HilbertTransform looks 47 samples back = 24 even-indexed and 12 odd-indexed ones
constexpr float kRatio = 0.01f;  // AirspySampleConverter.java:31

// Filters.HALF_BAND_FILTER_47T (J/dsp/filter/Filters.java:1708-1722), even taps left of the centre, turned into
// Hilbert coefficients by HilbertTransform.convertHalfBandToHilbert (:223-244): 2.0f * -|c|
__constant__ float c_hilbert[12];
const float kHalfBandLeft[12] = {-0.000998606272947510f, 0.001695637278417295f, -0.003054430179754289f, 0.005055504379767936f,
                                 -0.007901319195893647f, 0.011873357051047719f, -0.017411159379930066f, 0.025304817427568772f,
                                 -0.037225225204559217f, 0.057533286997004301f, -0.102327462004259350f, 0.317034472508947400f};

struct AirspyState {
    float average;   // DCRemovalFilter.mAverage
    int inverted;    // HilbertTransform.mInvertFlag
    int mismatches;  // segments whose guessed start turned out wrong in the last call
    int pad;
};

// AirspySampleConverter.scale of sample i of the raw buffer (convertUnpacked :92-110 / convertPacked :118-149)
__device__ __forceinline__ float raw_sample(const uint8_t *__restrict__ raw, size_t i, bool packed)
{
    int v;
    if (!packed) {
        v = (int)raw[2 * i] | ((int)raw[2 * i + 1] << 8);
    } else {
        const uint8_t *p = raw + 3 * (i >> 1);
        v = (i & 1) ? ((((int)p[1] << 8) & 0xF00) | (int)p[2]) : ((((int)p[0] << 4) & 0xFF0) | (((int)p[1] >> 4) & 0xF));
    }
    return __fmul_rn((float)((v & 0xFFF) - 2048), 1.0f / 2048.0f);
}

// 32 samples starting at sample i (a multiple of 32) are 64 (unpacked) or 48 (packed) bytes: four / three 16-byte loads.
// Loading and unpacking are separate so that the next group's loads are in flight while this group runs the recursion.
struct RawGroup {
    uint4 w[4];
};

__device__ __forceinline__ RawGroup load_group(const uint8_t *__restrict__ raw, size_t i, bool packed)
{
    RawGroup g;
    if (!packed) {
        const uint4 *p = reinterpret_cast<const uint4 *>(raw + 2 * i);
#pragma unroll
        for (int q = 0; q < 4; q++) g.w[q] = __ldg(p + q);
    } else {
        const uint4 *p = reinterpret_cast<const uint4 *>(raw + 3 * (i >> 1));
#pragma unroll
        for (int q = 0; q < 3; q++) g.w[q] = __ldg(p + q);
        g.w[3] = make_uint4(0, 0, 0, 0);
    }
    return g;
}

__device__ __forceinline__ void unpack_group(const RawGroup &g, bool packed, float *x)
{
    unsigned u[16];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        u[4 * q] = g.w[q].x;
        u[4 * q + 1] = g.w[q].y;
        u[4 * q + 2] = g.w[q].z;
        u[4 * q + 3] = g.w[q].w;
    }
    if (!packed) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            x[2 * j] = __fmul_rn((float)((int)(u[j] & 0xFFF) - 2048), 1.0f / 2048.0f);
            x[2 * j + 1] = __fmul_rn((float)((int)((u[j] >> 16) & 0xFFF) - 2048), 1.0f / 2048.0f);
        }
    } else {
        // three little-endian words hold eight samples: bytes b0 .. b11, pair j = bytes 3j .. 3j+2
#pragma unroll
        for (int g3 = 0; g3 < 4; g3++) {
            const unsigned w0 = u[3 * g3], w1 = u[3 * g3 + 1], w2 = u[3 * g3 + 2];
            const unsigned b[12] = {w0 & 255, (w0 >> 8) & 255, (w0 >> 16) & 255, w0 >> 24, w1 & 255, (w1 >> 8) & 255,
                                    (w1 >> 16) & 255, w1 >> 24, w2 & 255, (w2 >> 8) & 255, (w2 >> 16) & 255, w2 >> 24};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int first = (int)((b[3 * j] << 4) | (b[3 * j + 1] >> 4));
                const int second = (int)(((b[3 * j + 1] & 0xF) << 8) | b[3 * j + 2]);
                x[8 * g3 + 2 * j] = __fmul_rn((float)(first - 2048), 1.0f / 2048.0f);
                x[8 * g3 + 2 * j + 1] = __fmul_rn((float)(second - 2048), 1.0f / 2048.0f);
            }
        }
    }
}

// DCRemovalFilter.filter(float): filtered = sample - average; average += ratio * filtered
__device__ __forceinline__ float dc_step(float &average, float x)
{
    const float filtered = __fsub_rn(x, average);
    average = __fadd_rn(average, __fmul_rn(kRatio, filtered));
    return filtered;
}

// filtered sample i of the call goes to the even plane (i even) or the odd plane, position i / 2
struct Planes {
    float *even, *odd;   // this call's first entries; kPlaneHistory older ones sit in front of each
};
__device__ __forceinline__ void store_filtered(const Planes &pl, int i, float v) { ((i & 1) ? pl.odd : pl.even)[i >> 1] = v; }

// The guess a segment's thread starts its warm-up from.  The DC filter is the linear recursion average = (1 - ratio)
// average + ratio x with a rounding per step; without the roundings the average at the start of segment k is
//     (1 - ratio)^512 average(k - 1) + A[k - 1],   A[k] = sum_j ratio (1 - ratio)^(511 - j) x[512 k + j]
// and (1 - ratio)^512 = 0.0058, so four segments back nothing is left at float precision.  airspy_guess_kernel computes
// A[k] (one warp per segment: coalesced loads, independent multiply-adds, a shuffle reduction -- it is a guess, its
// rounding does not matter), and a thread that used to warm up for 4 096 samples from a guessed 0.0f now starts within
// ~1e-6 of the true average and runs the exact recursion for one segment before its own: 9 x less sequential work per
// thread.  The result is still only accepted if every segment's start equals its predecessor's end bit for bit.
// eight samples from three little-endian words of the packed format (bytes b0 .. b11, pair j = bytes 3j .. 3j+2)
__device__ __forceinline__ void unpack8_packed(unsigned w0, unsigned w1, unsigned w2, float *x)
{
    const unsigned b[12] = {w0 & 255, (w0 >> 8) & 255, (w0 >> 16) & 255, w0 >> 24, w1 & 255, (w1 >> 8) & 255,
                            (w1 >> 16) & 255, w1 >> 24, w2 & 255, (w2 >> 8) & 255, (w2 >> 16) & 255, w2 >> 24};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int first = (int)((b[3 * j] << 4) | (b[3 * j + 1] >> 4));
        const int second = (int)(((b[3 * j + 1] & 0xF) << 8) | b[3 * j + 2]);
        x[2 * j] = __fmul_rn((float)(first - 2048), 1.0f / 2048.0f);
        x[2 * j + 1] = __fmul_rn((float)(second - 2048), 1.0f / 2048.0f);
    }
}

// weights[j] = ratio (1 - ratio)^(511 - j), in global memory: every lane reads its own 16 (a __constant__ table would
// serialise the 32 different addresses of a warp: measured 85 us for 20 M samples against ~10 us of HBM time)
__global__ void __launch_bounds__(256) airspy_guess_kernel(const uint8_t *__restrict__ raw, int n, int packed, int aligned,
                                                            int n_segments, const float *__restrict__ weights,
                                                            float *__restrict__ seg_a)
{
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= n_segments) return;
    const bool pk = packed != 0;
    double acc = 0.0;   // (double: the guess then differs from the float recursion only by that recursion's own roundings)
    const int base = k * kSegment + 16 * lane;   // this lane's 16 samples: 24 / 32 contiguous bytes
    float wt[16];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(weights + 16 * lane) + q);
        wt[4 * q] = v.x;
        wt[4 * q + 1] = v.y;
        wt[4 * q + 2] = v.z;
        wt[4 * q + 3] = v.w;
    }
    if (aligned && base + 16 <= n) {
        float x[16];
        if (pk) {
            const uint2 *p = reinterpret_cast<const uint2 *>(raw + 3 * (size_t)(base >> 1));   // 24 bytes, 8-byte aligned
            const uint2 a0 = __ldg(p), a1 = __ldg(p + 1), a2 = __ldg(p + 2);
            unpack8_packed(a0.x, a0.y, a1.x, x);
            unpack8_packed(a1.y, a2.x, a2.y, x + 8);
        } else {
            const uint4 *p = reinterpret_cast<const uint4 *>(raw + 2 * (size_t)base);
            const uint4 a0 = __ldg(p), a1 = __ldg(p + 1);
            const unsigned u[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int j = 0; j < 8; j++) {
                x[2 * j] = __fmul_rn((float)((int)(u[j] & 0xFFF) - 2048), 1.0f / 2048.0f);
                x[2 * j + 1] = __fmul_rn((float)((int)((u[j] >> 16) & 0xFFF) - 2048), 1.0f / 2048.0f);
            }
        }
#pragma unroll
        for (int j = 0; j < 16; j++) acc = fma((double)wt[j], (double)x[j], acc);
    } else {
        for (int j = 0; j < 16; j++) {
            const int i = base + j;
            if (i < n) acc = fma((double)wt[j], (double)raw_sample(raw, (size_t)i, pk), acc);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) seg_a[k] = (float)acc;
}

// average at the start of segment j (j >= 1) from the carried state and the segment sums (see above)
__device__ __forceinline__ float dc_guess(const float *__restrict__ seg_a, int j, float state_average, float decay)
{
    float g = 0.0f, w = 1.0f;
    const int first = j > kGuessSegments ? j - kGuessSegments : 0;
    for (int i = j - 1; i >= first; i--) {
        g = fmaf(w, seg_a[i], g);
        w *= decay;
    }
    if (first == 0) g = fmaf(w, state_average, g);   // (w = decay^j: what is left of the state the call started from)
    return g;
}

// The four per-segment logs of a call: what airspy_dc_kernel saw (start0 / end0, read-only afterwards) and what holds after
// the repair walk (start / end).
struct SegmentLog {
    float *start0, *end0;   // average on the segment's first sample after the warm-up / after its last sample
    float *start, *end;     // the same after airspy_walk_kernel
};

// samples [i, end) of the call from `average` on: filtered values into the planes, returns the average behind `end`.
// i is a multiple of 32 (segment boundaries are).
__device__ __forceinline__ float run_samples(const uint8_t *__restrict__ raw, int i, int end, bool pk, int aligned, float average,
                                             const Planes &filtered)
{
    if (aligned) {
        for (; i + 32 <= end; i += 32) {
            float x[32];
            unpack_group(load_group(raw, (size_t)i, pk), pk, x);
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                float4 fe, fo;
                fe.x = dc_step(average, x[j]);
                fo.x = dc_step(average, x[j + 1]);
                fe.y = dc_step(average, x[j + 2]);
                fo.y = dc_step(average, x[j + 3]);
                fe.z = dc_step(average, x[j + 4]);
                fo.z = dc_step(average, x[j + 5]);
                fe.w = dc_step(average, x[j + 6]);
                fo.w = dc_step(average, x[j + 7]);
                *reinterpret_cast<float4 *>(filtered.even + ((i + j) >> 1)) = fe;
                *reinterpret_cast<float4 *>(filtered.odd + ((i + j) >> 1)) = fo;
            }
        }
    }
    for (; i < end; i++) store_filtered(filtered, i, dc_step(average, raw_sample(raw, (size_t)i, pk)));
    return average;
}

// One thread per segment.
__global__ void __launch_bounds__(32) airspy_dc_kernel(const uint8_t *__restrict__ raw, int n, int packed, int aligned,
                                                         const AirspyState *__restrict__ state, Planes filtered, SegmentLog log,
                                                         const float *__restrict__ seg_a, float decay)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const long long begin = (long long)k * kSegment;
    if (begin >= n) return;
    const int end = (int)min((long long)n, begin + kSegment);
    const int from = begin > kWarmup ? (int)begin - kWarmup : 0;
    float average = from == 0 ? state->average : dc_guess(seg_a, from / kSegment, state->average, decay);
    const bool pk = packed != 0;
    // kSegment and kWarmup are multiples of 32, so whole groups of 32 samples cover everything but the call's tail
    int i = from;
    if (aligned && i + 32 <= end) {
        // groups [from, groups_end) in steps of 32; those before `begin` only warm the average up
        const int groups_end = from + (end - from) / 32 * 32;
        RawGroup next = load_group(raw, (size_t)i, pk);
        for (; i < groups_end; i += 32) {
            const RawGroup cur = next;
            if (i + 32 < groups_end) next = load_group(raw, (size_t)i + 32, pk);
            // the lines two groups further on, into L1 (a group is 48 / 64 bytes)
            if (i + 96 < groups_end) asm volatile("prefetch.global.L1 [%0];" ::"l"(raw + (pk ? 3 * (size_t)((i + 96) >> 1) : 2 * (size_t)(i + 96))));
            float x[32];
            unpack_group(cur, pk, x);
            if (i == (int)begin) log.start0[k] = log.start[k] = average;
            if (i < (int)begin) {
#pragma unroll
                for (int j = 0; j < 32; j++) dc_step(average, x[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {   // eight samples = four of each plane (i is a multiple of 32)
                    float4 fe, fo;
                    fe.x = dc_step(average, x[j]);
                    fo.x = dc_step(average, x[j + 1]);
                    fe.y = dc_step(average, x[j + 2]);
                    fo.y = dc_step(average, x[j + 3]);
                    fe.z = dc_step(average, x[j + 4]);
                    fo.z = dc_step(average, x[j + 5]);
                    fe.w = dc_step(average, x[j + 6]);
                    fo.w = dc_step(average, x[j + 7]);
                    *reinterpret_cast<float4 *>(filtered.even + ((i + j) >> 1)) = fe;
                    *reinterpret_cast<float4 *>(filtered.odd + ((i + j) >> 1)) = fo;
                }
            }
        }
    }
    // unaligned buffers and the call's tail: sample by sample
    for (; i < end; i++) {
        if (i == (int)begin) log.start0[k] = log.start[k] = average;
        const float v = dc_step(average, raw_sample(raw, (size_t)i, pk));
        if (i >= (int)begin) store_filtered(filtered, i, v);
    }
    log.end0[k] = log.end[k] = average;
}

// Repair walk.  Segment j "disagrees" when the value its thread reached on its first sample differs from what its
// predecessor ended on (the warm-up had not merged with the true trajectory yet: with a guess a few ulps off and one
// segment of warm-up that is ~1 % of the segments on noise, whole runs of them on slow noiseless inputs).  The first
// segment of a run of disagreeing ones is a head: its predecessor's end is the true average there (by induction from
// segment 0, provided no walk further up overruns into it -- the final check decides), so its thread walks forward from
// that value: a segment whose logged start equals the walker's value was right after all and is skipped in one step
// (its logged end is the true one); any other is redone.  The walk ends where the walker's value meets a logged start
// and the segment behind agrees with that one's end, or at the next head.  Everything is decided on the read-only logs
// of airspy_dc_kernel, so walkers do not race; a constant input (nothing ever merges) makes the first head walk the
// whole call, i.e. the sequential algorithm.
__global__ void __launch_bounds__(32) airspy_walk_kernel(const uint8_t *__restrict__ raw, int n, int packed, int aligned,
                                                           int n_segments, Planes filtered, SegmentLog log,
                                                           int *__restrict__ repaired)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (k >= n_segments) return;
    auto disagrees = [&](int j) {
        // (segments up to the warm-up length started from the true state at sample 0)
        return j < n_segments && (long long)j * kSegment > kWarmup &&
               __float_as_uint(log.start0[j]) != __float_as_uint(log.end0[j - 1]);
    };
    if (!disagrees(k) || disagrees(k - 1)) return;   // not a head
    float v = log.end0[k - 1];
    int redone = 0;
    for (int j = k; j < n_segments; j++) {
        if (j > k && disagrees(j) && !disagrees(j - 1)) break;   // the next head walks from here
        if (__float_as_uint(v) == __float_as_uint(log.start0[j])) {
            if (!disagrees(j + 1)) break;
            v = log.end0[j];
            continue;
        }
        log.start[j] = v;
        v = run_samples(raw, j * kSegment, min(n, (j + 1) * kSegment), packed != 0, aligned, v, filtered);
        log.end[j] = v;
        redone++;
    }
    atomicAdd(repaired, redone);
}

// every segment's start against its predecessor's end (bit patterns)
__global__ void airspy_check_kernel(int n_segments, const float *__restrict__ seg_start, const float *__restrict__ seg_end,
                                    AirspyState *state)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (k >= n_segments) return;
    if ((long long)k * kSegment <= kWarmup) return;   // started from the true state at sample 0
    if (__float_as_uint(seg_start[k]) != __float_as_uint(seg_end[k - 1])) atomicAdd(&state->mismatches, 1);
}

// commits the new average; after a mismatch (not seen in practice) redoes the stream sequentially from the carried state
__global__ void airspy_commit_kernel(const uint8_t *__restrict__ raw, int n, int packed, int n_segments,
                                     const float *__restrict__ seg_end, AirspyState *state, Planes filtered,
                                     int *__restrict__ total_mismatches)
{
    if (state->mismatches == 0) {
        if (n_segments > 0) state->average = seg_end[n_segments - 1];
        return;
    }
    float average = state->average;
    for (int i = 0; i < n; i++) store_filtered(filtered, i, dc_step(average, raw_sample(raw, (size_t)i, packed != 0)));
    state->average = average;
    atomicAdd(total_mismatches, state->mismatches);
    state->mismatches = 0;
}

// Four complex outputs per thread, k0 = 4 t: e[k0 - 24 .. k0 + 3] in seven 16-byte loads, o[k0 - 12 .. k0 - 9] in one;
// consecutive threads read consecutive 16 bytes.  Both planes are 16-byte aligned at entry 0.
__global__ void airspy_hilbert_kernel(Planes f, int n_pairs, const AirspyState *__restrict__ state, float2 *__restrict__ out,
                                      int out_aligned)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int k0 = 4 * t;
    if (k0 >= n_pairs) return;
    float w[28];   // w[j] = e[k0 - 24 + j]
    const float4 *src = reinterpret_cast<const float4 *>(f.even + k0 - 24);
#pragma unroll
    for (int q = 0; q < 7; q++) {
        const float4 v = __ldg(src + q);   // the last load may run up to 3 entries past the call: inside the padding
        w[4 * q] = v.x;
        w[4 * q + 1] = v.y;
        w[4 * q + 2] = v.z;
        w[4 * q + 3] = v.w;
    }
    const float4 centre4 = __ldg(reinterpret_cast<const float4 *>(f.odd + k0 - 12));
    const float centre[4] = {centre4.x, centre4.y, centre4.z, centre4.w};
    const int inverted = state->inverted;
    const int have = min(4, n_pairs - k0);
    float2 r[4];
#pragma unroll
    for (int p = 0; p < 4; p++) {
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < 12; j++) acc = __fadd_rn(acc, __fmul_rn(c_hilbert[j], __fsub_rn(w[p + 1 + j], w[p + 24 - j])));
        const bool invert = ((inverted + k0 + p) & 1) != 0;
        r[p] = invert ? make_float2(-centre[p], -acc) : make_float2(centre[p], acc);
    }
    if (have == 4 && out_aligned) {
        reinterpret_cast<float4 *>(out + k0)[0] = make_float4(r[0].x, r[0].y, r[1].x, r[1].y);
        reinterpret_cast<float4 *>(out + k0)[1] = make_float4(r[2].x, r[2].y, r[3].x, r[3].y);
    } else {
#pragma unroll
        for (int p = 0; p < 4; p++)
            if (p < have) out[k0 + p] = r[p];
    }
}

// the last kPlaneHistory entries of [history | this call] of each plane become the next call's history; flips the
// invert flag.  plane pointers here are the allocation starts (history first).
__global__ void airspy_carry_kernel(float *__restrict__ even_with_history, float *__restrict__ odd_with_history, int n_pairs,
                                    AirspyState *state)
{
    __shared__ float keep[2][kPlaneHistory];
    const int t = threadIdx.x;
    if (t < kPlaneHistory) {
        keep[0][t] = even_with_history[n_pairs + t];
        keep[1][t] = odd_with_history[n_pairs + t];
    }
    __syncthreads();
    if (t < kPlaneHistory) {
        even_with_history[t] = keep[0][t];
        odd_with_history[t] = keep[1][t];
    }
    if (t == 0) state->inverted = (state->inverted + n_pairs) & 1;
}

}  // namespace

// ------------------------------------------------------------------------------------------------- converter object
struct sdrgpu_airspy {
    int device = 0;
    int max_samples = 0;
    int packed = 0;
    cudaStream_t stream = nullptr;   // own stream of the stand-alone converter
    AirspyState *d_state = nullptr;
    float *d_even = nullptr, *d_odd = nullptr;   // filtered planes: [kPlaneHistory older | max_samples / 2 | 8 pad] each
    float *d_seg_log = nullptr;      // SegmentLog: four rows of one entry per segment
    float *d_weights = nullptr;      // airspy_guess_kernel: ratio (1 - ratio)^(kSegment - 1 - j)
    int segments_cap = 0;
    bool walk = true;                // SDRGPU_AIRSPY_WALK=0: no repair walk (tests of the sequential fallback)
    float *d_seg_a = nullptr;        // per segment: its samples' contribution to the average at its end (airspy_guess_kernel)
    int *d_total_mismatches = nullptr, *d_repaired = nullptr;
    uint8_t *d_raw = nullptr;        // staging for host input
    float2 *d_iq = nullptr;          // staging for host output
};

namespace sdrgpu {

sdrgpu_status airspy_create(sdrgpu_airspy **out, int max_samples)
{
    static bool coefficients_loaded[64] = {false};
    auto *a = new sdrgpu_airspy();
    cudaGetDevice(&a->device);
    a->max_samples = max_samples;
    const int segments = (max_samples + kSegment - 1) / kSegment + 1;
    auto bail = [&](sdrgpu_status s) {
        airspy_destroy(a);
        return s;
    };
#define CHK(call)                                                                                                   \
    do {                                                                                                            \
        cudaError_t e__ = (call);                                                                                   \
        if (e__ != cudaSuccess) return bail(fail(SDRGPU_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__))); \
    } while (0)
    if (a->device < 64 && !coefficients_loaded[a->device]) {
        float h[12];
        for (int j = 0; j < 12; j++) h[j] = 2.0f * -fabsf(kHalfBandLeft[j]);
        CHK(cudaMemcpyToSymbol(c_hilbert, h, sizeof(h)));
        coefficients_loaded[a->device] = true;
    }
    {
        std::vector<float> w(kSegment);
        for (int j = 0; j < kSegment; j++) w[j] = (float)((double)kRatio * pow(1.0 - (double)kRatio, (double)(kSegment - 1 - j)));
        CHK(cudaMalloc(&a->d_weights, sizeof(float) * kSegment));
        CHK(cudaMemcpy(a->d_weights, w.data(), sizeof(float) * kSegment, cudaMemcpyHostToDevice));
    }
    a->segments_cap = segments;
    if (const char *e = getenv("SDRGPU_AIRSPY_WALK")) a->walk = atoi(e) != 0;
    CHK(cudaMalloc(&a->d_state, sizeof(AirspyState)));
    CHK(cudaMemset(a->d_state, 0, sizeof(AirspyState)));
    // + 8 behind: the Hilbert kernel's last 16-byte load
    const size_t plane = sizeof(float) * ((size_t)max_samples / 2 + kPlaneHistory + 8);
    CHK(cudaMalloc(&a->d_even, plane));
    CHK(cudaMemset(a->d_even, 0, plane));
    CHK(cudaMalloc(&a->d_odd, plane));
    CHK(cudaMemset(a->d_odd, 0, plane));
    CHK(cudaMalloc(&a->d_seg_a, sizeof(float) * (size_t)segments));
    CHK(cudaMalloc(&a->d_seg_log, sizeof(float) * 4 * (size_t)segments));
    CHK(cudaMalloc(&a->d_total_mismatches, sizeof(int)));
    CHK(cudaMemset(a->d_total_mismatches, 0, sizeof(int)));
    CHK(cudaMalloc(&a->d_repaired, sizeof(int)));
    CHK(cudaMemset(a->d_repaired, 0, sizeof(int)));
#undef CHK
    *out = a;
    return SDRGPU_OK;
}

void airspy_destroy(sdrgpu_airspy *a)
{
    if (!a) return;
    cudaFree(a->d_state);
    cudaFree(a->d_even);
    cudaFree(a->d_odd);
    cudaFree(a->d_seg_a);
    cudaFree(a->d_seg_log);
    cudaFree(a->d_weights);
    cudaFree(a->d_total_mismatches);
    cudaFree(a->d_repaired);
    cudaFree(a->d_raw);
    cudaFree(a->d_iq);
    if (a->stream) cudaStreamDestroy(a->stream);
    delete a;
}

size_t airspy_raw_bytes(int n_samples, int packed) { return packed ? (size_t)(n_samples / 2) * 3 : (size_t)n_samples * 2; }

// n_samples (even) raw real samples at d_raw -> n_samples / 2 complex samples at d_out, on `stream`
sdrgpu_status airspy_enqueue(sdrgpu_airspy *a, const uint8_t *d_raw, int n_samples, int packed, float2 *d_out, cudaStream_t stream)
{
    if (n_samples == 0) return SDRGPU_OK;
    if (n_samples > a->max_samples) return fail(SDRGPU_ERR_OVERFLOW, "%d samples exceed the converter's capacity %d", n_samples, a->max_samples);
    const int segments = (n_samples + kSegment - 1) / kSegment;
    const Planes f{a->d_even + kPlaneHistory, a->d_odd + kPlaneHistory};   // 96 bytes in: 16-byte aligned
    const int aligned = ((uintptr_t)d_raw & 15) == 0;
    const float decay = (float)pow(1.0 - (double)kRatio, (double)kSegment);
    const size_t cap = (size_t)a->segments_cap;
    const SegmentLog log{a->d_seg_log, a->d_seg_log + cap, a->d_seg_log + 2 * cap, a->d_seg_log + 3 * cap};
    if (segments > 1)
        airspy_guess_kernel<<<(segments + 7) / 8, 256, 0, stream>>>(d_raw, n_samples, packed, aligned, segments, a->d_weights, a->d_seg_a);
    airspy_dc_kernel<<<(segments + 31) / 32, 32, 0, stream>>>(d_raw, n_samples, packed, aligned, a->d_state, f, log, a->d_seg_a, decay);
    if (segments > 1) {
        if (a->walk)
            airspy_walk_kernel<<<(segments + 31) / 32, 32, 0, stream>>>(d_raw, n_samples, packed, aligned, segments, f, log, a->d_repaired);
        airspy_check_kernel<<<(segments + 255) / 256, 256, 0, stream>>>(segments, log.start, log.end, a->d_state);
    }
    airspy_commit_kernel<<<1, 1, 0, stream>>>(d_raw, n_samples, packed, segments, log.end, a->d_state, f, a->d_total_mismatches);
    const int pairs = n_samples / 2;
    airspy_hilbert_kernel<<<((pairs + 3) / 4 + 127) / 128, 128, 0, stream>>>(f, pairs, a->d_state, d_out,
                                                                             ((uintptr_t)d_out & 15) == 0);
    airspy_carry_kernel<<<1, 64, 0, stream>>>(a->d_even, a->d_odd, pairs, a->d_state);
    count_launch(segments > 1 ? (a->walk ? 7 : 6) : 4);
    SDRGPU_CUDA(cudaGetLastError());
    return SDRGPU_OK;
}

sdrgpu_status airspy_reset(sdrgpu_airspy *a, cudaStream_t stream)
{
    SDRGPU_CUDA(cudaMemsetAsync(a->d_state, 0, sizeof(AirspyState), stream));
    SDRGPU_CUDA(cudaMemsetAsync(a->d_even, 0, sizeof(float) * kPlaneHistory, stream));
    SDRGPU_CUDA(cudaMemsetAsync(a->d_odd, 0, sizeof(float) * kPlaneHistory, stream));
    return SDRGPU_OK;
}

}  // namespace sdrgpu

extern "C" {

sdrgpu_status sdrgpu_airspy_create(sdrgpu_airspy **out, int max_samples)
{
    if (!out || max_samples <= 0 || (max_samples & 1)) return fail(SDRGPU_ERR_INVALID_ARG, "max_samples must be positive and even");
    *out = nullptr;
    sdrgpu_airspy *a = nullptr;
    SDRGPU_TRY(sdrgpu::airspy_create(&a, max_samples));
    if (cudaStreamCreateWithFlags(&a->stream, cudaStreamNonBlocking) != cudaSuccess) {
        sdrgpu::airspy_destroy(a);
        return fail(SDRGPU_ERR_CUDA, "cannot create a stream");
    }
    *out = a;
    return SDRGPU_OK;
}

sdrgpu_status sdrgpu_airspy_destroy(sdrgpu_airspy *a)
{
    if (a) {
        cudaSetDevice(a->device);
        if (a->stream) cudaStreamSynchronize(a->stream);
    }
    sdrgpu::airspy_destroy(a);
    return SDRGPU_OK;
}

sdrgpu_status sdrgpu_airspy_set_sample_packing(sdrgpu_airspy *a, int enabled)
{
    if (!a) return fail(SDRGPU_ERR_INVALID_ARG, "NULL handle");
    a->packed = enabled != 0;
    return SDRGPU_OK;
}

sdrgpu_status sdrgpu_airspy_convert(sdrgpu_airspy *a, const void *raw, int raw_mem, int n_samples, float *iq, int iq_mem)
{
    if (!a) return fail(SDRGPU_ERR_INVALID_ARG, "NULL handle");
    if (n_samples < 0 || (n_samples & 1)) return fail(SDRGPU_ERR_INVALID_ARG, "n_samples must be even (two real samples per complex sample)");
    if (n_samples > a->max_samples) return fail(SDRGPU_ERR_OVERFLOW, "%d samples exceed max_samples %d", n_samples, a->max_samples);
    if (n_samples == 0) return SDRGPU_OK;
    if (!raw || !iq) return fail(SDRGPU_ERR_INVALID_ARG, "NULL buffer");
    SDRGPU_CUDA(cudaSetDevice(a->device));
    const uint8_t *d_raw = reinterpret_cast<const uint8_t *>(raw);
    const size_t bytes = sdrgpu::airspy_raw_bytes(n_samples, a->packed);
    if (raw_mem == SDRGPU_HOST) {
        if (!a->d_raw) SDRGPU_CUDA(cudaMalloc(&a->d_raw, sdrgpu::airspy_raw_bytes(a->max_samples, 0)));
        SDRGPU_CUDA(cudaMemcpyAsync(a->d_raw, raw, bytes, cudaMemcpyHostToDevice, a->stream));
        d_raw = a->d_raw;
    }
    float2 *d_out = reinterpret_cast<float2 *>(iq);
    if (iq_mem == SDRGPU_HOST) {
        if (!a->d_iq) SDRGPU_CUDA(cudaMalloc(&a->d_iq, sizeof(float2) * (size_t)(a->max_samples / 2)));
        d_out = a->d_iq;
    } else if ((uintptr_t)iq & 7) {
        return fail(SDRGPU_ERR_INVALID_ARG, "device output must be 8-byte aligned");
    }
    SDRGPU_TRY(sdrgpu::airspy_enqueue(a, d_raw, n_samples, a->packed, d_out, a->stream));
    if (iq_mem == SDRGPU_HOST)
        SDRGPU_CUDA(cudaMemcpyAsync(iq, d_out, sizeof(float) * (size_t)n_samples, cudaMemcpyDeviceToHost, a->stream));
    SDRGPU_CUDA(cudaStreamSynchronize(a->stream));
    return SDRGPU_OK;
}

sdrgpu_status sdrgpu_airspy_mismatches(sdrgpu_airspy *a, int *count)
{
    if (!a || !count) return fail(SDRGPU_ERR_INVALID_ARG, "NULL argument");
    SDRGPU_CUDA(cudaSetDevice(a->device));
    SDRGPU_CUDA(cudaStreamSynchronize(a->stream));
    SDRGPU_CUDA(cudaMemcpy(count, a->d_total_mismatches, sizeof(int), cudaMemcpyDeviceToHost));
    return SDRGPU_OK;
}

sdrgpu_status sdrgpu_airspy_repaired(sdrgpu_airspy *a, int *count)
{
    if (!a || !count) return fail(SDRGPU_ERR_INVALID_ARG, "NULL argument");
    SDRGPU_CUDA(cudaSetDevice(a->device));
    SDRGPU_CUDA(cudaStreamSynchronize(a->stream));
    SDRGPU_CUDA(cudaMemcpy(count, a->d_repaired, sizeof(int), cudaMemcpyDeviceToHost));
    return SDRGPU_OK;
}

}  // extern "C"
