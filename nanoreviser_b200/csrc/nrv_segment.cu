// K1: re-segmentation of the raw signal on basecalled event boundaries.
//   read_stats      : per-read shift = median(signal), scale = median(|signal - shift|)
//                     (preprocessing.py:95-97) by exact radix selection in the int16 domain,
//                     plus the per-read status checks of the boundary.
//   base_features   : per-base raw mean / std (preprocessing.py:134-137) and the six feature
//                     columns (nanorevtrainutils.py:162-169, NanoReviser.py:124-125).
//   sig_windows     : the normalised, symmetrically zero-padded 50-sample windows
//                     (preprocessing.py:111-131).  Only materialised for nrv_segment(); the
//                     model path gathers windows inside the CNN kernel instead.
// All arithmetic that the reference does in float64 is done in float64 here and rounded to
// fp32 once, exactly where Keras casts its inputs.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "nrv_common.cuh"

namespace nrv {

// ------------------------------------------------------------------------------------------
// exact median / MAD from ONE pass over the signal: a full-resolution histogram of the int16 values
// ------------------------------------------------------------------------------------------
// A read is cut into segments of RS_SEG samples, one CTA each (a 2.7 M-sample read of cfg5 is 83 CTAs, a 100 k-sample read of
// cfg2 four).  Every CTA counts its samples into a 65,536-bin histogram in shared memory (16-bit counters, two per word: a
// segment has fewer than 65,536 samples, so a counter cannot overflow into its neighbour) with 128-bit loads.  A read that
// fits one segment is finished by its CTA; otherwise the CTAs add their non-zero bins into the read's 32-bit histogram in global
// memory and the LAST CTA to arrive (a counter per read) finishes it and leaves histogram and counter zeroed for the next batch.
// The median is a rank search in the histogram.  The MAD needs no second pass over the data either: the number of samples with
// |2 key - shift2| <= d is C(hi(d)) - C(lo(d) - 1) for the cumulative counts C, so the k-th smallest deviation is a search over d
// (1024-ary: two rounds).  Results are exact: shift is a multiple of 0.5, scale a multiple of 0.25 (np.median of an even count
// averages the two middle order statistics).
constexpr int RS_SEG = 32768;
constexpr int RS_PENDING = -1;        // status of a read the compact kernel hands to the full-range kernel

// NBINS histogram bins searched by THREADS threads: chunks of 16 bins, CPT chunks per thread; the deviation search runs over
// d in [0, 2 NBINS) in two rounds of DSTEP
template <int NBINS_, int THREADS_>
struct RsCfg {
    static constexpr int NBINS = NBINS_, THREADS = THREADS_, CHUNKS = NBINS_ / 16, CPT = CHUNKS / THREADS_, DSTEP = 2 * NBINS_ / THREADS_;
};
template <class Cfg, int HWORDS>
struct RsSmemT {
    uint32_t h[HWORDS];                   // the CTA's histogram
    uint32_t p16[Cfg::CHUNKS + 1];        // exclusive cumulative count before every chunk of 16 bins; [CHUNKS] = n
    uint32_t fr[Cfg::THREADS];
    uint32_t fr2[2][Cfg::DSTEP];
    uint32_t warp_tot[Cfg::THREADS / 32];
    int res[4];
    int bad, is_last;
};
using RsFull = RsCfg<65536, 1024>;        // every int16 value its own bin; 16-bit counters, two per word (bin k in the (k & 1) half)
using RsFullSmem = RsSmemT<RsFull, 65536 / 2>;
using RsCompact = RsCfg<16384, 512>;      // values -8191 .. 8190 their own bin, the rest clamped into bins 0 / 16383; 32-bit counters
using RsCompactSmem = RsSmemT<RsCompact, 16384>;
constexpr int RC_OFF = 8192;

// the 16 bins of chunk c in registers: 2 x LDS.128 (16-bit counters) resp. 4 x 128-bit loads (no dependent scalar loads)
struct RsSrcShared {
    const uint32_t* h;
    __device__ __forceinline__ void chunk(int c, uint32_t (&b)[16]) const {
        const uint4 a0 = *reinterpret_cast<const uint4*>(h + c * 8), a1 = *reinterpret_cast<const uint4*>(h + c * 8 + 4);
        const uint32_t w[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) { b[2 * i] = w[i] & 0xffffu; b[2 * i + 1] = w[i] >> 16; }
    }
};
struct RsSrcShared32 {
    const uint32_t* h;
    __device__ __forceinline__ void chunk(int c, uint32_t (&b)[16]) const {
        const uint4* p = reinterpret_cast<const uint4*>(h + c * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint4 v = p[i];
            b[4 * i] = v.x; b[4 * i + 1] = v.y; b[4 * i + 2] = v.z; b[4 * i + 3] = v.w;
        }
    }
};
struct RsSrcGlobal {
    const uint32_t* g;
    __device__ __forceinline__ void chunk(int c, uint32_t (&b)[16]) const {
        const uint4* p = reinterpret_cast<const uint4*>(g + c * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint4 v = __ldcg(p + i);
            b[4 * i] = v.x; b[4 * i + 1] = v.y; b[4 * i + 2] = v.z; b[4 * i + 3] = v.w;
        }
    }
};

template <class Src>
__device__ __forceinline__ uint32_t rs_chunk_sum(const Src& src, int chunk, int upto) {     // bins chunk*16 .. chunk*16 + upto
    uint32_t b[16];
    src.chunk(chunk, b);
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += (i <= upto) ? b[i] : 0u;
    return s;
}
// cumulative count of keys <= x
template <class Cfg, class Src, class Smem>
__device__ __forceinline__ uint32_t rs_cum(const Src& src, const Smem& sm, int x, uint32_t n) {
    if (x < 0) return 0;
    if (x >= Cfg::NBINS) return n;
    return sm.p16[x >> 4] + rs_chunk_sum(src, x >> 4, x & 15);
}
// number of samples with |2 key - shift2| <= d
template <class Cfg, class Src, class Smem>
__device__ __forceinline__ uint32_t rs_within(const Src& src, const Smem& sm, int shift2, int d, uint32_t n) {
    const int lo = (shift2 - d + 1) >> 1, hi = (shift2 + d) >> 1;      // ceil / floor (arithmetic shifts)
    return rs_cum<Cfg>(src, sm, hi, n) - rs_cum<Cfg>(src, sm, lo - 1, n);
}

// whole CTA: sel = {keys of the two middle order statistics, the two middle |2 key - (key1 + key2)|}
template <class Cfg, class Src, class Smem>
__device__ void rs_select(const Src& src, Smem& sm, uint32_t n, int (&sel)[4]) {
    constexpr int CPT = Cfg::CPT, DSTEP = Cfg::DSTEP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t k1 = (n - 1) / 2, k2 = n / 2;
    // cumulative counts per chunk of 16 bins: thread t owns chunks CPT t .. CPT t + CPT - 1
    uint32_t cs[CPT], local = 0;
#pragma unroll
    for (int j = 0; j < CPT; ++j) { cs[j] = rs_chunk_sum(src, CPT * tid + j, 15); local += cs[j]; }
    uint32_t v = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    if (lane == 31) sm.warp_tot[warp] = v;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += sm.warp_tot[w];
    uint32_t run = base + v - local;
#pragma unroll
    for (int j = 0; j < CPT; ++j) { sm.p16[CPT * tid + j] = run; run += cs[j]; }
    if (tid == Cfg::THREADS - 1) sm.p16[Cfg::CHUNKS] = run;
    __syncthreads();
    // the two middle order statistics
    run = base + v - local;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const uint32_t k = which ? k2 : k1;
            if (k >= run && k < run + cs[j]) {
                uint32_t b[16];
                src.chunk(CPT * tid + j, b);
                uint32_t r2 = run;
                int found = 15;
#pragma unroll
                for (int i = 15; i >= 0; --i) {          // first bin whose cumulative count exceeds k
                    uint32_t upto = 0;
#pragma unroll
                    for (int q = 0; q <= i; ++q) upto += b[q];
                    if (k < r2 + upto) found = i;
                }
                sm.res[which] = (CPT * tid + j) * 16 + found;
            }
        }
        run += cs[j];
    }
    __syncthreads();
    const int m1 = sm.res[0], m2 = sm.res[1];
    const int shift2 = m1 + m2;
    // k-th smallest deviation: smallest d with within(d) >= k + 1; d in [0, 2 NBINS), round 1 at d = DSTEP t + DSTEP - 1
    sm.fr[tid] = rs_within<Cfg>(src, sm, shift2, DSTEP * tid + DSTEP - 1, n);
    __syncthreads();
#pragma unroll
    for (int which = 0; which < 2; ++which) {
        const uint32_t need = (which ? k2 : k1) + 1;
        if (sm.fr[tid] >= need && (tid == 0 || sm.fr[tid - 1] < need)) sm.res[2 + which] = tid;
    }
    __syncthreads();
    if (tid < 2 * DSTEP) {
        const int which = tid / DSTEP, i = tid % DSTEP;
        sm.fr2[which][i] = rs_within<Cfg>(src, sm, shift2, DSTEP * sm.res[2 + which] + i, n);
    }
    __syncthreads();
    if (tid < 2 * DSTEP) {
        const int which = tid / DSTEP, i = tid % DSTEP, ts = sm.res[2 + which];
        const uint32_t need = (which ? k2 : k1) + 1;
        const uint32_t prev = i ? sm.fr2[which][i - 1] : (ts ? sm.fr[ts - 1] : 0u);
        if (sm.fr2[which][i] >= need && prev < need) sm.res[which] = DSTEP * ts + i;      // res[0 / 1] are free again
    }
    __syncthreads();
    sel[0] = m1; sel[1] = m2; sel[2] = sm.res[0]; sel[3] = sm.res[1];
    __syncthreads();
}

// status checks of one read by the whole CTA (boundary error convention, include/nrv.h); sm_bad is a shared flag, zero on entry
__device__ __forceinline__ int rs_status_checks(int* sm_bad, const int64_t* __restrict__ base_off, const int32_t* __restrict__ starts,
                                                const int32_t* __restrict__ last_dur, int r, long long n, int window, int nthreads) {
    const long long b0 = base_off[r], nb = base_off[r + 1] - b0;
    for (long long j = threadIdx.x; j < nb; j += nthreads) {
        const long long st = starts[b0 + j];
        const long long en = (j + 1 < nb) ? (long long)starts[b0 + j + 1] : st + last_dur[r];
        if (st < 0 || en <= st || en > n) *sm_bad = 1;
    }
    __syncthreads();
    int st_code = NRV_READ_OK;
    if (*sm_bad || n <= 0 || nb <= 0 || n >= (1ll << 32)) st_code = NRV_READ_BAD_EVENTS;
    else if (nb <= window) st_code = NRV_READ_TOO_SHORT;
    return st_code;
}

// Full-range kernel (round-2 first version, now the fall-back): only reads whose status is RS_PENDING are processed.
__global__ void __launch_bounds__(RsFull::THREADS, 1)
read_stats_kernel(const int16_t* __restrict__ signal, const int64_t* __restrict__ sig_off,
                  const int64_t* __restrict__ base_off, const int32_t* __restrict__ starts,
                  const int32_t* __restrict__ last_dur, int window, const int32_t* __restrict__ hist_slot,
                  uint32_t* __restrict__ ghist, uint32_t* __restrict__ gdone,
                  double* __restrict__ shift_out, double* __restrict__ scale_out, int32_t* __restrict__ status) {
    constexpr int RS_THREADS = RsFull::THREADS, RS_BINS = RsFull::NBINS;
    extern __shared__ __align__(16) unsigned char rs_raw[];
    RsFullSmem& sm = *reinterpret_cast<RsFullSmem*>(rs_raw);
    const int r = blockIdx.x, seg = blockIdx.y;
    const int tid = threadIdx.x;
    if (status[r] != RS_PENDING) return;              // (every CTA of the read sees the same value: the last CTA writes it at its very end,
                                                      //  after all others have arrived at gdone and left)
    const long long o0 = sig_off[r];
    const long long n = sig_off[r + 1] - o0;
    const long long nseg = n <= RS_SEG ? 1 : (n + RS_SEG - 1) / RS_SEG;
    if (seg >= nseg) return;
    const int slot = hist_slot ? hist_slot[r] : -1;
    if (nseg > 1 && slot < 0) return;                 // cannot happen: the host assigns a slot to every multi-segment read

    for (int i = tid; i < RS_BINS / 2; i += RS_THREADS) sm.h[i] = 0;
    if (tid == 0) { sm.bad = 0; sm.is_last = 0; }
    __syncthreads();
    {   // ---- the one pass: this segment's samples into the 16-bit histogram; 128-bit loads on the 16-byte aligned body ----
        const int16_t* p0 = signal + o0 + (long long)seg * RS_SEG;
        const long long cnt = min((long long)RS_SEG, n - (long long)seg * RS_SEG);
        const int16_t* pend = p0 + cnt;
        const int16_t* body = reinterpret_cast<const int16_t*>((reinterpret_cast<uintptr_t>(p0) + 15) & ~(uintptr_t)15);
        if (body > pend) body = pend;
        const int16_t* bend = body + ((pend - body) & ~(long long)7);
        auto add = [&](int sv) {
            const unsigned key = (unsigned)(sv + 32768);
            atomicAdd(&sm.h[key >> 1], 1u << ((key & 1) * 16));
        };
        if (p0 + tid < body) add(p0[tid]);                                   // head: fewer than 8 samples
        for (const int16_t* q = body + (long long)tid * 8; q < bend; q += (long long)RS_THREADS * 8) {
            const uint4 v4 = __ldg(reinterpret_cast<const uint4*>(q));
            const uint32_t w[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) { add((int)(int16_t)(w[i] & 0xffffu)); add((int)(int16_t)(w[i] >> 16)); }
        }
        if (bend + tid < pend) add(bend[tid]);                               // tail
    }
    __syncthreads();
    if (nseg > 1) {
        uint32_t* g = ghist + (size_t)slot * RS_BINS;
        for (int i = tid; i < RS_BINS / 2; i += RS_THREADS) {
            const uint32_t v = sm.h[i];
            if (v & 0xffffu) atomicAdd(&g[2 * i], v & 0xffffu);
            if (v >> 16) atomicAdd(&g[2 * i + 1], v >> 16);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) sm.is_last = atomicAdd(&gdone[slot], 1u) == (unsigned)(nseg - 1);
        __syncthreads();
        if (!sm.is_last) return;
        __threadfence();
    }
    // ---- this CTA finishes the read: status checks, then the selection ----------
    int st_code = rs_status_checks(&sm.bad, base_off, starts, last_dur, r, n, window, RS_THREADS);
    if (n <= 0 || n >= (1ll << 32)) {
        if (tid == 0) { shift_out[r] = 0.0; scale_out[r] = 0.0; status[r] = st_code; }
    } else {
        int sel[4];
        if (nseg > 1) rs_select<RsFull>(RsSrcGlobal{ghist + (size_t)slot * RS_BINS}, sm, (uint32_t)n, sel);
        else rs_select<RsFull>(RsSrcShared{sm.h}, sm, (uint32_t)n, sel);
        if (tid == 0) {
            const double shift = (double)(sel[0] + sel[1]) * 0.5 - 32768.0;       // mean of the two middles, exact
            const double scale = (double)(sel[2] + sel[3]) * 0.25;                 // mean of two |x - shift|, exact
            shift_out[r] = shift;
            scale_out[r] = scale;
            if (st_code == NRV_READ_OK && !(scale > 0.0)) st_code = NRV_READ_SCALE_ZERO;
        }
    }
    if (nseg > 1) {            // leave the read's global histogram and counter zeroed for the next batch
        uint4* g4 = reinterpret_cast<uint4*>(ghist + (size_t)slot * RS_BINS);
        for (int i = tid; i < RS_BINS / 4; i += RS_THREADS) g4[i] = make_uint4(0, 0, 0, 0);
        if (tid == 0) gdone[slot] = 0;
    }
    __syncthreads();
    if (tid == 0 && n > 0 && n < (1ll << 32)) status[r] = st_code;          // last: the other CTAs of the read test it on entry
}

// Compact kernel (the one that normally does the work).  Raw nanopore samples are 13-bit ADC codes, so the histogram covers
// -8191 .. 8190 with one 32-bit counter per value (64 KB instead of 128 KB: three CTAs per SM, a quarter of the bins to clear,
// merge, search and re-zero) and CLAMPS everything else into its two end bins.  Cumulative counts are exact for every threshold
// inside the covered range, so the two middle order statistics and the smallest d with #{|2 key - shift2| <= d} >= k + 1 are
// exact whenever they and the band [lo(d), hi(d)] lie strictly inside it; if not (more than half of a read beyond +-8191, or a
// MAD in the thousands), the read is marked RS_PENDING and the full-range kernel, launched right behind, redoes it.
__global__ void __launch_bounds__(RsCompact::THREADS, 3)
read_stats_compact_kernel(const int16_t* __restrict__ signal, const int64_t* __restrict__ sig_off,
                          const int64_t* __restrict__ base_off, const int32_t* __restrict__ starts,
                          const int32_t* __restrict__ last_dur, int window, const int32_t* __restrict__ hist_slot,
                          uint32_t* __restrict__ ghist, uint32_t* __restrict__ gdone,
                          double* __restrict__ shift_out, double* __restrict__ scale_out, int32_t* __restrict__ status,
                          int32_t* __restrict__ base_read) {
    constexpr int TH = RsCompact::THREADS, NB = RsCompact::NBINS;
    extern __shared__ __align__(16) unsigned char rs_raw[];
    RsCompactSmem& sm = *reinterpret_cast<RsCompactSmem*>(rs_raw);
    const int r = blockIdx.x, seg = blockIdx.y;
    const int tid = threadIdx.x;
    const long long o0 = sig_off[r];
    const long long n = sig_off[r + 1] - o0;
    const long long nseg = n <= RS_SEG ? 1 : (n + RS_SEG - 1) / RS_SEG;
    if (seg >= nseg) return;
    const int slot = hist_slot ? hist_slot[r] : -1;
    if (nseg > 1 && slot < 0) return;                 // cannot happen: the host assigns a slot to every multi-segment read

    // ---- the one pass over this segment's samples: 128-bit loads on the 16-byte aligned body ----
    const int16_t* p0 = signal + o0 + (long long)seg * RS_SEG;
    const long long cnt = n > 0 ? min((long long)RS_SEG, n - (long long)seg * RS_SEG) : 0;
    const int16_t* pend = p0 + cnt;
    const int16_t* body = reinterpret_cast<const int16_t*>((reinterpret_cast<uintptr_t>(p0) + 15) & ~(uintptr_t)15);
    if (body > pend) body = pend;
    const int16_t* bend = body + ((pend - body) & ~(long long)7);
    // the first two loads of every thread go out before anything else (software pipeline: two more are issued before these are used)
    const int16_t* q = body + (long long)tid * 8;
    constexpr long long QS = (long long)TH * 8;
    uint4 va = q < bend ? __ldg(reinterpret_cast<const uint4*>(q)) : make_uint4(0, 0, 0, 0);
    uint4 vb = q + QS < bend ? __ldg(reinterpret_cast<const uint4*>(q + QS)) : make_uint4(0, 0, 0, 0);
    {
        uint4* h4 = reinterpret_cast<uint4*>(sm.h);
        for (int i = tid; i < NB / 4; i += TH) h4[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) { sm.bad = 0; sm.is_last = 0; }
    __syncthreads();
    // this CTA's share of the read's bases (the segments of a read split them evenly): the base -> read map and the status checks
    // (boundary error convention, include/nrv.h), under the loads
    const long long b0 = base_off[r], nb = base_off[r + 1] - b0;
    {
        const long long s0 = nb * seg / nseg, s1 = nb * (seg + 1) / nseg;
        const int ldur = last_dur[r];
        bool lbad = false;
#pragma unroll 4
        for (long long j = s0 + tid; j < s1; j += TH) {
            const long long st = starts[b0 + j];
            const long long en = (j + 1 < nb) ? (long long)starts[b0 + j + 1] : st + ldur;
            lbad |= (st < 0 || en <= st || en > n);
            if (base_read) base_read[b0 + j] = r;
        }
        if (lbad) sm.bad = 1;
    }
    {
        auto add = [&](int sv) { atomicAdd(&sm.h[min(max(sv + RC_OFF, 0), NB - 1)], 1u); };
        auto add8 = [&](const uint4& v4) {
            const uint32_t w[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) { add((int)(int16_t)(w[e] & 0xffffu)); add((int)(int16_t)(w[e] >> 16)); }
        };
        if (p0 + tid < body) add(p0[tid]);                                   // head: fewer than 8 samples
        for (; q < bend; q += 2 * QS) {
            const int16_t* q2 = q + 2 * QS;
            const uint4 na = q2 < bend ? __ldg(reinterpret_cast<const uint4*>(q2)) : make_uint4(0, 0, 0, 0);
            const uint4 nb2 = q2 + QS < bend ? __ldg(reinterpret_cast<const uint4*>(q2 + QS)) : make_uint4(0, 0, 0, 0);
            add8(va);
            if (q + QS < bend) add8(vb);
            va = na; vb = nb2;
        }
        if (bend + tid < pend) add(bend[tid]);                               // tail
    }
    __syncthreads();
    int bad = sm.bad;
    if (nseg > 1) {
        uint32_t* g = ghist + (size_t)slot * RsFull::NBINS;                  // the compact histogram uses the head of the read's slot
        uint4* h4 = reinterpret_cast<uint4*>(sm.h);
        for (int i = tid; i < NB / 4; i += TH) {
            const uint4 v = h4[i];
            if (v.x) atomicAdd(&g[4 * i], v.x);
            if (v.y) atomicAdd(&g[4 * i + 1], v.y);
            if (v.z) atomicAdd(&g[4 * i + 2], v.z);
            if (v.w) atomicAdd(&g[4 * i + 3], v.w);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {                               // arrival counter in the low half, "a share of the bases was bad" above it
            const uint32_t old = atomicAdd(&gdone[slot], 1u + (bad ? 0x10000u : 0u));
            sm.is_last = (old & 0xffffu) == (unsigned)(nseg - 1);
            sm.bad = bad || (old >> 16) != 0;
        }
        __syncthreads();
        if (!sm.is_last) return;
        __threadfence();
        bad = sm.bad;
        // the merged histogram comes into shared memory (one round trip, all loads independent) and leaves global memory zeroed
        // for the fall-back and the next batch; the search below then runs on shared memory
        uint4* g4 = reinterpret_cast<uint4*>(g);
#pragma unroll
        for (int i = 0; i < NB / 4 / TH; ++i) h4[i * TH + tid] = __ldcg(g4 + i * TH + tid);     // (independent loads: not one loop with the stores)
#pragma unroll
        for (int i = 0; i < NB / 4 / TH; ++i) g4[i * TH + tid] = make_uint4(0, 0, 0, 0);
        if (tid == 0) gdone[slot] = 0;
        __syncthreads();
    }
    // ---- this CTA finishes the read ----------
    int st_code = NRV_READ_OK;
    if (bad || n <= 0 || nb <= 0 || n >= (1ll << 32)) st_code = NRV_READ_BAD_EVENTS;
    else if (nb <= window) st_code = NRV_READ_TOO_SHORT;
    if (n <= 0 || n >= (1ll << 32)) {
        if (tid == 0) { shift_out[r] = 0.0; scale_out[r] = 0.0; status[r] = st_code; }
        return;
    }
    int sel[4];
    rs_select<RsCompact>(RsSrcShared32{sm.h}, sm, (uint32_t)n, sel);
    if (tid == 0) {
        const int shift2 = sel[0] + sel[1];
        const int lo = (shift2 - sel[3] + 1) >> 1, hi = (shift2 + sel[3]) >> 1;     // band of the larger of the two deviations
        const bool exact = sel[0] >= 1 && sel[1] <= NB - 2 && lo >= 1 && hi <= NB - 2;
        if (!exact) {
            status[r] = RS_PENDING;
        } else {
            const double shift = (double)shift2 * 0.5 - (double)RC_OFF;    // mean of the two middles, exact
            const double scale = (double)(sel[2] + sel[3]) * 0.25;          // mean of two |x - shift|, exact
            shift_out[r] = shift;
            scale_out[r] = scale;
            if (st_code == NRV_READ_OK && !(scale > 0.0)) st_code = NRV_READ_SCALE_ZERO;
            status[r] = st_code;
        }
    }
}

size_t read_stats_hist_bytes(int64_t n_multi) { return (size_t)n_multi * RsFull::NBINS * 4 + (size_t)n_multi * 4 + 16; }
int read_stats_segment() { return RS_SEG; }

// hist_slot[r] = index of read r among the reads with more than RS_SEG samples (else -1); ghist = zeroed scratch of
// read_stats_hist_bytes(n_multi) bytes (histograms, then the arrival counters); max_segs = segments of the longest read
int launch_read_stats(const int16_t* signal, const int64_t* sig_off, const int64_t* base_off,
                      const int32_t* starts, const int32_t* last_dur, int window, int64_t n_reads,
                      const int32_t* hist_slot, void* ghist, int64_t n_multi, int max_segs,
                      double* shift, double* scale, int32_t* status, cudaStream_t st, int32_t* base_read, int64_t n_bases) {
    if (n_reads <= 0) return 0;
    static PerDevice attr_set;
    if (attr_set.first()) {
        cudaFuncSetAttribute(read_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsFullSmem));
        cudaFuncSetAttribute(read_stats_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsCompactSmem));
    }
    uint32_t* g = reinterpret_cast<uint32_t*>(ghist);
    uint32_t* done = g ? g + (size_t)n_multi * RsFull::NBINS : nullptr;
    const dim3 grid((unsigned)n_reads, (unsigned)std::max(max_segs, 1));
    const bool full_only = getenv("NRV_READ_STATS") && !strcmp(getenv("NRV_READ_STATS"), "full");       // tests: the fall-back on every read
    int launches = 2;
    if (full_only) {
        cudaMemsetAsync(status, 0xff, (size_t)n_reads * 4, st);                                         // RS_PENDING everywhere
        if (base_read) launches += launch_base_read_map(base_off, n_reads, n_bases, base_read, st);
    } else {
        read_stats_compact_kernel<<<grid, RsCompact::THREADS, sizeof(RsCompactSmem), st>>>(
            signal, sig_off, base_off, starts, last_dur, window, hist_slot, g, done, shift, scale, status, base_read);
    }
    read_stats_kernel<<<grid, RsFull::THREADS, sizeof(RsFullSmem), st>>>(
        signal, sig_off, base_off, starts, last_dur, window, hist_slot, g, done, shift, scale, status);
    return launches;
    return 1;
}

// ------------------------------------------------------------------------------------------
// base -> read map and window -> first-base map
// ------------------------------------------------------------------------------------------
__global__ void base_read_map_kernel(const int64_t* __restrict__ base_off, int64_t n_reads, int64_t n_bases,
                                     int32_t* __restrict__ base_read) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_bases) return;
    int64_t lo = 0, hi = n_reads;            // largest r with base_off[r] <= j
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (base_off[mid] <= j) lo = mid; else hi = mid;
    }
    base_read[j] = (int32_t)lo;
}

int launch_base_read_map(const int64_t* base_off, int64_t n_reads, int64_t n_bases, int32_t* base_read,
                         cudaStream_t st) {
    if (n_bases <= 0) return 0;
    base_read_map_kernel<<<(unsigned)((n_bases + 255) / 256), 256, 0, st>>>(base_off, n_reads, n_bases, base_read);
    return 1;
}

// win_off[r] .. win_off[r+1] are the windows of read r (N_r - W of them, nanorevtrainutils.py:198);
// window i of read r starts at base base_off[r] + i.  Reads with a non-zero status keep their window
// slots (so that window numbering is status independent) -- their labels are simply never used.
__global__ void window_map_kernel(const int64_t* __restrict__ base_off, const int64_t* __restrict__ win_off,
                                  int64_t n_reads, int32_t* __restrict__ win_base) {
    const int r = blockIdx.x;
    const int64_t w0 = win_off[r], nw = win_off[r + 1] - w0, b0 = base_off[r];
    for (int64_t i = threadIdx.x; i < nw; i += blockDim.x) win_base[w0 + i] = (int32_t)(b0 + i);
}

int launch_window_map(const int64_t* base_off, const int64_t* win_off, const int32_t* /*status*/,
                      int64_t n_reads, int /*window*/, int32_t* win_base, cudaStream_t st) {
    if (n_reads <= 0) return 0;
    window_map_kernel<<<(unsigned)n_reads, 256, 0, st>>>(base_off, win_off, n_reads, win_base);
    return 1;
}

// ------------------------------------------------------------------------------------------
// per-base raw mean / std and the six feature columns
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float base_colour(uint8_t b) {   // preprocessing.py:173-175: colour / 300.0 in float64, cast to fp32
    return (b == 'A') ? (float)(250.0 / 300.0) : (b == 'G') ? (float)(180.0 / 300.0) : (b == 'T') ? (float)(100.0 / 300.0)
         : (b == 'C') ? (float)(30.0 / 300.0) : 0.0f;
}

constexpr int LONG_SEG = 128;
constexpr int BF_THREADS = 256;
constexpr int BF_STAGE = 8192;       // samples of one CTA's 256 bases staged in shared memory (mean 2,300; more: read from global)

// (double)x for a 32-bit integer without the conversion instruction (I2F.F64 runs on the quarter-rate XU pipe and was 51 % of this
// kernel's time, one per sample): 2^52 + 2^31 + x is exactly representable, its bits are 0x43300000 : (x ^ 0x80000000)
__device__ __forceinline__ double int_to_double(int x) {
    return __hiloint2double(0x43300000, (int)((unsigned)x ^ 0x80000000u)) - 4503601774854144.0;
}
// np.mean (exact integer sum / n) and np.std**2 (mean(|x - mean|^2), ddof = 0) of n samples, in the order the reference walks them
__device__ __forceinline__ void bf_moments(const int16_t* p, int n, double* mean_out, double* var_out) {
    long long s = 0;
    for (int i = 0; i < n; ++i) s += p[i];
    const double mean = (double)s / (double)n;
    double q = 0.0;
    for (int i = 0; i < n; ++i) { const double d = int_to_double(p[i]) - mean; q += d * d; }
    *mean_out = mean;
    *var_out = q / (double)n;
}

// Round 2: the samples of the CTA's 256 consecutive bases are one contiguous range of the read (~4.6 KB): it is brought into
// shared memory with 128-bit loads and the per-base loops (two passes, ~9 samples each, a different length in every lane) run
// from there instead of issuing scattered 2-byte global loads.  CTAs that straddle two reads, or whose range is longer than
// BF_STAGE samples, read from global memory as before.
__global__ void __launch_bounds__(BF_THREADS)
base_features_kernel(const int16_t* __restrict__ signal, const int64_t* __restrict__ sig_off,
                     const int32_t* __restrict__ starts, const int64_t* __restrict__ base_off,
                     const uint8_t* __restrict__ bases, const float* __restrict__ ev_mean,
                     const float* __restrict__ ev_std, const int32_t* __restrict__ last_dur,
                     const int32_t* __restrict__ base_read, const double* __restrict__ shift,
                     const double* __restrict__ scale, int64_t n_bases, float* __restrict__ x,
                     double* __restrict__ seg_mean, double* __restrict__ seg_std) {
    __shared__ __align__(16) int16_t s_sig[BF_STAGE + 16];
    __shared__ long long s_lo, s_hi;
    __shared__ int s_r[2];
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = j < n_bases;
    const int lane = threadIdx.x & 31;
    int r = 0;
    long long st = 0, en = 0, S = 0;
    long long dur = 0;
    const int16_t* sig = signal;
    float evm = 0.f, evs = 0.f;
    uint8_t bch = 0;
    if (active) {
        r = base_read[j];
        sig = signal + sig_off[r];
        S = sig_off[r + 1] - sig_off[r];
        st = starts[j];
        en = (j + 1 < base_off[r + 1]) ? (long long)starts[j + 1] : st + last_dur[r];
        dur = en - st;                                   // before clamping: nanorevtrainutils.py:164 uses the event table's length
        if (x) { evm = ev_mean[j]; evs = ev_std[j]; bch = bases[j]; }
        // memory safety for reads flagged NRV_READ_BAD_EVENTS (their features are never used)
        if (st < 0) st = 0;
        if (en > S) en = S;
        if (en < st) en = st;
    }
    const int64_t j_last = min((int64_t)(blockIdx.x + 1) * blockDim.x, n_bases) - 1;
    if (threadIdx.x == 0) { s_r[0] = r; s_lo = st; }
    if (j == j_last) { s_r[1] = r; s_hi = en; }
    __syncthreads();
    const long long lo = s_lo, hi = s_hi;
    const bool staged = s_r[0] == s_r[1] && hi > lo && hi - lo <= BF_STAGE;
    int off0 = 0;
    if (staged) {
        // the CTA's read is the one of thread 0 (= of every thread); chunks of 16 bytes at their global alignment
        const int16_t* rsig = signal + sig_off[s_r[0]];
        const long long rS = sig_off[s_r[0] + 1] - sig_off[s_r[0]];
        const uintptr_t A = reinterpret_cast<uintptr_t>(rsig + lo), A0 = A & ~(uintptr_t)15;
        const uintptr_t B = reinterpret_cast<uintptr_t>(rsig + hi);
        const uintptr_t R0 = reinterpret_cast<uintptr_t>(rsig), R1 = reinterpret_cast<uintptr_t>(rsig + rS);
        off0 = (int)((A - A0) >> 1);
        const int nchunks = (int)((B - A0 + 15) >> 4);
        for (int c = threadIdx.x; c < nchunks; c += BF_THREADS) {
            const uintptr_t a = A0 + (uintptr_t)c * 16;
            if (a >= R0 && a + 16 <= R1) {
                *reinterpret_cast<uint4*>(&s_sig[c * 8]) = __ldg(reinterpret_cast<const uint4*>(a));
            } else {                                     // first / last chunk of the read: only the samples that belong to it
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const uintptr_t ae = a + 2 * e;
                    s_sig[c * 8 + e] = (ae >= R0 && ae + 2 <= R1) ? __ldg(reinterpret_cast<const int16_t*>(ae)) : (int16_t)0;
                }
            }
        }
    }
    __syncthreads();
    const long long n = en - st;
    double mean = 0.0, var = 0.0;
    const bool is_long = active && n > LONG_SEG;
    if (active && !is_long && n > 0) {
        if (staged && st >= lo && en <= hi) bf_moments(&s_sig[(int)(st - lo) + off0], (int)n, &mean, &var);
        else bf_moments(sig + st, (int)n, &mean, &var);
    }
    // stalled bases (hundreds to ~1e5 samples): the whole warp walks the segment together
    unsigned long_mask = __ballot_sync(0xffffffffu, is_long);
    while (long_mask) {
        const int src = __ffs(long_mask) - 1;
        long_mask &= long_mask - 1;
        const long long sst = __shfl_sync(0xffffffffu, st, src);
        const long long sen = __shfl_sync(0xffffffffu, en, src);
        const int16_t* ssig = (const int16_t*)__shfl_sync(0xffffffffu, (unsigned long long)sig, src);
        long long s = 0;
        for (long long i = sst + lane; i < sen; i += 32) s += ssig[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const double m = (double)s / (double)(sen - sst);
        double q = 0.0;
        for (long long i = sst + lane; i < sen; i += 32) { double d = int_to_double(ssig[i]) - m; q += d * d; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        if (lane == src) { mean = m; var = q / (double)(sen - sst); }
    }
    if (!active) return;
    const double sd = sqrt(var);
    if (seg_mean) seg_mean[j] = mean;
    if (seg_std) seg_std[j] = sd;
    if (x) {
        const double sh = shift[r], sc = scale[r];
        float* o = x + j * 6;
        o[0] = base_colour(bch);
        o[1] = (float)(mean / sh);                       // NanoReviser.py:124
        o[2] = (float)(sd / sc);                         // NanoReviser.py:125
        o[3] = (float)((double)dur / 10.0);              // nanorevtrainutils.py:164
        o[4] = evm;
        o[5] = evs;
    }
}

int launch_base_features(const int16_t* signal, const int64_t* sig_off, const int32_t* starts,
                         const int64_t* base_off, const uint8_t* bases, const float* ev_mean,
                         const float* ev_std, const int32_t* last_dur, const int32_t* base_read,
                         const double* shift, const double* scale, int64_t n_bases,
                         float* x, double* seg_mean, double* seg_std, cudaStream_t st) {
    if (n_bases <= 0) return 0;
    base_features_kernel<<<(unsigned)((n_bases + BF_THREADS - 1) / BF_THREADS), BF_THREADS, 0, st>>>(
        signal, sig_off, starts, base_off, bases, ev_mean, ev_std, last_dur, base_read, shift, scale,
        n_bases, x, seg_mean, seg_std);
    return 1;
}

// ------------------------------------------------------------------------------------------
// materialised windows (parity entry point only)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sig_windows_kernel(const int16_t* __restrict__ signal, const int64_t* __restrict__ sig_off,
                   const int32_t* __restrict__ starts, const int32_t* __restrict__ base_read,
                   const double* __restrict__ shift, const double* __restrict__ scale, int64_t n_bases,
                   float* __restrict__ sig_win) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_bases * NRV_SIG) return;
    const int64_t j = idx / NRV_SIG;
    const int p = (int)(idx - j * NRV_SIG);
    const int r = base_read[j];
    const int16_t* sig = signal + sig_off[r];
    const long long S = sig_off[r + 1] - sig_off[r];
    const long long st = starts[j];
    const long long lo = (st - 25 <= 0) ? 0 : st - 25;           // preprocessing.py:111-114
    long long hi = (st + 25 >= S) ? S : st + 25;                 // :115-118
    if (hi < lo) hi = lo;
    const int len = (int)(hi - lo);
    const int pad = NRV_SIG - len;
    const int left = (pad + 1) / 2;                              // :120-131 (odd pad: one more on the left)
    float v = 0.f;
    if (p >= left && p < left + len) v = (float)(((double)sig[lo + (p - left)] - shift[r]) / scale[r]);
    sig_win[idx] = v;
}

int launch_sig_windows(const int16_t* signal, const int64_t* sig_off, const int32_t* starts,
                       const int64_t* /*base_off*/, const int32_t* base_read, const double* shift,
                       const double* scale, int64_t n_bases, float* sig_win, cudaStream_t st) {
    if (n_bases <= 0) return 0;
    const int64_t tot = n_bases * NRV_SIG;
    sig_windows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(signal, sig_off, starts, base_read, shift,
                                                                     scale, n_bases, sig_win);
    return 1;
}

}  // namespace nrv
