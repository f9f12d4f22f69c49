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
constexpr int RS_THREADS = 1024;
constexpr int RS_BINS = 65536;

struct RsSmem {
    uint32_t h[RS_BINS / 2];          // 16-bit counters, bin k in the (k & 1) half of word k >> 1
    uint32_t p16[RS_BINS / 16 + 1];   // exclusive cumulative count before every chunk of 16 bins; [4096] = n
    uint32_t fr[RS_THREADS];
    uint32_t fr2[2][128];
    uint32_t warp_tot[RS_THREADS / 32];
    int res[4];
    int bad, is_last;
};

// the 16 bins of chunk c in registers: 2 x LDS.128 (16-bit counters) resp. 4 x 128-bit L2 loads (no dependent scalar loads)
struct RsSrcShared {
    const uint32_t* h;
    __device__ __forceinline__ void chunk(int c, uint32_t (&b)[16]) const {
        const uint4 a0 = *reinterpret_cast<const uint4*>(h + c * 8), a1 = *reinterpret_cast<const uint4*>(h + c * 8 + 4);
        const uint32_t w[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) { b[2 * i] = w[i] & 0xffffu; b[2 * i + 1] = w[i] >> 16; }
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
template <class Src>
__device__ __forceinline__ uint32_t rs_cum(const Src& src, const RsSmem& sm, int x, uint32_t n) {
    if (x < 0) return 0;
    if (x >= RS_BINS) return n;
    return sm.p16[x >> 4] + rs_chunk_sum(src, x >> 4, x & 15);
}
// number of samples with |2 key - shift2| <= d
template <class Src>
__device__ __forceinline__ uint32_t rs_within(const Src& src, const RsSmem& sm, int shift2, int d, uint32_t n) {
    const int lo = (shift2 - d + 1) >> 1, hi = (shift2 + d) >> 1;      // ceil / floor (arithmetic shifts)
    return rs_cum(src, sm, hi, n) - rs_cum(src, sm, lo - 1, n);
}

// whole CTA: shift2 = sum of the keys of the two middle order statistics, dev4 = sum of the two middle |2 key - shift2|
template <class Src>
__device__ void rs_select(const Src& src, RsSmem& sm, uint32_t n, int* shift2_out, int* dev4_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t k1 = (n - 1) / 2, k2 = n / 2;
    // cumulative counts per chunk of 16 bins: thread t owns chunks 4t .. 4t + 3
    uint32_t cs[4], local = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { cs[j] = rs_chunk_sum(src, 4 * tid + j, 15); local += cs[j]; }
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
    for (int j = 0; j < 4; ++j) { sm.p16[4 * tid + j] = run; run += cs[j]; }
    if (tid == RS_THREADS - 1) sm.p16[RS_BINS / 16] = run;
    __syncthreads();
    // the two middle order statistics
    run = base + v - local;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const uint32_t k = which ? k2 : k1;
            if (k >= run && k < run + cs[j]) {
                uint32_t b[16];
                src.chunk(4 * tid + j, b);
                uint32_t r2 = run;
                int found = 15;
#pragma unroll
                for (int i = 15; i >= 0; --i) {          // first bin whose cumulative count exceeds k
                    uint32_t upto = 0;
#pragma unroll
                    for (int q = 0; q <= i; ++q) upto += b[q];
                    if (k < r2 + upto) found = i;
                }
                sm.res[which] = (4 * tid + j) * 16 + found;
            }
        }
        run += cs[j];
    }
    __syncthreads();
    const int shift2 = sm.res[0] + sm.res[1];
    // k-th smallest deviation: smallest d with within(d) >= k + 1; d in [0, 131071], round 1 at d = 128 t + 127
    sm.fr[tid] = rs_within(src, sm, shift2, 128 * tid + 127, n);
    __syncthreads();
#pragma unroll
    for (int which = 0; which < 2; ++which) {
        const uint32_t need = (which ? k2 : k1) + 1;
        if (sm.fr[tid] >= need && (tid == 0 || sm.fr[tid - 1] < need)) sm.res[2 + which] = tid;
    }
    __syncthreads();
    if (tid < 256) {
        const int which = tid >> 7, i = tid & 127;
        sm.fr2[which][i] = rs_within(src, sm, shift2, 128 * sm.res[2 + which] + i, n);
    }
    __syncthreads();
    if (tid < 256) {
        const int which = tid >> 7, i = tid & 127, ts = sm.res[2 + which];
        const uint32_t need = (which ? k2 : k1) + 1;
        const uint32_t prev = i ? sm.fr2[which][i - 1] : (ts ? sm.fr[ts - 1] : 0u);
        if (sm.fr2[which][i] >= need && prev < need) sm.res[which] = 128 * ts + i;      // res[0 / 1] are free again
    }
    __syncthreads();
    *shift2_out = shift2;
    *dev4_out = sm.res[0] + sm.res[1];
    __syncthreads();
}

__global__ void __launch_bounds__(RS_THREADS, 1)
read_stats_kernel(const int16_t* __restrict__ signal, const int64_t* __restrict__ sig_off,
                  const int64_t* __restrict__ base_off, const int32_t* __restrict__ starts,
                  const int32_t* __restrict__ last_dur, int window, const int32_t* __restrict__ hist_slot,
                  uint32_t* __restrict__ ghist, uint32_t* __restrict__ gdone,
                  double* __restrict__ shift_out, double* __restrict__ scale_out, int32_t* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char rs_raw[];
    RsSmem& sm = *reinterpret_cast<RsSmem*>(rs_raw);
    const int r = blockIdx.x, seg = blockIdx.y;
    const int tid = threadIdx.x;
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
    // ---- this CTA finishes the read: status checks (boundary error convention, include/nrv.h), then the selection ----------
    const long long b0 = base_off[r], nb = base_off[r + 1] - b0;
    for (long long j = tid; j < nb; j += RS_THREADS) {
        const long long st = starts[b0 + j];
        const long long en = (j + 1 < nb) ? (long long)starts[b0 + j + 1] : st + last_dur[r];
        if (st < 0 || en <= st || en > n) sm.bad = 1;
    }
    __syncthreads();
    int st_code = NRV_READ_OK;
    if (sm.bad || n <= 0 || nb <= 0 || n >= (1ll << 32)) st_code = NRV_READ_BAD_EVENTS;
    else if (nb <= window) st_code = NRV_READ_TOO_SHORT;
    if (n <= 0 || n >= (1ll << 32)) {
        if (tid == 0) { shift_out[r] = 0.0; scale_out[r] = 0.0; status[r] = st_code; }
    } else {
        int shift2, dev4;
        if (nseg > 1) rs_select(RsSrcGlobal{ghist + (size_t)slot * RS_BINS}, sm, (uint32_t)n, &shift2, &dev4);
        else rs_select(RsSrcShared{sm.h}, sm, (uint32_t)n, &shift2, &dev4);
        if (tid == 0) {
            const double shift = (double)shift2 * 0.5 - 32768.0;       // mean of the two middles, exact
            const double scale = (double)dev4 * 0.25;                  // mean of two |x - shift|, exact
            shift_out[r] = shift;
            scale_out[r] = scale;
            if (st_code == NRV_READ_OK && !(scale > 0.0)) st_code = NRV_READ_SCALE_ZERO;
            status[r] = st_code;
        }
    }
    if (nseg > 1) {            // leave the read's global histogram and counter zeroed for the next batch
        uint4* g4 = reinterpret_cast<uint4*>(ghist + (size_t)slot * RS_BINS);
        for (int i = tid; i < RS_BINS / 4; i += RS_THREADS) g4[i] = make_uint4(0, 0, 0, 0);
        if (tid == 0) gdone[slot] = 0;
    }
}

size_t read_stats_hist_bytes(int64_t n_multi) { return (size_t)n_multi * RS_BINS * 4 + (size_t)n_multi * 4 + 16; }
int read_stats_segment() { return RS_SEG; }

// hist_slot[r] = index of read r among the reads with more than RS_SEG samples (else -1); ghist = zeroed scratch of
// read_stats_hist_bytes(n_multi) bytes (histograms, then the arrival counters); max_segs = segments of the longest read
int launch_read_stats(const int16_t* signal, const int64_t* sig_off, const int64_t* base_off,
                      const int32_t* starts, const int32_t* last_dur, int window, int64_t n_reads,
                      const int32_t* hist_slot, void* ghist, int64_t n_multi, int max_segs,
                      double* shift, double* scale, int32_t* status, cudaStream_t st) {
    if (n_reads <= 0) return 0;
    static PerDevice attr_set;
    if (attr_set.first()) cudaFuncSetAttribute(read_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem));
    uint32_t* g = reinterpret_cast<uint32_t*>(ghist);
    uint32_t* done = g ? g + (size_t)n_multi * RS_BINS : nullptr;
    read_stats_kernel<<<dim3((unsigned)n_reads, (unsigned)std::max(max_segs, 1)), RS_THREADS, sizeof(RsSmem), st>>>(
        signal, sig_off, base_off, starts, last_dur, window, hist_slot, g, done, shift, scale, status);
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
__device__ __forceinline__ float base_colour(uint8_t b) {   // preprocessing.py:173-175
    int c = (b == 'A') ? 250 : (b == 'G') ? 180 : (b == 'T') ? 100 : (b == 'C') ? 30 : 0;
    return (float)((double)c / 300.0);
}

constexpr int LONG_SEG = 128;

// (double)x for a 32-bit integer without the conversion instruction (I2F.F64 runs on the quarter-rate XU pipe and was 51 % of this
// kernel's time, one per sample): 2^52 + 2^31 + x is exactly representable, its bits are 0x43300000 : (x ^ 0x80000000)
__device__ __forceinline__ double int_to_double(int x) {
    return __hiloint2double(0x43300000, (int)((unsigned)x ^ 0x80000000u)) - 4503601774854144.0;
}

__global__ void __launch_bounds__(256)
base_features_kernel(const int16_t* __restrict__ signal, const int64_t* __restrict__ sig_off,
                     const int32_t* __restrict__ starts, const int64_t* __restrict__ base_off,
                     const uint8_t* __restrict__ bases, const float* __restrict__ ev_mean,
                     const float* __restrict__ ev_std, const int32_t* __restrict__ last_dur,
                     const int32_t* __restrict__ base_read, const double* __restrict__ shift,
                     const double* __restrict__ scale, int64_t n_bases, float* __restrict__ x,
                     double* __restrict__ seg_mean, double* __restrict__ seg_std) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = j < n_bases;
    const int lane = threadIdx.x & 31;
    int r = 0;
    long long st = 0, en = 0, S = 0;
    const int16_t* sig = signal;
    if (active) {
        r = base_read[j];
        sig = signal + sig_off[r];
        S = sig_off[r + 1] - sig_off[r];
        st = starts[j];
        en = (j + 1 < base_off[r + 1]) ? (long long)starts[j + 1] : st + last_dur[r];
        // memory safety for reads flagged NRV_READ_BAD_EVENTS (their features are never used)
        if (st < 0) st = 0;
        if (en > S) en = S;
        if (en < st) en = st;
    }
    const long long n = en - st;
    double mean = 0.0, var = 0.0;
    const bool is_long = active && n > LONG_SEG;
    if (active && !is_long && n > 0) {
        long long s = 0;
        for (long long i = st; i < en; ++i) s += sig[i];
        mean = (double)s / (double)n;                    // np.mean: exact integer sum / n
        double q = 0.0;
        for (long long i = st; i < en; ++i) { double d = int_to_double(sig[i]) - mean; q += d * d; }
        var = q / (double)n;                             // np.std: sqrt(mean(|x - mean|^2)), ddof = 0
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
        o[0] = base_colour(bases[j]);
        o[1] = (float)(mean / sh);                       // NanoReviser.py:124
        o[2] = (float)(sd / sc);                         // NanoReviser.py:125
        const long long dur = (j + 1 < base_off[r + 1]) ? (long long)starts[j + 1] - starts[j] : (long long)last_dur[r];
        o[3] = (float)((double)dur / 10.0);              // nanorevtrainutils.py:164
        o[4] = ev_mean[j];
        o[5] = ev_std[j];
    }
}

int launch_base_features(const int16_t* signal, const int64_t* sig_off, const int32_t* starts,
                         const int64_t* base_off, const uint8_t* bases, const float* ev_mean,
                         const float* ev_std, const int32_t* last_dur, const int32_t* base_read,
                         const double* shift, const double* scale, int64_t n_bases,
                         float* x, double* seg_mean, double* seg_std, cudaStream_t st) {
    if (n_bases <= 0) return 0;
    base_features_kernel<<<(unsigned)((n_bases + 255) / 256), 256, 0, st>>>(
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
