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
#include "nrv_common.cuh"

namespace nrv {

// ------------------------------------------------------------------------------------------
// exact median / MAD by two-level radix selection, one CTA per read
// ------------------------------------------------------------------------------------------
constexpr int STAT_THREADS = 256;
constexpr int STAT_BINS = 4096;

// Find, for ranks k1 <= k2, the bins holding them and the ranks inside those bins.
// hist[STAT_BINS] in shared memory; result in res[4] = {bin1, rank_in_bin1, bin2, rank_in_bin2}.
__device__ void select_bins(const unsigned* hist, unsigned* part, long long k1, long long k2, long long* res) {
    const int tid = threadIdx.x;
    constexpr int PER = STAT_BINS / STAT_THREADS;   // 16
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) s += hist[tid * PER + i];
    part[tid] = s;
    __syncthreads();
    // exclusive prefix of part[] (256 entries): simple two-level warp scan
    unsigned v = s;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    __shared__ unsigned warp_tot[STAT_THREADS / 32];
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    unsigned base = 0;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    long long excl = (long long)base + v - s;
    long long run = excl;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        unsigned c = hist[tid * PER + i];
        if (k1 >= run && k1 < run + c) { res[0] = tid * PER + i; res[1] = k1 - run; }
        if (k2 >= run && k2 < run + c) { res[2] = tid * PER + i; res[3] = k2 - run; }
        run += c;
    }
    __syncthreads();
}

__device__ __forceinline__ void select_small(const unsigned* h, int nb, long long k, int* out) {
    long long run = 0;
    for (int i = 0; i < nb; ++i) {
        if (k >= run && k < run + h[i]) { *out = i; return; }
        run += h[i];
    }
    *out = nb - 1;
}

__global__ void __launch_bounds__(STAT_THREADS)
read_stats_kernel(const int16_t* __restrict__ signal, const int64_t* __restrict__ sig_off,
                  const int64_t* __restrict__ base_off, const int32_t* __restrict__ starts,
                  const int32_t* __restrict__ last_dur, int window,
                  double* __restrict__ shift_out, double* __restrict__ scale_out, int32_t* __restrict__ status) {
    __shared__ unsigned hist[STAT_BINS];
    __shared__ unsigned part[STAT_THREADS];
    __shared__ unsigned small1[32], small2[32];
    __shared__ long long res[4];
    __shared__ int bad;
    const int r = blockIdx.x;
    const int tid = threadIdx.x;
    const int16_t* sig = signal + sig_off[r];
    const long long n = sig_off[r + 1] - sig_off[r];

    // ---- status checks (boundary error convention, include/nrv.h) -------------------------
    if (tid == 0) bad = 0;
    __syncthreads();
    const long long b0 = base_off[r], nb = base_off[r + 1] - b0;
    for (long long j = tid; j < nb; j += STAT_THREADS) {
        const long long st = starts[b0 + j];
        const long long en = (j + 1 < nb) ? (long long)starts[b0 + j + 1] : st + last_dur[r];
        if (st < 0 || en <= st || en > n) bad = 1;
    }
    __syncthreads();
    int st_code = NRV_READ_OK;
    if (bad || n <= 0 || nb <= 0) st_code = NRV_READ_BAD_EVENTS;
    else if (nb <= window) st_code = NRV_READ_TOO_SHORT;
    if (n <= 0) {
        if (tid == 0) { shift_out[r] = 0.0; scale_out[r] = 0.0; status[r] = st_code; }
        return;
    }
    const long long k1 = (n - 1) / 2, k2 = n / 2;   // the two middle order statistics (equal if n odd)

    // ---- pass 1: histogram of the high 12 bits of the order-preserving key -----------------
    for (int i = tid; i < STAT_BINS; i += STAT_THREADS) hist[i] = 0;
    __syncthreads();
    for (long long i = tid; i < n; i += STAT_THREADS) {
        unsigned key = (unsigned)((int)sig[i] + 32768);
        atomicAdd(&hist[key >> 4], 1u);
    }
    __syncthreads();
    select_bins(hist, part, k1, k2, res);
    const unsigned binA = (unsigned)res[0], binB = (unsigned)res[2];
    const long long rA = res[1], rB = res[3];
    // ---- pass 2: low 4 bits inside the selected bins ---------------------------------------
    if (tid < 32) { small1[tid] = 0; small2[tid] = 0; }
    __syncthreads();
    for (long long i = tid; i < n; i += STAT_THREADS) {
        unsigned key = (unsigned)((int)sig[i] + 32768);
        unsigned hi = key >> 4;
        if (hi == binA) atomicAdd(&small1[key & 15], 1u);
        if (hi == binB) atomicAdd(&small2[key & 15], 1u);
    }
    __syncthreads();
    __shared__ int lo1, lo2;
    if (tid == 0) { select_small(small1, 16, rA, &lo1); select_small(small2, 16, rB, &lo2); }
    __syncthreads();
    const int v1 = (int)((binA << 4) | lo1), v2 = (int)((binB << 4) | lo2);   // keys of the two middles
    const int shift2 = v1 + v2;                                               // 2 * median, key space
    // ---- pass 3/4: D = |2*key - shift2| in [0, 131070] -> 12 high bits, 5 low bits ---------
    for (int i = tid; i < STAT_BINS; i += STAT_THREADS) hist[i] = 0;
    __syncthreads();
    for (long long i = tid; i < n; i += STAT_THREADS) {
        int key = (int)sig[i] + 32768;
        unsigned D = (unsigned)abs(2 * key - shift2);
        atomicAdd(&hist[D >> 5], 1u);
    }
    __syncthreads();
    select_bins(hist, part, k1, k2, res);
    const unsigned binC = (unsigned)res[0], binD = (unsigned)res[2];
    const long long rC = res[1], rD = res[3];
    if (tid < 32) { small1[tid] = 0; small2[tid] = 0; }
    __syncthreads();
    for (long long i = tid; i < n; i += STAT_THREADS) {
        int key = (int)sig[i] + 32768;
        unsigned D = (unsigned)abs(2 * key - shift2);
        unsigned hi = D >> 5;
        if (hi == binC) atomicAdd(&small1[D & 31], 1u);
        if (hi == binD) atomicAdd(&small2[D & 31], 1u);
    }
    __syncthreads();
    if (tid == 0) {
        int l1, l2;
        select_small(small1, 32, rC, &l1);
        select_small(small2, 32, rD, &l2);
        const int D1 = (int)((binC << 5) | l1), D2 = (int)((binD << 5) | l2);
        const double shift = (double)shift2 * 0.5 - 32768.0;       // mean of the two middles, exact
        const double scale = (double)(D1 + D2) * 0.25;             // mean of two |x - shift|, exact
        shift_out[r] = shift;
        scale_out[r] = scale;
        if (st_code == NRV_READ_OK && !(scale > 0.0)) st_code = NRV_READ_SCALE_ZERO;
        status[r] = st_code;
    }
}

int launch_read_stats(const int16_t* signal, const int64_t* sig_off, const int64_t* base_off,
                      const int32_t* starts, const int32_t* last_dur, int window, int64_t n_reads,
                      double* shift, double* scale, int32_t* status, cudaStream_t st) {
    if (n_reads <= 0) return 0;
    read_stats_kernel<<<(unsigned)n_reads, STAT_THREADS, 0, st>>>(signal, sig_off, base_off, starts, last_dur,
                                                                   window, shift, scale, status);
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
        for (long long i = st; i < en; ++i) { double d = (double)sig[i] - mean; q += d * d; }
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
        for (long long i = sst + lane; i < sen; i += 32) { double d = (double)ssig[i] - m; q += d * d; }
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
