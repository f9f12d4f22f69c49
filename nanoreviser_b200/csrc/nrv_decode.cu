// K4: two-model merge + stream compaction (output_handeler.py:83,104-122 applied per read with the
// pass-through edges of composition D4, see include/nrv.h / DESIGN.md):
//   every base of every read emits 0..3 characters -> tile sums -> scan of tile sums -> scatter.
// Integer / byte work only; HBM-bound (reads 1 B base + 2 B labels, writes <= 2 B per base).
#include <math.h>

#include "nrv_common.cuh"

namespace nrv {

constexpr int DEC_THREADS = 256;
constexpr int DEC_PER = 4;
constexpr int DEC_TILE = DEC_THREADS * DEC_PER;

int64_t decode_tile_count(int64_t n_bases) { return (n_bases + DEC_TILE - 1) / DEC_TILE; }

// label_to_base = {5:'A', 4:'G', 3:'T', 2:'C', 1:'-', 0:'D'}   (output_handeler.py:83)
__device__ __forceinline__ uint8_t label_char(int l) {
    const uint32_t lo = ('D') | ('-' << 8) | ('C' << 16) | ('T' << 24);
    return (l < 4) ? (uint8_t)(lo >> (8 * l)) : (l == 4 ? 'G' : 'A');
}

struct Emit { int n; uint8_t c[3]; uint8_t q[3]; };

// Quality definition D6' (include/nrv.h nrv_result.revised_qual; oracle/nanorev_oracle.py get_qual_1): Phred of a window and
// model = #{k in 1..60 : 1 - p_argmax <= 10^(-k/10)} (fp32), i.e. floor(-10 log10(1 - p)) capped at 60, from a threshold table
constexpr int PHRED_MAX = 60;
constexpr int PHRED_PASS = 40;
struct PhredTable { float t[PHRED_MAX]; };      // kernel parameter (constant bank): no per-device symbol to initialise

__global__ void window_phred_kernel(const float* __restrict__ probs, const uint8_t* __restrict__ labels, int nc, int64_t n_win,
                                    uint8_t* __restrict__ q, const PhredTable tab) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_win) return;
    const float e = 1.0f - probs[w * nc + labels[w]];
    int cnt = 0;
#pragma unroll 4
    for (int k = 0; k < PHRED_MAX; ++k) cnt += (e <= tab.t[k]) ? 1 : 0;
    q[w] = (uint8_t)cnt;
}

__device__ __forceinline__ Emit emit_for_base(int64_t j, const int64_t* __restrict__ base_off,
                                              const int64_t* __restrict__ win_off,
                                              const int32_t* __restrict__ base_read,
                                              const uint8_t* __restrict__ bases, const uint8_t* __restrict__ y1,
                                              const uint8_t* __restrict__ y2, const int32_t* __restrict__ status,
                                              int window, int* read_out, int64_t* idx_in_read,
                                              const uint8_t* __restrict__ q1 = nullptr, const uint8_t* __restrict__ q2 = nullptr,
                                              const uint8_t* __restrict__ qual_in = nullptr, bool want_q = false) {
    Emit e; e.n = 0;
    // quality of a base that passes through: the basecaller's, when known (capped at Phred 93), else Phred 40
    const uint8_t qpass = want_q ? (qual_in ? (uint8_t)min((int)qual_in[j], 93) : (uint8_t)PHRED_PASS) : (uint8_t)0;
    const int r = base_read[j];
    const int64_t i = j - base_off[r];
    const int64_t N = base_off[r + 1] - base_off[r];
    const int64_t M = N - window;                       // number of windows (nanorevtrainutils.py:198)
    const int bef = (window - 1) / 2;                   // SET_BEF (nanorevtrainutils.py:210)
    *read_out = r; *idx_in_read = i;
    const uint8_t base = bases[j];
    const bool ok = (status == nullptr) || (status[r] == NRV_READ_OK);
    if (!ok || M <= 0 || i < bef || i >= bef + M) { e.q[e.n] = qpass; e.c[e.n++] = base; return e; }   // pass-through
    const int64_t w = win_off[r] + (i - bef);
    const uint8_t qa = want_q ? q1[w] : (uint8_t)0;
    const uint8_t qm = want_q ? (uint8_t)min((int)qa, (int)q2[w]) : (uint8_t)0;
    const int l1 = y1[w];                               // label space 0..5
    const int l2 = (int)y2[w] + 1;                      // class k of model2 == label k+1
    if (i == bef) {                                     // output_handeler.py:107: leading label_to_base[y_pre[0]]
        const uint8_t lead = label_char(l1);
        if (lead != '-') { e.q[e.n] = qa; e.c[e.n++] = lead; }
    }
    if (l1 == l2 && l1 >= 2) {                          // both models agree on a base
        e.q[e.n] = qm; e.c[e.n++] = label_char(l1);
    } else if (l1 == 0 && l2 >= 2) {                    // 'D': a base is missing after this one -> insert
        if (base != '-') { e.q[e.n] = qpass; e.c[e.n++] = base; }
        e.q[e.n] = qm; e.c[e.n++] = label_char(l2);
    } else if (l1 == 1 && l2 == 1) {                    // both say '-': this base is an insertion -> drop
    } else {
        if (base != '-') { e.q[e.n] = qpass; e.c[e.n++] = base; }
    }
    return e;
}

__global__ void __launch_bounds__(DEC_THREADS)
decode_count_kernel(const int64_t* __restrict__ base_off, const int64_t* __restrict__ win_off,
                    const int32_t* __restrict__ base_read, const uint8_t* __restrict__ bases,
                    const uint8_t* __restrict__ y1, const uint8_t* __restrict__ y2,
                    const int32_t* __restrict__ status, int64_t n_bases, int window,
                    int32_t* __restrict__ tile_sum) {
    __shared__ int warp_sum[DEC_THREADS / 32];
    const int64_t j0 = (int64_t)blockIdx.x * DEC_TILE + (int64_t)threadIdx.x * DEC_PER;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < DEC_PER; ++k) {
        const int64_t j = j0 + k;
        if (j < n_bases) {
            int r; int64_t i;
            cnt += emit_for_base(j, base_off, win_off, base_read, bases, y1, y2, status, window, &r, &i).n;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < DEC_THREADS / 32; ++w) s += warp_sum[w];
        tile_sum[blockIdx.x] = s;
    }
}

// single CTA: exclusive scan of tile sums -> tile_off (int64); total -> out_off[n_reads]
__global__ void __launch_bounds__(1024)
decode_scan_kernel(const int32_t* __restrict__ tile_sum, int64_t n_tiles, int64_t* __restrict__ tile_off,
                   int64_t* __restrict__ out_off, int64_t n_reads, int64_t revised_cap, int* __restrict__ overflow) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < n_tiles; base += 1024) {
        const int64_t i = base + threadIdx.x;
        long long v = (i < n_tiles) ? tile_sum[i] : 0;
        long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        long long wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += warp_tot[w];
        const long long c = carry;
        if (i < n_tiles) tile_off[i] = c + wbase + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + wbase + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out_off[n_reads] = carry;
        if (carry > revised_cap) *overflow = 1;
    }
}

__global__ void __launch_bounds__(DEC_THREADS)
decode_scatter_kernel(const int64_t* __restrict__ base_off, const int64_t* __restrict__ win_off,
                      const int32_t* __restrict__ base_read, const uint8_t* __restrict__ bases,
                      const uint8_t* __restrict__ y1, const uint8_t* __restrict__ y2,
                      const int32_t* __restrict__ status, int64_t n_bases, int window,
                      const int64_t* __restrict__ tile_off, uint8_t* __restrict__ revised, int64_t revised_cap,
                      int64_t* __restrict__ out_off, const uint8_t* __restrict__ q1, const uint8_t* __restrict__ q2,
                      const uint8_t* __restrict__ qual_in, uint8_t* __restrict__ revised_qual) {
    __shared__ int warp_sum[DEC_THREADS / 32];
    const int64_t j0 = (int64_t)blockIdx.x * DEC_TILE + (int64_t)threadIdx.x * DEC_PER;
    Emit em[DEC_PER];
    int rd[DEC_PER];
    int64_t ii[DEC_PER];
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < DEC_PER; ++k) {
        const int64_t j = j0 + k;
        em[k].n = 0; rd[k] = -1; ii[k] = -1;
        if (j < n_bases) {
            em[k] = emit_for_base(j, base_off, win_off, base_read, bases, y1, y2, status, window, &rd[k], &ii[k], q1, q2, qual_in,
                                  revised_qual != nullptr);
            cnt += em[k].n;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += warp_sum[w];
    int64_t pos = tile_off[blockIdx.x] + wbase + inc - cnt;
#pragma unroll
    for (int k = 0; k < DEC_PER; ++k) {
        if (rd[k] >= 0 && ii[k] == 0) out_off[rd[k]] = pos;      // first base of a read
        for (int c = 0; c < em[k].n; ++c) {
            if (pos < revised_cap) {
                revised[pos] = em[k].c[c];
                if (revised_qual) revised_qual[pos] = (uint8_t)(em[k].q[c] + 33);      // Phred+33
            }
            ++pos;
        }
    }
}

// reads with zero bases own no position: their offset is the next read's
__global__ void decode_fix_empty_kernel(const int64_t* __restrict__ base_off, int64_t n_reads,
                                        int64_t* __restrict__ out_off) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    if (base_off[r + 1] != base_off[r]) return;
    int64_t q = r + 1;
    while (q < n_reads && base_off[q + 1] == base_off[q]) ++q;
    out_off[r] = out_off[q];     // q == n_reads -> total
}

int launch_decode(const int64_t* base_off, const int64_t* win_off, const int32_t* base_read,
                  const uint8_t* bases, const uint8_t* y1, const uint8_t* y2, const int32_t* status,
                  int64_t n_reads, int64_t n_bases, int window, int32_t* counts_tmp, int64_t* tile_tmp,
                  uint8_t* revised, int64_t revised_cap, int64_t* out_off, int* overflow_flag, cudaStream_t st,
                  const uint8_t* q1, const uint8_t* q2, const uint8_t* qual_in, uint8_t* revised_qual) {
    const int64_t n_tiles = decode_tile_count(n_bases);
    if (revised_qual && (!q1 || !q2)) return -1;
    int n = 0;
    if (n_tiles > 0) {
        decode_count_kernel<<<(unsigned)n_tiles, DEC_THREADS, 0, st>>>(base_off, win_off, base_read, bases, y1, y2,
                                                                      status, n_bases, window, counts_tmp);
        ++n;
    }
    decode_scan_kernel<<<1, 1024, 0, st>>>(counts_tmp, n_tiles, tile_tmp, out_off, n_reads, revised_cap, overflow_flag);
    ++n;
    if (n_tiles > 0) {
        decode_scatter_kernel<<<(unsigned)n_tiles, DEC_THREADS, 0, st>>>(base_off, win_off, base_read, bases, y1, y2,
                                                                        status, n_bases, window, tile_tmp, revised,
                                                                        revised_cap, out_off, q1, q2, qual_in, revised_qual);
        ++n;
    }
    if (n_reads > 0) {
        decode_fix_empty_kernel<<<(unsigned)((n_reads + 255) / 256), 256, 0, st>>>(base_off, n_reads, out_off);
        ++n;
    }
    return n;
}

// Phred score of every window of one model from its softmax output and argmax labels (see window_phred_kernel)
int launch_window_phred(const float* probs, const uint8_t* labels, int n_class, int64_t n_win, uint8_t* q, cudaStream_t st) {
    if (n_win <= 0) return 0;
    PhredTable tab;
    for (int k = 1; k <= PHRED_MAX; ++k) tab.t[k - 1] = (float)pow(10.0, -k / 10.0);
    window_phred_kernel<<<(unsigned)((n_win + 255) / 256), 256, 0, st>>>(probs, labels, n_class, n_win, q, tab);
    return 1;
}

}  // namespace nrv
