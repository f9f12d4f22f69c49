// K4: two-model merge + stream compaction (output_handeler.py:83,104-122 applied per read with the
// pass-through edges of composition D4, see include/nrv.h / DESIGN.md):
//   every base of every read emits 0..3 characters -> ONE kernel: per-tile count, decoupled look-back scan over the tiles, scatter.
// Integer / byte work only; HBM-bound (reads 1 B base + 2 B labels, writes <= 2 B per base).
#include <math.h>

#include "nrv_common.cuh"

namespace nrv {

constexpr int DEC_THREADS = 256;
constexpr int DEC_PER = 8;
constexpr int DEC_TILE = DEC_THREADS * DEC_PER;

int64_t decode_tile_count(int64_t n_bases) { return (n_bases + DEC_TILE - 1) / DEC_TILE; }

// label_to_base = {5:'A', 4:'G', 3:'T', 2:'C', 1:'-', 0:'D'}   (output_handeler.py:83)
__device__ __forceinline__ uint8_t label_char(int l) {
    const uint32_t lo = ('D') | ('-' << 8) | ('C' << 16) | ('T' << 24);
    return (l < 4) ? (uint8_t)(lo >> (8 * l)) : (l == 4 ? 'G' : 'A');
}

struct Emit { int n; uint32_t c, q; };     // up to 3 characters / Phred scores, byte i = the i-th emitted symbol
__device__ __forceinline__ void emit_push(Emit& e, uint8_t ch, uint8_t q) { e.c |= (uint32_t)ch << (8 * e.n); e.q |= (uint32_t)q << (8 * e.n); ++e.n; }

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
    Emit e; e.n = 0; e.c = 0; e.q = 0;
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
    if (!ok || M <= 0 || i < bef || i >= bef + M) { emit_push(e, base, qpass); return e; }   // pass-through
    const int64_t w = win_off[r] + (i - bef);
    const uint8_t qa = want_q ? q1[w] : (uint8_t)0;
    const uint8_t qm = want_q ? (uint8_t)min((int)qa, (int)q2[w]) : (uint8_t)0;
    const int l1 = y1[w];                               // label space 0..5
    const int l2 = (int)y2[w] + 1;                      // class k of model2 == label k+1
    if (i == bef) {                                     // output_handeler.py:107: leading label_to_base[y_pre[0]]
        const uint8_t lead = label_char(l1);
        if (lead != '-') { emit_push(e, lead, qa); }
    }
    if (l1 == l2 && l1 >= 2) {                          // both models agree on a base
        emit_push(e, label_char(l1), qm);
    } else if (l1 == 0 && l2 >= 2) {                    // 'D': a base is missing after this one -> insert
        if (base != '-') { emit_push(e, base, qpass); }
        emit_push(e, label_char(l2), qm);
    } else if (l1 == 1 && l2 == 1) {                    // both say '-': this base is an insertion -> drop
    } else {
        if (base != '-') { emit_push(e, base, qpass); }
    }
    return e;
}

// ---- single pass: count + decoupled look-back scan + scatter in ONE kernel ---------------------------------------------------
// Tiles of DEC_TILE bases are handed out in launch order by a ticket counter (a tile only ever waits for tiles with smaller
// tickets, which are running or done).  tile_state[t] = {status:2 | epoch:22 | value:40}: status 1 = the tile's own count,
// 2 = inclusive prefix; a tile publishes its count, looks back (one warp, 32 predecessors per round) until it meets an inclusive
// prefix, publishes its own and scatters.  The epoch (a per-launch number) makes stale states of earlier batches invisible, so
// nothing is cleared between batches.  The last tile writes the total and resets the ticket counter.
constexpr unsigned long long DEC_VAL_MASK = (1ull << 40) - 1;
__device__ __forceinline__ unsigned long long dec_pack(unsigned status, unsigned epoch, unsigned long long v) {
    return ((unsigned long long)status << 62) | ((unsigned long long)(epoch & 0x3fffffu) << 40) | (v & DEC_VAL_MASK);
}

__global__ void __launch_bounds__(DEC_THREADS)
decode_fused_kernel(const int64_t* __restrict__ base_off, const int64_t* __restrict__ win_off,
                    const int32_t* __restrict__ base_read, const uint8_t* __restrict__ bases,
                    const uint8_t* __restrict__ y1, const uint8_t* __restrict__ y2,
                    const int32_t* __restrict__ status, int64_t n_reads, int64_t n_bases, int window,
                    unsigned long long* __restrict__ tile_state, unsigned long long* __restrict__ ticket, unsigned epoch,
                    int64_t n_tiles, uint8_t* __restrict__ revised, int64_t revised_cap,
                    int64_t* __restrict__ out_off, int* __restrict__ overflow, const uint8_t* __restrict__ q1,
                    const uint8_t* __restrict__ q2, const uint8_t* __restrict__ qual_in, uint8_t* __restrict__ revised_qual) {
    __shared__ int warp_sum[DEC_THREADS / 32];
    __shared__ long long s_tile, s_excl;
    if (threadIdx.x == 0) s_tile = (long long)atomicAdd(ticket, 1ull);
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t j0 = tile * DEC_TILE + (int64_t)threadIdx.x * DEC_PER;
    Emit em[DEC_PER];
    int rd[DEC_PER];
    int64_t ii[DEC_PER];
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < DEC_PER; ++k) {
        const int64_t j = j0 + k;
        em[k].n = 0; em[k].c = 0; em[k].q = 0; rd[k] = -1; ii[k] = -1;
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
    int wbase = 0, total = 0;
    for (int w = 0; w < DEC_THREADS / 32; ++w) { if (w < warp) wbase += warp_sum[w]; total += warp_sum[w]; }
    if (warp == 0) {
        // publish this tile's count, look back for the exclusive prefix, publish the inclusive prefix
        volatile unsigned long long* st = tile_state;
        if (lane == 0 && tile > 0) { st[tile] = dec_pack(1, epoch, (unsigned long long)total); }
        long long excl = 0;
        int64_t look = tile - 1;
        while (look >= 0) {                                  // warp-uniform
            const int64_t t = look - lane;
            unsigned long long v = 0;
            bool ready = t < 0;
            if (t >= 0) {
                v = st[t];
                ready = (v >> 62) != 0 && ((unsigned)(v >> 40) & 0x3fffffu) == (epoch & 0x3fffffu);
            }
            // the contiguous run of ready predecessors starting at `look`, cut after the first inclusive prefix
            const unsigned not_ready = __ballot_sync(0xffffffffu, !ready);
            const unsigned incl = __ballot_sync(0xffffffffu, ready && t >= 0 && (v >> 62) == 2);
            const int n_ok = not_ready ? __ffs(not_ready) - 1 : 32;
            const int first_incl = incl ? __ffs(incl) - 1 : 32;
            const int take = min(n_ok, first_incl + 1);      // lanes [0, take) are consumed
            long long part = (lane < take && t >= 0) ? (long long)(v & DEC_VAL_MASK) : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            excl += part;
            if (first_incl < take) break;                    // met an inclusive prefix: done
            look -= take;                                    // take == 0: spin on the same predecessor
        }
        if (lane == 0) {
            st[tile] = dec_pack(2, epoch, (unsigned long long)(excl + total));
            __threadfence();
            s_excl = excl;
        }
    }
    __syncthreads();
    int64_t pos = s_excl + wbase + inc - cnt;
#pragma unroll
    for (int k = 0; k < DEC_PER; ++k) {
        if (rd[k] >= 0 && ii[k] == 0) {                      // first base of a read; reads with zero bases before it own no position
            out_off[rd[k]] = pos;
            for (int64_t q = rd[k] - 1; q >= 0 && base_off[q + 1] == base_off[q]; --q) out_off[q] = pos;
        }
        for (int c = 0; c < em[k].n; ++c) {
            if (pos < revised_cap) {
                revised[pos] = (uint8_t)(em[k].c >> (8 * c));
                if (revised_qual) revised_qual[pos] = (uint8_t)((uint8_t)(em[k].q >> (8 * c)) + 33);      // Phred+33
            }
            ++pos;
        }
    }
    if (tile == n_tiles - 1 && threadIdx.x == 0) {
        const long long tot = s_excl + total;
        out_off[n_reads] = tot;
        for (int64_t q = n_reads - 1; q >= 0 && base_off[q + 1] == base_off[q]; --q) out_off[q] = tot;     // trailing empty reads
        if (tot > revised_cap) *overflow = 1;
        *ticket = 0;                                         // every ticket of this launch has been taken
    }
}

// no bases at all: every read is empty
__global__ void decode_all_empty_kernel(int64_t n_reads, int64_t* __restrict__ out_off) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r <= n_reads) out_off[r] = 0;
}

// tile_tmp: 1 + decode_tile_count(n_bases) 64-bit words, ZEROED when allocated: the ticket counter (at a FIXED place: batches of
// different sizes share the scratch), then the tile states;
// epoch: a number that differs from the one of every earlier launch on the same tile_tmp (mod 2^22, never 0)
int launch_decode(const int64_t* base_off, const int64_t* win_off, const int32_t* base_read,
                  const uint8_t* bases, const uint8_t* y1, const uint8_t* y2, const int32_t* status,
                  int64_t n_reads, int64_t n_bases, int window, unsigned epoch, int64_t* tile_tmp,
                  uint8_t* revised, int64_t revised_cap, int64_t* out_off, int* overflow_flag, cudaStream_t st,
                  const uint8_t* q1, const uint8_t* q2, const uint8_t* qual_in, uint8_t* revised_qual) {
    const int64_t n_tiles = decode_tile_count(n_bases);
    if (revised_qual && (!q1 || !q2)) return -1;
    if (n_tiles == 0) {
        decode_all_empty_kernel<<<(unsigned)((n_reads + 256) / 256), 256, 0, st>>>(n_reads, out_off);
        return 1;
    }
    unsigned long long* state = reinterpret_cast<unsigned long long*>(tile_tmp);
    decode_fused_kernel<<<(unsigned)n_tiles, DEC_THREADS, 0, st>>>(base_off, win_off, base_read, bases, y1, y2, status, n_reads, n_bases,
                                                                  window, state + 1, state, epoch, n_tiles, revised, revised_cap,
                                                                  out_off, overflow_flag, q1, q2, qual_in, revised_qual);
    return 1;
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
