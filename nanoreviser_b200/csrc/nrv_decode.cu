// K4: two-model merge + stream compaction (output_handeler.py:83,104-122 applied per read with the
// pass-through edges of composition D4, see include/nrv.h / DESIGN.md):
//   every base of every read emits 0..3 characters -> ONE kernel: per-tile count, decoupled look-back scan over the tiles, scatter.
// Integer / byte work only; HBM-bound (reads 1 B base + 2 B labels, writes <= 2 B per base).
#include <math.h>

#include "nrv_common.cuh"

namespace nrv {

constexpr int DEC_THREADS = 256;
constexpr int DEC_PER = 16;
constexpr int DEC_TILE = DEC_THREADS * DEC_PER;
constexpr int DEC_LOOK = 4;              // groups of 32 predecessors read per look-back round

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

// what one base emits (output_handeler.py:104-122 with the pass-through edges), branch-free: `inwin` = the base has a window of
// its own (status OK, bef <= i < bef + M), `lead` = it is the first such base of its read (output_handeler.py:107: the leading
// label_to_base[y_pre[0]] unless it is '-').  Returns the characters in emission order (byte 0 first), *q their Phred scores,
// *n how many (0..3).
//   both models agree on a base (l1 == l2 >= 2)        -> that base
//   l1 == 0 ('D': a base is missing after this one)    -> the original base (unless '-'), then model 2's base
//   both say '-' (l1 == l2 == 1): an insertion          -> nothing
//   anything else, and every base without a window      -> the original base (a '-' is dropped only where a window exists)
__device__ __forceinline__ uint32_t emit_word(uint32_t base, bool inwin, bool lead, int l1, int l2, uint32_t qpass, uint32_t qa,
                                              uint32_t qm, uint32_t* q, uint32_t* n) {
    const bool agree = (l1 == l2) && (l1 >= 2);
    const bool ins = (l1 == 0) && (l2 >= 2);
    const bool drop = (l1 == 1) && (l2 == 1);
    const bool eb = !inwin || (!agree && !drop && base != (uint32_t)'-');
    const bool el = inwin && (agree || ins);
    const uint32_t lab = label_char(l2);
    uint32_t w = eb ? (base | (el ? lab << 8 : 0u)) : (el ? lab : 0u);
    uint32_t qq = eb ? (qpass | (el ? qm << 8 : 0u)) : (el ? qm : 0u);
    uint32_t cnt = (eb ? 1u : 0u) + (el ? 1u : 0u);
    if (lead && l1 != 1) { w = (w << 8) | label_char(l1); qq = (qq << 8) | qa; ++cnt; }
    *q = qq; *n = cnt;
    return w;
}

// largest r in [lo, n_reads) with base_off[r] <= j (the read that owns base j; empty reads own nothing)
__device__ __forceinline__ int64_t dec_read_of(const int64_t* __restrict__ base_off, int64_t n_reads, int64_t j, int64_t lo) {
    int64_t hi = n_reads;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (base_off[mid] <= j) lo = mid; else hi = mid;
    }
    return lo;
}
// 16 bytes starting at byte address p (any alignment) from five aligned 32-bit loads; the caller guarantees [p & ~3, +20) is readable
__device__ __forceinline__ void dec_load16(const uint8_t* p, uint32_t (&w)[4]) {
    const uint32_t* a = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
    const unsigned sh = ((unsigned)reinterpret_cast<uintptr_t>(p) & 3u) * 8u;
    uint32_t v[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) v[i] = __ldg(a + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = __funnelshift_r(v[i], v[i + 1], sh);
}
// the tile's `n` staged bytes -> dst (any alignment): bytes up to the first 16-byte boundary, 128-bit stores, bytes after the last one
__device__ __forceinline__ void dec_copy_out(uint8_t* __restrict__ dst, const uint8_t* src, int n) {
    const int tid = threadIdx.x;
    const int head = min(n, (int)((16u - ((unsigned)reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u));
    if (tid < head) dst[tid] = src[tid];
    const int nvec = (n - head) >> 4;
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(src) + (head >> 2);        // src is 16-byte aligned
    const unsigned sh = ((unsigned)head & 3u) * 8u;
    uint4* d4 = reinterpret_cast<uint4*>(dst + head);
    for (int c = tid; c < nvec; c += DEC_THREADS) {
        uint32_t v[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) v[i] = sw[4 * c + i];
        d4[c] = make_uint4(__funnelshift_r(v[0], v[1], sh), __funnelshift_r(v[1], v[2], sh), __funnelshift_r(v[2], v[3], sh),
                           __funnelshift_r(v[3], v[4], sh));
    }
    const int t0 = head + nvec * 16;
    if (tid < n - t0) dst[t0 + tid] = src[t0 + tid];
}

// ---- single pass: count + decoupled look-back scan + scatter in ONE kernel ---------------------------------------------------
// Tiles of DEC_TILE bases are handed out in launch order by a ticket counter (a tile only ever waits for tiles with smaller
// tickets, which are running or done).  tile_state[t] = {status:2 | epoch:22 | value:40}: status 1 = the tile's own count,
// 2 = inclusive prefix; a tile publishes its count, looks back (one warp, 32 predecessors per round) until it meets an inclusive
// prefix, publishes its own and scatters.  The epoch (a per-launch number) makes stale states of earlier batches invisible, so
// nothing is cleared between batches.  The last tile writes the total and resets the ticket counter.
// Memory side (round 2): a thread owns DEC_PER = 16 consecutive bases; their characters come in one 128-bit load and, when the
// 16 bases lie inside one read and all have windows (all but the threads at read edges), the labels of both models in five
// aligned 32-bit loads each (the window index advances with the base).  The owning read is found by one binary search per
// warp (no per-base read map is read).  The emitted symbols are staged in shared memory at their tile-local positions and
// leave as 128-bit stores.
constexpr unsigned long long DEC_VAL_MASK = (1ull << 40) - 1;
__device__ __forceinline__ unsigned long long dec_pack(unsigned status, unsigned epoch, unsigned long long v) {
    return ((unsigned long long)status << 62) | ((unsigned long long)(epoch & 0x3fffffu) << 40) | (v & DEC_VAL_MASK);
}

__global__ void __launch_bounds__(DEC_THREADS, 4)
decode_fused_kernel(const int64_t* __restrict__ base_off, const int64_t* __restrict__ win_off,
                    const uint8_t* __restrict__ bases,
                    const uint8_t* __restrict__ y1, const uint8_t* __restrict__ y2,
                    const int32_t* __restrict__ status, int64_t n_reads, int64_t n_bases, int64_t n_win, int window,
                    unsigned long long* __restrict__ tile_state, unsigned long long* __restrict__ ticket, unsigned epoch,
                    int64_t n_tiles, uint8_t* __restrict__ revised, int64_t revised_cap,
                    int64_t* __restrict__ out_off, int* __restrict__ overflow, const uint8_t* __restrict__ q1,
                    const uint8_t* __restrict__ q2, const uint8_t* __restrict__ qual_in, uint8_t* __restrict__ revised_qual) {
    __shared__ __align__(16) uint8_t s_c[DEC_TILE * 3 + 32];
    __shared__ __align__(16) uint8_t s_q[DEC_TILE * 3 + 32];
    __shared__ int warp_sum[DEC_THREADS / 32];
    __shared__ long long s_tile, s_excl;
    if (threadIdx.x == 0) s_tile = (long long)atomicAdd(ticket, 1ull);
    __syncthreads();
    const int64_t tile = s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool want_q = revised_qual != nullptr;
    const int bef = (window - 1) / 2;                   // SET_BEF (nanorevtrainutils.py:210)
    const int64_t j0 = tile * DEC_TILE + (int64_t)threadIdx.x * DEC_PER;
    const int nb = (int)max((int64_t)0, min((int64_t)DEC_PER, n_bases - j0));

    // the read of this thread's first base: one search per warp, a second one only where a read ends inside the warp's bases
    int64_t r = 0;
    {
        const int64_t jw = tile * DEC_TILE + (int64_t)warp * 32 * DEC_PER;
        int64_t rw = 0;
        if (lane == 0 && jw < n_bases) rw = dec_read_of(base_off, n_reads, jw, 0);
        rw = __shfl_sync(0xffffffffu, rw, 0);
        r = rw;
        if (nb > 0 && j0 >= base_off[r + 1]) r = dec_read_of(base_off, n_reads, j0, rw);
    }
    uint32_t ec[DEC_PER], eq[DEC_PER];                  // chars (3 bytes) | n << 24 | first-base-of-its-read << 26 ; Phred scores
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < DEC_PER; ++k) { ec[k] = 0; eq[k] = 0; }
    if (nb > 0) {
        // the read state in 32-bit numbers relative to this thread's first base: base j0 + k is base ib + k of its read, which has
        // `rem` bases from j0 on and Mi windows; the window of base j0 + k is yo + k
        int64_t b0 = base_off[r], e0 = base_off[r + 1];
        int rem = (int)min(e0 - j0, (int64_t)1 << 30);
        int ib = (int)(j0 - b0);
        int Mi = (int)max((e0 - b0) - window, (int64_t)0);        // number of windows of the read (nanorevtrainutils.py:198)
        int64_t yo = win_off[r] - b0 - bef + j0;
        bool okr = (status == nullptr) || (status[r] == NRV_READ_OK);
        uint32_t cw[4] = {0, 0, 0, 0}, l1w[4], l2w[4];
        const bool vec_c = nb == DEC_PER && ((reinterpret_cast<uintptr_t>(bases) & 15u) == 0);
        if (vec_c) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(bases + j0));
            cw[0] = v.x; cw[1] = v.y; cw[2] = v.z; cw[3] = v.w;
        }
        // all 16 bases inside this read, every one with a window, and the 20 bytes around their labels inside the label arrays
        const bool fast = vec_c && okr && rem >= DEC_PER && ib >= bef && ib + DEC_PER <= bef + Mi && yo + 20 <= n_win &&
                          ((reinterpret_cast<uintptr_t>(y1) | reinterpret_cast<uintptr_t>(y2)) & 3u) == 0;
        if (fast) {
            dec_load16(y1 + yo, l1w); dec_load16(y2 + yo, l2w);
#pragma unroll
            for (int k = 0; k < DEC_PER; ++k) {
                const uint32_t base = (cw[k >> 2] >> (8 * (k & 3))) & 0xffu;
                const int l1 = (int)((l1w[k >> 2] >> (8 * (k & 3))) & 0xffu);
                const int l2 = (int)((l2w[k >> 2] >> (8 * (k & 3))) & 0xffu) + 1;
                uint32_t qpass = 0, qa = 0, qm = 0;
                if (want_q) {
                    qpass = qual_in ? (uint32_t)min((int)qual_in[j0 + k], 93) : (uint32_t)PHRED_PASS;
                    qa = q1[yo + k]; qm = min(qa, (uint32_t)q2[yo + k]);
                }
                uint32_t n;
                ec[k] = emit_word(base, true, k == 0 && ib == bef, l1, l2, qpass, qa, qm, &eq[k], &n);
                ec[k] |= (n << 24);                     // (ib >= bef > 0 or bef == 0 and ... : the first base of a read is flagged below)
                if (k == 0 && ib == 0) ec[k] |= 1u << 26;
                cnt += (int)n;
            }
        } else {
#pragma unroll
            for (int k = 0; k < DEC_PER; ++k) {
                if (k < nb) {
                    if (k >= rem) {                     // next non-empty read
                        const int64_t j = j0 + k;
                        do { ++r; } while (base_off[r + 1] <= j);
                        b0 = base_off[r]; e0 = base_off[r + 1];
                        rem = (int)min(e0 - j0, (int64_t)1 << 30);
                        ib = (int)(j0 - b0);
                        Mi = (int)max((e0 - b0) - window, (int64_t)0);
                        yo = win_off[r] - b0 - bef + j0;
                        okr = (status == nullptr) || (status[r] == NRV_READ_OK);
                    }
                    const uint32_t base = vec_c ? ((cw[k >> 2] >> (8 * (k & 3))) & 0xffu) : (uint32_t)bases[j0 + k];
                    const int i = ib + k;
                    const bool inwin = okr && (unsigned)(i - bef) < (unsigned)Mi;
                    int l1 = 0, l2 = 0;
                    uint32_t qpass = 0, qa = 0, qm = 0;
                    if (inwin) {
                        l1 = y1[yo + k];                // label space 0..5
                        l2 = (int)y2[yo + k] + 1;       // class k of model2 == label k+1
                    }
                    if (want_q) {
                        // quality of a base that passes through: the basecaller's, when known (capped at Phred 93), else Phred 40
                        qpass = qual_in ? (uint32_t)min((int)qual_in[j0 + k], 93) : (uint32_t)PHRED_PASS;
                        if (inwin) { qa = q1[yo + k]; qm = min(qa, (uint32_t)q2[yo + k]); }
                    }
                    uint32_t n;
                    ec[k] = emit_word(base, inwin, inwin && i == bef, l1, l2, qpass, qa, qm, &eq[k], &n);
                    ec[k] |= (n << 24) | ((i == 0 ? 1u : 0u) << 26);
                    cnt += (int)n;
                }
            }
        }
    }
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
    // stage this thread's symbols at their tile-local positions (under the look-back of warp 0)
    int lpos = wbase + inc - cnt;
    const int lpos0 = lpos;
#pragma unroll
    for (int k = 0; k < DEC_PER; ++k) {
        const int n = (int)((ec[k] >> 24) & 3u);
        if (n > 0) s_c[lpos] = (uint8_t)ec[k];
        if (n > 1) s_c[lpos + 1] = (uint8_t)(ec[k] >> 8);
        if (n > 2) s_c[lpos + 2] = (uint8_t)(ec[k] >> 16);
        if (want_q) {                                   // Phred+33
            if (n > 0) s_q[lpos] = (uint8_t)(eq[k] + 33);
            if (n > 1) s_q[lpos + 1] = (uint8_t)((eq[k] >> 8) + 33);
            if (n > 2) s_q[lpos + 2] = (uint8_t)((eq[k] >> 16) + 33);
        }
        lpos += n;
    }
    if (warp == 0) {
        // publish this tile's count, look back for the exclusive prefix, publish the inclusive prefix
        volatile unsigned long long* st = tile_state;
        if (lane == 0 && tile > 0) { st[tile] = dec_pack(1, epoch, (unsigned long long)total); }
        long long excl = 0;
        int64_t look = tile - 1;
        // DEC_LOOK x 32 predecessors per round, all loads in flight together: the rounds of a look-back are dependent L2 round
        // trips, and with 32 per round the whole launch could not retire more than 32 tiles per round trip (measured: 18 tiles/us)
        while (look >= 0) {                                  // warp-uniform
            unsigned long long v[DEC_LOOK];
#pragma unroll
            for (int m = 0; m < DEC_LOOK; ++m) {
                const int64_t t = look - lane - 32 * m;
                v[m] = t >= 0 ? st[t] : 0ull;
            }
            bool done = false;
            int consumed = 0;
#pragma unroll
            for (int m = 0; m < DEC_LOOK; ++m) {
                if (done || consumed < 32 * m) continue;     // an earlier group ended the round (warp-uniform)
                const int64_t t = look - lane - 32 * m;
                const bool ready = t < 0 || ((v[m] >> 62) != 0 && ((unsigned)(v[m] >> 40) & 0x3fffffu) == (epoch & 0x3fffffu));
                // the contiguous run of ready predecessors, cut after the first inclusive prefix
                const unsigned not_ready = __ballot_sync(0xffffffffu, !ready);
                const unsigned incl = __ballot_sync(0xffffffffu, ready && t >= 0 && (v[m] >> 62) == 2);
                const int n_ok = not_ready ? __ffs(not_ready) - 1 : 32;
                const int first_incl = incl ? __ffs(incl) - 1 : 32;
                const int take = min(n_ok, first_incl + 1);  // lanes [0, take) are consumed
                long long part = (lane < take && t >= 0) ? (long long)(v[m] & DEC_VAL_MASK) : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                excl += part;
                consumed += take;
                if (first_incl < take) done = true;          // met an inclusive prefix
            }
            if (done) break;
            look -= consumed;                                // consumed == 0: spin on the same predecessor
        }
        if (lane == 0) {
            st[tile] = dec_pack(2, epoch, (unsigned long long)(excl + total));
            __threadfence();
            s_excl = excl;
        }
    }
    __syncthreads();
    const int64_t g0 = s_excl;
    {
        const int ncopy = (int)max((int64_t)0, min((int64_t)total, revised_cap - g0));
        dec_copy_out(revised + g0, s_c, ncopy);
        if (want_q) dec_copy_out(revised_qual + g0, s_q, ncopy);
    }
    // read offsets: the first base of a read knows where the read starts; reads with zero bases before it own no position
    {
        int lp = lpos0;
#pragma unroll
        for (int k = 0; k < DEC_PER; ++k) {
            if (ec[k] & (1u << 26)) {
                const int64_t j = j0 + k;
                const int64_t rr = dec_read_of(base_off, n_reads, j, 0);
                out_off[rr] = g0 + lp;
                for (int64_t q = rr - 1; q >= 0 && base_off[q + 1] == base_off[q]; --q) out_off[q] = g0 + lp;
            }
            lp += (int)((ec[k] >> 24) & 3u);
        }
    }
    if (tile == n_tiles - 1 && threadIdx.x == 0) {
        const long long tot = g0 + total;
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
int launch_decode(const int64_t* base_off, const int64_t* win_off,
                  const uint8_t* bases, const uint8_t* y1, const uint8_t* y2, const int32_t* status,
                  int64_t n_reads, int64_t n_bases, int64_t n_win, int window, unsigned epoch, int64_t* tile_tmp,
                  uint8_t* revised, int64_t revised_cap, int64_t* out_off, int* overflow_flag, cudaStream_t st,
                  const uint8_t* q1, const uint8_t* q2, const uint8_t* qual_in, uint8_t* revised_qual) {
    const int64_t n_tiles = decode_tile_count(n_bases);
    if (revised_qual && (!q1 || !q2)) return -1;
    if (n_tiles == 0) {
        decode_all_empty_kernel<<<(unsigned)((n_reads + 256) / 256), 256, 0, st>>>(n_reads, out_off);
        return 1;
    }
    unsigned long long* state = reinterpret_cast<unsigned long long*>(tile_tmp);
    decode_fused_kernel<<<(unsigned)n_tiles, DEC_THREADS, 0, st>>>(base_off, win_off, bases, y1, y2, status, n_reads, n_bases, n_win,
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
