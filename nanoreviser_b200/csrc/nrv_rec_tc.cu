// K3 recurrence on the 5th-generation tensor cores (tcgen05 + TMEM), the "persistent recurrence kernel that
// keeps the recurrent weights resident in shared memory" of BASELINE.json's north_star.
//
// The input projection x_t.Wk (+bias, BatchNormalization folded) has already been computed for all timesteps by
// the time-batched GEMM (nrv_gemm.cu).  What is left per step is  z = zin[t] + h_{t-1} . Wr ; gates ; c, h  --
// sequential in t, independent across windows.
//
// Data layout (all kernels of the tensor-core path):
//   * activations between layers: raw h as an fp16 (hi, lo) pair, TIME-MAJOR and padded:  row(t, w) = t*nwp + w,
//     nwp = windows of the chunk rounded up to 128.  A tile of 128 windows at one timestep is 128 consecutive rows,
//     so the recurrence publishes h with ONE bulk-tensor (TMA) store per 128x64 block straight from the swizzled
//     shared-memory tile that is also the A operand of the next MMA, and the GEMM reads it back with TMA.
//   * zin[dir][t][tile][col/4][128 rows][4]: "column-quad major inside a 128-row tile" -- the accumulator layout
//     of both kernels is one TMEM lane (= one thread) per row, so a warp touching one column quad reads/writes
//     32 x 16 B contiguous: coalesced 128-bit accesses without any staging.
//
// u = 64 variant (read_rnn11, total_rnn2; lstmmodel.py:46,51).  One CTA = one direction x TWO tiles of 128
// windows that ping-pong:
//   * Wr^T [256 gate columns][64] as an fp16 (hi, lo) pair stays in shared memory for the whole kernel
//     (64 KB, K-major, 128-byte swizzle -- the B operand of every MMA);
//   * h_{t-1} of each tile lives in shared memory as an fp16 (hi, lo) pair in the same swizzled K-major
//     layout (the A operand), written by that tile's epilogue warps;
//   * the accumulator of each tile is a 128-lane x 256-column fp32 block of TMEM;
//   * warp 0: per tile and step, one elected lane TMA-stores h_{t-1} to global, issues 12 tcgen05.mma
//     (4 K-steps x {lo*hi, hi*lo, hi*hi}) and commits to an mbarrier; warps 1-4 / 5-8 are the epilogue of
//     tile A / B: tcgen05.ld -> + zin -> hard_sigmoid / tanh -> c (registers) -> h -> shared memory.
// While the epilogue of tile A runs on the CUDA cores, the tensor core works on tile B, and vice versa.
#include <algorithm>

#include "nrv_cell.cuh"
#include "nrv_common.cuh"
#include "nrv_tc.cuh"

namespace nrv {

using namespace tc;

bool make_tmap_f16_k64(CUtensorMap* tm, const void* base, int64_t rows, int K, int box_rows);   // nrv_gemm.cu

#ifndef NRV_ZIN_PF_DIST
#define NRV_ZIN_PF_DIST 4     // u = 128 pair kernel: L2 prefetch distance (32-column blocks) ahead of the demand loads of zin;
#endif                        // measured on B200: 0 -> 46.4 ms, 2 -> 42.9, 4 -> 41.1, 7 -> 52.2 ms per step (far prefetch thrashes)
#ifndef NRV_REC_PASSES
#define NRV_REC_PASSES 3      // fp16 split passes of the recurrent product h.Wr: 3 = lo*hi + hi*lo + hi*hi (fp32-equivalent)
#endif
constexpr int RT_THREADS = 352;                 // 1 MMA warp + 2 x 4 epilogue warps + 2 zin producer warps
constexpr int RT_W_BYTES = 256 * 64 * 2;        // 32 KB: Wr^T hi (or lo)
constexpr int RT_H_BYTES = 128 * 64 * 2;        // 16 KB: h tile hi (or lo)
constexpr int RT_ZS = 3;                        // zin ring stages per tile
constexpr int RT_Z_BYTES = 8 * 128 * 16;        // 16 KB: one 32-column block of a zin tile (8 quads x 128 rows x 16 B)
constexpr size_t RT_SMEM = 2 * RT_W_BYTES + 4 * RT_H_BYTES + 2 * RT_ZS * RT_Z_BYTES + 1024 + 256;

__global__ void __launch_bounds__(RT_THREADS, 1)
lstm_rec_tc64_kernel(const __half* __restrict__ wr_hi, const __half* __restrict__ wr_lo, const float* __restrict__ zin,
                     const __grid_constant__ CUtensorMap tm_out_hi, const __grid_constant__ CUtensorMap tm_out_lo,
                     int64_t nwp, int T) {
    constexpr int U = 64, N = 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_whi = smem;
    uint8_t* s_wlo = smem + RT_W_BYTES;
    uint8_t* s_h = smem + 2 * RT_W_BYTES;               // [tile][hi|lo][16 KB]
    uint8_t* s_z = s_h + 4 * RT_H_BYTES;                // [tile][stage][16 KB] zin blocks landed by cp.async.bulk
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_z + 2 * RT_ZS * RT_Z_BYTES);
    uint64_t* h_ready = bars;                           // [2] count 4 (one arrive per epilogue warp)
    uint64_t* acc_ready = bars + 2;                     // [2] count 2 (tcgen05.commit + "h store has left smem")
    uint64_t* zfull = bars + 4;                         // [2][RT_ZS] count 1 + tx bytes
    uint64_t* zempty = bars + 4 + 2 * RT_ZS;            // [2][RT_ZS] count 4 (one arrive per epilogue warp of the tile)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 + 4 * RT_ZS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const int64_t ntw = nwp >> 7;
    const int64_t wt[2] = {(int64_t)blockIdx.x * 2, (int64_t)blockIdx.x * 2 + 1};
    const bool valid[2] = {wt[0] < ntw, wt[1] < ntw};

    if (threadIdx.x == 0) {
        mbar_init(&h_ready[0], 4); mbar_init(&h_ready[1], 4);
        mbar_init(&acc_ready[0], 2); mbar_init(&acc_ready[1], 2);
        for (int i = 0; i < 2 * RT_ZS; ++i) { mbar_init(&zfull[i], 1); mbar_init(&zempty[i], 4); }
        fence_mbar_init();
        tma_prefetch_desc(&tm_out_hi); tma_prefetch_desc(&tm_out_lo);
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    {   // resident recurrent weights: [256 rows][64 halves] -> K-major 128-byte-swizzled tiles
        const uint4* gh = reinterpret_cast<const uint4*>(wr_hi + (size_t)dir * N * U);
        const uint4* gl = reinterpret_cast<const uint4*>(wr_lo + (size_t)dir * N * U);
        for (int i = threadIdx.x; i < N * 8; i += RT_THREADS) {
            const uint32_t off = sw128_offset(i >> 3, i & 7);
            *reinterpret_cast<uint4*>(s_whi + off) = __ldg(gh + i);
            *reinterpret_cast<uint4*>(s_wlo + off) = __ldg(gl + i);
        }
    }
    fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the tensor core / TMA (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== h store + MMA issuer =====================
        constexpr uint32_t idesc = umma_idesc_f16_f32(128, N);
        const uint64_t b_hi = umma_desc_k_sw128(smem_u32(s_whi)), b_lo = umma_desc_k_sw128(smem_u32(s_wlo));
        for (int s = 1; s <= T; ++s) {
            const int t_prev = dir ? (T - s) : (s - 1);               // timestep whose h is in shared memory
            for (int X = 0; X < 2; ++X) {
                if (!valid[X]) continue;
                mbar_wait(&h_ready[X], (uint32_t)((s - 1) & 1));
                tc_fence_after();
                if (elect_one()) {
                    const uint8_t* hh = s_h + (X * 2 + 0) * RT_H_BYTES;
                    const uint8_t* hl = s_h + (X * 2 + 1) * RT_H_BYTES;
                    const int grow = (int)(t_prev * nwp + wt[X] * 128);
                    tma_store_2d(&tm_out_hi, hh, dir * U, grow);
                    tma_store_2d(&tm_out_lo, hl, dir * U, grow);
                    tma_store_commit();
                    if (s < T) {
                        const uint64_t a_hi = umma_desc_k_sw128(smem_u32(hh)), a_lo = umma_desc_k_sw128(smem_u32(hl));
                        const uint32_t d = tmem_base + (uint32_t)(X * N);
#pragma unroll
                        for (int k = 0; k < U / 16; ++k) {
                            const uint64_t adv = (uint64_t)(k * 2);        // +32 B along K (16-byte units)
#if NRV_REC_PASSES >= 3
                            umma_f16_ss(d, a_lo + adv, b_hi + adv, idesc, k != 0);
                            umma_f16_ss(d, a_hi + adv, b_lo + adv, idesc, 1);
#elif NRV_REC_PASSES == 2
                            umma_f16_ss(d, a_hi + adv, b_lo + adv, idesc, k != 0);
#endif
                            umma_f16_ss(d, a_hi + adv, b_hi + adv, idesc, (NRV_REC_PASSES > 1) || k != 0);
                        }
                        umma_commit(&acc_ready[X]);                        // arrival 1: MMAs retired
                        tma_store_wait_read();
                        mbar_arrive(&acc_ready[X]);                        // arrival 2: the store has read the h tile
                    }
                }
                __syncwarp();
            }
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    } else if (warp >= 9) {
        // ===================== zin producers: warp 9 -> tile 0, warp 10 -> tile 1 =====================
        // The whole zin stream of a tile (T steps x 8 blocks of 16 KB, each contiguous in the tiled layout) goes through a
        // 3-stage shared-memory ring with cp.async.bulk: bytes in flight are no longer limited by registers, which is what
        // bounded this kernel (ncu: ~64 KB in flight per SM against ~2 us of loaded HBM latency).
        const int X = warp - 9;
        if (valid[X] && elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int s = 0; s < T; ++s) {
                const int t = dir ? (T - 1 - s) : s;
                const uint8_t* src = reinterpret_cast<const uint8_t*>(zin + (((int64_t)dir * T + t) * ntw + wt[X]) * (N * 128));
                for (int cb = 0; cb < 8; ++cb) {
                    uint64_t* fb = &zfull[X * RT_ZS + stage];
                    mbar_wait(&zempty[X * RT_ZS + stage], phase ^ 1);
                    mbar_arrive_expect_tx(fb, RT_Z_BYTES);
                    bulk_load_1d(s_z + (X * RT_ZS + stage) * RT_Z_BYTES, src + (size_t)cb * RT_Z_BYTES, RT_Z_BYTES, fb);
                    if (++stage == RT_ZS) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue: warps 1-4 -> tile 0, warps 5-8 -> tile 1 =====================
        const int X = (warp - 1) >> 2;
        const int q = warp & 3;                          // TMEM lane quarter accessible to this warp
        const int row = q * 32 + lane;
        if (valid[X]) {
            const uint32_t hs_hi = smem_u32(s_h + (X * 2 + 0) * RT_H_BYTES);
            const uint32_t hs_lo = smem_u32(s_h + (X * 2 + 1) * RT_H_BYTES);
            float c[U];
#pragma unroll
            for (int j = 0; j < U; ++j) c[j] = 0.f;
            int stage = 0; uint32_t zphase = 0;
            for (int s = 0; s < T; ++s) {
                if (s > 0) {
                    mbar_wait(&acc_ready[X], (uint32_t)((s - 1) & 1));
                    tc_fence_after();
                }
#pragma unroll
                for (int cb = 0; cb < N / 32; ++cb) {
                    uint32_t v[32];
                    if (s > 0) tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(X * N + cb * 32), v);
                    // zin block from the ring: quad j of this row at j*2 KB + row*16 B (conflict-free 128-bit reads)
                    mbar_wait(&zfull[X * RT_ZS + stage], zphase);
                    const uint32_t zs = smem_u32(s_z + (X * RT_ZS + stage) * RT_Z_BYTES) + row * 16;
                    float4 z[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) z[j] = ld_shared_f4(zs + j * 2048);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&zempty[X * RT_ZS + stage]);
                    if (++stage == RT_ZS) { stage = 0; zphase ^= 1; }
                    if (s > 0) tmem_ld_wait();
                    uint4 phi, plo;
                    lstm_cell_block(v, s > 0, z, &c[cb * 8], phi, plo);
                    const uint32_t off = sw128_offset(row, cb);
                    st_shared_v4(hs_hi + off, phi);
                    st_shared_v4(hs_lo + off, plo);
                }
                tc_fence_before();           // our tcgen05.ld of this step precede the next MMA's writes
                fence_proxy_async_smem();    // our h writes are visible to the tensor core and to TMA
                __syncwarp();
                if (lane == 0) mbar_arrive(&h_ready[X]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int launch_lstm_rec_tc64(const LstmLayerDev& L, const LstmIo& io, int64_t nwp, int T, cudaStream_t st) {
    if (nwp <= 0) return 0;
    if (L.u != 64 || !L.rt_hi || !io.out_hi || (nwp & 127)) return -1;
    CUtensorMap tmh, tml;
    if (!make_tmap_f16_k64(&tmh, io.out_hi, (int64_t)T * nwp, io.out_ld, 128) ||
        !make_tmap_f16_k64(&tml, io.out_lo, (int64_t)T * nwp, io.out_ld, 128))
        return -2;
    static PerDevice attr;
    if (attr.first()) cudaFuncSetAttribute(lstm_rec_tc64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RT_SMEM);
    dim3 grid((unsigned)(((nwp >> 7) + 1) / 2), 2);
    lstm_rec_tc64_kernel<<<grid, RT_THREADS, RT_SMEM, st>>>(L.rt_hi, L.rt_lo, io.zin, tmh, tml, nwp, T);
    return 1;
}

// ============================================================================================================
// u = 64 FUSED variant (read_rnn11, lstmmodel.py:46): input projection AND recurrence in one kernel, no zin at all.
// The layer's input is only 32 wide (read_rnn1's BN'd output, zero-padded to K = 64, with column 32 fixed to 1.0 so
// that the bias rides along as a weight row), so Wk^T (64 KB as an fp16 pair) fits next to Wr^T (64 KB):
//   z_t = [x_t | 1 | 0..] . Wk  +  h_{t-1} . Wr          -- 12 + 12 tcgen05.mma per step into one TMEM accumulator
// One CTA = one direction x one tile of 128 windows.  x_t tiles ([128][64] hi + lo, 32 KB) arrive through a 2-stage TMA
// ring; the accumulator is double-buffered over timesteps so that the projection MMAs of step s+1 (independent of h)
// are issued right behind the recurrent MMAs of step s and run while the epilogue of step s is busy.  Compared with
// GEMM + recurrence this removes 2 x 22.5 KB of HBM traffic per window and a kernel launch.
// ============================================================================================================
constexpr int RF_EPI_WARPS = 16;                      // 4 per TMEM lane quarter, 16 units (2 blocks of 32 gate columns) each
constexpr int RF_THREADS = 64 + 32 * RF_EPI_WARPS;    // warp 0: store + MMA issue, warp 1: TMA producer, warps 2..: epilogue
constexpr int RF_W_BYTES = 256 * 64 * 2;              // 32 KB per weight tile (Wk hi, Wk lo, Wr hi, Wr lo)
constexpr int RF_H_BYTES = 128 * 64 * 2;              // 16 KB per h / x tile part
constexpr int RF_XS = 2;                              // x ring stages
constexpr size_t RF_SMEM = 4 * RF_W_BYTES + 2 * RF_H_BYTES + RF_XS * 2 * RF_H_BYTES + 1024 + 256;

__global__ void __launch_bounds__(RF_THREADS, 1)
lstm_fused_tc64_kernel(const __half* __restrict__ wk_hi, const __half* __restrict__ wk_lo, const __half* __restrict__ wr_hi,
                       const __half* __restrict__ wr_lo, const __grid_constant__ CUtensorMap tm_x_hi,
                       const __grid_constant__ CUtensorMap tm_x_lo, const __grid_constant__ CUtensorMap tm_out_hi,
                       const __grid_constant__ CUtensorMap tm_out_lo, int64_t nwp, int T) {
    constexpr int U = 64, N = 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w = smem;                                  // [Wk hi | Wk lo | Wr hi | Wr lo] x [256 rows][64]
    uint8_t* s_h = smem + 4 * RF_W_BYTES;                 // [hi | lo][128 rows][64]
    uint8_t* s_x = s_h + 2 * RF_H_BYTES;                  // [stage][hi | lo][128 rows][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_x + RF_XS * 2 * RF_H_BYTES);
    uint64_t* h_half = bars;                              // [2] count RF_EPI_WARPS: units 0-31 / 32-63 of h_t are in the shared-memory tile
    uint64_t* acc_ready = bars + 2;                       // count 2 (commit + "h store left smem"); step 0: see below
    uint64_t* xfull = bars + 3;                           // [RF_XS]
    uint64_t* xempty = bars + 3 + RF_XS;                  // [RF_XS] count 1 (commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 + 2 * RF_XS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;

    if (threadIdx.x == 0) {
        mbar_init(&h_half[0], RF_EPI_WARPS); mbar_init(&h_half[1], RF_EPI_WARPS);
        mbar_init(acc_ready, 2);
        for (int i = 0; i < RF_XS; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 1); }
        fence_mbar_init();
        tma_prefetch_desc(&tm_x_hi); tma_prefetch_desc(&tm_x_lo); tma_prefetch_desc(&tm_out_hi); tma_prefetch_desc(&tm_out_lo);
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    {
        const __half* srcs[4] = {wk_hi + (size_t)dir * N * U, wk_lo + (size_t)dir * N * U, wr_hi + (size_t)dir * N * U,
                                 wr_lo + (size_t)dir * N * U};
        for (int i = threadIdx.x; i < 4 * N * 8; i += RF_THREADS) {
            const int m = i / (N * 8), r = i - m * (N * 8);
            *reinterpret_cast<uint4*>(s_w + m * RF_W_BYTES + sw128_offset(r >> 3, r & 7)) = __ldg(reinterpret_cast<const uint4*>(srcs[m]) + r);
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Persistent over tiles: the weights are loaded once per CTA.  g = running step counter over all of this CTA's tiles;
    // every barrier completes exactly once per step, so all parities derive from g.
    const int64_t ntw = nwp >> 7;
    if (warp == 1) {
        // ===================== TMA producer: x_t tiles in step order =====================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t wtile = blockIdx.x; wtile < ntw; wtile += gridDim.x)
                for (int s = 0; s < T; ++s) {
                    const int t = dir ? (T - 1 - s) : s;
                    const int grow = (int)(t * nwp + wtile * 128);
                    mbar_wait(&xempty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&xfull[stage], 2 * RF_H_BYTES);
                    tma_load_2d(s_x + (stage * 2 + 0) * RF_H_BYTES, &tm_x_hi, &xfull[stage], 0, grow);
                    tma_load_2d(s_x + (stage * 2 + 1) * RF_H_BYTES, &tm_x_lo, &xfull[stage], 0, grow);
                    if (++stage == RF_XS) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp == 0) {
        // ===================== h store + MMA issuer =====================
        constexpr uint32_t idesc = umma_idesc_f16_f32(128, N);
        const uint32_t w_base = smem_u32(s_w);
        int xstage = 0; uint32_t xphase = 0;
        // projection of global step gs: acc[gs & 1] = x . Wk   (zero-initialises the accumulator)
        auto proj = [&](uint32_t gs) {
            mbar_wait(&xfull[xstage], xphase);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t xa = smem_u32(s_x + (xstage * 2) * RF_H_BYTES);
                const uint32_t d = tmem_base + (gs & 1) * N;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t a_hi = umma_desc_k_sw128(xa + k * 32), a_lo = umma_desc_k_sw128(xa + RF_H_BYTES + k * 32);
                    const uint64_t b_hi = umma_desc_k_sw128(w_base + 0 * RF_W_BYTES + k * 32), b_lo = umma_desc_k_sw128(w_base + 1 * RF_W_BYTES + k * 32);
                    umma_f16_ss(d, a_lo, b_hi, idesc, k != 0);
                    umma_f16_ss(d, a_hi, b_lo, idesc, 1);
                    umma_f16_ss(d, a_hi, b_hi, idesc, 1);
                }
                umma_commit(&xempty[xstage]);
            }
            __syncwarp();
            if (++xstage == RF_XS) { xstage = 0; xphase ^= 1; }
        };
        uint32_t g = 0;                                                    // global step of the current tile's step 0
        for (int64_t wtile = blockIdx.x; wtile < ntw; wtile += gridDim.x, g += (uint32_t)T) {
            // step 0 of this tile: projection only.  (The previous tile's last h store has been waited for below.)
            proj(g);
            if (elect_one()) { umma_commit(acc_ready); mbar_arrive(acc_ready); }
            __syncwarp();
            if (T > 1) proj(g + 1);
            for (int s = 1; s <= T; ++s) {
                const int t_prev = dir ? (T - s) : (s - 1);
                const uint32_t ha = smem_u32(s_h);
                const uint32_t d = tmem_base + ((g + (uint32_t)s) & 1) * N;
                // recurrent product in two halves of K: the epilogue delivers units 0-31 first (h_half[0]), so K-steps 0, 1 run on the
                // tensor core while it is still busy with units 32-63
                auto rec_half = [&](int hf) {
#pragma unroll
                    for (int k = 2 * hf; k < 2 * hf + 2; ++k) {
                        const uint64_t a_hi = umma_desc_k_sw128(ha + k * 32), a_lo = umma_desc_k_sw128(ha + RF_H_BYTES + k * 32);
                        const uint64_t b_hi = umma_desc_k_sw128(w_base + 2 * RF_W_BYTES + k * 32), b_lo = umma_desc_k_sw128(w_base + 3 * RF_W_BYTES + k * 32);
#if NRV_REC_PASSES >= 3
                        umma_f16_ss(d, a_lo, b_hi, idesc, 1);            // accumulate onto the projection of this step
#endif
#if NRV_REC_PASSES >= 2
                        umma_f16_ss(d, a_hi, b_lo, idesc, 1);
#endif
                        umma_f16_ss(d, a_hi, b_hi, idesc, 1);
                    }
                };
                mbar_wait(&h_half[0], (g + (uint32_t)s - 1) & 1);
                tc_fence_after();
                if (s < T && elect_one()) rec_half(0);
                __syncwarp();
                mbar_wait(&h_half[1], (g + (uint32_t)s - 1) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const int grow = (int)(t_prev * nwp + wtile * 128);
                    tma_store_2d(&tm_out_hi, s_h, dir * U, grow);
                    tma_store_2d(&tm_out_lo, s_h + RF_H_BYTES, dir * U, grow);
                    tma_store_commit();
                    if (s < T) {
                        rec_half(1);
                        umma_commit(acc_ready);
                    }
                }
                __syncwarp();
                if (s < T) {
                    if (s + 1 < T) proj(g + (uint32_t)s + 1);               // runs on the tensor core during epilogue(s)
                    if (elect_one()) { tma_store_wait_read(); mbar_arrive(acc_ready); }
                    __syncwarp();
                } else {
                    if (elect_one()) tma_store_wait_read();                // h tile may be overwritten by the next tile's step 0
                    __syncwarp();
                }
            }
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    } else {
        // ===================== epilogue: warps 2..17; lane quarter = warp % 4, column group = (warp - 2) / 4 =====================
        // 16 warps (4 per scheduler) instead of 8: the cell math is a chain of dependent MUFU / FMA ops, and with 2 warps per
        // scheduler it issued at ~0.5 IPC (ncu) -- the epilogue, not the tensor core, set the step time
        constexpr int NB = 8 / (RF_EPI_WARPS / 4);        // 32-column blocks per warp
        const int q = warp & 3;
        const int cg = (warp - 2) >> 2;                  // block cb: units cb*32 + cg*8 .. +8 (K-steps 2*cb, 2*cb + 1 of the recurrent product)
        static_assert(NB == 2, "unit blocks of 32: 16 epilogue warps");
        const int row = q * 32 + lane;
        const float4 zero4[8] = {};
        const uint32_t sh_base = smem_u32(s_h);
        uint32_t g = 0;
        for (int64_t wtile = blockIdx.x; wtile < ntw; wtile += gridDim.x) {
            float c[8 * NB];
#pragma unroll
            for (int j = 0; j < 8 * NB; ++j) c[j] = 0.f;
            for (int s = 0; s < T; ++s, ++g) {
                mbar_wait(acc_ready, g & 1);
                tc_fence_after();
#pragma unroll
                for (int cb = 0; cb < NB; ++cb) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (g & 1) * N + (uint32_t)((cb * 4 + cg) * 32), v);
                    tmem_ld_wait();
                    uint4 phi, plo;
                    lstm_cell_block(v, true, zero4, &c[cb * 8], phi, plo);
                    const uint32_t off = sw128_offset(row, cb * 4 + cg);
                    st_shared_v4(sh_base + off, phi);
                    st_shared_v4(sh_base + RF_H_BYTES + off, plo);
                    if (cb == NB - 1) tc_fence_before();      // all our tcgen05.ld of this step precede the next accumulator's MMAs
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&h_half[cb]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ============================================================================================================
// read_rnn11, PING-PONG variant (round 2): TWO window tiles in flight per CTA.  The kernel above is bound by the chain
// epilogue(s) -> last recurrent MMAs -> accumulator -> epilogue(s + 1) (tensor pipe 59 %): with a second, independent tile the
// MMAs of one tile (projection + recurrence of its next step: 18 instructions) run while the 16 epilogue warps work on the
// other, and the epilogue never waits.  Resources: two single-buffered accumulators = all 512 TMEM columns; two h tiles
// (2 x 32 KB); the x tiles shrink to the 32 real input columns ([128][32] fp16, 64-byte swizzle, tools/ts_probe/sw64_probe.cu)
// -- the constant-1 column that carried the bias is gone, the bias is added in the epilogue from shared memory -- which is what
// makes room: 128 KB weights + 64 KB h + 32 KB x ring.  Schedule (all three roles agree on it): tiles are taken in pairs
// (slot 0: 1st, 3rd, ... tile of this CTA; slot 1: 2nd, 4th, ...), steps alternate (slot 0, s), (slot 1, s), (slot 0, s + 1), ...
// ============================================================================================================
constexpr int PP_X_BYTES = 128 * 32 * 2;              // 8 KB per x tile part ([128 rows][32], SWIZZLE_64B)
constexpr size_t PP_SMEM = 4 * RF_W_BYTES + 4 * RF_H_BYTES + RF_XS * 2 * PP_X_BYTES + 1024 /*bias*/ + 256 /*barriers*/ + 1024 /*alignment*/;

// One 32-column block (8 units x gates i,f,c,o) of the cell with the bias read from shared memory unit by unit (keeps the register
// count of the two-tile epilogue at the 96 a 576-thread CTA allows): z = acc + b; c, h update; h -> fp16 (hi, lo) as 2 x 16 bytes.
__device__ __forceinline__ void cell_block_sbias(const uint32_t (&v)[32], uint32_t sbias, float* c8, uint4& phi, uint4& plo) {
    float hv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 b = ld_shared_f4(sbias + j * 16);
        const float ig = hsig(__uint_as_float(v[4 * j + 0]) + b.x), fg = hsig(__uint_as_float(v[4 * j + 1]) + b.y);
        const float gg = tanh_fast(__uint_as_float(v[4 * j + 2]) + b.z), og = hsig(__uint_as_float(v[4 * j + 3]) + b.w);
        const float cn = fmaf(fg, c8[j], ig * gg);
        c8[j] = cn;
        hv[j] = og * tanh_fast(cn);
    }
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const __half2 hi = __floats2half2_rn(hv[2 * p], hv[2 * p + 1]);
        const float2 hf = __half22float2(hi);
        const __half2 lo = __floats2half2_rn(hv[2 * p] - hf.x, hv[2 * p + 1] - hf.y);
        ph[p] = half2_bits(hi); pl[p] = half2_bits(lo);
    }
    phi = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    plo = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;                     // 8-row groups of 64-byte rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;                              // SWIZZLE_64B
    return d;
}

__global__ void __launch_bounds__(RF_THREADS, 1)
lstm_fused_tc64_pp_kernel(const __half* __restrict__ wk_hi, const __half* __restrict__ wk_lo, const __half* __restrict__ wr_hi,
                          const __half* __restrict__ wr_lo, const float* __restrict__ bias, const __grid_constant__ CUtensorMap tm_x_hi,
                          const __grid_constant__ CUtensorMap tm_x_lo, const __grid_constant__ CUtensorMap tm_out_hi,
                          const __grid_constant__ CUtensorMap tm_out_lo, int64_t nwp, int T) {
    constexpr int U = 64, N = 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w = smem;                                  // [Wk hi | Wk lo | Wr hi | Wr lo] x [256 rows][64] (of Wk only K < 32 is used)
    uint8_t* s_h = smem + 4 * RF_W_BYTES;                 // [slot][hi | lo][128 rows][64]
    uint8_t* s_x = s_h + 4 * RF_H_BYTES;                  // [stage][hi | lo][128 rows][32]
    float* s_bias = reinterpret_cast<float*>(s_x + RF_XS * 2 * PP_X_BYTES);     // [256] gate columns (unit * 4 + gate) of this direction
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 256);
    uint64_t* h_done = bars;                              // [2 slots] count RF_EPI_WARPS: h of the step is in the slot's tile, its accumulator is drained
    uint64_t* acc_ready = bars + 2;                       // [2 slots] count 2: MMAs complete (commit) + the store of the previous h has left the tile
    uint64_t* xfull = bars + 4;                           // [RF_XS]
    uint64_t* xempty = bars + 4 + RF_XS;                  // [RF_XS] count 1 (commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 + 2 * RF_XS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;

    if (threadIdx.x == 0) {
        for (int j = 0; j < 2; ++j) { mbar_init(&h_done[j], RF_EPI_WARPS); mbar_init(&acc_ready[j], 2); }
        for (int i = 0; i < RF_XS; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 1); }
        fence_mbar_init();
        tma_prefetch_desc(&tm_x_hi); tma_prefetch_desc(&tm_x_lo); tma_prefetch_desc(&tm_out_hi); tma_prefetch_desc(&tm_out_lo);
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    {
        const __half* srcs[4] = {wk_hi + (size_t)dir * N * U, wk_lo + (size_t)dir * N * U, wr_hi + (size_t)dir * N * U,
                                 wr_lo + (size_t)dir * N * U};
        for (int i = threadIdx.x; i < 4 * N * 8; i += RF_THREADS) {
            const int m = i / (N * 8), r = i - m * (N * 8);
            *reinterpret_cast<uint4*>(s_w + m * RF_W_BYTES + sw128_offset(r >> 3, r & 7)) = __ldg(reinterpret_cast<const uint4*>(srcs[m]) + r);
        }
        if (threadIdx.x < N) s_bias[threadIdx.x] = __ldg(bias + dir * N + threadIdx.x);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // this CTA's tiles: k-th tile = blockIdx.x + k * gridDim.x; slot j takes k = 2m + j
    const int64_t ntw = nwp >> 7;
    const int64_t n_mine = blockIdx.x < ntw ? (ntw - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int64_t n_slot[2] = {(n_mine + 1) / 2, n_mine / 2};
    auto tile_of = [&](int64_t m, int j) { return (int64_t)blockIdx.x + (2 * m + j) * (int64_t)gridDim.x; };

    if (warp == 1) {
        // ===================== TMA producer: x_t tiles in schedule order =====================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t m = 0; m < n_slot[0]; ++m)
                for (int s = 0; s < T; ++s)
                    for (int j = 0; j < 2; ++j) {
                        if (m >= n_slot[j]) continue;
                        const int t = dir ? (T - 1 - s) : s;
                        const int grow = (int)(t * nwp + tile_of(m, j) * 128);
                        mbar_wait(&xempty[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&xfull[stage], 2 * PP_X_BYTES);
                        tma_load_2d(s_x + (stage * 2 + 0) * PP_X_BYTES, &tm_x_hi, &xfull[stage], 0, grow);
                        tma_load_2d(s_x + (stage * 2 + 1) * PP_X_BYTES, &tm_x_lo, &xfull[stage], 0, grow);
                        if (++stage == RF_XS) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 0) {
        // ===================== h store + MMA issuer =====================
        constexpr uint32_t idesc = umma_idesc_f16_f32(128, N);
        const uint32_t w_base = smem_u32(s_w);
        int xstage = 0; uint32_t xphase = 0;
        int prev_row[2] = {-1, -1};                       // global row (t * nwp + tile * 128) of the h the slot's tile currently holds
        for (int64_t m = 0; m < n_slot[0]; ++m)
            for (int s = 0; s < T; ++s)
                for (int j = 0; j < 2; ++j) {
                    if (m >= n_slot[j]) continue;
                    const uint32_t q = (uint32_t)(m * T + s);                  // step counter of the slot: every barrier of the slot completes once per step
                    const uint32_t ha = smem_u32(s_h + j * 2 * RF_H_BYTES);
                    const uint32_t d = tmem_base + (uint32_t)j * N;
                    if (q > 0) {                                               // epilogue of the slot's previous step: h is in the tile, the accumulator drained
                        mbar_wait(&h_done[j], (q - 1) & 1);
                        tc_fence_after();
                    }
                    mbar_wait(&xfull[xstage], xphase);
                    tc_fence_after();
                    if (elect_one()) {
                        if (q > 0) {
                            tma_store_2d(&tm_out_hi, s_h + j * 2 * RF_H_BYTES, dir * U, prev_row[j]);
                            tma_store_2d(&tm_out_lo, s_h + j * 2 * RF_H_BYTES + RF_H_BYTES, dir * U, prev_row[j]);
                            tma_store_commit();
                        }
                        const uint32_t xa = smem_u32(s_x + (xstage * 2) * PP_X_BYTES);
#pragma unroll
                        for (int k = 0; k < 2; ++k) {                          // projection: K = 32
                            const uint64_t a_hi = umma_desc_k_sw64(xa + k * 32), a_lo = umma_desc_k_sw64(xa + PP_X_BYTES + k * 32);
                            const uint64_t b_hi = umma_desc_k_sw128(w_base + 0 * RF_W_BYTES + k * 32), b_lo = umma_desc_k_sw128(w_base + 1 * RF_W_BYTES + k * 32);
                            umma_f16_ss(d, a_lo, b_hi, idesc, k != 0);
                            umma_f16_ss(d, a_hi, b_lo, idesc, 1);
                            umma_f16_ss(d, a_hi, b_hi, idesc, 1);
                        }
                        umma_commit(&xempty[xstage]);
                        if (s > 0) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {                      // recurrence on h_{t-1} of this tile
                                const uint64_t a_hi = umma_desc_k_sw128(ha + k * 32), a_lo = umma_desc_k_sw128(ha + RF_H_BYTES + k * 32);
                                const uint64_t b_hi = umma_desc_k_sw128(w_base + 2 * RF_W_BYTES + k * 32), b_lo = umma_desc_k_sw128(w_base + 3 * RF_W_BYTES + k * 32);
                                umma_f16_ss(d, a_lo, b_hi, idesc, 1);
                                umma_f16_ss(d, a_hi, b_lo, idesc, 1);
                                umma_f16_ss(d, a_hi, b_hi, idesc, 1);
                            }
                        }
                        umma_commit(&acc_ready[j]);
                        tma_store_wait_read();                                 // the store of the previous h has read the tile ...
                        mbar_arrive(&acc_ready[j]);                            // ... and with the MMAs complete the epilogue may overwrite it
                    }
                    __syncwarp();
                    if (++xstage == RF_XS) { xstage = 0; xphase ^= 1; }
                    const int t = dir ? (T - 1 - s) : s;
                    prev_row[j] = (int)(t * nwp + tile_of(m, j) * 128);
                }
        // the last h of both slots
        for (int j = 0; j < 2; ++j) {
            if (n_slot[j] == 0) continue;
            const uint32_t q = (uint32_t)(n_slot[j] * T);
            mbar_wait(&h_done[j], (q - 1) & 1);
            if (elect_one()) {
                tma_store_2d(&tm_out_hi, s_h + j * 2 * RF_H_BYTES, dir * U, prev_row[j]);
                tma_store_2d(&tm_out_lo, s_h + j * 2 * RF_H_BYTES + RF_H_BYTES, dir * U, prev_row[j]);
                tma_store_commit();
            }
            __syncwarp();
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    } else {
        // ===================== epilogue: warps 2..17; lane quarter = warp % 4, column group = (warp - 2) / 4 =====================
        constexpr int NB = 2;                             // 32-column blocks per warp
        const int q4 = warp & 3;
        const int cg = (warp - 2) >> 2;                   // block cb: units cb*32 + cg*8 .. +8
        const int row = q4 * 32 + lane;
        float c[2][8 * NB];
        for (int64_t m = 0; m < n_slot[0]; ++m)
            for (int s = 0; s < T; ++s)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (m >= n_slot[j]) continue;
                    const uint32_t q = (uint32_t)(m * T + s);
                    if (s == 0) {
#pragma unroll
                        for (int i = 0; i < 8 * NB; ++i) c[j][i] = 0.f;
                    }
                    const uint32_t sh_base = smem_u32(s_h + j * 2 * RF_H_BYTES);
                    mbar_wait(&acc_ready[j], q & 1);
                    tc_fence_after();
#pragma unroll
                    for (int cb = 0; cb < NB; ++cb) {
                        uint32_t v[32];
                        tmem_ld_32x32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)j * N + (uint32_t)((cb * 4 + cg) * 32), v);
                        tmem_ld_wait();
                        uint4 phi, plo;
                        cell_block_sbias(v, smem_u32(s_bias + (cb * 4 + cg) * 32), &c[j][cb * 8], phi, plo);
                        const uint32_t off = sw128_offset(row, cb * 4 + cg);
                        st_shared_v4(sh_base + off, phi);
                        st_shared_v4(sh_base + RF_H_BYTES + off, plo);
                    }
                    tc_fence_before();                    // our tcgen05.ld of this step precede the slot's next MMAs
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&h_done[j]);
                }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

bool make_tmap_f16_k32_sw64(CUtensorMap* tm, const void* base, int64_t rows, int K, int box_rows);     // nrv_gemm.cu

int launch_lstm_fused_tc64(const LstmLayerDev& L, const __half* x_hi, const __half* x_lo, const LstmIo& io, int64_t nwp, int T,
                           cudaStream_t st) {
    if (nwp <= 0) return 0;
    if (L.u != 64 || !L.rt_hi || !L.pb_hi || !io.out_hi || (nwp & 127)) return -1;
    CUtensorMap txh, txl, tmh, tml;
    if (!make_tmap_f16_k64(&txh, x_hi, (int64_t)T * nwp, 64, 128) || !make_tmap_f16_k64(&txl, x_lo, (int64_t)T * nwp, 64, 128) ||
        !make_tmap_f16_k64(&tmh, io.out_hi, (int64_t)T * nwp, io.out_ld, 128) ||
        !make_tmap_f16_k64(&tml, io.out_lo, (int64_t)T * nwp, io.out_ld, 128))
        return -2;
    static const bool use_pp = !(getenv("NRV_RNN11") && !strcmp(getenv("NRV_RNN11"), "single"));
    if (use_pp && L.bias_l1) {
        CUtensorMap pxh, pxl;
        if (!make_tmap_f16_k32_sw64(&pxh, x_hi, (int64_t)T * nwp, 64, 128) || !make_tmap_f16_k32_sw64(&pxl, x_lo, (int64_t)T * nwp, 64, 128)) return -2;
        static PerDevice attr_pp;
        if (attr_pp.first()) cudaFuncSetAttribute(lstm_fused_tc64_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PP_SMEM);
        dim3 grid((unsigned)std::min<int64_t>(nwp >> 7, 74), 2);
        lstm_fused_tc64_pp_kernel<<<grid, RF_THREADS, PP_SMEM, st>>>(L.pb_hi, L.pb_lo, L.rt_hi, L.rt_lo, L.bias_l1, pxh, pxl, tmh, tml, nwp, T);
        return 1;
    }
    static PerDevice attr;
    if (attr.first()) cudaFuncSetAttribute(lstm_fused_tc64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RF_SMEM);
    dim3 grid((unsigned)std::min<int64_t>(nwp >> 7, 74), 2);       // persistent: 148 CTAs, weights loaded once each
    lstm_fused_tc64_kernel<<<grid, RF_THREADS, RF_SMEM, st>>>(L.pb_hi, L.pb_lo, L.rt_hi, L.rt_lo, txh, txl, tmh, tml, nwp, T);
    return 1;
}

// ============================================================================================================
// u = 128 variant (total_rnn1, lstmmodel.py:49): N = 512 gate columns, K = 128.
// Wr^T as an fp16 (hi, lo) pair is 2 x 128 KB -- more than one SM's shared memory -- so the hi half stays
// resident (128 KB, used by the lo*hi and hi*hi passes) and the lo half (used only by the hi*lo pass) is
// streamed from L2 every step through a 2-stage TMA ring of [128 columns][64 K] boxes (16 KB), issued by a
// dedicated producer warp that never waits on the recurrence (weights do not depend on h).
// One CTA = one direction x one tile of 128 windows; the accumulator fills all 512 TMEM columns, split in two
// 256-column halves H0 (units 0-63) / H1 (units 64-127) with their own "accumulator ready" barriers.  MMA order
// per step:  H1 x K-chunk 0  ->  H0 (all of K)  -> commit H0 ->  H1 x K-chunk 1 -> commit H1,  so that when H0's
// epilogue starts, nothing reads K-chunk 0 of h any more (it is the chunk H0's epilogue overwrites with its new
// h) and it overlaps the remaining H1 MMAs.  Eight epilogue warps: two per TMEM lane quarter, one per half.
// ============================================================================================================
constexpr int R2_THREADS = 320;                       // MMA warp, TMA warp, 8 epilogue warps
constexpr int R2_WHI_BYTES = 2 * 512 * 64 * 2;        // 128 KB: two K-chunks of [512 rows][64]
constexpr int R2_H_BYTES = 128 * 64 * 2;              // 16 KB: one K-chunk of h (hi or lo)
constexpr int R2_BOX_BYTES = 128 * 64 * 2;            // 16 KB: streamed W_lo box
constexpr int R2_STAGES = 2;
constexpr size_t R2_SMEM = (size_t)R2_WHI_BYTES + 4 * R2_H_BYTES + R2_STAGES * R2_BOX_BYTES + 1024 + 128;

// streamed W_lo boxes of one step, in consumption order: (n-quarter, K-chunk)
__device__ __forceinline__ int r2_box_nq(int p) { return (0x32110032 >> (4 * p)) & 0xF; }   // {2,3,0,0,1,1,2,3}
__device__ __forceinline__ int r2_box_kc(int p) { return (0xE8 >> p) & 1; }                 // {0,0,0,1,0,1,1,1}

__global__ void __launch_bounds__(R2_THREADS, 1)
lstm_rec_tc128_kernel(const __half* __restrict__ wr_hi, const __grid_constant__ CUtensorMap tm_wlo,
                      const float* __restrict__ zin, const __grid_constant__ CUtensorMap tm_out_hi,
                      const __grid_constant__ CUtensorMap tm_out_lo, int64_t nwp, int T) {
    constexpr int U = 128, N = 512;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_whi = smem;                                   // [kc][512 rows][64]
    uint8_t* s_h = smem + R2_WHI_BYTES;                      // [hi|lo][kc][128 rows][64]
    uint8_t* s_ring = s_h + 4 * R2_H_BYTES;                  // [stage][128 rows][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + R2_STAGES * R2_BOX_BYTES);
    uint64_t* h_ready = bars;                                // count 8
    uint64_t* acc_ready = bars + 1;                          // [2] count 2
    uint64_t* full = bars + 3;                               // [R2_STAGES]
    uint64_t* empty = bars + 3 + R2_STAGES;                  // [R2_STAGES]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 + 2 * R2_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const int64_t ntw = nwp >> 7;
    const int64_t wtile = blockIdx.x;

    if (threadIdx.x == 0) {
        mbar_init(h_ready, 8);
        mbar_init(&acc_ready[0], 2); mbar_init(&acc_ready[1], 2);
        for (int s = 0; s < R2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        fence_mbar_init();
        tma_prefetch_desc(&tm_wlo); tma_prefetch_desc(&tm_out_hi); tma_prefetch_desc(&tm_out_lo);
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    {
        const uint4* gh = reinterpret_cast<const uint4*>(wr_hi + (size_t)dir * N * U);   // [512][128] halves: 16 chunks per row
        for (int i = threadIdx.x; i < N * 16; i += R2_THREADS) {
            const int row = i >> 4, c16 = i & 15;
            *reinterpret_cast<uint4*>(s_whi + (c16 >> 3) * (512 * 128) + sw128_offset(row, c16 & 7)) = __ldg(gh + i);
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 1) {
        // ===================== TMA producer: W_lo boxes in consumption order ==========
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int s = 1; s < T; ++s)
                for (int p = 0; p < 8; ++p) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], R2_BOX_BYTES);
                    tma_load_2d(s_ring + stage * R2_BOX_BYTES, &tm_wlo, &full[stage], r2_box_kc(p) * 64, dir * N + r2_box_nq(p) * 128);
                    if (++stage == R2_STAGES) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp == 0) {
        // ===================== h store + MMA issuer =====================
        constexpr uint32_t idesc256 = umma_idesc_f16_f32(128, 256), idesc128 = umma_idesc_f16_f32(128, 128);
        int stage = 0; uint32_t phase = 0;
        const uint32_t a_base = smem_u32(s_h), w_base = smem_u32(s_whi);
        // resident passes (lo*hi, hi*hi) of column half hf over K-steps [k0, k1)
        auto resident = [&](int hf, int k0, int k1, bool zero_first) {
            const uint32_t d = tmem_base + (uint32_t)(hf * 256);
            for (int k = k0; k < k1; ++k) {
                const int kc = k >> 2, kk = k & 3;
                const uint64_t a_hi = umma_desc_k_sw128(a_base + (0 * 2 + kc) * R2_H_BYTES + kk * 32);
                const uint64_t a_lo = umma_desc_k_sw128(a_base + (1 * 2 + kc) * R2_H_BYTES + kk * 32);
                const uint64_t b_hi = umma_desc_k_sw128(w_base + kc * (512 * 128) + hf * (256 * 128) + kk * 32);
                umma_f16_ss(d, a_lo, b_hi, idesc256, (zero_first && k == k0) ? 0u : 1u);
                umma_f16_ss(d, a_hi, b_hi, idesc256, 1);
            }
        };
        for (int s = 1; s <= T; ++s) {
            const int t_prev = dir ? (T - s) : (s - 1);
            mbar_wait(h_ready, (uint32_t)((s - 1) & 1));
            tc_fence_after();
            if (elect_one()) {
                const int grow = (int)(t_prev * nwp + wtile * 128);
#pragma unroll
                for (int kc = 0; kc < 2; ++kc) {
                    tma_store_2d(&tm_out_hi, s_h + (0 * 2 + kc) * R2_H_BYTES, dir * U + kc * 64, grow);
                    tma_store_2d(&tm_out_lo, s_h + (1 * 2 + kc) * R2_H_BYTES, dir * U + kc * 64, grow);
                }
                tma_store_commit();
            }
            __syncwarp();
            if (s == T) break;
            for (int p = 0; p < 8; ++p) {
                // resident MMAs that precede streamed box p
                if (elect_one()) {
                    if (p == 0) resident(1, 0, 4, true);          // H1 x K-chunk 0
                    if (p == 2) resident(0, 0, 8, true);          // H0, all of K
                    if (p == 6) resident(1, 4, 8, false);         // H1 x K-chunk 1
                }
                __syncwarp();
                const int nq = r2_box_nq(p), kc = r2_box_kc(p);
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t bst = smem_u32(s_ring + stage * R2_BOX_BYTES);
                    const uint32_t d = tmem_base + (uint32_t)(nq * 128);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint64_t a_hi = umma_desc_k_sw128(a_base + (0 * 2 + kc) * R2_H_BYTES + kk * 32);
                        umma_f16_ss(d, a_hi, umma_desc_k_sw128(bst + kk * 32), idesc128, 1);
                    }
                    umma_commit(&empty[stage]);
                    if (p == 5) umma_commit(&acc_ready[0]);       // H0 complete (and every read of h K-chunk 0)
                    if (p == 7) umma_commit(&acc_ready[1]);       // everything complete
                }
                __syncwarp();
                if (++stage == R2_STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) {
                tma_store_wait_read();                            // the h_{s-1} stores have left shared memory
                mbar_arrive(&acc_ready[0]);
                mbar_arrive(&acc_ready[1]);
            }
            __syncwarp();
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    } else {
        // ===================== epilogue: warps 2..9; lane quarter = warp % 4, column half = (warp - 2) / 4 ===========
        const int q = warp & 3;
        const int hf = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        uint8_t* hs_hi = s_h + (0 * 2 + hf) * R2_H_BYTES;       // units hf*64.. -> K chunk hf of the h tile
        uint8_t* hs_lo = s_h + (1 * 2 + hf) * R2_H_BYTES;
        float c[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) c[j] = 0.f;
        auto ztile_of = [&](int s_) {
            const int t_ = dir ? (T - 1 - s_) : s_;
            return reinterpret_cast<const float4*>(zin + (((int64_t)dir * T + t_) * ntw + wtile) * (N * 128)) + (hf * 64) * 128 + row;
        };
        const float4* ztile = ztile_of(0);
        float4 z[8], zn[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] = __ldg(ztile + j * 128);
        for (int s = 0; s < T; ++s) {
            const float4* znext_tile = (s + 1 < T) ? ztile_of(s + 1) : ztile;
            if (s > 0) {
                mbar_wait(&acc_ready[hf], (uint32_t)((s - 1) & 1));
                tc_fence_after();
            }
#pragma unroll
            for (int cb = 0; cb < 8; ++cb) {
                if (cb + 1 < 8) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) zn[j] = __ldg(ztile + ((cb + 1) * 8 + j) * 128);
                } else if (s + 1 < T) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) zn[j] = __ldg(znext_tile + j * 128);
                }
                uint32_t v[32];
                if (s > 0) {
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * 256 + cb * 32), v);
                    tmem_ld_wait();
                }
                uint4 phi, plo;
                lstm_cell_block(v, s > 0, z, &c[cb * 8], phi, plo);
                const uint32_t off = sw128_offset(row, cb);
                *reinterpret_cast<uint4*>(hs_hi + off) = phi;
                *reinterpret_cast<uint4*>(hs_lo + off) = plo;
#pragma unroll
                for (int j = 0; j < 8; ++j) z[j] = zn[j];
            }
            ztile = znext_tile;
            tc_fence_before();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(h_ready);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int launch_lstm_rec_tc128(const LstmLayerDev& L, const LstmIo& io, int64_t nwp, int T, cudaStream_t st) {
    if (nwp <= 0) return 0;
    if (L.u != 128 || !L.rt_hi || !io.out_hi || (nwp & 127)) return -1;
    CUtensorMap tm, tmh, tml;
    if (!make_tmap_f16_k64(&tm, L.rt_lo, 2 * 512, 128, 128) ||
        !make_tmap_f16_k64(&tmh, io.out_hi, (int64_t)T * nwp, io.out_ld, 128) ||
        !make_tmap_f16_k64(&tml, io.out_lo, (int64_t)T * nwp, io.out_ld, 128))
        return -2;
    static PerDevice attr;
    if (attr.first()) cudaFuncSetAttribute(lstm_rec_tc128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R2_SMEM);
    dim3 grid((unsigned)(nwp >> 7), 2);
    lstm_rec_tc128_kernel<<<grid, R2_THREADS, R2_SMEM, st>>>(L.rt_hi, tm, io.zin, tmh, tml, nwp, T);
    return 1;
}

// ============================================================================================================
// u = 128, CTA-PAIR variant (tcgen05 cta_group::2, thread-block cluster of 2 on one TPC), persistent over tiles.
// One cluster = one direction x TWO tiles of 128 windows (one per CTA) at a time.  A single tcgen05.mma.cta_group::2
// covers M = 256 rows (both CTAs' windows) x N = 256 gate columns and reads the B operand half from each CTA's
// shared memory -- so each SM holds only HALF of Wr^T (hi and lo: 128 KB), all of it resident.  Every CTA keeps its
// own windows' h tile (A operand, 64 KB) and accumulator (its 128 TMEM lanes x 512 columns) and runs its own
// epilogue; no activation ever crosses the pair.
//
// The accumulator fills TMEM and the h tile fills the rest of shared memory, so two tiles cannot ping-pong.  Instead
// the step is software-pipelined over the two column halves H0 / H1 (units 0-63 / 64-127 = K-chunks 0 / 1 of h):
//     all 8 epilogue warps:  epi H0(s) ................ epi H1(s) ................ | epi H0(s+1) ...
//     tensor core         :                            K0->H0(s+1)                | K0->H1, K1->H0 (commit H0), K1->H1 (commit H1)
// i.e. as soon as the epilogue has produced K-chunk 0 of h_s (and thereby drained accumulator H0) the leader issues the
// quarter of step s+1 that needs nothing else, hidden under the H1 epilogue; after H1 only 2/4 of the MMAs stand between
// the epilogue and its next accumulator half, and the last quarter runs under epi H0(s+1).  (ncu before: epilogue warps
// 42 % of their time waiting for the accumulator, tensor pipe idle while they worked.)
//
// Barriers (all complete exactly once per step; parity from a running step counter):
//   h_half[hf]  (local, 8)    epilogue warps wrote K-chunk hf of h_s            -> store warp
//   h_pair[hf]  (leader, 16)  the same, from both CTAs                          -> MMA issuer
//   acc_ready[hf] (local, 2)  multicast tcgen05.commit + "the TMA store of the old K-chunk hf has left smem" -> epilogue
// ============================================================================================================
constexpr int RP_THREADS = 320;                        // warp 0: MMA issue (leader); warps 1..8: epilogue; warp 9: h stores
constexpr int RP_W_BYTES = 8 * 128 * 64 * 2;           // 128 KB: [hi|lo][half][kc][128 rows][64]
constexpr int RP_H_BYTES = 128 * 64 * 2;               // 16 KB: one K-chunk of h (hi or lo)
constexpr size_t RP_SMEM = (size_t)RP_W_BYTES + 6 * RP_H_BYTES + 1024 + 128;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(RP_THREADS, 1)
lstm_rec_tc128_pair_kernel(const __half* __restrict__ wr_hi, const __half* __restrict__ wr_lo, const float* __restrict__ zin,
                           const __grid_constant__ CUtensorMap tm_out_hi, const __grid_constant__ CUtensorMap tm_out_lo,
                           int64_t nwp, int T) {
    constexpr int U = 128, N = 512;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w = smem;                                     // [part][half][kc][128 rows][64]
    uint8_t* s_h = smem + RP_W_BYTES;                        // [part][kc][128 rows][64]; K-chunk 0 of odd steps: s_h0b[part]
    uint8_t* s_h0b = s_h + 4 * RP_H_BYTES;                   // K-chunk 0 is double-buffered over steps (see below)
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_h + 6 * RP_H_BYTES);
    uint64_t* h_half = bars;                                 // [2] count 8
    uint64_t* h_pair = bars + 2;                             // [2] count 16 (leader's copy is used)
    uint64_t* acc_ready = bars + 4;                          // [2] count 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int dir = blockIdx.y;
    const int64_t ntw = nwp >> 7;
    const int64_t n_pairs = (ntw + 1) >> 1;
    const int64_t cl0 = blockIdx.x >> 1, cl_stride = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&h_half[i], 8); mbar_init(&h_pair[i], 16); mbar_init(&acc_ready[i], 2); }
        fence_mbar_init();
        tma_prefetch_desc(&tm_out_hi); tma_prefetch_desc(&tm_out_lo);
    }
    if (warp == 0) tmem_alloc_pair(tmem_slot, 512);
    {   // this CTA's half of Wr^T: for column half hf, global rows hf*256 + rank*128 .. +128 (hi and lo)
        for (int i = threadIdx.x; i < 2 * 2 * 128 * 16; i += RP_THREADS) {
            const int c16 = i & 15, row = (i >> 4) & 127, hf = (i >> 11) & 1, part = i >> 12;
            const __half* src = (part ? wr_lo : wr_hi) + ((size_t)dir * N + hf * 256 + rank * 128 + row) * U;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + c16);
            *reinterpret_cast<uint4*>(s_w + (((part * 2 + hf) * 2 + (c16 >> 3)) * (128 * 128)) + sw128_offset(row, c16 & 7)) = v;
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();            // barriers initialised, weights in place and TMEM allocated in BOTH CTAs
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== MMA issue (leader CTA only) =====================
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc_f16_f32(256, 256);
            const uint32_t a_base = smem_u32(s_h), w_base = smem_u32(s_w);
            // acc[hf] (+)= h[K-chunk kc] . Wr[kc, hf]   (4 K-steps x 3 split passes); K-chunk 0 of h_{s-1} is in buffer (s-1) & 1
            auto mma_block = [&](int hf, int kc, int buf, bool zero_first) {
                const uint32_t d = tmem_base + (uint32_t)(hf * 256);
                const uint32_t a_hi0 = (kc == 0 && buf) ? smem_u32(s_h0b) : a_base + (0 * 2 + kc) * RP_H_BYTES;
                const uint32_t a_lo0 = (kc == 0 && buf) ? smem_u32(s_h0b) + RP_H_BYTES : a_base + (1 * 2 + kc) * RP_H_BYTES;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint64_t a_hi = umma_desc_k_sw128(a_hi0 + kk * 32);
                    const uint64_t a_lo = umma_desc_k_sw128(a_lo0 + kk * 32);
                    const uint64_t b_hi = umma_desc_k_sw128(w_base + ((0 * 2 + hf) * 2 + kc) * (128 * 128) + kk * 32);
                    const uint64_t b_lo = umma_desc_k_sw128(w_base + ((1 * 2 + hf) * 2 + kc) * (128 * 128) + kk * 32);
#if NRV_REC_PASSES >= 3
                    umma_f16_ss_pair(d, a_lo, b_hi, idesc, (zero_first && kk == 0) ? 0u : 1u);
                    umma_f16_ss_pair(d, a_hi, b_lo, idesc, 1);
                    umma_f16_ss_pair(d, a_hi, b_hi, idesc, 1);
#elif NRV_REC_PASSES == 2
                    umma_f16_ss_pair(d, a_hi, b_lo, idesc, (zero_first && kk == 0) ? 0u : 1u);
                    umma_f16_ss_pair(d, a_hi, b_hi, idesc, 1);
#else
                    umma_f16_ss_pair(d, a_hi, b_hi, idesc, (zero_first && kk == 0) ? 0u : 1u);
#endif
                }
            };
            uint32_t g = 0;                                       // phase counter of h_pair (one phase per step and half)
            for (int64_t tp = cl0; tp < n_pairs; tp += cl_stride)
                for (int s = 1; s <= T; ++s, ++g) {
                    const int buf = (s - 1) & 1;
                    mbar_wait(&h_pair[0], g & 1);                 // K-chunk 0 of h_{s-1} in both CTAs; accumulator H0 drained
                    tc_fence_after();
                    if (s < T && elect_one()) mma_block(0, 0, buf, true);
                    __syncwarp();
                    mbar_wait(&h_pair[1], g & 1);                 // K-chunk 1; accumulator H1 drained
                    tc_fence_after();
                    if (s < T && elect_one()) {
                        mma_block(0, 1, buf, false);
                        umma_commit_pair(&acc_ready[0]);          // H0 complete -> epi H0(s) starts (it writes the OTHER K-chunk-0 buffer)
                        mma_block(1, 0, buf, true);
                        mma_block(1, 1, buf, false);
                        umma_commit_pair(&acc_ready[1]);          // H1 complete, and every read of K-chunk 1 of h_{s-1}
                    }
                    __syncwarp();
                }
        }
    } else if (warp == 9) {
        // ===================== h stores (both CTAs): K-chunk hf of h_s as soon as the epilogue has written it =====================
        uint32_t g = 0;
        for (int64_t tp = cl0; tp < n_pairs; tp += cl_stride) {
            const int64_t wtile = min(tp * 2 + (int64_t)rank, ntw - 1);
            for (int s = 0; s < T; ++s, ++g) {
                const int t = dir ? (T - 1 - s) : s;
                const int grow = (int)(t * nwp + wtile * 128);
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    mbar_wait(&h_half[hf], g & 1);
                    if (elect_one()) {
                        const uint8_t* src_hi = (hf == 0 && (s & 1)) ? s_h0b : s_h + (0 * 2 + hf) * RP_H_BYTES;
                        const uint8_t* src_lo = (hf == 0 && (s & 1)) ? s_h0b + RP_H_BYTES : s_h + (1 * 2 + hf) * RP_H_BYTES;
                        tma_store_2d(&tm_out_hi, src_hi, dir * U + hf * 64, grow);
                        tma_store_2d(&tm_out_lo, src_lo, dir * U + hf * 64, grow);
                        tma_store_commit();
                        tma_store_wait_read();                    // the store has left shared memory: the chunk may be overwritten
                        mbar_arrive(&acc_ready[hf]);              // (the other arrival is the commit of step s+1's MMAs; at the last
                        if (s == T - 1) mbar_arrive(&acc_ready[hf]);   // step there are none, so this warp completes the phase alone)
                    }
                    __syncwarp();
                }
            }
        }
        if (elect_one()) tma_store_wait_all();
        __syncwarp();
    } else {
        // ===================== epilogue: warps 1..8; TMEM lane quarter = warp % 4; 4 of the 8 column blocks of each half ===========
        const int q = warp & 3;
        const int sub = (warp - 1) >> 2;
        const int row = q * 32 + lane;
        const int64_t zstep = (dir ? -1 : 1) * ntw * (int64_t)(N * 128 / 4);      // float4s between consecutive steps of a tile
        uint32_t g = 0;                                           // phase counter of acc_ready
        for (int64_t tp = cl0; tp < n_pairs; tp += cl_stride) {
            const int64_t wtile = min(tp * 2 + (int64_t)rank, ntw - 1);   // odd tile count: the last peer repeats the last tile
            float c[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) c[j] = 0.f;
            // this thread's zin quads of step s: block (hf, b) at zs + (hf*64 + b*8 + j) * 128, j = 0..7 (2 KB apart)
            const float4* zs = reinterpret_cast<const float4*>(zin + (((int64_t)dir * T + (dir ? T - 1 : 0)) * ntw + wtile) * (N * 128)) +
                               (sub * 32) * 128 + row;
            float4 z[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) z[j] = __ldg(zs + j * 128);
            for (int s = 0; s < T; ++s, zs += zstep) {
                const bool more = s + 1 < T;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    if (s > 0) {
                        mbar_wait(&acc_ready[hf], (g + (uint32_t)s - 1) & 1);
                        tc_fence_after();
                    } else if (tp != cl0) {
                        // first step of a later tile: the previous tile's last h store must have left shared memory
                        mbar_wait(&acc_ready[hf], (g - 1) & 1);
                    }
                    const uint32_t hs_hi = (hf == 0 && (s & 1)) ? smem_u32(s_h0b) : smem_u32(s_h + (0 * 2 + hf) * RP_H_BYTES);
                    const uint32_t hs_lo = (hf == 0 && (s & 1)) ? smem_u32(s_h0b) + RP_H_BYTES : smem_u32(s_h + (1 * 2 + hf) * RP_H_BYTES);
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
#if NRV_ZIN_PF_DIST > 0
                        {   // paced L2 prefetch one half (4 blocks) ahead of the demand loads
                            const float4* pt = (hf == 0) ? zs + (64 + b * 8) * 128 : zs + zstep + (b * 8) * 128;
                            if (hf == 0 || more) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) asm volatile("prefetch.global.L2 [%0];" ::"l"(pt + j * 128));
                            }
                        }
#endif
                        float4 zn[8];
                        if (b < 3) {             // next block of this half: in flight while this one is computed
#pragma unroll
                            for (int j = 0; j < 8; ++j) zn[j] = __ldg(zs + (hf * 64 + (b + 1) * 8 + j) * 128);
                        }
                        const int cb = sub * 4 + b;
                        uint32_t v[32];
                        if (s > 0) {
                            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * 256 + cb * 32), v);
                            tmem_ld_wait();
                        }
                        uint4 phi, plo;
                        lstm_cell_block(v, s > 0, z, &c[(hf * 4 + b) * 8], phi, plo);
                        const uint32_t off = sw128_offset(row, cb);
                        st_shared_v4(hs_hi + off, phi);
                        st_shared_v4(hs_lo + off, plo);
                        if (b < 3) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) z[j] = zn[j];
                        }
                    }
                    tc_fence_before();           // our tcgen05.ld of this half precede the MMAs that overwrite it
                    fence_proxy_async_smem();    // our h writes are visible to the tensor core (of the leader) and to TMA
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&h_half[hf]);
                        mbar_arrive_remote(&h_pair[hf], 0);
                    }
                    // first block of the next half: issued AFTER the proxy fence (its MEMBAR would wait for the loads) and, for
                    // hf == 1, in flight while this warp waits for the next accumulator
                    if (hf == 0) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) z[j] = __ldg(zs + (64 + j) * 128);
                    } else if (more) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) z[j] = __ldg(zs + zstep + j * 128);
                    }
                }
            }
            g += (uint32_t)T;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();            // the leader's MMAs read the peer's shared memory: nobody leaves early
    if (warp == 0) { tc_fence_after(); tmem_dealloc_pair(tmem_base, 512); }
}

int launch_lstm_rec_tc128_pair(const LstmLayerDev& L, const LstmIo& io, int64_t nwp, int T, cudaStream_t st) {
    if (nwp <= 0) return 0;
    if (L.u != 128 || !L.rt_hi || !io.out_hi || (nwp & 127)) return -1;
    CUtensorMap tmh, tml;
    if (!make_tmap_f16_k64(&tmh, io.out_hi, (int64_t)T * nwp, io.out_ld, 128) ||
        !make_tmap_f16_k64(&tml, io.out_lo, (int64_t)T * nwp, io.out_ld, 128))
        return -2;
    static PerDevice attr;
    if (attr.first()) cudaFuncSetAttribute(lstm_rec_tc128_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RP_SMEM);
    const int64_t n_pairs = ((nwp >> 7) + 1) / 2;
    dim3 grid((unsigned)(2 * std::min<int64_t>(n_pairs, 37)), 2);   // persistent: 74 clusters = 148 CTAs, weights loaded once each
    lstm_rec_tc128_pair_kernel<<<grid, RP_THREADS, RP_SMEM, st>>>(L.rt_hi, L.rt_lo, io.zin, tmh, tml, nwp, T);
    return 1;
}

}  // namespace nrv
