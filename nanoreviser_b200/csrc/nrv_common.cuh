// Shared declarations of libnrv.so (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "nrv.h"

#define NRV_MAX_T 13        // largest supported window W (odd 5..13); shipped weights: 11
#define NRV_SIG 50
#define NRV_CNN_CH 8
#define NRV_SIGFEAT 64

namespace nrv {

// cudaFuncSetAttribute / cluster occupancy are PER DEVICE: a process may hold one nrv_handle per GPU, so "done once" state of
// a launcher is kept per device ordinal (the launchers run on the handle's device: every entry point calls cudaSetDevice).
constexpr int NRV_MAX_DEVICES = 64;
struct PerDevice {
    int v[NRV_MAX_DEVICES] = {0};
    int& cur() { int d = 0; cudaGetDevice(&d); return v[d & (NRV_MAX_DEVICES - 1)]; }
    // true exactly once per device
    bool first() { int& x = cur(); if (x) return false; x = 1; return true; }
};

// ---- packed device weights of one model ---------------------------------------------------
struct LstmLayerDev {
    int in_a, in_b, u, k, k_pad;       // K = in_a + in_b + u (x_t rows, then h rows); k_pad = roundup(K, 16)
    float* wcat[2];                    // [k_pad][4u], column = unit*4 + gate (i,f,c,o), rows [x | h]
    float* wrec[2];                    // [u][4u] recurrent rows only (used when the projection runs on tensor cores)
    float* bias[2];                    // [4u] same column order
    float* bn_scale;                   // [2u] gamma/sqrt(var+eps)  (identity for the last layer)
    float* bn_shift;                   // [2u] beta - mean*scale
    // tensor-core projection operands (layers 2, 3): B^T = [2*4u][k_in] fp16 (hi, lo), rows = dir-major interleaved
    // gate columns, the preceding BatchNormalization folded in; bias_tc [2*4u] = bias + bn_shift . Wk
    __half* pb_hi; __half* pb_lo;      // main part (K = in_a)
    __half* sb_hi; __half* sb_lo;      // per-base part (K = in_b), layer 2 only
    float* bias_tc;
    float* bias_l1 = nullptr;          // read_rnn11 only: the bias itself (the ping-pong kernel adds it in the epilogue; bias_tc is 0 there)
    // tensor-core recurrence operand: Wr^T [2 dirs][4u][u] fp16 (hi, lo), row = unit*4 + gate (layers 1..3)
    __half* rt_hi; __half* rt_lo;
    // fused layers with e4m3 correction passes (nrv_fused_pair.cu, F8): weights with a power-of-two scale, accumulator = 2^S z.
    //   f8_wk_hi / f8_wk_lo = fp16 pair of Wk 2^S [2*4u][in] (rows as pb_hi),  f8_wr_hi / f8_wr_lo = fp16 pair of Wr 2^S [2*4u][u]
    //   f8_wk8 / f8_wr8: the 8-bit copies for the two correction passes, [rows][2 in] / [rows][2 u] bytes, interleaved in groups
    //   of 4 inputs: bytes 8g..8g+3 = e4m3(W_hi 2^(S-12)) (meets x_lo 2^12), bytes 8g+4..8g+7 = e4m3(W_lo 2^S) (meets x_hi)
    __half* f8_wk_hi = nullptr; __half* f8_wk_lo = nullptr; uint8_t* f8_wk8 = nullptr;
    __half* f8_wr_hi = nullptr; __half* f8_wr_lo = nullptr; uint8_t* f8_wr8 = nullptr;
    float f8_acc_scale = 1.f;                                                // 2^-S
};

struct LstmIo {
    const float* act_in = nullptr;     // [n_win][T][in_a] fp32 (fused variants)
    const float* base_in = nullptr;    // [n_bases][in_b]
    const int32_t* win_base = nullptr;
    float* act_out = nullptr;          // fp32 [n_win][T][2u]
    const float* zin = nullptr;        // [2][T][n_win][4u]
    const float* zsig = nullptr;       // [n_bases][2][4u]
    __half* out_hi = nullptr;          // fp16 pair [n_win][T][out_ld] (columns [0, 2u))
    __half* out_lo = nullptr;
    int out_ld = 0;
    int64_t out_nwp = 0;               // != 0: time-major padded rows, row(t, w) = t*out_nwp + w
    // fused total_rnn1: CNN-feature columns of boundary-free window tiles come straight from the per-base table [sf_rows][64]
    // (fp16 hi / lo); tile_base[tile] = table row of the tile's first window at t = 0, or -1 (then a2's columns [128, 192) are used)
    const __half* sf_hi = nullptr; const __half* sf_lo = nullptr; int64_t sf_rows = 0; const int32_t* tile_base = nullptr;
    bool rec_f8 = false;               // fused total_rnn1: its recurrence runs with e4m3 correction passes
    bool out_f8 = false;               // fused total_rnn1 feeding an F8 total_rnn2: out_hi = fp16(h) and out_lo holds, per 4 units,
                                       // the 8 bytes {e4m3(h_lo 2^12) x 4, e4m3(h_hi) x 4} (same bytes per row as the fp16 lo part)
};

struct CnnDev {
    // conv weights with the batch-norms kept separate (order is conv -> relu -> BN)
    float* blob;                       // one allocation, offsets below (floats)
    // [0..24) conv1_k[3][8]  [24..32) conv1_b  [32..40) bn1_scale  [40..48) bn1_shift
    // [48..240) conv2_k[3][8][8]  [240..248) conv2_b  [248..256) bn2_scale  [256..264) bn2_shift
    float* dense_k;                    // [400][64]
    float* dense_b;                    // [64]
    __half* dt_hi;                     // Dense(400->64) kernel transposed, W^T [64][400] (K-major), fp16 hi / lo parts: the A operand of the
    __half* dt_lo;                     // tcgen05 dense stage (TMA-streamed)
};

struct HeadsDev {
    float *d1k, *d1b, *d2k, *d2b, *mk, *mb, *fk, *fb, *ok, *ob;
    __half *d1t_hi, *d1t_lo;           // dense1 kernel transposed [128 out][128 in] fp16 (hi, lo): tensor-core B operand
    __half *d2t_hi, *d2t_lo;           // dense2 kernel transposed [32 out][128 in]
    int n_class;
};

struct ModelDev {
    int window, n_class;
    CnnDev cnn;
    LstmLayerDev lstm[4];
    HeadsDev heads;
};

// ---- kernel launchers (each returns the number of kernels it launched) -----------------------
// nrv_segment.cu
int launch_read_stats(const int16_t* signal, const int64_t* sig_off, const int64_t* base_off,
                      const int32_t* starts, const int32_t* last_dur, int window, int64_t n_reads,
                      const int32_t* hist_slot, void* ghist, int64_t n_multi, int max_segs,
                      double* shift, double* scale, int32_t* status, cudaStream_t st, int32_t* base_read = nullptr,
                      int64_t n_bases = 0);   // base_read != null: also fills the base -> read map (launch_base_read_map not needed)
size_t read_stats_hist_bytes(int64_t n_multi);      // zeroed scratch for the reads that span several segments
int read_stats_segment();                            // samples per CTA
int launch_base_features(const int16_t* signal, const int64_t* sig_off, const int32_t* starts,
                         const int64_t* base_off, const uint8_t* bases, const float* ev_mean,
                         const float* ev_std, const int32_t* last_dur, const int32_t* base_read,
                         const double* shift, const double* scale, int64_t n_bases,
                         float* x, double* seg_mean, double* seg_std, cudaStream_t st);
int launch_sig_windows(const int16_t* signal, const int64_t* sig_off, const int32_t* starts,
                       const int64_t* base_off, const int32_t* base_read, const double* shift,
                       const double* scale, int64_t n_bases, float* sig_win, cudaStream_t st);
int launch_base_read_map(const int64_t* base_off, int64_t n_reads, int64_t n_bases, int32_t* base_read,
                         cudaStream_t st);
int launch_window_map(const int64_t* base_off, const int64_t* win_off, const int32_t* status,
                      int64_t n_reads, int window, int32_t* win_base, cudaStream_t st);

// nrv_cnn.cu: sig_feat[m][base][64] for both models.  If explicit_win != nullptr the windows are read
// from it ([n_bases][50] fp32) instead of being gathered from the raw signal.
int launch_cnn(const ModelDev* m1, const ModelDev* m2, const int16_t* signal, const int64_t* sig_off,
               const int32_t* starts, const int64_t* base_off, const int32_t* base_read,
               const double* shift, const double* scale, const float* explicit_win, int64_t n_bases,
               float* sig_feat1, float* sig_feat2, __half* const sf_hi[2], __half* const sf_lo[2], cudaStream_t st);

// nrv_lstm.cu: one Bi-LSTM layer over a chunk of windows.
int launch_lstm_layer(int layer, int variant, const LstmLayerDev& L, const LstmIo& io, int64_t n_win, int T,
                      cudaStream_t st);

// nrv_gemm.cu: C[M][N] = A[M][K] . B[N][K]^T (+bias) with split-fp16 operands on tcgen05 (see file header)
int launch_gemm_f16x3(const __half* a_hi, const __half* a_lo, const __half* b_hi, const __half* b_lo, int64_t M, int N, int K,
                      float* c, const float* bias, int mode, int T, int64_t nw, int n_per_dir, int relu, int num_sms,
                      cudaStream_t st, const __half* w2t_hi = nullptr, const __half* w2t_lo = nullptr, const float* b2 = nullptr,
                      float in_scale = 1.f);      // mode 2: the A operand carries this power-of-two scale (undone in the epilogue)
int launch_split_f16(const float* x, __half* hi, __half* lo, int64_t n, cudaStream_t st);

// nrv_rec_tc.cu: tcgen05 recurrence (u = 64) consuming the projection GEMM's zin
int launch_lstm_rec_tc64(const LstmLayerDev& L, const LstmIo& io, int64_t n_win, int T, cudaStream_t st);
// fused projection + recurrence for read_rnn11 (K_in = 32 padded to 64 with a constant-1 bias column)
int launch_lstm_fused_tc64(const LstmLayerDev& L, const __half* x_hi, const __half* x_lo, const LstmIo& io, int64_t nwp, int T,
                           cudaStream_t st);
int launch_read_rnn1(const LstmLayerDev& L, const LstmIo& io, int64_t n_win, int T, cudaStream_t st);
// nrv_fused_pair.cu: projection + recurrence of total_rnn2 (cluster of 2) / total_rnn1 (cluster of 4), drained accumulator, h in TMEM
int launch_lstm_fused_pair64(const LstmLayerDev& L, const __half* x_hi, const __half* x_lo, const LstmIo& io, int64_t nwp, int T, int num_sms,
                             cudaStream_t st, int f8 = 0);
int launch_lstm_fused_pair128(const LstmLayerDev& L, const __half* x_hi, const __half* x_lo, const LstmIo& io, int64_t nwp, int T, int num_sms,
                              cudaStream_t st);
int launch_lstm_rec_tc128(const LstmLayerDev& L, const LstmIo& io, int64_t n_win, int T, cudaStream_t st);
int launch_lstm_rec_tc128_pair(const LstmLayerDev& L, const LstmIo& io, int64_t n_win, int T, cudaStream_t st);

// nrv_heads.cu: dense heads + flatten + feature + final softmax + argmax.
// stage = 0: act_in is total_rnn2's output [..][128]; 1: act_in is relu(Dense(128)) [..][128];
// 2: act_in is relu(Dense(32)) [..][32] (both dense layers ran in the tensor-core GEMM and its epilogue)
int launch_heads(const HeadsDev& H, const float* act_in /*[n_win][T][128]*/, int64_t n_win, int T,
                 float* probs /*[n_win][n_class] or null*/, uint8_t* labels /*[n_win] or null*/, int stage,
                 int64_t in_nwp /* != 0: act_in rows are time-major, row(t, w) = t*in_nwp + w */, cudaStream_t st);

// nrv_decode.cu
int launch_decode(const int64_t* base_off, const int64_t* win_off,
                  const uint8_t* bases, const uint8_t* y1, const uint8_t* y2, const int32_t* status,
                  int64_t n_reads, int64_t n_bases, int64_t n_win, int window, unsigned epoch, int64_t* tile_tmp,
                  uint8_t* revised, int64_t revised_cap, int64_t* out_off, int* overflow_flag, cudaStream_t st,
                  const uint8_t* q1 = nullptr, const uint8_t* q2 = nullptr, const uint8_t* qual_in = nullptr,
                  uint8_t* revised_qual = nullptr);
int launch_window_phred(const float* probs, const uint8_t* labels, int n_class, int64_t n_win, uint8_t* q, cudaStream_t st);
int64_t decode_tile_count(int64_t n_bases);

}  // namespace nrv
