// C-ABI of libnrv.so (include/nrv.h): handle, grow-only device arenas, weight re-packing, and the
// orchestration of K1..K4 on one CUDA stream.  No allocation on the hot path after warm-up.
#include <cuda_fp8.h>
#include <math.h>
#include <string.h>

#include <algorithm>

#include "nrv_common.cuh"
#include <nvtx3/nvToolsExt.h>    // header-only: ranges around the stages when NRV_NVTX=1 (ncu --nvtx, Nsight Systems)

using namespace nrv;

namespace {

std::string g_create_error;
constexpr int NRV_SLOTS = 2;

struct Arena {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

struct PinnedArena {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

enum { ST_STATS = 0, ST_FEAT, ST_CNN, ST_L0, ST_PROJ1, ST_REC1, ST_PROJ2, ST_REC2, ST_PROJ3, ST_REC3, ST_HEADS_GEMM,
       ST_HEADS, ST_DECODE, ST_COUNT };
const char* const kStageNames[ST_COUNT] = {"read_stats", "base_features", "cnn", "lstm0", "proj1", "rec1", "proj2", "rec2",
                                           "proj3", "rec3", "heads_gemm", "heads", "decode"};

}  // namespace

struct nrv_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;       // side stream: read_rnn1 of the NEXT (chunk, model) runs under the fused layers of the current one
    cudaEvent_t ev_a1_free = nullptr, ev_l0_done = nullptr, ev_hg_done = nullptr, ev_tail_done = nullptr;
    int overlap_l0 = 1;                   // NRV_OVERLAP=0 keeps everything on one stream
    std::string err;
    bool sticky = false;
    int window = 11;
    ModelDev m[2];
    std::vector<void*> weight_allocs;
    int64_t launches = 0;
    int64_t chunk_windows = 148 * 1024;   // 592 tile pairs: 37 rounds on the 16 clusters per direction of total_rnn1, 16 on the 37 of total_rnn2,
                                          // 8 on the 148 CTAs of read_rnn11 -- all exact (NRV_CHUNK_WINDOWS; 37,888 measured 4 % slower)
    // inputs / per-batch arenas
    // Per-batch I/O lives in a ring of NRV_SLOTS slots so that two batches can be in flight: the H2D copies of batch i+1 (copy
    // stream) run under the kernels of batch i (main stream) and the D2H of batch i under the kernels of batch i+1.  Everything
    // BETWEEN input and output (features, activations, labels) is shared: the kernels of consecutive batches are serialised on
    // the main stream anyway.
    struct IoSlot {
        Arena d_signal, d_starts, d_bases, d_evm, d_evs, d_lastdur, d_off, d_qual_in;      // inputs
        Arena d_status, d_revised, d_outoff, d_revq, d_flag;                               // outputs
        PinnedArena h_off, h_flag;
        cudaEvent_t ev_h2d = nullptr, ev_compute = nullptr;
        // the ticket that owns the slot (0 = free) and what nrv_wait needs to finish it
        int64_t ticket = 0;
        nrv_result res = {};
        int64_t n_reads = 0;
        bool want_q = false;
    };
    IoSlot slot[NRV_SLOTS];
    int64_t next_ticket = 1;
    int cur = 0;                          // slot of the batch being enqueued
    cudaStream_t copy_stream = nullptr;
    Arena d_shift, d_scale, d_base_read,
        d_win_base, d_x, d_sigfeat[2], d_act[4], d_probs[2], d_y[2], d_counts, d_tiles, d_wq[2],
        d_segmean, d_segstd, d_sigwin, d_sfh[2], d_sfl[2], d_a1[2], d_a2[2], d_a3[2], d_a4[2], d_zin, d_tile_base, d_ghist, d_ref, d_ref_y, d_ref_p;
    IoSlot& io() { return slot[cur]; }
    int in_flight() const { int n = 0; for (const IoSlot& s : slot) n += s.ticket != 0; return n; }
    int path = 1;           // 0: fp32 SIMT everywhere; 1: tcgen05 projections for total_rnn1/total_rnn2 (NRV_PATH)
    int num_sms = 148;
    int trnn2_fused = 1;    // total_rnn2: 1 = fused CTA-pair kernel (nrv_fused_pair.cu); 0 = GEMM + recurrence (NRV_TRNN2=split)
    int trnn1_fused = 1;    // total_rnn1: 1 = fused cluster-of-4 kernel (nrv_fused_pair.cu); 0 = GEMM + recurrence (NRV_TRNN1=split)
    int f8_rnn2 = 1;        // total_rnn2's correction passes in e4m3 (kind::f8f6f4); NRV_F8=0 keeps them fp16
    int f8_rnn1 = 1;        // total_rnn1's RECURRENT correction passes in e4m3 as well; NRV_F8=2 limits F8 to total_rnn2
    int refine = 1;              // F8 path: windows whose top-2 softmax margin is below refine_tau are evaluated again in fp16 x 3 (NRV_REFINE=0: off)
    float refine_tau = 1e-3f;    // NRV_REFINE_TAU
    int refine_cap = 4096;       // windows per model and batch that can be refined (NRV_REFINE_CAP)
    unsigned decode_epoch = 0;   // launch number of the single-pass decode (nrv_decode.cu), 1 .. 2^22 - 2
    int sig_table = 1;      // fused total_rnn1 reads the CNN features of boundary-free tiles straight from the per-base table; NRV_SIGTAB=0: gather all
    int rec128_pair = 1;    // u = 128 recurrence on CTA pairs (tcgen05 cta_group::2); NRV_REC128=single selects the 1-CTA kernel
    // stage timing: CUDA-event pairs recorded on the stream around every stage launch, never synchronised
    // on the hot path; folded into per-stage totals by nrv_get_stage_ms().
    bool timing = false;
    bool nvtx = false;                    // NRV_NVTX=1: an NVTX range per stage scope (named like nrv_stage_name)
    float stage_ms[ST_COUNT] = {0};
    int64_t stage_launches[ST_COUNT] = {0};
    struct EvPair { cudaEvent_t a, b; int stage; };
    std::vector<EvPair> ev_pool;
    size_t ev_used = 0;
    int timers_open = 0;                  // live StageTimer scopes (they nest): the pool is never folded under them
    void fold_events() {
        if (ev_used == 0 || timers_open > 0) return;
        cudaStreamSynchronize(stream);
        if (stream2) cudaStreamSynchronize(stream2);      // side-stream stages (read_rnn1, heads tail) record there
        for (size_t i = 0; i < ev_used; ++i) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ev_pool[i].a, ev_pool[i].b) == cudaSuccess) stage_ms[ev_pool[i].stage] += ms;
        }
        ev_used = 0;
    }
};

namespace {

int fail(nrv_handle* h, int code, const std::string& msg) {
    if (h) { h->err = msg; if (code == NRV_E_CUDA) h->sticky = true; }
    else g_create_error = msg;
    return code;
}

#define CU(h, call)                                                                                   \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return fail(h, NRV_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));          \
    } while (0)

template <class T>
T* upload(nrv_handle* h, const std::vector<T>& v, cudaError_t* err) {
    void* p = nullptr;
    *err = cudaMalloc(&p, std::max<size_t>(v.size() * sizeof(T), 16));
    if (*err != cudaSuccess) return nullptr;
    *err = cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    h->weight_allocs.push_back(p);
    return reinterpret_cast<T*>(p);
}

void bn_fold(const float* bn, int n, std::vector<float>& scale, std::vector<float>& shift) {
    // tf.nn.batch_normalization: inv = gamma * rsqrt(var + eps); y = x*inv + (beta - mean*inv); eps = 1e-3 (Keras)
    scale.resize(n); shift.resize(n);
    for (int i = 0; i < n; ++i) {
        const float gamma = bn[i], beta = bn[n + i], mean = bn[2 * n + i], var = bn[3 * n + i];
        const float inv = gamma / sqrtf(var + 1e-3f);
        scale[i] = inv;
        shift[i] = beta - mean * inv;
    }
}

int pack_model(nrv_handle* h, const nrv_model_weights* w, ModelDev* out) {
    static const int IN_A[4] = {0, 32, 128, 256}, IN_B[4] = {6, 0, 64, 0}, UU[4] = {16, 64, 128, 64};
    cudaError_t e = cudaSuccess;
    out->window = w->window;
    out->n_class = w->n_class;
    // ---- CNN ----
    std::vector<float> blob(264), s, t;
    memcpy(&blob[0], w->conv1_k, 24 * sizeof(float));
    memcpy(&blob[24], w->conv1_b, 8 * sizeof(float));
    bn_fold(w->bn1, 8, s, t);
    memcpy(&blob[32], s.data(), 32); memcpy(&blob[40], t.data(), 32);
    memcpy(&blob[48], w->conv2_k, 192 * sizeof(float));
    memcpy(&blob[240], w->conv2_b, 8 * sizeof(float));
    bn_fold(w->bn2, 8, s, t);
    memcpy(&blob[248], s.data(), 32); memcpy(&blob[256], t.data(), 32);
    out->cnn.blob = upload(h, blob, &e); if (e) goto cuda_fail;
    out->cnn.dense_k = upload(h, std::vector<float>(w->sig_dense_k, w->sig_dense_k + 400 * 64), &e); if (e) goto cuda_fail;
    out->cnn.dense_b = upload(h, std::vector<float>(w->sig_dense_b, w->sig_dense_b + 64), &e); if (e) goto cuda_fail;
    {   // Dense(400 -> 64) for the tcgen05 stage of nrv_cnn.cu: W^T [64 features][400] K-major, split fp16 (hi, lo); TMA streams it
        // chunk by chunk (6 x 64 columns with the 128-byte swizzle, then 16 columns with the 32-byte swizzle)
        std::vector<__half> th((size_t)64 * 400), tl(th.size());
        for (int n = 0; n < 64; ++n)
            for (int k = 0; k < 400; ++k) {
                const float wv = w->sig_dense_k[(size_t)k * 64 + n];
                const __half a = __float2half_rn(wv);
                th[(size_t)n * 400 + k] = a;
                tl[(size_t)n * 400 + k] = __float2half_rn(wv - __half2float(a));
            }
        out->cnn.dt_hi = upload(h, th, &e); if (e) goto cuda_fail;
        out->cnn.dt_lo = upload(h, tl, &e); if (e) goto cuda_fail;
    }
    // ---- LSTM layers ----
    for (int l = 0; l < 4; ++l) {
        LstmLayerDev& L = out->lstm[l];
        const int in = IN_A[l] + IN_B[l], u = UU[l];
        L.in_a = IN_A[l]; L.in_b = IN_B[l]; L.u = u; L.k = in + u; L.k_pad = (L.k + 15) / 16 * 16;
        for (int d = 0; d < 2; ++d) {
            const nrv_lstm_dir& src = w->lstm[l][d];
            std::vector<float> wc((size_t)L.k_pad * 4 * u, 0.f), bb(4 * u);
            for (int r = 0; r < in; ++r)
                for (int g = 0; g < 4; ++g)
                    for (int j = 0; j < u; ++j) wc[(size_t)r * 4 * u + j * 4 + g] = src.kernel[(size_t)r * 4 * u + g * u + j];
            for (int r = 0; r < u; ++r)
                for (int g = 0; g < 4; ++g)
                    for (int j = 0; j < u; ++j)
                        wc[(size_t)(in + r) * 4 * u + j * 4 + g] = src.recurrent[(size_t)r * 4 * u + g * u + j];
            for (int g = 0; g < 4; ++g)
                for (int j = 0; j < u; ++j) bb[j * 4 + g] = src.bias[g * u + j];
            L.wcat[d] = upload(h, wc, &e); if (e) goto cuda_fail;
            L.bias[d] = upload(h, bb, &e); if (e) goto cuda_fail;
            std::vector<float> wr(wc.begin() + (size_t)in * 4 * u, wc.begin() + (size_t)(in + u) * 4 * u);
            L.wrec[d] = upload(h, wr, &e); if (e) goto cuda_fail;
        }
        if (l < 3) {
            bn_fold(w->bn_rnn[l], 2 * u, s, t);
            L.bn_scale = upload(h, s, &e); if (e) goto cuda_fail;
            L.bn_shift = upload(h, t, &e); if (e) goto cuda_fail;
        } else {
            L.bn_scale = nullptr; L.bn_shift = nullptr;
        }
        L.pb_hi = L.pb_lo = L.sb_hi = L.sb_lo = nullptr; L.bias_tc = nullptr;
        L.rt_hi = L.rt_lo = nullptr;
        if (l >= 1) {
            std::vector<__half> rh((size_t)2 * 4 * u * u), rl(rh.size());
            for (int d = 0; d < 2; ++d)
                for (int g = 0; g < 4; ++g)
                    for (int j = 0; j < u; ++j)
                        for (int k = 0; k < u; ++k) {
                            const float wv = w->lstm[l][d].recurrent[(size_t)k * 4 * u + g * u + j];
                            const size_t idx = ((size_t)d * 4 * u + j * 4 + g) * u + k;
                            rh[idx] = __float2half_rn(wv);
                            rl[idx] = __float2half_rn(wv - __half2float(rh[idx]));
                        }
            L.rt_hi = upload(h, rh, &e); if (e) goto cuda_fail;
            L.rt_lo = upload(h, rl, &e); if (e) goto cuda_fail;
        }
        if (l >= 1) {
            // Tensor-core projection operand B^T [2*4u][kin] (K-major, kin = in rounded up to 64 with zero columns),
            // rows = dir*4u + unit*4 + gate.  The input of this layer is BN(h_prev) (+ the raw CNN features for layer
            // 2): y = h*s + t  =>  fold s into the rows of Wk that multiply h_prev and t . Wk into the bias, so the GEMM
            // consumes the raw h in [-1, 1].
            std::vector<float> ps, pt;
            bn_fold(w->bn_rnn[l - 1], IN_A[l], ps, pt);
            if (l == 1) {   // read_rnn1's BN is applied by the producing kernel (see nrv_lstm.cu OMODE 2): no fold
                std::fill(ps.begin(), ps.end(), 1.f);
                std::fill(pt.begin(), pt.end(), 0.f);
            }
            const int N2 = 2 * 4 * u;
            const int kin = (in + 63) / 64 * 64;
            std::vector<float> bt((size_t)N2 * kin, 0.f), bias_tc(N2), bias_l1(N2);
            for (int d = 0; d < 2; ++d) {
                const nrv_lstm_dir& src = w->lstm[l][d];
                for (int g = 0; g < 4; ++g)
                    for (int j = 0; j < u; ++j) {
                        const int n = d * 4 * u + j * 4 + g;
                        double bacc = src.bias[g * u + j];
                        for (int r = 0; r < in; ++r) {
                            const float wv = src.kernel[(size_t)r * 4 * u + g * u + j];
                            if (r < IN_A[l]) {
                                bt[(size_t)n * kin + r] = ps[r] * wv;
                                bacc += (double)pt[r] * (double)wv;
                            } else {
                                bt[(size_t)n * kin + r] = wv;
                            }
                        }
                        bias_tc[n] = (float)bacc;
                        if (l == 1) {   // fused kernel: the operand's column 32 is the constant 1.0, its weight row is the bias
                            bt[(size_t)n * kin + 32] = (float)bacc;
                            bias_tc[n] = 0.f;
                            bias_l1[n] = (float)bacc;      // the ping-pong kernel multiplies only the 32 real columns and adds this in its epilogue
                        }
                    }
            }
            std::vector<__half> bh(bt.size()), bl(bt.size());
            for (size_t i = 0; i < bt.size(); ++i) {
                bh[i] = __float2half_rn(bt[i]);
                bl[i] = __float2half_rn(bt[i] - __half2float(bh[i]));
            }
            L.pb_hi = upload(h, bh, &e); if (e) goto cuda_fail;
            L.pb_lo = upload(h, bl, &e); if (e) goto cuda_fail;
            L.bias_tc = upload(h, bias_tc, &e); if (e) goto cuda_fail;
            if (l == 1) { L.bias_l1 = upload(h, bias_l1, &e); if (e) goto cuda_fail; }
            if (l >= 2) {
                // fused layers with e4m3 correction passes (nrv_fused_pair.cu, F8): the ACTIVATIONS keep their ordinary fp16 hi part
                // (8-bit copies: x_lo 2^12 and x_hi), the WEIGHTS carry one power-of-two scale per layer so that every pass accumulates
                // 2^S z.  b = floor(log2(448 / max|W|)) puts the largest weight just below the e4m3 maximum; S = 7 + b:
                //   W16 (hi, lo) = fp16 pair of W 2^S                          (|W| 2^S <= 448 * 128 < 65504)
                //   W_hi8 = e4m3(W_hi 2^(S-12)) = e4m3(W_hi 2^(b-5))            meets x_lo8 = e4m3(x_lo 2^12)
                //   W_lo8 = e4m3(W_lo 2^S)       (|W_lo| <= 2^-11 |W| => <= 28)  meets x_hi8 = e4m3(x_hi)
                // The 8-bit copies of one row are interleaved in groups of 4 inputs (LstmLayerDev::f8_wk8) exactly like the 8-bit copies
                // of the activations, so both correction passes are ONE K-contiguous e4m3 product of twice the length.
                // total_rnn2 (l = 3) runs projection and recurrence this way, total_rnn1 (l = 2) its recurrence (fp16 x 3 projection on
                // the scaled fp16 pair).
                double wmax = 1e-30;
                for (float v : bt) wmax = std::max(wmax, (double)fabsf(v));
                for (int d = 0; d < 2; ++d)
                    for (size_t i = 0; i < (size_t)u * 4 * u; ++i) wmax = std::max(wmax, (double)fabsf(w->lstm[l][d].recurrent[i]));
                const int bexp = (int)floor(log2(448.0 / wmax));
                const int S = 7 + bexp;
                const double sq = ldexp(1.0, S), s_hi8 = ldexp(1.0, S - 12), s_lo8 = ldexp(1.0, S);
                auto e4m3 = [](double v) { return (uint8_t)__nv_cvt_float_to_fp8((float)v, __NV_SATFINITE, __NV_E4M3); };
                // row-major [rows][K] fp32 values (already in the kernel's row order) -> fp16 pair of W sq and the interleaved 8-bit row
                auto split8 = [&](const std::vector<float>& src, size_t rows, int K, std::vector<__half>& h16, std::vector<__half>& l16,
                                  std::vector<uint8_t>& q8) {
                    h16.resize(rows * K); l16.resize(rows * K); q8.resize(rows * 2 * K);
                    for (size_t r = 0; r < rows; ++r)
                        for (int k = 0; k < K; ++k) {
                            const double wv = (double)src[r * K + k];
                            const __half hs = __float2half_rn((float)(wv * sq));
                            const double his = (double)__half2float(hs);              // hi part, scaled
                            h16[r * K + k] = hs;
                            l16[r * K + k] = __float2half_rn((float)(wv * sq - his));
                            uint8_t* g = &q8[r * 2 * K + (size_t)(k >> 2) * 8 + (k & 3)];
                            g[0] = e4m3(his / sq * s_hi8);
                            g[4] = e4m3((wv - his / sq) * s_lo8);
                        }
                };
                std::vector<__half> fh, fl, rh, rl;
                std::vector<uint8_t> f8, r8;
                split8(bt, (size_t)2 * 4 * u, kin, fh, fl, f8);
                std::vector<float> rt((size_t)2 * 4 * u * u);
                for (int d = 0; d < 2; ++d)
                    for (int g = 0; g < 4; ++g)
                        for (int j = 0; j < u; ++j)
                            for (int k = 0; k < u; ++k)
                                rt[((size_t)d * 4 * u + j * 4 + g) * u + k] = w->lstm[l][d].recurrent[(size_t)k * 4 * u + g * u + j];
                split8(rt, (size_t)2 * 4 * u, u, rh, rl, r8);
                L.f8_wk_hi = upload(h, fh, &e); if (e) goto cuda_fail;
                L.f8_wk_lo = upload(h, fl, &e); if (e) goto cuda_fail;
                L.f8_wk8 = upload(h, f8, &e); if (e) goto cuda_fail;
                L.f8_wr_hi = upload(h, rh, &e); if (e) goto cuda_fail;
                L.f8_wr_lo = upload(h, rl, &e); if (e) goto cuda_fail;
                L.f8_wr8 = upload(h, r8, &e); if (e) goto cuda_fail;
                L.f8_acc_scale = (float)ldexp(1.0, -S);
            }
        }
    }
    // ---- heads ----
    {
        HeadsDev& H = out->heads;
        const int W = w->window, nc = w->n_class;
        H.n_class = nc;
        H.d1k = upload(h, std::vector<float>(w->dense1_k, w->dense1_k + 128 * 128), &e); if (e) goto cuda_fail;
        H.d1b = upload(h, std::vector<float>(w->dense1_b, w->dense1_b + 128), &e); if (e) goto cuda_fail;
        H.d2k = upload(h, std::vector<float>(w->dense2_k, w->dense2_k + 128 * 32), &e); if (e) goto cuda_fail;
        H.d2b = upload(h, std::vector<float>(w->dense2_b, w->dense2_b + 32), &e); if (e) goto cuda_fail;
        H.mk = upload(h, std::vector<float>(w->main_k, w->main_k + 32 * 6), &e); if (e) goto cuda_fail;
        H.mb = upload(h, std::vector<float>(w->main_b, w->main_b + 6), &e); if (e) goto cuda_fail;
        H.fk = upload(h, std::vector<float>(w->feat_k, w->feat_k + W * 6 * 16), &e); if (e) goto cuda_fail;
        H.fb = upload(h, std::vector<float>(w->feat_b, w->feat_b + 16), &e); if (e) goto cuda_fail;
        H.ok = upload(h, std::vector<float>(w->final_k, w->final_k + 16 * nc), &e); if (e) goto cuda_fail;
        H.ob = upload(h, std::vector<float>(w->final_b, w->final_b + nc), &e); if (e) goto cuda_fail;
        std::vector<__half> th(128 * 128), tl(128 * 128);
        for (int o = 0; o < 128; ++o)
            for (int i = 0; i < 128; ++i) {
                const float wv = w->dense1_k[i * 128 + o];
                th[o * 128 + i] = __float2half_rn(wv);
                tl[o * 128 + i] = __float2half_rn(wv - __half2float(th[o * 128 + i]));
            }
        H.d1t_hi = upload(h, th, &e); if (e) goto cuda_fail;
        H.d1t_lo = upload(h, tl, &e); if (e) goto cuda_fail;
        std::vector<__half> uh(32 * 128), ul(32 * 128);
        for (int o = 0; o < 32; ++o)
            for (int i = 0; i < 128; ++i) {
                const float wv = w->dense2_k[i * 32 + o];
                uh[o * 128 + i] = __float2half_rn(wv);
                ul[o * 128 + i] = __float2half_rn(wv - __half2float(uh[o * 128 + i]));
            }
        H.d2t_hi = upload(h, uh, &e); if (e) goto cuda_fail;
        H.d2t_lo = upload(h, ul, &e); if (e) goto cuda_fail;
    }
    return NRV_OK;
cuda_fail:
    return fail(h, NRV_E_CUDA, std::string("weight upload: ") + cudaGetErrorString(e));
}

// Scoped stage timer.  Scopes nest (ST_PROJ2 encloses launch_l0's ST_L0) and the pool is a std::vector that grows under
// them, so a timer keeps the INDEX of its event pair, never a pointer; a full pool is folded only when no timer is open.
struct StageTimer {
    nrv_handle* h; int stage; int64_t launches0; ptrdiff_t slot = -1; cudaStream_t st;
    StageTimer(nrv_handle* h_, int s, cudaStream_t st_ = nullptr) : h(h_), stage(s), launches0(h_->launches), st(st_ ? st_ : h_->stream) {
        if (h->nvtx) nvtxRangePushA(kStageNames[s]);
        if (!h->timing) return;
        if (h->ev_used == h->ev_pool.size()) {
            if (h->ev_pool.size() >= 8192 && h->timers_open == 0) h->fold_events();
            if (h->ev_used == h->ev_pool.size()) {
                nrv_handle::EvPair p; p.stage = 0;
                if (cudaEventCreate(&p.a) != cudaSuccess) return;
                if (cudaEventCreate(&p.b) != cudaSuccess) { cudaEventDestroy(p.a); return; }
                h->ev_pool.push_back(p);
            }
        }
        slot = (ptrdiff_t)h->ev_used++;
        h->ev_pool[slot].stage = stage;
        cudaEventRecord(h->ev_pool[slot].a, st);
        ++h->timers_open;
    }
    ~StageTimer() {
        if (h->nvtx) nvtxRangePop();
        h->stage_launches[stage] += h->launches - launches0;
        if (slot >= 0) { cudaEventRecord(h->ev_pool[slot].b, st); --h->timers_open; }
    }
};

__global__ void iota_mul_kernel(int32_t* out, int64_t n, int mul) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int32_t)(i * mul);
}

__global__ void set_half_column_kernel(__half* a, int64_t rows, int ld, int col, float v) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) a[r * ld + col] = __float2half_rn(v);
}

// tile_base[tile] = first base of the tile's 128 windows if their bases are consecutive (no read boundary inside the tile), else -1.
// For such a tile the rows (t, w) of total_rnn1's CNN-feature columns are 128 CONSECUTIVE rows of the per-base feature table
// starting at tile_base + t: the fused layer kernel loads them with one TMA box straight from the table (nrv_fused_pair.cu) and
// the gather below skips the tile.  win_base is increasing inside a batch (jumps of W at read boundaries), so "consecutive" is one
// subtraction.
__global__ void tile_base_kernel(const int32_t* __restrict__ win_base, int64_t n_win, int32_t* __restrict__ tile_base, int64_t n_tiles) {
    const int64_t tile = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= n_tiles) return;
    const int64_t w0 = tile * 128, w1 = min(w0 + 127, n_win - 1);
    int32_t tb = -1;
    if (w0 < n_win) {
        const int32_t b0 = win_base[w0];
        if ((int64_t)win_base[w1] - b0 == w1 - w0) tb = b0;
    }
    tile_base[tile] = tb;
}

// columns [128, 192) of total_rnn1's input row (t, w) = CNN features of base win_base[w] + t; 8 x 16 B per half array.
// One CTA per (tile of 128 windows, t); tile_base != nullptr: tiles with consecutive bases are skipped (read from the table instead)
__global__ void __launch_bounds__(256) gather_sig_kernel(const __half* __restrict__ sf_hi, const __half* __restrict__ sf_lo,
                                  const int32_t* __restrict__ win_base, int64_t n_win, int64_t nwp, int ld,
                                  const int32_t* __restrict__ tile_base, __half* __restrict__ a_hi, __half* __restrict__ a_lo) {
    const int64_t tile = blockIdx.x;
    const int t = blockIdx.y;
    if (tile_base && tile_base[tile] >= 0) return;
    for (int i = threadIdx.x; i < 128 * 8; i += 256) {
        const int64_t w = tile * 128 + (i >> 3);
        const int q = i & 7;
        if (w >= n_win) break;
        const int64_t b = (int64_t)win_base[w] + t;
        const int64_t row = (int64_t)t * nwp + w;
        reinterpret_cast<uint4*>(a_hi + row * ld + 128)[q] = __ldg(reinterpret_cast<const uint4*>(sf_hi + b * NRV_SIGFEAT) + q);
        reinterpret_cast<uint4*>(a_lo + row * ld + 128)[q] = __ldg(reinterpret_cast<const uint4*>(sf_lo + b * NRV_SIGFEAT) + q);
    }
}

// Both models over all windows, chunk by chunk.  x [n_bases][6]; sig_feat[m] [n_bases][64] (+ fp16 pairs).
// only_model >= 0: that model alone (the refinement pass of near-tie windows); f8_off: every product in three fp16 passes.
int run_models(nrv_handle* h, int64_t n_win, const int32_t* win_base, const float* x, float* const sig_feat[2],
               float* const probs[2], uint8_t* const labels[2], int64_t n_bases, int only_model = -1, bool f8_off = false) {
    const int T = h->window;
    const int64_t CH = std::min<int64_t>(h->chunk_windows, std::max<int64_t>(n_win, 1));
    const int64_t rows = ((CH + 127) / 128 * 128) * T;     // padded time-major rows of one chunk
    static const int widths[4] = {32, 128, 256, 128};
    if (h->path == 0) {
        for (int l = 0; l < 4; ++l) CU(h, h->d_act[l].ensure((size_t)rows * widths[l] * sizeof(float)));
    } else {
        CU(h, h->d_act[0].ensure((size_t)rows * 32 * sizeof(float)));
        CU(h, h->d_act[3].ensure((size_t)rows * 128 * sizeof(float)));
        for (int k = 0; k < 2; ++k) {
            const size_t before = h->d_a1[k].cap;
            CU(h, h->d_a1[k].ensure((size_t)rows * 64 * 2));
            // columns [32, 64) of read_rnn11's operand are zero padding that no kernel ever writes
            if (h->d_a1[k].cap != before) {
                CU(h, cudaMemsetAsync(h->d_a1[k].p, 0, h->d_a1[k].cap, h->stream));
                if (k == 0) {   // hi part: column 32 = 1.0 (bias column of read_rnn11's fused projection)
                    const int64_t nrows = (int64_t)(h->d_a1[k].cap / (64 * 2));
                    set_half_column_kernel<<<(unsigned)((nrows + 255) / 256), 256, 0, h->stream>>>(h->d_a1[k].as<__half>(), nrows, 64, 32, 1.0f);
                    h->launches += 1;
                }
            }
            CU(h, h->d_a4[k].ensure((size_t)rows * 128 * 2));
        }
        CU(h, h->d_a2[0].ensure((size_t)rows * 192 * 2)); CU(h, h->d_a2[1].ensure((size_t)rows * 192 * 2));
        CU(h, h->d_tile_base.ensure((size_t)(rows / T / 128 + 1) * 4));
        CU(h, h->d_a3[0].ensure((size_t)rows * 256 * 2)); CU(h, h->d_a3[1].ensure((size_t)rows * 256 * 2));
        // fp32 gate pre-activations exist only on the split (GEMM + recurrence) paths; the fused layer kernels never materialise them
        if (!h->trnn1_fused || !h->trnn2_fused) CU(h, h->d_zin.ensure((size_t)rows * 1024 * sizeof(float)));
    }
    bool l0_prefetched = false;      // read_rnn1 of the current iteration was already launched on the side stream
    bool tail_pending = false;       // a heads tail is (possibly) still running on the side stream
    for (int64_t c0 = 0; c0 < n_win; c0 += CH) {
        const int64_t nw = std::min(CH, n_win - c0);
        for (int mi = 0; mi < 2; ++mi) {
            if (only_model >= 0 && mi != only_model) continue;
            const ModelDev& M = h->m[mi];
            const float* heads_in = nullptr;
            int heads_stage = 0;
            int64_t heads_nwp = 0;
            if (h->path == 0) {
                // ---- fp32 SIMT path: projection fused into every recurrence step ----
                const float* in_prev = nullptr;
                for (int l = 0; l < 4; ++l) {
                    static const int simt_stage[4] = {ST_L0, ST_REC1, ST_REC2, ST_REC3};
                    StageTimer tm(h, simt_stage[l]);
                    LstmIo io;
                    io.act_in = in_prev; io.base_in = (l == 0) ? x : (l == 2 ? sig_feat[mi] : nullptr);
                    io.win_base = win_base + c0; io.act_out = h->d_act[l].as<float>();
                    int n = launch_lstm_layer(l, 0, M.lstm[l], io, nw, T, h->stream);
                    if (n < 0) return fail(h, NRV_E_INVALID, "lstm variant");
                    h->launches += n;
                    in_prev = h->d_act[l].as<float>();
                }
                heads_in = in_prev;
            } else {
                // ---- tensor-core path: every projection is a tcgen05 GEMM, every recurrence with u >= 64 a tcgen05
                //      recurrence kernel; activations travel between layers as fp16 (hi, lo) pairs in the padded
                //      time-major layout row(t, w) = t*nwp + w (see nrv_rec_tc.cu) ----
                const int64_t nwp = (nw + 127) / 128 * 128;
                const int64_t R = nwp * T;
                __half *a1h = h->d_a1[0].as<__half>(), *a1l = h->d_a1[1].as<__half>();
                __half *a2h = h->d_a2[0].as<__half>(), *a2l = h->d_a2[1].as<__half>();
                __half *a3h = h->d_a3[0].as<__half>(), *a3l = h->d_a3[1].as<__half>();
                __half *a4h = h->d_a4[0].as<__half>(), *a4l = h->d_a4[1].as<__half>();
                float* zin = h->d_zin.as<float>();
                int n;
                const int f8 = !f8_off && h->f8_rnn2 && h->trnn1_fused && h->trnn2_fused;
                const bool sig_table = h->sig_table && h->trnn1_fused && n_bases > 0;
                int32_t* tile_base = h->d_tile_base.as<int32_t>();
                // read_rnn1 (u = 16, K = 6: fp32 SIMT, fused) -> BN(h), columns [0,32) of a 64-wide zero-padded operand.
                // The cluster-of-4 kernel of total_rnn1 can only use 128 of the 148 SMs (32 co-resident clusters) and read_rnn1 is a
                // grid of small CTAs, so the read_rnn1 of the NEXT (chunk, model) is launched on a side stream as soon as read_rnn11 of
                // this one has consumed a1: it runs on the idle SMs under the fused layers.
                auto launch_l0 = [&](int64_t c0_, int mi_, cudaStream_t st) -> int {
                    const int64_t nw_ = std::min(CH, n_win - c0_);
                    StageTimer tm(h, ST_L0, st);
                    LstmIo io; io.base_in = x; io.win_base = win_base + c0_; io.out_hi = a1h; io.out_lo = a1l; io.out_ld = 64;
                    io.out_nwp = (nw_ + 127) / 128 * 128;
                    const int k = launch_read_rnn1(h->m[mi_].lstm[0], io, nw_, T, st);
                    if (k > 0) h->launches += k;
                    return k;
                };
                const bool overlap = h->overlap_l0 && h->stream2 && h->trnn1_fused && only_model < 0;
                if (!overlap || !l0_prefetched) {
                    if (launch_l0(c0, mi, h->stream) < 0) return fail(h, NRV_E_CUDA, "read_rnn1 kernel could not be launched");
                } else {
                    CU(h, cudaStreamWaitEvent(h->stream, h->ev_l0_done, 0));     // launched during the previous iteration
                }
                l0_prefetched = false;
                {   // read_rnn11: projection (K = 32 -> 64, bias as the weight row of a constant-1 column) and recurrence
                    // (u = 64) fused in one tcgen05 kernel -- no zin round trip for this layer
                    StageTimer tm(h, ST_REC1);
                    LstmIo io; io.out_hi = a2h; io.out_lo = a2l; io.out_ld = 192;
                    n = launch_lstm_fused_tc64(M.lstm[1], a1h, a1l, io, nwp, T, h->stream);
                    if (n < 0) return fail(h, NRV_E_CUDA, "tcgen05 fused layer (read_rnn11) could not be launched");
                    h->launches += n;
                }
                {   // total_rnn1 (K = 192, u = 128): gather the CNN features, then the fused layer kernel (or projection GEMM + recurrence)
                    {
                        StageTimer tm(h, ST_PROJ2);
                        // tiles without a read boundary take their CNN features straight from the per-base table (TMA in the fused
                        // kernel); only the others are gathered into columns [128, 192) of a2
                        const int64_t n_tiles = nwp >> 7;
                        if (sig_table && (mi == 0 || only_model >= 0)) {
                            tile_base_kernel<<<(unsigned)((n_tiles + 255) / 256), 256, 0, h->stream>>>(win_base + c0, nw, tile_base, n_tiles);
                            h->launches += 1;
                        }
                        gather_sig_kernel<<<dim3((unsigned)n_tiles, (unsigned)T), 256, 0, h->stream>>>(
                            h->d_sfh[mi].as<__half>(), h->d_sfl[mi].as<__half>(), win_base + c0, nw, nwp, 192, sig_table ? tile_base : nullptr, a2h, a2l);
                        h->launches += 1;
                        if (overlap) {
                            // next iteration in launch order: the other model of this chunk, or model 1 of the next chunk.  The side stream has the
                            // lowest priority and becomes eligible together with the cluster kernel (after the gather): the cluster kernel takes
                            // its 128 SMs first, read_rnn1 gets the rest
                            const int mi_n = mi == 0 ? 1 : 0;
                            const int64_t c0_n = mi == 0 ? c0 : c0 + CH;
                            if (c0_n < n_win) {
                                CU(h, cudaEventRecord(h->ev_a1_free, h->stream));
                                CU(h, cudaStreamWaitEvent(h->stream2, h->ev_a1_free, 0));
                                if (launch_l0(c0_n, mi_n, h->stream2) < 0) return fail(h, NRV_E_CUDA, "read_rnn1 kernel could not be launched");
                                CU(h, cudaEventRecord(h->ev_l0_done, h->stream2));
                                l0_prefetched = true;
                            }
                        }
                        if (!h->trnn1_fused) {
                            n = launch_gemm_f16x3(a2h, a2l, M.lstm[2].pb_hi, M.lstm[2].pb_lo, R, 1024, 192, zin, M.lstm[2].bias_tc,
                                                  1, T, nwp, 512, 0, h->num_sms, h->stream);
                            if (n < 0) return fail(h, NRV_E_CUDA, "tcgen05 projection (total_rnn1) could not be launched");
                            h->launches += n;
                        }
                    }
                    StageTimer tm(h, ST_REC2);
                    LstmIo io; io.zin = zin; io.out_hi = a3h; io.out_lo = a3l; io.out_ld = 256;
                    io.out_f8 = f8 != 0;       // total_rnn2 runs its correction passes in e4m3: a3h = fp16(h), a3l = the 8-bit copies
                    io.rec_f8 = f8 != 0 && h->f8_rnn1;     // ... and so does total_rnn1's recurrence (the copies are then made once for both)
                    if (sig_table) { io.sf_hi = h->d_sfh[mi].as<__half>(); io.sf_lo = h->d_sfl[mi].as<__half>(); io.sf_rows = n_bases; io.tile_base = tile_base; }
                    if (h->trnn1_fused) n = launch_lstm_fused_pair128(M.lstm[2], a2h, a2l, io, nwp, T, h->num_sms, h->stream);
                    else n = h->rec128_pair ? launch_lstm_rec_tc128_pair(M.lstm[2], io, nwp, T, h->stream)
                                            : launch_lstm_rec_tc128(M.lstm[2], io, nwp, T, h->stream);
                    if (n < 0) return fail(h, NRV_E_CUDA, "tcgen05 layer kernel (total_rnn1) could not be launched");
                    h->launches += n;
                }
                if (h->trnn2_fused) {
                    StageTimer tm(h, ST_REC3);
                    LstmIo io; io.out_hi = a4h; io.out_lo = a4l; io.out_ld = 128;
                    n = launch_lstm_fused_pair64(M.lstm[3], a3h, a3l, io, nwp, T, h->num_sms, h->stream, f8);
                    if (n < 0) return fail(h, NRV_E_CUDA, "tcgen05 fused layer (total_rnn2) could not be launched");
                    h->launches += n;
                } else {   // total_rnn2: projection (K = 256), recurrence (u = 64)
                    {
                        StageTimer tm(h, ST_PROJ3);
                        n = launch_gemm_f16x3(a3h, a3l, M.lstm[3].pb_hi, M.lstm[3].pb_lo, R, 512, 256, zin, M.lstm[3].bias_tc, 1,
                                              T, nwp, 256, 0, h->num_sms, h->stream);
                        if (n < 0) return fail(h, NRV_E_CUDA, "tcgen05 projection (total_rnn2) could not be launched");
                        h->launches += n;
                    }
                    StageTimer tm(h, ST_REC3);
                    LstmIo io; io.zin = zin; io.out_hi = a4h; io.out_lo = a4l; io.out_ld = 128;
                    n = launch_lstm_rec_tc64(M.lstm[3], io, nwp, T, h->stream);
                    if (n < 0) return fail(h, NRV_E_CUDA, "tcgen05 recurrence (total_rnn2) could not be launched");
                    h->launches += n;
                }
                {   // heads: relu(Dense(128 -> 128)) as a tcgen05 GEMM with relu(Dense(128 -> 32)) fused into its epilogue
                    if (tail_pending) CU(h, cudaStreamWaitEvent(h->stream, h->ev_tail_done, 0));   // the side-stream tail has read d_act[3]
                    StageTimer tm(h, ST_HEADS_GEMM);
                    n = launch_gemm_f16x3(a4h, a4l, M.heads.d1t_hi, M.heads.d1t_lo, R, 128, 128, h->d_act[3].as<float>(),
                                          M.heads.d1b, 2, T, nwp, 128, 1, h->num_sms, h->stream, M.heads.d2t_hi, M.heads.d2t_lo,
                                          M.heads.d2b);
                    if (n < 0) return fail(h, NRV_E_CUDA, "tcgen05 dense head could not be launched");
                    h->launches += n;
                }
                heads_in = h->d_act[3].as<float>();
                heads_stage = 2;
                heads_nwp = nwp;
            }
            {
                // heads tail (small CTAs, 33 us): on the side stream it runs under the next iteration's persistent kernels (they leave
                // room for small CTAs on every SM) instead of between them
                const bool side = h->path != 0 && h->overlap_l0 && h->stream2 && h->trnn1_fused;
                cudaStream_t st = side ? h->stream2 : h->stream;
                if (side) {
                    CU(h, cudaEventRecord(h->ev_hg_done, h->stream));
                    CU(h, cudaStreamWaitEvent(h->stream2, h->ev_hg_done, 0));
                }
                {
                    StageTimer tm(h, ST_HEADS, st);
                    int n = launch_heads(M.heads, heads_in, nw, T, probs[mi] ? probs[mi] + c0 * M.n_class : nullptr,
                                         labels[mi] ? labels[mi] + c0 : nullptr, heads_stage, heads_nwp, st);
                    if (n < 0) return fail(h, NRV_E_INVALID, "unsupported window length");
                    h->launches += n;
                }
                if (side) {
                    CU(h, cudaEventRecord(h->ev_tail_done, h->stream2));
                    tail_pending = true;
                }
            }
        }
    }
    if (tail_pending) CU(h, cudaStreamWaitEvent(h->stream, h->ev_tail_done, 0));     // labels / probabilities of the last chunk
    CU(h, cudaGetLastError());
    return NRV_OK;
}

// ---- refinement of near-tie windows ------------------------------------------------------------------------------------
// The e4m3 correction passes (nrv_fused_pair.cu, F8) keep the softmax outputs within the 1e-3 tolerance (measured 1.5e-4), but an
// argmax can flip where the two best classes are closer than the error.  Every window whose top-2 margin is below `tau` (default
// 1e-3 = the tolerance: a margin >= 2 max|dP| cannot flip) is therefore evaluated AGAIN with every product in three fp16 passes --
// one extra pass per model over a fixed-capacity list (no host synchronisation: unused slots hold window 0 and are ignored).
// On the unitest set this turns 81,769 / 81,770 identical labels into 81,770 (profiles/r02_gpu_precision.md); 0.01-0.1 % of the
// windows qualify.
__global__ void refine_init_kernel(int32_t* count, int32_t* list, int32_t* base, int32_t cap, const int32_t* __restrict__ win_base) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *count = 0;
    if (i < cap) { list[i] = -1; base[i] = win_base[0]; }
}
__global__ void near_tie_kernel(const float* __restrict__ probs, int nc, int64_t n_win, float tau, const int32_t* __restrict__ win_base,
                                int32_t cap, int32_t* count, int32_t* list, int32_t* base) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_win) return;
    float a = -1.f, b = -1.f;                 // the two largest probabilities
    for (int k = 0; k < nc; ++k) {
        const float p = probs[w * nc + k];
        if (p > a) { b = a; a = p; } else if (p > b) b = p;
    }
    if (!(a - b >= tau)) {                    // also catches NaN
        const int i = atomicAdd(count, 1);
        if (i < cap) { list[i] = (int32_t)w; base[i] = win_base[w]; }
    }
}
__global__ void refine_scatter_kernel(const int32_t* __restrict__ list, int32_t cap, int nc, const uint8_t* __restrict__ ty,
                                      const float* __restrict__ tp, uint8_t* __restrict__ y, float* __restrict__ p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    const int32_t w = list[i];
    if (w < 0) return;
    y[w] = ty[i];
    for (int k = 0; k < nc; ++k) p[(int64_t)w * nc + k] = tp[(int64_t)i * nc + k];
}

int refine_near_ties(nrv_handle* h, int64_t n_win, const int32_t* win_base, const float* x, float* const sig_feat[2],
                     float* const probs[2], uint8_t* const labels[2], int64_t n_bases) {
    const int32_t cap = (int32_t)std::min<int64_t>(n_win, h->refine_cap);
    if (cap <= 0) return NRV_OK;
    CU(h, h->d_ref.ensure((size_t)(2 * cap + 4) * 4));
    CU(h, h->d_ref_y.ensure((size_t)cap + 16));
    CU(h, h->d_ref_p.ensure((size_t)cap * 6 * 4 + 16));
    int32_t* count = h->d_ref.as<int32_t>();
    int32_t* list = count + 4;
    int32_t* base = list + cap;
    for (int mi = 0; mi < 2; ++mi) {
        const int nc = h->m[mi].n_class;
        refine_init_kernel<<<(cap + 255) / 256, 256, 0, h->stream>>>(count, list, base, cap, win_base);
        near_tie_kernel<<<(unsigned)((n_win + 255) / 256), 256, 0, h->stream>>>(probs[mi], nc, n_win, h->refine_tau, win_base, cap, count, list, base);
        h->launches += 2;
        float* tp[2] = {nullptr, nullptr};
        uint8_t* ty[2] = {nullptr, nullptr};
        tp[mi] = h->d_ref_p.as<float>(); ty[mi] = h->d_ref_y.as<uint8_t>();
        const int rc = run_models(h, cap, base, x, sig_feat, tp, ty, n_bases, mi, true);
        if (rc) return rc;
        refine_scatter_kernel<<<(cap + 255) / 256, 256, 0, h->stream>>>(list, cap, nc, ty[mi], tp[mi], labels[mi], probs[mi]);
        h->launches += 1;
    }
    CU(h, cudaGetLastError());
    return NRV_OK;
}

struct Offsets {
    int64_t n_reads = 0, n_samples = 0, n_bases = 0, n_win = 0;
    const int64_t *d_sig_off = nullptr, *d_base_off = nullptr, *d_win_off = nullptr;
    // read_stats: reads longer than one segment merge their histograms in global memory (slot = index among those reads, else -1)
    const int32_t* d_hist_slot = nullptr;
    int64_t n_multi = 0;
    int max_segs = 1;
};

// host offsets -> pinned staging -> device; also derives the window CSR and the histogram slots of read_stats.
int stage_offsets(nrv_handle* h, int64_t R, const int64_t* sig_off, const int64_t* base_off, Offsets* o, cudaStream_t st) {
    const size_t n = (size_t)R + 1;
    nrv_handle::IoSlot& S = h->io();
    // the upload that last used this slot's pinned staging buffer must be done before it is overwritten: wait for THAT copy
    // (an event recorded right after it, NRV_SLOTS batches ago) -- not for the stream, which would drain the batch in flight
    CU(h, cudaEventSynchronize(S.ev_h2d));
    const size_t bytes = 3 * n * sizeof(int64_t) + n * sizeof(int32_t);
    CU(h, S.h_off.ensure(bytes));
    CU(h, S.d_off.ensure(bytes));
    int64_t* hs = S.h_off.as<int64_t>();
    int64_t* hb = hs + n;
    int64_t* hw = hb + n;
    int32_t* hslot = reinterpret_cast<int32_t*>(hw + n);
    for (size_t i = 0; i < n; ++i) { hs[i] = sig_off ? sig_off[i] : 0; hb[i] = base_off[i]; }
    hw[0] = 0;
    const int64_t seg = read_stats_segment();
    int64_t n_multi = 0, max_segs = 1;
    for (int64_t r = 0; r < R; ++r) {
        if (hb[r + 1] < hb[r] || hs[r + 1] < hs[r]) return fail(h, NRV_E_INVALID, "offsets must be non-decreasing");
        const int64_t N = hb[r + 1] - hb[r];
        hw[r + 1] = hw[r] + std::max<int64_t>(N - h->window, 0);
        const int64_t nseg = (hs[r + 1] - hs[r] + seg - 1) / seg;
        hslot[r] = nseg > 1 ? (int32_t)n_multi++ : -1;
        max_segs = std::max(max_segs, nseg);
    }
    hslot[R] = -1;
    if (hb[0] != 0 || hs[0] != 0) return fail(h, NRV_E_INVALID, "offsets must start at 0");
    if (max_segs > 65535) return fail(h, NRV_E_INVALID, "a read has more than 2^31 samples");
    CU(h, cudaMemcpyAsync(S.d_off.p, hs, bytes, cudaMemcpyHostToDevice, st));
    CU(h, cudaEventRecord(S.ev_h2d, st));
    o->n_reads = R; o->n_samples = hs[R]; o->n_bases = hb[R]; o->n_win = hw[R];
    o->d_sig_off = S.d_off.as<int64_t>();
    o->d_base_off = o->d_sig_off + n;
    o->d_win_off = o->d_base_off + n;
    o->d_hist_slot = reinterpret_cast<const int32_t*>(o->d_win_off + n);
    o->n_multi = n_multi; o->max_segs = (int)max_segs;
    if (o->n_bases >= (int64_t)INT32_MAX || o->n_samples >= ((int64_t)1 << 40))
        return fail(h, NRV_E_INVALID, "batch too large (total bases must be < 2^31)");
    return NRV_OK;
}

int check_batch(nrv_handle* h, const nrv_batch* b) {
    if (!h) return NRV_E_INVALID;
    if (h->sticky) return NRV_E_CUDA;
    if (!b || b->n_reads < 0 || !b->sig_off || !b->base_off) return fail(h, NRV_E_INVALID, "bad batch");
    if (b->n_reads > 0 && (!b->signal || !b->starts || !b->bases || !b->ev_mean || !b->ev_std || !b->last_dur))
        return fail(h, NRV_E_INVALID, "batch has NULL arrays");
    return NRV_OK;
}

struct DevBatch {
    const int16_t* signal; const int32_t* starts; const uint8_t* bases; const float* evm; const float* evs;
    const int32_t* last_dur;
};

int upload_batch(nrv_handle* h, const nrv_batch* b, const Offsets& o, DevBatch* d, cudaStream_t st) {
    nrv_handle::IoSlot& S = h->io();
    CU(h, S.d_signal.ensure((size_t)o.n_samples * 2 + 64));
    CU(h, S.d_starts.ensure((size_t)o.n_bases * 4 + 16));
    CU(h, S.d_bases.ensure((size_t)o.n_bases + 16));
    CU(h, S.d_evm.ensure((size_t)o.n_bases * 4 + 16));
    CU(h, S.d_evs.ensure((size_t)o.n_bases * 4 + 16));
    CU(h, S.d_lastdur.ensure((size_t)o.n_reads * 4 + 16));
    CU(h, cudaMemcpyAsync(S.d_signal.p, b->signal, (size_t)o.n_samples * 2, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(S.d_starts.p, b->starts, (size_t)o.n_bases * 4, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(S.d_bases.p, b->bases, (size_t)o.n_bases, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(S.d_evm.p, b->ev_mean, (size_t)o.n_bases * 4, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(S.d_evs.p, b->ev_std, (size_t)o.n_bases * 4, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(S.d_lastdur.p, b->last_dur, (size_t)o.n_reads * 4, cudaMemcpyHostToDevice, st));
    d->signal = S.d_signal.as<int16_t>(); d->starts = S.d_starts.as<int32_t>(); d->bases = S.d_bases.as<uint8_t>();
    d->evm = S.d_evm.as<float>(); d->evs = S.d_evs.as<float>(); d->last_dur = S.d_lastdur.as<int32_t>();
    return NRV_OK;
}

// K1: stats + maps (+ features).  Fills shift/scale/status/base_read arenas.
int run_segment(nrv_handle* h, const DevBatch& d, const Offsets& o, bool want_x, double* seg_mean, double* seg_std) {
    CU(h, h->d_shift.ensure((size_t)o.n_reads * 8 + 16));
    CU(h, h->d_scale.ensure((size_t)o.n_reads * 8 + 16));
    CU(h, h->io().d_status.ensure((size_t)o.n_reads * 4 + 16));
    CU(h, h->d_base_read.ensure((size_t)o.n_bases * 4 + 16));
    if (want_x) CU(h, h->d_x.ensure((size_t)o.n_bases * 6 * 4 + 16));
    {
        StageTimer tm(h, ST_STATS);
        // scratch histograms of the reads that span several segments: zero on (re)allocation, the kernel leaves them zeroed
        const size_t before = h->d_ghist.cap;
        CU(h, h->d_ghist.ensure(read_stats_hist_bytes(o.n_multi)));
        if (h->d_ghist.cap != before) CU(h, cudaMemsetAsync(h->d_ghist.p, 0, h->d_ghist.cap, h->stream));
        h->launches += launch_read_stats(d.signal, o.d_sig_off, o.d_base_off, d.starts, d.last_dur, h->window, o.n_reads,
                                         o.d_hist_slot, h->d_ghist.p, o.n_multi, o.max_segs,
                                         h->d_shift.as<double>(), h->d_scale.as<double>(), h->io().d_status.as<int32_t>(),
                                         h->stream, h->d_base_read.as<int32_t>(), o.n_bases);      // + the base -> read map
    }
    {
        StageTimer tm(h, ST_FEAT);
        h->launches += launch_base_features(d.signal, o.d_sig_off, d.starts, o.d_base_off, d.bases, d.evm, d.evs,
                                            d.last_dur, h->d_base_read.as<int32_t>(), h->d_shift.as<double>(),
                                            h->d_scale.as<double>(), o.n_bases, want_x ? h->d_x.as<float>() : nullptr,
                                            seg_mean, seg_std, h->stream);
    }
    CU(h, cudaGetLastError());
    return NRV_OK;
}

// K4 scratch: tile states + ticket counter of the single-pass decode (zeroed when (re)allocated and when the epoch wraps)
int decode_scratch(nrv_handle* h, int64_t n_tiles, unsigned* epoch) {
    const size_t before = h->d_tiles.cap;
    CU(h, h->d_tiles.ensure((size_t)(n_tiles + 1) * 8 + 16));
    h->decode_epoch = h->decode_epoch % 0x3ffffeu + 1;
    if (h->d_tiles.cap != before || h->decode_epoch == 1) CU(h, cudaMemsetAsync(h->d_tiles.p, 0, h->d_tiles.cap, h->stream));
    *epoch = h->decode_epoch;
    return NRV_OK;
}


// Pick the slot for a new batch.  Sync entry points (nrv_segment, nrv_decode, nrv_predict_windows) need an idle handle.
int begin_batch(nrv_handle* h, bool need_idle) {
    if (need_idle && h->in_flight()) return fail(h, NRV_E_INVALID, "call nrv_wait on the batches in flight first");
    int c = -1;
    for (int k = 0; k < NRV_SLOTS; ++k) {
        const int cand = (h->cur + 1 + k) % NRV_SLOTS;
        if (h->slot[cand].ticket == 0) { c = cand; break; }
    }
    if (c < 0) return fail(h, NRV_E_INVALID, "too many batches in flight (nrv_wait the oldest ticket first)");
    h->cur = c;
    return NRV_OK;
}

// Enqueue one batch: H2D (host_io: on the copy stream, so that it runs under the previous batch's kernels), K1..K4 on the main
// stream.  host_io: the results stay in the slot's device arenas until finish_batch() copies them out.
int enqueue_batch(nrv_handle* h, const nrv_batch* b, nrv_result* r, bool host_io) {
    int rc = check_batch(h, b);
    if (rc) return rc;
    if (!r || !r->revised || !r->out_off || !r->status) return fail(h, NRV_E_INVALID, "result needs revised/out_off/status");
    CU(h, cudaSetDevice(h->device));
    rc = begin_batch(h, false);
    if (rc) return rc;
    nrv_handle::IoSlot& S = h->io();
    cudaStream_t up = host_io ? h->copy_stream : h->stream;
    Offsets o;
    rc = stage_offsets(h, b->n_reads, b->sig_off, b->base_off, &o, up);
    if (rc) return rc;
    if (r->revised_cap < o.n_bases) return fail(h, NRV_E_CAPACITY, "revised_cap smaller than the number of bases");
    DevBatch d;
    const bool want_q = r->revised_qual != nullptr;      // -F fastq: per-window Phred scores need the softmax outputs
    const uint8_t* d_qual_in = nullptr;
    if (host_io) {
        rc = upload_batch(h, b, o, &d, up);
        if (rc) return rc;
        if (want_q && b->qual) {
            CU(h, S.d_qual_in.ensure((size_t)o.n_bases + 16));
            CU(h, cudaMemcpyAsync(S.d_qual_in.p, b->qual, (size_t)o.n_bases, cudaMemcpyHostToDevice, up));
            d_qual_in = S.d_qual_in.as<uint8_t>();
        }
        CU(h, cudaEventRecord(S.ev_h2d, up));
        CU(h, cudaStreamWaitEvent(h->stream, S.ev_h2d, 0));
    } else {
        d.signal = b->signal; d.starts = b->starts; d.bases = b->bases; d.evm = b->ev_mean; d.evs = b->ev_std;
        d.last_dur = b->last_dur;
        d_qual_in = b->qual;
    }
    rc = run_segment(h, d, o, true, nullptr, nullptr);
    if (rc) return rc;
    // ---- K2: CNN per base, both models -------------------------------------------------------
    for (int mi = 0; mi < 2; ++mi) {
        if (h->path == 0) CU(h, h->d_sigfeat[mi].ensure((size_t)o.n_bases * NRV_SIGFEAT * 4 + 16));   // fp32 copy: SIMT path only
        CU(h, h->d_sfh[mi].ensure((size_t)o.n_bases * NRV_SIGFEAT * 2 + 16));
        CU(h, h->d_sfl[mi].ensure((size_t)o.n_bases * NRV_SIGFEAT * 2 + 16));
    }
    {
        StageTimer tm(h, ST_CNN);
        __half* sfh[2] = {h->d_sfh[0].as<__half>(), h->d_sfh[1].as<__half>()};
        __half* sfl[2] = {h->d_sfl[0].as<__half>(), h->d_sfl[1].as<__half>()};
        const int nc = launch_cnn(&h->m[0], &h->m[1], d.signal, o.d_sig_off, d.starts, o.d_base_off,
                                  h->d_base_read.as<int32_t>(), h->d_shift.as<double>(), h->d_scale.as<double>(), nullptr,
                                  o.n_bases, h->path == 0 ? h->d_sigfeat[0].as<float>() : nullptr,
                                  h->path == 0 ? h->d_sigfeat[1].as<float>() : nullptr, sfh, sfl, h->stream);
        if (nc < 0) return fail(h, NRV_E_CUDA, "CNN kernel could not be launched (tensor maps)");
        h->launches += nc;
    }
    // ---- K3: window map + Bi-LSTM stack + heads --------------------------------------------------
    CU(h, h->d_win_base.ensure((size_t)o.n_win * 4 + 16));
    h->launches += launch_window_map(o.d_base_off, o.d_win_off, nullptr, o.n_reads, h->window, h->d_win_base.as<int32_t>(),
                                     h->stream);
    float* probs[2] = {nullptr, nullptr};
    uint8_t* labels[2];
    if (host_io) {
        if (r->p1 || want_q) { CU(h, h->d_probs[0].ensure((size_t)o.n_win * 6 * 4 + 16)); probs[0] = h->d_probs[0].as<float>(); }
        if (r->p2 || want_q) { CU(h, h->d_probs[1].ensure((size_t)o.n_win * 5 * 4 + 16)); probs[1] = h->d_probs[1].as<float>(); }
        for (int mi = 0; mi < 2; ++mi) { CU(h, h->d_y[mi].ensure((size_t)o.n_win + 16)); labels[mi] = h->d_y[mi].as<uint8_t>(); }
    } else {
        probs[0] = r->p1; probs[1] = r->p2;
        if (want_q && !probs[0]) { CU(h, h->d_probs[0].ensure((size_t)o.n_win * 6 * 4 + 16)); probs[0] = h->d_probs[0].as<float>(); }
        if (want_q && !probs[1]) { CU(h, h->d_probs[1].ensure((size_t)o.n_win * 5 * 4 + 16)); probs[1] = h->d_probs[1].as<float>(); }
        uint8_t* user[2] = {r->y1, r->y2};
        for (int mi = 0; mi < 2; ++mi) {
            if (user[mi]) labels[mi] = user[mi];
            else { CU(h, h->d_y[mi].ensure((size_t)o.n_win + 16)); labels[mi] = h->d_y[mi].as<uint8_t>(); }
        }
    }
    // near-tie windows of the F8 path are evaluated again in three fp16 passes: needs the softmax outputs of every window
    const bool refine = h->refine && h->path != 0 && h->f8_rnn2 && h->trnn1_fused && h->trnn2_fused && o.n_win > 0;
    if (refine) {
        if (!probs[0]) { CU(h, h->d_probs[0].ensure((size_t)o.n_win * 6 * 4 + 16)); probs[0] = h->d_probs[0].as<float>(); }
        if (!probs[1]) { CU(h, h->d_probs[1].ensure((size_t)o.n_win * 5 * 4 + 16)); probs[1] = h->d_probs[1].as<float>(); }
    }
    float* sf[2] = {h->d_sigfeat[0].as<float>(), h->d_sigfeat[1].as<float>()};
    rc = run_models(h, o.n_win, h->d_win_base.as<int32_t>(), h->d_x.as<float>(), sf, probs, labels, o.n_bases);
    if (rc) return rc;
    if (refine) {
        rc = refine_near_ties(h, o.n_win, h->d_win_base.as<int32_t>(), h->d_x.as<float>(), sf, probs, labels, o.n_bases);
        if (rc) return rc;
    }
    // ---- K4: decode --------------------------------------------------------------------------------
    unsigned dec_epoch = 0;
    rc = decode_scratch(h, decode_tile_count(o.n_bases), &dec_epoch);
    if (rc) return rc;
    CU(h, S.d_flag.ensure(16));
    CU(h, cudaMemsetAsync(S.d_flag.p, 0, 4, h->stream));
    uint8_t* d_rev; int64_t* d_outoff;
    uint8_t* d_revq = nullptr;
    uint8_t* wq[2] = {nullptr, nullptr};
    if (host_io) {
        CU(h, S.d_revised.ensure((size_t)r->revised_cap + 16));
        CU(h, S.d_outoff.ensure((size_t)(o.n_reads + 1) * 8));
        d_rev = S.d_revised.as<uint8_t>(); d_outoff = S.d_outoff.as<int64_t>();
        if (want_q) {
            CU(h, S.d_revq.ensure((size_t)r->revised_cap + 16));
            d_revq = S.d_revq.as<uint8_t>();
        }
    } else {
        d_rev = r->revised; d_outoff = r->out_off;
        d_revq = r->revised_qual;
    }
    {
        StageTimer tm(h, ST_DECODE);
        if (want_q) {
            for (int mi = 0; mi < 2; ++mi) {
                CU(h, h->d_wq[mi].ensure((size_t)o.n_win + 16));
                wq[mi] = h->d_wq[mi].as<uint8_t>();
                const int n = launch_window_phred(probs[mi], labels[mi], mi == 0 ? 6 : 5, o.n_win, wq[mi], h->stream);
                if (n < 0) return fail(h, NRV_E_CUDA, "window_phred kernel could not be launched");
                h->launches += n;
            }
        }
        const int n = launch_decode(o.d_base_off, o.d_win_off, d.bases, labels[0], labels[1],
                                    S.d_status.as<int32_t>(), o.n_reads, o.n_bases, o.n_win, h->window, dec_epoch,
                                    h->d_tiles.as<int64_t>(), d_rev, r->revised_cap, d_outoff, S.d_flag.as<int>(),
                                    h->stream, wq[0], wq[1], d_qual_in, d_revq);
        if (n < 0) return fail(h, NRV_E_INVALID, "decode: qualities requested without per-window scores");
        h->launches += n;
    }
    CU(h, cudaGetLastError());
    if (!host_io) {
        CU(h, cudaMemcpyAsync(r->status, S.d_status.p, (size_t)o.n_reads * 4, cudaMemcpyDeviceToDevice, h->stream));
        return NRV_OK;
    }
    // labels / probabilities (diagnostics) live in arenas the next batch overwrites: copied out in stream order, now
    if (r->y1) CU(h, cudaMemcpyAsync(r->y1, labels[0], (size_t)o.n_win, cudaMemcpyDeviceToHost, h->stream));
    if (r->y2) CU(h, cudaMemcpyAsync(r->y2, labels[1], (size_t)o.n_win, cudaMemcpyDeviceToHost, h->stream));
    if (r->p1) CU(h, cudaMemcpyAsync(r->p1, probs[0], (size_t)o.n_win * 6 * 4, cudaMemcpyDeviceToHost, h->stream));
    if (r->p2) CU(h, cudaMemcpyAsync(r->p2, probs[1], (size_t)o.n_win * 5 * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaEventRecord(S.ev_compute, h->stream));
    S.ticket = h->next_ticket++;
    S.res = *r;
    S.n_reads = o.n_reads;
    S.want_q = want_q;
    return NRV_OK;
}

// D2H of a host-I/O batch on the copy stream (under the kernels of the batch enqueued after it); blocks until the results are
// in the caller's buffers.
int finish_batch(nrv_handle* h, nrv_handle::IoSlot& S) {
    const nrv_result& r = S.res;
    const int64_t R = S.n_reads;
    S.ticket = 0;
    CU(h, cudaSetDevice(h->device));
    CU(h, S.h_flag.ensure(16));
    CU(h, cudaStreamWaitEvent(h->copy_stream, S.ev_compute, 0));
    CU(h, cudaMemcpyAsync(r.out_off, S.d_outoff.p, (size_t)(R + 1) * 8, cudaMemcpyDeviceToHost, h->copy_stream));
    CU(h, cudaMemcpyAsync(r.status, S.d_status.p, (size_t)R * 4, cudaMemcpyDeviceToHost, h->copy_stream));
    CU(h, cudaMemcpyAsync(S.h_flag.p, S.d_flag.p, 4, cudaMemcpyDeviceToHost, h->copy_stream));
    CU(h, cudaStreamSynchronize(h->copy_stream));
    if (*S.h_flag.as<int>()) return fail(h, NRV_E_CAPACITY, "revised_cap too small for the revised sequences");
    const int64_t total = r.out_off[R];
    CU(h, cudaMemcpyAsync(r.revised, S.d_revised.p, (size_t)total, cudaMemcpyDeviceToHost, h->copy_stream));
    if (S.want_q) CU(h, cudaMemcpyAsync(r.revised_qual, S.d_revq.p, (size_t)total, cudaMemcpyDeviceToHost, h->copy_stream));
    CU(h, cudaStreamSynchronize(h->copy_stream));
    return NRV_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" {

const char* nrv_version(void) { return "nanoreviser-b200 0.1 (sm_100a)"; }

int nrv_create(int device, const nrv_model_weights* m1, const nrv_model_weights* m2, nrv_handle** out) {
    if (!out || !m1 || !m2) return fail(nullptr, NRV_E_INVALID, "NULL argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(nullptr, NRV_E_NODEVICE, "no CUDA device available (libnrv has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(nullptr, NRV_E_INVALID, "bad device index");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
        return fail(nullptr, NRV_E_NODEVICE, "libnrv is built for sm_100a (B200) only");
    if (m1->window != m2->window || m1->window < 5 || m1->window > NRV_MAX_T || (m1->window & 1) == 0)
        return fail(nullptr, NRV_E_INVALID, "window must be odd, 5..13, and equal for both models");
    if (m1->n_class != 6 || m2->n_class != 5) return fail(nullptr, NRV_E_INVALID, "model1 must have 6 classes, model2 5");
    nrv_handle* h = new nrv_handle();
    h->device = device;
    h->window = m1->window;
    cudaError_t e = cudaSetDevice(device);
    int prio_lo = 0, prio_hi = 0;
    if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, prio_lo);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    for (int k = 0; k < NRV_SLOTS && e == cudaSuccess; ++k) {
        e = cudaEventCreateWithFlags(&h->slot[k].ev_h2d, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->slot[k].ev_compute, cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_a1_free, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_l0_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_hg_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_tail_done, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        g_create_error = std::string("nrv_create: ") + cudaGetErrorString(e);
        delete h;
        return NRV_E_CUDA;
    }
    int rc = pack_model(h, m1, &h->m[0]);
    if (rc == NRV_OK) rc = pack_model(h, m2, &h->m[1]);
    if (rc != NRV_OK) { g_create_error = h->err; nrv_destroy(h); return rc; }
    const char* ch = getenv("NRV_CHUNK_WINDOWS");
    if (ch && atoll(ch) > 0) h->chunk_windows = atoll(ch);
    const char* pa = getenv("NRV_PATH");
    if (pa && !strcmp(pa, "simt")) h->path = 0;
    const char* t2 = getenv("NRV_TRNN2");
    if (t2 && !strcmp(t2, "split")) h->trnn2_fused = 0;
    const char* ov = getenv("NRV_OVERLAP");
    if (ov && !strcmp(ov, "0")) h->overlap_l0 = 0;
    const char* t1 = getenv("NRV_TRNN1");
    if (t1 && !strcmp(t1, "split")) h->trnn1_fused = 0;
    if (t1 && !strcmp(t1, "fused")) h->trnn1_fused = 1;
    const char* rfe = getenv("NRV_REFINE");
    if (rfe && !strcmp(rfe, "0")) h->refine = 0;
    if (getenv("NRV_REFINE_TAU") && atof(getenv("NRV_REFINE_TAU")) > 0) h->refine_tau = (float)atof(getenv("NRV_REFINE_TAU"));
    if (getenv("NRV_REFINE_CAP") && atoi(getenv("NRV_REFINE_CAP")) > 0) h->refine_cap = atoi(getenv("NRV_REFINE_CAP"));
    h->nvtx = getenv("NRV_NVTX") && atoi(getenv("NRV_NVTX")) != 0;
    const char* sge = getenv("NRV_SIGTAB");
    if (sge && !strcmp(sge, "0")) h->sig_table = 0;
    const char* f8e = getenv("NRV_F8");
    if (f8e && !strcmp(f8e, "0")) h->f8_rnn2 = 0;
    if (f8e && !strcmp(f8e, "2")) h->f8_rnn1 = 0;
    const char* r128 = getenv("NRV_REC128");
    if (r128 && !strcmp(r128, "single")) h->rec128_pair = 0;
    h->num_sms = prop.multiProcessorCount;
    *out = h;
    return NRV_OK;
}

void nrv_destroy(nrv_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->stream2) cudaStreamSynchronize(h->stream2);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    for (void* p : h->weight_allocs) cudaFree(p);
    Arena* arenas[] = {&h->d_shift, &h->d_scale, &h->d_base_read, &h->d_win_base, &h->d_x, &h->d_sigfeat[0],
                       &h->d_sigfeat[1], &h->d_act[0], &h->d_act[1], &h->d_act[2], &h->d_act[3], &h->d_probs[0],
                       &h->d_probs[1], &h->d_y[0], &h->d_y[1], &h->d_counts, &h->d_tiles, &h->d_wq[0], &h->d_wq[1],
                       &h->d_segmean, &h->d_segstd, &h->d_sigwin, &h->d_sfh[0], &h->d_sfh[1], &h->d_sfl[0],
                       &h->d_sfl[1], &h->d_a1[0], &h->d_a1[1], &h->d_a2[0], &h->d_a2[1], &h->d_a3[0], &h->d_a3[1], &h->d_a4[0],
                       &h->d_a4[1], &h->d_zin, &h->d_tile_base, &h->d_ghist, &h->d_ref, &h->d_ref_y, &h->d_ref_p};
    for (Arena* a : arenas) a->release();
    for (nrv_handle::IoSlot& S : h->slot) {
        Arena* io[] = {&S.d_signal, &S.d_starts, &S.d_bases, &S.d_evm, &S.d_evs, &S.d_lastdur, &S.d_off, &S.d_qual_in,
                       &S.d_status, &S.d_revised, &S.d_outoff, &S.d_revq, &S.d_flag};
        for (Arena* a : io) a->release();
        S.h_off.release(); S.h_flag.release();
        if (S.ev_h2d) cudaEventDestroy(S.ev_h2d);
        if (S.ev_compute) cudaEventDestroy(S.ev_compute);
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (auto& p : h->ev_pool) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    if (h->ev_a1_free) cudaEventDestroy(h->ev_a1_free);
    if (h->ev_l0_done) cudaEventDestroy(h->ev_l0_done);
    if (h->ev_hg_done) cudaEventDestroy(h->ev_hg_done);
    if (h->ev_tail_done) cudaEventDestroy(h->ev_tail_done);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

const char* nrv_last_error(const nrv_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }
int64_t nrv_launch_count(const nrv_handle* h) { return h ? h->launches : 0; }
int nrv_set_stage_timing(nrv_handle* h, int enable) {
    if (!h) return NRV_E_INVALID;
    h->fold_events();
    h->timing = enable != 0;
    for (int i = 0; i < ST_COUNT; ++i) { h->stage_ms[i] = 0.f; h->stage_launches[i] = 0; }
    return NRV_OK;
}
int nrv_stage_count(void) { return ST_COUNT; }
const char* nrv_stage_name(int i) { return (i >= 0 && i < ST_COUNT) ? kStageNames[i] : ""; }
int nrv_get_stage_ms(nrv_handle* h, float* out, int n) {
    if (!h || !out || n < ST_COUNT) return NRV_E_INVALID;
    h->fold_events();
    for (int i = 0; i < ST_COUNT; ++i) out[i] = h->stage_ms[i];
    return NRV_OK;
}
int nrv_get_stage_launches(const nrv_handle* h, int64_t* out, int n) {
    if (!h || !out || n < ST_COUNT) return NRV_E_INVALID;
    for (int i = 0; i < ST_COUNT; ++i) out[i] = h->stage_launches[i];
    return NRV_OK;
}
void* nrv_stream(const nrv_handle* h) { return h ? (void*)h->stream : nullptr; }
int nrv_synchronize(nrv_handle* h) {
    if (!h) return NRV_E_INVALID;
    CU(h, cudaStreamSynchronize(h->stream));
    return NRV_OK;
}

int nrv_segment(nrv_handle* h, const nrv_batch* b, double* shift, double* scale, double* seg_mean, double* seg_std,
                float* x, float* sig_win, int32_t* status) {
    int rc = check_batch(h, b);
    if (rc) return rc;
    CU(h, cudaSetDevice(h->device));
    rc = begin_batch(h, true);
    if (rc) return rc;
    Offsets o;
    rc = stage_offsets(h, b->n_reads, b->sig_off, b->base_off, &o, h->stream);
    if (rc) return rc;
    DevBatch d;
    rc = upload_batch(h, b, o, &d, h->stream);
    if (rc) return rc;
    if (seg_mean) CU(h, h->d_segmean.ensure((size_t)o.n_bases * 8 + 16));
    if (seg_std) CU(h, h->d_segstd.ensure((size_t)o.n_bases * 8 + 16));
    rc = run_segment(h, d, o, true, seg_mean ? h->d_segmean.as<double>() : nullptr,
                     seg_std ? h->d_segstd.as<double>() : nullptr);
    if (rc) return rc;
    if (sig_win) {
        CU(h, h->d_sigwin.ensure((size_t)o.n_bases * NRV_SIG * 4 + 16));
        h->launches += launch_sig_windows(d.signal, o.d_sig_off, d.starts, o.d_base_off, h->d_base_read.as<int32_t>(),
                                          h->d_shift.as<double>(), h->d_scale.as<double>(), o.n_bases,
                                          h->d_sigwin.as<float>(), h->stream);
        CU(h, cudaMemcpyAsync(sig_win, h->d_sigwin.p, (size_t)o.n_bases * NRV_SIG * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    if (shift) CU(h, cudaMemcpyAsync(shift, h->d_shift.p, (size_t)o.n_reads * 8, cudaMemcpyDeviceToHost, h->stream));
    if (scale) CU(h, cudaMemcpyAsync(scale, h->d_scale.p, (size_t)o.n_reads * 8, cudaMemcpyDeviceToHost, h->stream));
    if (status) CU(h, cudaMemcpyAsync(status, h->io().d_status.p, (size_t)o.n_reads * 4, cudaMemcpyDeviceToHost, h->stream));
    if (seg_mean) CU(h, cudaMemcpyAsync(seg_mean, h->d_segmean.p, (size_t)o.n_bases * 8, cudaMemcpyDeviceToHost, h->stream));
    if (seg_std) CU(h, cudaMemcpyAsync(seg_std, h->d_segstd.p, (size_t)o.n_bases * 8, cudaMemcpyDeviceToHost, h->stream));
    if (x) CU(h, cudaMemcpyAsync(x, h->d_x.p, (size_t)o.n_bases * 6 * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, cudaGetLastError());
    return NRV_OK;
}

int nrv_predict_windows(nrv_handle* h, int64_t n, const float* S, const float* X, float* p1, float* p2) {
    if (!h) return NRV_E_INVALID;
    if (h->sticky) return NRV_E_CUDA;
    if (n < 0 || (n > 0 && (!S || !X))) return fail(h, NRV_E_INVALID, "bad arguments");
    if (n == 0) return NRV_OK;
    CU(h, cudaSetDevice(h->device));
    if (h->in_flight()) return fail(h, NRV_E_INVALID, "call nrv_wait on the batches in flight first");
    const int W = h->window;
    const int64_t nb = n * W;
    if (nb >= (int64_t)INT32_MAX) return fail(h, NRV_E_INVALID, "too many windows in one call");
    CU(h, h->d_sigwin.ensure((size_t)nb * NRV_SIG * 4));
    CU(h, h->d_x.ensure((size_t)nb * 6 * 4));
    CU(h, h->d_win_base.ensure((size_t)n * 4));
    for (int mi = 0; mi < 2; ++mi) {
        if (h->path == 0) CU(h, h->d_sigfeat[mi].ensure((size_t)nb * NRV_SIGFEAT * 4));
        CU(h, h->d_sfh[mi].ensure((size_t)nb * NRV_SIGFEAT * 2 + 16));
        CU(h, h->d_sfl[mi].ensure((size_t)nb * NRV_SIGFEAT * 2 + 16));
    }
    __half* sfh[2] = {h->d_sfh[0].as<__half>(), h->d_sfh[1].as<__half>()};
    __half* sfl[2] = {h->d_sfl[0].as<__half>(), h->d_sfl[1].as<__half>()};
    CU(h, h->d_probs[0].ensure((size_t)n * 6 * 4));
    CU(h, h->d_probs[1].ensure((size_t)n * 5 * 4));
    CU(h, cudaMemcpyAsync(h->d_sigwin.p, S, (size_t)nb * NRV_SIG * 4, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->d_x.p, X, (size_t)nb * 6 * 4, cudaMemcpyHostToDevice, h->stream));
    iota_mul_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->d_win_base.as<int32_t>(), n, W);
    h->launches += 1;
    {
        const int nc = launch_cnn(&h->m[0], &h->m[1], nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                  h->d_sigwin.as<float>(), nb, h->path == 0 ? h->d_sigfeat[0].as<float>() : nullptr,
                                  h->path == 0 ? h->d_sigfeat[1].as<float>() : nullptr, sfh, sfl, h->stream);
        if (nc < 0) return fail(h, NRV_E_CUDA, "CNN kernel could not be launched (tensor maps)");
        h->launches += nc;
    }
    float* sf[2] = {h->d_sigfeat[0].as<float>(), h->d_sigfeat[1].as<float>()};
    float* probs[2] = {h->d_probs[0].as<float>(), h->d_probs[1].as<float>()};
    uint8_t* labels[2] = {nullptr, nullptr};
    int rc = run_models(h, n, h->d_win_base.as<int32_t>(), h->d_x.as<float>(), sf, probs, labels, nb);
    if (rc) return rc;
    if (p1) CU(h, cudaMemcpyAsync(p1, probs[0], (size_t)n * 6 * 4, cudaMemcpyDeviceToHost, h->stream));
    if (p2) CU(h, cudaMemcpyAsync(p2, probs[1], (size_t)n * 5 * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, cudaGetLastError());
    return NRV_OK;
}

int nrv_decode(nrv_handle* h, int64_t n_reads, const int64_t* base_off, const uint8_t* bases, const uint8_t* y1,
               const uint8_t* y2, const int32_t* status, uint8_t* revised, int64_t revised_cap, int64_t* out_off) {
    if (!h) return NRV_E_INVALID;
    if (h->sticky) return NRV_E_CUDA;
    if (n_reads < 0 || !base_off || !revised || !out_off) return fail(h, NRV_E_INVALID, "bad arguments");
    CU(h, cudaSetDevice(h->device));
    int rc = begin_batch(h, true);
    if (rc) return rc;
    nrv_handle::IoSlot& S = h->io();
    Offsets o;
    rc = stage_offsets(h, n_reads, nullptr, base_off, &o, h->stream);
    if (rc) return rc;
    if (o.n_bases > 0 && !bases) return fail(h, NRV_E_INVALID, "bases is NULL");
    if (o.n_win > 0 && (!y1 || !y2)) return fail(h, NRV_E_INVALID, "labels are NULL");
    if (revised_cap < o.n_bases) return fail(h, NRV_E_CAPACITY, "revised_cap smaller than the number of bases");
    CU(h, S.d_bases.ensure((size_t)o.n_bases + 16));
    CU(h, h->d_y[0].ensure((size_t)o.n_win + 16));
    CU(h, h->d_y[1].ensure((size_t)o.n_win + 16));
    CU(h, S.d_status.ensure((size_t)n_reads * 4 + 16));
    CU(h, h->d_base_read.ensure((size_t)o.n_bases * 4 + 16));
    unsigned dec_epoch = 0;
    {
        const int rc2 = decode_scratch(h, decode_tile_count(o.n_bases), &dec_epoch);
        if (rc2) return rc2;
    }
    CU(h, S.d_flag.ensure(16));
    CU(h, S.h_flag.ensure(16));
    CU(h, S.d_revised.ensure((size_t)revised_cap + 16));
    CU(h, S.d_outoff.ensure((size_t)(n_reads + 1) * 8));
    CU(h, cudaMemsetAsync(S.d_flag.p, 0, 4, h->stream));
    CU(h, cudaMemcpyAsync(S.d_bases.p, bases, (size_t)o.n_bases, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->d_y[0].p, y1, (size_t)o.n_win, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->d_y[1].p, y2, (size_t)o.n_win, cudaMemcpyHostToDevice, h->stream));
    if (status) CU(h, cudaMemcpyAsync(S.d_status.p, status, (size_t)n_reads * 4, cudaMemcpyHostToDevice, h->stream));
    else CU(h, cudaMemsetAsync(S.d_status.p, 0, (size_t)n_reads * 4 + 16, h->stream));
    {
        StageTimer tm(h, ST_DECODE);
        h->launches += launch_decode(o.d_base_off, o.d_win_off, S.d_bases.as<uint8_t>(),
                                     h->d_y[0].as<uint8_t>(), h->d_y[1].as<uint8_t>(), S.d_status.as<int32_t>(), n_reads,
                                     o.n_bases, o.n_win, h->window, dec_epoch, h->d_tiles.as<int64_t>(),
                                     S.d_revised.as<uint8_t>(), revised_cap, S.d_outoff.as<int64_t>(), S.d_flag.as<int>(),
                                     h->stream);
    }
    CU(h, cudaMemcpyAsync(out_off, S.d_outoff.p, (size_t)(n_reads + 1) * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(S.h_flag.p, S.d_flag.p, 4, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (*S.h_flag.as<int>()) return fail(h, NRV_E_CAPACITY, "revised_cap too small for the revised sequences");
    CU(h, cudaMemcpyAsync(revised, S.d_revised.p, (size_t)out_off[n_reads], cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, cudaGetLastError());
    return NRV_OK;
}

int nrv_debug_gemm(nrv_handle* h, int64_t M, int N, int K, const float* A, const float* Bt, const float* bias, float* C) {
    if (!h) return NRV_E_INVALID;
    if (h->sticky) return NRV_E_CUDA;
    if (M <= 0 || N <= 0 || K <= 0 || !A || !Bt || !C) return fail(h, NRV_E_INVALID, "bad arguments");
    CU(h, cudaSetDevice(h->device));
    Arena a32, b32, ah, al, bh, bl, c32, bi;
    int rc = NRV_OK;
    auto cleanup = [&]() { a32.release(); b32.release(); ah.release(); al.release(); bh.release(); bl.release(); c32.release(); bi.release(); };
    do {
        if (a32.ensure((size_t)M * K * 4) || b32.ensure((size_t)N * K * 4) || ah.ensure((size_t)M * K * 2) ||
            al.ensure((size_t)M * K * 2) || bh.ensure((size_t)N * K * 2) || bl.ensure((size_t)N * K * 2) ||
            c32.ensure((size_t)M * N * 4) || bi.ensure((size_t)N * 4)) { rc = fail(h, NRV_E_CUDA, "debug gemm: cudaMalloc"); break; }
        cudaMemcpyAsync(a32.p, A, (size_t)M * K * 4, cudaMemcpyHostToDevice, h->stream);
        cudaMemcpyAsync(b32.p, Bt, (size_t)N * K * 4, cudaMemcpyHostToDevice, h->stream);
        if (bias) cudaMemcpyAsync(bi.p, bias, (size_t)N * 4, cudaMemcpyHostToDevice, h->stream);
        cudaMemsetAsync(c32.p, 0xFF, (size_t)M * N * 4, h->stream);      // NaN pattern: unwritten outputs are visible
        h->launches += launch_split_f16(a32.as<float>(), ah.as<__half>(), al.as<__half>(), M * K, h->stream);
        h->launches += launch_split_f16(b32.as<float>(), bh.as<__half>(), bl.as<__half>(), (int64_t)N * K, h->stream);
        int n = launch_gemm_f16x3(ah.as<__half>(), al.as<__half>(), bh.as<__half>(), bl.as<__half>(), M, N, K, c32.as<float>(),
                                  bias ? bi.as<float>() : nullptr, 0, 1, M, N, 0, h->num_sms, h->stream);
        if (n < 0) { rc = fail(h, NRV_E_INVALID, "debug gemm: unsupported shape or tensor-map failure"); break; }
        h->launches += n;
        cudaMemcpyAsync(C, c32.p, (size_t)M * N * 4, cudaMemcpyDeviceToHost, h->stream);
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { rc = fail(h, NRV_E_CUDA, std::string("debug gemm: ") + cudaGetErrorString(e)); break; }
    } while (0);
    cleanup();
    return rc;
}

int nrv_submit_batch(nrv_handle* h, const nrv_batch* b, nrv_result* r, int64_t* ticket) {
    if (!h || !ticket) return NRV_E_INVALID;
    *ticket = 0;
    const int rc = enqueue_batch(h, b, r, true);
    if (rc == NRV_OK) *ticket = h->io().ticket;
    return rc;
}

int nrv_wait_batch(nrv_handle* h, int64_t ticket) {
    if (!h) return NRV_E_INVALID;
    if (h->sticky) return NRV_E_CUDA;
    for (nrv_handle::IoSlot& S : h->slot)
        if (ticket != 0 && S.ticket == ticket) return finish_batch(h, S);
    return fail(h, NRV_E_INVALID, "unknown ticket (already waited for?)");
}

int nrv_revise_batch(nrv_handle* h, const nrv_batch* b, nrv_result* r) {
    int64_t t = 0;
    const int rc = nrv_submit_batch(h, b, r, &t);
    return rc != NRV_OK ? rc : nrv_wait_batch(h, t);
}

int nrv_revise_batch_device(nrv_handle* h, const nrv_batch* b, nrv_result* r) { return enqueue_batch(h, b, r, false); }

}  // extern "C"
