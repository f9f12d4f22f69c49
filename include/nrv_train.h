/* Training operators of the two NanoReviser networks: the device kernels behind
 * nanoreviser_b200/train.py, which mirrors Keras' Model.fit for the graphs of nanorevutils/lstmmodel.py:32-133 and is what
 * NanoReviser_train.py:164-199 calls (att_model_train.fit / save_weights).
 *
 * C ABI: plain pointers and sizes; every pointer is a DEVICE pointer to float32 (labels / masks as noted), every matrix is
 * row-major with an explicit leading dimension where one is taken, `stream` is a cudaStream_t passed as void* (NULL = default
 * stream).  All calls are asynchronous on that stream and return 0, or a negative value when the launch was rejected
 * (nrvt_last_error() has the text).  Nothing here falls back to the CPU.
 *
 * A first, correct form: one launch per operator and LSTM timestep (the GEMM on the tensor cores as 3 x TF32, the rest fp32 SIMT) (a training step is ~600 calls, which
 * train.py captures into a CUDA graph: nothing here allocates or synchronises, per-step scalars are read from device memory);
 * gradients are checked against an fp64 autograd graph of the same network in tests/test_train_gpu.py.
 */
#ifndef NRV_TRAIN_H
#define NRV_TRAIN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* nrvt_last_error(void);

/* C[M,N] = alpha * op(A) . op(B) + beta * C.  ta == 0: A is [M,K] (lda >= K), else A is stored [K,M] (lda >= M); likewise tb for
 * B ([K,N] resp. stored [N,K]).  Replaces K.dot inside Dense / LSTM (keras/layers/core.py Dense.call, recurrent.py LSTMCell.call). */
int nrvt_gemm(void* stream, int ta, int tb, int M, int N, int K, float alpha, const float* A, int lda, const float* B, int ldb,
              float beta, float* C, int ldc);
/* Y[M,N] (ldy) += bias[N]; relu != 0 applies max(., 0) afterwards (Dense(activation='relu')). */
int nrvt_bias_act(void* stream, float* Y, int M, int N, int ldy, const float* bias, int relu);
/* dY *= (Y > 0), element-wise over n values (gradient of relu through its output). */
int nrvt_relu_bwd(void* stream, float* dY, const float* Y, int64_t n);
/* out[N] = beta * out + column sums of X[M,N] (ldx): bias gradients. */
int nrvt_colsum(void* stream, const float* X, int M, int N, int ldx, float* out, float beta);
/* dst[rows, cols] (ldd) = src (lds), or += when accumulate != 0: concatenation / slicing of feature blocks. */
int nrvt_copy2d(void* stream, float* dst, int ldd, const float* src, int lds, int rows, int cols, int accumulate);

/* Conv1D(kernel 3, padding 'same', activation relu) over n sequences of L positions, channels last (nanorevcnn.py:24):
 * X [n,L,cin], W [3,cin,cout] (Keras layout), Y [n,L,cout]; cin, cout <= 8. */
int nrvt_conv1d_fwd(void* stream, const float* X, const float* W, const float* b, float* Y, int n, int L, int cin, int cout);
/* dY = gradient w.r.t. the relu output Y.  dX may be NULL (first layer).  dW [3,cin,cout] and db [cout] are OVERWRITTEN.
 * work: >= 3*cin*cout + cout doubles of scratch. */
int nrvt_conv1d_bwd(void* stream, const float* X, const float* W, const float* Y, const float* dY, float* dX, float* dW, float* db,
                    double* work, int n, int L, int cin, int cout);
/* Y[n,L,C] += X[n,L,1] broadcast over the channels (the Add() of identity_Block on a 1-channel input, nanorevcnn.py:37). */
int nrvt_add_bcast(void* stream, float* Y, const float* X, int64_t rows, int C);

/* BatchNormalization in training mode over the rows of X[rows,C] (C a power of two <= 256): batch mean and biased variance
 * are written to mean[C] / var[C], Y = gamma (X - mean) / sqrt(var + eps) + beta.  work: >= 2*C doubles of scratch. */
int nrvt_bn_fwd(void* stream, const float* X, const float* gamma, const float* beta, float eps, float* Y, float* mean, float* var,
                double* work, int64_t rows, int C);
/* Y = gamma (X - mean) / sqrt(var + eps) + beta with GIVEN statistics (inference form: the moving averages; validation passes). */
int nrvt_bn_apply(void* stream, const float* X, const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                  float* Y, int64_t rows, int C);
/* dX, dgamma[C], dbeta[C] (overwritten) from dY and the saved batch statistics. */
int nrvt_bn_bwd(void* stream, const float* X, const float* dY, const float* gamma, const float* mean, const float* var, float eps,
                float* dX, float* dgamma, float* dbeta, double* work, int64_t rows, int C);

/* One LSTM timestep (Keras 2.2.4 LSTMCell, gates i,f,c,o, recurrent_activation hard_sigmoid, activation tanh):
 * z[B,4u] holds x_t.Wk + b + h_{t-1}.Wr on entry and the ACTIVATED gates (i, f, g, o) on return; c_prev may be NULL (zero
 * state); c[B,u] and h (row stride ldh) are written. */
int nrvt_lstm_cell_fwd(void* stream, float* z, const float* c_prev, float* c, float* h, int ldh, int B, int u);
/* Backward of one timestep: gates = what nrvt_lstm_cell_fwd left in z; dh_out (row stride ldh) is the gradient arriving from
 * the layer above, dh_rec[B,u] the one from the next timestep (may be NULL); dc[B,u] carries dL/dc_t in and dL/dc_{t-1} out;
 * dz[B,4u] receives the gradient w.r.t. the pre-activations. */
int nrvt_lstm_cell_bwd(void* stream, const float* gates, const float* c_prev, const float* c, const float* dh_out, int ldh,
                       const float* dh_rec, float* dc, float* dz, int B, int u);

/* sparse_categorical_crossentropy on softmax(logits[B,nc]) with per-class weights cw[nc] (NULL = 1): probs[B,nc] and
 * dlogits[B,nc] = scale * w_y (p - onehot(y)) are written; stats[0] += sum of w_y * ce, stats[1] += #(argmax == y). */
int nrvt_softmax_ce(void* stream, const float* logits, const int32_t* labels, const float* cw, float* probs, float* dlogits,
                    float* stats, int B, int nc, float scale);
/* Center loss (lstmmodel.py:65-67): l2_i = sum_k (feat[i,k] - centers[y_i,k])^2.  dfeat += scale * 2 (feat - c_y);
 * dcenters[y_i] -= the same (dcenters must be zeroed by the caller); stats[2] += sum_i l2_i; stats[3] += #(round(l2_i) == 0), the
 * numerator of the 'accuracy' Keras reports for this output against its all-zero target (dim == 16 only). */
int nrvt_center_loss(void* stream, const float* feat, const int32_t* labels, const float* centers, float* dfeat, float* dcenters,
                     float* stats, int B, int dim, float scale);
/* keep-mask of Dropout(rate) for n elements: mask[i] = 1 with probability 1 - rate, a pure function of (seed, *step, i); the step
 * number is read from DEVICE memory so that a captured CUDA graph of the training step draws a new mask at every replay. */
int nrvt_dropout_mask(void* stream, uint8_t* mask, int64_t n, uint64_t seed, const int64_t* step, float rate);
/* X *= mask * scale (mask: uint8 0/1): Dropout forward and backward. */
int nrvt_dropout(void* stream, float* X, const uint8_t* mask, int64_t n, float scale);
/* Keras 2.2.4 Adam: m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr_t m / (sqrt(v) + eps); the bias-corrected step size
 * lr_t = lr sqrt(1 - b2^t) / (1 - b1^t) is read from DEVICE memory (one float, written by the caller before the step: graph replays). */
int nrvt_adam(void* stream, float* p, const float* g, float* m, float* v, int64_t n, const float* lr_t, float b1, float b2, float eps);
/* moving = momentum * moving + (1 - momentum) * batch (the BatchNormalization update ops). */
int nrvt_ema(void* stream, float* moving, const float* batch, int64_t n, float momentum, float batch_scale);

#ifdef __cplusplus
}
#endif
#endif
