/*
 * unispeech_b200 -- C ABI of the B200-native WavLM / UniSpeech-SAT encoder hot path.
 *
 * Every entry point enqueues hand-written sm_100a kernels on the given CUDA stream and returns
 * without synchronising.  All pointers are DEVICE pointers owned by the caller (PyTorch tensors);
 * the library never allocates or frees caller memory on the hot path.
 * Return value: 0 on success, negative on error (message via b200s_last_error(), thread-local).
 * There is no CPU fallback: on a device that is not compute capability 10.x every call fails.
 *
 * The reference (microsoft/UniSpeech) has no FFI for this path: it is plain PyTorch module code.
 * Each function below cites the reference lines whose library calls (cuDNN conv, cuBLAS GEMM,
 * F.multi_head_attention_forward, F.layer_norm, F.group_norm, F.gelu) it replaces.
 * Activations are bf16, accumulation fp32, parameters/gradients fp32 masters.
 */
#ifndef UNISPEECH_B200_H_
#define UNISPEECH_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b200s_stream; /* cudaStream_t */

int b200s_version(void);
const char* b200s_last_error(void);
/* 0 if the current device can run the kernels (sm_100), negative otherwise */
int b200s_check_device(void);

/* ---- fused GEMM epilogue description (all tensors optional unless noted) ---------------------
 * value = acc (+ bias[col]);  if gelu: out_pre <- value (optional), value = gelu(value)
 *         if dgelu: value *= gelu'(gelu_aux[row,col]);  value += res1 + res2;  out <- value
 * bs = batch stride, ld = row stride, in elements.  colsum (fp32[N]) accumulates column sums of
 * the stored values (bias gradients). */
typedef struct {
  const float* bias;
  const void* res1; long long res1_bs, res1_ld;
  const void* res2; long long res2_bs, res2_ld;
  const void* gelu_aux; long long aux_bs, aux_ld;
  void* out_pre; long long pre_bs, pre_ld;
  float* colsum;
  int gelu;
  int dgelu;
} b200s_epilogue;

/* out[b, r, 0:N] = epilogue( A[b, r, 0:K] . W[N,K]^T ),  bf16 in / bf16 out, fp32 accumulate.
 * A rows live at a + b*a_bs + r*a_rs (elements) and may OVERLAP (a_rs < K): this is how the strided
 * Conv1d layers of ConvFeatureExtractionModel (WavLM/WavLM.py:400-403,485-504) become GEMMs on a
 * channels-last [B,T,C] activation (row = k*C window, row stride = stride*C).  Also used for every
 * nn.Linear forward / input-gradient (q,k,v,out_proj: WavLM/modules.py:540-563; fc1/fc2:
 * WavLM/WavLM.py:706-739; post_extract_proj: WavLM/WavLM.py:347-348). */
int b200s_gemm_rows(const void* a, long long a_bs, long long a_rs, int rows, int batches, int K,
                    const void* w, int N, void* out, long long out_bs, long long out_ld,
                    const b200s_epilogue* epi, b200s_stream stream);

/* dW[n, k] += sum_{b,r} Y[b,r,n] * X[b,r,k]   (fp32 atomic accumulation into dw, row stride dw_ld).
 * Weight gradient of the GEMMs above (autograd of nn.Linear / nn.Conv1d in the reference).
 * X rows may overlap like A above (conv im2col view).  N and K are multiples of 8. */
int b200s_gemm_wgrad(const void* y, long long y_bs, long long y_rs, const void* x, long long x_bs,
                     long long x_rs, int rows, int batches, int N, int K, float* dw, long long dw_ld,
                     b200s_stream stream);

/* Grouped positional convolution as an implicit GEMM (TransformerEncoder.pos_conv,
 * WavLM/WavLM.py:514-527,577-579; SamePad WavLM/modules.py:72-83).
 *   out[b,t,g*Cg+n] = epilogue( sum_{j<taps} sum_{c<Cg} xpad[b, t+j, g*Cg+c] * wp[g*64+n, j*64+c] )
 * xpad: [B, Tpad, D] bf16 with row stride D (zero rows around the T valid frames; the caller offsets
 * the pointer so that tap j of output frame t reads row t+j);  wp: [G*64, taps*64] bf16, zero padded. */
int b200s_posconv_gemm(const void* xpad, long long xpad_bs, int T, int B, int D, int G, int taps,
                       const void* wp, void* out, long long out_bs, long long out_ld,
                       const b200s_epilogue* epi, b200s_stream stream);

/* dWp[g*Cg+n... ] : dwp[g, n, j, c] += sum_{b,t} dy[b,t,g*Cg+n] * xpad[b,t+j,g*Cg+c]
 * dwp: fp32 [G, Cg, taps, 64] (c padded to 64; columns >= Cg hold garbage-free zeros are NOT guaranteed:
 * only c < Cg is meaningful). */
int b200s_posconv_wgrad(const void* dy, long long dy_bs, long long dy_rs, const void* xpad, long long xpad_bs,
                        int T, int B, int D, int G, int taps, float* dwp, b200s_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* UNISPEECH_B200_H_ */
