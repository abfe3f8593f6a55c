/*
 * aniportrait_b200 — C ABI of the B200-native (sm_100a) AniPortrait denoising hot path.
 *
 * The reference (Zejun-Yang/AniPortrait) is pure Python/PyTorch and has no FFI: its "operator interface" for this
 * path is the set of torch/diffusers library calls issued by src/models/*.py. Each entry point below replaces one such
 * family of calls (cited per function as reference file:line); the Python host mirror in aniportrait_b200/models and
 * aniportrait_b200/pipelines binds them with ctypes (see INTEGRATION.md for the binding a reference maintainer adds).
 *
 * Conventions
 *   - plain C types only; all tensors are raw CUDA device pointers owned by the caller (torch in practice)
 *   - activations are fp16, channels-last: images [frames, H, W, C] == token matrices [frames*H*W, C]
 *   - `stream` is a cudaStream_t passed as void*; functions only enqueue work: no allocation, no synchronisation
 *   - return 0 on success, a negative AP_ERR_* otherwise; ap_last_error() gives a thread-local message
 *   - there is NO CPU fallback: every function fails if the CUDA device is not a compute-capability 10.x GPU
 */
#ifndef ANIPORTRAIT_B200_H_
#define ANIPORTRAIT_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define AP_VERSION 100

#define AP_OK 0
#define AP_ERR_INVALID (-1)  /* bad argument / unsupported shape */
#define AP_ERR_CUDA (-2)     /* CUDA runtime / driver error      */
#define AP_ERR_DEVICE (-3)   /* not an sm_100 device             */

/* flags for ap_gemm_f16 */
#define AP_GEMM_GEGLU 1 /* weight rows interleaved [16 value | 16 gate]; out = value * gelu_erf(gate), N/2 columns */

int ap_version(void);
const char* ap_last_error(void);
/* Binds the library to `device` (cudaSetDevice), verifies sm_100, resolves the driver entry points it needs. */
int ap_init(int device);

/*
 * out[M, N] = A[M, K1 (+K2)] . W[N, K1+K2]^T (+ bias) (+ residual)            -- fp16 in, fp32 accumulate, fp16 out
 * Replaces nn.Linear / 1x1 Conv2d / diffusers Attention.to_{q,k,v,out} / FeedForward projections:
 *   reference src/models/transformer_3d.py:64-66,93-95,124-160; src/models/attention.py:323-361;
 *   src/models/motion_module.py:122,144,163-170,233; src/models/resnet.py:207-209 (1x1 conv_shortcut, two-source
 *   K = torch.cat([hidden, skip]) of src/models/unet_3d_blocks.py:697,826 without materialising the concat).
 * a/a2: row-major fp16, leading dims lda/lda2 (elements); a2 may be NULL. w: [N, K1+K2] row-major contiguous.
 * bias: fp32 [groups, N] or NULL; output row m uses bias row m / bias_group_rows (<=0: one shared row).
 * residual: fp16 [M, ldr] or NULL. n_valid: columns >= n_valid are not written (<=0: all).
 * block_n: 0 = auto (N must be a multiple of 32).
 */
int ap_gemm_f16(const void* a, long long lda, int K1, const void* a2, long long lda2, int K2, const void* w,
                long long M, int N, const float* bias, long long bias_group_rows, const void* residual,
                long long ldr, void* out, long long ldo, int n_valid, int flags, int block_n, void* stream);

/*
 * 3x3 convolution, zero padding 1, stride 1|2, channels-last fp16, as an implicit GEMM (no im2col buffer).
 * Replaces InflatedConv3d / Downsample3D / Upsample3D.conv (reference src/models/resnet.py:10-18,52,107,166,195)
 * and conv_in/conv_out (src/models/unet_3d.py:90,250).
 * x: [Nf, H, W, C1]; x2: optional [Nf, H, W, C2] concatenated after x along channels; C1, C2 multiples of 64.
 * w: [Cout, 3, 3, C1+C2] (tap-major, channel-minor) fp16. out/residual: [Nf, H/stride, W/stride, ldo].
 */
int ap_conv3x3_nhwc_f16(const void* x, int C1, const void* x2, int C2, int Nf, int H, int W, int stride,
                        const void* w, int Cout, const float* bias, long long bias_group_rows,
                        const void* residual, void* out, long long ldo, int n_valid, int block_n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ANIPORTRAIT_B200_H_ */
