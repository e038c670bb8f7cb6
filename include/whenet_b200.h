/*
 * whenet_b200.h - C ABI of the B200-native WHENet per-crop forward.
 *
 * The reference has no FFI of its own: its whole hot path is the Python class
 * in reference whenet.py (WHENet.__init__ :7-20, WHENet.get_angle :22-34,
 * self.model.predict :27) sitting on Keras/TensorFlow.  This header is the
 * boundary a host in any language binds instead of Keras; each entry point
 * cites the reference line(s) it replaces.  Plain C types only, no C++
 * exceptions cross it.  Every function returns 0 on success or a negative
 * WHENET_E* code; whenet_last_error() returns the message of the last failure
 * on the calling thread.
 *
 * A context is bound to one device and one stream and is NOT thread-safe
 * (the reference is single-threaded and synchronous as well: demo_video.py:49-63).
 * Use one context per GPU / per host thread.
 */
#ifndef WHENET_B200_H
#define WHENET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WHENET_OK              0
#define WHENET_EINVAL         -1   /* bad argument (NULL, n<1, n>max_batch, wrong shape)      */
#define WHENET_ECUDA          -2   /* a CUDA runtime/driver call failed (message has details) */
#define WHENET_ENOWEIGHTS     -3   /* forward called before whenet_load_weights               */
#define WHENET_ESHAPE         -4   /* a weight tensor is missing or has the wrong shape       */
#define WHENET_ENOTFOUND      -5   /* unknown tap / kernel name                               */

#define WHENET_PRECISION_FP32  0   /* fp32 storage + fp32 FMA: the parity mode                */
#define WHENET_PRECISION_BF16  1   /* bf16 activations, fp32 accumulate: the throughput mode  */
#define WHENET_PRECISION_FP16  2   /* fp16 activations, fp32 accumulate                       */

#define WHENET_IMG        224
#define WHENET_N_YAW      120      /* reference whenet.py:11 */
#define WHENET_N_PITCH     66      /* reference whenet.py:12 */
#define WHENET_N_ROLL      66      /* reference whenet.py:13 */
#define WHENET_N_LOGITS   252

typedef struct whenet_ctx whenet_ctx;

/* One named float32 tensor of the Keras weights file ("conv2d_1/kernel:0", ...),
 * row-major in the file's own layout (conv HWIO, depthwise [kh,kw,C,1], BN [C],
 * Dense [in,out]).  Borrowed for the duration of whenet_load_weights only. */
typedef struct {
    const char*  name;
    const float* data;
    int32_t      ndim;
    int64_t      dims[4];
} whenet_tensor;

/* Per-kernel device timings of the last profiled forward (CUDA events on the
 * context's stream).  bytes/flops are the ALGORITHMIC figures of DESIGN.md. */
typedef struct {
    char   name[48];
    float  ms;          /* summed over `launches` launches                 */
    int    launches;
    double bytes;       /* algorithmic HBM bytes those launches must move   */
    double flops;       /* algorithmic flops (2*MAC) of those launches      */
} whenet_kernel_stat;

/* replaces: graph construction, reference whenet.py:8-14 (+ device choice via
 * CUDA_VISIBLE_DEVICES, demo_video.py:79-80).  `max_batch` bounds n of one
 * forward call (activation workspace is sized for min(max_batch, chunk)). */
int whenet_create(whenet_ctx** out, int device, int max_batch, int precision);

/* replaces: self.model.load_weights(snapshot), reference whenet.py:15-16.
 * Takes the file's raw tensors by name; folds BatchNorm (eps 1e-3) into the
 * preceding conv, re-lays weights out for the kernels and uploads them. */
int whenet_load_weights(whenet_ctx* ctx, const whenet_tensor* tensors, int n_tensors);

/* Persisted packed artefact: the device image whenet_load_weights built - BatchNorm folded, 1x1 kernels
 * transposed to K-major and rounded to the context's storage type, K1 constants pre-halved - as two flat arenas plus an
 * index of offsets.  Export: call with NULL buffers to get sizes[3] = {fp32 elements, 16-bit elements, index entries}, then
 * with buffers of those sizes.  Import replaces whenet_load_weights (reference whenet.py:15-16) for a context of the SAME
 * precision: no HDF5 walk, no folding, no repacking - one validation pass over the index and two uploads. */
int whenet_export_packed(whenet_ctx* ctx, float* arena_f32, uint16_t* arena_16, int64_t* index, int64_t sizes[3]);
int whenet_import_packed(whenet_ctx* ctx, const float* arena_f32, int64_t n_f32, const uint16_t* arena_16, int64_t n_16,
                         const int64_t* index, int64_t n_index);

/* Run on an existing CUDA stream (cudaStream_t / CUstream) instead of the
 * context's own; NULL restores the internal stream. */
int whenet_set_stream(whenet_ctx* ctx, void* cuda_stream);

/* replaces: WHENet.get_angle(img), reference whenet.py:22-34, for uint8 input.
 * `nhwc_rgb`: n x 224 x 224 x 3 RGB bytes (host memory, or device memory when
 * in_is_device != 0).  Normalisation (whenet.py:25-26) happens on the device
 * through a 3x256 table holding float32(((v/255)-mean)/std) evaluated in
 * float64 exactly as numpy does there.  `angles_out`: n x 3 floats
 * (yaw, pitch, roll in degrees) ; `logits_out`: NULL or n x 252 floats
 * (yaw 120 | pitch 66 | roll 66) = what self.model.predict returns
 * (whenet.py:27).  Outputs go to host memory, or to device memory when
 * out_is_device != 0 (then the call is asynchronous on the context's stream). */
int whenet_forward_u8(whenet_ctx* ctx, const uint8_t* nhwc_rgb, int n, int in_is_device,
                      float* angles_out, float* logits_out, int out_is_device);

/* Asynchronous form of whenet_forward_u8 for HOST buffers (both must be pinned, see whenet_host_alloc): queues the
 * H2D copy (copy stream), the forward and the D2H of the results, then returns.  Up to TWO calls may be in flight
 * (input staging and result buffers are double-buffered), so the upload of batch i+1 overlaps the compute of batch i.
 * Call whenet_synchronize before reading `angles_host` / `logits_host` or reusing `in_host`. */
int whenet_forward_u8_async(whenet_ctx* ctx, const uint8_t* in_host, int n, float* angles_host, float* logits_host);

/* replaces: self.model.predict(img, batch_size=8), reference whenet.py:27, for
 * an already normalised float32 n x 224 x 224 x 3 input. */
int whenet_forward_f32(whenet_ctx* ctx, const float* nhwc_normalised, int n, int in_is_device,
                       float* angles_out, float* logits_out, int out_is_device);

/* Crop front-end of the stream path; replaces, for ALL heads of a frame at once, the per-head
 * slice + cv2.cvtColor(BGR2RGB) + cv2.resize(.., (224,224)) of reference demo_video.py:21-23 (and demo.py:10-11).
 * `frame`: H x W x 3 uint8 (host, or device when frame_is_device); `rects`: m x 4 int32 host array of slice
 * bounds (y0, y1, x0, x1) with 0 <= y0 < y1 <= H, 0 <= x0 < x1 <= W (the margin arithmetic of
 * demo_video.py:13-19 is host-side, see whenet_b200/crops.py); `swap_rb` != 0 converts BGR -> RGB.
 * `crops_out`: m x 224 x 224 x 3 uint8 in DEVICE memory (feed it to whenet_forward_u8 with in_is_device=1),
 * bit-identical to OpenCV's 8-bit INTER_LINEAR resize.  Asynchronous on the context's stream. */
int whenet_crop_resize_u8(whenet_ctx* ctx, const uint8_t* frame, int H, int W, int frame_is_device,
                          const int32_t* rects, int m, int swap_rb, uint8_t* crops_out);

/* Block until everything queued by this context has finished. */
int whenet_synchronize(whenet_ctx* ctx);

/* Pinned host memory for asynchronous input staging (optional). */
void* whenet_host_alloc(size_t bytes);
void  whenet_host_free(void* p);

/* ---- test / measurement hooks (no reference counterpart) ---- */

/* Keep float32 copies of intermediate tensors of the NEXT forwards
 * ("stem", "dw1".."dw16", "gate1".."gate16", "block1".."block16", "head",
 * "pooled"); only for n <= 8. */
int whenet_debug_enable_taps(whenet_ctx* ctx, int enable);
/* Copy a tap to host; *n_elems receives its element count (call with out=NULL to query). */
int whenet_debug_tap(whenet_ctx* ctx, const char* name, float* out, size_t cap_elems, size_t* n_elems);

/* Run ONE 1x1 convolution through the kernel family chosen by use_tc (0 CUDA-core, 1 tcgen05):
 * out[m,n] = act(bias[n] + sum_k A[m,k]*gate[m/hw,k]*W[k,n]) (+ resid[m,n]).  All arrays are host
 * float32 (converted to the context's storage type on the way in and back on the way out);
 * gate / resid may be NULL.  Returns WHENET_EINVAL when the family cannot run the shape. */
int whenet_debug_conv1x1(whenet_ctx* ctx, int use_tc, const float* A, const float* W, const float* bias,
                         const float* gate, const float* resid, float* out,
                         int64_t M, int K, int N, int hw, int swish);

/* Tuning hook: force the K1 (fused expand+depthwise) tile plan of block `block` (2..16): output tile
 * th x tw, strips of r outputs, cc expanded channels per chunk, nt threads per CTA (256 or 512), nb crops
 * per CTA (2 only where one tile is the whole image).  WHENET_EINVAL if the plan cannot run. */
int whenet_debug_set_k1_plan(whenet_ctx* ctx, int block, int th, int tw, int r, int cc, int nt, int nb);

/* Decode only: n x 252 host logits -> n x 3 host angles through the SAME device function the head kernel uses
 * (softmax of reference utils.py:7-11, expectation of whenet.py:31-33).  Synchronous. */
int whenet_debug_decode(whenet_ctx* ctx, const float* logits_host, int n, float* angles_host);

/* A device kernel raises the context's mbarrier-timeout flag (what a tcgen05 kernel does when a bounded wait expires):
 * the next synchronising call (host-output forward, whenet_synchronize) must return WHENET_ECUDA. */
int whenet_debug_raise_timeout(whenet_ctx* ctx);

/* Where the warp roles of the K1W kernel waited: set option "k1w_trace" to a block index (2..16), run a forward, then read
 * n_rows x 16 int64 cycle counters (one row per CTA): [0] CTA cycles; producer [1] wait a_empty [3] busy span; MMA [4] wait
 * a_full [5] wait t_empty; epilogue warp 0 [7] wait t_full [8] wait e_empty [9] span; depthwise warp 0 [10] wait e_full
 * [11] group barrier [12] span. */
int whenet_debug_read_trace(whenet_ctx* ctx, int64_t* out, int n_rows);

/* Same for K1W (option "k1_variant" = 4, every block with an expand conv): output tile, strip length, channels per CTA,
 * crops per item (2 only where one tile is the whole image), epilogue warps (4 or 8), threads per CTA (768). */
int whenet_debug_set_k1w_plan(whenet_ctx* ctx, int block, int th, int tw, int r, int cc, int nb, int n_epi, int nt);

/* Time every kernel of the NEXT forwards with CUDA events. */
int whenet_profile_enable(whenet_ctx* ctx, int enable);
/* Read (and reset) the accumulated per-kernel statistics; returns the count written. */
int whenet_profile_read(whenet_ctx* ctx, whenet_kernel_stat* out, int cap);

/* Kernels launched by this context since creation (the bench's gpu_launches). */
int64_t whenet_launch_count(whenet_ctx* ctx);

/* Tuning / ablation switches (all have measured defaults; see DESIGN.md and profiles/README.md):
 *   "chunk"          crops per pass through the network (default max_batch: one pass)
 *   "streams"        1..4 batch parts running concurrently on internal streams (default 2)
 *   "graph"          1: replay device-resident forwards from a captured CUDA graph (default 0)
 *   "tensor_cores"   16-bit modes: 0 = CUDA-core kernels for every 1x1 conv, 1 = tcgen05 (default 1).
 *                    fp32 mode: 0 = fp32 FMA kernels (default), 1 = the 1x1 convs on tcgen05 through the bf16 hi/lo split
 *                    (three MMAs per product, fp32 accumulation: 6e-4 deg from the float64 oracle on the golden crops)
 *   "fused"          1: K1 (expand + depthwise fused, expanded tensor in shared memory) for blocks 2..fused_max_block
 *   "k1_variant"     1 = K1 (one tile per CTA), 4 = K1W (weight-stationary persistent CTAs, TMA input tiles, warp roles)
 *   "pw_variant"     2 = pw_tc2 (one tile per CTA, cp.async ring), 3 = K2 (persistent, TMA, warp-specialised), 4 = per layer (default):
 *                    K2 for the ungated / small-map convs that have at least 2 x 148 tiles, pw_tc2 otherwise
 *   "kd_from"        bf16: blocks >= this (default 7) run expand GEMM (fp16 E) + KD (depthwise + squeeze over TMA tiles) instead of K1;
 *                    0 = K1 on every block.  "kd_tail" 1: KD computes the SE gate and gates its output itself (default 0: se_gate +
 *                    gated project); "kd_expand_k2" 0: that expand GEMM on pw_tc2; "dw1_kd" 0: block 1's depthwise on K1's depthwise
 *                    half over a bf16 stem output (default 1: KD over an fp16 stem output); "pw3" 0: block-1 project on pw_tc2
 *   "se_batch", "head_batch"   batches >= 64: four crops per CTA in the SE gate / in the Dense + decode head (default 1; same bits)
 *   "stem_tc"        bf16, uint8 input: 1 = the stem as an im2col GEMM on the tensor core (default 0: not faster)
 *   "stage_threads"  host threads that stage PAGEABLE inputs of 8 MB and more into the context's pinned buffer (default 8; 0 = plain
 *                    cudaMemcpyAsync from the caller's buffer)
 *   "fused_max_block", "dw1_fused", "k1_split_ctas", "k1w_trace" (block whose K1W launch records its role waits),
 *   "se_fused", "se_tail" (K1 CTAs that hold whole crops compute the SE gate themselves, default 1), "se_scale_out", "se_wide",
 *   "pw_stage_cap", "pw_smem_kb", "pw_min_ctas" (split N until the grid has this many CTAs),
 *   "dw_variant", "stem_variant", "host_chunk".
 * Unknown keys return WHENET_ENOTFOUND. */
int whenet_set_option(whenet_ctx* ctx, const char* key, int value);

const char* whenet_last_error(void);
const char* whenet_version(void);
void whenet_destroy(whenet_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* WHENET_B200_H */
