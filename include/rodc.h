/*
 * rodc.h -- C ABI of librodc.so: the B200 (sm_100a) cascade-inference hot path of
 * Johnson145/RapidObjectDetectionUsingCascadedCNNs.
 *
 * The reference is pure Python; it has no FFI of its own.  Each entry point below
 * replaces the third-party native call (or Python loop) the reference makes at the
 * cited file:line, so a maintainer can bind it with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; every `*_dev` / device pointer is caller-owned
 *     (PyTorch allocates); `stream` is a cudaStream_t passed as void*.
 *   - return 0 on success, <0 on error; the message is in rodc_last_error()
 *     (thread-local).  No exceptions cross the ABI, no CPU fallback exists.
 *   - the library keeps no global mutable state; the only allocations it owns are
 *     the per-handle tables of rodc_plan (resampling coefficients, window table)
 *     and rodc_stage (prepared weights), freed by the matching *_destroy.
 */
#ifndef RODC_H
#define RODC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RODC_VERSION 100

#define RODC_OK            0
#define RODC_ERR_ARG      -1
#define RODC_ERR_CUDA     -2
#define RODC_ERR_NOMEM    -3
#define RODC_ERR_UNSUP    -4

typedef struct rodc_plan  rodc_plan;    /* pyramid plan + window grid of one frame shape */
typedef struct rodc_stage rodc_stage;   /* one stage net (weights prepared on the device) */

int         rodc_version(void);
const char* rodc_last_error(void);
/* SM count and compute capability of the current device (fails loudly without a GPU). */
int         rodc_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------- planner (host only)
 * Replaces the scale loop and grid of Window.extract_windows
 * (data/rectangles.py:338-425) and the *_norm properties (:248-262); float64
 * arithmetic in the reference's exact order.  Level sizes follow
 * scipy.misc.imresize(arr, ratio) (data/image_info.py:240). */
int     rodc_plan_create(int img_w, int img_h, int win, double factor, double min_window_length,
                         rodc_plan** out);
void    rodc_plan_destroy(rodc_plan* p);
int     rodc_plan_n_levels(const rodc_plan* p);
int64_t rodc_plan_n_windows(const rodc_plan* p);
/* 1 if every window lies inside its level; 0 reproduces the ValueError of Window.raw
 * (data/rectangles.py:286-289) -> the image is skipped (abstract_inference_app.py:165-183). */
int     rodc_plan_fits(const rodc_plan* p);
/* arrays of n_levels entries each (any may be NULL) */
int     rodc_plan_levels(const rodc_plan* p, double* scale, int32_t* width, int32_t* height,
                         int32_t* step_x, int32_t* step_y, int32_t* nx, int32_t* ny,
                         int64_t* first_window, int64_t* byte_offset);
/* arrays of n_windows entries (norm: [n_windows,4] = xmin,ymin,xmax,ymax in the original) */
int     rodc_plan_windows(const rodc_plan* p, int32_t* level, int32_t* x, int32_t* y, int32_t* norm);
int64_t rodc_plan_pyramid_bytes(const rodc_plan* p);   /* packed levels of ONE frame */
int64_t rodc_plan_scratch_bytes(const rodc_plan* p);   /* horizontal-pass rows of ONE frame */

/* ---------------------------------------------------------------- K1 pyramid
 * Replaces ImageInfo.raw_scaled -> scipy.misc.imresize -> PIL resize(BILINEAR)
 * (data/image_info.py:201-250), bit-exact: two passes, 22-bit fixed point, u8 rounding
 * after each pass, every level from the ORIGINAL frame.
 *   frames  [n_frames, H, W, 3] u8      pyramid [n_frames, pyramid_bytes] u8
 *   scratch [n_frames, scratch_bytes] u8 */
int rodc_pyramid(rodc_plan* p, const uint8_t* frames, int n_frames, uint8_t* pyramid,
                 uint8_t* scratch, void* stream);

/* Host frames -> one flat device buffer: one cudaMemcpyAsync per frame (frame i: nbytes[i] bytes from
 * host_frames[i] to dev_base + dev_off[i]) issued from ONE call, so that a batch of hundreds of small frames costs
 * one trip through the binding instead of one per frame.  The copies are asynchronous only for page-locked host
 * memory (cudaHostAlloc / cudaHostRegister); the reference has no counterpart (it keeps images on the host,
 * data/image_info.py:197-213). */
int rodc_upload_frames(const void* const* host_frames, const int64_t* nbytes, const int64_t* dev_off, int n,
                       uint8_t* dev_base, void* stream);

/* Raw stage-0 windows, order frame -> level -> x -> y (data/rectangles.py:291,409):
 *   windows [n_frames * n_windows, win, win, 3] u8 */
int rodc_extract_windows(rodc_plan* p, const uint8_t* pyramid, int n_frames, uint8_t* windows,
                         void* stream);
/* Window metadata rows (frame_id, xmin_norm, ymin_norm, xmax_norm, ymax_norm):
 *   meta [n_frames * n_windows, 5] i32, frame ids start at frame_id_base */
int rodc_window_meta(rodc_plan* p, int n_frames, int frame_id_base, int32_t* meta, void* stream);
/* GATHER rows of the same windows: byte offset (pyramid_base_off + frame * pyramid_bytes + level offset +
 * (y * level_w + x) * 3) and row pitch (level_w * 3), so that windows of several frame shapes
 * (several plans, several pyramids in one buffer) can share one stage launch -- the reference's merge
 * mode concatenates the windows of all images the same way (abstract_inference_app.py:66-81).
 *   win_off [n_frames * n_windows] i64, win_pitch [n_frames * n_windows] i32 */
int rodc_window_gather_rows(rodc_plan* p, int n_frames, int64_t pyramid_base_off, int64_t* win_off,
                            int32_t* win_pitch, void* stream);

/* Level index (+ slot_base) of every stage-0 window, same order as rodc_window_meta:
 *   out [n_frames * n_windows] i32.  It is the window's key into a rodc_scaled_slot table. */
int rodc_window_levels(rodc_plan* p, int n_frames, int slot_base, int32_t* out, void* stream);
/* out[i] = in[idx[i]] for rows of row_bytes (multiple of 4): carries per-window bookkeeping (level slots,
 * GATHER rows) through the survivor lists that rodc_compact returns, like the reference's
 * windows_info_new.append(windows_info[window_index_original]) (inference_cascade_app.py:219-240). */
int rodc_select_rows(const void* in, const int32_t* idx, int64_t n, int row_bytes, void* out, void* stream);

/* ---------------------------------------------------------------- resize-then-crop branch
 * cf cascade_scale_patches_individually=False (inference_cascade_app.py:374-394): the next stage's input
 * of a survivor is a size x size crop at (int(xmin_norm * next_scale), int(ymin_norm * next_scale)) of the
 * WHOLE frame resized to next_scale = window.scale * size / window.width (image.raw_scaled ->
 * scipy.misc.imresize, data/image_info.py:201-250).  rodc_plan_create_scales + rodc_pyramid build those
 * frames (same integer-exact Pillow resampler as the stage-0 pyramid); rodc_window_scaled_rows turns the
 * survivors into GATHER rows over them. */
int rodc_plan_create_scales(int img_w, int img_h, int n_scales, const double* scales, rodc_plan** out);
typedef struct rodc_scaled_slot {
    double  scale;          /* next_scale of this (frame group, level)                                   */
    int64_t base_off;       /* byte offset of the level of the group's FIRST frame from gather_base      */
    int64_t frame_stride;   /* bytes between consecutive frames of the group (0: a single frame)         */
    int32_t width, height;  /* size of the resized frame                                                 */
    int32_t frame_base;     /* frame id (meta column 0) of the group's first frame                       */
    int32_t individually;   /* 1: this window keeps the crop + cv2.resize input (iff_not_cached option): */
                            /*    row i points at dense_rel_off + i * size*size*3, pitch size*3          */
} rodc_scaled_slot;
/*   meta [n,5] i32, slot [n] i32 (index into slots), slots: DEVICE array of n_slots entries,
 *   win_off [n] i64 / win_pitch [n] i32 out, err: DEVICE i32, bit 0 set when a crop leaves its resized
 *   frame (Window.raw raises ValueError, data/rectangles.py:286-289), bit 1 on a bad slot index. */
int rodc_window_scaled_rows(const int32_t* meta, const int32_t* slot, int64_t n, const rodc_scaled_slot* slots,
                            int n_slots, int size, int64_t dense_rel_off, int64_t* win_off, int32_t* win_pitch,
                            int32_t* err, void* stream);

/* ---------------------------------------------------------------- K4 recrop
 * Replaces _get_increased_raw_window_for_next_net (inference_cascade_app.py:363-373):
 * cv2.resize(window.raw_norm, (size,size)), INTER_LINEAR 8U fixed point incl. the 2x
 * INTER_AREA shortcut, bit-exact.
 *   frames_base + frame_off[f] -> [frame_hw[f][0], frame_hw[f][1], 3] u8
 *   meta [n,5] as above        out [n, size, size, 3] u8
 * The kernel fetches taps with aligned 8-byte loads: it may read up to 7 bytes around a frame byte inside
 * that byte's aligned 8-byte word, i.e. never outside the 256-byte allocation granule of device memory. */
int rodc_recrop(const uint8_t* frames_base, const int64_t* frame_off, const int32_t* frame_hw,
                const int32_t* meta, int64_t n, int size, uint8_t* out, void* stream);

/* ---------------------------------------------------------------- K2 stage net
 * Replaces session.run of one frozen stage graph (inference_cascade_app.py:196-215,
 * inference_app.py:156-166; graph = network/net.py:101-240):
 * conv3x3(3->C)+b+ReLU -> maxpool3x3/1 SAME -> NHWC flatten -> fc1+b+ReLU ->
 * concat(prev bottleneck) -> fc2+b -> softmax.  Weights in the reference's layouts:
 * conv HWIO [3,3,3,C], fc1 [in*in*C, fc1_size], fc2 [fc1_size + n_prev, 2]. */
int  rodc_stage_create(int in_size, int conv_channels, int fc1_size, int n_prev,
                       const float* conv_w, const float* conv_b,
                       const float* fc1_w, const float* fc1_b,
                       const float* fc2_w, const float* fc2_b, rodc_stage** out);
/* General stage architecture (network/net.py:189-221): n_conv conv layers with channels[l] filters of
 * ksize x ksize, stride conv_stride, SAME padding, each followed by ReLU and a SAME max-pool of pool_size /
 * pool_stride; weights conv_w[l] HWIO [k,k,Cin,Cout].  The reference default (one 3x3x32 conv, 3x3/1 pool) is
 * routed to the specialised kernels; everything else runs a generic fp32 kernel (correctness path). */
int  rodc_stage_create_general(int in_size, int n_conv, const int32_t* channels, int ksize, int conv_stride,
                               int pool_size, int pool_stride,
                               const float* const* conv_w, const float* const* conv_b,
                               int fc1_size, int n_prev, const float* fc1_w, const float* fc1_b,
                               const float* fc2_w, const float* fc2_b, rodc_stage** out);
void rodc_stage_destroy(rodc_stage* s);

#define RODC_SRC_DENSE   0   /* windows already materialised: dense [n,S,S,3] u8            */
#define RODC_SRC_PYRAMID 1   /* stage 0: windows inside the packed pyramid (K1 output)       */
#define RODC_SRC_RECROP  2   /* stages >=1: crop + cv2.resize from the original frames       */
#define RODC_SRC_GATHER  3   /* stage 0 over frames of DIFFERENT shapes: per-window byte offset + */
                             /* row pitch into one buffer holding several packed pyramids          */
typedef struct {
    int32_t          kind;
    const uint8_t*   dense;       /* DENSE                                                   */
    const rodc_plan* plan;        /* PYRAMID: window w -> frame w / n_windows                */
    const uint8_t*   pyramid;
    const uint8_t*   frames_base; /* RECROP (see rodc_recrop)                                */
    const int64_t*   frame_off;
    const int32_t*   frame_hw;
    const int32_t*   meta;
    const uint8_t*   gather_base; /* GATHER: window w starts at gather_base + win_off[w],    */
    const int64_t*   win_off;     /*         rows are win_pitch[w] bytes apart               */
    const int32_t*   win_pitch;
} rodc_window_source;

#define RODC_IMPL_TENSOR 0   /* tcgen05 split-fp16 GEMM (3 MMAs, fp32-grade)                 */
#define RODC_IMPL_SIMT   1   /* fp32 CUDA-core GEMM (cross-check implementation)             */

/* The u8 input of a stage for any window source: out [n,size,size,3] u8 = the reference's windows_raw
 * slice fed to X (inference_cascade_app.py:200-202).  rodc_stage_forward runs the same kernels per chunk
 * into its workspace before the streaming feature kernel (the generic feature kernels read every source in
 * place); for RODC_SRC_PYRAMID it equals rodc_extract_windows, for RODC_SRC_RECROP rodc_recrop. */
int rodc_source_windows(const rodc_window_source* src, int64_t first, int64_t n, int size, uint8_t* out,
                        void* stream);

/* Bytes of workspace rodc_stage_forward wants for n windows: per window of a chunk 4*K bytes of features
 * (two fp16 planes or one fp32 plane) + size*size*3 bytes of materialised u8 input; smaller workspaces
 * are legal and make the call run in more chunks. */
int64_t rodc_stage_workspace_bytes(const rodc_stage* s, int64_t n, int impl);
/* Forward of windows [first, first+n) of `src`.
 *   bottleneck [n, ld_bott] f32: columns [0,fc1_size) are written; columns
 *       [fc1_size, fc1_size+n_prev) must already hold the gathered previous bottleneck
 *       (written by rodc_compact) when n_prev > 0.
 *   logits, probs [n,2] f32 (probs = "final_softmax"). */
int rodc_stage_forward(const rodc_stage* s, const rodc_window_source* src, int64_t first, int64_t n,
                       float* bottleneck, int ld_bott, float* logits, float* probs,
                       void* workspace, int64_t workspace_bytes, int impl, void* stream);
/* Device-side counts ("_dyn" entry points).  n is an UPPER BOUND that sizes buffers and grids; the number of
 * windows actually processed is min(n, *n_dev), read on the device when the kernels run (n_dev = the count_out of
 * the previous stage's rodc_compact_dyn; NULL = n).  Rows past the count are neither read nor written.  With
 * these three calls a whole cascade is enqueued without a host round trip per stage and can be captured in a
 * CUDA graph (the reference learns every survivor count on the host, inference_cascade_app.py:219-247).
 * Tensor fc1 only. */
int rodc_stage_forward_dyn(const rodc_stage* s, const rodc_window_source* src, int64_t first, int64_t n,
                           const int64_t* n_dev, float* bottleneck, int ld_bott, float* logits, float* probs,
                           void* workspace, int64_t workspace_bytes, int impl, void* stream);
int rodc_compact_dyn(const float* probs, int64_t n, const int64_t* n_dev, double threshold,
                     const int32_t* meta_in, int32_t* meta_out,
                     const float* bott_in, int ld_in, int width, float* bott_out, int ld_out, int col_off,
                     const float* conf_acc_in, float* conf_out, int conf_mode,
                     int32_t* keep_idx_out, int64_t* count_out, void* scratch, void* stream);
int rodc_group_rectangles_dyn(const int32_t* cand, int64_t k, const int64_t* k_dev, int n_frames, int min_neighbors,
                              double eps, int32_t* out, int64_t* out_count, void* scratch, void* stream);

/* Test hook: pooled conv features of the windows, fp32 [n, in*in*C] (NHWC flatten order). */
int rodc_stage_features_f32(const rodc_stage* s, const rodc_window_source* src, int64_t first,
                            int64_t n, float* features, void* stream);

/* ---------------------------------------------------------------- K3 compaction
 * Replaces the per-row Python filter (inference_cascade_app.py:219-247): keep window w iff
 * (double)probs[w,1] > threshold (strict; threshold = NaN selects the single-net rule
 * argmax != background, i.e. probs[w,1] > probs[w,0], inference_app.py:140-149), order preserved; gathers metadata rows, the
 * bottleneck rows (into columns [col_off, col_off+width) of bott_out) and the confidence
 * (mode 0 = p, 1 = acc + p, 2 = acc * p).  count_out (device int64) receives the number
 * kept; keep_idx_out [<=n] i32 the kept positions.  scratch: rodc_compact_scratch_bytes(n). */
int64_t rodc_compact_scratch_bytes(int64_t n);
int rodc_compact(const float* probs, int64_t n, double threshold,
                 const int32_t* meta_in, int32_t* meta_out,
                 const float* bott_in, int ld_in, int width, float* bott_out, int ld_out, int col_off,
                 const float* conf_acc_in, float* conf_out, int conf_mode,
                 int32_t* keep_idx_out, int64_t* count_out, void* scratch, void* stream);

/* ---------------------------------------------------------------- K5 group rectangles
 * Replaces cv2.groupRectangles(rects, min_neighbors, eps=0.2) per image
 * (inference_app.py:187-212): union-find on the SimilarRects predicate, per-class
 * float32 mean + cvRound, size filter, nested filter; classes in order of first member.
 *   cand [K,5] i32 (frame_id, xmin,ymin,xmax,ymax), grouped by ascending frame id
 *   out  [K,6] i32 (frame_id, x, y, w, h, weight), frames ascending, class order inside
 *   out_count: device int64.  scratch: rodc_group_scratch_bytes(K, n_frames). */
int64_t rodc_group_scratch_bytes(int64_t k, int n_frames);
int rodc_group_rectangles(const int32_t* cand, int64_t k, int n_frames, int min_neighbors, double eps,
                          int32_t* out, int64_t* out_count, void* scratch, void* stream);

/* ---------------------------------------------------------------- launch accounting + measurement hook
 * rodc_launch_count: kernel launches the library has issued since it was loaded (monotonic, process wide).
 * bench.py reports the difference of two reads around its timed region as "gpu_launches".
 *
 * rodc_profile_enable(level): 0 = off (default); 1 = every fc1 launch (the dominant kernel) is bracketed by
 * cudaEvents on the stream it is launched on; 2 = every kernel class is.  rodc_profile_collect synchronises
 * the recorded events and returns the summed device time, the algorithmic FLOPs (2*m*K*N per launch) and the
 * launch count of fc1 since the last collect -- once over all fc1 launches and once over the launches of the
 * dominant shape (largest FLOP count, i.e. full stage-2 chunks), which is what the roofline is quoted on.
 * rodc_profile_table then returns one row per (kernel class, window side) seen by that collect: summed device
 * ms, algorithmic bytes and FLOPs by SURVEY.md §8(d)'s definitions, kernel launches and recorded scopes;
 * the return value is the number of rows written (<= max_rows).  rodc_kernel_class_name names a class.
 * The counters and the record list are the library's only process-wide state. */
int64_t     rodc_launch_count(void);
const char* rodc_kernel_class_name(int kernel_class);
int rodc_profile_enable(int level);
int rodc_profile_collect(double* fc1_ms, double* fc1_flops, int64_t* fc1_launches,
                         double* dom_ms, double* dom_flops, int64_t* dom_launches);
int rodc_profile_table(int max_rows, int32_t* kernel_class, int32_t* window_side, double* ms, double* bytes,
                       double* flops, int64_t* launches, int64_t* records);

#ifdef __cplusplus
}
#endif
#endif /* RODC_H */
