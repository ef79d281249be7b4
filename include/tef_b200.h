/*
 * tef_b200.h -- C ABI of the B200 (sm_100a) contrast-maximization library.
 *
 * The reference (tudelft/taming_event_flow) is pure Python/PyTorch and has no FFI;
 * the boundary it offers is the Python surface of utils/iwe.py, loss/flow.py and
 * dataloader/encodings.py.  Each entry point below names the
 * reference function it replaces; the Python host mirror in
 * taming_event_flow_b200/ binds them with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name says host; buffers are
 *     owned and allocated by the caller, nothing is allocated inside;
 *   - all arithmetic is fp32; tensors are contiguous, row-major, laid out as the
 *     reference's tensors unless stated;
 *   - image and flow-gradient sums use red.global.add.f32, which sm_100a implements as REDG...F32.FTZ: a
 *     DENORMAL partial sum or addend is flushed to zero where the CPU reference keeps it.  Weights are products
 *     of two factors >= 2^-24, so an addend is >= 2^-48 and never denormal; flow-gradient addends can be, at
 *     magnitudes (< 1.2e-38) far below the 1e-5 norm-relative tolerance.  The deterministic mode (integer
 *     reductions) has no flush;
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous;
 *   - return value: 0 on success, a positive cudaError_t from the launch, or a
 *     negative TEF_E* argument error.  tef_strerror() explains both.
 */
#ifndef TEF_B200_H
#define TEF_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define TEF_MAX_PASSES 31   /* passes_loss (after the mode-"four" doubling); in-image bits fit 32 bits */
#define TEF_MAX_SCALES 6    /* scales_loss                                            */
#define TEF_MAX_FLOWS 8     /* flow maps per pass (RecEVFlowNet emits 4)              */

#define TEF_EINVAL (-1)     /* bad size / null pointer                                */
#define TEF_ELIMIT (-2)     /* configuration beyond the static limits above           */
#define TEF_EEMPTY (-3)     /* empty window list: reference raises RuntimeError (torch.cat of []) */
#define TEF_EMODE4 (-4)     /* mode "four" + border compensation: reference raises TypeError       */

int tef_version(void);
const char *tef_strerror(int code);

/* ------------------------------------------------------------------------- */
/* Fused CM loss (loss/flow.py: Iterative :415-746, Linear :216-412)          */
/* ------------------------------------------------------------------------- */
typedef struct tef_cm_desc {
    int B, H, W;           /* batch, resolution (config["loader"]["resolution"])      */
    int P;                 /* passes in the loss window = max(passes_loss)            */
    int F;                 /* number of flow maps per pass (len(flow_list))           */
    int S;                 /* config["data"]["scales_loss"]                           */
    int mode;              /* iterative_mode: 1 one, 2 two, 4 four (ignored by Linear) */
    int border_comp;       /* BaseEventWarping.border_compensation                    */
    int loss_scaling;      /* BaseEventWarping.loss_scaling                           */
    int deterministic;     /* 0: fp32 vector reductions; 1: 64-bit fixed-point integer reductions -- order-independent, bit-reproducible.
                              img then holds two int64 words per value (high 2^-40, low 2^-88: every fp32 addend exactly; four times
                              the bytes), gflow one int64 per value at 2^-40 relative to a power-of-two scale (twice the bytes) */
    /* staged events, set 0 = with gradient, set 1 = detached; pass t holds [B][n] rows:
       ev = float4 (ts + t, y, x, p), mk = float2 (pos, neg).  Written by tef_stage_events. */
    const void *ev[2][TEF_MAX_PASSES];
    const void *mk[2][TEF_MAX_PASSES];
    int n[2][TEF_MAX_PASSES];
    const void *flow;      /* packed flow maps, float2 (x, y), dual-phase and zero-padded: [F][P][B][phase][H+1][Wp]
                              (written by tef_pack_flow / tef_update_pass; Wp = (W+3)&~1)  */
    void *gflow;           /* gradient of the packed maps, dual-phase: [F][P][B][phase][H][Wp] float2 (backward) */
    void *img;             /* accumulation images, float2 (count, time-weighted): [F][B][slots][phase][pol][H][Wp];
                              pixel x lives at column x of phase 0 plus column x+1 of phase 1 (see csrc/tef_cm_common.cuh);
                              after backward the planes hold (dL/dcount, dL/dtime-weighted) in both phases       */
    double *acc_sum;       /* [F][B][slots][chunks] partial sums of squared normalised timestamps (fixed-order reduction) */
    int *acc_nnz;          /* [F][B][slots][chunks] partial counts of pixels with at least one event */
    float *den;            /* [F][B][slots] nnz + 1e-9 (or 1), followed by 2 floats: scale and 1/scale of the deterministic mode's
                              flow-gradient words (written by the backward call)                                  */
    float *loss;           /* [1] scalar loss                                          */
    const float *grad_out; /* [1] upstream gradient of the loss (backward)             */
    /* workspace written by the forward call and read by the backward call (sizes from tef_cm_sizes) */
    void *sort_bins;       /* int [nbins + 1]  tile-sort histogram / offsets (16-byte aligned) */
    void *sort_sums;       /* int scan scratch                                         */
    void *sorted_ev;       /* 32-byte records [rows]: (ts, y, x, sample index bits, mask+, mask-, 0, 0), tile-sorted */
    void *posbuf;          /* float2 [F][P+1][rows_grad] chain positions (x, y) (Iterative) */
    void *alivebuf;        /* uint32 [F][rows_grad] cumulative in-image bits (bit tref) */
    void *gimg;            /* deterministic mode only: gradient images float2 [F][B][slots][phase][pol][H][Wp] */
    int hist_done;         /* 1: every tef_update_pass of this window already counted its events into sort_bins (fused histogram),
                              the forward call then skips its own histogram pass */
    int reserved_;
    /* optional quad-cell copies (NULL: not used).  The 2x2 neighbourhood of an in-image position is one 32-byte cell
       {row y0: (left, right), row y0+1: (left, right)} of float2 pixels, stored once per parity (x0 & 1, y0 & 1):
       [py*2+px][H/2+1][W/2+1] cells per map -- a bilinear sample / corner quad is a single 256-bit gather (one sector) */
    const void *flowq;     /* [F][P][B] quad-cell flow maps, written by tef_update_pass (packedq); Iterative only       */
    void *gimgq;           /* [F][B][slots][pol] quad-cell gradient images, written by tef_iterative_backward; sizes: out[11] */
} tef_cm_desc;

/* buffer sizes for the events currently described by `d`; out[13] =
   { slots, floats in img, floats in gflow, ints in sort_bins, ints in sort_sums, rows of sorted_ev,
     gradient-carrying rows, floats in posbuf, padded row length Wp, chunks per image (acc_sum / acc_nnz),
     floats in gimg, floats in gimgq (Iterative, non-deterministic; else 0), floats in flowq (likewise) } */
int tef_cm_sizes(const tef_cm_desc *d, int linear, long *out);
/* number of image slots (temporal scale x sub-window x reference time, loss/flow.py:657-668) of the configuration in `d`,
   or the negative TEF_E* code the forward call would return for it */
int tef_cm_num_slots(const tef_cm_desc *d, int linear);

/* Iterative.update / Linear.update, event part (loss/flow.py:456-473, :246-263):
   ts += pass_index IN PLACE in the caller's [B,n,4] tensor, then (ts | ts_override), y, x, p
   and the mask are copied into the staging buffers.  ts_override: device scalar for
   round_ts (loss/flow.py:461-463) or NULL. */
int tef_stage_events(void *events_inout, const void *pol_mask, void *ev_out, void *mk_out,
                     long rows, float pass_index, const float *ts_override, void *stream);

/* update_base (loss/flow.py:46-66): one pass of F flow maps [B,2,H,W] (ch0 = x, ch1 = y)
   interleaved into the packed buffer at pass `t`. */
int tef_pack_flow(const void *const *flow_maps_host, int F, int t, int P, int B, int H, int W,
                  void *packed, void *stream);
/* backward counterpart: packed gradient -> [P][F][B][2][H][W] */
/* det_scale: deterministic mode only -- the power-of-two scale of the fixed-point words, den[F*B*slots] of the backward call
   (NULL = 1; ignored otherwise) */
int tef_unpack_flow_grad(const void *packed, void *out, int F, int P, int B, int H, int W, int deterministic, const float *det_scale, void *stream);

/* The whole of Iterative.update / Linear.update in one call and one launch: tef_pack_flow for the F maps of pass `t`
   plus tef_stage_events for the gradient-carrying set [0] and the detached set [1]; optionally the staging kernel also
   builds the histogram of the tile sort, which saves the forward call a pass over the events. */
typedef struct tef_update_desc {
    int F, t, P, B, H, W;
    const void *flow_maps[TEF_MAX_FLOWS];  /* [B][2][H][W] each                                   */
    void *packed;                          /* [F][P][B][2][H+1][Wp] float2                        */
    void *packedq;                         /* optional quad-cell copy [F][P][B][4][H/2+1][W/2+1][8] floats (tef_cm_desc.flowq) or NULL */
    void *events[2];                       /* caller's [B][n][4]; ts += pass_index in place       */
    const void *masks[2];                  /* [B][n][2]                                           */
    void *ev_out[2];                       /* staged rows float4                                  */
    void *mk_out[2];                       /* staged masks float2                                 */
    long rows[2];                          /* B * n                                               */
    float pass_index[2];
    const float *ts_override[2];           /* round_ts device scalars or NULL                     */
    void *sort_bins;                       /* optional fused tile-sort histogram: int [2*P*B*tiles*256 + 1] of the window, where
                                              tiles = ceil(W/16)*ceil(H/8); segment (set k, pass t) owns bins from
                                              (k*P + t)*B*tiles*256 (the layout tef_*_forward expects with hist_done = 1) */
    int hist;                              /* 1: count this pass' events into sort_bins            */
    int zero_bins;                         /* 1: clear sort_bins first (first update of a window)  */
    /* strided[k] = 1: events[k] / masks[k] are NOT contiguous rows; element (b, row, col) of the [B][n][4] event tensor is at
       ev_strides[k][0]*b + ev_strides[k][1]*row + ev_strides[k][2]*col floats (masks likewise, [B][n][2]) -- the transposed
       views the reference's custom_collate returns (dataloader/base.py:414-431); rows[k] must be a multiple of B */
    int strided[2];
    long ev_strides[2][3];
    long mk_strides[2][3];
} tef_update_desc;
int tef_update_pass(const tef_update_desc *u, void *stream);

/* Iterative.forward (loss/flow.py:588-746) and its analytic backward */
int tef_iterative_forward(const tef_cm_desc *d, void *stream);
int tef_iterative_backward(const tef_cm_desc *d, void *stream);

/* Linear.forward (loss/flow.py:306-412) and backward; the per-event flow of Linear.update (:266-285)
   is sampled inside the kernels from the packed map of the event's own pass */
int tef_linear_forward(const tef_cm_desc *d, void *stream);
int tef_linear_backward(const tef_cm_desc *d, void *stream);

/* ------------------------------------------------------------------------- */
/* utils/iwe.py primitives                                                    */
/* ------------------------------------------------------------------------- */
/* event_propagation (utils/iwe.py:5-14): out = loc + (tref - ts) * flow; ts [n], others [n][2] */
int tef_event_propagation(const float *ts, const float *loc, const float *flow, float tref, float *out, long n, void *stream);
int tef_event_propagation_bwd(const float *gout, const float *ts, const float *flow, float tref,
                              float *g_ts, float *g_loc, float *g_flow, long n, void *stream);
/* get_event_flow (utils/iwe.py:17-40): maps [B][H][W], loc [B][N][2] (y,x) -> out [B][N][2] (y,x) */
int tef_get_event_flow(const float *mapx, const float *mapy, const float *loc, float *out, int B, int N, int H, int W, void *stream);
/* g_mapx/g_mapy must be zeroed by the caller; g_loc is written */
int tef_get_event_flow_bwd(const float *gout, const float *mapx, const float *mapy, const float *loc,
                           float *g_mapx, float *g_mapy, float *g_loc, int B, int N, int H, int W, void *stream);
/* purge_unfeasible (utils/iwe.py:43-60); rows of [n][2]; also usable as its own backward (g * in) */
int tef_purge_unfeasible(const float *loc, const float *mask, float *out_loc, float *out_mask, long n, int H, int W, void *stream);
int tef_purge_unfeasible_bwd(const float *loc, const float *g_loc_out, const float *g_mask_out, float *g_loc, float *g_mask, long n, int H, int W, void *stream);
/* get_interpolation (utils/iwe.py:63-113): warped [B][N][2] -> idx, w [B][4N] (corner-major) or [B][N] when round_idx */
int tef_get_interpolation(const float *warped, float *idx, float *w, int B, int N, int H, int W, int round_idx, void *stream);
int tef_get_interpolation_bwd(const float *warped, const float *g_w, float *g_warped, int B, int N, int H, int W, void *stream);
/* interpolate (utils/iwe.py:116-136): iwe [B][H*W] must hold zeros or the `zeros` start image; pol may be NULL */
int tef_interpolate(const float *idx, const float *w, const float *pol, float *iwe, int B, long M, int H, int W, void *stream);
int tef_interpolate_bwd(const float *idx, const float *pol, const float *w, const float *g_iwe, float *g_w, float *g_pol, int B, long M, int H, int W, void *stream);
/* deblur_events (utils/iwe.py:139-224), fused: flow [B][2][H][W], events [B][N][4]; pol may be NULL, else the
   polarity of event i of the whole batch is pol[i * pol_stride] (a [B,N,2] mask column has stride 2);
   sample b is accumulated into iwe + b * iwe_batch_stride (H*W floats, zeroed inside), so compute_pol_iwe
   (utils/iwe.py:227-257) writes its two channels with two calls */
int tef_deblur_events(const float *flow, const float *events, const float *pol, long pol_stride, float *iwe, long iwe_batch_stride,
                      int B, int N, int H, int W, int round_idx, int round_flow, void *stream);

/* ------------------------------------------------------------------------- */
/* dataloader/encodings.py                                                    */
/* ------------------------------------------------------------------------- */
/* `oob` (device int, may be NULL) is set to 1 when an event's truncated (and, if negative, wrapped) coordinates fall outside
   the sensor -- where the reference's index_put_ raises IndexError (:23-27); such events are skipped.  The caller zeroes it. */
/* events_to_image (:8-29): img [H][W] initialised inside.  accumulate=0 is the reference's plain put: the LAST event of a
   pixel wins (the order of its CPU kernel), computed deterministically in two passes (highest event index per pixel first) */
int tef_events_to_image(const float *xs, const float *ys, const float *ps, float *img, long n, int H, int W, int accumulate, int *oob,
                        void *stream);
/* events_to_channels (:59-81): out [2][H][W] */
int tef_events_to_channels(const float *xs, const float *ys, const float *ps, float *out, long n, int H, int W, int *oob, void *stream);
/* batched events_to_channels over the loader's zero-padded [B][N][4] (ts, y, x, p) rows: out [B][2][H][W] */
int tef_events_to_channels_batched(const float *events, float *out, int B, int N, int H, int W, void *stream);
/* get_hot_event_mask: NOT in this reference (north_star names it) -- "parity unpinned".  Implements the
   published routine of tudelft/event_flow's dataloader/encodings.py: with idx > min_obvs, up to max_px times the arg-max of
   event_rate [H][W] is zeroed and masked while it exceeds max_rate.  event_rate is modified in place like there. */
int tef_get_hot_event_mask(float *event_rate, float *mask, int H, int W, int idx, int max_px, int min_obvs, float max_rate, void *stream);
/* events_to_voxel (:32-56): out [bins][H][W] */
int tef_events_to_voxel(const float *xs, const float *ys, const float *ts, const float *ps, float *out, long n, int bins, int H, int W, int *oob,
                        void *stream);

/* ------------------------------------------------------------------------- */
/* dataloader/base.py -- the loader -> loss contract          */
/* ------------------------------------------------------------------------- */
/* create_polarity_mask (:264-278): ps [n] -> mask [2][n] */
int tef_create_polarity_mask(const float *ps, float *mask, long n, void *stream);
/* create_mask_encoding (:302-314): cnt [B][2][H][W] -> out [B][1][H][W] */
int tef_create_mask_encoding(const float *cnt, float *out, int B, int H, int W, void *stream);
/* custom_collate (:391-434) for one sample: src [C][n] -> dst rows [N][C], zero rows n..N (the caller offsets dst per sample) */
int tef_collate_events(const float *src, float *dst, long n, long N, int C, void *stream);
/* event_formatting (:139-170) + create_list_encoding (:247-262) + create_polarity_mask + custom_collate for a ragged batch
   that crossed PCIe as 8-byte packed events {fp32 raw ts, x | y << 14 | pol << 28}: window b = packed[offsets[b] : offsets[b+1]]
   (offsets: B+1 longs in device memory).  Writes event_list [B][n_pad][4] = (ts normalised, y, x, p = pol*2-1), pol_mask
   [B][n_pad][2], zero rows as padding and, when cnt is not NULL, events_to_channels of every window: cnt [B][2][H][W] */
int tef_format_events(const void *packed, const long *offsets, int B, long n_pad, float *event_list, float *pol_mask,
                      float *cnt, int H, int W, void *stream);
/* split_event_list (:347-377) on the padded batch: every sample's events are ranked by a keyed pseudo-random permutation
   (Feistel network, cycle-walked; `seed` picks it); rank < k -> gradient list [B][Ng] at row rank, else detached list
   [B][Nd] at row rank-k; samples with at most k events keep their order and detach nothing.  Outputs are zeroed inside
   (padding rows). */
int tef_split_events(const float *event_list, const float *pol_mask, const long *offsets, unsigned long long seed, int B, long N, long k,
                     float *g_events, float *g_mask, long Ng, float *d_events, float *d_mask, long Nd, void *stream);

/* ------------------------------------------------------------------------- */
/* loss/flow.py:131-209 -- the smoothness priors on the packed   */
/* flow maps written by tef_pack_flow / tef_update_pass: [F][P][B] maps, of which  */
/* the first n passes have been given to update()                                 */
/* ------------------------------------------------------------------------- */
/* floats of scratch either prior needs */
long tef_flow_smoothing_scratch(int B, int H, int W, int n, int F);
/* flow_spatial_smoothing (:170-209): out [B] = per-sample value before the weight (the caller sums and scales) */
int tef_flow_spatial_smoothing(const float *packed_flow, int B, int H, int W, int P, int F, int n, float *scratch, float *out, void *stream);
/* its gradient: gout [B] (device) -> packed_grad, the dual-phase layout tef_unpack_flow_grad reads (zeroed inside) */
int tef_flow_spatial_smoothing_bwd(const float *packed_flow, const float *gout, float *packed_grad, int B, int H, int W, int P, int F, int n,
                                   void *stream);
/* flow_temporal_smoothing (:131-168): out [B]; sums [F][n-1][B][2] (masked sum, valid pixels) is kept for the backward */
int tef_flow_temporal_smoothing(const float *packed_flow, int B, int H, int W, int P, int F, int n, float *scratch, float *sums, float *out,
                                void *stream);
int tef_flow_temporal_smoothing_bwd(const float *packed_flow, const float *sums, const float *gout, float *packed_grad, int B, int H, int W,
                                    int P, int F, int n, void *stream);

/* ------------------------------------------------------------------------- */
/* loss/flow_val.py -- fused stages of the validation update; */
/* batch size 1 like upstream; maps are planar [H][W] (x and y flow separately) */
/* ------------------------------------------------------------------------- */
/* Everything `update` appends, in one launch (update_base :75-114 and, for Iterative, :483-487, :519-528, :558-562): the window's
   events [n][4] (ts += pass_index IN PLACE, like upstream) and mask [n][2] are copied to row `offset` of the caller's row arrays
   (the pointers below already point at that row), the newest flow map [2][H][W] and event mask [H][W] to slot `now` of the
   per-window stacks (the pointers point at that slot).  ts_override: device scalar for round_ts or NULL.  Linear: fw_* / bw_* /
   prop_* are NULL. */
typedef struct tef_val_append {
    void *events; const void *pol_mask; long n; float pass_index; const float *ts_override;
    float *ev_ts; float *ev_loc; float *ev_mask;
    float *fw_ts; float *fw_loc; float *fw_mask;
    float *bw_loc; float *bw_mask;
    const float *flow; const float *event_mask; int H, W;
    float *map_x; float *map_y; float *map_e; float *prop_x; float *prop_y;
} tef_val_append;
int tef_val_append_window(const tef_val_append *d, void *stream);
/* _pol_images of the criteria (:116-131 and the images of FWL / RSAT, :189-274): get_interpolation + optional per-event weight
   `extra` [n] (NULL: none) + one interpolate per polarity, fused: loc [n][2] (y, x), mask [n][2] -> out [2][H][W] (zeroed inside) */
int tef_val_pol_images(const float *loc, const float *mask, const float *extra, float *out, long n, int H, int W, int round_idx, void *stream);
/* Iterative.update :483-517: every accumulated event one window forward with the newest map, in place:
   loc [n][2] (y, x), ts [n] (set to tref), mask [n][2] */
int tef_val_forward_step(const float *mapx, const float *mapy, float *loc, float *ts, float *mask, float tref, long n, int H, int W,
                         void *stream);
/* Iterative.update :519-556: the new window back to time 0 through maps n_maps-1 .. 0 (mapsx/mapsy [n_maps][H][W]); loc and
   mask in place, ts [n] = the window's timestamps (read only) */
int tef_val_backward_chain(const float *mapsx, const float *mapsy, int n_maps, float *loc, const float *ts, float *mask, long n,
                           int H, int W, void *stream);
/* forward_prop_flow :43-74 for n_maps consecutive maps (the i-th has time index first+i): each is carried to time first+i+1
   (tref_is_next) or to `tref` by splatting it along itself.  acc: scratch [n_maps][3][H][W] (zeroed inside); outx/outy
   [n_maps][H][W] may alias the inputs */
int tef_val_forward_prop_flow(const float *mapsx, const float *mapsy, int first, int n_maps, int tref_is_next, float tref, float *acc,
                              float *outx, float *outy, int H, int W, void *stream);
/* Iterative.update :579-605: pixel trajectories through the newest map: idx [2][H][W] (y, x planes) and out_mask [H][W] in
   place, accx/accy [H][W] = accumulated displacement */
int tef_val_trajectory_step(const float *mapx, const float *mapy, float *idx, float *out_mask, float *accx, float *accy, int H, int W,
                            void *stream);

/* ------------------------------------------------------------------------- */
/* The training step around the loss: element-wise stages of   */
/* the recurrent flow network, fused.  The convolutions stay on cuDNN.  NHWC     */
/* ("channels_last") fp32: M = B*H*W pixel rows, channel counts multiples of 4.  */
/* ------------------------------------------------------------------------- */
/* ConvGRU.forward (models/submodules.py:134-152) between its convolutions.  zr [M][2C]: output of the merged update|reset
   convolution on xh, replaced IN PLACE by (z, r) = sigmoid(zr + bias_zr);  xh [M][Cx+C] = (input, previous state);
   xrh [M][Cx+C] = (input, state * r) -- the input of the candidate convolution (:149) */
int tef_gru_gates(float *zr, const float *bias_zr, const float *xh, float *xrh, long M, int Cx, int C, void *stream);
/* c [M][C]: output of the candidate convolution, replaced IN PLACE by cand = tanh(c + bias_c);
   out [M][C] = state * (1 - z) + cand * z (:150) */
int tef_gru_output(float *c, const float *bias_c, const float *xh, const float *zr, float *out, long M, int Cx, int C, void *stream);
/* their reverse, bias gradients included (gbias_* are ADDED to; NULL: skipped).  gout [M][C] ->
   gc [M][C] (gradient of the candidate convolution's output), gzr[:, :C] (update gate), gh [M][C] = gout * (1 - z) */
int tef_gru_output_bwd(const float *gout, const float *cand, const float *xh, const float *zr, float *gc, float *gzr, float *gh, float *gbias_c,
                       float *gbias_zr, long M, int Cx, int C, void *stream);
/* gxrh [M][Cx+C] (gradient of the candidate convolution's input) -> gzr[:, C:] (reset gate), gh += gxrh[:, Cx:] * r */
int tef_gru_gates_bwd(const float *gxrh, const float *xh, const float *zr, float *gzr, float *gh, float *gbias_zr, long M, int Cx, int C,
                      void *stream);
/* gxh [M][Cx+C] (gradient of the gate convolution's input): gx [M][Cx] = gxrh[:, :Cx] + gxh[:, :Cx];  gh += gxh[:, Cx:] */
int tef_gru_input_grads(const float *gxrh, const float *gxh, float *gx, float *gh, long M, int Cx, int C, void *stream);
/* ConvLayer / ResidualBlock (models/submodules.py): y [M][C] = act(y + bias + residual) in place; act 0 none, 1 ReLU, 2 tanh;
   bias / residual may be NULL */
int tef_bias_act(float *y, const float *bias, const float *residual, int act, long M, int C, void *stream);
/* gpre = gy * act'(y) (y = the activated output; gpre may alias gy), gbias [C] += column sums of gpre (NULL: skipped) */
int tef_bias_act_bwd(const float *gy, const float *y, float *gpre, float *gbias, int act, long M, int C, void *stream);
/* Flow head (models/model.py:65-85 with train_flow.py:106-108): bilinear up-sampling (align_corners = False) of the 2-channel
   prediction pred [B][2][h][w] (element strides {batch, channel, row, column}) to out [B][2][H][W] (contiguous), times `scale`
   (the head's 2^k and flow_scaling in one factor); and its adjoint (deterministic gather) */
int tef_upsample_scale(const float *pred, const long *strides, int h, int w, float *out, int B, int H, int W, float scale, void *stream);
int tef_upsample_scale_bwd(const float *gout, int B, int H, int W, float scale, float *gpred, const long *strides, int h, int w, void *stream);

/* Decoder stage input (models/arch.py, the decoder loop of the recurrent U-Net): out [B][H][W][Cp] = bilinear up-sampling
   (align_corners = False; H/h, W/w up to 2.5) of cat(pred, x + skip): x, skip [B][h][w][C] NHWC (skip may be NULL, C even), pred
   [B][2][h][w] through element strides {batch, channel, row, column} or NULL (then Cp = C, else C + 2); and the adjoint:
   gout -> gx [B][h][w][C] (the gradient of x and of skip alike) and gpred (pred's strides; NULL when there was no pred) */
int tef_decoder_up(const float *x, const float *skip, const float *pred, const long *pred_strides, float *out, int B, int h, int w, int C, int H,
                   int W, void *stream);
int tef_decoder_up_bwd(const float *gout, float *gx, float *gpred, const long *pred_strides, int B, int h, int w, int C, int H, int W, void *stream);

/* L2 rate micro-benchmarks (what bounds the CM kernels): kind 0 = red.global.add.v4.f32, 1 = 8-byte gathers, 2 = 16-byte gathers,
   4 = 2x2 neighbourhood fetches of tile-sorted positions (mode 0: two 16-byte gathers in the dual-phase layout, 1: one 32-byte gather);
   mode 0 = uniformly random addresses, 1 = a 4 KB window per warp; buf = `bytes` (power of two) of device memory */
int tef_microbench(int kind, int mode, void *buf, long bytes, int iters, long *ops, void *stream);
/* the shared-memory counterpart (north_star's smem tiles): 16-byte updates accumulated in a CTA-private shared-memory patch with
   shared-memory atomics (atom 0 = fp32 compare-and-swap loop, 1 = native u32 fixed point, 2 = u64 fixed point), flushed with coalesced
   red.v4 every k updates per thread; patch_slots = 16-byte slots per patch; pattern 0 = random slots, 1 = tile-sorted-like */
int tef_microbench_smem(int atom, int patch_slots, int k, int pattern, void *buf, long bytes, int iters, long *ops, void *stream);

/* ------------------------------------------------------------------------- */
/* launch accounting / per-kernel timing (used by bench.py for gpu_launches   */
/* and the roofline; CUDA events are recorded on the launching stream)        */
/* ------------------------------------------------------------------------- */
void tef_prof_enable(int on);
void tef_prof_reset(void);
int tef_prof_num_kernels(void);
const char *tef_prof_name(int id);
int tef_prof_read(int id, double *ms_total, long *timed_launches, long *launches);
long tef_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* TEF_B200_H */
