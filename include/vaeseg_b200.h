/*
 * vaeseg_b200 -- C ABI of the B200-native training hot path of yyNoBug/VAE_segmentation.
 *
 * The reference has no FFI of its own (it is pure PyTorch): each entry
 * point below replaces the library kernel that the cited reference line reaches through
 * torch.nn on the hot path.  Host side (vae_segmentation_b200/*.py) binds these with ctypes.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (PyTorch caching allocator);
 *    the library allocates no device memory.
 *  - activations are NDHWC ("channels last"), dtype selected by `dtype`:
 *      VS_F32  (check mode: fp32 storage, fp32 accumulate)   VS_BF16 (bf16 storage, fp32 accumulate)
 *    tensors called "planar" are NCDHW fp32 (the nn.Module boundary layout).
 *  - parameters and parameter gradients are fp32 in PyTorch's own layouts
 *    (Conv3d [Cout,Cin,kD,kH,kW], ConvTranspose3d [Cin,Cout,2,2,2], Linear [out,in]).
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), performs no
 *    host synchronisation and is CUDA-graph capturable.
 *  - return value: 0 on success, negative vs_status otherwise; vs_last_error_string()
 *    describes the last failure on the calling thread.  There is no CPU fallback.
 */
#ifndef VAESEG_B200_H
#define VAESEG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { VS_OK = 0, VS_ERR_SHAPE = -1, VS_ERR_UNSUPPORTED = -2, VS_ERR_ALIGN = -3, VS_ERR_CUDA = -4 } vs_status;
typedef enum { VS_F32 = 0, VS_BF16 = 1 } vs_dtype;
/* flags: VS_FLAG_PREZEROED = the caller already zeroed the statistics / shift (or sums) words the call would
 * otherwise memset -- the host side zeroes ONE arena per network pass instead of one memset per layer. */
enum { VS_FLAG_PREZEROED = 1 };

const char* vs_last_error_string(void);
int vs_version(void);
/* 1 when the library was built with the tcgen05/TMA (sm_100a) conv kernels */
int vs_has_tcgen05(void);
/* Programmatic dependent launch for the kernels of the hot chain (convolutions, k2s2, InstanceNorm passes): launches
 * of at most `max_ctas` CTAs may be scheduled while their predecessor in the stream drains (the prologue overlaps the
 * predecessor's tail; the kernel waits before touching global memory).  0 disables.  Returns the previous setting. */
int vs_set_pdl(int max_ctas);
/* 1: the kd-in-N convolution issues its accumulating MMAs from one warp in a fixed order, which makes the forward pass
 * bit-reproducible from run to run (with 0 four warps issue concurrently and the fp32 accumulation order, hence the
 * last bit of some outputs, varies).  The reference has no equivalent switch; torch.use_deterministic_algorithms is
 * the nearest notion.  Statistics and weight gradients still use atomics (order differences ~1e-16 / ~1e-7). */
void vs_set_kdn_ordered(int on);

/* ---- weight repacking (derived caches of the fp32 master weights) ------------------- */
/* Conv3d 3x3x3 weight [Cout,Cin,27] -> wf[27][Cin][Cout] (fprop) and, when wd != NULL,
 * wd[27 (flipped)][Cout][Cin] (dgrad operand).  joint_model.py:40,43,46,106,224,366 */
int vs_pack_conv3_weight(const float* w, float* wf, float* wd, int cin, int cout, void* stream);
/* bf16 UMMA B-operand pack of the same weight for the tcgen05 path (dgrad=1: flipped taps, channels swapped).
 * vs_conv3_tc_pack_bytes() returns 0 when the tensor-core path does not take the shape
 * (it takes Cin == 8 or Cin %% 16 == 0, Cout %% 8 == 0).                                                   */
size_t vs_conv3_tc_pack_bytes(int cin, int cout, int dgrad);
int vs_pack_conv3_weight_tc(const float* w, void* out, int cin, int cout, int dgrad, void* stream);

/* All derived packs of many layers in one launch.  jobs_dev: DEVICE array of njobs vs_pack_job; NULL outputs are
 * skipped.  cout_pad > cout zero-pads the output channels of the dgrad tensor-core pack (the 2-class head is
 * run as an 8-channel layer in backward); otherwise cout_pad == cout.  *_elems = vs_conv3_tc_pack_bytes()/2.   */
typedef struct {
    const void* w;        /* fp32 [Cout][Cin][27] master weight                         */
    void* wf;             /* fp32 [27][Cin][Cout] or NULL                               */
    void* wd;             /* fp32 [27 flipped][Cout][Cin] or NULL                       */
    void* tcf;            /* bf16 tensor-core fprop pack or NULL                        */
    void* tcd;            /* bf16 tensor-core dgrad pack (of [cout_pad][Cin][27]) or NULL */
    long long tcf_elems, tcd_elems;
    int cin, cout, cout_pad, cin_pad;   /* cin_pad > cin zero-pads the input channels of the dgrad pack (0 = none) */
    void* kdn;            /* bf16 kd-in-N fprop pack (vs_conv3x3x3_tc_kdn) or NULL      */
    long long kdn_elems;
    int kind;             /* 0: 3x3x3 layer (fields above).  1: 2x2x2 stride-2 layer: w = wt[A][B][8] with A = cout,
                           * B = cin; tcf = gather pack, tcd = scatter pack (vs_k2s2_tc_pack_bytes()/2 elements)   */
    int kdn_dgrad;        /* the kd-in-N pack is the dgrad one (flipped taps, channels swapped)         */
} vs_pack_job;
int vs_pack_conv3_batched(const void* jobs_dev, int njobs, void* stream);
/* One padded pack: master weight [cout][cin][27], pack built for cin_pad >= cin / cout_pad >= cout channels (zeros);
 * out holds vs_conv3_tc_pack_bytes(cin_pad, cout_pad, dgrad) bytes.                                             */
int vs_pack_conv3_weight_tc_padded(const float* w, void* out, int cin, int cout, int cin_pad, int cout_pad, int dgrad,
                                   void* stream);

/* ---- 3x3x3 convolution, padding 1 (Conv3d at joint_model.py:40-46,106,224,366) ------- */
/* x: NDHWC `in_dtype` (or planar fp32 when in_planar=1, used by the in_blocks whose input is
 * the module's NCDHW fp32 tensor); wpk: fp32 [27][Cin][Cout]; wtc: bf16 tensor-core pack or NULL
 * (non-NULL + bf16 NDHWC in/out + no bias -> tcgen05/TMEM/TMA implicit GEMM, else the CUDA-core
 * direct kernel); bias fp32 [Cout] or NULL;
 * y: NDHWC `out_dtype` (or planar fp32 when out_planar=1); stats: fp64 [N][Cout][2]
 * (sum, sum of squares of the fp32 accumulators, accumulated in double because
 * E[x^2]-E[x]^2 cancels in fp32; zeroed by the call) or NULL.            */
/* shift: fp32 scratch [N][Cout] or NULL.  When given, the call first evaluates the convolution
 * at voxel (1,1,1) of every (n, co) and subtracts that constant from the whole output channel
 * before statistics and storage: InstanceNorm is invariant to a per-(n,c) shift, and storing
 * deviations keeps bf16 precision on channels with |mean| >> sigma (VAE layers on masks).     */
int vs_conv3x3x3_fprop(int in_dtype, int out_dtype, int in_planar, int out_planar, int flags,
                       const void* x, const float* wpk, const void* wtc, const float* bias, void* y, double* stats,
                       float* shift, int n, int d, int h, int w, int cin, int cout, void* stream);
/* dgrad is the same contraction with the flipped/transposed pack (wd of vs_pack_conv3_weight):
 * dx = conv3(dy, wd), Cin/Cout swapped.  Provided as its own symbol for the binding's clarity.
 * Fused epilogue (tensor-core path only, else VS_ERR_UNSUPPORTED): when sums_prev != NULL, dx is the gradient w.r.t.
 * the activation of the PREVIOUS Conv3d+InstanceNorm+ReLU layer whose raw output is y_prev (NDHWC bf16, dx's shape)
 * with statistics stats_prev; the call then also accumulates that layer's InstanceNorm-backward sums
 * sums_prev[N][Cin][2] += (sum dx*mask, sum dx*mask*xhat) -- what vs_inorm_relu_bwd_reduce would compute in a
 * separate pass over dx and y_prev.  sums_prev must be zeroed by the caller.
 * out_planar=1 with Cin == 2 and wdtc = vs_pack_conv3_weight_tc_padded(w, 2, cout, 8, cout, dgrad=1): tensor-core
 * path storing the two real input-gradient channels as planar fp32 (the VAE in-block).                          */
int vs_conv3x3x3_dgrad(int in_dtype, int out_dtype, int out_planar,
                       const void* dy, const float* wdpk, const void* wdtc, void* dx,
                       const void* y_prev, const double* stats_prev, double* sums_prev,
                       int n, int d, int h, int w, int cin, int cout, void* stream);
/* 2-class head in one launch: probs[N][2][D][H][W] (planar fp32) = softmax_c(conv3(x, w) + bias).  x: NDHWC bf16,
 * Cin == 8 or Cin %% 16 == 0; wtc8: vs_pack_conv3_weight_tc_padded(w, cin, 2, cin, 8, dgrad=0) (output channels
 * zero-padded to 8); bias fp32 [2] or NULL.  Replaces Conv3d + nn.Softmax(dim=1) at joint_model.py:224-225,366-367. */
int vs_head_conv_softmax2_fwd(const void* x, const void* wtc8, const float* bias, float* probs,
                              int n, int d, int h, int w, int cin, void* stream);
/* "kd-in-N" variant of the tensor-core convolution for GEMM outputs of 8 or 16 channels and D >= 4 (the full-resolution
 * layers, forward and input gradient) -- MMAs issued per input plane with the three kd taps side by side in N, halving the
 * shared-memory operand traffic.  Same contract as the tensor-core branch of vs_conv3x3x3_fprop / _dgrad (bf16 NDHWC in
 * and out, optional shift + statistics; _ex: the fused norm-backward reduction of the previous layer, psums pre-zeroed
 * by the caller), with its own weight pack.  vs_conv3_tc_kdn_pack_bytes() returns 0 for shapes it does not take.   */
size_t vs_conv3_tc_kdn_pack_bytes(int cin, int cout, int dgrad);
int vs_pack_conv3_weight_tc_kdn(const float* w, void* out, int cin, int cout, int dgrad, void* stream);
int vs_conv3x3x3_tc_kdn(const void* x, const void* wkdn, void* y, double* stats, float* shift, int prezeroed,
                        int n, int d, int h, int w, int gin, int gout, void* stream);
int vs_conv3x3x3_tc_kdn_ex(const void* x, const void* wkdn, void* y, double* stats, float* shift, int prezeroed,
                           const void* yprev, const double* pstats, double* psums,
                           int n, int d, int h, int w, int gin, int gout, void* stream);
/* Conv3d(3, padding 1) -> InstanceNorm3d(eps 1e-5, biased, no affine) -> ReLU (+ skip add) in ONE launch
 * (joint_model.py:40-46,106 `Conv` / `DoubleConv`): the kernel is launched cooperatively, accumulates the statistics,
 * passes a grid barrier and then turns the raw outputs it just stored (still in L2) into the activation
 * a = relu((y - mean) * rstd) + skip -- the separate vs_inorm_relu_apply pass disappears.  y (raw, shifted; kept for the
 * backward pass), a and skip are bf16 NDHWC [N,D,H,W,gout]; stats [N][gout][2] fp64, shift [N][gout] fp32 and the
 * 32-bit word gbar must be ZERO on entry.  _tc_: tap-per-MMA kernel (pack of vs_pack_conv3_weight_tc); _tc_kdn_: kd-in-N
 * kernel (pack of vs_pack_conv3_weight_tc_kdn; gout 8 or 16, D >= 4).  Same channel constraints as those kernels. */
int vs_conv3x3x3_tc_in_relu(const void* x, const void* wtc, void* y, void* a, const void* skip, double* stats, float* shift,
                            unsigned* gbar, int n, int d, int h, int w, int gin, int gout, void* stream);
int vs_conv3x3x3_tc_kdn_in_relu(const void* x, const void* wkdn, void* y, void* a, const void* skip, double* stats,
                                float* shift, unsigned* gbar, int n, int d, int h, int w, int gin, int gout, void* stream);
/* kd-in-N pack of a layer whose channels are zero-padded to cin_pad x cout_pad (w is [cout][cin][27]): the 2-class head
 * (cout 2 -> 8) and the 2-channel in-block's input gradient (cin 2 -> 8).  Size: vs_conv3_tc_kdn_pack_bytes(cin_pad, cout_pad, dgrad). */
int vs_pack_conv3_weight_tc_kdn_padded(const float* w, void* out, int cin, int cout, int cin_pad, int cout_pad, int dgrad,
                                       void* stream);
/* Planar fp32 output out[N][2][D][H][W] of GEMM output channels 0..1 of an 8-output-channel kd-in-N convolution:
 * mode 1 = the 2-class head, softmax_c(conv3(x, w) + bias) in one launch (as vs_head_conv_softmax2_fwd;
 * joint_model.py:224-225,366-367); mode 2 = plain values (the planar 2-channel input gradient of the VAE in-block,
 * as vs_conv3x3x3_dgrad with out_planar = 1).  x bf16 NDHWC with gin = 8 or a multiple of 16 channels, D >= 4. */
int vs_conv3x3x3_tc_kdn_planar(const void* x, const void* wkdn8, float* out, const float* bias, int mode,
                               int n, int d, int h, int w, int gin, void* stream);
/* dw[Cout,Cin,27] (+)= sum_v dy[v,co] * x[v+tap,ci]; db[Cout] (+)= sum_v dy (db may be NULL).
 * x may be planar fp32 (in_planar=1).  accumulate=0 overwrites.  workspace: fp32
 * vs_conv3_wgrad_workspace_bytes() bytes (partial sums).                                 */
size_t vs_conv3_wgrad_workspace_bytes(int n, int d, int h, int w, int cin, int cout);
int vs_conv3x3x3_wgrad(int dtype, int in_planar, const void* x, const void* dy, float* dw, float* db,
                       void* workspace, size_t ws_bytes, int accumulate,
                       int n, int d, int h, int w, int cin, int cout, void* stream);

/* ---- 2x2x2 stride-2 convolution and transposed convolution (joint_model.py:118,130) --- */
/* Both layers are one of two contractions over a weight viewed as wt[A][B][8]:
 *   gather : coarse[o, a] = bias[a] + sum_{k,b} wt[a][b][k] * fine[2o+k, b]
 *   scatter: fine[2o+k, b] = bias[b] + sum_a    wt[a][b][k] * coarse[o, a]
 * Conv3d(k2,s2)       : fprop = gather (A=Cout,B=Cin), dgrad = scatter.
 * ConvTranspose3d(k2) : fprop = scatter (A=Cin,B=Cout), dgrad = gather.
 * (dc,hc,wc) are the COARSE spatial dims; the fine tensor is (2dc,2hc,2wc).            */
int vs_k2s2_gather(int dtype, const void* fine, const float* wt, const float* bias, void* coarse,
                   int n, int dc, int hc, int wc, int a, int b, void* stream);
int vs_k2s2_scatter(int dtype, const void* coarse, const float* wt, const float* bias, void* fine,
                    int n, int dc, int hc, int wc, int a, int b, void* stream);
/* The same two contractions on the tensor cores (tcgen05.mma / TMEM / TMA; bf16 NDHWC in and out; A, B in {8, 16k},
 * <= 256): GEMM over the coarse voxels, M tile = 16 x 8 voxels of a coarse d-plane; gather K = (kd,kh,kw,b) read
 * through a [2B][Wc][kh][Hc][n*Df] view of the fine tensor, scatter N = (kd,kh,kw,b) stored as contiguous (kw,b)
 * runs.  wpack: bf16 UMMA B-operand pack of wt (scatter = 0 / 1), vs_k2s2_tc_pack_bytes() bytes; bias may be NULL. */
size_t vs_k2s2_tc_pack_bytes(int a, int b, int scatter);
int vs_pack_k2s2_weight_tc(const float* wt, void* out, int a, int b, int scatter, void* stream);
int vs_k2s2_gather_tc(const void* fine, const void* wpack, const float* bias, void* coarse,
                      int n, int dc, int hc, int wc, int a, int b, void* stream);
int vs_k2s2_scatter_tc(const void* coarse, const void* wpack, const float* bias, void* fine,
                       int n, int dc, int hc, int wc, int a, int b, void* stream);
/* dwt[A][B][8] (+)= sum_o coarse[o,a]*fine[2o+k,b]; optional bias grads over either side. */
int vs_k2s2_wgrad(int dtype, const void* coarse, const void* fine, float* dwt,
                  float* dbias_coarse, float* dbias_fine, int accumulate,
                  int n, int dc, int hc, int wc, int a, int b, void* stream);

/* ---- InstanceNorm3d(affine=False, eps 1e-5) + ReLU (joint_model.py:9-15,38,104) ------- */
/* a = relu((y-mean)*rstd) [+ skip]; mean/rstd derived from the fp64 stats[N][C][2] over `s` voxels.  */
int vs_inorm_relu_apply(int dtype, const void* y, const double* stats, const void* skip, void* a,
                        int n, long long s, int c, void* stream);
/* sums[N][C][2] (fp64) = (sum g*mask, sum g*mask*xhat) with mask = [y > mean]; zeroed by the call unless
 * flags & VS_FLAG_PREZEROED. */
int vs_inorm_relu_bwd_reduce(int dtype, const void* g, const void* y, const double* stats, double* sums,
                             int n, long long s, int c, int flags, void* stream);
/* dy = rstd * (g*mask - sums0/s - xhat * sums1/s)                                         */
int vs_inorm_relu_bwd_apply(int dtype, const void* g, const void* y, const double* stats, const double* sums,
                            void* dy, int n, long long s, int c, void* stream);
/* dst += src (gradient fan-in at the additive skips, joint_model.py:381,383)               */
int vs_add_inplace(int dtype, void* dst, const void* src, long long count, void* stream);
/* ---- per-(n,c) affine + ReLU passes: BatchNorm3d(affine, running statistics) + ReLU, joint_model.py:12-13 -------
 * The statistics of BatchNorm are pooled over the batch, which the host does on the [N][C] outputs of the convolution
 * kernels (a few hundred numbers); gamma / beta / the stored shift are folded into per-(n,c) coefficient tables and
 * these kernels do the tensor passes.  y, skip, a, g, dy: NDHWC [N,S,C] of `dtype`, C a power of two in [8,256].
 *   apply:       a  = relu(k*y + b) + skip                    kb   [N][C][2] fp32 = (k, b); skip may be NULL
 *   bwd_reduce:  sums[n][c] = (sum gm, sum gm*xhat)           gm = g*[k*y+b > 0], xhat = k2*y + b2; k2b2 [N][C][2];
 *                                                              sums [N][C][2] fp64, zeroed by the call
 *   bwd_apply:   dy = c0*gm + c1 + c2*y                        coef [N][C][3] fp32                                  */
int vs_affine_relu_apply(int dtype, const void* y, const float* kb, const void* skip, void* a, int n, long long s, int c,
                         void* stream);
int vs_affine_relu_bwd_reduce(int dtype, const void* g, const void* y, const float* kb, const float* k2b2, double* sums,
                              int n, long long s, int c, void* stream);
int vs_affine_relu_bwd_apply(int dtype, const void* g, const void* y, const float* kb, const float* coef, void* dy,
                             int n, long long s, int c, void* stream);

/* ---- softmax over n_class=2 (joint_model.py:225,367) ---------------------------------- */
/* logits: NDHWC fp32 [N][S][2]; probs: planar fp32 [N][2][S]                               */
int vs_softmax2_fwd(const float* logits, float* probs, int n, long long s, void* stream);
/* dlogits (NDHWC `dtype`) from planar dprobs and saved planar probs                        */
int vs_softmax2_bwd(int dtype, const float* dprobs, const float* probs, void* dlogits,
                    int n, long long s, void* stream);

/* same gradient written as an 8-channel bf16 NDHWC tensor [N][S][8] (channels 2..7 zero) so the head's wgrad / dgrad
 * take the tensor-core kernels; db[2] += sum_v dlogits (the out_block bias gradient) when non-NULL.       */
int vs_softmax2_bwd_pad8(const float* dprobs, const float* probs, void* dlogits8, float* db,
                         int n, long long s, void* stream);
/* planar fp32 [N][C][S], C <= 8 -> NDHWC bf16 [N][S][8] (zero channels C..7): the in-block input as a
 * tensor-core wgrad operand (joint_model.py:210,355 in_block Conv3d(n_channels, 8))                          */
int vs_planar_to_ndhwc8(const float* x, void* out, int n, int c, long long s, void* stream);

/* ---- Linear layers + reparameterisation of the VAE (joint_model.py:216-218,241-250) --- */
/* x: NDHWC `dtype` [B][S3][C] read in the reference's NCDHW flatten order i = c*S3 + v.
 * mean = Wm x + bm; std = relu(Ws x + bs); lat = mean + z*std*scale (use_z) or mean.       */
int vs_fc_encode_fwd(int dtype, const void* x, const float* wm, const float* bm, const float* ws,
                     const float* bs, const float* z, float scale, int use_z,
                     float* mean, float* std, float* lat, int batch, int s3, int c, int dim, void* stream);
/* h = W2 lat + b2 written NDHWC `dtype` [B][S3][C]                                         */
int vs_fc_decode_fwd(int dtype, const float* lat, const float* w2, const float* b2, void* h,
                     int batch, int s3, int c, int dim, void* stream);
/* dlat[B][dim] = dh W2 ; dw2/db2 (+)= when non-NULL                                        */
int vs_fc_decode_bwd(int dtype, const void* dh, const float* lat, const float* w2, float* dlat,
                     float* dw2, float* db2, int accumulate, int batch, int s3, int c, int dim, void* stream);
/* gm = dlat + gmean_ext ; gs = (dlat*z*scale*use_z + gstd_ext) * [std>0]   (ext may be NULL)
 * dx = gm Wm + gs Ws (NDHWC `dtype`); dwm/dbm/dws/dbs (+)= when non-NULL.
 * gbuf: fp32 scratch [2][B][dim].                                                          */
int vs_fc_encode_bwd(int dtype, const void* x, const float* wm, const float* ws, const float* z,
                     float scale, int use_z, const float* std, const float* dlat,
                     const float* gmean_ext, const float* gstd_ext, float* gbuf, void* dx,
                     float* dwm, float* dbm, float* dws, float* dbs, int accumulate,
                     int batch, int s3, int c, int dim, void* stream);

/* Generic Linear (+ activation) of the encoder / discriminator head (joint_model.py:287-304: fc1 -> ReLU -> fc2 -> ReLU ->
 * fc_mean -> sigmoid).  x [batch][in_f], w [out_f][in_f], bias [out_f] or NULL, y [batch][out_f]; act 0 = identity,
 * 1 = ReLU, 2 = sigmoid.  Backward: gbuf [batch][out_f] scratch; dx / dw / db may be NULL; accumulate adds onto dw / db. */
int vs_linear_fwd(const float* x, const float* w, const float* bias, float* y, int batch, int in_f, int out_f, int act,
                  void* stream);
int vs_linear_bwd(const float* x, const float* w, const float* y, const float* dy, float* gbuf, float* dx, float* dw,
                  float* db, int accumulate, int batch, int in_f, int out_f, int act, void* stream);

/* ---- losses (utils/evaluation.py:6-18,42-80; main_source.py:150-182) ------------------ */
typedef enum { VS_TGT_TENSOR = 0, VS_TGT_BINARIZE = 1, VS_TGT_CONFIDENT = 2, VS_TGT_LABEL = 3,
               VS_TGT_ARGMAX = 4 } vs_target_mode;
/* sums[N][C][3] = (sum s*t', sum s', sum t') over voxels, with t' = f_mode(t):
 *   TENSOR t, BINARIZE [t>=.5], CONFIDENT (t>.8 ->1, t<.2 ->0), LABEL onehot of label[N][1][S]
 *   (fp32 class index), ARGMAX: both s and t replaced by onehot(argmax) (binary=True).
 * src, tgt planar fp32 [N][C][S].  zeroed by the call.                                     */
int vs_dice_sums(const float* src, const float* tgt, int mode, float* sums,
                 int n, int c, long long s, void* stream);
/* gsrc = gper[n,c]*(2 t' D - 2I)/D^2, gtgt = gper[n,c]*(2 s D - 2I)/D^2, D = S+T+eps;
 * gsrc/gtgt may be NULL; accumulate adds into them.                                        */
int vs_dice_bwd(const float* src, const float* tgt, int mode, const float* sums, const float* gper,
                float eps, float* gsrc, float* gtgt, int accumulate,
                int n, int c, long long s, void* stream);
/* out[0] = mean_b 0.5*(sum std^2 + sum mean^2 - 2 sum log(std+1e-5))                       */
int vs_kl_fwd(const float* mean, const float* std, float* out, int batch, int dim, void* stream);
int vs_kl_bwd(const float* mean, const float* std, const float* gout, float* gmean, float* gstd,
              int batch, int dim, void* stream);
/* elementwise thresholding: mode BINARIZE or CONFIDENT                                     */
int vs_binarize(const float* a, float* out, int mode, long long count, void* stream);
/* Input intensity normalisation on the device: out = (clip(x, lo, hi) - sub) / div, IEEE fp32 --
 * the reference's Clip + CenterIntensities transforms (utils/utils.py:508-533,572-618; main_target.py:223-224).
 * in_kind 0: x fp32; 2: x int16 (raw Hounsfield units).  out: fp32, same element count (the in-blocks' planar input). */
int vs_clip_center(int in_kind, const void* x, float* out, long long count, float lo, float hi, float sub, float div,
                   void* stream);
/* CropResize (utils/utils.py:220-293) on the device: `src` [sd][sh][sw] fp32 is cropped to crop9 = {start[3], length[3],
 * leading zero padding[3]} (HOST array of 9 ints; the caller derives it from the label's bounding box exactly as the
 * reference does) inside a zero cube of side `side`, optionally anti-alias filtered, and resampled to out [od][oh][ow]:
 * order 1 = skimage.transform.resize(img, size) (Gaussian pre-filter sigma = max(0, (side/out - 1)/2), linear, mode
 * 'mirror'), order 0 + anti_alias 0 = resize(label, size, order=0, anti_aliasing=False).  tmp0 / tmp1: side^3 fp32
 * workspaces (may be NULL when no axis is downsampled or anti_alias = 0).  The crop cube is never materialised.    */
int vs_crop_resize(const float* src, int sd, int sh, int sw, const int* crop9_host, int side, float* out, int od, int oh,
                   int ow, int order, int anti_alias, float* tmp0, float* tmp1, void* stream);
int vs_gauss_radius(double sigma);
/* label[N][1][S] (fp32 class index) -> planar one-hot [N][C][S] (main_target.py:520-522)   */
int vs_one_hot(const float* label, float* out, int n, int c, long long s, void* stream);

/* ---- optimiser / teacher update on flat fp32 arenas (main_target.py:347-352,512-516) --- */
/* SGD(momentum, dampening 0, no wd, no nesterov): first!=0 initialises buf=g.  gscale scales g. */
int vs_sgd_step(float* p, const float* g, float* buf, long long count, float lr, float momentum,
                int first, float gscale, void* stream);
int vs_adam_step(float* p, const float* g, float* m, float* v, long long count, float lr, float beta1,
                 float beta2, float eps, int step, float gscale, void* stream);
/* teacher = alpha*teacher + (1-alpha)*student                                              */
int vs_ema_update(float* teacher, const float* student, long long count, float alpha, void* stream);
/* device-side loss composition for the target-domain step incl. dynamic lambda (type 8,
 * main_target.py:550-560) without a host sync: terms = (recon_loss, dsc_loss_fake, klloss);
 * writes final loss and the two scalar weights (d final/d recon, d final/d fake, d final/d kl). */
int vs_compose_target_loss(const float* terms, float lambda_vae, int loss_type, int use_kl,
                           float* final_loss, float* weights, void* stream);
/* dst[r][0..row_len) += src[r * src_row_stride + 0..row_len) with atomics (gradient fan-in from concurrent streams) */
int vs_atomic_add_rows(float* dst, const float* src, long long rows, long long row_len, long long src_row_stride,
                       void* stream);
/* The scalar tail of a teacher-student step in ONE launch (main_target.py:543-546,550-560,588-590): from the three
 * vs_dice_sums results (student vs reconstruction, vs ground truth [monitor, may be NULL], vs pseudo label) and the
 * teacher KL value (may be NULL): out5 = (final, recon_loss, dice_loss, dice_loss_fake, kl) with
 * loss = 1 - mean_{n, c in [bot,top)} Dice, and gper2[2][N][C] = d final / d Dice[n][c] of the reconstruction and the
 * pseudo-label term (the gper arguments of the two vs_dice_bwd launches).  only_pseudo: final = dice_loss_fake.   */
int vs_joint_target_finish(const float* sums_recon, const float* sums_gt, const float* sums_fake, const float* kl,
                           int n, int c, int bot, int top, float eps, float lambda_vae, int loss_type, int use_kl,
                           int only_pseudo, float* out5, float* gper2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VAESEG_B200_H */
