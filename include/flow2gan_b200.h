/* flow2gan_b200 -- C ABI of the B200-native (sm_100a) Flow2GAN hot path.
 *
 * The reference (k2-fsa/Flow2GAN) has no FFI / plugin layer: its boundary is the Python
 * nn.Module surface.  This header is therefore the NEW native boundary
 * that the Python host layer (flow2gan_b200/*.py, mirroring flow2gan/models/*.py) binds through
 * ctypes.  Every entry point:
 *   - takes plain device pointers (fp32 unless noted), sizes and an explicit cudaStream_t
 *     (passed as void*); the caller owns all memory (torch tensors on the Python side);
 *   - launches asynchronously on that stream, spawns no threads, keeps no per-call state;
 *   - returns 0 on success, a positive cudaError_t or a negative F2G_E* code otherwise;
 *     f2g_last_error() returns the message (the Python layer raises RuntimeError with it).
 *
 * Activation layout: "token-major / channel-last" -- a (B, C, T) tensor of the reference is
 * held as rows = b*T + t, columns = channel, leading dimension `ld` (multiple of 4 floats).
 * Each function cites the reference code it replaces (paths relative to /root/reference).
 */
#ifndef FLOW2GAN_B200_H_
#define FLOW2GAN_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define F2G_ABI_VERSION 8
#define F2G_GEMM_MAX_PROBLEMS 8

enum { F2G_ACT_NONE = 0, F2G_ACT_PRELU = 1, F2G_ACT_LEAKY = 2, F2G_ACT_SILU = 3 };
enum { F2G_SPEC_PACKED = 0, F2G_SPEC_MAG = 1, F2G_SPEC_POWER = 2, F2G_SPEC_COMPLEX_BANDS = 3 };

int f2g_abi_version(void);
const char* f2g_last_error(void);
/* Fails (F2G_EARCH) unless the current device is compute capability 10.x. */
int f2g_check_device(void);

/* ---------------------------------------------------------------------------------------
 * Tensor-core contraction  C[M,N] = epi(alpha * A[M,K] . B[N,K]^T)   (tcgen05 kind::tf32 / kind::f16)
 * replaces: nn.Conv1d(kernel_size=1) / nn.Linear (flow2gan/models/modules.py:443-451,
 * 563,570-579,593) and, via overlapping-row operands, nn.Conv2d of the discriminators
 * (flow2gan/models/discriminators.py:65-76,171-184) plus their dgrad / wgrad.
 *   a_mn = 0: A is K-major, element (m,k) at a[m*lda + k];  a_mn = 1: MN-major, a[k*lda + m].
 *   b_mn = 0: B is K-major, element (n,k) at b[n*ldb + k];  b_mn = 1: MN-major, b[k*ldb + n].
 *   epilogue, in this order: x = alpha*acc + bias[n]; act (PReLU slope[n] / leaky / SiLU);
 *   x *= (gate[m,n] > 0 ? 1 : slope[n]) if gate; x += res_scale[n] * res[m,n] if res;
 *   x *= row_scale[m] if row_scale; x += C[m,n] if accumulate; round-to-nearest TF32 if asked.
 * Up to F2G_GEMM_MAX_PROBLEMS problems sharing (bn, a_mn, b_mn) run as ONE persistent launch.
 * ------------------------------------------------------------------------------------- */
typedef struct F2GGemm {
  const float* a;
  const float* b;
  float* c;
  int M, N, K;
  int lda, ldb, ldc;
  int a_mn, b_mn;
  int bn; /* N tile: 64, 128 or 256 */
  const float* bias;
  const float* slope;
  const float* res;
  const float* res_scale;
  const float* row_scale;
  const float* gate;
  int ld_res, ld_gate;
  int act;
  float leaky;
  float alpha; /* 0 means 1 */
  int round_tf32;
  int accumulate;
  float* c_pre; /* optional: alpha*acc + bias before the activation, (M, ld_pre) */
  int ld_pre;
  int split_k;  /* > 1: K is split over that many CTAs per tile, results atomically added into a
                   pre-zeroed C (plain epilogue only) -- used by weight-gradient GEMMs with huge K */
  /* Windowed ("implicit im2col") A operand: A is an OVERLAPPING-row view of a padded channel-last
   * activation buffer -- row r starts at a + r*lda and is a_seg_len floats long (lda < a_seg_len:
   * consecutive output pixels share taps), and the contraction index is cut into segments of
   * a_seg_len (one per kernel row); segment s reads buffer row r + s*a_seg_shift.  Rows outside
   * [0, a_rows) read as zero (TMA out-of-bounds fill).  a_seg_len == 0: plain operand.
   *   a_mn = 0: element (m, k) = a[(m + (k / a_seg_len) * a_seg_shift) * lda + k % a_seg_len]
   *   a_mn = 1: element (m, k) = a[(k + (m / a_seg_len) * a_seg_shift) * lda + m % a_seg_len]
   * a_seg_len must be a multiple of 32.  CTA-pair kernel only (max M > 128). */
  int a_seg_len, a_seg_shift, a_rows;
  /* Half-width operands (tcgen05 kind::f16, fp32 accumulate).  IEEE fp16 carries the same 11-bit
   * significand as TF32, so for operands inside the fp16 range the products are the ones the TF32
   * path forms -- at half the operand bytes through shared memory and twice the tensor issue rate
   * (measured: both operand types sit at the same distance from the fp32 oracle).  ab_f16 = 1: a and b point to fp16 (__half) matrices, K-major
   * only (a_mn = b_mn = 0), lda / ldb in ELEMENTS (multiples of 8), 16 B-aligned pointers.
   * c_f16 = 1: C is stored as fp16 (round-to-nearest, clamped to +-65504), ldc in elements
   * (multiple of 8); needs a bias+activation or bias-only epilogue (no res / gate / accumulate /
   * c_pre / split_k).  The epilogue itself always runs in fp32.  CTA-pair kernel only. */
  int ab_f16, c_f16;
  /* Producer -> consumer chaining INSIDE one launch (pwconv1 -> PReLU -> pwconv2 of a ConvNeXt
   * block without a kernel boundary): a problem with done_counter != NULL adds 1 (release, gpu
   * scope) to done_counter[m / 256] every time one CTA has stored its 128 rows of one output
   * tile; a problem with wait_counter != NULL does not read the A rows of the 256-row tile
   * m / 256 before wait_counter[m / 256] has reached 2 * (number of N tiles of the producer), the
   * producer being the problem of the SAME group whose done_counter equals this wait_counter
   * (same M, split_k = 1).  All producer tiles are scheduled before any consumer tile on every
   * CTA pair, so the launch cannot deadlock as long as its <= 148 CTAs are co-resident -- do not
   * run two chained groups concurrently on different streams.  The caller zeroes the counters
   * (ceil(M / 256) ints) before every launch (f2g_block_pre_group can do it: zero_ptr).
   * CTA-pair kernel only. */
  int* done_counter;
  const int* wait_counter;
  /* fp16 range guard (c_f16 = 1 only): when any value written to C lies outside +-65504 BEFORE the
   * saturating conversion (or is not finite), bit 0 of *sat_flag is set (atomicOr).  NULL: no check.
   * The host layer clears the flag per call and falls back to TF32 operands when it comes back set
   * (engine.py: fp16 shares TF32's 11-bit significand but not its 8-bit exponent). */
  int* sat_flag;
} F2GGemm;

/* Watchdog of chained launches (done_counter / wait_counter): a consumer tile that does not see its
 * producer counter within ~1 s of polling writes {1, problem, row tile, counter value} to a mapped host
 * record and traps (the launch fails instead of hanging; f2g_last_error() of the next failing call
 * carries the record).  Chained launches rely on all their CTAs becoming resident: two of them must
 * not run concurrently on one device (serialise them across streams -- the Python host layer does,
 * engine.py::_chain_guard).  Returns the flag; out = the record. */
int f2g_chain_watchdog(int out[4]);

/* Host logic of a CTA-pair launch without a device: plans the group exactly as f2g_gemm_group would (N tiles,
 * chaining, the per-pair longest-processing-time tile schedule for `pairs` CTA pairs) and writes it to `out`:
 * {scheduled (0/1), problems, tiles, pairs used}, 12 ints per problem in launch order {M, N, K, N tile, row
 * tiles, column tiles, first tile, K splits, waits, expected count, publishes, TMA-store epilogue}, pairs + 1
 * list offsets, then one packed entry per tile: problem | K split << 3 | row tile << 12 | column tile << 22.
 * Returns the number of ints written (> 0) or a negative F2G_E* code.  Test / inspection entry: no reference
 * counterpart (the reference leaves scheduling to cuBLAS). */
int f2g_gemm_plan(const F2GGemm* problems, int n_problems, int pairs, int* out, int out_ints);

int f2g_gemm_tf32(const F2GGemm* problems, int n_problems, void* stream);

/* ---------------------------------------------------------------------------------------
 * STFT family.  replaces torch.stft(center=True, reflect, periodic hann, onesided) as used by
 * STFT.forward + fft_to_real (modules.py:31-38,68-84), torchaudio MelSpectrogram /
 * Spectrogram (modules.py:131-143,180-214; gan.py:47-54; discriminators.py:163,186-196).
 * audio: (B, T) rows of stride ld_audio.  frames = 1 + T/hop.  One CTA per frame.
 *   mode PACKED : out[(b*frames+f)*ld_out + c] = Re(bin c), c<=n/2 ; Im at c + n/2+1
 *   mode MAG    : |S|   (n/2+1 columns)          mode POWER : |S|^2
 *   mode COMPLEX_BANDS : interleaved (re, im) per bin = channel-last (rows, freq, 2) for the MRD
 * pre (optional, (B,2)): sample -> (x - pre[b][0]) * pre[b][1] before windowing (MRD).
 * fb (optional, (n/2+1, n_filt) row-major): if given, the MAG/POWER spectrum is contracted
 * with fb in fp32 inside the kernel and out gets n_filt columns; log_clip > 0 applies
 * log(max(., log_clip)) (safe_log, utils.py:221-232).
 * ------------------------------------------------------------------------------------- */
/* fb_ranges (optional, with fb): (n_filt + n_fft/2 + 1) int pairs -- per filter m the bin range [lo, hi)
 * outside which fb[:, m] is exactly zero, then per bin k the filter range [lo, hi) outside which
 * fb[k, :] is exactly zero.  Triangular mel / linear filterbanks are banded (2-3 filters per bin): the
 * kernels then skip the zero products (results unchanged bit for bit).  NULL: dense contraction. */
int f2g_stft(const float* audio, int B, int T, int ld_audio, int n_fft, int hop, int mode,
             const float* pre, const float* fb, int n_filt, float log_clip, float* out,
             int ld_out, int round_tf32, const int* fb_ranges, void* stream);

/* Grouped forms: up to 4 resolutions of the same (B, T) batch in one launch (the branch STFTs /
 * inverse transforms of AudioConvNeXt.forward, modules.py:699-719).
 *   stft  : in = audio (B, ld_in), out = packed rows (rows = B * (1 + T/hop), ld_out), PACKED mode
 *   irfft : in = packed rows (rows, ld_in), out = windowed frames (rows, n_fft); hop/frames unused */
typedef struct F2GSpecProblem {
  const float* in;
  float* out;
  int n_fft, hop, frames, rows, ld_in, ld_out;
} F2GSpecProblem;
int f2g_stft_group(const F2GSpecProblem* problems, int n_problems, int B, int T, int round_tf32,
                   void* stream);
int f2g_irfft_group(const F2GSpecProblem* problems, int n_problems, void* stream);

/* per-row DC removal + peak normalisation constants (discriminators.py:187-190):
 * pre[b] = (mean_t x, 0.8 / (max_t |x - mean| + 1e-9)). */
int f2g_dc_peak(const float* audio, int B, int T, int ld_audio, float* pre, void* stream);

/* iSTFT stage 1: packed rows -> windowed time frames  fr[(b*frames+f)*n + i] =
 * w[i] * irfft(spec)[i]   (real_to_fft + torch.istft inner part, modules.py:41-49,105-116). */
int f2g_irfft_frames(const float* packed, int rows, int ld, int n_fft, float* frames_out,
                     void* stream);

/* iSTFT stage 2 for up to 3 branches + branch fusion + Euler update
 * (modules.py:717-719, generator.py:136-168,263-265):
 *   pred[b,s] = sum_j weight[b,j] * OLA_j(s) / env_j(s)   (0 beyond hop_j*(frames_j-1))
 *   out = euler ? x + ((pred - x) / (1 - t)) * dt : pred ; optional clamp to [-1,1].
 * weight: (B, nb) or NULL (=> 1/nb each, "mean").  x may alias out. */
int f2g_ola_combine(const float* const* frames, const int* n_ffts, const int* hops,
                    const int* n_frames, int nb, const float* weight, const float* x, float* out,
                    int B, int T, int euler, float t, float dt, int clamp, void* stream);

/* ---------------------------------------------------------------------------------------
 * ConvNeXt-block pieces (modules.py:286-416,456-495), channel-last rows.
 * ------------------------------------------------------------------------------------- */
/* BiasNorm over the channel dim of each row: y = x * mean((x-bias)^2)^-1/2 * exp(log_scale);
 * log_scale is a device scalar.  y may alias x.  inv_out (optional, rows) receives the row scale. */
int f2g_biasnorm(const float* x, int rows, int C, int ld, const float* bias,
                 const float* log_scale, float* y, int ld_y, float* inv_out, void* stream);

/* Fused block prologue: mask -> depthwise conv k=7 (zero pad 3, bias) -> BiasNorm ->
 * + cond_proj row -> * (1 + time scale) -> TF32 round.  (ConvNeXtBlock.forward :473-485)
 *   x: (B*T, ld_x) rows; dw_wT: (7, C) transposed depthwise weights; row_mask: (B*T) or NULL;
 *   cond: rows of ld_cond (or NULL); cond row for (b,t) = t < cond_T*factor ? b*cond_T +
 *   t/factor : zero_row;  tscale: (B, ld_ts) holds time_embed_proj(te) (the "1 +" is added
 *   here) or NULL.  Optionally also writes the pre-norm conv output (needed by backward). */
int f2g_block_pre(const float* x, int B, int T, int C, int ld_x, const float* dw_wT,
                  const float* dw_b, const float* bn_bias, const float* bn_log_scale,
                  const float* row_mask, const float* cond, int ld_cond, int cond_T, int factor,
                  int zero_row, const float* tscale, int ld_ts, float* out, int ld_out,
                  float* conv_out, float* inv_rms_out, void* stream);

/* The same for up to 4 independent problems in ONE launch (the three branches' blocks of equal
 * depth): fields as the arguments of f2g_block_pre. */
typedef struct F2GBlockPre {
  const float* x;
  const float* dw_wT;
  const float* dw_b;
  const float* bn_bias;
  const float* bn_log_scale;
  const float* row_mask;
  const float* cond;
  const float* tscale;
  float* out;
  float* conv_out;
  float* inv_rms_out;
  int B, T, C, ld_x, ld_cond, cond_T, factor, zero_row, ld_ts, ld_out;
  int* zero_ptr; /* problem 0 only: zero_n ints are cleared by this launch (the chaining counters of */
  int zero_n;    /* the GEMM group that follows in the stream), or NULL */
  int out_f16; /* F2G_PRE_OUT_F16 (1): `out` points to fp16 rows (ld_out in elements, multiple of 4),
                  RN-rounded and clamped to +-65504 instead of TF32-rounded fp32 -- operand of an
                  ab_f16 GEMM */
  int* sat_flag; /* fp16 range guard (out_f16 only): bit 1 of *sat_flag is set when a value lies outside
                    +-65504 before the clamp (or is not finite); NULL: no check */
} F2GBlockPre;
enum { F2G_PRE_OUT_F16 = 1 };
int f2g_block_pre_group(const F2GBlockPre* problems, int n_problems, void* stream);

/* Small dense layers for few-row inputs, fp32 SIMT, up to 4 independent problems per launch:
 * out[b,o] = act(bias[o] + in[b,:].W[o,:]), b < B (time_mlp / time_embed_proj of the three
 * branches, modules.py:569-573,451,485).  K multiple of 128 and <= 1536. */
typedef struct F2GLinear {
  const float* in;
  const float* W;
  const float* bias;
  float* out;
  int K, O, ld_in, ldw, ld_out;
} F2GLinear;
int f2g_linear_small(const F2GLinear* problems, int n, int B, int act, void* stream);

/* SinusoidalPosEmb (modules.py:223-232): emb[b] = [sin(scale t f_i) ; cos(scale t f_i)], f_i
 * from the caller's table freqs[dim/2] = exp(-i ln(10000)/(dim/2-1)). */
int f2g_time_sinusoid(const float* t, int B, int dim, const float* freqs, float scale, float* out,
                      void* stream);

/* Generic strided 2-D gather used for weight packing / layout changes:
 * dst[r*ld + c] = src[r*src_rs + c*src_cs] for c < cols, 0 for cols <= c < ld_fill. */
int f2g_pack2d(const float* src, long long src_rs, long long src_cs, int rows, int cols,
               float* dst, int ld, int ld_fill, int round_tf32, void* stream);

/* (B, C, T) channel-first -> rows (b*T+t) x [k*C + c] im2col for a k-tap 'same' conv, zero
 * padded in time; columns K*C..ld zeroed (CondEncoder.in_proj, modules.py:511,536). */
int f2g_im2col_cf(const float* x, int B, int C, int T, int ktaps, float* out, int ld,
                  int round_tf32, void* stream);

/* frame mask from lengths: m[b*frames+f] = f < 1 + lens[b]/hop (modules.py:79-82,706-707). */
int f2g_frame_mask(const int* lens, int B, int frames, int hop, float* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Backward pieces of the generator path (adjoints of the kernels above; the contractions of
 * the backward pass are f2g_gemm_tf32 calls with MN-major operands).
 * ------------------------------------------------------------------------------------- */
/* Block prologue backward, stage A (BiasNormFunction.backward, modules.py:321-339, fused with the
 * time-scale / cond adjoints): du = da1*(1+ts), dy = inv*du - coef*(y-beta); per-row coef, gs. */
int f2g_block_bwd_a(const float* da1, int ld_da, const float* y, const float* inv,
                    const float* bn_bias, const float* log_scale, const float* tscale, int ld_ts,
                    int B, int T, int C, float* dy, float* du, float* coef, float* gs, void* stream);

/* Stage C: every per-channel parameter-gradient reduction of one block in one pass (atomicAdd
 * into the g_* accumulators; NULL skips a term): depthwise weights (7,C) / bias, BiasNorm bias /
 * log_scale, time-scale (B,C), residual ChannelScale, pwconv2 bias. */
typedef struct F2GBlockBwdC {
  const float* dy; const float* x; int ld_x; const float* row_mask; const float* y;
  const float* coef; const float* gs; const float* bn_bias; const float* da1; int ld_da;
  const float* inv; const float* cond; int ld_cond; int cond_T; int factor; int zero_row;
  const float* dxo; int ld_dxo;
  float* g_dww; float* g_dwb; float* g_beta; float* g_ls; float* g_ts; int ld_gts; float* g_rs;
  float* g_b2;
  int B, T, C;
} F2GBlockBwdC;
int f2g_block_bwd_c(const F2GBlockBwdC* args, void* stream);

/* Stage B: dx = mask * dwconv^T(dy) + rs * dxo  (either term may be absent). */
int f2g_block_bwd_b(const float* dy, const float* dw_wT, const float* row_mask, const float* dxo,
                    int ld_dxo, const float* rs, int B, int T, int C, float* dx, int ld_dx,
                    void* stream);

/* Activation backward over rows with fused column reductions: dz = dh * act'(z);
 * g_bias[c] += sum dz; g_slope[c] += sum dh*min(z,0) (PReLU).  act = F2G_ACT_*. */
int f2g_act_bwd(const float* dh, int ld_dh, const float* z, int ld_z, const float* slope, float leaky,
                int act, int rows, int cols, float* dz, int ld_dz, float* g_bias, float* g_slope,
                int round_tf32, void* stream);

/* The same for the windowed convs (F2GGemm::a_seg_len): GEMM rows are (n, line, r) with Hl lines of
 * R columns per image, real outputs at line < Ho, r < Wo.  dy is the (Nb, Ho, Wo, cols) gradient
 * view with element strides s_n / s_h / s_w (channel stride 1); z the saved GEMM output rows.
 * Writes ALL Nb*Hl*R rows of dz (zeros on padding rows, on columns [cols, cols_total) and on the
 * guard_rows rows in front of dz) and adds the column sums of the real rows to g_bias.
 * act = F2G_ACT_NONE or F2G_ACT_LEAKY; columns, strides, leading dimensions multiples of 4. */
int f2g_act_bwd_win(const float* dy, long long s_n, long long s_h, long long s_w, int Nb, int Hl, int R,
                    int Ho, int Wo, const float* z, int ld_z, float leaky, int act, int cols, int cols_total,
                    float* dz, int ld_dz, int guard_rows, float* g_bias, int round_tf32, void* stream);

/* Adjoint of upsample_cond (modules.py:668-680): frame-rate gradient rows -> mel-rate rows
 * (+ the shared zero row). */
int f2g_cond_reduce(const float* du, int B, int T, int C, int cond_T, int factor, int zero_row,
                    float* out, int ld_out, void* stream);

/* iSTFT adjoint: g (B,T) -> padded-domain gs (B, n + hop*(frames-1)) = g*scale/env, then
 * f2g_istft_bwd_spec: framed rFFT * c_k/n (* row mask) -> packed (rows, ld) gradient. */
int f2g_istft_bwd_prep(const float* g, int B, int T, int n_fft, int hop, int frames, float scale,
                       float* gs, void* stream);
int f2g_istft_bwd_spec(const float* gs, int B, int Lp, int n_fft, int hop, const float* row_mask,
                       float* dpacked, int ld, int round_tf32, void* stream);

/* STFT adjoint: packed spectrum gradient -> windowed frame gradients -> fold onto the signal
 * (overlap-add + reflect-padding adjoint). */
int f2g_stft_bwd_frames(const float* dpacked, int rows, int ld, int n_fft, float* frames_out,
                        int interleaved, void* stream);
int f2g_stft_bwd_fold(const float* frames_grad, int B, int T, int n_fft, int hop, int frames,
                      float* dx, int accumulate, void* stream);

/* Fused backward of a [log-]filterbank spectrogram (MelSpectrogram power=1 / LinearFilter-
 * Spectrogram power=2): dF (rows, ld_dF) -> windowed frame gradients (rows, n_fft). */
int f2g_spec_loss_bwd(const float* audio, int B, int T, int ld_audio, int n_fft, int hop, int mode,
                      const float* fb, int n_filt, float log_clip, const float* dF, int ld_dF,
                      float* frames_out, const int* fb_ranges, void* stream);

/* Direct convolution of the discriminators' FIRST layers (torch.nn.Conv2d(1, 32, (5,1), (3,1)) of
 * DiscriminatorP, discriminators.py:65, run here along the contiguous axis; Conv2d(2, 32, (3,9)) of
 * DiscriminatorR, :171) with the LeakyReLU fused, fp32 SIMT: Cin in {1, 2}, Cout = 32, kh*kw*Cin <= 64.
 * x: channel-last (Nb, H, W, Cin) view with element strides (pitch_n, pitch_h, pitch_w), channel stride 1;
 * w, bias: the parameter tensors (Co, Cin, kh, kw) / (Co); y: (Nb*Ho*Wo, 32) contiguous; leaky < 0: no
 * activation.  bwd: dy / y contiguous (rows, 32); gw_packed (kh*kw*Cin, 32) [k = (s*kw + t)*Cin + ci][co] and
 * gb (32) are ACCUMULATED into (zero them first; NULL gw_packed: no weight gradient); dx (Nb, H, W, Cin)
 * contiguous or NULL.  The weight gradient is reduced without atomics (bit-reproducible): `scratch` receives one
 * partial record of kh*kw*Cin*32 + 32 floats per block (as many blocks as fit, <= 2368). */
int f2g_conv_small_fwd(const float* x, int Nb, int H, int W, int Cin, long long pitch_n, long long pitch_h,
                       long long pitch_w, const float* w, const float* bias, int Co, int kh, int kw, int sh, int sw,
                       int ph, int pw, float leaky, float* y, void* stream);
int f2g_conv_small_bwd(const float* x, int Nb, int H, int W, int Cin, long long pitch_n, long long pitch_h,
                       long long pitch_w, const float* w, int Co, int kh, int kw, int sh, int sw, int ph, int pw,
                       float leaky, const float* dy, const float* y, float* gw_packed, float* gb, float* dx,
                       float* scratch, long long scratch_floats, void* stream);

/* out[c] += sum_r x[r*ld + c]  (bias gradients). */
int f2g_colsum(const float* x, int ld, int rows, int cols, float* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Conv2d of the discriminators (discriminators.py:65-76,171-184) = gather + f2g_gemm_tf32 on
 * channel-last tensors.  The input may be a W-band [w0, w0+W) of a wider (Nb, H, Wfull, C)
 * buffer: pass the band base pointer and the buffer pitches (elements).
 * ------------------------------------------------------------------------------------- */
typedef struct F2GConv2d {
  int Nb, H, W, C;
  long long pitch_h, pitch_n;
  int kh, kw, sh, sw, ph, pw;
  int ldk; /* leading dimension of the column matrix: >= kh*kw*C, multiple of 4 */
} F2GConv2d;
/* col[(n,ho,wo), (ih,iw,c)] = x[n, ho*sh-ph+ih, wo*sw-pw+iw, c] (0 outside), padding columns 0 */
int f2g_im2col2d(const float* x, const F2GConv2d* geom, float* col, int round_tf32, void* stream);
/* adjoint of im2col2d: dx (+)= sum of the taps that read each input element */
int f2g_col2im2d(const float* dcol, const F2GConv2d* geom, float* dx, int accumulate, void* stream);
/* Zero-padded, TF32-rounded copy for the windowed convs: out is (Nb, Hl, Wp, C) contiguous plus
 * `slack` trailing floats (zeroed); out[n, h, w, :] = x[n, h-ph, w-pw, :] inside, 0 outside.
 * x element (n, h, w, c) at x[n*pitch_n + h*pitch_h + w*pitch_w + c].  C % 4 == 0. */
int f2g_pad2d(const float* x, int Nb, int H, int W, int C, long long pitch_n, long long pitch_h,
              long long pitch_w, int Hl, int Wp, int ph, int pw, long long slack, int round_tf32,
              float* out, void* stream);
/* dir 0: (Co, Ci, taps) parameter -> (Co_pad, ld) GEMM operand [co][tap*Ci + ci], TF32-rounded;
 * dir 1: packed gradient -> parameter layout. */
int f2g_conv_w_pack(const float* src, int Co, int Ci, int taps, int Co_pad, int ld, float* dst, int dir,
                    void* stream);
/* Weights of the phase-decomposed transposed convolution = input gradient of a stride-(1, sw)
 * windowed conv: for every phase p < sw (ntp = ceil((kw-p)/sw) taps) the matrix
 * out_p[ci][(s, jj, co)] = W[co][ci][s][p + sw*(ntp-1-jj)] (zero for co >= Co), TF32-rounded; phase
 * blocks back to back, block p at float offset Ci*kh*Cop*sum_{q<p} ntq.  w: (Co, Ci, kh, kw). */
int f2g_conv_w_pack_dgrad(const float* w, int Co, int Ci, int kh, int kw, int sw, int Cop, float* out,
                          void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused multi-tensor ScaledAdam step (flow2gan/optim.py:125-255,451-619) for ONE param group.
 * tab (device): one record per parameter tensor; chunks (device): int2 {tensor, chunk} covering
 * every tensor in 4096-element pieces; acc: 3 floats per tensor (scratch); tensor_state: 8
 * floats per tensor {param_rms, scale_exp_avg_sq, scale_grads[4], scale_step, -}; group_state:
 * {grad_norm (out), clip (out), model_norm_threshold (in, < 0 = unset)}; model_norms: ring of
 * clipping_update_period floats.  phase 0 = reductions + norm, phase 1 = clip + update; the
 * host refreshes the threshold between the phases on the (rare) steps the reference does.
 * ------------------------------------------------------------------------------------- */
typedef struct F2GAdamTensor {
  float* p;
  const float* g; /* NULL = no gradient this step (treated as zeros, optim.py:107-109) */
  float* v;       /* exp_avg_sq */
  float* d;       /* delta (momentum) */
  long long numel;
  int is_scalar;  /* numel == 1: no rms scaling, lr * scalar_lr_scale, clamp to +-scalar_max */
  int reserved;
} F2GAdamTensor;

typedef struct F2GAdamHyper {
  float lr, scalar_lr_scale, beta1, beta2, eps, param_min_rms, param_max_rms, scalar_max;
  int size_update_period, clipping_update_period, use_clipping;
} F2GAdamHyper;

int f2g_scaled_adam_step(const F2GAdamTensor* tab_dev, int n_tensors, const int* chunks_dev,
                         int n_chunks, float* acc_dev, float* tensor_state_dev,
                         float* group_state_dev, float* model_norms_dev, int step, int phase,
                         const F2GAdamHyper* hyper, void* stream);

/* ---------------------------------------------------------------------------------------
 * Data path either side of the generator: wav payload -> training /
 * inference waveform, waveform -> wav payload, and the fp64 running model average.
 * ------------------------------------------------------------------------------------- */
enum { F2G_PCM_F32 = 1, F2G_PCM_S16 = 16, F2G_PCM_S24 = 24, F2G_PCM_S32 = 32 };

/* mono[i] = mean over channels of frame first_frame + i of an interleaved little-endian PCM
 * payload (device memory), i < n_frames; integer formats are scaled by 2^-(bits-1) like
 * libsndfile / torchaudio.load.  stats (optional, 2 floats, zeroed by the caller): [0] += sum
 * mono^2 (the loader's silence test), [1] = max |mono| (sox `norm`).
 * replaces: lhotse Recording.load_audio(offset, duration) + is_silence + y.mean(dim=0)
 * (flow2gan/dataset.py:122-160), torchaudio.load + torch.mean (flow2gan/bin/infer_dir.py:217-220,
 * test_from_wav.py:62-66). */
int f2g_pcm_decode(const void* pcm, int sample_format, int channels, long long first_frame,
                   long long n_frames, float* mono, float* stats, void* stream);

/* out = resample(g * x): g = 10^(norm_db/20) / stats[1] when `stats` is given (sox effect
 * ["norm", dB], dataset.py:164-168), else 1; polyphase sinc interpolation with the
 * (new_r, 2*width + orig_r) tap table of torchaudio.functional.resample (orig_r/new_r = the
 * rates divided by their gcd; dataset.py:170-173); n_out <= (n_in / orig_r + 1) * new_r, the
 * reference keeps ceil(new_r * n_in / orig_r).  orig_r = new_r = 1, width = 0, taps = {1} is the
 * gain-only pass. */
int f2g_gain_resample(const float* x, long long n_in, const float* stats, float norm_db, int orig_r,
                      int new_r, int width, const float* taps, float* out, long long n_out,
                      void* stream);

/* out[i] = (int16) lrintf(x[i] * 32767) -- soundfile.write's default PCM_16 conversion
 * (flow2gan/bin/infer.py:208-212, bin/infer_dir.py:237); clamp != 0 bounds x to [-1, 1] first. */
int f2g_pcm16_encode(const float* x, long long n, int clamp, short* out, void* stream);

/* avg = (avg * w_avg + cur * w_cur) * scale, every tensor of a model in one launch (chunks: int
 * pairs {tensor, 4096-element chunk}), with the rounding sequence of the reference's in-place
 * torch ops for each dtype pair; replaces average_state_dict as used by update_averaged_model /
 * update_ema_model / average_checkpoints_with_averaged_model
 * (flow2gan/checkpoint.py:378-409,411-441,443-531). */
typedef struct F2GAvgTensor {
  void* avg;       /* accumulator: fp64 (model_avg as created) or fp32 (after save_checkpoint's in-place downcast) */
  const void* cur; /* fp32 (model parameters / buffers) or fp64 (another averaged model) */
  long long numel;
  int cur_is_f64;
  int avg_is_f32;
} F2GAvgTensor;
int f2g_average_update(const F2GAvgTensor* tab_dev, const int* chunks_dev, int n_chunks, double w_avg,
                       double w_cur, double scale, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused multi-tensor loss reductions of GAN.forward (flow2gan/models/gan.py:57-99):
 *   F2G_LOSS_L1    term = scale * sum |a - b|              (feature matching :72-87, a = real.detach();
 *                                                          log-mel reconstruction :89-99); gradient w.r.t. b
 *   F2G_LOSS_HINGE term = scale * sum clamp(1 + sign*a, 0) (discriminator_loss :57-63, generator_loss
 *                                                          :65-70; sign = -1 for real scores in the D loss and
 *                                                          for fake scores in the G loss, +1 for fake scores in
 *                                                          the D loss); gradient w.r.t. a
 * with scale = 1 / numel (torch.mean).  `terms` is a HOST array (descriptors travel as kernel
 * parameters, 24 per launch); tensors are addressed through 4-D dims / element strides (permuted or
 * channel-sliced views).  backward = 0: out[0] = sum of all terms.  backward = 1: every term's
 * `grad` (contiguous, logical order) = gout[0] * d(term)/d(input).
 * ------------------------------------------------------------------------------------- */
enum { F2G_LOSS_L1 = 0, F2G_LOSS_HINGE = 1 };
typedef struct F2GLossTerm {
  const float* a;
  const float* b;
  float* grad;
  long long numel;
  long long stride_a[4];
  long long stride_b[4];
  int dims[4];
  float scale;
  float sign;
  int mode;
  int reserved;
} F2GLossTerm;
int f2g_loss_terms(const F2GLossTerm* terms, int n_terms, int backward, float* out, const float* gout,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLOW2GAN_B200_H_ */
