/*
 * apgs_b200 -- C ABI of the B200 (sm_100a) kernels behind the APG reweight -> resample hot path.
 *
 * The reference (hao-w/apgs) is pure Python/PyTorch and has NO FFI boundary:
 * the path sits behind plain Python objects.  Each entry point below names the reference
 * interface it replaces (file:line, relative to the reference tree); the Python binding a
 * maintainer would add (ctypes) is shown in INTEGRATION.md and implemented in apgs_b200/_lib.py.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless stated; tensors are dense row-major fp32 in the
 *     reference layout (S, B, ...) -- S particles, B instances;
 *   - no allocation, no ownership transfer, no global mutable state; every call is asynchronous
 *     on `stream` (a cudaStream_t passed as void*) and re-entrant from any host thread;
 *   - return value: 0 ok, < 0 bad argument (APGS_E_*), > 0 a cudaError_t;
 *     apgs_last_error() returns a thread-local message for the last non-zero return;
 *   - nothing throws across the ABI.
 */
#ifndef APGS_B200_H
#define APGS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APGS_ABI_VERSION 2

#if defined(__GNUC__)
#define APGS_API __attribute__((visibility("default")))
#else
#define APGS_API
#endif

#define APGS_E_BADARG   (-1)
#define APGS_E_SHAPE    (-2)
#define APGS_E_ALIGN    (-3)
#define APGS_E_WORKSPACE (-4)
#define APGS_E_UNSUPPORTED (-5)

/* resampling strategy (apgs/resampler.py:8-11) */
#define APGS_SYSTEMATIC  0
#define APGS_MULTINOMIAL 1

/* how the CDF over the particle axis is accumulated (DESIGN.md "index exactness")
 *   AUTO    : EXACT when S <= 8192, BLOCKED otherwise
 *   EXACT   : one sequential chain per instance -- systematic: fp64 running sum rounded to fp32
 *             per element (== ATen CPU cumsum, apgs/resampler.py:31); multinomial: sequential
 *             fp32 running sum (== ATen CPU multinomial_with_replacement, apgs/resampler.py:41)
 *   BLOCKED : strip-parallel fp64 scan (bit-equal to EXACT-systematic except where fp64
 *             re-association moves a prefix across an fp32 rounding boundary, ~2^-29 / element) */
#define APGS_CDF_AUTO    0
#define APGS_CDF_EXACT   1
#define APGS_CDF_BLOCKED 2

APGS_API int         apgs_version(void);
APGS_API const char* apgs_last_error(void);

/* ----------------------------------------------------------------------------------------------
 * W1  weight normalisation + ESS + log-sum-exp over the particle axis.
 * replaces: F.softmax(log_w, 0) / 1/(w**2).sum(0)   apgs/gmm/objectives.py:93,98,141,146,181,186;
 *           apgs/dmm/objectives.py:94,100,137,143; apgs/bmnist/objectives.py:150,156,207,242,248.
 *   log_w (S,B) in;  w_out (S,B), ess (B), lse (B) out -- each output may be NULL.
 * workspace: apgs_ancestors_workspace_bytes(S, B) bytes (may be 0 -> NULL allowed).           */
APGS_API int apgs_weight_stats_f32(const float* log_w, int64_t S, int64_t B,
                          float* w_out, float* ess, float* lse,
                          void* workspace, size_t workspace_bytes, void* stream);

/* W2  per-point weights: w_local = softmax over S of log_w_local (S,B,N); log_w = sum_n.
 * replaces: apgs/gmm/objectives.py:216-218, apgs/dmm/objectives.py:173-175.
 *   w_local_out (S,B,N) and log_w_out (S,B) may each be NULL.                                  */
APGS_API int apgs_local_weight_stats_f32(const float* log_w_local, int64_t S, int64_t B, int64_t N,
                                float* w_local_out, float* log_w_out, void* stream);

/* ----------------------------------------------------------------------------------------------
 * R2/R3  ancestor sampling.   replaces Resampler.sample_ancestral_index, apgs/resampler.py:22-44.
 *   exactly one of log_w / w_norm is non-NULL ((S,B) fp32).  With log_w the kernel normalises
 *   itself (and can emit w_out/ess/lse like W1, so the reference's duplicated softmax is done
 *   once); with w_norm (already-normalised weights) ancestors are bit-exact with the reference
 *   run on CPU given the same uniforms.
 *   strategy APGS_SYSTEMATIC : uniforms = float32[B], one per instance (apgs/resampler.py:28);
 *                              u_stride_* ignored.
 *   strategy APGS_MULTINOMIAL: uniforms = float64, element (s,b) at uniforms[s*u_stride_s +
 *                              b*u_stride_b] (the reference's generator order is (B,S):
 *                              u_stride_s = 1, u_stride_b = S).
 *   anc64 (S,B) int64 and/or anc32 (S,B) int32 receive the ancestors (either may be NULL).
 *   workspace: 0 bytes while one instance's CDF fits a shared-memory tile (S <= ~17k); beyond that the transposed
 *   CDF (S*B floats) plus the per-chunk partials of the split-S pipeline, 256-byte aligned.          */
APGS_API size_t apgs_ancestors_workspace_bytes(int64_t S, int64_t B);
APGS_API int apgs_sample_ancestors_f32(const float* log_w, const float* w_norm, int64_t S, int64_t B,
                              int strategy, const void* uniforms,
                              int64_t u_stride_s, int64_t u_stride_b, int cdf_mode,
                              int64_t* anc64, int32_t* anc32,
                              float* w_out, float* ess, float* lse,
                              void* workspace, size_t workspace_bytes, void* stream);

/* ----------------------------------------------------------------------------------------------
 * R4  ancestor gather.  replaces Resampler.resample_4dims / resample_5dims,
 *     apgs/resampler.py:46-54 (no materialised int64 gather index).
 *   for every latent l:  dst_l[s, b, :] = src_l[anc[s, b], b, :]   with inner_bytes_l bytes per
 *   (s,b) row; src/dst dense, non-overlapping.  Exactly one of anc32 / anc64 is non-NULL.
 *   Up to APGS_MAX_LATENTS latents are gathered by ONE launch.                                 */
#define APGS_MAX_LATENTS 8
typedef struct apgs_latent {
    const void* src;
    void*       dst;
    int64_t     inner_bytes;
} apgs_latent_t;   /* HOST array */

APGS_API int apgs_gather_ancestors(const int32_t* anc32, const int64_t* anc64, int64_t S, int64_t B,
                          const apgs_latent_t* latents, int n_latents, void* stream);

/* R4 through an earlier resampling step that was never materialised (deferred gather): a block update that only READS the
 *   latents of the previous step can take them through that step's ancestors (apgs_gmm_z_step_fused_tab_f32:
 *   src_ancestors) instead of having them gathered first; the gather after it then composes the two index maps:
 *     first[l] == NULL:  dst_l[s, b, :] = src_l[anc32[s, b], b, :]                         (latent made by the update)
 *     first[l] != NULL:  dst_l[s, b, :] = src_l[first[l][anc32[s, b], b], b, :]            (latent still in the old order)
 *   first: HOST array of n_latents device pointers to (S,B) int32 ancestor maps.  One launch, same kernel as R4.
 *   Saves one apgs_gather_ancestors per sweep in apgs/gmm/objectives.py:37-45 (eta step -> resample -> z step -> resample). */
APGS_API int apgs_gather_ancestors_via(const int32_t* anc32, int64_t S, int64_t B, const apgs_latent_t* latents,
                              const int32_t* const* first, int n_latents, void* stream);

/* Composition of two ancestor maps, out[s, b] = prev[next[s, b], b]: what two consecutive resampling steps do to a latent
 *   that was never gathered in between.  With it a whole sweep runs WITHOUT gathers: every block update reads its inputs
 *   through (composed) ancestor maps -- apgs_gmm_z_step_fused_tab_f32 (src_ancestors, src_ancestors_z),
 *   apgs_gmm_eta_posterior_stats_labels_via_f32, apgs_gmm_draw_logw_eta_stats_rng_via_f32 -- and writes its outputs in place
 *   order; a latent is at most two resampling steps old when it is read (apgs/gmm/objectives.py:37-45).             */
APGS_API int apgs_compose_ancestors(const int32_t* prev, const int32_t* next, int64_t S, int64_t B, int32_t* out, void* stream);

/* R5  fused resample = R2/R3 + R4.  replaces resample_variables,
 *     apgs/gmm/objectives.py:229-234, apgs/dmm/objectives.py:190-195, apgs/bmnist/objectives.py:5-9.
 *   workspace: apgs_resample_workspace_bytes(S,B) (holds the int32 ancestors + CDF scratch).  */
APGS_API size_t apgs_resample_workspace_bytes(int64_t S, int64_t B);
APGS_API int apgs_resample_f32(const float* log_w, const float* w_norm, int64_t S, int64_t B,
                      int strategy, const void* uniforms, int64_t u_stride_s, int64_t u_stride_b,
                      int cdf_mode, const apgs_latent_t* latents, int n_latents,
                      int64_t* anc64, float* w_out, float* ess, float* lse,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ==============================================================================================
 * GMM reweighting.  K clusters, D data dims, N points.
 * Instantiated (K,D): (3,2) (2,2) (4,2) (5,2) (3,3) (4,3) (3,1); anything else -> APGS_E_UNSUPPORTED.
 * ============================================================================================== */

/* G1  neural sufficient statistics -> Normal-Gamma posterior parameters.
 * replaces data_to_stats + posterior_eta, apgs/gmm/kls_gmm.py:92-119 (called from
 * Enc_rws_eta.forward apgs/gmm/models.py:27 and Enc_apg_eta.forward :66), incl. quirks Q1/Q2.
 *   xp (S,B,N,D) = Linear(ob|z), gamma (S,B,N,K) = softmax(Linear(ob|z)); prior_* (K,D);
 *   post_* (S,B,K,D) out.  No (S,B,N,K,D) temporaries are built.                              */
APGS_API int apgs_gmm_nss_posterior_f32(const float* xp, const float* gamma,
                               int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                               const float* prior_alpha, const float* prior_beta,
                               const float* prior_mu, const float* prior_nu,
                               float* post_alpha, float* post_beta, float* post_mu, float* post_nu,
                               void* stream);

/* Which (K clusters, D data dims) the templated GMM kernels are compiled for.  Bit 0: G1-G4 (nss_posterior, logw_*);
 * bit 1: the fused z step (apgs_gmm_z_step_fused*); bit 2: the fused eta heads (apgs_gmm_eta_posterior_*); bit 3: the
 * label-mode entries (*_labels_*, K*D <= 8).
 * Entry points called with other shapes return APGS_E_UNSUPPORTED; the Python host side consults this and routes such
 * shapes through stock PyTorch CUDA ops (apgs_b200/gmm/torch_path.py).  apgs_dmm_supported: bit 0 the fused local
 * encoder, bit 1 the decoder kernel (the other DMM kernels take K, D at run time).                                   */
APGS_API int apgs_gmm_supported(int64_t K, int64_t D);
APGS_API int apgs_dmm_supported(int64_t K, int64_t D);

typedef struct apgs_gmm_prior {       /* HOST struct of DEVICE pointers; apgs/gmm/models.py:124-128 */
    const float *alpha, *beta, *mu, *nu;   /* (K,D) each */
    const float *pi;                       /* (K)        */
} apgs_gmm_prior_t;
typedef struct apgs_gmm_post {        /* proposal parameters from G1, (S,B,K,D) each */
    const float *alpha, *beta, *mu, *nu;
} apgs_gmm_post_t;
typedef struct apgs_gmm_post_out {    /* the same four tensors as OUTPUTS of a kernel */
    float *alpha, *beta, *mu, *nu;
} apgs_gmm_post_out_t;
typedef struct apgs_gmm_out {         /* every field optional (NULL = not wanted) */
    float *log_w;        /* (S,B)   log importance weight of the step                               */
    float *log_q_f;      /* (S,B)   forward proposal log-density (the loss term)                    */
    float *density;      /* (S,B)   the step's contribution to trace['density']                     */
    float *log_w_f;      /* (S,B)   forward / backward halves of log_w (parity checks)              */
    float *log_w_b;
    float *log_w_local;  /* (S,B,N) z step only: per-point weight (objectives.py:216)                */
    float *log_q_local;  /* (S,B,N) z step only: per-point forward log q(z_n)                       */
} apgs_gmm_out_t;

/* x is (B,N,D): ONE copy per instance -- the reference's S-fold replication (apg_training.py:19)
 * is not needed.  z tensors are one-hot fp32 (S,B,N,K); tau/mu (S,B,K,D); q_probs (S,B,N,K).       */

/* G4a one-shot weight.  replaces the arithmetic of oneshot, apgs/gmm/objectives.py:74-110.         */
APGS_API int apgs_gmm_logw_oneshot_f32(const float* x, const apgs_gmm_post_t* post, const apgs_gmm_prior_t* prior,
                              const float* tau, const float* mu, const float* q_probs, const float* z,
                              int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                              const apgs_gmm_out_t* out, void* stream);
/* G2  eta step.  replaces apg_update_eta arithmetic, apgs/gmm/objectives.py:161-195
 *     (+ Generative.eta_prior models.py:139-150, Generative.log_prob(aggregate=True) :172-185).    */
APGS_API int apgs_gmm_logw_eta_f32(const float* x, const float* z, const apgs_gmm_post_t* post,
                          const apgs_gmm_prior_t* prior, const float* tau, const float* mu,
                          const float* tau_old, const float* mu_old,
                          int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                          const apgs_gmm_out_t* out, void* stream);
/* G3  z step (+W2 sum over points).  replaces apg_update_z arithmetic, objectives.py:197-227
 *     (+ z_prior models.py:167-170, log_prob(aggregate=False)), eps-clamped categorical (Q4).      */
APGS_API int apgs_gmm_logw_z_f32(const float* x, const float* tau, const float* mu, const float* q_probs,
                        const apgs_gmm_prior_t* prior, const float* z, const float* z_old,
                        int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                        const apgs_gmm_out_t* out, void* stream);
/* G4b joint step.  replaces apg_update_joint arithmetic, objectives.py:112-158.                    */
APGS_API int apgs_gmm_logw_joint_f32(const float* x, const apgs_gmm_post_t* post, const apgs_gmm_prior_t* prior,
                            const float* tau, const float* mu, const float* tau_old, const float* mu_old,
                            const float* q_probs, const float* z, const float* z_old,
                            int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                            const apgs_gmm_out_t* out, void* stream);

/* N1+G3  fused z step: Enc_apg_z.forward (apgs/gmm/models.py:102-115: per-(n,k)
 *     MLP 3D -> H -> 1 with tanh, softmax over k) + categorical draw by inverse CDF from the supplied
 *     uniforms + the z-step weight of G3, in one pass (no (S,B,N,K,H) activations, no q_probs round trip).
 *   W1 (H,3D) b1 (H) W2 (H) b2 (1): pi_log_prob.{0,2}.{weight,bias} of the reference module.
 *   z_old    (S,B,N,K) one-hot or NULL (NULL: sample-only mode, used by the one-shot / joint steps).
 *   z_forced (S,B,N,K) one-hot or NULL (teacher forcing for parity tests: use instead of sampling).
 *   uniforms (S,B,N) fp32 in [0,1) (required unless z_forced): z_n = #{k : cdf_k <= u_n * sum_k q_nk}.
 *   z_new    (S,B,N,K) out;  q_probs (S,B,N,K) out or NULL;  out: as G3 (may be NULL).             */
APGS_API int apgs_gmm_z_step_fused_f32(const float* x, const float* tau, const float* mu,
                              const float* W1, const float* b1, const float* W2, const float* b2, int64_t H,
                              const apgs_gmm_prior_t* prior, const float* z_old, const float* z_forced,
                              const float* uniforms,
                              int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                              float* z_new, float* q_probs, const apgs_gmm_out_t* out, void* stream);

/* N1b+G1  fused eta heads: the two Linear layers of Enc_apg_eta / Enc_rws_eta (apgs/gmm/models.py:17-22,56-65),
 *     the softmax over k and G1 in one pass over z -- cat(ob, z), gamma and x' never touch HBM.
 *   Wg (K, D+K) bg (K) Wo (D, D+K) bo (D) = gamma.0.{weight,bias}, ob.0.{weight,bias};
 *   z (S,B,N,K) or NULL (Enc_rws_eta: the heads see x only; then Wg is (K,D), Wo (D,D)).        */
APGS_API int apgs_gmm_eta_posterior_fused_f32(const float* x, const float* z,
                                     const float* Wg, const float* bg, const float* Wo, const float* bo,
                                     int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                                     const float* prior_alpha, const float* prior_beta,
                                     const float* prior_mu, const float* prior_nu,
                                     float* post_alpha, float* post_beta, float* post_mu, float* post_nu,
                                     void* stream);

/* G1+G2 from statistics.  The eta step scores the SAME assignment z under the new and the old (tau, mu)
 *     (apgs/gmm/objectives.py:163-180), so its two log-likelihoods (apgs/gmm/models.py:172-185) are functions
 *     of the one-hot statistics of z.  apgs_gmm_eta_posterior_stats_f32 is apgs_gmm_eta_posterior_fused_f32
 *     that also writes, from the same pass over z,
 *       stats (S,B,K,1+2D):  [k] = { sum_n z_nk,  sum_n z_nk (x_nd - c_kd) (d<D),  sum_n z_nk (x_nd - c_kd)^2 (d<D) }
 *     centred at `center` (S,B,K,D) -- pass the OLD mu, which makes the backward term cancellation-free;
 *     apgs_gmm_logw_eta_stats_f32 is apgs_gmm_logw_eta_f32 evaluated from those statistics: O(K D) work and
 *     ~300 bytes per particle, no second pass over z.                                                 */
APGS_API int apgs_gmm_eta_posterior_stats_f32(const float* x, const float* z,
                                     const float* Wg, const float* bg, const float* Wo, const float* bo,
                                     int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                                     const float* prior_alpha, const float* prior_beta,
                                     const float* prior_mu, const float* prior_nu,
                                     float* post_alpha, float* post_beta, float* post_mu, float* post_nu,
                                     const float* center, float* stats, void* stream);
APGS_API int apgs_gmm_logw_eta_stats_f32(const float* stats, const float* center, const apgs_gmm_post_t* post,
                                const apgs_gmm_prior_t* prior, const float* tau, const float* mu,
                                const float* tau_old, const float* mu_old,
                                int64_t S, int64_t B, int64_t K, int64_t D,
                                const apgs_gmm_out_t* out, void* stream);

/* apgs_gmm_logw_eta_stats_f32 with the draw fused in: given standard variates (torch._standard_gamma(alpha),
 *   torch.randn) it forms tau = max(g / beta, tiny) and mu = e / sqrt(nu tau) + q_mu with the IEEE operations of
 *   Gamma(alpha, beta).sample() / Normal(q_mu, 1/sqrt(q_nu tau)).sample() (apgs/gmm/models.py:69-78), writes them to
 *   tau / mu (S,B,K,D) and scores them -- seven elementwise launches per eta step less, bit-identical draws.          */
APGS_API int apgs_gmm_draw_logw_eta_stats_f32(const float* stats, const float* center, const apgs_gmm_post_t* post,
                                const apgs_gmm_prior_t* prior, const float* gamma_std, const float* normal_std,
                                float* tau, float* mu, const float* tau_old, const float* mu_old,
                                int64_t S, int64_t B, int64_t K, int64_t D,
                                const apgs_gmm_out_t* out, void* stream);

/* ... and with the variates made in the kernel: standard Gamma(alpha) by Marsaglia-Tsang (the method of ATen's sample_gamma)
 *   and standard normals by Box-Muller from Philox4x32-10 words keyed by philox_seed[0] (+ stream word philox_seed[1]) and
 *   counted by (particle, element) -- philox_seed is a DEVICE pointer to two 64-bit words the host draws from its own
 *   generator per launch (torch.randint: fresh on every CUDA-graph replay).  Saves the _standard_gamma / randn launches and
 *   both variate tensors; the draws are distribution-equal (not stream-equal) to torch's.                              */
APGS_API int apgs_gmm_draw_logw_eta_stats_rng_f32(const float* stats, const float* center, const apgs_gmm_post_t* post,
                                const apgs_gmm_prior_t* prior, const uint64_t* philox_seed,
                                float* tau, float* mu, const float* tau_old, const float* mu_old,
                                int64_t S, int64_t B, int64_t K, int64_t D,
                                const apgs_gmm_out_t* out, void* stream);

/* Label mode of the two per-sweep GMM kernels.  Inside a sweep the assignments never need their one-hot (S,B,N,K) fp32
 *   form: these entries take / return them as one byte per point, z_labels uint8 in [0,K) with `label_stride` >= N bytes
 *   between the rows of consecutive particles (pad it to 16 so the gather moves 16-byte vectors), which cuts the z row
 *   the resampling gather moves from 4*K*N to ~N bytes.  Same arithmetic and outputs as apgs_gmm_z_step_fused_f32 (with
 *   z_old given, sampling from `uniforms`) and apgs_gmm_eta_posterior_stats_f32; center / stats may both be NULL;
 *   z_old_labels may be NULL (sample-only: out->density = log p(x, z | eta) and out->log_q_f = log q(z) of the draw,
 *   out->log_w = 0) and z_new_onehot (S,B,N,K), when given, receives the one-hot form of the same draw.  The launch a
 *   sweep repeats (labels in, labels out, no one-hot / q_probs / per-point / log_w_f / log_w_b outputs) runs a lean
 *   compile-time variant, specialised further for the reference's hidden width 32 and N = 60 points.                 */
APGS_API int apgs_gmm_z_step_fused_labels_f32(const float* x, const float* tau, const float* mu,
                              const float* W1, const float* b1, const float* W2, const float* b2, int64_t H,
                              const apgs_gmm_prior_t* prior, const uint8_t* z_old_labels, const float* uniforms,
                              int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                              uint8_t* z_new_labels, int64_t label_stride, float* z_new_onehot,
                              float* q_probs, const apgs_gmm_out_t* out, void* stream);
/* The label-mode z step that ALSO emits the next eta step's inputs (G3 + W2 + N1 + G1 in one launch): besides z_new_labels
 *   and out->{log_w, log_q_f, density} it writes, for the labels just drawn, Enc_apg_eta's Normal-Gamma posterior
 *   parameters (apgs/gmm/models.py:56-66 -> kls_gmm.py:92-119; heads Wg (K,D+K), bg (K), Wo (D,D+K), bo (D); prior->
 *   alpha/beta/mu/nu) to post_out (S,B,K,D each) and the one-hot statistics of z_new centred at this particle's mu to
 *   stats_out (S,B,K,1+2D) -- bit-identical to apgs_gmm_eta_posterior_stats_labels_f32(z_new_labels, center = mu).
 *   They are per-particle functions of (z, mu), so the resampling gather that follows the z step moves them with the
 *   particle (apgs_gather_ancestors copies rows) and the eta step of the next sweep needs no pass over the labels:
 *   apgs/gmm/objectives.py:161-166 costs no launch of its own.  K*D <= 8; lean label-mode launch only.               */
APGS_API int apgs_gmm_z_step_fused_labels_eta_f32(const float* x, const float* tau, const float* mu,
                              const float* W1, const float* b1, const float* W2, const float* b2, int64_t H,
                              const apgs_gmm_prior_t* prior, const uint8_t* z_old_labels, const float* uniforms,
                              int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                              uint8_t* z_new_labels, int64_t label_stride,
                              const float* Wg, const float* bg, const float* Wo, const float* bo,
                              const apgs_gmm_post_out_t* post_out, float* stats_out,
                              const apgs_gmm_out_t* out, void* stream);
/* The z step with the draw's uniforms made INSIDE the kernel (replaces `Categorical(probs).sample()`,
 *   apgs/gmm/models.py:115-117, whose random stream is torch's own; the draw is distributed identically, the stream is
 *   Philox4x32-10 keyed by *philox_seed -- a device word the host fills from torch's generator per launch, so that manual
 *   seeds reproduce and CUDA-graph replays get fresh draws -- and counted by (particle row, lane, pass)).  One entry for every
 *   mode of the three above: z_old as one-hot OR labels OR neither (sample-only), z_new as one-hot and / or labels, and
 *   the eta statistics when Wg..stats_out are given (then labels in / labels out only).  Saves the (S,B,N) uniforms:
 *   one launch writing and one pass reading 4 N bytes per particle.                                                   */
APGS_API int apgs_gmm_z_step_fused_rng_f32(const float* x, const float* tau, const float* mu,
                              const float* W1, const float* b1, const float* W2, const float* b2, int64_t H,
                              const apgs_gmm_prior_t* prior, const float* z_old_onehot, const uint8_t* z_old_labels,
                              const uint64_t* philox_seed, int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                              float* z_new_onehot, uint8_t* z_new_labels, int64_t label_stride, float* q_probs,
                              const float* Wg, const float* bg, const float* Wo, const float* bo,
                              const apgs_gmm_post_out_t* post_out, float* stats_out,
                              const apgs_gmm_out_t* out, void* stream);
/* Per-instance tables of the z step.  Everything Enc_apg_z.forward (apgs/gmm/models.py:102-115) computes from the DATA
 *   and the encoder WEIGHTS alone -- the data half of the first Linear (W1[:, :D] x_n, tabulated as 2^(2 log2(e) .)), the
 *   output weights, b2 + sum_h W2[h], the clamped log prior (apgs/gmm/models.py:168-170) and the points -- is the same for
 *   every particle of an instance and for every sweep of an objective (the reference re-evaluates it S times per sweep
 *   on the repeated data, apgs/gmm/apg_training.py:19).  apgs_gmm_z_tables_f32 writes it once per instance
 *   (apgs_gmm_z_tables_bytes(B, N, H, K, D) bytes, 16-byte aligned) and apgs_gmm_z_step_fused_tab_f32 -- the in-kernel-
 *   uniform entry with one more argument -- fetches an instance's block with a single bulk copy per CTA instead of
 *   rebuilding it.  tables = NULL: identical to apgs_gmm_z_step_fused_rng_f32.  Results are bit-identical either way.
 *   src_ancestors (S,B) int32 or NULL: the INPUT latents (tau, mu, z_old) of particle (s,b) are read from row
 *   (src_ancestors[s,b], b) -- the ancestors of a resampling step whose gather is deferred (apgs_gather_ancestors_via);
 *   outputs and the random stream are those of particle (s,b).  Bit-identical to gathering first.  src_ancestors_z: the map
 *   for z_old alone when the assignments went through other resampling steps than (tau, mu) (NULL: src_ancestors).
 *   The tables depend on (x, W1, b1, W2, b2, prior->pi): rebuild them when any of these changes.                       */
APGS_API size_t apgs_gmm_z_tables_bytes(int64_t B, int64_t N, int64_t H, int64_t K, int64_t D);
APGS_API int apgs_gmm_z_tables_f32(const float* x, const float* W1, const float* b1, const float* W2, const float* b2,
                              int64_t H, const apgs_gmm_prior_t* prior, int64_t B, int64_t N, int64_t K, int64_t D,
                              float* tables, void* stream);
APGS_API int apgs_gmm_z_step_fused_tab_f32(const float* x, const float* tau, const float* mu,
                              const float* W1, const float* b1, const float* W2, const float* b2, int64_t H,
                              const apgs_gmm_prior_t* prior, const float* z_old_onehot, const uint8_t* z_old_labels,
                              const uint64_t* philox_seed, int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                              float* z_new_onehot, uint8_t* z_new_labels, int64_t label_stride, float* q_probs,
                              const float* Wg, const float* bg, const float* Wo, const float* bo,
                              const apgs_gmm_post_out_t* post_out, float* stats_out,
                              const apgs_gmm_out_t* out, const float* tables, const int32_t* src_ancestors,
                              const int32_t* src_ancestors_z, void* stream);
/* Measurement only: the main loop of the fused z step alone (K = 3, H = 32; `ctas` CTAs of 32 particles x 8 lanes, `passes`
 *   passes of eight quads of hidden units each, synthetic tables) -- the ceiling of the step's MLP part on this GPU, timed
 *   by bench.py next to the step itself (roofline.mlp_loop_alone).  out: ctas * 256 floats (a checksum per thread).      */
APGS_API int apgs_gmm_z_loop_probe_f32(int64_t ctas, int64_t passes, float* out, void* stream);
/* The two eta-step entries with their INPUT latents read through ancestor maps (S,B) int32 (NULL = in place), see
 *   apgs_compose_ancestors: labels from row (z_ancestors[s,b], b), centre from row (center_ancestors[s,b], b);
 *   center / tau_old / mu_old from row (old_ancestors[s,b], b).  Outputs, statistics, posterior parameters and the random
 *   stream belong to particle (s,b).  Bit-identical to gathering first.                                                 */
APGS_API int apgs_gmm_eta_posterior_stats_labels_via_f32(const float* x, const uint8_t* z_labels,
                                     const float* Wg, const float* bg, const float* Wo, const float* bo,
                                     int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                                     const float* prior_alpha, const float* prior_beta,
                                     const float* prior_mu, const float* prior_nu,
                                     float* post_alpha, float* post_beta, float* post_mu, float* post_nu,
                                     const float* center, float* stats, int64_t label_stride,
                                     const int32_t* z_ancestors, const int32_t* center_ancestors, const float* tables,
                                     void* stream);
/* Per-instance tables of the eta step: Enc_apg_eta's two Linear heads on cat(x_n, e_label) (apgs/gmm/models.py:56-65) take
 *   only K values per data point when z is one-hot; { x'_n, gamma_n } per (point, label) and the points themselves are the
 *   same for every particle of an instance and every sweep of an objective.  apgs_gmm_eta_tables_f32 writes them once
 *   (apgs_gmm_eta_tables_bytes(B, N, K, D) bytes, 16-byte aligned; 0: shape not supported) and
 *   apgs_gmm_eta_posterior_stats_labels_via_f32(tables = ...) fetches an instance's block with one bulk copy per CTA instead
 *   of tabulating it; tables = NULL: the CTA builds its own.  Bit-identical either way.  Depends on (x, Wg, bg, Wo, bo).  */
APGS_API size_t apgs_gmm_eta_tables_bytes(int64_t B, int64_t N, int64_t K, int64_t D);
APGS_API int apgs_gmm_eta_tables_f32(const float* x, const float* Wg, const float* bg, const float* Wo, const float* bo,
                                     int64_t B, int64_t N, int64_t K, int64_t D, float* tables, void* stream);
APGS_API int apgs_gmm_draw_logw_eta_stats_rng_via_f32(const float* stats, const float* center, const apgs_gmm_post_t* post,
                                     const apgs_gmm_prior_t* prior, const uint64_t* philox_seed, float* tau, float* mu,
                                     const float* tau_old, const float* mu_old, const int32_t* old_ancestors,
                                     int64_t S, int64_t B, int64_t K, int64_t D, const apgs_gmm_out_t* out, void* stream);
APGS_API int apgs_gmm_eta_posterior_stats_labels_f32(const float* x, const uint8_t* z_labels,
                                     const float* Wg, const float* bg, const float* Wo, const float* bo,
                                     int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                                     const float* prior_alpha, const float* prior_beta,
                                     const float* prior_mu, const float* prior_nu,
                                     float* post_alpha, float* post_beta, float* post_mu, float* post_nu,
                                     const float* center, float* stats, int64_t label_stride, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Training mode.  The APG loss is  sum_s w_s.detach() * (-log q_s)
 * (apgs/gmm/objectives.py:183,220), so autograd needs only the gradients of the FORWARD proposal log-densities.
 *
 * apgs_gmm_z_logq_backward_f32: gradients of  sum_{s,b,n} grad_logq[s,b,n] * log q(z_n | x_n, tau, mu)  w.r.t. the
 *   Enc_apg_z MLP (apgs/gmm/models.py:97-115: W1 (H,3D), b1 (H), W2 (H), b2 (1)), where log q is
 *   OneHotCategorical(probs=q_probs).log_prob(z) (torch's eps clamp: zero gradient outside [eps, 1-eps]).  q_probs
 *   (S,B,N,K) are the forward kernel's (apgs_gmm_z_step_fused*_f32); the assignment comes one-hot (S,B,N,K) or as labels;
 *   grad_logq element (r = s*B+b, n) sits at grad_logq[r*g_stride_particle + n*g_stride_point] (g_stride_point = 0: one
 *   value per particle, the one-shot's sum over points, objectives.py:92-95).  Any of gW1/gb1/gW2/gb2 may be NULL.
 *   Deterministic (no atomics).  workspace: apgs_gmm_z_logq_backward_workspace_bytes(S, B, H, D), 256-byte aligned.
 *
 * apgs_gmm_label_table_sum_f32 / _backward_f32: with one-hot z the two Linear heads of Enc_apg_eta on cat(x_n, z_n)
 *   (models.py:56-66) take only K values per data point; the host tabulates any per-(instance, point, label) row
 *   table (B,N,K,C) with differentiable PyTorch ops and these kernels do the only S-sized work of the eta step's graph:
 *       out[s,b,c] = sum_n table[b, n, label(s,b,n), c]            (the neural sufficient statistics, kls_gmm.py:92-104)
 *       grad_table[b,n,l,c] = sum_s [label(s,b,n) == l] grad_out[s,b,c]
 *   C <= 64, K <= 255.  The adjoint splits the particles of an instance over P = apgs_gmm_label_table_sum_backward_splits
 *   (S, B) CTAs: grad_table is (P,B,N,K,C), P partial tables the caller adds up (deterministic, no atomics).           */
/* log q(tau, mu) of Enc_apg_eta from its neural statistics nss (S,B,K+2KD) = {s1_k, s2_kd, s3_kd} -- the Normal-Gamma update
 * (apgs/gmm/kls_gmm.py:106-119, quirks Q1 / Q2) followed by Gamma / Normal log-densities summed over (K,D)
 * (apgs/gmm/objectives.py:167) -- and its Jacobian jac (S,B,K+2KD) = d log_q / d nss, in one launch: the eta step's graph
 * is then  head table (PyTorch, per instance) -> apgs_gmm_label_table_sum_f32 -> this.                                   */
APGS_API int apgs_gmm_eta_logq_jac_f32(const float* nss, const float* tau, const float* mu,
                              const float* prior_alpha, const float* prior_beta, const float* prior_mu, const float* prior_nu,
                              int64_t S, int64_t B, int64_t K, int64_t D, float* log_q, float* jac, void* stream);
APGS_API size_t apgs_gmm_z_logq_backward_workspace_bytes(int64_t S, int64_t B, int64_t H, int64_t D);
APGS_API int apgs_gmm_z_logq_backward_f32(const float* x, const float* tau, const float* mu,
                              const float* W1, const float* b1, const float* W2, int64_t H,
                              const float* q_probs, const float* z_onehot, const uint8_t* z_labels, int64_t label_stride,
                              const float* grad_logq, int64_t g_stride_particle, int64_t g_stride_point,
                              int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                              float* gW1, float* gb1, float* gW2, float* gb2,
                              void* workspace, size_t workspace_bytes, void* stream);
APGS_API int apgs_gmm_label_table_sum_f32(const float* table, const float* z_onehot, const uint8_t* z_labels,
                              int64_t label_stride, int64_t S, int64_t B, int64_t N, int64_t K, int64_t C,
                              float* out, void* stream);
APGS_API int apgs_gmm_label_table_sum_backward_splits(int64_t S, int64_t B);
APGS_API int apgs_gmm_label_table_sum_backward_f32(const float* grad_out, const float* z_onehot, const uint8_t* z_labels,
                              int64_t label_stride, int64_t S, int64_t B, int64_t N, int64_t K, int64_t C,
                              float* grad_table, void* stream);

/* ==============================================================================================
 * Bouncing-MNIST.  FP frame pixels (96), DP digit pixels (28), K digits,
 * T frames.  The transformer is the reference's: nearest sampling, align_corners=True, diagonal theta
 * (SURVEY.md F5).  lin_* = torch.linspace(-1, 1, n) (affine_grid's base axis) on the device.
 * ============================================================================================== */

/* B1  crop.  replaces Affine_Transformer.frame_to_digit, apgs/bmnist/affine_transformer.py:54-65.
 *   frames: element (s,b,t,i,j) at frames[s*stride_s + b*stride_b + t*stride_t + i*FP + j] -- stride_s may be 0
 *   (the reference's S-fold repeat, apgs/bmnist/apg_training.py:25, is not needed);
 *   z_where (S,B,T,K,2) (not modified); lin_dp (DP); digits_out (S,B,T,K,DP,DP).                   */
APGS_API int apgs_bmnist_crop_f32(const float* frames, int64_t stride_s, int64_t stride_b, int64_t stride_t,
                         const float* z_where, const float* lin_dp,
                         int64_t S, int64_t B, int64_t T, int64_t K, int64_t FP, int64_t DP,
                         float* digits_out, void* stream);
/* B2  paste.  replaces Affine_Transformer.digit_to_frame, affine_transformer.py:41-52.
 *   digits (S,B,K,DP,DP); lin_fp (FP); frames_out (S,B,T,K,FP,FP) (zero outside the digit).         */
APGS_API int apgs_bmnist_paste_f32(const float* digits, const float* z_where, const float* lin_fp,
                          int64_t S, int64_t B, int64_t T, int64_t K, int64_t FP, int64_t DP,
                          float* frames_out, void* stream);
/* Bilinear mode of B1 / B2 (north_star item 4; the reference itself samples with mode='nearest', SURVEY.md F5, which is
 * the parity-checked mode above): grid_sample(mode='bilinear', padding_mode='zeros', align_corners=True) semantics on
 * the same diagonal affine grid.  Same arguments and layouts as apgs_bmnist_crop_f32 / apgs_bmnist_paste_f32.         */
APGS_API int apgs_bmnist_crop_bilinear_f32(const float* frames, int64_t stride_s, int64_t stride_b, int64_t stride_t,
                         const float* z_where, const float* lin_dp,
                         int64_t S, int64_t B, int64_t T, int64_t K, int64_t FP, int64_t DP,
                         float* digits_out, void* stream);
APGS_API int apgs_bmnist_paste_bilinear_f32(const float* digits, const float* z_where, const float* lin_fp,
                          int64_t S, int64_t B, int64_t T, int64_t K, int64_t FP, int64_t DP,
                          float* frames_out, void* stream);
/* B2+B3  paste + sum over digits + clamp + Bernoulli-like pixel log-likelihood, fused.
 * replaces the tail of Dec_digit.forward, apgs/bmnist/models.py:148-154, and MBern_log_prob :156-162.
 *   ll_out (S,B,T); recon_out (S,B,T,FP,FP) or NULL.  The (S,B,T,K,FP,FP) pasted tensor never exists.  */
APGS_API int apgs_bmnist_paste_ll_f32(const float* digits, const float* z_where, const float* lin_fp,
                             const float* frames, int64_t stride_s, int64_t stride_b, int64_t stride_t,
                             int64_t S, int64_t B, int64_t T, int64_t K, int64_t FP, int64_t DP,
                             float* ll_out, float* recon_out, void* stream);
/* B4a  z_where step of one frame.  replaces the log-density bookkeeping of propose_one_movement,
 * apgs/bmnist/objectives.py:60-107 (+ Dec_coor.forward, apgs/bmnist/models.py:97-104) and the weight at
 * objectives.py:198-206.   q_mean/q_std/z_new (S,B,K,D) of frame t; z_prev = z_new of frame t-1 or NULL (t=0);
 * z_old / z_old_prev: the old trajectory at t and t-1 (NULL: no backward term / t=0); prior_* (D);
 * ll_f / ll_b (S,B) or NULL.  Outputs (S,B), each optional: log_w, log_q_f, log_p_f, log_q_b, log_p_b. */
APGS_API int apgs_bmnist_logw_where_f32(const float* q_mean, const float* q_std, const float* z_new,
                               const float* z_prev, const float* z_old, const float* z_old_prev,
                               const float* prior_mu0, const float* prior_sigma0, const float* prior_sigmat,
                               const float* ll_f, const float* ll_b,
                               int64_t S, int64_t B, int64_t K, int64_t D,
                               float* log_w, float* log_q_f, float* log_p_f, float* log_q_b, float* log_p_b,
                               void* stream);
/* B4b  z_what step / one-shot.  replaces objectives.py:241 (apg_what) and :147-149 (oneshot).
 *   q_mu/q_std/z_new (S,B,K,Z); z_old or NULL; ll_f/ll_b (S,B,T); extra_log_p / extra_log_q (S,B) or NULL
 *   (the one-shot's accumulated log p / log q of z_where).  Outputs (S,B), each optional.           */
APGS_API int apgs_bmnist_logw_what_f32(const float* q_mu, const float* q_std, const float* z_new, const float* z_old,
                              const float* ll_f, const float* ll_b,
                              const float* extra_log_p, const float* extra_log_q,
                              int64_t S, int64_t B, int64_t K, int64_t Z, int64_t T,
                              float* log_w, float* log_q_f, float* density, float* ll_sum, void* stream);

/* ==============================================================================================
 * DMM ring mixture.  K <= 16 clusters, D <= 8 data dims, N points.
 * ============================================================================================== */

/* D1  NSS pooling.  replaces apgs/dmm/models.py:134-139 (Enc_apg_mu) and :38-43 (Enc_rws_mu):
 *   out[s,b,k,:J] = sum_n nss1[s,b,n,:] * nss2[s,b,n,k] / (sum_n nss2[s,b,n,k] + 1e-8),
 *   out[s,b,k,J:] = cat(prior_mu_mu, prior_mu_sigma);   nss1 (S,B,N,J), nss2 (S,B,N,K), out (S,B,K,J+2D).
 *   No (S,B,N,K,J) temporaries.                                                                     */
APGS_API int apgs_dmm_nss_pool_f32(const float* nss1, const float* nss2,
                          const float* prior_mu_mu, const float* prior_mu_sigma,
                          int64_t S, int64_t B, int64_t N, int64_t K, int64_t J, int64_t D,
                          float* out, void* stream);

typedef struct apgs_dmm_prior {       /* HOST struct of DEVICE pointers; apgs/dmm/models.py:163-172 */
    const float *mu_mu, *mu_sigma;    /* (D)  prior of the centres                    */
    const float *pi;                  /* (K)  prior of the assignments                */
    const float *con1, *con0;         /* (1)  Beta prior of the angles                */
    const float *radi;                /* (1)  decoder radius parameter                */
    const float *recon_sigma;         /* (1)  likelihood scale                        */
} apgs_dmm_prior_t;
typedef struct apgs_dmm_out {         /* every field optional */
    float *log_w, *log_q_f, *density, *ll_f, *log_w_f, *log_w_b;     /* (S,B)   */
    float *log_w_local, *log_q_local, *ll_local;                      /* (S,B,N) */
} apgs_dmm_out_t;

/* x (B,N,D) one copy per instance; z one-hot (S,B,N,K); hidden2 (S,B,N,D) = unit-normalised decoder MLP
 * output (apgs/dmm/models.py:202-203); beta / con1 / con0 (S,B,N[,1]); mu / q_mu / q_sigma (S,B,K,D).      */

/* D2  mu step.  replaces the arithmetic of apg_update_mu, apgs/dmm/objectives.py:116-151.          */
APGS_API int apgs_dmm_logw_mu_f32(const float* x, const float* z, const float* hidden2,
                         const float* q_mu, const float* q_sigma, const float* mu, const float* mu_old,
                         const apgs_dmm_prior_t* prior,
                         int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                         const apgs_dmm_out_t* out, void* stream);
/* D3  {z, beta} step per point (+ sum over points).  replaces apg_update_local, objectives.py:153-188.  */
APGS_API int apgs_dmm_logw_local_f32(const float* x, const float* mu, const float* q_probs,
                            const float* z, const float* beta, const float* con1, const float* con0,
                            const float* hidden2,
                            const float* z_old, const float* beta_old, const float* con1_old,
                            const float* con0_old, const float* hidden2_old,
                            const apgs_dmm_prior_t* prior,
                            int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                            const apgs_dmm_out_t* out, void* stream);
/* one-shot weight.  replaces the arithmetic of oneshot, objectives.py:80-114.                       */
APGS_API int apgs_dmm_logw_oneshot_f32(const float* x, const float* q_mu, const float* q_sigma, const float* mu,
                              const float* q_probs, const float* z, const float* beta,
                              const float* con1, const float* con0, const float* hidden2,
                              const apgs_dmm_prior_t* prior,
                              int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                              const apgs_dmm_out_t* out, void* stream);

/* ---- fused DMM encoders / decoder ("the analogous Enc_apg_local") -------------- */
typedef struct apgs_dmm_mlp {         /* HOST struct of DEVICE pointers: nn.Sequential(Linear(i,H), act, Linear(H,o)) */
    const float *W0, *b0;             /* (H, i), (H)   = <name>.0.weight / .0.bias   */
    const float *W2, *b2;             /* (o, H), (o)   = <name>.2.weight / .2.bias   */
} apgs_dmm_mlp_t;

/* D1 for Enc_apg_mu with one-hot z (apgs/dmm/models.py:129-139).  z is one-hot, so nss1/nss2 of cat(x_n, z_n)
 *   take K distinct values per data point; the caller tabulates them once per INSTANCE,
 *     tab1 (B,N,K,J) = nss1(cat(x_n, e_l)),  tab2 (B,N,K,K) = softmax(nss2(cat(x_n, e_l))),  l = label,
 *   and this entry pools them over the labels argmax(z) (first maximum): out (S,B,K,J+2D) as apgs_dmm_nss_pool_f32. */
APGS_API int apgs_dmm_nss_pool_onehot_f32(const float* tab1, const float* tab2, const float* z,
                                 const float* prior_mu_mu, const float* prior_mu_sigma,
                                 int64_t S, int64_t B, int64_t N, int64_t K, int64_t J, int64_t D,
                                 float* out, void* stream);

/* Enc_apg_local.forward (apgs/dmm/models.py:76-101) in one pass:
 *   q_probs (S,B,N,K) = softmax_k pi_log_prob(x_n - mu_k)                               [:78-79]
 *   z       (S,B,N,K) = z_forced, or the inverse-CDF draw of q_probs at uniforms (S,B,N) [:81]
 *   con1, con0 (S,B,N) = exp(angle_log_con{1,0}(x_n - mu_{z_n}))                         [:86-88]
 *   con1_old, con0_old = the same at z_old (S,B,N,K), when given (backward kernel)       [:99-101]
 *   nets: pi (D->H->ReLU->1), con1 / con0 (D->H->tanh->1), all with the same H.                              */
APGS_API int apgs_dmm_local_encoder_fused_f32(const float* x, const float* mu,
                                 const apgs_dmm_mlp_t* pi, const apgs_dmm_mlp_t* con1_net, const apgs_dmm_mlp_t* con0_net,
                                 int64_t H, const float* uniforms, const float* z_forced, const float* z_old,
                                 int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                                 float* q_probs, float* z, float* con1, float* con0,
                                 float* con1_old, float* con0_old, void* stream);

/* The same kernel with the draws made inside it: beta ~ Beta(con1, con0) (`Beta(...).sample()`, apgs/dmm/models.py:90-91; as
 *   torch draws it -- two standard Gammas by Marsaglia-Tsang, normalised, clamped to [tiny, 1 - eps] -- from Philox4x32-10
 *   keyed by the two 64-bit words at philox_seed, which the host takes from its generator per launch) -> beta (S,B,N), and,
 *   when `decoder` (recon_mu, 1 -> decoder_hidden -> tanh -> D) and hidden2 (S,B,N,D) are given, the decoder's unit-normalised
 *   output of that beta (apgs/dmm/models.py:202-203; apgs_dmm_decoder_hidden2_f32's arithmetic).  uniforms may be NULL: the
 *   categorical draw then takes its uniform from the same Philox stream.  Distribution-equal, not stream-equal, to torch.   */
APGS_API int apgs_dmm_local_encoder_draw_f32(const float* x, const float* mu,
                                 const apgs_dmm_mlp_t* pi, const apgs_dmm_mlp_t* con1_net, const apgs_dmm_mlp_t* con0_net,
                                 int64_t H, const float* uniforms, const float* z_forced, const float* z_old,
                                 const uint64_t* philox_seed, int64_t S, int64_t B, int64_t N, int64_t K, int64_t D,
                                 float* q_probs, float* z, float* con1, float* con0, float* con1_old, float* con0_old,
                                 float* beta, const apgs_dmm_mlp_t* decoder, int64_t decoder_hidden, float* hidden2,
                                 void* stream);

/* The two heads of Enc_rws_mu / Enc_apg_mu on the pooled statistics in one launch (apgs/dmm/models.py: mean_mu,
 *   mean_log_sigma -- Linear(num_nss + 2 D, H2) + Tanh + Linear(H2, D); forward :43-44, :88-89):
 *     q_mu (M,Z) = W2m tanh(W1m v + b1m) + b2m,   q_sigma (M,Z) = exp(W2s tanh(W1s v + b1s) + b2s),   v (M,H), M = S*B*K rows
 *   weights as nn.Linear holds them: w1_* (H2,H), b1_* (H2), w2_* (Z,H2), b2_* (Z).  A thread owns a row; fp32, sums in
 *   ascending k.  H <= 32, Z <= 4 (else APGS_E_UNSUPPORTED: the host keeps the nn.Sequential modules).                      */
APGS_API int apgs_dmm_mu_heads_f32(const float* v, const float* w1_mean, const float* b1_mean, const float* w2_mean,
                                   const float* b2_mean, const float* w1_sigma, const float* b1_sigma, const float* w2_sigma,
                                   const float* b2_sigma, int64_t M, int64_t H, int64_t H2, int64_t Z, float* q_mu,
                                   float* q_sigma, void* stream);

/* Decoder.recon_mu + unit normalisation (apgs/dmm/models.py:202-203): hidden2 (M,D) from beta (M), M = S*B*N;
 *   net = recon_mu (1->H->tanh->D).                                                                          */
APGS_API int apgs_dmm_decoder_hidden2_f32(const float* beta, const apgs_dmm_mlp_t* net, int64_t H, int64_t D,
                                 int64_t M, float* hidden2, void* stream);

/* frame_left update of the K sequential matches (apgs/bmnist/objectives.py:87-89):
 *     out[s,b] = frame[s,b] - digit_to_frame(digit[s,b], z_where[s,b])           (FP x FP, nearest paste, zero outside)
 *   frame (s,b) at frame + s*frame_stride_s + b*frame_stride_b (dense FP*FP; stride 0 allowed), digit (s,b) likewise
 *   (dense DP*DP, e.g. template[:, :, k] of (S,B,K,DP,DP)), z_where (S,B,2), out (S,B,FP,FP) dense.                  */
APGS_API int apgs_bmnist_paste_subtract_f32(const float* frame, int64_t frame_stride_s, int64_t frame_stride_b,
                                   const float* digit, int64_t digit_stride_s, int64_t digit_stride_b,
                                   const float* z_where, const float* lin_fp,
                                   int64_t S, int64_t B, int64_t FP, int64_t DP, float* out, void* stream);

/* out (M,N) = act(x (M,K) W (N,K)^T + bias (N)), act = ReLU when relu != 0: nn.Linear (+ nn.ReLU) for the one dense contraction
 *   on the bouncing-MNIST path, Enc_coor.enc_hidden = Linear(4761 -> 400) + ReLU on the softmax of the template-match map
 *   (apgs/bmnist/models.py:13-16,28; 150 calls per 5-sweep objective with M = S*B rows).  Tensor cores with fp32-level
 *   accuracy: 3xTF32 error compensation (operands split into two TF32 halves in registers, three mma.sync.m16n8k8 per
 *   product, fp32 accumulation) and deterministic split-K over the SMs (partial tiles in `workspace`,
 *   apgs_linear_tf32x3_workspace_bytes(M, N, K) bytes, summed in a fixed order by a second launch).  Rows need only
 *   4-byte alignment.  Results agree with an fp32 GEMM to ~1e-6 of sum_k |x_k w_k| (tests/test_bmnist_gpu.py).            */
/* The same 3xTF32 product through tcgen05 (5th-generation tensor cores): tcgen05.mma kind::tf32 on shared-memory operand halves,
 *   fp32 accumulator in tensor memory, split-K with the same deterministic reduction.  Workspace:
 *   apgs_linear_tc5_workspace_bytes(M, N, K).  Returns through the same status codes; an MMA pipeline that does not
 *   complete raises an error word in the workspace instead of hanging, and the reduction launch then writes NaN to `out`. */
APGS_API size_t apgs_linear_tc5_workspace_bytes(int64_t M, int64_t N, int64_t K);
APGS_API int apgs_linear_tc5_f32(const float* x, const float* W, const float* bias, int64_t M, int64_t K, int64_t N,
                                 int relu, float* out, void* workspace, size_t workspace_bytes, void* stream);
/* Packed operands for the tcgen05 form: both operands already split into their TF32 halves and laid out tile by tile in the
 *   MMA's shared-memory order, so a pipeline stage of the GEMM is two bulk copies and no register touches an operand.
 *   Layout (apgs_b200/csrc/tf32_pack.cuh): tiles of TR rows x 32 columns, TR = 128 for the activations (which = 0) and 80 for
 *   the weights (which = 1); tile (rt, c) at float offset (rt * ceil(K/32) + c) * 2*TR*32 as { hi[TR x 32], lo[TR x 32] },
 *   element (r, k) of a half at ((r % TR) / 8) * 256 + ((k % 32) / 4) * 32 + (r % 8) * 4 + k % 4; columns K .. 32 ceil(K/32) - 1
 *   are zero.  apgs_linear_tc5_packed_bytes(rows, K, which) is the size of such a buffer (16-byte aligned).
 *   apgs_linear_tc5_pack_f32 packs a row-major matrix (row stride ld >= K elements): once per set of weights -- Enc_coor's
 *   weight is constant over the 150 calls of an objective (apgs/bmnist/models.py:13-16) -- while the activations are written in
 *   this layout by apgs_bmnist_template_match_packed_f32.  apgs_linear_tc5_packed_f32 is then bit-identical to
 *   apgs_linear_tc5_f32 on the same operands (same MMAs in the same order, same split-K; the measurement knob
 *   APGS_TC5PK_STAGES=4 selects fewer, longer K slices).  Rows >= `rows` of the last activation tile may hold anything finite. */
APGS_API size_t apgs_linear_tc5_packed_bytes(int64_t rows, int64_t K, int which);
APGS_API int apgs_linear_tc5_pack_f32(const float* src, int64_t rows, int64_t K, int64_t ld, int which, float* dst, void* stream);
APGS_API size_t apgs_linear_tc5_packed_workspace_bytes(int64_t M, int64_t N, int64_t K);
APGS_API int apgs_linear_tc5_packed_f32(const float* x_packed, const float* w_packed, const float* bias, int64_t M, int64_t K,
                                        int64_t N, int relu, float* out, void* workspace, size_t workspace_bytes, void* stream);
APGS_API size_t apgs_linear_tf32x3_workspace_bytes(int64_t M, int64_t N, int64_t K);
APGS_API int apgs_linear_tf32x3_f32(const float* x, const float* W, const float* bias, int64_t M, int64_t K, int64_t N,
                                    int relu, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* Enc_coor's two heads on the hidden layer in one launch (apgs/bmnist/models.py:16-26 definitions, :31-32 forward):
 *     q_mean (M,Z) = tanh(W2m relu(W1m h + b1m) + b2m),   q_std (M,Z) = exp(W2s relu(W1s h + b1s) + b2s),   h = hidden (M,H)
 *   w1t: the two first layers concatenated, transposed and 4-k interleaved, [ceil(H/4)][2*H2][4] floats with
 *     w1t[k4][u][i] = W1cat[u][4*k4 + i], W1cat = [W1m (H2,H); W1s (H2,H)], zero for 4*k4 + i >= H (16-byte aligned; built once
 *     per set of weights by the host); b1 (2*H2) = [b1m; b1s]; w2_mean / w2_std (Z,H2); b2_mean / b2_std (Z).
 *   fp32, every sum in a fixed order: agrees with the nn.Sequential modules to the rounding of a different summation order.   */
APGS_API int apgs_bmnist_coor_heads_f32(const float* hidden, const float* w1t, const float* b1, const float* w2_mean,
                                        const float* b2_mean, const float* w2_std, const float* b2_std, int64_t M, int64_t H,
                                        int64_t H2, int64_t Z, float* q_mean, float* q_std, void* stream);

/* N3  bmnist template matching (apgs/bmnist/objectives.py:72-74):
 *     conv2d(frame_left (1, S*B, FP, FP), template_k (S*B, 1, DP, DP), groups = S*B) -> (S, B, CP*CP), CP = FP-DP+1,
 *     i.e. image (s,b) cross-correlated with ITS OWN template, optionally followed by the softmax over the CP*CP scores.
 *   image (s,b) at frame + s*frame_stride_s + b*frame_stride_b (FP*FP dense; stride 0 allowed for S-shared frames),
 *   template (s,b) at tmpl + s*tmpl_stride_s + b*tmpl_stride_b (DP*DP dense), strides in elements.              */
APGS_API int apgs_bmnist_template_match_f32(const float* frame, int64_t frame_stride_s, int64_t frame_stride_b,
                                   const float* tmpl, int64_t tmpl_stride_s, int64_t tmpl_stride_b,
                                   int64_t S, int64_t B, int64_t FP, int64_t DP, int softmax,
                                   float* out, void* stream);
/* The same with the softmax written as the packed activation operand of apgs_linear_tc5_packed_f32 (rows = S*B, K = CP*CP,
 *   which = 0; apgs_linear_tc5_packed_bytes(S*B, CP*CP, 0) bytes): the map only feeds Enc_coor's first layer
 *   (apgs/bmnist/objectives.py:74-75), so it never exists in fp32.  FP = 96, DP = 28 only (else APGS_E_UNSUPPORTED).        */
APGS_API int apgs_bmnist_template_match_packed_f32(const float* frame, int64_t frame_stride_s, int64_t frame_stride_b,
                                   const float* tmpl, int64_t tmpl_stride_s, int64_t tmpl_stride_b,
                                   int64_t S, int64_t B, int64_t FP, int64_t DP, float* out_packed, void* stream);
/* The match on the RESIDUAL frame of the sequential proposals (apgs/bmnist/objectives.py:72-89): the image of (s,b) is
 *     (frame - digit_to_frame(digit0, z_where0)) - digit_to_frame(digit1, z_where1)         (digit1 / digit0 may be NULL)
 *   formed pixel by pixel, in that order, while the image is staged in shared memory -- bit-identical to matching the output
 *   of apgs_bmnist_paste_subtract_f32, without the frame_left tensors ever existing.  digit p of (s,b) at
 *   digit_p + s*stride_s + b*stride_b (DP*DP dense), z_where_p (S,B,2) contiguous, lin_fp = affine_grid's base axis at the
 *   frame resolution (FP floats; Affine_Transformer's nearest mode, affine_transformer.py:41-52).  packed != 0 writes the
 *   softmax as the packed operand (as apgs_bmnist_template_match_packed_f32).  FP = 96, DP = 28 when a digit is given.        */
APGS_API int apgs_bmnist_template_match_residual_f32(const float* frame, int64_t frame_stride_s, int64_t frame_stride_b,
                                   const float* tmpl, int64_t tmpl_stride_s, int64_t tmpl_stride_b,
                                   const float* digit0, int64_t digit0_stride_s, int64_t digit0_stride_b, const float* z_where0,
                                   const float* digit1, int64_t digit1_stride_s, int64_t digit1_stride_b, const float* z_where1,
                                   const float* lin_fp, int64_t S, int64_t B, int64_t FP, int64_t DP, int softmax, int packed,
                                   float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* APGS_B200_H */
