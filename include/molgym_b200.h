/* molgym_b200.h — C ABI of the B200-native PPO policy/value hot path of gncs/molgym.
 *
 * Plain pointers and sizes only (no torch types).  All `const float*` / `float*` / `int*` arguments named d_* are
 * DEVICE pointers on the current CUDA device; `stream` is a cudaStream_t passed as void*.  Every entry point
 * returns 0 on success, a negative mgb_status otherwise; mgb_last_error() returns a message for the calling thread.
 *
 * Which reference interface each entry replaces (paths relative to the reference tree):
 *   mgb_cov_forward / mgb_cov_backward   — CovariantAC.step(observations, actions) in evaluate mode and its autograd
 *                                          backward: molgym/agents/covariant/agent.py:209-334 (cormorant stack
 *                                          molgym/agents/covariant/modules.py:97-135,180-190; invariants
 *                                          so3_tools.py:147-190; spherical distribution spherical_dists.py:182-286)
 *   mgb_ppo_loss                         — ppo.compute_loss arithmetic, molgym/ppo.py:28-52
 *   mgb_pack_observations                — CovariantAC.parse_observations + tools.process_atoms_list + spaces parse:
 *                                          agent.py:165-197, covariant/tools.py:8-49, spaces.py:55-61,106-107
 *   mgb_cov_param_count / layout         — the nn.Module parameter inventory of CovariantAC (agent.py:59-143)
 */
#ifndef MOLGYM_B200_H
#define MOLGYM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGB_MAX_SPECIES 16
#define MGB_MAX_LEVELS 4
#define MGB_MAXL 4

typedef enum {
  MGB_OK = 0,
  MGB_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
  MGB_ERR_CUDA = -2,      /* a CUDA runtime call or kernel launch failed */
  MGB_ERR_WORKSPACE = -3, /* workspace too small */
  MGB_ERR_STATE = -4      /* reserved */
} mgb_status;

/* Hyper-parameters of the covariant agent — same meaning as the CovariantAC constructor (agent.py:21-35). */
typedef struct {
  int32_t canvas_size;                 /* N: observation_space.canvas_space.size */
  int32_t num_species;                 /* Z: len(zs) */
  int32_t zs[MGB_MAX_SPECIES];         /* atomic numbers, zs[k] == 0 is the null symbol */
  int32_t maxl;                        /* must be 4 in this build */
  int32_t num_cg_levels;               /* 1..MGB_MAX_LEVELS */
  int32_t num_channels_hidden;         /* C */
  int32_t num_channels_per_element;    /* CPE; output channels = Z * CPE */
  int32_t num_gaussians;               /* G */
  int32_t network_width;               /* MLP width */
  float min_distance, max_distance;    /* min_max_distance */
  float bag_scale;
  int32_t has_beta;                    /* 0: SO3Distribution (beta=None); 1: ExpSO3Distribution */
  float beta;
  int32_t rel_sh_normalize;            /* `normalize` of the relative spherical harmonics (oracle switch #1; 0) */
} mgb_cov_config;

typedef struct mgb_cov_plan mgb_cov_plan;

const char* mgb_last_error(void);
int mgb_version(void);
/* 1 when the library was built by nvcc for sm_100a, 0 for the CPU kernel emulator used by the test-suite. */
int mgb_is_cuda_build(void);
/* <j1 m1 j2 m2 | j m> exactly as the plan builder evaluates it for its Clebsch-Gordan tables (host arithmetic, Racah's formula;
 * replaces cormorant.cg_lib.CGDict, molgym/agents/covariant/agent.py:59).  Exported so that the tables can be pinned against an
 * independent implementation (tests/test_clebsch_gordan.py compares with sympy.physics.quantum.cg.CG for all l <= 4). */
double mgb_clebsch_gordan(int32_t j1, int32_t m1, int32_t j2, int32_t m2, int32_t j, int32_t m);

/* Plan: immutable per-(config, device) state — Clebsch-Gordan term tables, Lebedev-71 quadrature harmonics.
 * lebedev_xyz[G*3], lebedev_w[G] (weights summing to 1) are HOST arrays (quadpy lebedev_071 ==
 * scipy.integrate.lebedev_rule(71) / 4 pi, spherical_dists.py:208-215). */
int mgb_cov_plan_create(const mgb_cov_config* cfg, const double* lebedev_xyz, const double* lebedev_w, int32_t n_grid,
                        mgb_cov_plan** out);
void mgb_cov_plan_destroy(mgb_cov_plan* plan);

/* Flat fp32 parameter buffer.  Tensors appear in this order, each in the reference's own shape/row-major layout:
 *   input.weight[2C,4Z] input.bias[2C]
 *   per level k: rad.scales[8] rad.phases[8] rad.linear[l].weight[2C,32] (l=0..L) rad.linear[l].bias[2C] (l=0..L)
 *                edge.weights[l][C,catE_kl,2] (l=0..L) atom.weights[l][Cout_k,catA_kl,2] (l=0..L)
 *   mixer.weights[l][CPE,catM_l,2] (l=0..L)
 *   phi_focus, phi_element, phi_d, phi_trans, phi_v: each weight0,bias0,weight1,bias1 ; distance_log_stds[G]
 * mgb_cov_param_layout fills offsets[i], numels[i] (floats) for i < mgb_cov_param_count(). */
int mgb_cov_param_count(const mgb_cov_plan* plan);
int mgb_cov_param_layout(const mgb_cov_plan* plan, int64_t* offsets, int64_t* numels, int64_t* total);
/* cat sizes: out[(k*(L+1)+l)*2+0] = catE_kl, +1 = catA_kl ; then L+1 entries catM_l */
int mgb_cov_cat_sizes(const mgb_cov_plan* plan, int32_t* out);

size_t mgb_cov_workspace_bytes(const mgb_cov_plan* plan, int32_t batch);

/* Per-canvas outputs of the forward (all device, fp32 unless noted).  Any pointer except logp/ent/v may be NULL. */
typedef struct {
  float* logp;            /* [B]  sum of the four sub-action log-probabilities (agent.py:295-301) */
  float* ent;             /* [B]  focus + element entropies (agent.py:304-308) */
  float* v;               /* [B]  state value (agent.py:313-316) */
  float* logp_parts;      /* [B,4] focus, element, distance, orientation */
  float* focus_probs;     /* [B,N] */
  float* element_probs;   /* [B,Z] */
  float* gmm;             /* [B,3,G] normalised log mixture weights, means, stds */
  float* coefficients;    /* [B,25,CPE,2] normalised a_lm (lm-major, channel, re/im)  (so3_dist.coefficients) */
  float* log_z;           /* [B] (has_beta only) */
  float* covariats;       /* [B,N,25,Z*CPE,2] last-level atom representations (lm-major, channel-minor) */
} mgb_cov_outputs;

/* Forward in evaluate mode.  d_positions[B,N,3] f32, d_charges[B,N] i32 (0 = padding, real atoms first),
 * d_bags[B,Z] f32, d_actions[B,6] f32 = focus, element, distance, ox, oy, oz (agent.py:230,249,271,288).
 * Saves what backward needs inside the workspace: the backward (or mgb_cov_policy) call must be handed the same
 * workspace, untouched, together with the same inputs and parameters. */
int mgb_cov_forward(mgb_cov_plan* plan, int32_t batch, const float* d_positions, const int32_t* d_charges,
                    const float* d_bags, const float* d_actions, const float* d_params, void* d_workspace,
                    size_t workspace_bytes, const mgb_cov_outputs* out, void* stream);

/* Policy heads only, on the workspace a previous mgb_cov_forward of the same batch filled (the Cormorant body is not
 * recomputed).  Used by rollout mode (agent.py:229-292), where the element head needs the sampled focus, the distance
 * head the sampled element and the orientation head the sampled distance: the caller re-evaluates the heads after
 * each sub-action is drawn.  Same outputs as mgb_cov_forward except `covariats`. */
/* Rollout mode of CovariantAC.step (actions=None, agent.py:229-292): the body, then one kernel that draws focus, element, distance
 * and orientation on the device — mode 1: samples (self.training; Categorical / mixture / rejection sampling on the sphere,
 * spherical_dists.py:116-150,227-262), mode 2: greedy (argmax, gmm.py:20-27, spherical_dists.py:152-158,264-271) — with Philox random
 * numbers keyed by `seed`, writes the chosen actions [batch, 6] and evaluates them (logp / ent / v and the optional outputs are
 * exactly what mgb_cov_forward returns for those actions).  No host round trip between the sub-actions. */
int mgb_cov_rollout(mgb_cov_plan* plan, int32_t batch, const float* d_positions, const int32_t* d_charges, const float* d_bags,
                    const float* d_params, void* d_workspace, size_t workspace_bytes, int32_t mode, uint64_t seed,
                    float* d_actions, const mgb_cov_outputs* outputs, void* stream);
int mgb_cov_policy(mgb_cov_plan* plan, int32_t batch, const float* d_bags, const float* d_actions, const float* d_params,
                   void* d_workspace, size_t workspace_bytes, const mgb_cov_outputs* out, void* stream);

/* Backward of the same call: cotangents d_g_logp/d_g_ent/d_g_v [B] -> d_grad_params (flat, same layout as params).
 * accumulate != 0 adds into d_grad_params, else it is overwritten. */
int mgb_cov_backward(mgb_cov_plan* plan, int32_t batch, const float* d_positions, const int32_t* d_charges,
                     const float* d_bags, const float* d_actions, const float* d_params, void* d_workspace,
                     size_t workspace_bytes, const float* d_g_logp, const float* d_g_ent, const float* d_g_v,
                     float* d_grad_params, int32_t accumulate, void* stream);

/* PPO-clip loss (ppo.py:28-52) and its cotangents in one launch.  logp/ent/v/old_logp fp32, adv/ret fp64 (the
 * reference hands float64 advantages/returns to torch, buffer.py:106-114, so the loss is float64).
 * inv_global_batch = 1 / (minibatch size across all ranks) so that sharded ranks produce gradient *sums*.
 * d_info[8] (fp64): loss, policy_loss, entropy_loss, vf_loss, approx_kl, clip_fraction, 0, 0 — local partial sums
 * already scaled by inv_global_batch.  Gradient outputs may be NULL (loss only). */
int mgb_ppo_loss(int32_t batch, const float* d_logp, const float* d_ent, const float* d_v, const float* d_old_logp,
                 const double* d_adv, const double* d_ret, double clip_ratio, double vf_coef, double entropy_coef,
                 double inv_global_batch, double* d_info, float* d_g_logp, float* d_g_ent, float* d_g_v, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Internal-coordinate (SchNet) actor-critic: SchNetAC.step(observations, actions) in evaluate mode and its backward,
 * molgym/agents/internal/agent.py:181-353 (make_atomic_tensors :112-151, surrogate_features :153-179) on
 * schnetpack 0.3's representation.SchNet(n_atom_basis = network_width / 2).
 *
 * Every canvas contributes three "molecules" of up to M = canvas_size + 1 atoms: [0] the canvas itself, [1] and [2] the
 * canvas plus the hypothetical new atom placed by the z-matrix for dihedral and -dihedral (the placement itself,
 * zmat.py:99-133, is float64 host arithmetic and stays with the caller).  d_numbers[B,3,M] i32 atomic numbers (0 = empty
 * slot, real atoms first), d_positions[B,3,M,3] f32, d_bags[B,Z] f32, d_actions[B,7] f32 = stop, focus, element,
 * distance, angle, dihedral, kappa.
 * Flat parameters, in this order (reference shapes): embedding[100,F]; per interaction t=0..2: filter.0.weight[128,25],
 * filter.0.bias, filter.1.weight[128,128], filter.1.bias, in2f.weight[128,F], f2out.weight[F,128], f2out.bias,
 * dense.weight[F,F], dense.bias; phi_beta, phi_focus, phi_element, phi_continuous, phi_kappa (weight0,bias0,weight1,
 * bias1 each); critic (three layers); log_stds[3]. */
typedef struct {
  int32_t canvas_size;
  int32_t num_species;
  int32_t zs[MGB_MAX_SPECIES];
  int32_t network_width;
  float min_distance, max_distance;
} mgb_int_config;
typedef struct mgb_int_plan mgb_int_plan;
typedef struct {
  float* logp;           /* [B] */
  float* ent;            /* [B] */
  float* v;              /* [B] */
  float* logp_terms;     /* [B,6] masked sub-action log-probabilities (may be NULL, like everything below) */
  float* focus_probs;    /* [B,N] */
  float* element_probs;  /* [B,Z] */
  float* means;          /* [B,3] distance, angle, dihedral means */
  float* kappa_logits;   /* [B,2] */
} mgb_int_outputs;
int mgb_int_plan_create(const mgb_int_config* cfg, mgb_int_plan** out);
void mgb_int_plan_destroy(mgb_int_plan* plan);
int mgb_int_param_count(const mgb_int_plan* plan);
int mgb_int_param_layout(const mgb_int_plan* plan, int64_t* offsets, int64_t* numels, int64_t* total);
size_t mgb_int_workspace_bytes(const mgb_int_plan* plan, int32_t batch);
int mgb_int_forward(mgb_int_plan* plan, int32_t batch, const int32_t* d_numbers, const float* d_positions, const float* d_bags,
                    const float* d_actions, const float* d_params, void* d_workspace, size_t workspace_bytes,
                    const mgb_int_outputs* out, void* stream);
int mgb_int_backward(mgb_int_plan* plan, int32_t batch, const int32_t* d_numbers, const float* d_positions, const float* d_bags,
                     const float* d_actions, const float* d_params, void* d_workspace, size_t workspace_bytes,
                     const float* d_g_logp, const float* d_g_ent, const float* d_g_v, float* d_grad_params, int32_t accumulate,
                     void* stream);

/* Launch accounting and per-kernel timing, for bench.py (no counterpart in the reference).
 * mgb_launch_count: kernels launched by this library in this process so far.
 * mgb_profile_kernel(substr): from now on bracket every launch whose kernel name contains `substr` with CUDA events on
 *   the launching stream (NULL or "" switches it off).  mgb_profile_read: synchronise those events, return the summed
 *   duration in milliseconds and the number of launches timed, and reset. */
/* dst = (accumulate ? dst : 0) + scale * src over n floats; `scale` points to ONE device scalar (float64 when scale_is_double,
 * else float32).  Used by the fused PPO step to fold the ready scratch gradient into the parameters' .grad with the cotangent
 * autograd hands over (replaces torch's cast + multiply + add; reference: the gradient accumulation of loss.backward(),
 * molgym/ppo.py:131).  dst / src 16-byte aligned. */
int mgb_scale_accumulate(float* dst, const float* src, const void* scale, int32_t scale_is_double, int64_t n, int32_t accumulate,
                         void* stream);
/* Optimizer tail of one PPO epoch on the flat buffers (replaces tools/util.py:61-69 compute_gradient_norm,
 * torch.nn.utils.clip_grad_norm_ at ppo.py:144 and the torch.optim.Adam step of tools/util.py:197-205 / ppo.py:145):
 *   mgb_grad_norm : norm[0] = ||grad||_2, norm[1] = ||grad||_2^2 (float64, deterministic), one launch; `scratch` = device buffer of
 *                   mgb_optim_scratch_bytes() bytes, zeroed once by the caller before the first use;
 *   mgb_adam_step : torch.optim.Adam(amsgrad, weight_decay, maximize) on float32 state, gradients scaled by
 *                   min(1, max_norm / (norm[0] + 1e-6)) when max_norm > 0 (clip_grad_norm_ semantics, norm read on the device);
 *                   `step` is the 1-based update count of this call.  All pointers are device pointers. */
size_t mgb_optim_scratch_bytes(void);
int mgb_grad_norm(const float* d_grad, int64_t n, void* d_scratch, double* d_norm, void* stream);
int mgb_adam_step(float* d_params, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, float* d_max_exp_avg_sq, int64_t n,
                  double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step, int32_t amsgrad,
                  int32_t maximize, const double* d_norm, double max_norm, void* stream);
int64_t mgb_launch_count(void);
int mgb_profile_kernel(const char* substr);
int mgb_profile_read(double* total_ms, int64_t* launches);
/* Per-launch report of the timed launches since the last read, one "<kernel> <milliseconds>\n" line each in launch order
 * (truncated at cap); clears the list like mgb_profile_read. */
int mgb_profile_report(char* buf, int64_t cap);

/* Host-side observation packer (no device work).  labels[B,N] are indices into zs, xyz[B,N,3] float64 canvas
 * coordinates, as found in ObservationType tuples.  Null-symbol items are dropped and the rest compacted to the
 * front (spaces.py:55-61), padding is zero (covariant/tools.py:18-31). Outputs are HOST arrays
 * positions[B,N,3] f32, charges[B,N] i32.  Returns MGB_ERR_INVALID for a negative or out-of-range label. */
int mgb_pack_observations(const mgb_cov_config* cfg, int32_t batch, const int32_t* labels, const double* xyz,
                          float* positions, int32_t* charges);

#ifdef __cplusplus
}
#endif
#endif /* MOLGYM_B200_H */
