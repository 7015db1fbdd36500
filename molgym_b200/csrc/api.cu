// api.cu — the C ABI (include/molgym_b200.h): plan construction, workspace carving, kernel launch sequences.
// Compiled by nvcc for sm_100a (product) or by g++ -DMGB_CUSIM against tests/cusim/cusim.h (kernel-logic tests).
#include "plan.cuh"
#include "mlp_tc.cuh"
#include "mix_tc.cuh"
#include "cov_backward.cuh"
#include "internal.cuh"
#include "optim.cuh"

using namespace mgb;

constexpr size_t kMaxDynSmemMix = 227 * 1024;

#define MGB_LAUNCH_OK(what)                                                                             \
  do {                                                                                                  \
    cudaError_t e_ = cudaGetLastError();                                                                \
    if (e_ != cudaSuccess) return fail(MGB_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e_)); \
  } while (0)

template <int CO, bool BACKWARD, int KS, int THREADS>
static int launch_mix_rows_kt(const mgb_cov_plan* plan, int level, int B, const CovWs& w, const float* A_out, float* out, cudaStream_t st) {
  const CovDesc& d = plan->desc;
  const LevelDesc& L = d.lv[level];
  int kmax = 0;
  for (int l = 0; l < kNL; ++l) kmax = std::max(kmax, L.catA[l]);
  const size_t smem = sizeof(float2) * (size_t)kmax * kMixStride<CO>;
  MGB_CUDA_OK(cudaFuncSetAttribute((k_mix_rows<CO, BACKWARD, KS, THREADS>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long rows_per_cta = THREADS / KS;
  const long long groups = ((long long)B * d.N * 9 + rows_per_cta - 1) / rows_per_cta;
  dim3 grid((unsigned)std::min<long long>(groups, 148 * (KS == 32 ? 3 : 2)), kNL);   // ~10-15 CTAs per SM over the five ells; each loops over row groups
  MGB_LAUNCH((k_mix_rows<CO, BACKWARD, KS, THREADS>), grid, THREADS, smem, st, plan->d_desc, level, w.Wt, w.atom_off, w.atom_list, B,
             w.cat[level], A_out, out);
  MGB_LAUNCH_OK("k_mix_rows");
  return MGB_OK;
}
template <int CO, bool BACKWARD, int KS>
static int launch_mix_rows_ks(const mgb_cov_plan* plan, int level, int B, const CovWs& w, const float* A_out, float* out, cudaStream_t st) {
  // the dcat mix of large minibatches runs 512-thread CTAs (fewer stagings of W_l per row); small minibatches are latency-bound and
  // faster with 128-thread CTAs (C2: 21 against 28 us per launch)
  if (BACKWARD && large_atoms(B, plan->desc.N)) return launch_mix_rows_kt<CO, BACKWARD, KS, kMixThreadsLarge>(plan, level, B, w, A_out, out, st);
  return launch_mix_rows_kt<CO, BACKWARD, KS, kMixThreads>(plan, level, B, w, A_out, out, st);
}
template <int CO, bool BACKWARD>
static int launch_mix_rows_co(const mgb_cov_plan* plan, int level, int B, const CovWs& w, const float* A_out, float* out, cudaStream_t st) {
  // 8 adjacent lanes share a row: its cat entries are read / written as contiguous 64-byte pieces (a whole warp per row was
  // measured slower at C2: every CTA stages the ell's weights for fewer rows)
  return launch_mix_rows_ks<CO, BACKWARD, 8>(plan, level, B, w, A_out, out, st);
}
// tensor-core channel mix (mix_tc.cuh): even channel counts whose real-expanded output fits 3 or 4 n-tiles (Cout 9..16), wide
// enough inner dimensions.  Measured on B200 (profiles/r2n_mix_tc.md): the forward wins from a few thousand atoms per minibatch
// on (C5 b256: 130 / 144 us against 182 / 243 us), the backward does not (its FFMA version is already bound by the dcat
// writes), small minibatches are launch-latency bound either way.  Default: forward only, large minibatches.
// MGB_MIX_TC=0: never; MGB_MIX_TC=1: always, both directions (parity tests).
static int mix_tc_ntiles(const LevelDesc& L, bool large, bool backward) {
  const char* e = std::getenv("MGB_MIX_TC");
  if (e && e[0] == '0') return 0;
  const bool forced = e && e[0] == '1';
  if (!forced && (backward || !large)) return 0;
  const int nt = (2 * L.Cout + 7) / 8;
  if (nt != 3 && nt != 4) return 0;
  int kmax = 0;
  if (L.totA % 2) return 0;
  for (int l = 0; l < kNL; ++l) {
    if (L.catA[l] % 2 || L.offA[l] % 2 || L.catA[l] <= 0) return 0;   // 16-byte aligned cat rows
    kmax = std::max(kmax, L.catA[l]);
  }
  if (kmax < 32 || mix_tc_ksteps(kmax) > kMixTcWarps * kMixTcKsw) return 0;   // the forward keeps a warp's weight fragments in registers
  if (mix_tc_fwd_smem_bytes(kmax, nt) > kMaxDynSmemMix || mix_tc_bwd_smem_bytes(kmax, nt) > kMaxDynSmemMix) return 0;
  return nt;
}
template <int NT, bool BACKWARD>
static int launch_mix_rows_tc(const mgb_cov_plan* plan, int level, int B, const float* P, const CovWs& w, const float* A_out, float* out,
                              cudaStream_t st) {
  const CovDesc& d = plan->desc;
  const LevelDesc& L = d.lv[level];
  int kmax = 0;
  for (int l = 0; l < kNL; ++l) kmax = std::max(kmax, L.catA[l]);
  const long long tiles = ((long long)B * d.N * kM + kMixTcRows - 1) / kMixTcRows + kNL;
  static const int c0 = [] { const char* e = std::getenv("MGB_MIX_C0"); return e ? std::atoi(e) : 256; }();   // per-tile constant of the CTA deal
  if (!BACKWARD) {
    const size_t smem = mix_tc_fwd_smem_bytes(kmax, NT);
    MGB_CUDA_OK(cudaFuncSetAttribute(k_mix_rows_tc_fwd<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)std::max<long long>(kNL, std::min<long long>(tiles, 148));
    MGB_LAUNCH(k_mix_rows_tc_fwd<NT>, grid, kMixTcThreads, smem, st, plan->d_desc, level, P, w.atom_off, w.atom_list, B, w.cat[level], out, c0);
  } else {
    const size_t smem = mix_tc_bwd_smem_bytes(kmax, NT);
    MGB_CUDA_OK(cudaFuncSetAttribute(k_mix_rows_tc_bwd<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)std::max<long long>(kNL, std::min<long long>(tiles, 148));
    MGB_LAUNCH(k_mix_rows_tc_bwd<NT>, grid, kMixTcThreads, smem, st, plan->d_desc, level, P, w.atom_off, w.atom_list, B, A_out, out, c0);
  }
  MGB_LAUNCH_OK("k_mix_rows_tc");
  return MGB_OK;
}
template <bool BACKWARD>
static int launch_mix_rows(const mgb_cov_plan* plan, int level, int B, const float* P, const CovWs& w, const float* A_out, float* out,
                           cudaStream_t st) {
  switch (mix_tc_ntiles(plan->desc.lv[level], large_atoms(B, plan->desc.N), BACKWARD)) {
    case 3: return launch_mix_rows_tc<3, BACKWARD>(plan, level, B, P, w, A_out, out, st);
    case 4: return launch_mix_rows_tc<4, BACKWARD>(plan, level, B, P, w, A_out, out, st);
    default: break;
  }
  switch (pick_co_rows(plan->desc.lv[level].Cout)) {
    case 4: return launch_mix_rows_co<4, BACKWARD>(plan, level, B, w, A_out, out, st);
    case 8: return launch_mix_rows_co<8, BACKWARD>(plan, level, B, w, A_out, out, st);
    case 10: return launch_mix_rows_co<10, BACKWARD>(plan, level, B, w, A_out, out, st);
    case 12: return launch_mix_rows_co<12, BACKWARD>(plan, level, B, w, A_out, out, st);
    case 16: return launch_mix_rows_co<16, BACKWARD>(plan, level, B, w, A_out, out, st);
    case 20: return launch_mix_rows_co<20, BACKWARD>(plan, level, B, w, A_out, out, st);
    default: return launch_mix_rows_co<32, BACKWARD>(plan, level, B, w, A_out, out, st);
  }
}

template <int NLM2, int CT>
static int launch_atom_cat_ct(const mgb_cov_plan* plan, int level, int B, const float* pos, const CovWs& w, int phases, cudaStream_t st) {
  const CovDesc& d = plan->desc;
  const size_t smem = sizeof(float) * atom_smem_floats(d.lv[level], d.N);
  MGB_CUDA_OK(cudaFuncSetAttribute((k_atom_cat<NLM2, CT>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MGB_LAUNCH((k_atom_cat<NLM2, CT>), B * d.N, kAtomThreads, smem, st, plan->d_desc, level, pos, w.n_atoms, w.atom_off, w.atom_list, B,
             w.A[level], w.E[level], w.cat[level], phases);
  MGB_LAUNCH_OK("k_atom_cat");
  return MGB_OK;
}
template <int NLM2>
static int launch_atom_cat(const mgb_cov_plan* plan, int level, int B, const float* pos, const CovWs& w, int phases, cudaStream_t st) {
  // the default hidden width gets compile-time channel strides
  return plan->desc.lv[level].C == 10 ? launch_atom_cat_ct<NLM2, 10>(plan, level, B, pos, w, phases, st)
                                      : launch_atom_cat_ct<NLM2, 0>(plan, level, B, pos, w, phases, st);
}

template <int NLM2>
static int launch_atom_fwd(const mgb_cov_plan* plan, int level, int B, const float* P, const float* pos, const CovWs& w,
                           cudaStream_t st) {
  const CovDesc& d = plan->desc;
  int rc = launch_atom_cat<NLM2>(plan, level, B, pos, w, small_atoms(B, d.N) ? kAtomPhaseA : kAtomPhaseA | kAtomPhaseB, st);
  if (rc != MGB_OK) return rc;
  if (small_atoms(B, d.N)) MGB_CUDA_OK(cudaStreamWaitEvent(st, plan->ev_join3[level], 0));   // the square / pass-through blocks (side3)
  return launch_mix_rows<false>(plan, level, B, P, w, nullptr, w.A[level + 1], st);
}

// Row MLPs (focus head, value transform): shared-memory resident weights when they fit, else the generic kernels.
constexpr size_t kMaxDynSmem = 227 * 1024;
static bool rows_mlp_smem_ok(const CovDesc& d) {
  return d.lat % 4 == 0 && d.Wd % 4 == 0 && d.focus.in == d.trans.in && d.focus.hidden == d.trans.hidden &&
         rows_mlp_fwd_smem_bytes<32>(d.lat, d.Wd, d.Wd) <= kMaxDynSmem && rows_mlp_bwd_smem_bytes<32>(d.lat, d.Wd) <= kMaxDynSmem;
}
template <int RT>
static int launch_rows_mlp_fwd_rt(const mgb_cov_plan* plan, int B, const float* P, const CovWs& w, cudaStream_t st) {
  const CovDesc& d = plan->desc;
  const long long rows = (long long)B * d.N;
  const size_t sm = rows_mlp_fwd_smem_bytes<RT>(d.lat, d.Wd, d.Wd);
  MGB_CUDA_OK(cudaFuncSetAttribute(k_rows_mlp_fwd_smem<RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  dim3 grid((unsigned)((rows + RT - 1) / RT), 2);
  MGB_LAUNCH(k_rows_mlp_fwd_smem<RT>, grid, kHeadThreads, sm, st, plan->d_desc, P, w.Wt, B, w.act_off, w.act_list, w.atom_off, w.atom_list,
             w.inv, w.hf, w.flogit, w.ht0, w.trans);
  MGB_LAUNCH_OK("k_rows_mlp_fwd_smem");
  return MGB_OK;
}
// tensor-core row MLPs (mlp_tc.cuh): the default widths; anything else takes the FFMA kernels above
static bool rows_mlp_tc_ok(const CovDesc& d) {
  const char* e = std::getenv("MGB_MLP_TC");
  if (e && e[0] == '0') return false;
  return d.lat % 8 == 0 && d.Wd % 64 == 0 && d.Wd <= 256 && d.lat <= 256 && d.focus.in == d.trans.in && d.focus.hidden == d.trans.hidden &&
         d.trans.out == d.Wd && d.focus.out == 1 && (d.focus.W0 % 4) == 0 && (d.trans.W0 % 4) == 0 && (d.trans.W1 % 4) == 0 &&
         rows_mlp_tc_fwd_smem_bytes(d.lat, d.Wd, true) <= kMaxDynSmem && rows_mlp_tc_bwd_smem_bytes(d.lat, d.Wd, true) <= kMaxDynSmem;
}
template <int NT>
static int launch_rows_mlp_fwd_tc(const mgb_cov_plan* plan, int B, const float* P, const CovWs& w, cudaStream_t st) {
  const CovDesc& d = plan->desc;
  const size_t sm = rows_mlp_tc_fwd_smem_bytes(d.lat, d.Wd, true);
  MGB_CUDA_OK(cudaFuncSetAttribute(k_rows_mlp_fwd_tc<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  const long long tiles = ((long long)B * d.N + kTcRows - 1) / kTcRows;
  dim3 grid((unsigned)std::min<long long>(tiles, 148), 2);   // one CTA per SM and role keeps the weights for its row tiles
  MGB_LAUNCH(k_rows_mlp_fwd_tc<NT>, grid, kTcThreads, sm, st, plan->d_desc, P, B, w.act_off, w.act_list, w.atom_off, w.atom_list, w.inv, w.hf,
             w.flogit, w.ht0, w.trans);
  MGB_LAUNCH_OK("k_rows_mlp_fwd_tc");
  return MGB_OK;
}
static int launch_rows_mlp_fwd(const mgb_cov_plan* plan, int B, const float* P, const CovWs& w, cudaStream_t st) {
  const CovDesc& d = plan->desc;
  const long long rows = (long long)B * d.N;
  if (rows_mlp_tc_ok(d)) {
    switch (d.Wd / 64) {
      case 1: return launch_rows_mlp_fwd_tc<1>(plan, B, P, w, st);
      case 2: return launch_rows_mlp_fwd_tc<2>(plan, B, P, w, st);
      case 3: return launch_rows_mlp_fwd_tc<3>(plan, B, P, w, st);
      case 4: return launch_rows_mlp_fwd_tc<4>(plan, B, P, w, st);
      default: break;
    }
  }
  if (rows_mlp_smem_ok(d)) return rows <= 8ll * 148 * 4 ? launch_rows_mlp_fwd_rt<8>(plan, B, P, w, st) : launch_rows_mlp_fwd_rt<32>(plan, B, P, w, st);
  dim3 grid((unsigned)((rows + kRowTile - 1) / kRowTile), 2);
  const size_t sm = sizeof(float) * kRowTile * (d.lat + d.Wd);
  MGB_CUDA_OK(cudaFuncSetAttribute(k_rows_mlp_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  MGB_LAUNCH(k_rows_mlp_fwd, grid, kHeadThreads, sm, st, plan->d_desc, P, w.Wt, w.n_atoms, rows, w.inv, w.hf, w.flogit, w.ht0, w.trans);
  MGB_LAUNCH_OK("k_rows_mlp_fwd");
  return MGB_OK;
}
template <int RT>
static int launch_rows_mlp_bwd_rt(const mgb_cov_plan* plan, int B, const float* P, const CovWs& w, cudaStream_t st) {
  const CovDesc& d = plan->desc;
  const long long rows = (long long)B * d.N;
  const size_t sm = rows_mlp_bwd_smem_bytes<RT>(d.lat, d.Wd);
  MGB_CUDA_OK(cudaFuncSetAttribute(k_rows_mlp_bwd_smem<RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  dim3 grid((unsigned)((rows + RT - 1) / RT), 2);
  MGB_LAUNCH(k_rows_mlp_bwd_smem<RT>, grid, kHeadThreads, sm, st, plan->d_desc, P, B, w.act_off, w.act_list, w.atom_off, w.atom_list, w.hf,
             w.dflogit, w.dhf, w.ht0, w.dvf, w.dtrans, w.dht0, w.dinv);
  MGB_LAUNCH_OK("k_rows_mlp_bwd_smem");
  return MGB_OK;
}
template <int NT, int NTX>
static int launch_rows_mlp_bwd_tc(const mgb_cov_plan* plan, int B, const float* P, const CovWs& w, cudaStream_t st) {
  const CovDesc& d = plan->desc;
  const size_t sm = rows_mlp_tc_bwd_smem_bytes(d.lat, d.Wd, true);
  MGB_CUDA_OK(cudaFuncSetAttribute((k_rows_mlp_bwd_tc<NT, NTX>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  const long long tiles = ((long long)B * d.N + kTcRows - 1) / kTcRows;
  dim3 grid((unsigned)std::min<long long>(tiles, 148), 2);
  MGB_LAUNCH((k_rows_mlp_bwd_tc<NT, NTX>), grid, kTcThreads, sm, st, plan->d_desc, P, B, w.act_off, w.act_list, w.atom_off, w.atom_list, w.hf,
             w.dflogit, w.dhf, w.ht0, w.dvf, w.dtrans, w.dht0, w.dinv);
  MGB_LAUNCH_OK("k_rows_mlp_bwd_tc");
  return MGB_OK;
}
static int launch_rows_mlp_bwd(const mgb_cov_plan* plan, int B, const float* P, const CovWs& w, cudaStream_t st) {
  const CovDesc& d = plan->desc;
  const long long rows = (long long)B * d.N;
  if (rows_mlp_tc_ok(d)) {
    // input n-tiles per warp: ceil(lat / 64) <= 4 for lat <= 256
    if (d.Wd / 64 == 2) return launch_rows_mlp_bwd_tc<2, 4>(plan, B, P, w, st);
    if (d.Wd / 64 == 1) return launch_rows_mlp_bwd_tc<1, 4>(plan, B, P, w, st);
    if (d.Wd / 64 == 3) return launch_rows_mlp_bwd_tc<3, 4>(plan, B, P, w, st);
    if (d.Wd / 64 == 4) return launch_rows_mlp_bwd_tc<4, 4>(plan, B, P, w, st);
  }
  if (rows_mlp_smem_ok(d)) return rows <= 8ll * 148 * 4 ? launch_rows_mlp_bwd_rt<8>(plan, B, P, w, st) : launch_rows_mlp_bwd_rt<32>(plan, B, P, w, st);
  const size_t sm = sizeof(float) * kRowTile * d.Wd * 3;
  MGB_CUDA_OK(cudaFuncSetAttribute(k_rows_mlp_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  MGB_LAUNCH(k_rows_mlp_bwd, (unsigned)((rows + kRowTile - 1) / kRowTile), kHeadThreads, sm, st, plan->d_desc, P, w.n_atoms, rows, w.hf,
             w.dflogit, w.dhf, w.ht0, w.dvf, w.dtrans, w.dht0, w.dinv);
  MGB_LAUNCH_OK("k_rows_mlp_bwd");
  return MGB_OK;
}

static int launch_policy_fwd(const mgb_cov_plan* plan, int B, const float* bags, const float* actions, const float* P,
                             const CovWs& w, const mgb_cov_outputs* out, cudaStream_t st) {
  const CovDesc& d = plan->desc;
  const size_t sm = sizeof(float) * policy_smem_floats(d);
  const int grid = std::min(B, 148 * 4);
  if (B <= 148) {   // at most one canvas per SM: the spill-free instantiation
    MGB_CUDA_OK(cudaFuncSetAttribute(k_policy_fwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    MGB_LAUNCH(k_policy_fwd<1>, grid, kPolicyThreads, sm, st, plan->d_desc, P, w.Wt, B, w.n_atoms, bags, actions, w.A[d.K], w.inv,
               w.flogit, w.trans, reinterpret_cast<float2*>(w.lse), w.pol_state, *out);
  } else {          // two resident CTAs per SM
    MGB_CUDA_OK(cudaFuncSetAttribute(k_policy_fwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    MGB_LAUNCH(k_policy_fwd<2>, grid, kPolicyThreads, sm, st, plan->d_desc, P, w.Wt, B, w.n_atoms, bags, actions, w.A[d.K], w.inv,
               w.flogit, w.trans, reinterpret_cast<float2*>(w.lse), w.pol_state, *out);
  }
  MGB_LAUNCH_OK("k_policy_fwd");
  return MGB_OK;
}

extern "C" {

const char* mgb_last_error(void) { return g_err; }
int mgb_version(void) { return 100; }
int mgb_is_cuda_build(void) {
#ifdef MGB_CUSIM
  return 0;
#else
  return 1;
#endif
}

double mgb_clebsch_gordan(int32_t j1, int32_t m1, int32_t j2, int32_t m2, int32_t j, int32_t m) { return clebsch_gordan(j1, m1, j2, m2, j, m); }

int mgb_cov_plan_create(const mgb_cov_config* cfg, const double* leb_xyz, const double* leb_w, int32_t n_grid,
                        mgb_cov_plan** out) {
  if (!cfg || !out) return fail(MGB_ERR_INVALID, "null argument");
  if (cfg->maxl != kL) return fail(MGB_ERR_INVALID, "this build supports maxl == %d only (got %d)", kL, cfg->maxl);
  if (cfg->num_cg_levels < 1 || cfg->num_cg_levels > kMaxLevels) return fail(MGB_ERR_INVALID, "num_cg_levels out of range");
  if (cfg->num_species < 1 || cfg->num_species > MGB_MAX_SPECIES) return fail(MGB_ERR_INVALID, "num_species out of range");
  if (cfg->canvas_size < 1 || cfg->canvas_size > 64) return fail(MGB_ERR_INVALID, "canvas_size must be in 1..64");
  if (cfg->num_channels_hidden < 1 || cfg->num_channels_hidden > 10)
    return fail(MGB_ERR_INVALID, "num_channels_hidden must be in 1..10 in this build");
  if (cfg->num_channels_per_element < 1 || cfg->num_channels_per_element > 4)
    return fail(MGB_ERR_INVALID, "num_channels_per_element must be in 1..4");
  if (cfg->num_species * cfg->num_channels_per_element > 32) return fail(MGB_ERR_INVALID, "too many output channels");
  if (cfg->num_gaussians < 1 || cfg->num_gaussians > 8) return fail(MGB_ERR_INVALID, "num_gaussians must be in 1..8");
  if (cfg->network_width < 1 || cfg->network_width > 1024) return fail(MGB_ERR_INVALID, "network_width out of range");
  if (cfg->rel_sh_normalize) return fail(MGB_ERR_INVALID, "rel_sh_normalize=1 is not supported");
  if (cfg->has_beta && (n_grid <= 0 || !leb_xyz || !leb_w)) return fail(MGB_ERR_INVALID, "beta needs the Lebedev grid");

  std::unique_ptr<mgb_cov_plan> plan(new mgb_cov_plan());
  plan->cfg = *cfg;
  CovDesc& d = plan->desc;
  std::memset(&d, 0, sizeof(d));
  d.N = cfg->canvas_size; d.Z = cfg->num_species; d.K = cfg->num_cg_levels;
  d.C = cfg->num_channels_hidden; d.CPE = cfg->num_channels_per_element; d.Cout = d.Z * d.CPE;
  d.G = cfg->num_gaussians; d.Wd = cfg->network_width;
  d.lat = (kL + 2) * d.Cout * 2; d.latE = (kL + 2) * d.CPE * 2;
  d.S_in = d.Z * 3 + d.Z;
  int zmax = 1;
  for (int z = 0; z < d.Z; ++z) { d.zs[z] = cfg->zs[z]; zmax = std::max(zmax, cfg->zs[z]); }
  d.charge_scale = (float)zmax;                                   // agent.py:71 charge_scale=max(zs)
  d.bag_scale = cfg->bag_scale;
  d.cut_rad = std::max(1e-3f, std::fabs(std::min(cfg->max_distance, 2.1f)));   // agent.py:66-69 + MaskLevel eps
  d.cut_width = std::max(1e-3f, 0.2f);
  d.dmin = cfg->min_distance; d.dmax = cfg->max_distance;
  d.has_beta = cfg->has_beta; d.beta = cfg->beta;
  d.n_grid = cfg->has_beta ? n_grid : 0;

  // ---- parameter layout (see include/molgym_b200.h) + transposed-weights scratch layout
  long long p = 0, wt = 0;
  auto param = [&](long long numel) { plan->p_offsets.push_back(p); plan->p_numels.push_back(numel); long long o = p; p += numel; return o; };
  const int C = d.C, C2 = 2 * C;
  d.p_inW = param((long long)C2 * d.S_in);
  d.p_inb = param(C2);
  TableArena arena;
  std::vector<PendingTable> pending;
  struct PendingGather { GatherTables* dst; size_t o_ag, o_sq, o_ag_slot, o_sq_slot; };
  std::vector<PendingGather> pending_g;
  HostCgTable sq_full = build_cg_table(kNL, kNL);
  for (int k = 0; k < d.K; ++k) {
    LevelDesc& L = d.lv[k];
    L.nLin = k == 0 ? 1 : kNL;
    L.nlm_in = L.nLin * L.nLin;
    L.C = C;
    L.Cout = (k == d.K - 1) ? d.Cout : C;
    L.has_prev = k > 0;
    HostCgTable ag = build_cg_table(kNL, L.nLin), sq = build_cg_table(L.nLin, L.nLin);
    int eo = 0, ao = 0, wo = 0;
    L.sumCatE = 0;
    for (int l = 0; l < kNL; ++l) {
      L.catE[l] = (L.has_prev ? C : 0) + (l < L.nLin ? L.nLin * C : 0) + C;
      L.offE[l] = eo; eo += C * L.catE[l];
      L.sumCatE += L.catE[l];
      const bool has_in = l < L.nLin;
      L.in_block[l] = has_in ? ag.n_blocks[l] : -1;
      L.sq_block[l] = ag.n_blocks[l] + (has_in ? 1 : 0);
      L.catA[l] = C * (ag.n_blocks[l] + (has_in ? 1 : 0) + sq.n_blocks[l]);
      L.offA[l] = ao; ao += L.catA[l] * (2 * l + 1);
      L.offWA[l] = wo; wo += L.Cout * L.catA[l];
    }
    L.totE = eo; L.totA = ao; L.totWA = wo;
    L.p_scales = param(kTrig);
    L.p_phases = param(kTrig);
    L.p_radW = p;
    for (int l = 0; l < kNL; ++l) param((long long)C2 * kRadFeat);
    L.p_radb = p;
    for (int l = 0; l < kNL; ++l) param(C2);
    L.p_edgeW = p;
    for (int l = 0; l < kNL; ++l) param((long long)C * L.catE[l] * 2);
    L.p_atomW = p;
    for (int l = 0; l < kNL; ++l) param((long long)L.Cout * L.catA[l] * 2);
    // transposed copies: edge weights [l][k][c'][2] then radial weights [l][t][2C]
    d.wt_edge[k] = wt;
    for (int l = 0; l < kNL; ++l)
      plan->segs.push_back(TransposeSeg{L.p_edgeW + 2ll * L.offE[l], wt + 2ll * L.offE[l], C, L.catE[l], 2, C});
    wt += 2ll * L.totE;
    for (int l = 0; l < kNL; ++l)
      plan->segs.push_back(TransposeSeg{L.p_radW + (long long)l * C2 * kRadFeat, wt + (long long)l * C2 * kRadFeat, C2, kRadFeat, 1, C2});
    wt += (long long)kNL * C2 * kRadFeat;
    wt = (wt + 3) & ~3ll;   // 16-byte aligned: the row-mix kernels bulk-copy these matrices into shared memory
    d.wt_atom[k] = wt;      // atom-mix weights transposed and zero-padded to [l][k][mixCS][2]
    L.mixCS = mix_stride_of(pick_co_rows(L.Cout));
    {
      int wo_t = 0;
      for (int l = 0; l < kNL; ++l) {
        L.offWAt[l] = wo_t;
        plan->segs.push_back(TransposeSeg{L.p_atomW + 2ll * L.offWA[l], wt + 2ll * wo_t, L.Cout, L.catA[l], 2, L.mixCS});
        wo_t += L.mixCS * L.catA[l];
      }
      wt += 2ll * wo_t;
    }
    {
      const int zero[kNL] = {0, 0, 0, 0, 0};
      resolve_cg_table(ag, L.catA, L.offA, zero, C, false);
      resolve_cg_table(sq, L.catA, L.offA, L.sq_block, C, true);
      if (!finalize_cg_table(ag, kAtomThreads / C, false) || !finalize_cg_table(sq, kAtomThreads / C, true))
        return fail(MGB_ERR_INVALID, "internal: CG pair table wider than kCgPad");
    }
    pending.push_back(stage_table(arena, ag, &L.ag));
    pending.push_back(stage_table(arena, sq, &L.sq));
    std::memset(&L.gt, 0, sizeof(L.gt));
    if (L.nLin == kNL) {
      HostGather h;
      if (!build_gather_tables(ag, sq, C, h)) return fail(MGB_ERR_INVALID, "internal: gather tables do not fit their encoding");
      pending_g.push_back(PendingGather{&L.gt, arena.add(h.ag_flat8.data(), h.ag_flat8.size() * sizeof(int)),
                                        arena.add(h.sq_flat8.data(), h.sq_flat8.size() * sizeof(int)),
                                        arena.add(h.ag_slot.data(), h.ag_slot.size() * sizeof(int)),
                                        arena.add(h.sq_slot.data(), h.sq_slot.size() * sizeof(int))});
    }
  }
  {  // mixer
    int ao = 0, wo = 0;
    for (int l = 0; l < kNL; ++l) {
      d.catM[l] = d.CPE * (1 + sq_full.n_blocks[l] + 1);
      d.inM_block[l] = 1 + sq_full.n_blocks[l];
      d.offM[l] = ao; ao += d.catM[l] * (2 * l + 1);
      d.offWM[l] = wo; wo += d.CPE * d.catM[l];
    }
    d.totM = ao; d.totWM = wo;
    d.p_mixW = p;
    for (int l = 0; l < kNL; ++l) param((long long)d.CPE * d.catM[l] * 2);
    {
      const int one[kNL] = {1, 1, 1, 1, 1};
      resolve_cg_table(sq_full, d.catM, d.offM, one, d.CPE, true);
      if (!finalize_cg_table(sq_full, kPolicyThreads / d.CPE, true)) return fail(MGB_ERR_INVALID, "internal: CG pair table wider than kCgPad");
    }
    pending.push_back(stage_table(arena, sq_full, &d.mix_sq));
  }
  auto mlp = [&](MlpDesc& m, int in, int hidden, int outn) {
    fill_mlp(m, in, hidden, outn, p, wt);
    plan->p_offsets.push_back(m.W0); plan->p_numels.push_back((long long)hidden * in);
    plan->p_offsets.push_back(m.b0); plan->p_numels.push_back(hidden);
    plan->p_offsets.push_back(m.W1); plan->p_numels.push_back((long long)outn * hidden);
    plan->p_offsets.push_back(m.b1); plan->p_numels.push_back(outn);
    plan->segs.push_back(TransposeSeg{m.W0, m.W0t, hidden, in, 1, hidden});
    plan->segs.push_back(TransposeSeg{m.W1, m.W1t, outn, hidden, 1, outn});
  };
  mlp(d.focus, d.lat, d.Wd, 1);
  mlp(d.element, d.lat, d.Wd, d.Z);
  mlp(d.dist, d.latE, d.Wd, 2 * d.G);
  mlp(d.trans, d.lat, d.Wd, d.Wd);
  mlp(d.value, d.Wd, d.Wd, 1);
  d.p_logstd = param(d.G);
  d.n_params = p;
  d.n_wt = wt;
  d.n_units_hidden = build_mix_units(d.units_hidden, 2);
  d.n_units_out = build_mix_units(d.units_out, 2);

  // ---- Lebedev tables
  size_t o_leb_y = 0, o_leb_w = 0;
  if (d.has_beta) {
    std::vector<float> ly((size_t)n_grid * kM * 2), lw(n_grid);
    for (int g = 0; g < n_grid; ++g) {
      double y[kM * 2];
      double x = leb_xyz[g * 3], yy = leb_xyz[g * 3 + 1], z = leb_xyz[g * 3 + 2];
      const double nr = std::sqrt(x * x + yy * yy + z * z);
      if (nr > 0) { x /= nr; yy /= nr; z /= nr; }
      host_sph_harm(x, yy, z, y);
      for (int q = 0; q < kM; ++q) {   // layout [25][n_grid] complex
        ly[((size_t)q * n_grid + g) * 2 + 0] = (float)y[2 * q];
        ly[((size_t)q * n_grid + g) * 2 + 1] = (float)y[2 * q + 1];
      }
      lw[g] = (float)std::log((double)(float)leb_w[g]);   // torch.log(weights) on the float32 weights
    }
    o_leb_y = arena.add(ly.data(), ly.size() * sizeof(float));
    o_leb_w = arena.add(lw.data(), lw.size() * sizeof(float));
  }
  MGB_CUDA_OK(cudaMalloc(&plan->d_tables, arena.host.size()));
  plan->table_bytes = arena.host.size();
  MGB_CUDA_OK(cudaMemcpy(plan->d_tables, arena.host.data(), arena.host.size(), cudaMemcpyHostToDevice));
  for (auto& pt : pending) resolve_table(pt, (const unsigned char*)plan->d_tables);
  for (auto& pg : pending_g) {
    const unsigned char* base = (const unsigned char*)plan->d_tables;
    pg.dst->ag_flat8 = (const int2*)(base + pg.o_ag); pg.dst->ag_slot = (const int*)(base + pg.o_ag_slot);
    pg.dst->sq_flat8 = (const int2*)(base + pg.o_sq); pg.dst->sq_slot = (const int*)(base + pg.o_sq_slot);
  }
  if (d.has_beta) {
    d.leb_y = (const float*)((const unsigned char*)plan->d_tables + o_leb_y);
    d.leb_logw = (const float*)((const unsigned char*)plan->d_tables + o_leb_w);
  }
  MGB_CUDA_OK(cudaMalloc((void**)&plan->d_desc, sizeof(CovDesc)));
  MGB_CUDA_OK(cudaMemcpy(plan->d_desc, &d, sizeof(CovDesc), cudaMemcpyHostToDevice));
  MGB_CUDA_OK(cudaMalloc((void**)&plan->d_segs, sizeof(TransposeSeg) * plan->segs.size()));
  MGB_CUDA_OK(cudaMemcpy(plan->d_segs, plan->segs.data(), sizeof(TransposeSeg) * plan->segs.size(), cudaMemcpyHostToDevice));
  MGB_CUDA_OK(cudaStreamCreateWithFlags(&plan->side, cudaStreamNonBlocking));
  MGB_CUDA_OK(cudaStreamCreateWithFlags(&plan->side2, cudaStreamNonBlocking));
  {
    // side3 carries the half of the atom kernels that the NEXT level waits for: above the weight-gradient streams, below a
    // high-priority main stream (the Python layer captures its graphs on one)
    int lo = 0, hi = 0;
    MGB_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // lo: least priority (numerically largest), hi: greatest
    int mid = hi < lo - 1 ? lo - (lo - hi) / 2 : lo;
    if (const char* e = std::getenv("MGB_SIDE3_PRIO")) mid = std::atoi(e);   // tuning knob
    MGB_CUDA_OK(cudaStreamCreateWithPriority(&plan->side3, cudaStreamNonBlocking, mid));
  }
  for (int q = 0; q <= kMaxLevels; ++q) {
    MGB_CUDA_OK(cudaEventCreateWithFlags(&plan->ev_fork[q], cudaEventDisableTiming));
    MGB_CUDA_OK(cudaEventCreateWithFlags(&plan->ev_join[q], cudaEventDisableTiming));
    MGB_CUDA_OK(cudaEventCreateWithFlags(&plan->ev_fork2[q], cudaEventDisableTiming));
    MGB_CUDA_OK(cudaEventCreateWithFlags(&plan->ev_join2[q], cudaEventDisableTiming));
    MGB_CUDA_OK(cudaEventCreateWithFlags(&plan->ev_fork3[q], cudaEventDisableTiming));
    MGB_CUDA_OK(cudaEventCreateWithFlags(&plan->ev_join3[q], cudaEventDisableTiming));
  }
  *out = plan.release();
  return MGB_OK;
}

void mgb_cov_plan_destroy(mgb_cov_plan* plan) {
  if (!plan) return;
  if (plan->side) cudaStreamDestroy(plan->side);
  if (plan->side2) cudaStreamDestroy(plan->side2);
  if (plan->side3) cudaStreamDestroy(plan->side3);
  for (int q = 0; q <= kMaxLevels; ++q) {
    if (plan->ev_fork[q]) cudaEventDestroy(plan->ev_fork[q]);
    if (plan->ev_join[q]) cudaEventDestroy(plan->ev_join[q]);
    if (plan->ev_fork2[q]) cudaEventDestroy(plan->ev_fork2[q]);
    if (plan->ev_join2[q]) cudaEventDestroy(plan->ev_join2[q]);
    if (plan->ev_fork3[q]) cudaEventDestroy(plan->ev_fork3[q]);
    if (plan->ev_join3[q]) cudaEventDestroy(plan->ev_join3[q]);
  }
  cudaFree(plan->d_tables);
  cudaFree(plan->d_desc);
  cudaFree(plan->d_segs);
  delete plan;
}

int mgb_cov_param_count(const mgb_cov_plan* plan) { return plan ? (int)plan->p_offsets.size() : 0; }

int mgb_cov_param_layout(const mgb_cov_plan* plan, int64_t* offsets, int64_t* numels, int64_t* total) {
  if (!plan) return fail(MGB_ERR_INVALID, "null plan");
  for (size_t i = 0; i < plan->p_offsets.size(); ++i) {
    if (offsets) offsets[i] = plan->p_offsets[i];
    if (numels) numels[i] = plan->p_numels[i];
  }
  if (total) *total = plan->desc.n_params;
  return MGB_OK;
}

int mgb_cov_cat_sizes(const mgb_cov_plan* plan, int32_t* out) {
  if (!plan || !out) return fail(MGB_ERR_INVALID, "null argument");
  const CovDesc& d = plan->desc;
  int q = 0;
  for (int k = 0; k < d.K; ++k)
    for (int l = 0; l < kNL; ++l) { out[q++] = d.lv[k].catE[l]; out[q++] = d.lv[k].catA[l]; }
  for (int l = 0; l < kNL; ++l) out[q++] = d.catM[l];
  return MGB_OK;
}

size_t mgb_cov_workspace_bytes(const mgb_cov_plan* plan, int32_t batch) {
  if (!plan || batch <= 0) return 0;
  return carve_workspace(plan->desc, batch, nullptr).bytes;
}

}  // extern "C"

// The Cormorant body + per-atom heads (everything that does not depend on the action): fills the workspace up to inv / flogit / trans.
static int forward_body(mgb_cov_plan* plan, int32_t B, const float* pos, const int32_t* charges, const float* bags, const float* P,
                        const CovWs& w, const mgb_cov_outputs* out, cudaStream_t st) {
  const CovDesc& d = plan->desc;
  const int N = d.N;
  // the weight transposes (+ L2 prefetch of parameters and tables) run on the side stream beside the input kernels; the first
  // reader of Wt is the level-0 edge kernel
  MGB_CUDA_OK(cudaEventRecord(plan->ev_fork[0], st));
  MGB_CUDA_OK(cudaStreamWaitEvent(plan->side, plan->ev_fork[0], 0));
  MGB_LAUNCH(k_prep_params, (int)((d.n_wt + 255) / 256), 256, 0, plan->side, plan->d_segs, (int)plan->segs.size(), (long long)d.n_wt, P,
             w.Wt, (long long)d.n_params, (const char*)plan->d_tables, (long long)plan->table_bytes);
  MGB_LAUNCH_OK("k_prep_params");
  MGB_CUDA_OK(cudaEventRecord(plan->ev_join[0], plan->side));
  MGB_LAUNCH(k_input_fwd, B, 128, sizeof(float) * N * d.S_in, st, plan->d_desc, P, charges, bags, w.n_atoms, w.X, w.A[0]);
  MGB_LAUNCH_OK("k_input_fwd");
  if (out->covariats)   // padded atoms carry zero representations in the reference; the level kernels skip them
    MGB_CUDA_OK(cudaMemsetAsync(w.A[d.K], 0, sizeof(float) * (size_t)B * N * kM * d.Cout * 2, st));
  MGB_LAUNCH(k_pair_offsets, 1, 1024, 0, st, B, N, w.n_atoms, w.pair_off, w.atom_off, w.atom_list, w.act_off, w.act_list);
  MGB_LAUNCH_OK("k_pair_offsets");
  const unsigned pair_blocks = (unsigned)(((long long)B * N * N + kPairThreads - 1) / kPairThreads);
  for (int k = 0; k < d.K; ++k) {
    const LevelDesc& L = d.lv[k];
    if (small_atoms(B, N)) {
      // the CG-square and pass-through blocks of cat_k only need A_k: side3, beside the dot matrix and the edge kernel
      MGB_CUDA_OK(cudaEventRecord(plan->ev_fork3[k], st));
      MGB_CUDA_OK(cudaStreamWaitEvent(plan->side3, plan->ev_fork3[k], 0));
      int rc3 = k == 0 ? launch_atom_cat<1>(plan, k, B, pos, w, kAtomPhaseB, plan->side3) : launch_atom_cat<kM>(plan, k, B, pos, w, kAtomPhaseB, plan->side3);
      if (rc3 != MGB_OK) return rc3;
      MGB_CUDA_OK(cudaEventRecord(plan->ev_join3[k], plan->side3));
    }
    const size_t dsm = sizeof(float2) * (size_t)N * L.nlm_in * L.C;
    const size_t esm = sizeof(float2) * 70 * kEdgeC + sizeof(float) * (2 * L.C * (kRadFeat + 1));
    dim3 egrid(pair_blocks, kNL);
    const bool small = edge_small(B, N);   // few pairs: five threads per (pair, ell)
    dim3 sgrid((unsigned)(((long long)B * N * N + kPairCsPairs - 1) / kPairCsPairs), kNL);
    const size_t ssm = esm + sizeof(float2) * kPairCsPairs * kEdgeKMaxFwd;
    if (k == 0) {
      MGB_CUDA_OK(cudaFuncSetAttribute(k_dot_fwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
      MGB_LAUNCH(k_dot_fwd<1>, B, 256, dsm, st, plan->d_desc, k, w.n_atoms, w.A[k], w.D[k]);
      MGB_LAUNCH_OK("k_dot_fwd");
      MGB_CUDA_OK(cudaStreamWaitEvent(st, plan->ev_join[0], 0));   // Wt is ready
      if (small) {
        MGB_LAUNCH(k_edge_pairs_fwd_cs<1>, sgrid, kPairCsThreads, ssm, st, plan->d_desc, k, B, P, w.Wt, pos, w.n_atoms, w.pair_off,
                   w.D[k], (const float*)nullptr, w.E[k], w.pair_slot);
      } else {
        MGB_LAUNCH(k_edge_pairs_fwd<1>, egrid, kPairThreads, esm, st, plan->d_desc, k, B, P, w.Wt, pos, w.n_atoms, w.pair_off, w.D[k],
                   (const float*)nullptr, w.E[k], w.pair_slot);
      }
    } else {
      MGB_CUDA_OK(cudaFuncSetAttribute(k_dot_fwd<kNL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
      MGB_LAUNCH(k_dot_fwd<kNL>, B, 256, dsm, st, plan->d_desc, k, w.n_atoms, w.A[k], w.D[k]);
      MGB_LAUNCH_OK("k_dot_fwd");
      if (small) {
        MGB_LAUNCH(k_edge_pairs_fwd_cs<kNL>, sgrid, kPairCsThreads, ssm, st, plan->d_desc, k, B, P, w.Wt, pos, w.n_atoms, w.pair_off,
                   w.D[k], w.E[k - 1], w.E[k], (int*)nullptr);
      } else {
        MGB_LAUNCH(k_edge_pairs_fwd<kNL>, egrid, kPairThreads, esm, st, plan->d_desc, k, B, P, w.Wt, pos, w.n_atoms, w.pair_off, w.D[k],
                   w.E[k - 1], w.E[k], (int*)nullptr);
      }
    }
    MGB_LAUNCH_OK("k_edge_pairs_fwd");
    int rc = k == 0 ? launch_atom_fwd<1>(plan, k, B, P, pos, w, st) : launch_atom_fwd<kM>(plan, k, B, P, pos, w, st);
    if (rc != MGB_OK) return rc;
  }
  MGB_LAUNCH(k_scalars_fwd, B * N, 64, 0, st, plan->d_desc, w.n_atoms, w.A[d.K], w.inv);
  MGB_LAUNCH_OK("k_scalars_fwd");
  return launch_rows_mlp_fwd(plan, B, P, w, st);
}

extern "C" {

int mgb_cov_forward(mgb_cov_plan* plan, int32_t B, const float* pos, const int32_t* charges, const float* bags,
                    const float* actions, const float* P, void* workspace, size_t workspace_bytes,
                    const mgb_cov_outputs* out, void* stream) {
  if (!plan || !pos || !charges || !bags || !actions || !P || !workspace || !out) return fail(MGB_ERR_INVALID, "null argument");
  if (!out->logp || !out->ent || !out->v) return fail(MGB_ERR_INVALID, "logp/ent/v outputs are required");
  if (B <= 0) return fail(MGB_ERR_INVALID, "batch must be positive");
  const CovDesc& d = plan->desc;
  const CovWs w = carve_workspace(d, B, workspace);
  if (w.bytes > workspace_bytes) return fail(MGB_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = forward_body(plan, B, pos, charges, bags, P, w, out, st);
  if (rc != MGB_OK) return rc;
  rc = launch_policy_fwd(plan, B, bags, actions, P, w, out, st);
  if (rc != MGB_OK) return rc;
  if (out->covariats)
    MGB_CUDA_OK(cudaMemcpyAsync(out->covariats, w.A[d.K], sizeof(float) * (size_t)B * d.N * kM * d.Cout * 2, cudaMemcpyDeviceToDevice, st));
  return MGB_OK;
}

int mgb_cov_rollout(mgb_cov_plan* plan, int32_t B, const float* pos, const int32_t* charges, const float* bags, const float* P,
                    void* workspace, size_t workspace_bytes, int32_t mode, uint64_t seed, float* actions, const mgb_cov_outputs* out,
                    void* stream) {
  if (!plan || !pos || !charges || !bags || !P || !workspace || !actions || !out) return fail(MGB_ERR_INVALID, "null argument");
  if (!out->logp || !out->ent || !out->v) return fail(MGB_ERR_INVALID, "logp/ent/v outputs are required");
  if (B <= 0) return fail(MGB_ERR_INVALID, "batch must be positive");
  if (mode != 1 && mode != 2) return fail(MGB_ERR_INVALID, "mode must be 1 (sample) or 2 (greedy)");
  const CovDesc& d = plan->desc;
  const CovWs w = carve_workspace(d, B, workspace);
  if (w.bytes > workspace_bytes) return fail(MGB_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = forward_body(plan, B, pos, charges, bags, P, w, out, st);
  if (rc != MGB_OK) return rc;
  const size_t sm = sizeof(float) * policy_smem_floats(d);
  MGB_CUDA_OK(cudaFuncSetAttribute(k_policy_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  MGB_LAUNCH(k_policy_sample, std::min(B, 148 * 4), kPolicyThreads, sm, st, plan->d_desc, P, w.Wt, B, w.n_atoms, bags, w.A[d.K], w.inv,
             w.flogit, w.trans, mode, (unsigned long long)seed, actions, reinterpret_cast<float2*>(w.lse), *out);
  MGB_LAUNCH_OK("k_policy_sample");
  return MGB_OK;
}

int mgb_cov_policy(mgb_cov_plan* plan, int32_t B, const float* bags, const float* actions, const float* P, void* workspace,
                   size_t workspace_bytes, const mgb_cov_outputs* out, void* stream) {
  if (!plan || !bags || !actions || !P || !workspace || !out) return fail(MGB_ERR_INVALID, "null argument");
  if (!out->logp || !out->ent || !out->v) return fail(MGB_ERR_INVALID, "logp/ent/v outputs are required");
  if (B <= 0) return fail(MGB_ERR_INVALID, "batch must be positive");
  const CovWs w = carve_workspace(plan->desc, B, workspace);
  if (w.bytes > workspace_bytes) return fail(MGB_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
  return launch_policy_fwd(plan, B, bags, actions, P, w, out, (cudaStream_t)stream);
}

}  // extern "C"

#include "api_backward.inl"
