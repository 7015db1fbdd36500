// api_backward.inl — backward launch sequence + PPO loss + host packer (part of api.cu).

template <int NLM2, int CT>
static int launch_atom_bwd_ct(const mgb_cov_plan* plan, int level, int B, const float* pos, const CovWs& w, int accumulate_dE,
                              cudaStream_t st) {
  const CovDesc& d = plan->desc;
  const LevelDesc& L = d.lv[level];
  const size_t smem = sizeof(float) * atom_bwd_smem_floats(L, d.N);
  MGB_CUDA_OK(cudaFuncSetAttribute((k_atom_bwd<NLM2, CT>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MGB_LAUNCH((k_atom_bwd<NLM2, CT>), B * d.N, kAtomBwdThreads, smem, st, plan->d_desc, level, pos, w.n_atoms, w.atom_off, w.atom_list, B,
             w.A[level], w.E[level], w.dcat, w.dA[level & 1], w.dE[level & 1], accumulate_dE,
             small_atoms(B, d.N) ? kAtomPhaseA : kAtomPhaseA | kAtomPhaseB);
  MGB_LAUNCH_OK("k_atom_bwd");
  if (small_atoms(B, d.N)) {
    // column pass + own-atom terms on side3, BEHIND the row pass (the two would only compete for the same SMs) and beside the
    // edge backward -> dot backward part of the main stream
    MGB_CUDA_OK(cudaEventRecord(plan->ev_fork3[level], st));
    MGB_CUDA_OK(cudaStreamWaitEvent(plan->side3, plan->ev_fork3[level], 0));
    MGB_LAUNCH((k_atom_bwd<NLM2, CT>), B * d.N, kAtomBwdThreads, smem, plan->side3, plan->d_desc, level, pos, w.n_atoms, w.atom_off,
               w.atom_list, B, w.A[level], w.E[level], w.dcat, w.dA[level & 1], w.dE[level & 1], accumulate_dE, kAtomPhaseB);
    MGB_LAUNCH_OK("k_atom_bwd");
    MGB_CUDA_OK(cudaEventRecord(plan->ev_join3[level], plan->side3));
  }
  return MGB_OK;
}

template <int NLM2>
static int launch_atom_bwd(const mgb_cov_plan* plan, int level, int B, const float* P, const float* pos, const CovWs& w,
                           int accumulate_dE, bool mix_done, cudaStream_t st) {
  // dcat = W^H dA_{level+1}, row-parallel, into HBM; the atom kernel stages its atom's slice in shared memory
  if (!mix_done) {
    int rc = launch_mix_rows<true>(plan, level, B, P, w, w.dA[(level + 1) & 1], w.dcat, st);
    if (rc != MGB_OK) return rc;
  }

  return plan->desc.lv[level].C == 10 ? launch_atom_bwd_ct<NLM2, 10>(plan, level, B, pos, w, accumulate_dE, st)
                                      : launch_atom_bwd_ct<NLM2, 0>(plan, level, B, pos, w, accumulate_dE, st);
}

extern "C" {

int mgb_cov_backward(mgb_cov_plan* plan, int32_t B, const float* pos, const int32_t* charges, const float* bags,
                     const float* actions, const float* P, void* workspace, size_t workspace_bytes, const float* g_logp,
                     const float* g_ent, const float* g_v, float* grad, int32_t accumulate, void* stream) {
  if (!plan || !pos || !charges || !bags || !actions || !P || !workspace || !g_logp || !g_ent || !g_v || !grad)
    return fail(MGB_ERR_INVALID, "null argument");
  if (B <= 0) return fail(MGB_ERR_INVALID, "batch must be positive");
  const CovDesc& d = plan->desc;
  const CovWs w = carve_workspace(d, B, workspace);
  if (w.bytes > workspace_bytes) return fail(MGB_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int N = d.N, K = d.K;
  const size_t BN = (size_t)B * N;
  const int cmax = std::max(d.C, d.Cout);
  if (!accumulate) MGB_CUDA_OK(cudaMemsetAsync(grad, 0, sizeof(float) * d.n_params, st));
  MGB_CUDA_OK(cudaMemsetAsync(w.dinv, 0, sizeof(float) * BN * d.lat, st));
  MGB_CUDA_OK(cudaMemsetAsync(w.mix_stage, 0, sizeof(float) * 2 * d.totWM, st));
  MGB_CUDA_OK(cudaMemsetAsync(w.dA[K & 1], 0, sizeof(float) * BN * kM * cmax * 2, st));
  {
    PolicyBwdOut o{w.finv, w.he, w.einv, w.hd, w.vf, w.hv, w.dhe, w.dye, w.dhd, w.dyd, w.dhv, w.dyv, w.dvf, w.dflogit, w.dinv, w.dA[K & 1]};
    const size_t sm = sizeof(float) * (policy_smem_floats(d) + policy_bwd_extra_floats(d));
    MGB_CUDA_OK(cudaFuncSetAttribute(k_policy_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const int grid = std::min(B, 148 * MGB_POLICY_BWD_MIN_CTAS);
    MGB_LAUNCH(k_policy_bwd, grid, kPolicyBwdThreads, sm, st, plan->d_desc, P, w.Wt, B, w.n_atoms, bags, actions, w.A[K], w.inv, w.flogit,
               w.trans, w.pol_state, g_logp, g_ent, g_v, o, w.mix_stage, grad);
    MGB_LAUNCH_OK("k_policy_bwd");
  }
  {
    int rc = launch_rows_mlp_bwd(plan, B, P, w, st);
    if (rc != MGB_OK) return rc;
  }
  MGB_LAUNCH(k_scalars_bwd, B * N, 64, 0, st, plan->d_desc, w.n_atoms, w.A[K], w.dinv, w.dA[K & 1]);
  MGB_LAUNCH_OK("k_scalars_bwd");
  cudaStream_t side = plan->side;
  // fork: the MLP weight gradients only need the head kernels' outputs; they run beside the CG levels
  MGB_CUDA_OK(cudaEventRecord(plan->ev_fork[K], st));
  MGB_CUDA_OK(cudaStreamWaitEvent(side, plan->ev_fork[K], 0));
  MGB_LAUNCH(k_mixer_dw_finish, 2, 256, 0, side, plan->d_desc, w.mix_stage, grad);   // expands the compact mixer cotangent
  MGB_LAUNCH_OK("k_mixer_dw_finish");
  {
    DwProblemList list;
    int q = 0;
    const long long rows = (long long)BN;
    auto add = [&](const float* X, const float* dY, long long r, int Kin, int No, int mode, long long dW, long long db) {
      list.p[q++] = DwProblem{X, dY, r, Kin, No, mode, dW, db};
    };
    add(w.inv, w.dhf, rows, d.lat, d.Wd, kRowsActive, d.focus.W0, d.focus.b0);
    add(w.hf, w.dflogit, rows, d.Wd, 1, kRowsActive, d.focus.W1, d.focus.b1);
    add(w.inv, w.dht0, rows, d.lat, d.Wd, kRowsValid, d.trans.W0, d.trans.b0);
    add(w.ht0, w.dtrans, rows, d.Wd, d.Wd, kRowsValid, d.trans.W1, d.trans.b1);
    add(w.finv, w.dhe, B, d.lat, d.Wd, kRowsAll, d.element.W0, d.element.b0);
    add(w.he, w.dye, B, d.Wd, d.Z, kRowsAll, d.element.W1, d.element.b1);
    add(w.einv, w.dhd, B, d.latE, d.Wd, kRowsAll, d.dist.W0, d.dist.b0);
    add(w.hd, w.dyd, B, d.Wd, 2 * d.G, kRowsAll, d.dist.W1, d.dist.b1);
    add(w.vf, w.dhv, B, d.Wd, d.Wd, kRowsAll, d.value.W0, d.value.b0);
    add(w.hv, w.dyv, B, d.Wd, 1, kRowsAll, d.value.W1, d.value.b1);
    list.n = q;
    int nw = 0;
    for (int pi = 0; pi < q; ++pi)
      for (int o0 = 0; o0 < list.p[pi].No; o0 += kDwTileO) list.w[nw++] = DwWork{pi, o0};
    list.nw = nw;
    const int chunks = (int)std::max<size_t>(1, std::min<size_t>((BN + 63) / 64, 148));
    dim3 grid(chunks, nw);
    MGB_LAUNCH(k_dw_grouped, grid, kDwThreads, 0, side, list, w.n_atoms, N, grad);
    MGB_LAUNCH_OK("k_dw_grouped");
  }
  for (int k = K - 1; k >= 0; --k) {
    const LevelDesc& L = d.lv[k];
    if (k < K - 1) MGB_CUDA_OK(cudaStreamWaitEvent(st, plan->ev_join[k + 1], 0));   // mix_dw(k+1) still reads dA[(k+2)&1] == dA[k&1]
    if (k < K - 1 && small_atoms(B, N)) MGB_CUDA_OK(cudaStreamWaitEvent(st, plan->ev_join3[k + 1], 0));   // column pass of level k+1: dA_{k+1} complete, dcat free
    MGB_CUDA_OK(cudaMemsetAsync(w.dA[k & 1], 0, sizeof(float) * BN * kM * cmax * 2, st));
    // large minibatches: the weight gradient (streams cat_k from HBM) is forked BEHIND the dcat mix (streams dcat to HBM), so that it
    // runs beside the FP32-bound atom kernel instead of competing with the other bandwidth-bound kernel of the level
    bool mix_first = !small_atoms(B, N);
    if (const char* e = std::getenv("MGB_MIXDW_LATE")) mix_first = e[0] == '1';
    if (mix_first) {
      int rc0 = launch_mix_rows<true>(plan, k, B, P, w, w.dA[(k + 1) & 1], w.dcat, st);
      if (rc0 != MGB_OK) return rc0;
    }
    {
      // fork: the atom-mix weight gradient (reads cat_k and dA_{k+1}, both complete here) runs beside the level
      MGB_CUDA_OK(cudaEventRecord(plan->ev_fork[k], st));
      MGB_CUDA_OK(cudaStreamWaitEvent(side, plan->ev_fork[k], 0));
      const int chunks = (int)std::max<size_t>(1, std::min<size_t>((BN + kMixDwAtoms - 1) / kMixDwAtoms, 148 * 2));
      int kmax = 0;
      for (int l = 0; l < kNL; ++l) kmax = std::max(kmax, L.catA[l]);
      dim3 grid(chunks, kNL, (kmax + kMixDwThreads - 1) / kMixDwThreads);
      int co = std::min(pick_co_rows(L.Cout), 16 + 4 * (L.Cout == 20));   // more than 20 channels: passes of 16
      // 20 channels in one pass need 238 registers (8 warps per SM): from a few thousand atoms on two passes of 10 (96 registers)
      // are faster although cat is read twice (C3 b1024 step 6.27 -> 6.11 ms, C4 b1024 12.74 -> 12.48 ms; no difference at b128)
      if (L.Cout == 20 && large_atoms(B, N)) co = 10;
      const size_t sm = sizeof(float2) * kMixDwAtoms * 9 * co;
#define MGB_MIXDW_CASE(CO)                                                                                                        \
  case CO:                                                                                                                        \
    for (int cb = 0; cb < L.Cout; cb += CO) {                                                                                     \
      MGB_LAUNCH(k_mix_dw<CO>, grid, kMixDwThreads, sm, side, plan->d_desc, k, B, w.atom_off, w.atom_list, w.cat[k],              \
                 w.dA[(k + 1) & 1], cb, grad);                                                                                    \
    }                                                                                                                             \
    break;
      switch (co) {
        MGB_MIXDW_CASE(4)
        MGB_MIXDW_CASE(8)
        MGB_MIXDW_CASE(10)
        MGB_MIXDW_CASE(12)
        MGB_MIXDW_CASE(16)
        MGB_MIXDW_CASE(20)
      }
#undef MGB_MIXDW_CASE
      MGB_LAUNCH_OK("k_mix_dw");
      MGB_CUDA_OK(cudaEventRecord(plan->ev_join[k], side));
    }
    const int acc_dE = (k < K - 1) ? 1 : 0;
    int rc = k == 0 ? launch_atom_bwd<1>(plan, k, B, P, pos, w, acc_dE, mix_first, st) : launch_atom_bwd<kM>(plan, k, B, P, pos, w, acc_dE, mix_first, st);
    if (rc != MGB_OK) return rc;
    {
      const unsigned pair_blocks = (unsigned)((BN * N + kPairThreads - 1) / kPairThreads);
      const size_t esm = sizeof(float2) * (size_t)L.sumCatE * kEdgeC + sizeof(float) * (kNL * 2 * L.C * (kRadFeat + 1));
      EdgeScratch sc{w.e_dpre, w.e_R, w.e_dR, w.e_f};
      const int dw_chunks = (int)std::max<size_t>(1, std::min<size_t>((BN * N + kEdgeDwTile - 1) / kEdgeDwTile, 148 * 2));
      dim3 dwgrid(dw_chunks, kNL);
      const bool split = edge_bwd_split(B, N), small = edge_small(B, N);
      dim3 sgrid((unsigned)((BN * N + kPairCsPairs - 1) / kPairCsPairs), kNL);
      const size_t ssm = sizeof(float2) * 70 * kEdgeC + sizeof(float) * (2 * L.C * (kRadFeat + 1));
      const long long slice = (long long)BN * N * kNL * L.C;
      dim3 pgrid(pair_blocks, split ? kNL : 1);
      // the scratch of the previous (higher) level's edge backward is still being reduced by k_edge_dw on side2
      if (k < K - 1) MGB_CUDA_OK(cudaStreamWaitEvent(st, plan->ev_join2[k + 1], 0));
#define MGB_EDGE_BWD(NL, EPREV, DEPREV, DOTTHREADS)                                                                              \
  if (small) {                                                                                                                      \
    MGB_LAUNCH(k_edge_pairs_bwd_cs<NL>, sgrid, kPairCsThreads, ssm, st, plan->d_desc, k, B, P, pos, w.n_atoms, w.pair_off,          \
               w.dE[k & 1], DEPREV, w.dD, slice, sc, grad);                                                                         \
  } else if (split) {                                                                                                                    \
    MGB_CUDA_OK(cudaFuncSetAttribute((k_edge_pairs_bwd<NL, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esm));         \
    MGB_LAUNCH((k_edge_pairs_bwd<NL, true>), pgrid, kPairThreads, esm, st, plan->d_desc, k, B, P, pos, w.n_atoms, w.pair_off,       \
               w.dE[k & 1], DEPREV, w.dD, slice, sc, grad);                                                                         \
  } else {                                                                                                                          \
    MGB_CUDA_OK(cudaFuncSetAttribute((k_edge_pairs_bwd<NL, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esm));        \
    MGB_LAUNCH((k_edge_pairs_bwd<NL, false>), pgrid, kPairThreads, esm, st, plan->d_desc, k, B, P, pos, w.n_atoms, w.pair_off,      \
               w.dE[k & 1], DEPREV, w.dD, slice, sc, grad);                                                                         \
  }                                                                                                                                 \
  MGB_LAUNCH_OK("k_edge_pairs_bwd");                                                                                                \
  /* fork: the edge weight gradients (reductions over pairs of the scratch just written) run on side2 */                            \
  MGB_CUDA_OK(cudaEventRecord(plan->ev_fork2[k], st));                                                                              \
  MGB_CUDA_OK(cudaStreamWaitEvent(plan->side2, plan->ev_fork2[k], 0));                                                              \
  MGB_CUDA_OK(cudaFuncSetAttribute(k_edge_dw<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * sizeof(float) * kEdgeDwBufFloats))); \
  MGB_LAUNCH(k_edge_dw<NL>, dwgrid, kEdgeDwThreads, 2 * sizeof(float) * kEdgeDwBufFloats, plan->side2, plan->d_desc, k, B, w.pair_off, w.pair_slot, EPREV, w.D[k], sc, grad); \
  MGB_LAUNCH_OK("k_edge_dw");                                                                                                       \
  MGB_CUDA_OK(cudaEventRecord(plan->ev_join2[k], plan->side2));                                                                     \
  MGB_LAUNCH(k_dot_bwd<NL>, B * N, DOTTHREADS, 0, st, plan->d_desc, k, w.n_atoms, w.A[k], w.dD, split ? NL : 1, slice, w.dA[k & 1]);
      if (k == 0) {
        MGB_EDGE_BWD(1, (const float*)nullptr, (float*)nullptr, 64)
      } else {
        MGB_EDGE_BWD(kNL, w.E[k - 1], w.dE[(k - 1) & 1], 256)
      }
#undef MGB_EDGE_BWD
      MGB_LAUNCH_OK("k_dot_bwd");
    }
  }
  {
    // InputLinear weight gradient (needs the complete dA_0): the tail of the main stream
    if (small_atoms(B, N)) MGB_CUDA_OK(cudaStreamWaitEvent(st, plan->ev_join3[0], 0));
    if (small_atoms(B, N)) {
      const int per_cta = (B + 147) / 148;
      MGB_LAUNCH(k_input_dw, (B + per_cta - 1) / per_cta, 256, 0, st, plan->d_desc, B, per_cta, w.n_atoms, w.X, w.dA[0], grad);
      MGB_LAUNCH_OK("k_input_dw");
    } else {   // many rows: the tiled grouped weight-gradient kernel
      DwProblemList list;
      list.p[0] = DwProblem{w.X, w.dA[0], (long long)BN, d.S_in, 2 * d.C, kRowsValid, d.p_inW, d.p_inb};
      list.n = 1;
      int nw = 0;
      for (int o0 = 0; o0 < list.p[0].No; o0 += kDwTileO) list.w[nw++] = DwWork{0, o0};
      list.nw = nw;
      const int chunks = (int)std::max<size_t>(1, std::min<size_t>((BN + 63) / 64, 148));
      dim3 grid(chunks, nw);
      MGB_LAUNCH(k_dw_grouped, grid, kDwThreads, 0, st, list, w.n_atoms, N, grad);
      MGB_LAUNCH_OK("k_dw_grouped");
    }
  }
  MGB_CUDA_OK(cudaStreamWaitEvent(st, plan->ev_join[0], 0));    // join the side streams (their last work is mix_dw(0), edge_dw(0))
  MGB_CUDA_OK(cudaStreamWaitEvent(st, plan->ev_join2[0], 0));
  return MGB_OK;
}

int mgb_ppo_loss(int32_t B, const float* logp, const float* ent, const float* v, const float* old_logp, const double* adv,
                 const double* ret, double clip_ratio, double vf_coef, double entropy_coef, double inv_global_batch,
                 double* info, float* g_logp, float* g_ent, float* g_v, void* stream) {
  if (!logp || !ent || !v || !old_logp || !adv || !ret || !info) return fail(MGB_ERR_INVALID, "null argument");
  if (B <= 0) return fail(MGB_ERR_INVALID, "batch must be positive");
  if ((g_logp || g_ent || g_v) && !(g_logp && g_ent && g_v)) return fail(MGB_ERR_INVALID, "gradient outputs must be all set or all NULL");
  MGB_LAUNCH(k_ppo_loss, 1, 256, 0, (cudaStream_t)stream, B, logp, ent, v, old_logp, adv, ret, clip_ratio, vf_coef, entropy_coef,
             inv_global_batch, info, g_logp, g_ent, g_v);
  MGB_LAUNCH_OK("k_ppo_loss");
  return MGB_OK;
}

int mgb_scale_accumulate(float* dst, const float* src, const void* scale, int32_t scale_is_double, int64_t n, int32_t accumulate,
                         void* stream) {
  if (!dst || !src || !scale) return fail(MGB_ERR_INVALID, "null argument");
  if (n <= 0) return MGB_OK;
  if (((uintptr_t)dst & 15) || ((uintptr_t)src & 15)) return fail(MGB_ERR_INVALID, "dst / src must be 16-byte aligned");
  const int grid = (int)std::min<int64_t>((n / 4 + 255) / 256 + 1, 148 * 4);
  MGB_LAUNCH(k_scale_accumulate, grid, 256, 0, (cudaStream_t)stream, dst, src, scale, scale_is_double, (long long)n, accumulate);
  MGB_LAUNCH_OK("k_scale_accumulate");
  return MGB_OK;
}

size_t mgb_optim_scratch_bytes(void) { return sizeof(double) * (kNormMaxBlocks + 4); }

int mgb_grad_norm(const float* grad, int64_t n, void* scratch, double* norm, void* stream) {
  if (!grad || !scratch || !norm) return fail(MGB_ERR_INVALID, "null argument");
  if (n <= 0) return fail(MGB_ERR_INVALID, "n must be positive");
  if ((uintptr_t)grad & 15) return fail(MGB_ERR_INVALID, "grad must be 16-byte aligned");
  const int grid = (int)std::min<int64_t>((n / 4 + kNormThreads - 1) / kNormThreads + 1, kNormMaxBlocks);
  MGB_LAUNCH(k_grad_norm, grid, kNormThreads, 0, (cudaStream_t)stream, grad, (long long)n, (double*)scratch, norm);
  MGB_LAUNCH_OK("k_grad_norm");
  return MGB_OK;
}

int mgb_adam_step(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq, int64_t n, double lr,
                  double beta1, double beta2, double eps, double weight_decay, int64_t step, int32_t amsgrad, int32_t maximize,
                  const double* norm, double max_norm, void* stream) {
  if (!params || !grad || !exp_avg || !exp_avg_sq) return fail(MGB_ERR_INVALID, "null argument");
  if (amsgrad && !max_exp_avg_sq) return fail(MGB_ERR_INVALID, "amsgrad needs max_exp_avg_sq");
  if (n <= 0 || step <= 0) return fail(MGB_ERR_INVALID, "n and step must be positive");
  if (max_norm > 0 && !norm) return fail(MGB_ERR_INVALID, "clipping needs the device norm (mgb_grad_norm)");
  AdamArgs a;
  a.lr = (float)lr; a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps; a.weight_decay = (float)weight_decay;
  a.step_size = (float)(lr / (1.0 - std::pow(beta1, (double)step)));
  a.bias2_sqrt = (float)std::sqrt(1.0 - std::pow(beta2, (double)step));
  a.max_norm = (float)max_norm; a.amsgrad = amsgrad; a.maximize = maximize;
  const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  MGB_LAUNCH(k_adam_step, grid, 256, 0, (cudaStream_t)stream, params, grad, exp_avg, exp_avg_sq, max_exp_avg_sq, (long long)n, norm, a);
  MGB_LAUNCH_OK("k_adam_step");
  return MGB_OK;
}

int64_t mgb_launch_count(void) { return g_prof.launches; }

int mgb_profile_kernel(const char* substr) {
#ifndef MGB_CUSIM
  for (cudaEvent_t e : g_prof.ev) cudaEventDestroy(e);
  g_prof.ev.clear();
  g_prof.names.clear();
  g_prof.active = substr && substr[0];
  std::snprintf(g_prof.pattern, sizeof(g_prof.pattern), "%s", substr ? substr : "");
#endif
  return MGB_OK;
}

int mgb_profile_read(double* total_ms, int64_t* launches) {
  double tot = 0.0;
  int64_t n = 0;
#ifndef MGB_CUSIM
  for (size_t i = 0; i + 1 < g_prof.ev.size(); i += 2) {
    MGB_CUDA_OK(cudaEventSynchronize(g_prof.ev[i + 1]));
    float ms = 0.f;
    MGB_CUDA_OK(cudaEventElapsedTime(&ms, g_prof.ev[i], g_prof.ev[i + 1]));
    tot += ms;
    ++n;
  }
  for (cudaEvent_t e : g_prof.ev) cudaEventDestroy(e);
  g_prof.ev.clear();
  g_prof.names.clear();
#endif
  if (total_ms) *total_ms = tot;
  if (launches) *launches = n;
  return MGB_OK;
}

int mgb_profile_report(char* buf, int64_t cap) {
  if (!buf || cap <= 0) return fail(MGB_ERR_INVALID, "null buffer");
  buf[0] = 0;
#ifndef MGB_CUSIM
  // one line per timed launch, in launch order: "<kernel> <milliseconds>"
  int64_t used = 0;
  for (size_t i = 0; i + 1 < g_prof.ev.size(); i += 2) {
    MGB_CUDA_OK(cudaEventSynchronize(g_prof.ev[i + 1]));
    float ms = 0.f;
    MGB_CUDA_OK(cudaEventElapsedTime(&ms, g_prof.ev[i], g_prof.ev[i + 1]));
    const int n = std::snprintf(buf + used, (size_t)(cap - used), "%s %.6f\n", g_prof.names[i / 2], ms);
    if (n < 0 || used + n >= cap) break;
    used += n;
  }
  for (cudaEvent_t e : g_prof.ev) cudaEventDestroy(e);
  g_prof.ev.clear();
  g_prof.names.clear();
#endif
  return MGB_OK;
}

int mgb_pack_observations(const mgb_cov_config* cfg, int32_t B, const int32_t* labels, const double* xyz, float* positions,
                          int32_t* charges) {
  if (!cfg || !labels || !xyz || !positions || !charges) return fail(MGB_ERR_INVALID, "null argument");
  const int N = cfg->canvas_size, Z = cfg->num_species;
  for (int b = 0; b < B; ++b) {
    int k = 0;
    for (int i = 0; i < N; ++i) {
      const int lab = labels[(size_t)b * N + i];
      if (lab < 0 || lab >= Z) return fail(MGB_ERR_INVALID, "canvas %d item %d: label %d outside [0, %d)", b, i, lab, Z);
      if (cfg->zs[lab] == 0) continue;
      charges[(size_t)b * N + k] = cfg->zs[lab];
      for (int a = 0; a < 3; ++a) positions[((size_t)b * N + k) * 3 + a] = (float)xyz[((size_t)b * N + i) * 3 + a];
      ++k;
    }
    for (; k < N; ++k) {
      charges[(size_t)b * N + k] = 0;
      for (int a = 0; a < 3; ++a) positions[((size_t)b * N + k) * 3 + a] = 0.f;
    }
  }
  return MGB_OK;
}

}  // extern "C"

#include "api_internal.inl"
