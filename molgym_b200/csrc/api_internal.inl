// api_internal.inl — C ABI of the internal-coordinate (SchNet) actor-critic (part of api.cu).

struct mgb_int_plan {
  mgb_int_config cfg;
  mgb::IntDesc desc;
  mgb::IntDesc* d_desc = nullptr;
  std::vector<long long> p_offsets, p_numels;
};

namespace mgb {
struct IntWs {
  SchWs s;
  IntHeadBufs h;
  DwProblem* dw_probs;
  DwWork* dw_work;
  void* zero_begin;     // [zero_begin, zero_end) is cleared at the start of every backward
  size_t zero_bytes;
  size_t bytes;
};
inline IntWs carve_int_workspace(const IntDesc& d, int B, void* base) {
  IntWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? (void*)((unsigned char*)base + off) : nullptr;
    off += (bytes + 255) & ~size_t(255);
    return p;
  };
  const size_t nm = (size_t)3 * B, R = nm * d.M, Pn = R * d.M, F = d.F, NF = kSchFilters;
  const size_t f4 = sizeof(float);
  w.s.n_mol = (int*)take(sizeof(int) * nm);
  w.s.xs = (float*)take(f4 * 4 * R * F);
  w.s.agg = (float*)take(f4 * 3 * R * NF);
  w.s.u = (float*)take(f4 * 3 * R * F);
  w.dw_probs = (DwProblem*)take(sizeof(DwProblem) * 32);
  w.dw_work = (DwWork*)take(sizeof(DwWork) * 96);
  const size_t z0 = off;
  w.s.z = (float*)take(f4 * 3 * R * F);
  w.s.dv = (float*)take(f4 * 3 * R * F);
  w.s.du = (float*)take(f4 * 3 * R * F);
  w.s.dy = (float*)take(f4 * 3 * R * NF);
  w.s.gauss = (float*)take(f4 * Pn * kSchGauss);
  w.s.h1 = (float*)take(f4 * 3 * Pn * NF);
  w.s.dw2o = (float*)take(f4 * 3 * Pn * NF);
  w.s.dpre1 = (float*)take(f4 * 3 * Pn * NF);
  w.s.dxf = (float*)take(f4 * R * F);
  const size_t b2 = (size_t)2 * B, bn = (size_t)B * d.N, b1 = (size_t)B;
  w.h.Xb = (float*)take(f4 * b2 * d.Z); w.h.Hb = (float*)take(f4 * b2 * d.Wd); w.h.dHb = (float*)take(f4 * b2 * d.Wd); w.h.dOb = (float*)take(f4 * b2 * d.LB);
  w.h.Xf = (float*)take(f4 * bn * d.lat); w.h.Hf = (float*)take(f4 * bn * d.Wd); w.h.dHf = (float*)take(f4 * bn * d.Wd); w.h.dOf = (float*)take(f4 * bn);
  w.h.Xe = (float*)take(f4 * b1 * d.lat); w.h.He = (float*)take(f4 * b1 * d.Wd); w.h.dHe = (float*)take(f4 * b1 * d.Wd); w.h.dOe = (float*)take(f4 * b1 * d.Z);
  w.h.Xc = (float*)take(f4 * b1 * (d.lat + d.Z)); w.h.Hc = (float*)take(f4 * b1 * d.Wd); w.h.dHc = (float*)take(f4 * b1 * d.Wd); w.h.dOc = (float*)take(f4 * b1 * 3);
  w.h.Xk = (float*)take(f4 * b2 * d.lat); w.h.Hk = (float*)take(f4 * b2 * d.Wd); w.h.dHk = (float*)take(f4 * b2 * d.Wd); w.h.dOk = (float*)take(f4 * b2);
  w.h.Xv = (float*)take(f4 * b1 * d.lat); w.h.H1 = (float*)take(f4 * b1 * d.Wd); w.h.H2 = (float*)take(f4 * b1 * d.Wd);
  w.h.dH1 = (float*)take(f4 * b1 * d.Wd); w.h.dH2 = (float*)take(f4 * b1 * d.Wd); w.h.dOv = (float*)take(f4 * b1);
  w.zero_begin = base ? (void*)((unsigned char*)base + z0) : nullptr;
  w.zero_bytes = off - z0;
  w.bytes = off;
  return w;
}
inline size_t sch_fwd_smem(const IntDesc& d) { return sizeof(float) * (2 * d.M * d.F + 2 * d.M * kSchFilters + kSchFilters + 32 + d.M * 3 + 8); }
inline size_t sch_bwd_smem(const IntDesc& d) { return sizeof(float) * (3 * d.M * d.F + 3 * d.M * kSchFilters + 3 * kSchFilters + 32 + d.M * 3 + 8); }
}  // namespace mgb

extern "C" {

int mgb_int_plan_create(const mgb_int_config* cfg, mgb_int_plan** out) {
  if (!cfg || !out) return fail(MGB_ERR_INVALID, "null argument");
  if (cfg->num_species < 1 || cfg->num_species > MGB_MAX_SPECIES) return fail(MGB_ERR_INVALID, "num_species out of range");
  if (cfg->canvas_size < 1 || cfg->canvas_size > 63) return fail(MGB_ERR_INVALID, "canvas_size must be in 1..63");
  if (cfg->network_width < 4 || cfg->network_width > 256 || cfg->network_width % 4) return fail(MGB_ERR_INVALID, "network_width must be a multiple of 4 in 4..256");
  std::unique_ptr<mgb_int_plan> plan(new mgb_int_plan());
  plan->cfg = *cfg;
  IntDesc& d = plan->desc;
  std::memset(&d, 0, sizeof(d));
  d.N = cfg->canvas_size; d.M = d.N + 1; d.Z = cfg->num_species; d.Wd = cfg->network_width;
  d.F = d.Wd / 2; d.LB = d.Wd / 4; d.lat = d.F + d.LB;
  for (int z = 0; z < d.Z; ++z) d.zs[z] = cfg->zs[z];
  d.dmin = cfg->min_distance; d.dmax = cfg->max_distance;
  long long p = 0, wt = 0;
  auto param = [&](long long numel) { plan->p_offsets.push_back(p); plan->p_numels.push_back(numel); long long o = p; p += numel; return o; };
  d.p_emb = param(100ll * d.F);
  for (int t = 0; t < kSchIters; ++t) {
    SchLayer& L = d.it[t];
    L.W1 = param((long long)kSchFilters * kSchGauss); L.b1 = param(kSchFilters);
    L.W2 = param((long long)kSchFilters * kSchFilters); L.b2 = param(kSchFilters);
    L.in2f = param((long long)kSchFilters * d.F);
    L.Wo = param((long long)d.F * kSchFilters); L.bo = param(d.F);
    L.Wd = param((long long)d.F * d.F); L.bd = param(d.F);
  }
  auto mlp = [&](MlpDesc& m, int in, int hidden, int outn) {
    fill_mlp(m, in, hidden, outn, p, wt);
    plan->p_offsets.push_back(m.W0); plan->p_numels.push_back((long long)hidden * in);
    plan->p_offsets.push_back(m.b0); plan->p_numels.push_back(hidden);
    plan->p_offsets.push_back(m.W1); plan->p_numels.push_back((long long)outn * hidden);
    plan->p_offsets.push_back(m.b1); plan->p_numels.push_back(outn);
  };
  mlp(d.beta, d.Z, d.Wd, d.LB);
  mlp(d.focus, d.lat, d.Wd, 1);
  mlp(d.element, d.lat, d.Wd, d.Z);
  mlp(d.cont, d.lat + d.Z, d.Wd, 3);
  mlp(d.kappa, d.lat, d.Wd, 1);
  d.crW0 = param((long long)d.Wd * d.lat); d.crb0 = param(d.Wd);
  d.crW1 = param((long long)d.Wd * d.Wd); d.crb1 = param(d.Wd);
  d.crW2 = param(d.Wd); d.crb2 = param(1);
  d.p_logstd = param(3);
  d.n_params = p;
  MGB_CUDA_OK(cudaMalloc((void**)&plan->d_desc, sizeof(IntDesc)));
  MGB_CUDA_OK(cudaMemcpy(plan->d_desc, &d, sizeof(IntDesc), cudaMemcpyHostToDevice));
  *out = plan.release();
  return MGB_OK;
}

void mgb_int_plan_destroy(mgb_int_plan* plan) {
  if (!plan) return;
  cudaFree(plan->d_desc);
  delete plan;
}
int mgb_int_param_count(const mgb_int_plan* plan) { return plan ? (int)plan->p_offsets.size() : 0; }
int mgb_int_param_layout(const mgb_int_plan* plan, int64_t* offsets, int64_t* numels, int64_t* total) {
  if (!plan) return fail(MGB_ERR_INVALID, "null plan");
  for (size_t i = 0; i < plan->p_offsets.size(); ++i) {
    if (offsets) offsets[i] = plan->p_offsets[i];
    if (numels) numels[i] = plan->p_numels[i];
  }
  if (total) *total = plan->desc.n_params;
  return MGB_OK;
}
size_t mgb_int_workspace_bytes(const mgb_int_plan* plan, int32_t batch) {
  if (!plan || batch <= 0) return 0;
  return carve_int_workspace(plan->desc, batch, nullptr).bytes;
}

int mgb_int_forward(mgb_int_plan* plan, int32_t B, const int32_t* numbers, const float* pos, const float* bags, const float* actions,
                    const float* P, void* workspace, size_t workspace_bytes, const mgb_int_outputs* out, void* stream) {
  if (!plan || !numbers || !pos || !bags || !actions || !P || !workspace || !out) return fail(MGB_ERR_INVALID, "null argument");
  if (!out->logp || !out->ent || !out->v) return fail(MGB_ERR_INVALID, "logp/ent/v outputs are required");
  if (B <= 0) return fail(MGB_ERR_INVALID, "batch must be positive");
  const IntDesc& d = plan->desc;
  const IntWs w = carve_int_workspace(d, B, workspace);
  if (w.bytes > workspace_bytes) return fail(MGB_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  {
    const size_t sm = sch_fwd_smem(d);
    MGB_CUDA_OK(cudaFuncSetAttribute(k_sch_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    MGB_LAUNCH(k_sch_fwd, 3 * B, kSchThreads, sm, st, plan->d_desc, P, numbers, pos, w.s, 3 * B);
    MGB_LAUNCH_OK("k_sch_fwd");
  }
  {
    IntOutputs o{out->logp, out->ent, out->v, out->logp_terms, out->focus_probs, out->element_probs, out->means, out->kappa_logits};
    const size_t sm = sizeof(float) * int_smem_floats(d);
    MGB_CUDA_OK(cudaFuncSetAttribute(k_int_heads_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    MGB_LAUNCH(k_int_heads_fwd, std::min(B, 148 * 4), kIntHeadThreads, sm, st, plan->d_desc, P, B, w.s, bags, actions, o);
    MGB_LAUNCH_OK("k_int_heads_fwd");
  }
  return MGB_OK;
}

int mgb_int_backward(mgb_int_plan* plan, int32_t B, const int32_t* numbers, const float* pos, const float* bags, const float* actions,
                     const float* P, void* workspace, size_t workspace_bytes, const float* g_logp, const float* g_ent,
                     const float* g_v, float* grad, int32_t accumulate, void* stream) {
  if (!plan || !numbers || !pos || !bags || !actions || !P || !workspace || !g_logp || !g_ent || !g_v || !grad)
    return fail(MGB_ERR_INVALID, "null argument");
  if (B <= 0) return fail(MGB_ERR_INVALID, "batch must be positive");
  const IntDesc& d = plan->desc;
  const IntWs w = carve_int_workspace(d, B, workspace);
  if (w.bytes > workspace_bytes) return fail(MGB_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) MGB_CUDA_OK(cudaMemsetAsync(grad, 0, sizeof(float) * d.n_params, st));
  MGB_CUDA_OK(cudaMemsetAsync(w.zero_begin, 0, w.zero_bytes, st));
  {
    const size_t sm = sizeof(float) * (int_smem_floats(d) + int_bwd_extra_floats(d));
    MGB_CUDA_OK(cudaFuncSetAttribute(k_int_heads_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    MGB_LAUNCH(k_int_heads_bwd, std::min(B, 148 * 2), kIntHeadThreads, sm, st, plan->d_desc, P, B, w.s, bags, actions, g_logp, g_ent,
               g_v, w.h, grad);
    MGB_LAUNCH_OK("k_int_heads_bwd");
  }
  {
    const size_t sm = sch_bwd_smem(d);
    MGB_CUDA_OK(cudaFuncSetAttribute(k_sch_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    MGB_LAUNCH(k_sch_bwd, 3 * B, kSchThreads, sm, st, plan->d_desc, P, numbers, pos, w.s, 3 * B, grad);
    MGB_LAUNCH_OK("k_sch_bwd");
  }
  {
    DwProblemList list;
    int q = 0;
    const long long R = 3ll * B * d.M, Pn = R * d.M, F = d.F, NF = kSchFilters;
    auto add = [&](const float* X, const float* dY, long long r, int Kin, int No, long long dW, long long db) {
      list.p[q++] = DwProblem{X, dY, r, Kin, No, kRowsAll, dW, db};
    };
    for (int t = 0; t < kSchIters; ++t) {
      const SchLayer& L = d.it[t];
      add(w.s.gauss, w.s.dpre1 + t * Pn * NF, Pn, kSchGauss, kSchFilters, L.W1, L.b1);
      add(w.s.h1 + t * Pn * NF, w.s.dw2o + t * Pn * NF, Pn, kSchFilters, kSchFilters, L.W2, L.b2);
      add(w.s.xs + t * R * F, w.s.dy + t * R * NF, R, d.F, kSchFilters, L.in2f, -1);
      add(w.s.agg + t * R * NF, w.s.du + t * R * F, R, kSchFilters, d.F, L.Wo, L.bo);
      add(w.s.z + t * R * F, w.s.dv + t * R * F, R, d.F, d.F, L.Wd, L.bd);
    }
    const long long b2 = 2ll * B, bn = (long long)B * d.N;
    add(w.h.Xb, w.h.dHb, b2, d.Z, d.Wd, d.beta.W0, d.beta.b0);       add(w.h.Hb, w.h.dOb, b2, d.Wd, d.LB, d.beta.W1, d.beta.b1);
    add(w.h.Xf, w.h.dHf, bn, d.lat, d.Wd, d.focus.W0, d.focus.b0);   add(w.h.Hf, w.h.dOf, bn, d.Wd, 1, d.focus.W1, d.focus.b1);
    add(w.h.Xe, w.h.dHe, B, d.lat, d.Wd, d.element.W0, d.element.b0); add(w.h.He, w.h.dOe, B, d.Wd, d.Z, d.element.W1, d.element.b1);
    add(w.h.Xc, w.h.dHc, B, d.lat + d.Z, d.Wd, d.cont.W0, d.cont.b0); add(w.h.Hc, w.h.dOc, B, d.Wd, 3, d.cont.W1, d.cont.b1);
    add(w.h.Xk, w.h.dHk, b2, d.lat, d.Wd, d.kappa.W0, d.kappa.b0);   add(w.h.Hk, w.h.dOk, b2, d.Wd, 1, d.kappa.W1, d.kappa.b1);
    add(w.h.Xv, w.h.dH1, B, d.lat, d.Wd, d.crW0, d.crb0);
    add(w.h.H1, w.h.dH2, B, d.Wd, d.Wd, d.crW1, d.crb1);
    add(w.h.H2, w.h.dOv, B, d.Wd, 1, d.crW2, d.crb2);
    list.n = q;
    int nw = 0;
    for (int pi = 0; pi < q; ++pi)
      for (int o0 = 0; o0 < list.p[pi].No; o0 += kDwTileO) list.w[nw++] = DwWork{pi, o0};
    list.nw = nw;
    const int chunks = (int)std::max<long long>(1, std::min<long long>((Pn + 255) / 256, 64));
    dim3 grid(chunks, nw);
    MGB_LAUNCH(k_dw_grouped, grid, kDwThreads, 0, st, list, (const int*)nullptr, d.N, grad);
    MGB_LAUNCH_OK("k_dw_grouped");
  }
  return MGB_OK;
}

}  // extern "C"
