// plan.cuh — host-side plan: descriptors, Clebsch-Gordan tables, parameter layout, workspace carving.
#pragma once
#include <cstdlib>
#include <cstdarg>
#include <memory>

#include "heads.cuh"
#include "edge.cuh"

struct mgb_cov_plan {
  mgb_cov_config cfg;
  mgb::CovDesc desc;                 // host copy (device pointers inside)
  mgb::CovDesc* d_desc = nullptr;    // device copy
  void* d_tables = nullptr;          // one allocation holding every table
  size_t table_bytes = 0;
  std::vector<mgb::TransposeSeg> segs;
  mgb::TransposeSeg* d_segs = nullptr;
  std::vector<long long> p_offsets, p_numels;
  // fork/join of independent backward kernels (weight-gradient reductions next to the edge level)
  cudaStream_t side = nullptr, side2 = nullptr, side3 = nullptr;
  cudaEvent_t ev_fork[mgb::kMaxLevels + 1] = {}, ev_join[mgb::kMaxLevels + 1] = {};
  cudaEvent_t ev_fork2[mgb::kMaxLevels + 1] = {}, ev_join2[mgb::kMaxLevels + 1] = {};   // edge weight gradients (side2)
  cudaEvent_t ev_fork3[mgb::kMaxLevels + 1] = {}, ev_join3[mgb::kMaxLevels + 1] = {};   // small minibatches: half of the atom level (side3)
};

namespace mgb {

inline thread_local char g_err[512] = "";
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define MGB_CUDA_OK(expr)                                                                              \
  do {                                                                                                 \
    cudaError_t e_ = (expr);                                                                           \
    if (e_ != cudaSuccess) return mgb::fail(MGB_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

// Bump allocator over one host staging buffer mirrored to the device.
struct TableArena {
  std::vector<unsigned char> host;
  size_t add(const void* p, size_t bytes) {
    size_t off = (host.size() + 127) & ~size_t(127);   // whole 128-byte lines: some tables are read one line per warp load
    host.resize(off + bytes);
    std::memcpy(host.data() + off, p, bytes);
    return off;
  }
};
struct PendingTable {
  CgTable* dst;
  size_t o_out_l, o_out_m, o_out_block, o_term_start, o_lm1, o_lm2, o_coef, o_pstart, o_pout, o_pcoef, o_dst, o_src, o_ent, o_flat, o_slot, o_pad, o_sym;
  bool has_sym;
};
inline PendingTable stage_table(TableArena& a, const HostCgTable& h, CgTable* dst) {
  dst->n_out = h.n_out; dst->n_pair = h.n_pair; dst->nlm2 = h.nlm2;
  PendingTable p;
  p.dst = dst;
  auto vi = [&](const std::vector<int>& v) { return a.add(v.data(), v.size() * sizeof(int)); };
  auto vf = [&](const std::vector<float>& v) { return a.add(v.data(), v.size() * sizeof(float)); };
  p.o_out_l = vi(h.out_l); p.o_out_m = vi(h.out_m); p.o_out_block = vi(h.out_block); p.o_term_start = vi(h.term_start);
  p.o_lm1 = vi(h.term_lm1); p.o_lm2 = vi(h.term_lm2); p.o_coef = vf(h.term_coef);
  p.o_pstart = vi(h.pair_start); p.o_pout = vi(h.pair_out); p.o_pcoef = vf(h.pair_coef);
  p.o_dst = vi(h.out_dst); p.o_src = vi(h.term_src); p.o_ent = vi(h.pair_ent);
  p.o_flat = vi(h.flat); p.o_slot = vi(h.slot_start); p.o_pad = vi(h.pad_pair);
  p.has_sym = !h.pad_sym.empty();
  p.o_sym = p.has_sym ? vi(h.pad_sym) : 0;
  dst->n_slots = h.n_slots;
  return p;
}
inline void resolve_table(const PendingTable& p, const unsigned char* base) {
  CgTable* t = p.dst;
  t->out_l = (const int*)(base + p.o_out_l); t->out_m = (const int*)(base + p.o_out_m);
  t->out_block = (const int*)(base + p.o_out_block); t->term_start = (const int*)(base + p.o_term_start);
  t->term_lm1 = (const int*)(base + p.o_lm1); t->term_lm2 = (const int*)(base + p.o_lm2);
  t->term_coef = (const float*)(base + p.o_coef); t->pair_start = (const int*)(base + p.o_pstart);
  t->pair_out = (const int*)(base + p.o_pout); t->pair_coef = (const float*)(base + p.o_pcoef);
  t->out_dst = (const int*)(base + p.o_dst); t->term_src = (const int2*)(base + p.o_src);
  t->pair_ent = (const int2*)(base + p.o_ent);
  t->flat = (const int4*)(base + p.o_flat); t->slot_start = (const int*)(base + p.o_slot);
  t->pad_pair = (const int2*)(base + p.o_pad);
  t->pad_sym = p.has_sym ? (const int2*)(base + p.o_sym) : nullptr;
}

// Complex spherical harmonics on the host in double (same closed forms as sph_harm_l4), 'qm' norm, no conjugation,
// unit-vector argument: used for the Lebedev table.
inline void host_sph_harm(double x, double y, double z, double* out /* [25][2] */) {
  const double r2 = x * x + y * y + z * z, z2 = z * z;
  double e[5][2] = {{1, 0}, {x, y}, {x * x - y * y, 2 * x * y}, {0, 0}, {0, 0}};
  e[3][0] = e[2][0] * x - e[2][1] * y; e[3][1] = e[2][0] * y + e[2][1] * x;
  e[4][0] = e[2][0] * e[2][0] - e[2][1] * e[2][1]; e[4][1] = 2 * e[2][0] * e[2][1];
  double dlm[5][5] = {};
  dlm[0][0] = 1;
  dlm[1][0] = z; dlm[1][1] = 1;
  dlm[2][0] = 0.5 * (3 * z2 - r2); dlm[2][1] = 3 * z; dlm[2][2] = 3;
  dlm[3][0] = 0.5 * z * (5 * z2 - 3 * r2); dlm[3][1] = 0.5 * (15 * z2 - 3 * r2); dlm[3][2] = 15 * z; dlm[3][3] = 15;
  dlm[4][0] = 0.125 * (35 * z2 * z2 - 30 * z2 * r2 + 3 * r2 * r2); dlm[4][1] = 0.5 * z * (35 * z2 - 15 * r2);
  dlm[4][2] = 0.5 * (105 * z2 - 15 * r2); dlm[4][3] = 105 * z; dlm[4][4] = 105;
  const double pi = 3.14159265358979323846;
  for (int l = 0; l <= 4; ++l)
    for (int m = 0; m <= l; ++m) {
      const double nrm = std::sqrt((2 * l + 1) / (4 * pi) * factorial_d(l - m) / factorial_d(l + m));
      const double a = nrm * dlm[l][m];
      const double sg = (m & 1) ? -1.0 : 1.0;
      out[(l * l + l + m) * 2 + 0] = sg * a * e[m][0];
      out[(l * l + l + m) * 2 + 1] = sg * a * e[m][1];
      if (m > 0) {
        out[(l * l + l - m) * 2 + 0] = a * e[m][0];
        out[(l * l + l - m) * 2 + 1] = -a * e[m][1];
      }
    }
}

inline void fill_mlp(MlpDesc& m, int in, int hidden, int out, long long& p, long long& wt) {
  m.in = in; m.hidden = hidden; m.out = out;
  // every tensor starts on a 16-byte boundary of the flat buffer (bulk copies into shared memory need it)
  auto al = [](long long v) { return (v + 3) & ~3ll; };
  p = al(p); m.W0 = p; p += (long long)hidden * in;
  p = al(p); m.b0 = p; p += hidden;
  p = al(p); m.W1 = p; p += (long long)out * hidden;
  p = al(p); m.b1 = p; p += out;
  p = al(p);
  m.W0t = wt; wt += (long long)hidden * in;
  m.W1t = wt; wt += (long long)out * hidden;
}

// ------------------------------------------------------------------------------------------------------------
// Workspace carving (device pointers).  Everything a forward leaves for its backward lives here.
// ------------------------------------------------------------------------------------------------------------
struct CovWs {
  int* n_atoms;
  int* pair_off;                  // [B+1] prefix of n_b^2 (flat list of valid pairs)
  int* pair_slot;                 // [B*N*N] dense slot (b*N+i)*N+j of every flat pair (written by the level-0 edge kernel)
  int* atom_off;                  // [B+1] prefix of n_b
  int* atom_list;                 // [B*N] slot index b*N+i of every valid atom
  int* act_off;                   // [B+1] prefix of max(n_b, 1)
  int* act_list;                  // [B*N] slot index of every active row (i < max(n_b, 1): what the focus head sees)
  float* dcat;                    // [B,N,max totA,2] cotangent of the cat vectors of the level being differentiated
  float* Wt;                      // transposed weights scratch
  float* X;                       // [B,N,S_in]
  float* A[kMaxLevels + 1];       // A[0] [B,N,1,C,2]; A[k] [B,N,25,C_k,2]
  float* E[kMaxLevels];           // [B,N,N,5,C,2]
  float* cat[kMaxLevels];         // [B,N,totA_k,2]
  float* inv;                     // [B,N,lat]
  float* lse;                     // [B,2] running max / sum of the log Z quadrature
  float* pol_state;               // [B, policy_state_floats] intermediates of k_policy_fwd, read by k_policy_bwd
  float* hf, *flogit, *ht0, *trans;       // rows
  // backward
  float* finv, *he, *einv, *hd, *vf, *hv;  // per canvas activations saved by policy_bwd for the weight gradients
  float* dhe, *dye, *dhd, *dyd, *dhv, *dyv, *dvf;
  float* dflogit, *dhf, *dht0, *dtrans, *dinv;
  float* dA[2];                   // ping-pong, each [B,N,25,Cmax,2]
  float* dE[2];                   // ping-pong, each [B,N,N,5,C,2]
  float* dD;                      // [B,N,N,5C,2]
  float* D[kMaxLevels];           // [B,N,N,5C,2] dot matrix of every level (saved by the forward for the edge backward)
  float* e_dpre, *e_R, *e_dR, *e_f;   // per-pair scratch of the edge backward (flat pair index)
  double* loss_acc;               // [16]
  float* mix_stage;               // [sum_l catM, 2] compact mixer-weight cotangent
  DwProblem* dw_probs;            // [16]
  DwWork* dw_work;                // [64]
  size_t bytes;
};

// small minibatches: the per-pair edge backward is split over the five ells (one thread per (pair, ell))
// MGB_EDGE_MODE=0|1|2 forces the decomposition of the per-pair edge kernels (tests): 0 thread per pair, 1 thread per
// (pair, ell), 2 five threads per (pair, ell)
inline int edge_mode_override() {
  const char* e = std::getenv("MGB_EDGE_MODE");
  return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : -1;
}
inline bool edge_bwd_split(int B, int N) {
  const int m = edge_mode_override();
  return m >= 0 ? m >= 1 : (long long)B * N * N < 148ll * 2048;
}
// small minibatches (the grid of one-atom CTAs is about one wave): the atom kernels are launched as two half kernels, the
// half that is off the critical path on a side stream (forward: CG square + pass-through blocks of cat, which only need
// A_k; backward: column pass + own-atom terms, which the edge backward does not wait for)
inline bool small_atoms(int B, int N) {
  const char* e = std::getenv("MGB_SMALL_ATOMS");   // tests / tuning: override the slot threshold
  const long long limit = e ? std::atoll(e) : 2560;   // measured: helps up to ~1.5 k atom slots (C2/140, C3/128), hurts at C4/512
  return (long long)B * N < limit;
}
// large minibatches (a few thousand atom slots and more): tensor-core forward mix, 512-thread dcat mix, two-pass 20-channel mix weight
// gradient.  MGB_LARGE_ATOMS overrides the slot threshold (the parity tests force the large path on small batches with 1).
inline bool large_atoms(int B, int N) {
  const char* e = std::getenv("MGB_LARGE_ATOMS");
  const long long limit = e ? std::atoll(e) : 2048;
  return (long long)B * N >= limit;
}
// a few thousand pairs only: five threads per (pair, ell) (k_edge_pairs_*_cs)
inline bool edge_small(int B, int N) {
  const int m = edge_mode_override();
  return m >= 0 ? m == 2 : (long long)B * N * N * 5 < 148ll * 1024;
}

inline CovWs carve_workspace(const CovDesc& d, int B, void* base) {
  CovWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? (void*)((unsigned char*)base + off) : nullptr;
    off += (bytes + 255) & ~size_t(255);
    return p;
  };
  const size_t BN = (size_t)B * d.N, BNN = BN * d.N;
  const int C = d.C;
  int cmax = std::max(C, d.Cout);
  w.n_atoms = (int*)take(sizeof(int) * B);
  w.pair_off = (int*)take(sizeof(int) * (B + 1));
  w.pair_slot = (int*)take(sizeof(int) * BNN);
  w.atom_off = (int*)take(sizeof(int) * (B + 1));
  w.atom_list = (int*)take(sizeof(int) * BN);
  w.act_off = (int*)take(sizeof(int) * (B + 1));
  w.act_list = (int*)take(sizeof(int) * BN);
  {
    int tmax = 0;
    for (int k = 0; k < d.K; ++k) tmax = std::max(tmax, d.lv[k].totA);
    w.dcat = (float*)take(sizeof(float) * BN * tmax * 2);
  }
  w.Wt = (float*)take(sizeof(float) * d.n_wt);
  w.X = (float*)take(sizeof(float) * BN * d.S_in);
  w.A[0] = (float*)take(sizeof(float) * BN * C * 2);
  for (int k = 0; k < d.K; ++k) {
    w.A[k + 1] = (float*)take(sizeof(float) * BN * kM * d.lv[k].Cout * 2);
    w.E[k] = (float*)take(sizeof(float) * BNN * kNL * C * 2);
    w.cat[k] = (float*)take(sizeof(float) * BN * d.lv[k].totA * 2);
  }
  w.inv = (float*)take(sizeof(float) * BN * d.lat);
  w.lse = (float*)take(sizeof(float) * B * 2);
  w.pol_state = (float*)take(sizeof(float) * (size_t)B * policy_state_floats(d));
  w.hf = (float*)take(sizeof(float) * BN * d.Wd);
  w.flogit = (float*)take(sizeof(float) * BN);
  w.ht0 = (float*)take(sizeof(float) * BN * d.Wd);
  w.trans = (float*)take(sizeof(float) * BN * d.Wd);
  w.finv = (float*)take(sizeof(float) * B * d.lat);
  w.he = (float*)take(sizeof(float) * B * d.Wd);
  w.einv = (float*)take(sizeof(float) * B * d.latE);
  w.hd = (float*)take(sizeof(float) * B * d.Wd);
  w.vf = (float*)take(sizeof(float) * B * d.Wd);
  w.hv = (float*)take(sizeof(float) * B * d.Wd);
  w.dhe = (float*)take(sizeof(float) * B * d.Wd);
  w.dye = (float*)take(sizeof(float) * B * d.Z);
  w.dhd = (float*)take(sizeof(float) * B * d.Wd);
  w.dyd = (float*)take(sizeof(float) * B * 2 * d.G);
  w.dhv = (float*)take(sizeof(float) * B * d.Wd);
  w.dyv = (float*)take(sizeof(float) * B);
  w.dvf = (float*)take(sizeof(float) * B * d.Wd);
  w.dflogit = (float*)take(sizeof(float) * BN);
  w.dhf = (float*)take(sizeof(float) * BN * d.Wd);
  w.dht0 = (float*)take(sizeof(float) * BN * d.Wd);
  w.dtrans = (float*)take(sizeof(float) * BN * d.Wd);
  w.dinv = (float*)take(sizeof(float) * BN * d.lat);
  for (int q = 0; q < 2; ++q) w.dA[q] = (float*)take(sizeof(float) * BN * kM * cmax * 2);
  for (int q = 0; q < 2; ++q) w.dE[q] = (float*)take(sizeof(float) * BNN * kNL * C * 2);
  w.dD = (float*)take(sizeof(float) * BNN * kNL * C * 2 * (edge_bwd_split(B, d.N) ? kNL : 1));
  for (int k = 0; k < d.K; ++k) w.D[k] = (float*)take(sizeof(float) * BNN * kNL * C * 2);
  w.e_dpre = (float*)take(sizeof(float) * BNN * kNL * C * 2);
  w.e_R = (float*)take(sizeof(float) * BNN * kNL * C * 2);
  w.e_dR = (float*)take(sizeof(float) * BNN * kNL * C * 2);
  w.e_f = (float*)take(sizeof(float) * BNN * kRadFeat);
  w.loss_acc = (double*)take(sizeof(double) * 16);
  w.mix_stage = (float*)take(sizeof(float) * 2 * d.totWM);
  w.dw_probs = (DwProblem*)take(sizeof(DwProblem) * 32);
  w.dw_work = (DwWork*)take(sizeof(DwWork) * 96);
  w.bytes = off;
  return w;
}

inline int pick_co(int cout) {
  for (int co : {10, 8, 6, 5, 4}) if (cout % co == 0) return co;
  return 4;
}
// smallest instantiated register tile that holds all output channels of the row-parallel mix kernels
inline int pick_co_rows(int cout) {
  for (int co : {4, 8, 10, 12, 16, 20, 32}) if (cout <= co) return co;
  return 32;
}

}  // namespace mgb
