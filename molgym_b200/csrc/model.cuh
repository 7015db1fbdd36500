// model.cuh — descriptors shared by host and device code + the host-side plan builder.
//
// Reference structure being described (paths relative to the reference tree):
//   molgym/agents/covariant/agent.py:59-143   (module inventory of CovariantAC)
//   molgym/agents/covariant/modules.py:11-135 (Cormorant stack configuration), :138-190 (CormorantMixer)
// and the cormorant package semantics restated in oracle/thirdparty/cormorant (CG product path ordering,
// cat orderings, radial basis).
#pragma once
#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"

namespace mgb {

constexpr int kMaxLevels = MGB_MAX_LEVELS;
constexpr int kMaxMixUnits = 32;

// A unit of the channel-mixing contraction: rows (l, m0 .. m0+nm-1) handled by one warp pass.
struct MixUnit {
  int l, m0, nm;
};

struct MlpDesc {
  int in, hidden, out;
  long long W0, b0, W1, b1;      // float offsets into the flat parameter buffer (reference layout [out,in])
  long long W0t, W1t;            // float offsets into the transposed-weights scratch ([in,out])
};

struct CovDesc {
  int N, Z, K;                   // canvas size, species, number of CG levels
  int C, CPE, Cout;              // hidden channels, channels per element, last-level channels (Z*CPE)
  int G, Wd;                     // gaussians, MLP width
  int lat, latE;                 // invariant feature sizes (48Z and 48 for the defaults)
  int S_in;                      // input scalar features per atom: Z*3 + Z
  int zs[MGB_MAX_SPECIES];
  float charge_scale, bag_scale;
  float cut_rad, cut_width;      // soft cutoff (agent.py:66-69)
  float dmin, dmax;
  int has_beta;
  float beta;
  long long p_inW, p_inb;        // InputLinear
  LevelDesc lv[kMaxLevels];
  // mixer (CormorantMixer): cat = [ag(CPE) | sq blocks | in(CPE)]
  int catM[kNL], offM[kNL], totM, offWM[kNL], totWM, inM_block[kNL];
  long long p_mixW;
  CgTable mix_sq;
  MlpDesc focus, element, dist, trans, value;
  long long p_logstd;
  long long n_params;
  long long n_wt;                // floats of transposed-weight scratch
  long long wt_edge[kMaxLevels]; // float offset of transposed edge weights [l][k][c'][2] in the scratch
  long long wt_atom[kMaxLevels]; // float offset of transposed atom-mix weights [l][k][c'][2] in the scratch
  int n_grid;                    // Lebedev points
  const float* leb_y;            // [25][n_grid][2]  Y_lm(x_g), 'qm' norm, no conjugation
  const float* leb_logw;         // [n_grid]
  int n_units_hidden, n_units_out;
  MixUnit units_hidden[kMaxMixUnits], units_out[kMaxMixUnits];
};

// ------------------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------------------
inline double factorial_d(int n) {
  double r = 1.0;
  for (int i = 2; i <= n; ++i) r *= i;
  return r;
}

// <j1 m1 j2 m2 | j m>, Racah's formula (same convention as oracle/thirdparty/cormorant/cg_lib.py::clebsch).
inline double clebsch_gordan(int j1, int m1, int j2, int m2, int j, int m) {
  if (m1 + m2 != m || j < std::abs(j1 - j2) || j > j1 + j2) return 0.0;
  if (std::abs(m1) > j1 || std::abs(m2) > j2 || std::abs(m) > j) return 0.0;
  auto f = factorial_d;
  double pref = (2 * j + 1) * f(j + j1 - j2) * f(j - j1 + j2) * f(j1 + j2 - j) / f(j1 + j2 + j + 1);
  pref *= f(j + m) * f(j - m) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2);
  double tot = 0.0;
  for (int k = 0; k <= j1 + j2 - j; ++k) {
    int a[6] = {k, j1 + j2 - j - k, j1 - m1 - k, j2 + m2 - k, j - j2 + m1 + k, j - j1 - m2 + k};
    bool ok = true;
    double den = 1.0;
    for (int x : a) {
      if (x < 0) { ok = false; break; }
      den *= f(x);
    }
    if (!ok) continue;
    tot += ((k & 1) ? -1.0 : 1.0) / den;
  }
  return tot * std::sqrt(pref);
}

struct HostCgTable {
  int n_out = 0, n_pair = 0, nlm2 = 0;
  int n_blocks[kNL] = {0, 0, 0, 0, 0};  // number of channel blocks (paths) ending in each l
  std::vector<int> out_l, out_m, out_block, term_start, term_lm1, term_lm2;
  std::vector<float> term_coef;
  std::vector<int> pair_start, pair_out;
  std::vector<float> pair_coef;
  std::vector<int> out_dst, term_src, pair_ent;   // resolved (term_src / pair_ent hold 2 ints per term)
  std::vector<int> flat, slot_start, pad_pair, pad_sym;
  int n_slots = 0;
};

// Resolve a table against one use site: cat_l = [...blocks of C channels...] with per-l size catA[l], per-atom offset
// offA[l] (complex units), first block of this product block0[l].
inline void resolve_cg_table(HostCgTable& t, const int* catA, const int* offA, const int* block0, int C, bool square) {
  t.out_dst.resize(t.n_out);
  for (int o = 0; o < t.n_out; ++o) {
    const int l = t.out_l[o];
    t.out_dst[o] = offA[l] + t.out_m[o] * catA[l] + (block0[l] + t.out_block[o]) * C;
  }
  const size_t nt = t.term_lm1.size();
  t.term_src.resize(2 * nt);
  for (size_t q = 0; q < nt; ++q) {
    t.term_src[2 * q] = square ? t.term_lm1[q] * C : (t.term_lm1[q] * t.nlm2 + t.term_lm2[q]) * C;
    t.term_src[2 * q + 1] = square ? t.term_lm2[q] * C : 0;
  }
  t.pair_ent.resize(2 * nt);
  for (size_t q = 0; q < nt; ++q) {
    t.pair_ent[2 * q] = t.out_dst[t.pair_out[q]];
    int bits;
    std::memcpy(&bits, &t.pair_coef[q], 4);
    t.pair_ent[2 * q + 1] = bits;
  }
}

// Flat / padded forms.  n_slots = number of (thread / C) lanes that walk the output-major table.
inline bool finalize_cg_table(HostCgTable& t, int n_slots, bool square) {
  const size_t nt = t.term_lm1.size();
  t.flat.resize(4 * nt);
  for (int o = 0; o < t.n_out; ++o)
    for (int q = t.term_start[o]; q < t.term_start[o + 1]; ++q) {
      int bits;
      std::memcpy(&bits, &t.term_coef[q], 4);
      t.flat[4 * q + 0] = t.term_src[2 * q];
      t.flat[4 * q + 1] = t.term_src[2 * q + 1];
      t.flat[4 * q + 2] = (t.out_dst[o] << 1) | (q == t.term_start[o + 1] - 1 ? 1 : 0);
      t.flat[4 * q + 3] = bits;
    }
  t.n_slots = n_slots;
  t.slot_start.assign(n_slots + 1, (int)nt);
  t.slot_start[0] = 0;
  {
    int o = 0;
    for (int s = 1; s < n_slots; ++s) {
      const long long target = (long long)nt * s / n_slots;
      while (o < t.n_out && t.term_start[o] < target) ++o;
      t.slot_start[s] = o < t.n_out ? t.term_start[o] : (int)nt;
    }
  }
  t.pad_pair.assign((size_t)t.n_pair * kCgPad * 2, 0);
  for (int p = 0; p < t.n_pair; ++p) {
    const int n = t.pair_start[p + 1] - t.pair_start[p];
    if (n > kCgPad) return false;
    for (int q = 0; q < n; ++q) {
      t.pad_pair[((size_t)p * kCgPad + q) * 2 + 0] = t.pair_ent[2 * (t.pair_start[p] + q)];
      t.pad_pair[((size_t)p * kCgPad + q) * 2 + 1] = t.pair_ent[2 * (t.pair_start[p] + q) + 1];
    }
  }
  if (square) {
    const int n = t.nlm2;   // square: nlm1 == nlm2
    t.pad_sym.assign((size_t)t.n_pair * 2 * kCgPad * 2, 0);
    for (int x = 0; x < n; ++x)
      for (int y = 0; y < n; ++y)
        for (int h = 0; h < 2; ++h) {
          const int p = h == 0 ? x * n + y : y * n + x;
          for (int q = 0; q < t.pair_start[p + 1] - t.pair_start[p]; ++q) {
            const size_t dst = (((size_t)(x * n + y) * 2 + h) * kCgPad + q) * 2;
            t.pad_sym[dst + 0] = t.pair_ent[2 * (t.pair_start[p] + q)];
            t.pad_sym[dst + 1] = t.pair_ent[2 * (t.pair_start[p] + q) + 1];
          }
        }
  }
  return true;
}

// CG product table for rep1 with ells 0..n1-1 and rep2 with ells 0..n2-1, truncated at kL, paths enumerated
// l1 outer / l2 inner and outputs of one l concatenated in that order (cg_lib.py::cg_product).
inline HostCgTable build_cg_table(int n1, int n2) {
  HostCgTable t;
  const int nlm1 = n1 * n1, nlm2 = n2 * n2;
  t.nlm2 = nlm2;
  t.n_pair = nlm1 * nlm2;
  std::vector<std::vector<std::pair<int, float>>> by_pair(t.n_pair);
  t.term_start.push_back(0);
  for (int l1 = 0; l1 < n1; ++l1)
    for (int l2 = 0; l2 < n2; ++l2)
      for (int l = std::abs(l1 - l2); l <= std::min(l1 + l2, kL); ++l) {
        const int block = t.n_blocks[l]++;
        for (int m = -l; m <= l; ++m) {
          const int o = t.n_out++;
          t.out_l.push_back(l);
          t.out_m.push_back(m + l);
          t.out_block.push_back(block);
          for (int m1 = -l1; m1 <= l1; ++m1) {
            const int m2 = m - m1;
            if (std::abs(m2) > l2) continue;
            const double cg = clebsch_gordan(l1, m1, l2, m2, l, m);
            if (cg == 0.0) continue;
            const int a = lm_index(l1, m1), b = lm_index(l2, m2);
            t.term_lm1.push_back(a);
            t.term_lm2.push_back(b);
            t.term_coef.push_back((float)cg);
            by_pair[a * nlm2 + b].push_back({o, (float)cg});
          }
          t.term_start.push_back((int)t.term_lm1.size());
        }
      }
  t.pair_start.push_back(0);
  for (auto& v : by_pair) {
    for (auto& pr : v) {
      t.pair_out.push_back(pr.first);
      t.pair_coef.push_back(pr.second);
    }
    t.pair_start.push_back((int)t.pair_out.size());
  }
  return t;
}

// Forward gather tables (see GatherTables in common.cuh).
struct HostGather {
  std::vector<int> ag_flat8, sq_flat8;   // 2 ints per term
  std::vector<int> ag_slot, sq_slot;     // [kGatherSlots + 1]
};
inline bool build_gather_tables(const HostCgTable& ag, const HostCgTable& sq, int C, HostGather& out) {
  auto flat8 = [&](const HostCgTable& t, bool square, std::vector<int>& flat, std::vector<int>& slots) {
    const int nt = (int)t.term_lm1.size();
    for (int o = 0; o < t.n_out; ++o)
      for (int q = t.term_start[o]; q < t.term_start[o + 1]; ++q) {
        const int last = q == t.term_start[o + 1] - 1 ? 1 : 0, dst = t.out_dst[o];
        int w0;
        if (square) w0 = (t.term_lm1[q] * C) | ((t.term_lm2[q] * C) << 8) | (last << 16) | (dst << 17);
        else w0 = ((t.term_lm1[q] * t.nlm2 + t.term_lm2[q]) * C) | (last << 13) | (dst << 14);
        int bits;
        std::memcpy(&bits, &t.term_coef[q], 4);
        flat.push_back(w0);
        flat.push_back(bits);
      }
    slots.assign(kGatherSlots + 1, nt);
    slots[0] = 0;
    int o = 0;
    for (int s = 1; s < kGatherSlots; ++s) {
      const long long target = (long long)nt * s / kGatherSlots;
      while (o < t.n_out && t.term_start[o] < target) ++o;
      slots[s] = o < t.n_out ? t.term_start[o] : nt;
    }
  };
  if (kM * kM * C >= (1 << 13) || kM * C >= (1 << 8)) return false;
  for (int o = 0; o < ag.n_out; ++o) if (ag.out_dst[o] >= (1 << 13)) return false;
  for (int o = 0; o < sq.n_out; ++o) if (sq.out_dst[o] >= (1 << 13)) return false;
  flat8(ag, false, out.ag_flat8, out.ag_slot);
  flat8(sq, true, out.sq_flat8, out.sq_slot);
  return true;
}

// Greedy split of the 25 (l, m) rows into warp units of <= 3 consecutive m of the same l.
inline int build_mix_units(MixUnit* units, int max_nm) {
  int n = 0;
  for (int l = kL; l >= 0; --l) {  // heavy rows first
    int m = 0;
    while (m < 2 * l + 1) {
      int nm = std::min(max_nm, 2 * l + 1 - m);
      units[n++] = MixUnit{l, m, nm};
      m += nm;
    }
  }
  return n;
}

}  // namespace mgb
