// Group-major sharding plan: which rank / device owns which (metric, scale) group.
// Pure host logic, shared by the single-process multi-GPU path (seqj_sequence on a handle with several devices) and,
// through seqj_plan_shards, by the one-process-per-GPU driver (sharding.py).
//
// Units of work.  An EMD-type group (EMD, explicit-grid EMD, PeriodicEuclidean: direct tile kernels at every scale)
// is a unit of its own.  A contraction metric (L2 / KL / Energy) over scales whose chunks nest is a CHAIN: only its
// finest scale runs DMMA tiles, every coarser scale is derived on the same GPU (kernels_derive.cuh).  A chain can be
// cut into consecutive sub-chains -- each pays one tile launch for its own finest scale -- when that balances the
// ranks better: the greedy below adds the cut that lowers the longest rank until none does.  Units go to ranks by
// longest-processing-time-first.
//
// Cost model, in picoseconds per N^2 matrix entry (every term scales with N^2, so N drops out), from the measured
// kernel rates of profiles/r02_*: EMD tiles 16.2 Tinstr/s x cl / (cl + 18), DMMA tiles 37.06 TFLOP/s x min(0.9, cl / (cl + 110)) for
// chunk length cl, derivation and combine 4.5 / 5.9 TB/s, MST 5.5 TB/s (8 bytes per entry and graph).
#pragma once

namespace {

struct ShardUnit {
  double cost = 0.0;
  std::vector<int> groups;        // metric-major group indices k * nscales + l
};

inline double shard_cost_mst_combine(int nchunks) {
  const double mst = 8.0 / 5.5 * (nchunks + 1);                 // chunk graphs + the group graph
  const double comb = (4.0 * nchunks + 8.0) / 5.9;              // upper halves of the chunk matrices read, group matrix written
  return mst + comb;
}
inline double shard_cost_tiles(int mid, int64_t M, int cl) {
  const bool direct = mid == SEQJ_EMD || mid == SEQJ_EMD_GRID || mid == SEQJ_PERIODIC || mid >= SEQJ_CITYBLOCK;
  if (direct) return (double)M / (16.2 * (double)cl / ((double)cl + 18.0));   // per-tile fixed cost: 154 ms instead of 122 at cl = 64, N = 30,000
  const double eff = std::min(0.9, (double)cl / ((double)cl + 110.0));
  return (double)M / (37.06 * eff);
}
inline double shard_cost_derive(int nfine, int ncoarse) { return 8.0 * (0.5 * nfine + ncoarse) / 4.5; }

// owner[g] for every group; load[r] (optional) = modelled cost per rank in ps per N^2 entry
inline void plan_shards(int64_t M, const int32_t* metric_ids, int nmetrics, const int32_t* scales, int nscales, int world,
                        std::vector<int>& owner, std::vector<double>* load_out) {
  const int G = nmetrics * nscales;
  owner.assign(G, 0);
  std::vector<int> cl(nscales), nch(nscales);
  for (int l = 0; l < nscales; ++l) {
    const int64_t sc = std::max<int64_t>(1, scales[l]);
    cl[l] = (int)((M + sc - 1) / sc);
    nch[l] = (int)((M + cl[l] - 1) / cl[l]);
  }
  // scales of a chain: distinct chunk lengths, finest first, each a multiple of the previous one
  std::vector<int> order(nscales);
  for (int l = 0; l < nscales; ++l) order[l] = l;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return nch[a] > nch[b]; });
  auto linked = [&](int lf, int lc) {                           // lc derivable from lf
    return cl[lc] % cl[lf] == 0 && cl[lc] / cl[lf] <= 32 && nch[lc] < nch[lf];
  };
  // per contraction metric: maximal chains (vectors of scale indices) + the cut flags between consecutive members
  struct Chain { int k; std::vector<int> sc; std::vector<uint8_t> cut; };
  std::vector<Chain> chains;
  std::vector<ShardUnit> fixed;
  for (int k = 0; k < nmetrics; ++k) {
    const int mid = metric_ids[k];
    const bool contraction = mid == SEQJ_L2 || mid == SEQJ_KLD || mid == SEQJ_ENERGY;
    if (!contraction || nscales < 2) {
      for (int l = 0; l < nscales; ++l)
        fixed.push_back({shard_cost_tiles(mid, M, cl[l]) + shard_cost_mst_combine(nch[l]), {k * nscales + l}});
      continue;
    }
    Chain cur{k, {}, {}};
    for (int l : order) {
      if (!cur.sc.empty() && !linked(cur.sc.back(), l)) {
        chains.push_back(cur);
        cur = Chain{k, {}, {}};
      }
      cur.sc.push_back(l);
    }
    if (!cur.sc.empty()) chains.push_back(cur);
  }
  for (auto& c : chains) c.cut.assign(c.sc.size(), 0);           // cut[t] = 1: scale t starts a new unit (t >= 1)
  auto build_units = [&](std::vector<ShardUnit>& units) {
    units = fixed;
    for (auto& c : chains) {
      ShardUnit u;
      for (size_t t = 0; t < c.sc.size(); ++t) {
        const int l = c.sc[t];
        if (t == 0 || c.cut[t]) {
          if (!u.groups.empty()) units.push_back(u);
          u = ShardUnit();
          u.cost += shard_cost_tiles(metric_ids[c.k], M, cl[l]);
        } else {
          u.cost += shard_cost_derive(nch[c.sc[t - 1]], nch[l]);
        }
        u.cost += shard_cost_mst_combine(nch[l]);
        u.groups.push_back(c.k * nscales + l);
      }
      if (!u.groups.empty()) units.push_back(u);
    }
  };
  // LPT assignment; returns the rank loads sorted in descending order (compared lexicographically: a cut that leaves
  // the longest rank alone but shortens the next one is still progress -- the following cut can then help the longest)
  auto lpt = [&](std::vector<ShardUnit>& units, std::vector<int>* own, std::vector<double>* load) {
    std::stable_sort(units.begin(), units.end(), [](const ShardUnit& a, const ShardUnit& b) { return a.cost > b.cost; });
    std::vector<double> ld(world, 0.0);
    for (auto& u : units) {
      const int r = (int)(std::min_element(ld.begin(), ld.end()) - ld.begin());
      ld[r] += u.cost;
      if (own) for (int g : u.groups) (*own)[g] = r;
    }
    if (load) *load = ld;
    std::sort(ld.begin(), ld.end(), std::greater<double>());
    return ld;
  };
  auto better = [](const std::vector<double>& a, const std::vector<double>& b) {   // a strictly better than b
    for (size_t i = 0; i < a.size(); ++i) {
      if (a[i] < b[i] * (1.0 - 1e-3)) return true;
      if (a[i] > b[i] * (1.0 + 1e-3)) return false;
    }
    return false;
  };
  std::vector<ShardUnit> units;
  build_units(units);
  std::vector<double> best = lpt(units, nullptr, nullptr);
  while (world > 1) {                                            // greedy: the single cut that helps most, until none does
    std::vector<double> cand_best = best;
    Chain* cand_chain = nullptr;
    size_t cand_t = 0;
    for (auto& c : chains)
      for (size_t t = 1; t < c.sc.size(); ++t) {
        if (c.cut[t]) continue;
        c.cut[t] = 1;
        build_units(units);
        const std::vector<double> v = lpt(units, nullptr, nullptr);
        c.cut[t] = 0;
        if (better(v, cand_best)) { cand_best = v; cand_chain = &c; cand_t = t; }
      }
    if (!cand_chain) break;
    cand_chain->cut[cand_t] = 1;
    best = cand_best;
  }
  build_units(units);
  lpt(units, &owner, load_out);
}

}  // namespace
