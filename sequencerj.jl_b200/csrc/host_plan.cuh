// Static schedule of one sequence() call (pure host logic, no CUDA): which chunk matrices are produced in which wave,
// in which slot of the device matrix pool, by which path (tile kernels or derivation from the next finer scale), and
// which accumulator every (metric, scale) group sums into.  Replaces the order of the reference's triple loop
// (src/sequencer.jl:194-258: metric, scale, chunk) by one that keeps the GPU's HBM full and the derivation sources
// resident; the results do not depend on the order (every chunk matrix, eta and group sum is the same).
//
// A *wave* is a set of chunk matrices that are resident at once: they are produced, measured by ONE batched MST + tree
// launch sequence, combined into their groups' accumulators, and their slots are reused by the next wave.
//
//   flat waves    groups computed by the tile kernels (EMD; every group when the scales do not nest; chunk-major
//                 partial calls), packed scale-major as round 1 did.
//   chain waves   a contraction metric (L2 / KL / Energy) over scales whose chunk boundaries nest: only the finest
//                 scale runs tiles, each coarser chunk is derived from the `ratio` chunks below it
//                 (kernels_derive.cuh).  When all of it fits one wave, that is the whole story (BASELINE config 3).
//                 Otherwise (config 4: 63 matrices of 7.2 GB per metric) the rows are cut into BLOCKS = chunks of a
//                 block-level scale B: one wave per block holds the block's chunks of every scale up to B, derived
//                 depth-first inside the wave; the block-level matrices stay behind in CARRY slots and the scales
//                 coarser than B are derived from them in a final top wave.  Live matrices: one block + carry.
#pragma once

namespace {

struct PlanSeg {
  int g = 0, k = 0, l = 0;        // group (metric-major index k * nscales + l), metric index, scale index
  int c0 = 0, c1 = 0;             // chunk range of this segment
  int slot0 = 0;                  // first slot in the matrix pool (chunk c at slot0 + c - c0)
  int first = 0, last = 0;        // first / last segment of the group in this call
  int out0 = 0;                   // index of chunk c0 in the task-ordered output arrays
  int src_slot0 = -1;             // derived segment: pool slot of fine chunk `fbase` of scale `src_l`; -1 = tile kernels
  int src_l = -1, fbase = 0;
  int sslot = -1;                 // accumulator slot of the group in the pool (-1: caller's buffer, partial mode)
  int gp = -1;                    // group index in finishing order
};

struct PlanWave {
  std::vector<PlanSeg> segs;      // in execution order (finer scales first inside a chain wave)
  int ntasks = 0, out0 = 0;
  int gp_done0 = 0, gp_done1 = 0; // groups [gp_done0, gp_done1) (finishing order) are complete after this wave
};

struct Plan {
  std::vector<PlanWave> waves;
  int pool_slots = 0;             // matrices in the pool: accumulators + wave + carry, maximum over the phases
  int max_tasks = 0, max_done = 0;
  int total_chunks = 0, ngp = 0, n_derived_groups = 0;
  std::vector<int> gp_of, g_of_gp, chunk_off;      // per group g: finishing index (-1 = not computed), first output position
  std::vector<int> sslot_of_gp;
  std::vector<int> task_slot, task_pos;            // per task (wave order): pool slot, position chunk_off[g] + c - lo
  std::vector<uint8_t> derived;                    // per group: 1 = at least one segment is derived
  std::string error;
};

struct PlanInput {
  int M = 0, nmetrics = 0, nscales = 0;
  const int32_t* metric_ids = nullptr;
  const std::vector<ScaleLayout>* layouts = nullptr;
  std::function<bool(int, int)> mask;              // (k, l) -> group is computed by this call
  std::function<int(int, int)> c_lo, c_hi;         // (k, l) -> chunk range (whole group unless partial)
  bool partial = false, hier = true;
  int64_t mats = 0;                                // N x N matrices that fit the memory budget
  int wave_target = 0;                             // flat waves close at scale boundaries beyond this many tasks (0 = never)
  int max_fine = 32;                               // kDeriveMaxFine
};

inline bool plan_metric_derivable(int mid) { return mid == SEQJ_L2 || mid == SEQJ_KLD || mid == SEQJ_ENERGY; }

// Flat waves over the groups selected by `sel`, scale-major (one prep per scale and wave), chunks innermost.  Pool
// layout of the phase: accumulators [0, nS), wave slots [nS, nS + W).
inline void plan_flat(const PlanInput& in, const std::vector<int>& scale_order, const std::function<bool(int, int)>& sel,
                      int nS, int W, Plan& P) {
  const auto& Ls = *in.layouts;
  PlanWave w;
  int open_groups = 0;
  std::vector<int> open_slot;                      // accumulator slots in use: index = slot, value = group or -1
  open_slot.assign(std::max(1, nS), -1);
  std::vector<int> slot_of_g(in.nmetrics * in.nscales, -1);
  auto close = [&]() {
    if (w.ntasks) P.waves.push_back(w);
    // groups that finished in this wave release their accumulators
    for (auto& sg : w.segs)
      if (sg.last && slot_of_g[sg.g] >= 0) { open_slot[slot_of_g[sg.g]] = -1; slot_of_g[sg.g] = -1; }
    w = PlanWave();
    open_groups = 0;
    for (int s : open_slot) if (s >= 0) open_groups++;
  };
  for (int l : scale_order) {
    if (in.wave_target > 0 && w.ntasks >= in.wave_target) close();
    for (int k = 0; k < in.nmetrics; ++k) {
      if (!in.mask(k, l) || !sel(k, l)) continue;
      const int g = k * in.nscales + l, lo = in.c_lo(k, l), nch = in.c_hi(k, l);
      int c0 = lo;
      while (c0 < nch) {
        if (w.ntasks >= W) close();
        if (slot_of_g[g] < 0 && !in.partial) {     // the group needs an accumulator from this wave on
          int s = -1;
          for (int q = 0; q < (int)open_slot.size(); ++q) if (open_slot[q] < 0) { s = q; break; }
          if (s < 0) { close(); for (int q = 0; q < (int)open_slot.size(); ++q) if (open_slot[q] < 0) { s = q; break; } }
          if (s < 0) { P.error = "internal: no free accumulator slot"; return; }
          open_slot[s] = g; slot_of_g[g] = s;
        }
        const int take = std::min(nch - c0, W - w.ntasks);
        PlanSeg sg;
        sg.g = g; sg.k = k; sg.l = l; sg.c0 = c0; sg.c1 = c0 + take; sg.slot0 = nS + w.ntasks;
        sg.first = (c0 == lo); sg.last = (c0 + take == nch);
        sg.sslot = in.partial ? -1 : slot_of_g[g];
        w.segs.push_back(sg);
        w.ntasks += take; c0 += take;
        (void)Ls;
      }
    }
  }
  close();
  P.pool_slots = std::max(P.pool_slots, (in.partial ? 0 : nS) + W);
}

// Chain waves of metric k over the nested scales `chain` (finest first).  Pool layout of the phase: accumulators
// [0, T+1), wave slots after them, carry slots after those.
inline bool plan_chain(const PlanInput& in, int k, const std::vector<int>& chain, Plan& P) {
  const auto& Ls = *in.layouts;
  const int T = (int)chain.size() - 1;
  const int nacc = T + 1;
  const int64_t room = in.mats - nacc;
  auto nch = [&](int t) { return Ls[chain[t]].nchunks; };
  auto cl = [&](int t) { return Ls[chain[t]].cl; };
  // block level: the coarsest B whose block wave + carry, and top wave + carry, fit
  int B = -1;
  for (int b = T; b >= 1; --b) {
    int64_t blockmats = 0;
    for (int t = 0; t <= b; ++t) blockmats += cl(b) / cl(t);
    const int64_t carry = (b < T) ? nch(b) : 0;
    int64_t top = 0;
    for (int t = b + 1; t <= T; ++t) top += nch(t);
    if (blockmats + carry <= room && top + carry <= room) { B = b; break; }
  }
  if (B < 0) return false;
  const int carry = (B < T) ? nch(B) : 0;
  int64_t blockmats = 0;
  for (int t = 0; t <= B; ++t) blockmats += cl(B) / cl(t);
  int64_t top = 0;
  for (int t = B + 1; t <= T; ++t) top += nch(t);
  const int wave0 = nacc;                                        // first wave slot
  const int wave_slots = (int)std::max<int64_t>(blockmats - (carry ? 1 : 0), top);
  const int carry0 = wave0 + wave_slots;                         // block-level chunks live here when a top wave follows
  P.pool_slots = std::max(P.pool_slots, carry0 + carry);
  const int nblocks = nch(B);
  for (int b = 0; b < nblocks; ++b) {
    PlanWave w;
    int next = wave0, prev_slot0 = -1, prev_c0 = 0;
    for (int t = 0; t <= B; ++t) {
      const int R = cl(B) / cl(t);
      PlanSeg sg;
      sg.k = k; sg.l = chain[t]; sg.g = k * in.nscales + sg.l;
      sg.c0 = b * R; sg.c1 = std::min((b + 1) * R, nch(t));
      if (t == B && carry) sg.slot0 = carry0 + b;
      else { sg.slot0 = next; next += sg.c1 - sg.c0; }
      sg.first = (b == 0); sg.last = (b == nblocks - 1);
      sg.sslot = t;
      if (t > 0) { sg.src_slot0 = prev_slot0; sg.src_l = chain[t - 1]; sg.fbase = prev_c0; }
      prev_slot0 = sg.slot0; prev_c0 = sg.c0;
      w.segs.push_back(sg);
      w.ntasks += sg.c1 - sg.c0;
    }
    P.waves.push_back(w);
  }
  if (B < T) {
    PlanWave w;
    int next = wave0, prev_slot0 = carry0, prev_c0 = 0;
    for (int t = B + 1; t <= T; ++t) {
      PlanSeg sg;
      sg.k = k; sg.l = chain[t]; sg.g = k * in.nscales + sg.l;
      sg.c0 = 0; sg.c1 = nch(t); sg.slot0 = next; next += nch(t);
      sg.first = 1; sg.last = 1; sg.sslot = t;
      sg.src_slot0 = prev_slot0; sg.src_l = chain[t - 1]; sg.fbase = prev_c0;
      prev_slot0 = sg.slot0; prev_c0 = 0;
      w.segs.push_back(sg);
      w.ntasks += nch(t);
    }
    P.waves.push_back(w);
  }
  return true;
}

inline Plan make_plan(const PlanInput& in) {
  Plan P;
  const auto& Ls = *in.layouts;
  const int G = in.nmetrics * in.nscales;
  P.gp_of.assign(G, -1); P.chunk_off.assign(G, 0); P.derived.assign(G, 0);
  // scales with many chunks (short chunks) first
  std::vector<int> scale_order(in.nscales);
  for (int l = 0; l < in.nscales; ++l) scale_order[l] = l;
  std::stable_sort(scale_order.begin(), scale_order.end(), [&](int a, int b) { return Ls[a].nchunks > Ls[b].nchunks; });
  int ngroups = 0;
  for (int g = 0; g < G; ++g) {
    const int k = g / in.nscales, l = g % in.nscales;
    P.chunk_off[g] = P.total_chunks;
    if (in.mask(k, l)) { P.total_chunks += in.c_hi(k, l) - in.c_lo(k, l); ngroups++; }
  }
  if (P.total_chunks == 0) return P;
  // chains: per derivable metric, maximal runs of consecutive (in scale_order) computed scales that nest
  std::vector<std::vector<std::vector<int>>> chains(in.nmetrics);
  std::vector<uint8_t> in_chain(G, 0);
  if (in.hier && !in.partial && in.nscales > 1)
    for (int k = 0; k < in.nmetrics; ++k) {
      if (!plan_metric_derivable(in.metric_ids[k])) continue;
      std::vector<int> cur;
      auto flush = [&]() { if (cur.size() >= 2) chains[k].push_back(cur); cur.clear(); };
      int l_prev = -1;                                            // previous scale of the order that any metric computes
      for (int l : scale_order) {
        bool any = false;
        for (int kk = 0; kk < in.nmetrics; ++kk) any |= in.mask(kk, l);
        if (!any) continue;
        bool link = false;
        if (l_prev >= 0 && in.mask(k, l) && in.mask(k, l_prev) && !cur.empty() && cur.back() == l_prev) {
          const int ratio = Ls[l].cl / Ls[l_prev].cl;
          link = (Ls[l].cl % Ls[l_prev].cl == 0) && ratio <= in.max_fine && Ls[l].nchunks < Ls[l_prev].nchunks;
        }
        if (!link) { flush(); if (in.mask(k, l)) cur.push_back(l); }
        else cur.push_back(l);
        l_prev = l;
      }
      flush();
    }
  // does everything fit one wave?  Then one flat-ordered wave with in-wave derivation (the config-3 schedule).
  const int nS_all = in.partial ? 0 : (int)std::min<int64_t>(std::max(1, ngroups), std::max<int64_t>(1, in.mats / 6));
  const bool single = (int64_t)P.total_chunks + nS_all <= in.mats && (in.partial || nS_all >= ngroups) &&
                      (in.wave_target <= 0 || in.wave_target >= P.total_chunks);
  if (single) {
    plan_flat(in, scale_order, [](int, int) { return true; }, nS_all, P.total_chunks, P);
    if (!P.error.empty()) return P;
    // mark the derived segments: source = the same metric's segment of the previous chain scale, same (only) wave
    if (!P.waves.empty()) {
      PlanWave& w = P.waves[0];
      for (int k = 0; k < in.nmetrics; ++k)
        for (auto& ch : chains[k])
          for (size_t t = 1; t < ch.size(); ++t) {
            PlanSeg *dst = nullptr, *src = nullptr;
            for (auto& sg : w.segs) {
              if (sg.k == k && sg.l == ch[t]) dst = &sg;
              if (sg.k == k && sg.l == ch[t - 1]) src = &sg;
            }
            if (dst && src) { dst->src_slot0 = src->slot0; dst->src_l = src->l; dst->fbase = 0; }
          }
    }
  } else {
    // chains that fit the block schedule leave the flat phase
    for (int k = 0; k < in.nmetrics; ++k)
      for (auto& ch : chains[k]) {
        Plan probe;                                               // dry run: does a block level exist?
        if ((int64_t)ch.size() + 3 <= in.mats && plan_chain(in, k, ch, probe))
          for (int l : ch) in_chain[k * in.nscales + l] = 1;
      }
    int nflat = 0, flat_chunks = 0;
    for (int g = 0; g < G; ++g) {
      const int k = g / in.nscales, l = g % in.nscales;
      if (in.mask(k, l) && !in_chain[g]) { nflat++; flat_chunks += in.c_hi(k, l) - in.c_lo(k, l); }
    }
    if (nflat) {
      const int nS = in.partial ? 0 : (int)std::min<int64_t>(nflat, std::max<int64_t>(1, in.mats / 6));
      const int64_t W = std::min<int64_t>(in.mats - nS, flat_chunks);
      if (W < 1) { P.error = "not enough device memory for two N x N chunk matrices plus the accumulator"; return P; }
      plan_flat(in, scale_order, [&](int k, int l) { return !in_chain[k * in.nscales + l]; }, nS, (int)W, P);
      if (!P.error.empty()) return P;
    }
    for (int k = 0; k < in.nmetrics; ++k)
      for (auto& ch : chains[k])
        if (in_chain[k * in.nscales + ch[0]]) plan_chain(in, k, ch, P);
  }
  // finishing order of the groups, output positions, task tables
  int out = 0;
  for (auto& w : P.waves) {
    w.out0 = out;
    w.gp_done0 = P.ngp;
    for (auto& sg : w.segs) {
      sg.out0 = out;
      out += sg.c1 - sg.c0;
      if (sg.src_slot0 >= 0) P.derived[sg.g] = 1;
      for (int c = sg.c0; c < sg.c1; ++c) {
        P.task_slot.push_back(sg.slot0 + c - sg.c0);
        P.task_pos.push_back(P.chunk_off[sg.g] + c - in.c_lo(sg.k, sg.l));
      }
    }
    if (!in.partial)
      for (auto& sg : w.segs)
        if (sg.last) { P.gp_of[sg.g] = P.ngp++; P.g_of_gp.push_back(sg.g); P.sslot_of_gp.push_back(sg.sslot); }
    w.gp_done1 = P.ngp;
    P.max_tasks = std::max(P.max_tasks, w.ntasks);
    P.max_done = std::max(P.max_done, w.gp_done1 - w.gp_done0);
  }
  if (in.partial)                                                 // no group finishes here: index the groups in g order
    for (int g = 0; g < G; ++g)
      if (in.mask(g / in.nscales, g % in.nscales)) { P.gp_of[g] = P.ngp++; P.g_of_gp.push_back(g); P.sslot_of_gp.push_back(-1); }
  for (auto& w : P.waves)
    for (auto& sg : w.segs) sg.gp = P.gp_of[sg.g];
  for (int g = 0; g < G; ++g) P.n_derived_groups += P.derived[g];
  return P;
}

}  // namespace
