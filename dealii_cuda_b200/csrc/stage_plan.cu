// stage_plan.cu -- builds the plan of the staged Laplace cell kernel from the index array (host code; see stage_plan.h).
#include <algorithm>
#include <atomic>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <unordered_map>
#include "common.cuh"
#include "stage_plan.h"

namespace mfg {

namespace {

constexpr uint32_t CBIT = 0x80000000u;
constexpr uint16_t RAW_NONE = 0xffffu;   // constrained entry / cell beyond the mesh / idle lane

struct GroupRaw
{
  // pattern part (identical for all groups of the same shape): entries of the load list are numbered own range first,
  // then the halo list; pos holds entry | STAGE_DEAD, or RAW_NONE
  std::vector<uint16_t> pos;     // [ns][32]
  std::vector<uint8_t>  oflag;   // [own_total] 0 skip, 1 plain store, 2 red.add
  uint32_t              own_total = 0, n_halo = 0, mm = 0;
  // group part
  uint32_t              own_base = 0;
  std::vector<uint32_t> halo;
  bool                  ok = false;
  std::string key() const
  {
    std::string k;
    auto put = [&](const void *p, size_t n) { k.append((const char *)p, n); };
    const uint32_t hdr[3] = {own_total, n_halo, mm};
    put(hdr, sizeof(hdr));
    put(pos.data(), pos.size() * 2);
    put(oflag.data(), oflag.size());
    return k;
  }
};

struct Lane { int c, i; };

struct Builder
{
  const StagePlanIn &in;
  const int          n, ns, npc, cw;
  std::vector<Lane>  lanes;      // [32]
  std::vector<uint8_t>  mult;    // unconstrained entries per DoF over all cells (saturating)
  std::vector<uint32_t> first;   // first cell that holds the DoF
  explicit Builder(const StagePlanIn &i) : in(i), n(i.n), ns(i.n * i.n), npc(i.n * i.n * i.n), cw(i.cw)
  {
    lanes.resize(32);
    const bool split = cw % 2 == 0;
    for (int l = 0; l < 32; ++l)
      {
        const int ch = split ? l / 16 : 0, l16 = split ? l % 16 : l;
        if (l16 >= (split ? in.hc : cw) * n) lanes[l] = Lane{-1, 0};
        else lanes[l] = Lane{(split ? in.hc * ch : 0) + l16 / n, l16 % n};
      }
  }

  void count()
  {
    mult.assign(in.n_dofs, 0);
    first.assign(in.n_dofs, 0xffffffffu);
    const size_t total = (size_t)in.n_cells * npc;
    for (size_t t = 0; t < total; ++t)
      {
        const uint32_t e = in.idx[t], d = e & ~CBIT;
        if (first[d] == 0xffffffffu) first[d] = (uint32_t)(t / npc);
        if (!(e & CBIT) && mult[d] != 255) ++mult[d];
      }
  }

  // face-merge mask: bit (10 d + c) = cell c hands the entries of its upper face in direction d to cell c + 2^d.
  // Order in the kernel: x, z, y.  A merge into entries that an EARLIER merge of the receiver hands on, while the
  // sender does not take part in that earlier merge, would strand the contribution and is dropped.
  uint32_t merge_mask(uint32_t g) const
  {
    if (cw > 10) return 0;
    uint32_t m = 0;
    for (int d = 0; d < 3; ++d)
      {
        if (!((in.merge_dirs >> d) & 1)) continue;
        const int step = 1 << d, sd = d == 0 ? 1 : d == 1 ? n : n * n;
        for (int c = 0; c + step < cw; ++c)
          {
            const uint32_t a = g * cw + c, b = a + step;
            if (b >= in.n_plain) continue;
            bool same = true;
            for (int q = 0; q < npc && same; ++q)
              if ((q / sd) % n == n - 1) same = in.idx[(size_t)a * npc + q] == in.idx[(size_t)b * npc + q - (n - 1) * sd];
            if (same) m |= 1u << (10 * d + c);
          }
      }
    auto bit = [&](int d, int c) { return c >= 0 && c < 10 && ((m >> (10 * d + c)) & 1u); };
    // z (second): receiver c+4 hands its i = n-1 entries on in x (first)
    for (int c = 0; c + 4 < cw; ++c)
      if (bit(2, c) && bit(0, c + 4) && !bit(0, c)) m &= ~(1u << (20 + c));
    // y (last): receiver c+2 hands on in x or z
    for (int c = 0; c + 2 < cw; ++c)
      if (bit(1, c) && ((bit(0, c + 2) && !bit(0, c)) || (bit(2, c + 2) && !bit(2, c)))) m &= ~(1u << (10 + c));
    return m;
  }

  // entries handed over (dead) after the merges x, z, y with zeroing of what was handed over; returns false when a
  // contribution would end in a dead entry
  bool simulate(uint32_t g, uint32_t mm, std::vector<uint8_t> &dead) const
  {
    const int ne = cw * npc;
    std::vector<uint16_t> cnt(ne, 1);  // contributions currently held by the entry
    dead.assign(ne, 0);
    auto E = [&](int c, int i, int j, int k) { return c * npc + i + n * j + n * n * k; };
    auto move = [&](int from, int to) {
      cnt[to] = (uint16_t)(cnt[to] + cnt[from]);
      cnt[from] = 0;
      dead[from] = 1;
    };
    const int order[3] = {0, 2, 1};
    for (int o = 0; o < 3; ++o)
      {
        const int d = order[o], step = 1 << d;
        for (int c = 0; c + step < cw && c < 10; ++c)
          {
            if (!((mm >> (10 * d + c)) & 1u)) continue;
            for (int a = 0; a < n; ++a)
              for (int b = 0; b < n; ++b)
                {
                  if (d == 0) move(E(c, n - 1, a, b), E(c + 1, 0, a, b));
                  else if (d == 1) move(E(c, a, n - 1, b), E(c + 2, a, 0, b));
                  else move(E(c, a, b, n - 1), E(c + 4, a, b, 0));
                }
          }
      }
    for (int c = 0; c < cw; ++c)
      {
        const uint32_t cell = g * cw + c;
        if (cell >= in.n_plain) continue;
        for (int q = 0; q < npc; ++q)
          if (dead[c * npc + q] && cnt[c * npc + q] != 0 && !(in.idx[(size_t)cell * npc + q] & CBIT)) return false;
      }
    return true;
  }

  void build_group(uint32_t g, GroupRaw &r) const
  {
    r = GroupRaw();
    const uint32_t c0 = g * cw;
    const int      nv = (int)std::min<uint32_t>(cw, in.n_plain - c0);  // valid cells
    auto ent = [&](int c, int q) { return in.idx[(size_t)(c0 + c) * npc + q]; };
    // ---- own range: DoFs first touched by a cell of the group ----
    std::vector<uint32_t> own;
    own.reserve((size_t)nv * npc);
    for (int c = 0; c < nv; ++c)
      for (int q = 0; q < npc; ++q)
        {
          const uint32_t d = ent(c, q) & ~CBIT, f = first[d];
          if (f >= c0 && f < c0 + (uint32_t)nv) own.push_back(d);
        }
    std::sort(own.begin(), own.end());
    own.erase(std::unique(own.begin(), own.end()), own.end());
    uint32_t own_total = 0;
    if (!own.empty() && own.back() - own.front() + 1 == own.size() && own.size() <= (size_t)in.ocap)
      {
        r.own_base = own.front();
        own_total = (uint32_t)own.size();
      }
    r.own_total = own_total;
    // ---- merges ----
    std::vector<uint8_t> dead;
    r.mm = merge_mask(g);
    if (!simulate(g, r.mm, dead)) { r.mm = 0; simulate(g, 0, dead); }
    // ---- slots ----
    struct Info { uint32_t d; uint16_t primary; uint8_t live, cnt, constrained; };
    std::vector<Info> infos;
    std::unordered_map<uint32_t, int> where;
    where.reserve((size_t)nv * npc);
    r.halo.clear();
    std::vector<uint16_t> raw((size_t)cw * npc, RAW_NONE);
    for (int c = 0; c < nv; ++c)
      for (int q = 0; q < npc; ++q)
        {
          const uint32_t e = ent(c, q), d = e & ~CBIT;
          auto it = where.find(d);
          if (it == where.end())
            {
              Info f{d, 0, 0, 0, 0};
              if (own_total && d >= r.own_base && d - r.own_base < own_total) f.primary = (uint16_t)(d - r.own_base);
              else if (!(e & CBIT))
                {
                  if ((int)r.halo.size() >= in.hmax) return;
                  f.primary = (uint16_t)(own_total + r.halo.size());
                  r.halo.push_back(d);
                }
              it = where.emplace(d, (int)infos.size()).first;
              infos.push_back(f);
            }
          Info &f = infos[it->second];
          if (e & CBIT) { f.constrained = 1; continue; }
          ++f.cnt;
          const bool dd = dead[c * npc + q];
          uint16_t   slot = f.primary;
          if (!dd)
            {
              if (f.live)
                {  // a second cell keeps a partial sum of this DoF: its own copy in the halo list
                  if ((int)r.halo.size() >= in.hmax) return;
                  slot = (uint16_t)(own_total + r.halo.size());
                  r.halo.push_back(d);
                }
              ++f.live;
            }
          raw[c * npc + q] = (uint16_t)(slot | (dd ? STAGE_DEAD : 0));
        }
    r.n_halo = (uint32_t)r.halo.size();
    if ((int)r.n_halo > in.hmax) return;
    r.oflag.assign(own_total, 0);
    for (const Info &f : infos)
      {
        if (f.primary >= own_total) continue;
        if (f.constrained || f.cnt == 0 || f.live == 0) continue;
        r.oflag[f.primary] = (f.live == 1 && f.cnt == mult[f.d]) ? 1 : 2;
      }
    // ---- pos table in the lane order of the kernel ----
    r.pos.assign((size_t)ns * 32, RAW_NONE);
    for (int l = 0; l < 32; ++l)
      {
        const Lane &ln = lanes[l];
        if (ln.c < 0 || ln.c >= nv) continue;
        for (int s = 0; s < ns; ++s) r.pos[(size_t)s * 32 + l] = raw[ln.c * npc + ln.i + n * s];
      }
    r.ok = true;
  }

  // shared-memory wavefronts of a pattern with final slots: slab reads (all lanes), staged writes (live lanes; the others
  // hit the trash slot) and the copies into / out of the buffer (consecutive entries of the load list)
  void conflicts(const std::vector<uint16_t> &fin, const std::vector<uint16_t> &st, uint32_t own_total, uint32_t n_halo, int &rd, int &wr, int &cp) const
  {
    const int banks = 128 / in.wb, half = in.wb == 8 ? 16 : 32;
    auto worst_of = [&](const uint16_t *slots, int cnt) {
      int w = 0, nb[32] = {0};
      uint16_t seen[32][32];
      for (int t = 0; t < cnt; ++t)
        {
          const int b = slots[t] % banks;
          bool      dup = false;
          for (int k = 0; k < nb[b]; ++k) dup |= seen[b][k] == slots[t];
          if (!dup) seen[b][nb[b]++] = slots[t];
          w = std::max(w, nb[b]);
        }
      return w;
    };
    rd = wr = cp = 0;
    uint16_t tmp[32];
    for (int s = 0; s < ns; ++s)
      for (int h0 = 0; h0 < 32; h0 += half)
        for (int pass = 0; pass < 2; ++pass)
          {
            for (int l = 0; l < half; ++l)
              {
                const uint16_t v = fin[(size_t)s * 32 + h0 + l];
                tmp[l] = (pass == 1 && (v & STAGE_DEAD)) ? (uint16_t)(in.xcap - 1) : (uint16_t)(v & 0x7fffu);
              }
            (pass == 0 ? rd : wr) += worst_of(tmp, half);
          }
    for (uint32_t e0 = 0; e0 < own_total; e0 += half) cp += worst_of(&st[e0], (int)std::min<uint32_t>(half, own_total - e0));
    for (uint32_t e0 = 0; e0 < n_halo; e0 += half) cp += worst_of(&st[own_total + e0], (int)std::min<uint32_t>(half, n_halo - e0));
  }

  // Slots of the load-list entries: an entry's bank is chosen (greedy colouring + refinement) such that the entries one
  // shared-memory instruction touches lie in different banks; returns false if the group does not fit the buffer.
  bool finalize(const GroupRaw &r, std::vector<uint16_t> &st, std::vector<uint16_t> &fin, std::vector<uint16_t> &hperm) const
  {
    const int B = 128 / in.wb, half = in.wb == 8 ? 16 : 32;
    const int L = (int)(r.own_total + r.n_halo), Z = L;  // pseudo entry Z = the zero / trash slot
    if (L > in.xcap - 1 || L > in.lcap) return false;
    std::vector<std::vector<int>> instr;
    auto add_instr = [&](std::vector<int> &v) {
      std::sort(v.begin(), v.end());
      v.erase(std::unique(v.begin(), v.end()), v.end());
      if (v.size() > 1) instr.push_back(v);
    };
    for (int s = 0; s < ns; ++s)
      for (int h0 = 0; h0 < 32; h0 += half)
        {
          std::vector<int> rdv, wrv;
          for (int l = h0; l < h0 + half; ++l)
            {
              const uint16_t v = r.pos[(size_t)s * 32 + l];
              if (v == RAW_NONE) { rdv.push_back(Z); wrv.push_back(Z); continue; }
              rdv.push_back(v & 0x7fffu);
              wrv.push_back((v & STAGE_DEAD) ? Z : (v & 0x7fffu));
            }
          add_instr(rdv);
          add_instr(wrv);
        }
    const int n_slab_instr = (int)instr.size();
    for (int e0 = 0; e0 < (int)r.own_total; e0 += half)
      {
        std::vector<int> v;
        for (int e = e0; e < std::min<int>(e0 + half, r.own_total); ++e) v.push_back(e);
        add_instr(v);
      }
    for (int e0 = 0; e0 < (int)r.n_halo; e0 += half)
      {
        std::vector<int> v;
        for (int e = e0; e < std::min<int>(e0 + half, r.n_halo); ++e) v.push_back((int)r.own_total + e);
        add_instr(v);
      }
    std::vector<std::vector<int>> of(L + 1);  // instructions of an entry
    for (int k = 0; k < (int)instr.size(); ++k)
      for (int e : instr[k]) of[e].push_back(k);
    std::vector<int> colour(L + 1, -1), cap(B, 0), used(B, 0);
    for (int b = 0; b < B; ++b) cap[b] = (in.xcap - 1 - b + B - 1) / B;
    colour[Z] = (in.xcap - 1) % B;
    // cost of giving entry e colour b: wavefronts its instructions would need beyond one.  The slab reads and staged
    // writes sit on the dependent path of a warp, the list passes (cp.async in, stores out) do not: they weigh less
    std::vector<int> weight(instr.size(), 3);
    for (int k = n_slab_instr; k < (int)instr.size(); ++k) weight[k] = 2;
    auto costs = [&](int e, std::vector<int> &c) {
      c.assign(B, 0);
      for (int k : of[e])
        {
          int cnt[32] = {0}, worst = 0;
          for (int o : instr[k]) if (o != e && colour[o] >= 0) worst = std::max(worst, ++cnt[colour[o]]);
          for (int b = 0; b < B; ++b) c[b] += weight[k] * std::max(0, cnt[b] + 1 - std::max(worst, 1));
        }
    };
    std::vector<int> order(L);
    for (int e = 0; e < L; ++e) order[e] = e;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return of[a].size() > of[b].size(); });
    std::vector<int> c;
    for (int e : order)
      {
        costs(e, c);
        int best = -1;
        for (int b = 0; b < B; ++b)
          if (used[b] < cap[b] && (best < 0 || c[b] < c[best] || (c[b] == c[best] && used[b] < used[best]))) best = b;
        if (best < 0) return false;
        colour[e] = best; ++used[best];
      }
    // refinement on the exact cost: sum over the instructions of weight x (largest number of entries in one bank)
    auto worst_of = [&](int k) {
      int cnt[32] = {0}, w = 0;
      for (int o : instr[k]) w = std::max(w, ++cnt[colour[o]]);
      return w;
    };
    auto local = [&](int e) { int t = 0; for (int k : of[e]) t += weight[k] * worst_of(k); return t; };
    for (int pass = 0; pass < 10; ++pass)
      {
        bool changed = false;
        for (int e : order)
          {
            const int a = colour[e];
            int best = a, bestc = local(e);
            for (int b = 0; b < B; ++b)
              {
                if (b == a || used[b] >= cap[b]) continue;
                colour[e] = b;
                const int t = local(e);
                if (t < bestc) { bestc = t; best = b; }
              }
            colour[e] = best;
            if (best != a) { --used[a]; ++used[best]; changed = true; }
          }
        if (!changed) break;
      }
    // the order of the halo list is free: deal the entries out bank by bank, so that consecutive entries (one copy
    // instruction) lie in different banks.  hperm[new position] = old rank
    hperm.clear();
    {
      std::vector<std::vector<int>> bucket(B);
      for (int h = 0; h < (int)r.n_halo; ++h) bucket[colour[r.own_total + h]].push_back(h);
      size_t round = 0;
      while (hperm.size() < r.n_halo)
        {
          for (int b = 0; b < B; ++b)
            if (round < bucket[b].size()) hperm.push_back((uint16_t)bucket[b][round]);
          ++round;
        }
    }
    std::vector<uint16_t> newpos(r.n_halo);
    for (int k = 0; k < (int)r.n_halo; ++k) newpos[hperm[k]] = (uint16_t)k;
    st.assign(in.lcap, (uint16_t)(in.xcap - 1));
    std::vector<int>      nextk(B, 0);
    std::vector<uint16_t> slot_of(L);
    for (int e = 0; e < L; ++e) slot_of[e] = (uint16_t)(colour[e] + B * nextk[colour[e]]++);
    for (int e = 0; e < (int)r.own_total; ++e) st[e] = slot_of[e];
    for (int h = 0; h < (int)r.n_halo; ++h) st[r.own_total + newpos[h]] = slot_of[r.own_total + h];
    fin.resize(r.pos.size());
    for (size_t t = 0; t < r.pos.size(); ++t)
      {
        const uint16_t v = r.pos[t];
        fin[t] = v == RAW_NONE ? (uint16_t)((in.xcap - 1) | STAGE_DEAD) : (uint16_t)(slot_of[v & 0x7fffu] | (v & STAGE_DEAD));
      }
    return true;
  }
};

}  // namespace

void build_stage_plan(const StagePlanIn &in, StagePlan &out)
{
  MFG_REQUIRE(in.n >= 2 && in.cw >= 1 && in.cw * in.n <= 32, "stage plan: bad group shape");
  MFG_REQUIRE(in.ocap % 32 == 0 && in.ocap >= in.cw * in.n * in.n * in.n && in.lcap == in.ocap + in.hmax && in.lcap < 0x7fff, "stage plan: bad list capacities");
  MFG_REQUIRE(in.xcap >= 64 && in.xcap * in.wb <= 0x8000 && in.xcap % 2 == 0, "stage plan: bad staging capacity");
  out = StagePlan();
  for (int a = 0; a < 8; ++a) out.class_pat[a] = STAGE_NOPAT;
  Builder B(in);
  const uint32_t ng = (in.n_plain + in.cw - 1) / in.cw;
  out.n_groups = ng;
  const int ns2 = (B.ns + 1) / 2;  // position rows are stored in pairs
  out.pstride = ((STAGE_PH + 2 * ns2 * 32 + 2 * in.ocap + in.hmax) + 7) / 8 * 8;
  out.gdesc.assign((size_t)ng * 4, 0);
  if (ng == 0) return;
  B.count();
  std::unordered_map<std::string, int> pid;
  struct Pat { std::vector<uint16_t> st, fin, hperm; std::vector<uint8_t> oflag; bool fits; int rd, wr, cp; uint32_t own_total, n_halo, plain, red; };
  std::vector<Pat> pats;
  // groups in parallel (chunk by chunk: a raw group description is a few KB), patterns deduplicated sequentially
  constexpr uint32_t CHUNK = 16384;
  std::vector<GroupRaw> raws(std::min(ng, CHUNK));
  const unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
  for (uint32_t gs = 0; gs < ng; gs += CHUNK)
  {
    const uint32_t ge = std::min(ng, gs + CHUNK);
    {
      std::atomic<uint32_t> next(gs);
      auto work = [&]() {
        for (;;)
          {
            const uint32_t g0 = next.fetch_add(64);
            if (g0 >= ge) break;
            for (uint32_t g = g0; g < std::min(ge, g0 + 64); ++g) B.build_group(g, raws[g - gs]);
          }
      };
      std::vector<std::thread> th;
      for (unsigned t = 1; t < nt; ++t) th.emplace_back(work);
      work();
      for (auto &t : th) t.join();
    }
  for (uint32_t g = gs; g < ge; ++g)
    {
      GroupRaw &r = raws[g - gs];
      uint32_t *gd = &out.gdesc[(size_t)g * 4];
      gd[2] = STAGE_NOPAT << 16;
      if (!r.ok) { out.fallback.push_back(g); continue; }
      const std::string key = r.key();
      auto it = pid.find(key);
      if (it == pid.end())
        {
          Pat p;
          p.fits = B.finalize(r, p.st, p.fin, p.hperm);
          p.own_total = r.own_total; p.n_halo = r.n_halo;
          p.rd = p.wr = p.cp = 0;
          if (p.fits) B.conflicts(p.fin, p.st, r.own_total, r.n_halo, p.rd, p.wr, p.cp);
          p.oflag = r.oflag;
          p.plain = p.red = 0;
          for (uint8_t f : r.oflag) { p.plain += f == 1; p.red += f == 2; }
          it = pid.emplace(key, (int)pats.size()).first;
          pats.push_back(std::move(p));
        }
      const Pat &p = pats[it->second];
      if (!p.fits || it->second >= (int)STAGE_NOPAT) { out.fallback.push_back(g); continue; }
      gd[0] = r.own_base;
      gd[1] = (uint32_t)out.halo.size();
      gd[2] = r.n_halo | ((uint32_t)it->second << 16);
      gd[3] = r.mm;
      for (uint32_t k2 = 0; k2 < r.n_halo; ++k2) out.halo.push_back(r.halo[p.hperm[k2]]);
      ++out.n_staged;
      out.n_own += p.own_total;
      out.n_halo += r.n_halo;
      out.n_plain_dofs += p.plain;
      out.n_red_dofs += p.red + r.n_halo;
      out.rd_wavefronts += p.rd;
      out.wr_wavefronts += p.wr;
      out.cp_wavefronts += p.cp;
    }
  }
  // most frequent pattern of every class of groups: the kernel keeps its tables in shared memory
  {
    const int nc = std::max(1, std::min(in.nclass, 8));
    for (int a = 0; a < 8; ++a) out.class_pat[a] = STAGE_NOPAT;
    std::vector<std::vector<uint32_t>> hist(nc, std::vector<uint32_t>(pats.size(), 0));
    for (uint32_t g = 0; g < ng; ++g)
      {
        const uint32_t pat = out.gdesc[(size_t)g * 4 + 2] >> 16;
        if (pat != STAGE_NOPAT) ++hist[g % nc][pat];
      }
    for (int a = 0; a < nc; ++a)
      {
        uint32_t best = 0;
        for (size_t k = 0; k < pats.size(); ++k)
          if (hist[a][k] > best) { best = hist[a][k]; out.class_pat[a] = (uint32_t)k; }
      }
  }
  // the kernel reads whole 32-entry rows of the halo list
  out.halo.resize(out.halo.size() + 32, 0);
  out.n_patterns = (uint32_t)pats.size();
  out.ptab.assign((size_t)pats.size() * out.pstride, 0);
  for (size_t k = 0; k < pats.size(); ++k)
    {
      const Pat &p = pats[k];
      if (!p.fits) continue;
      uint16_t *t = &out.ptab[k * out.pstride];
      t[STAGE_H_OWN] = (uint16_t)p.own_total; t[STAGE_H_NHALO] = (uint16_t)p.n_halo;
      // pos [ns2][32] uint32: rows 2 s2 (low half) and 2 s2 + 1 (high half) of the position table
      uint32_t *pos32 = reinterpret_cast<uint32_t *>(t + STAGE_PH);
      for (int s2 = 0; s2 < ns2; ++s2)
        for (int l = 0; l < 32; ++l)
          {
            // (slot * sizeof(Number) | dead flag)
            auto enc = [&](uint16_t v) { return (uint32_t)(((v & 0x7fffu) * in.wb) | (v & STAGE_DEAD)); };
            const uint32_t lo = enc(p.fin[(size_t)(2 * s2) * 32 + l]);
            const uint32_t hi = enc(2 * s2 + 1 < B.ns ? p.fin[(size_t)(2 * s2 + 1) * 32 + l] : (uint16_t)((in.xcap - 1) | STAGE_DEAD));
            pos32[s2 * 32 + l] = lo | (hi << 16);
          }
      // own [ocap] uint32: slot | flag << 16 ; halo slots [hmax] uint16
      uint32_t *own32 = pos32 + ns2 * 32;
      for (int e = 0; e < in.ocap; ++e) own32[e] = (uint32_t)(e < (int)p.own_total ? (uint32_t)p.st[e] * in.wb | ((uint32_t)p.oflag[e] << 16) : (uint32_t)(in.xcap - 1) * in.wb);
      uint16_t *hs = t + STAGE_PH + 2 * ns2 * 32 + 2 * in.ocap;
      for (int h = 0; h < in.hmax; ++h) hs[h] = (uint16_t)((h < (int)p.n_halo ? p.st[p.own_total + h] : (uint16_t)(in.xcap - 1)) * in.wb);
    }
}

}  // namespace mfg
