// aqc_program.h -- compiled circuit programs (host side): gate units -> tile passes -> stages.
// Included by aqc_sv.cu.  See DESIGN.md section 3.
#pragma once
// ------------------------------------------------------------------------------------------
// compiled program: passes -> stages -> units
// ------------------------------------------------------------------------------------------
constexpr int kThreads = 128;  // threads per CTA of the pass kernel
constexpr int kMaxUnits = 3;   // units fused into one stage (a Trotter triplet)
constexpr int kStageUnits = 5; // dense-stage programs: two front gates + a triplet on one bit pair
constexpr int kMaxTileBits = 12;

struct UnitDesc {
  int32_t kind;
  int32_t flags;
  int32_t theta;  // index of the unit's first angle
  int32_t slot;   // first raw-gradient accumulator of this unit OCCURRENCE (5 per unit)
};
struct StageDesc {
  int32_t p, q;  // tile-local bit positions held in registers, p > q
  int32_t nunits;
  int32_t triplet;  // 1: Trotter triplet (ctrl hi / lo / hi, Rz(-pi/2) first, Rz(+pi/2) last)
  UnitDesc u[kStageUnits];  // legacy / scale-free programs use at most kMaxUnits of them
};
static_assert(sizeof(StageDesc) == 96, "StageDesc layout");

struct PassDesc {
  int32_t tb;       // tile bits
  int32_t nstages;  // stages in this pass
  int32_t stage0;   // first stage in the program's stage array
  int32_t nouter;   // number of index bits outside the tile
  uint8_t bitpos[16];    // global bit position of tile-local bit k
  uint8_t outerpos[48];  // global bit positions of the non-tile bits, ascending
};

struct Program {
  std::vector<PassDesc> passes;
  std::vector<StageDesc> stages;
  StageDesc* d_stages = nullptr;
  PassDesc* d_passes = nullptr;
  // sharded execution: passes [epoch_pass0[e], epoch_pass0[e+1]) need data layout epoch_layout[e]
  std::vector<int> epoch_pass0, epoch_layout;
};

struct HostUnit {
  int kind;  // 0 front, 1 block
  int qa;    // front qubit | control
  int qb;    // -1 | target
  int theta;
  int flags;
  int seq;  // index of the unit in forward circuit order (names its gradient accumulators)
};

struct aqc_circuit {
  int n = 0;
  int ent = 0;
  int trotter = 0;
  int nb = 0;  // blocks in full layers
  int half = 0;
  int tpb = 4;
  int nthetas = 0;
  std::vector<int> ctrl, targ;
};

// read-only accessors for aqc_mps.cu
int aqc_circ_n(const aqc_circuit* c) { return c->n; }
int aqc_circ_ent(const aqc_circuit* c) { return c->ent; }
int aqc_circ_trotter(const aqc_circuit* c) { return c->trotter; }
int aqc_circ_nb(const aqc_circuit* c) { return c->nb; }
int aqc_circ_half(const aqc_circuit* c) { return c->half; }
int aqc_circ_tpb(const aqc_circuit* c) { return c->tpb; }
int aqc_circ_nthetas(const aqc_circuit* c) { return c->nthetas; }
int aqc_circ_ctrl(const aqc_circuit* c, int i) { return c->ctrl[i]; }
int aqc_circ_targ(const aqc_circuit* c, int i) { return c->targ[i]; }

static void build_units(const aqc_circuit& c, bool reversed, std::vector<HostUnit>& out) {
  out.clear();
  for (int q = 0; q < c.n; ++q) out.push_back({0, q, -1, 3 * q, 0, q});
  const int total = c.nb + c.half;
  for (int i = 0; i < total; ++i) {
    const int im = c.nb > 0 ? i % c.nb : 0;
    int flags = 0;
    if (c.trotter != AQC_GENERIC) {
      if (i % 3 == 0) flags |= F_PRE;
      if (i % 3 == 2) flags |= F_POST;
    }
    out.push_back({1, c.ctrl[im], c.targ[im], 3 * c.n + c.tpb * im, flags, c.n + i});
  }
  if (reversed) std::reverse(out.begin(), out.end());
}

// Greedy tile-pass scheduler.  `units` is the gate-unit sequence in execution order; units on
// disjoint qubits commute, so a unit may run in the current pass iff all its qubits are inside
// the tile and none of them is touched by an earlier unit that had to be deferred.
// `units` carry PHYSICAL bit positions in qa / qb.  Passes are appended to `prog`.
// `max_units` caps the units of a stage; `merge_fronts` (dense-stage programs) lets a block unit
// join the front-gate stage that holds its qubits, so the front layer costs no stages of its own.
// `plan` (optional): the tile of pass i is the bit set plan[i] (see plan_tiles) instead of the greedy choice.
static void build_program_units_impl(const std::vector<HostUnit>& units, int nbits, int tb_max,
                                     int lowbits, Program& prog, int max_units, bool merge_fronts,
                                     const std::vector<unsigned long long>* plan) {
  const int qoff = 0;
  const size_t pass_begin = prog.passes.size();
  const int tb = std::min(nbits, tb_max);
  const int low = std::min(lowbits, tb);

  std::vector<char> done(units.size(), 0);
  size_t ndone = 0;
  while (ndone < units.size() || prog.passes.size() == pass_begin) {
    std::vector<char> intile(nbits, 0), blocked(nbits, 0);
    int ntile = 0;
    const size_t pass_no = prog.passes.size() - pass_begin;
    if (plan && pass_no < plan->size()) {
      for (int b = 0; b < nbits; ++b)
        if (((*plan)[pass_no] >> b) & 1ull) intile[b] = 1, ++ntile;
    } else {
      for (int b = 0; b < low; ++b) intile[b] = 1, ++ntile;
    }
    std::vector<int> picked;
    for (size_t k = 0; k < units.size(); ++k) {
      if (done[k]) continue;
      const HostUnit& u = units[k];
      const int ba = u.qa + qoff, bb = u.kind ? u.qb + qoff : -1;
      const bool blk = blocked[ba] || (bb >= 0 && blocked[bb]);
      int need = (intile[ba] ? 0 : 1) + ((bb >= 0 && !intile[bb]) ? 1 : 0);
      if (!blk && ntile + need <= tb) {
        if (!intile[ba]) intile[ba] = 1, ++ntile;
        if (bb >= 0 && !intile[bb]) intile[bb] = 1, ++ntile;
        picked.push_back((int)k);
      } else {
        blocked[ba] = 1;
        if (bb >= 0) blocked[bb] = 1;
      }
    }
    // pad the tile with the lowest free bits (longer contiguous runs)
    for (int b = 0; b < nbits && ntile < tb; ++b)
      if (!intile[b]) intile[b] = 1, ++ntile;

    PassDesc pd;
    memset(&pd, 0, sizeof(pd));
    pd.tb = tb;
    pd.stage0 = (int)prog.stages.size();
    std::vector<int> local(nbits, -1);
    int kt = 0, ko = 0;
    for (int b = 0; b < nbits; ++b) {
      if (intile[b]) {
        local[b] = kt;
        pd.bitpos[kt++] = (uint8_t)b;
      } else {
        pd.outerpos[ko++] = (uint8_t)b;
      }
    }
    pd.nouter = ko;

    // group the picked units into stages (register-resident quads on one bit pair)
    struct Open {
      int a, b;  // tile-local bits (b == -1: partner still free, front-only stage)
      bool front;
      StageDesc sd;
    };
    std::vector<Open> open;
    std::vector<int> last(tb, -1);  // last stage that touched tile-local bit
    auto put = [&](Open& o, int kind, const HostUnit& u) {
      UnitDesc& d = o.sd.u[o.sd.nunits++];
      d.kind = kind;
      d.flags = u.flags;
      d.theta = u.theta;
      d.slot = 5 * u.seq;
    };
    // units are stored with *qubit roles*; the LO/HI kind is fixed up when the stage closes
    struct Pending {
      int stage;
      int slot;
      int la, lb;
      bool front;
    };
    std::vector<Pending> pend;
    for (int k : picked) {
      const HostUnit& u = units[k];
      const int la = local[u.qa + qoff];
      const int lb = u.kind ? local[u.qb + qoff] : -1;
      int s = -1;
      if (u.kind == 0) {
        // front gate: join an open front stage that has a free partner seat
        // (legal iff no stage created after it has touched this bit)
        if (merge_fronts && last[la] >= 0 && !open[last[la]].front &&
            open[last[la]].sd.nunits < max_units)
          s = last[la];  // reversed sweeps: the front gate follows the last stage on its qubit
        for (size_t i = 0; s < 0 && i < open.size(); ++i)
          if (open[i].front && open[i].b < 0 && open[i].a != la && open[i].sd.nunits < 2 &&
              (int)i > last[la]) {
            s = (int)i;
            open[i].b = la;
            break;
          }
        if (s < 0) {
          Open o;
          memset(&o.sd, 0, sizeof(o.sd));
          o.a = la;
          o.b = -1;
          o.front = true;
          open.push_back(o);
          s = (int)open.size() - 1;
        }
        last[la] = s;
      } else {
        const int sa = last[la], sb = last[lb];
        if (sa >= 0 && sa == sb && (merge_fronts || !open[sa].front) && open[sa].sd.nunits < max_units &&
            ((open[sa].a == la && open[sa].b == lb) || (open[sa].a == lb && open[sa].b == la))) {
          s = sa;
          open[sa].front = false;  // no further front gate may take a seat here
        } else if (merge_fronts && sa >= 0 && sa > sb && open[sa].front && open[sa].b < 0 &&
                   open[sa].a == la && open[sa].sd.nunits < max_units) {
          // single front gate on la with a free partner seat; everything on lb happened earlier
          s = sa;
          open[sa].b = lb;
          open[sa].front = false;
        } else if (merge_fronts && sb >= 0 && sb > sa && open[sb].front && open[sb].b < 0 &&
                   open[sb].a == lb && open[sb].sd.nunits < max_units) {
          s = sb;
          open[sb].b = la;
          open[sb].front = false;
        } else {
          Open o;
          memset(&o.sd, 0, sizeof(o.sd));
          o.a = la;
          o.b = lb;
          o.front = false;
          open.push_back(o);
          s = (int)open.size() - 1;
        }
        last[la] = last[lb] = s;
      }
      pend.push_back({s, open[s].sd.nunits, la, lb, u.kind == 0});
      put(open[s], U_NONE, u);
    }
    // A front stage whose partner seat stayed free gets any other tile bit as a passive partner.
    for (auto& o : open)
      if (o.b < 0) o.b = (o.a == 0) ? 1 : 0;
    for (auto& pe : pend) {
      Open& o = open[pe.stage];
      const int hi = std::max(o.a, o.b), lo = std::min(o.a, o.b);
      o.sd.p = hi;
      o.sd.q = lo;
      UnitDesc& d = o.sd.u[pe.slot];
      if (pe.front)
        d.kind = (pe.la == hi) ? U_FRONT_HI : U_FRONT_LO;
      else
        d.kind = (pe.la == hi) ? U_BLOCK_CHI : U_BLOCK_CLO;
      (void)lo;
    }
    for (auto& o : open) {
      const StageDesc& d = o.sd;
      o.sd.triplet = (d.nunits == 3 && d.u[0].kind == U_BLOCK_CHI && d.u[1].kind == U_BLOCK_CLO &&
                      d.u[2].kind == U_BLOCK_CHI && d.u[0].flags == F_PRE && d.u[1].flags == 0 &&
                      d.u[2].flags == F_POST)
                         ? 1
                         : 0;
      prog.stages.push_back(o.sd);
    }
    pd.nstages = (int)open.size();
    prog.passes.push_back(pd);
    for (int k : picked) done[k] = 1;
    ndone += picked.size();
    if (picked.empty() && ndone < units.size()) break;  // cannot happen (tb >= 2)
  }
}

// Tile planner.  The greedy scheduler above grows each tile along the unit order, which follows the
// layers of the circuit: on a brick-wall circuit it sweeps one window after the other across the qubits
// and pays a pass per window and layer group.  Which units a pass can run depends only on its bit SET
// (light cones inside the set), so the sets are searched instead: candidates are the fixed low bits plus
// up to two contiguous runs of the other bits, a beam search over passes maximises the units done, and
// the plan is used when it needs FEWER passes than the greedy schedule (n = 20, L = 2 gradient program:
// 4 instead of 5 passes; n = 22: 4 instead of 6).  Fewer passes = fewer trips of the state through the
// memory system and fewer launches.  AQC_TILE_PLAN=0 keeps the greedy schedule.
static int env_int(const char* name, int dflt);
static bool plan_tiles(const std::vector<HostUnit>& units, int nbits, int tb, int low, int greedy_passes,
                       std::vector<unsigned long long>& plan) {
  const int U = (int)units.size();
  if (greedy_passes < 2 || U == 0 || nbits > 62 || tb >= nbits) return false;
  const int k = tb - low;
  if (k < 2) return false;
  // candidate sets
  std::vector<unsigned long long> cands;
  {
    std::set<unsigned long long> seen;
    const unsigned long long lowmask = (1ull << low) - 1ull;
    for (int a = low; a < nbits; ++a)
      for (int la = (k + 1) / 2; la <= k; ++la)
        for (int b = low; b < nbits; ++b) {
          unsigned long long m = lowmask;
          for (int q = a; q < a + la && q < nbits; ++q) m |= 1ull << q;
          for (int q = b; q < b + (k - la) && q < nbits; ++q) m |= 1ull << q;
          int cnt = __builtin_popcountll(m);
          for (int q = 0; q < nbits && cnt < tb; ++q)  // pad with the lowest free bits
            if (!((m >> q) & 1ull)) m |= 1ull << q, ++cnt;
          if (cnt == tb && seen.insert(m).second) cands.push_back(m);
        }
  }
  struct State {
    std::vector<char> done;
    int ndone, first;  // first: index of the first unit not yet done
    int parent;        // index in the previous level
    unsigned long long tile;
  };
  // the search visits beam x candidates x units per level (a few ns each, spread over up to 8 host threads):
  // the beam narrows for long unit lists, and deep circuits (thousands of blocks, hundreds of passes) keep the
  // greedy schedule instead of spending seconds here
  int beam = env_int("AQC_TILE_PLAN_BEAM", 24);
  auto effort = [&](int bm) { return (double)bm * (double)cands.size() * (double)U * (double)greedy_passes; };
  while (beam > 6 && effort(beam) > 6e8) beam /= 2;
  if (effort(beam) > 6e8) return false;
  const int nthreads = (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
  std::vector<std::vector<State>> levels;
  levels.push_back({State{std::vector<char>((size_t)U, 0), 0, 0, -1, 0ull}});
  std::vector<unsigned long long> bits((size_t)U);
  for (int i = 0; i < U; ++i)
    bits[(size_t)i] = (1ull << units[(size_t)i].qa) | (units[(size_t)i].kind ? (1ull << units[(size_t)i].qb) : 0ull);
  for (int level = 1; level < greedy_passes; ++level) {
    const std::vector<State>& prev = levels.back();
    // every (state, candidate) pair is simulated independently: candidates are dealt out to the threads
    std::vector<std::vector<State>> part((size_t)nthreads);
    auto work = [&](int t) {
      std::vector<State>& out = part[(size_t)t];
      for (size_t ci = (size_t)t; ci < cands.size(); ci += (size_t)nthreads) {
        const unsigned long long S = cands[ci];
        for (int pi = 0; pi < (int)prev.size(); ++pi) {
          const State& st = prev[(size_t)pi];
          unsigned long long blocked = 0ull;
          int nd = st.ndone;
          std::vector<char> d;
          for (int i = st.first; i < U; ++i) {
            if (st.done[(size_t)i]) continue;
            const unsigned long long b = bits[(size_t)i];
            if ((b & ~S) == 0ull && (b & blocked) == 0ull) {
              if (d.empty()) d = st.done;
              d[(size_t)i] = 1;
              ++nd;
            } else {
              blocked |= b;
              if ((~blocked & S) == 0ull) break;  // every tile bit is blocked: nothing else can run
            }
          }
          if (nd == st.ndone) continue;
          int first = st.first;
          while (first < U && d[(size_t)first]) ++first;
          out.push_back(State{std::move(d), nd, first, pi, S});
        }
      }
    };
    if (nthreads > 1 && (double)prev.size() * (double)cands.size() * (double)U > 2e6) {
      std::vector<std::thread> pool;
      for (int t = 1; t < nthreads; ++t) pool.emplace_back(work, t);
      work(0);
      for (std::thread& th : pool) th.join();
    } else {
      for (int t = 0; t < nthreads; ++t) work(t);
    }
    std::vector<State> all;
    for (std::vector<State>& v : part) {
      for (State& st : v) all.push_back(std::move(st));
      v.clear();
    }
    if (all.empty()) return false;
    // deterministic order whatever the thread count: most units done first, then parent, then tile set
    std::sort(all.begin(), all.end(), [](const State& x, const State& y) {
      if (x.ndone != y.ndone) return x.ndone > y.ndone;
      if (x.parent != y.parent) return x.parent < y.parent;
      return x.tile < y.tile;
    });
    std::vector<State> next;
    std::set<std::vector<char>> seen;
    for (State& st : all) {
      if ((int)next.size() >= beam) break;
      if (seen.insert(st.done).second) next.push_back(std::move(st));
    }
    levels.push_back(std::move(next));
    if (levels.back()[0].ndone == U) {
      plan.clear();
      int idx = 0;
      for (int l = (int)levels.size() - 1; l >= 1; --l) {
        plan.push_back(levels[(size_t)l][(size_t)idx].tile);
        idx = levels[(size_t)l][(size_t)idx].parent;
      }
      std::reverse(plan.begin(), plan.end());
      return true;
    }
  }
  return false;
}

static void build_program_units(const std::vector<HostUnit>& units, int nbits, int tb_max,
                                int lowbits, Program& prog, int max_units = kMaxUnits,
                                bool merge_fronts = false) {
  const bool planner = env_int("AQC_TILE_PLAN", 1) != 0;  // (read per program: the tests switch it)
  Program greedy, planned;
  build_program_units_impl(units, nbits, tb_max, lowbits, greedy, max_units, merge_fronts, nullptr);
  const Program* best = &greedy;
  if (planner) {
    const int tb = std::min(nbits, tb_max);
    std::vector<unsigned long long> plan;
    if (plan_tiles(units, nbits, tb, std::min(lowbits, tb), (int)greedy.passes.size(), plan)) {
      build_program_units_impl(units, nbits, tb_max, lowbits, planned, max_units, merge_fronts, &plan);
      if (planned.passes.size() < greedy.passes.size()) best = &planned;
    }
  }
  const int stage_off = (int)prog.stages.size();
  for (PassDesc pd : best->passes) {
    pd.stage0 += stage_off;
    prog.passes.push_back(pd);
  }
  prog.stages.insert(prog.stages.end(), best->stages.begin(), best->stages.end());
}

static void build_program(const aqc_circuit& c, int qoff, int nbits, int tb_max, int lowbits,
                          bool reversed, Program& prog, int max_units = kMaxUnits,
                          bool merge_fronts = false) {
  std::vector<HostUnit> units;
  build_units(c, reversed, units);
  for (HostUnit& u : units) {
    u.qa += qoff;
    if (u.kind) u.qb += qoff;
  }
  prog.passes.clear();
  prog.stages.clear();
  prog.epoch_pass0.assign(1, 0);
  prog.epoch_layout.assign(1, 0);
  build_program_units(units, nbits, tb_max, lowbits, prog, max_units, merge_fronts);
}

// ---- global-qubit sharding (one state over 2^g GPUs) --------------------------------------------
// The top g index bits select the rank.  Two data layouts alternate:
//   layout A: qubits n-g..n-1 are global; qubits 0..g-1 sit on the TOP g local bits;
//   layout B: qubits 0..g-1 are global; qubits n-g..n-1 sit on the top g local bits;
// the other qubits occupy local bits 0..nl-g-1 (q -> q - g) in both.  Switching layouts is the
// block transpose new[rank c][chunk r] = old[rank r][chunk c] over chunks of 2^(nl-g) amplitudes
// (all-to-all over NVLink).  An epoch runs every gate unit that is executable without touching a
// global qubit (a light-cone trapezoid of the brick-wall circuit); then the layout is switched.
static int phys_bit(int q, int n, int g, int layout) {
  const int nl = n - g;
  if (q < g) return layout == 0 ? nl - g + q : -1;
  if (q >= n - g) return layout == 0 ? -1 : nl - g + (q - (n - g));
  return q - g;
}

static int build_program_sharded(const aqc_circuit& c, int g, int tb_max, int lowbits, bool reversed,
                                 Program& prog, std::string& err, int max_units = kMaxUnits,
                                 bool merge_fronts = false) {
  const int n = c.n, nl = n - g;
  if (nl - g < 2 || 2 * g > n - 2) {
    err = "too few qubits for this number of GPUs";
    return AQC_EINVAL;
  }
  std::vector<HostUnit> units;
  build_units(c, reversed, units);
  prog.passes.clear();
  prog.stages.clear();
  prog.epoch_pass0.clear();
  prog.epoch_layout.clear();
  std::vector<char> done(units.size(), 0);
  size_t ndone = 0;
  int layout = 0, idle = 0;
  while (ndone < units.size()) {
    std::vector<char> blocked(n, 0);
    std::vector<HostUnit> now;
    std::vector<size_t> ids;
    for (size_t k = 0; k < units.size(); ++k) {
      if (done[k]) continue;
      const HostUnit& u = units[k];
      const int pa = phys_bit(u.qa, n, g, layout);
      const int pb = u.kind ? phys_bit(u.qb, n, g, layout) : 0;
      const bool blk = blocked[u.qa] || (u.kind && blocked[u.qb]);
      if (!blk && pa >= 0 && pb >= 0) {
        HostUnit v = u;
        v.qa = pa;
        if (u.kind) v.qb = pb;
        now.push_back(v);
        ids.push_back(k);
      } else {
        blocked[u.qa] = 1;
        if (u.kind) blocked[u.qb] = 1;
      }
    }
    if (now.empty()) {
      if (++idle > 1) {
        err = "circuit cannot be scheduled over global qubits (a unit couples the lowest and highest qubits)";
        return AQC_EINVAL;
      }
      layout ^= 1;
      continue;
    }
    idle = 0;
    prog.epoch_pass0.push_back((int)prog.passes.size());
    prog.epoch_layout.push_back(layout);
    build_program_units(now, nl, tb_max, lowbits, prog, max_units, merge_fronts);
    for (size_t k : ids) done[k] = 1;
    ndone += ids.size();
    layout ^= 1;
  }
  if (prog.epoch_pass0.empty()) {  // circuit without units cannot happen (front layer), keep safe
    prog.epoch_pass0.push_back(0);
    prog.epoch_layout.push_back(0);
  }
  return AQC_OK;
}

