// Plan compiler and C ABI of the B200 contraction engine.
//
// Turns a contraction list (contractn/einsum.py:341-344: positions to pop,
// per-step einsum string) into a device-resident plan: per-step loop modes with
// element strides (no transposes are ever materialised), kernel class per step,
// lifetime-based placement of intermediates in one workspace arena, fusion of
// same-layout Hadamard runs (copy-node hubs) into n-ary launches and of runs of
// tiny steps into a single-CTA interpreter launch.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <set>

#include "ctn_internal.h"

namespace ctn {

static thread_local std::string g_cuda_error;

void set_cuda_error(cudaError_t e, const char* where) {
  g_cuda_error = std::string(where) + ": " + cudaGetErrorString(e);
}

#define CTN_CUDA(expr)                      \
  do {                                      \
    cudaError_t _e = (expr);                \
    if (_e != cudaSuccess) {                \
      set_cuda_error(_e, #expr);            \
      return CTN_ERR_CUDA;                  \
    }                                       \
  } while (0)

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------
// arena: first-fit free list with coalescing (intermediates have nested or
// overlapping lifetimes; stream order makes reuse safe)
// ---------------------------------------------------------------------------
struct Arena {
  std::map<size_t, size_t> free_;  // offset -> size
  size_t top = 0;
  size_t alloc(size_t bytes) {
    bytes = align_up(std::max<size_t>(bytes, 1), 256);
    for (auto it = free_.begin(); it != free_.end(); ++it) {
      if (it->second >= bytes) {
        size_t off = it->first, rest = it->second - bytes;
        free_.erase(it);
        if (rest) free_[off + bytes] = rest;
        return off;
      }
    }
    // grow: extend a free block that touches the top if there is one
    auto last = free_.empty() ? free_.end() : std::prev(free_.end());
    if (last != free_.end() && last->first + last->second == top) {
      size_t off = last->first;
      free_.erase(last);
      top = off + bytes;
      return off;
    }
    size_t off = top;
    top += bytes;
    return off;
  }
  void release(size_t off, size_t bytes) {
    bytes = align_up(std::max<size_t>(bytes, 1), 256);
    auto it = free_.emplace(off, bytes).first;
    auto nx = std::next(it);
    if (nx != free_.end() && it->first + it->second == nx->first) {
      it->second += nx->second;
      free_.erase(nx);
    }
    if (it != free_.begin()) {
      auto pv = std::prev(it);
      if (pv->first + pv->second == it->first) {
        pv->second += it->second;
        free_.erase(it);
      }
    }
  }
};

// ---------------------------------------------------------------------------
// step lowering
// ---------------------------------------------------------------------------
struct LoweredStep {
  std::vector<int32_t> in_slots;
  int32_t out_slot;
  std::vector<ModeDev> out_modes, k_modes;  // fused, extent-1 dropped
  int64_t out_numel = 1, k_numel = 1;
  double flops = 0;
  bool elementwise = false;  // no summed mode, every input laid out like the output
  bool tiny = false;
  GemmDev gemm;
  bool has_gemm = false;
  // label-level view, kept for the consumer-fusion pass
  std::vector<std::vector<int32_t>> in_labels;
  std::vector<int32_t> out_labels;
  std::vector<std::pair<int32_t, GemmMode>> gm, gn, gk;  // unfused GEMM groups with their labels
  int32_t g_n_batch = 0;
  bool fused_consumer = false;
  int32_t desc_step = -1;                 // contraction-list entry whose result this step writes
  bool overridden = false;                // its output got a non-default physical layout
  std::vector<int32_t> row_labels, col_labels, batch_labels;  // GEMM orientation of the output labels
};

// Physical layouts chosen for intermediates: contraction-list entry -> output labels, outermost first.
typedef std::map<int32_t, std::vector<int32_t>> LayoutOverrides;

static void fuse_modes(std::vector<ModeDev>& modes, int n_in, bool out_modes) {
  // drop extent-1 modes, then merge neighbours (outer, inner) whenever
  // stride_outer == stride_inner * extent_inner in every input
  std::vector<ModeDev> kept;
  for (auto& m : modes)
    if (m.extent != 1) kept.push_back(m);
  if (!out_modes) {
    // summed modes may be visited in any order: sort by stride in the first
    // input that has a non-zero stride so that contiguous runs become adjacent
    std::stable_sort(kept.begin(), kept.end(), [n_in](const ModeDev& a, const ModeDev& b) {
      for (int i = 0; i < n_in; ++i)
        if (a.stride[i] != b.stride[i]) return a.stride[i] > b.stride[i];
      return false;
    });
  }
  std::vector<ModeDev> fused;
  for (auto& m : kept) {
    if (!fused.empty()) {
      ModeDev& o = fused.back();
      bool ok = true;
      for (int i = 0; i < n_in; ++i) ok = ok && (o.stride[i] == m.stride[i] * m.extent);
      if (ok) {
        o.extent *= m.extent;
        for (int i = 0; i < n_in; ++i) o.stride[i] = m.stride[i];
        continue;
      }
    }
    fused.push_back(m);
  }
  modes.swap(fused);
}

static void fuse_gemm_group(GemmMode* g, int32_t& n, bool keep_order) {
  // (sw takes part in the compatibility test so that fused-consumer strides stay exact)
  std::vector<GemmMode> kept;
  for (int i = 0; i < n; ++i)
    if (g[i].extent != 1) kept.push_back(g[i]);
  if (!keep_order) {
    std::stable_sort(kept.begin(), kept.end(), [](const GemmMode& a, const GemmMode& b) {
      if (a.sa != b.sa) return a.sa > b.sa;
      return a.sb > b.sb;
    });
  }
  std::vector<GemmMode> fused;
  for (auto& m : kept) {
    if (!fused.empty()) {
      GemmMode& o = fused.back();
      if (o.sa == m.sa * m.extent && o.sb == m.sb * m.extent && o.sc == m.sc * m.extent &&
          o.sw == m.sw * m.extent) {
        o.extent *= m.extent;
        o.sa = m.sa; o.sb = m.sb; o.sc = m.sc; o.sw = m.sw;
        continue;
      }
    }
    fused.push_back(m);
  }
  n = (int32_t)fused.size();
  for (int i = 0; i < n; ++i) g[i] = fused[i];
}

// The order in which the modes of the M (or N) group are flattened into a row index is free:
// offsets into A/B and C are computed from the same decomposition.  If the summed modes are not
// contiguous in an operand but one of its free modes is (unit stride, even extent), make that
// mode the innermost of its group so that tiles of the operand can be staged with 16-byte copies.
static void prefer_unit_stride_innermost(GemmMode* grp, int n, const GemmMode* ks, int n_k, bool use_b) {
  auto stride = [&](const GemmMode& m) { return use_b ? m.sb : m.sa; };
  for (int i = 0; i < n_k; ++i)
    if (stride(ks[i]) == 1) return;  // K-fast: nothing to gain
  for (int i = 0; i + 1 < n; ++i) {
    if (stride(grp[i]) == 1 && grp[i].extent % 4 == 0 && grp[i].extent >= 8) {
      GemmMode m = grp[i];
      for (int j = i; j + 1 < n; ++j) grp[j] = grp[j + 1];
      grp[n - 1] = m;
      return;
    }
  }
}

static int lower_step(Plan& p, const ctn_step_desc& sd, std::vector<int32_t>& live, LoweredStep& ls,
                      const std::vector<int32_t>* phys_order) {
  if (sd.n_inputs < 1 || !sd.positions || !sd.input_ranks) return CTN_ERR_INVALID;
  // pop operands exactly like `[operands.pop(x) for x in inds]` (einsum.py:344)
  for (int i = 0; i < sd.n_inputs; ++i) {
    int32_t pos = sd.positions[i];
    if (pos < 0 || pos >= (int32_t)live.size()) return CTN_ERR_INVALID;
    ls.in_slots.push_back(live[pos]);
    live.erase(live.begin() + pos);
  }
  const int n_in = sd.n_inputs;
  // label -> extent (broadcast-aware), per-input stride sums
  std::map<int32_t, int64_t> extent;
  std::vector<std::map<int32_t, int64_t>> in_stride(n_in);
  std::vector<std::map<int32_t, int64_t>> in_extent(n_in);
  std::vector<std::vector<int32_t>> in_labels(n_in);
  const int32_t* lab = sd.input_labels;
  for (int i = 0; i < n_in; ++i) {
    const Slot& s = p.slots[ls.in_slots[i]];
    if (sd.input_ranks[i] != (int32_t)s.extents.size()) return CTN_ERR_INVALID;
    for (int a = 0; a < sd.input_ranks[i]; ++a, ++lab) {
      int32_t l = *lab;
      in_labels[i].push_back(l);
      if (l < 0) return CTN_ERR_INVALID;
      int64_t e = s.extents[a];
      auto it = extent.find(l);
      if (it == extent.end()) extent[l] = e;
      else if (it->second == 1) it->second = e;
      else if (e != 1 && e != it->second) return CTN_ERR_SHAPE;
      // a repeated label inside one operand walks its diagonal: strides add up
      if (in_extent[i].count(l) && in_extent[i][l] != e) return CTN_ERR_SHAPE;
      in_extent[i][l] = e;
      in_stride[i][l] += s.strides[a];
    }
  }
  // output slot
  Slot out;
  std::set<int32_t> out_set;
  for (int a = 0; a < sd.out_rank; ++a) {
    int32_t l = sd.out_labels[a];
    if (!extent.count(l) || out_set.count(l)) return CTN_ERR_INVALID;
    out_set.insert(l);
    out.labels.push_back(l);
    out.extents.push_back(extent[l]);
  }
  out.strides.assign(sd.out_rank, 1);
  for (int a = sd.out_rank - 2; a >= 0; --a) out.strides[a] = out.strides[a + 1] * out.extents[a + 1];
  if (phys_order && (int)phys_order->size() == sd.out_rank) {
    // an intermediate may live in any dense layout (nobody but the plan's own kernels reads it):
    // `phys_order` lists its labels outermost first
    int64_t st = 1;
    bool ok = true;
    std::vector<int64_t> strides(sd.out_rank, 0);
    for (int a = sd.out_rank - 1; a >= 0 && ok; --a) {
      auto it = std::find(out.labels.begin(), out.labels.end(), (*phys_order)[a]);
      if (it == out.labels.end() || strides[it - out.labels.begin()] != 0) { ok = false; break; }
      strides[it - out.labels.begin()] = st;
      st *= out.extents[it - out.labels.begin()];
    }
    if (ok) { out.strides = strides; ls.overridden = true; }
  }
  for (auto e : out.extents) out.numel *= e;
  ls.out_slot = (int32_t)p.slots.size();
  p.slots.push_back(out);
  live.push_back(ls.out_slot);
  ls.in_labels = in_labels;
  ls.out_labels = out.labels;

  auto make_mode = [&](int32_t l) {
    ModeDev m{};
    m.extent = extent[l];
    for (int i = 0; i < n_in && i < kMaxIn; ++i) {
      auto it = in_stride[i].find(l);
      // absent, or broadcast (extent 1 against a larger one): stride 0
      m.stride[i] = (it == in_stride[i].end() || in_extent[i][l] != m.extent) ? 0 : it->second;
    }
    return m;
  };
  bool any_sum = false;
  double all = 1;
  for (auto& kv : extent) {
    all *= (double)kv.second;
    if (!out_set.count(kv.first)) any_sum = true;
  }
  ls.flops = all * (std::max(1, n_in - 1) + (any_sum ? 1 : 0));
  ls.out_numel = out.numel;

  if (n_in <= kMaxIn) {
    for (int a = 0; a < sd.out_rank; ++a) ls.out_modes.push_back(make_mode(sd.out_labels[a]));
    for (auto& kv : extent)
      if (!out_set.count(kv.first)) ls.k_modes.push_back(make_mode(kv.first));
    for (auto& m : ls.k_modes) ls.k_numel *= m.extent;
    fuse_modes(ls.out_modes, n_in, true);
    fuse_modes(ls.k_modes, n_in, false);
  }
  // same-layout Hadamard product?  every label in the output, every input
  // carrying all output labels in output order with identical (contiguous) strides
  ls.elementwise = !any_sum && !ls.overridden;
  if (ls.elementwise) {
    for (int i = 0; i < n_in && ls.elementwise; ++i) {
      const Slot& s = p.slots[ls.in_slots[i]];
      if (in_labels[i] != out.labels || s.extents != out.extents || s.strides != out.strides)
        ls.elementwise = false;
    }
  }
  if (!ls.elementwise && n_in > kMaxIn) return CTN_ERR_UNSUPPORTED;
  double work = (double)ls.out_numel * (double)ls.k_numel * n_in;
  ls.tiny = n_in <= kMaxIn && work <= 8192.0 && ls.out_numel <= 2048 &&
            ls.out_modes.size() + ls.k_modes.size() <= (size_t)kInterpRecModes;

  // GEMM view for pairwise steps with a summed index and no index that is
  // summed out of a single operand only
  if (n_in == 2 && any_sum) {
    GemmDev g{};
    bool ok = true;
    auto push = [&](GemmMode* grp, int32_t& n, GemmMode m) {
      if (n >= kGemmMaxModes) { ok = false; return; }
      grp[n++] = m;
    };
    std::map<int32_t, int64_t> c_stride;
    for (int a = 0; a < sd.out_rank; ++a) c_stride[sd.out_labels[a]] = p.slots[ls.out_slot].strides[a];
    auto stride_of = [&](int i, int32_t l) -> int64_t {
      auto it = in_stride[i].find(l);
      return (it == in_stride[i].end() || in_extent[i][l] != extent[l]) ? 0 : it->second;
    };
    // groups in output order for M, N, batch; K order is free
    for (int a = 0; a < sd.out_rank; ++a) {
      int32_t l = sd.out_labels[a];
      bool inA = in_stride[0].count(l) > 0, inB = in_stride[1].count(l) > 0;
      GemmMode m{extent[l], stride_of(0, l), stride_of(1, l), c_stride[l], 0};
      if (inA && inB) { push(g.batch, g.n_batch, m); ls.g_n_batch++; ls.batch_labels.push_back(l); }
      else if (inA) { push(g.m, g.n_m, m); ls.gm.emplace_back(l, m); ls.row_labels.push_back(l); }
      else { push(g.n, g.n_n, m); ls.gn.emplace_back(l, m); ls.col_labels.push_back(l); }
    }
    for (auto& kv : extent) {
      if (out_set.count(kv.first)) continue;
      int32_t l = kv.first;
      GemmMode m{kv.second, stride_of(0, l), stride_of(1, l), 0, 0};
      push(g.k, g.n_k, m);
      ls.gk.emplace_back(l, m);
    }
    if (ok) {
      fuse_gemm_group(g.batch, g.n_batch, true);
      fuse_gemm_group(g.m, g.n_m, true);
      fuse_gemm_group(g.n, g.n_n, true);
      fuse_gemm_group(g.k, g.n_k, false);
      prefer_unit_stride_innermost(g.m, g.n_m, g.k, g.n_k, false);
      prefer_unit_stride_innermost(g.n, g.n_n, g.k, g.n_k, true);
      g.B = g.M = g.N = g.K = 1;
      for (int i = 0; i < g.n_batch; ++i) g.B *= g.batch[i].extent;
      for (int i = 0; i < g.n_m; ++i) g.M *= g.m[i].extent;
      for (int i = 0; i < g.n_n; ++i) g.N *= g.n[i].extent;
      for (int i = 0; i < g.n_k; ++i) g.K *= g.k[i].extent;
      ls.gemm = g;
      ls.has_gemm = true;
    }
  }
  return CTN_OK;
}

// Consumer fusion (float32, tcgen05 path): a GEMM step X[m,(r,n')] = A*B whose result is
// consumed ONLY by the next step out[m,n'] = sum_r W[m,r] * X[m,(r,n')] (the hyperedge step
// "bps,bp->bs" of a batched MPS sweep, cfg 3) is merged with it: the GEMM epilogue applies the
// weights and sums over r, X is never written.  Stabilisation stays exact up to the gauge:
// the kernel still measures sum|X| and sum|out| and flags CTN_RESULT_DEGENERATE if either of the
// reference's two rescale decisions (contractn/einsum.py:102) would have been "skip".
static bool try_fuse_consumer(Plan& p, LoweredStep& cur, const LoweredStep& nxt) {
  if (p.dtype != CTN_F32 || !cur.has_gemm || cur.tiny || cur.g_n_batch != 0) return false;
  if (cur.in_slots.size() != 2 || nxt.in_slots.size() != 2 || nxt.elementwise) return false;
  int xi = nxt.in_slots[0] == cur.out_slot ? 0 : nxt.in_slots[1] == cur.out_slot ? 1 : -1;
  if (xi < 0 || nxt.in_slots[1 - xi] == cur.out_slot) return false;
  const int wi = 1 - xi;
  const std::vector<int32_t>& xl = nxt.in_labels[xi];
  const std::vector<int32_t>& wl = nxt.in_labels[wi];
  if (xl != cur.out_labels) return false;
  std::set<int32_t> xset(xl.begin(), xl.end()), wset(wl.begin(), wl.end()), oset(nxt.out_labels.begin(), nxt.out_labels.end());
  if (xset.size() != xl.size() || wset.size() != wl.size()) return false;  // no diagonals
  // exactly one summed label r, carried by W and by the GEMM's N side
  std::vector<int32_t> summed;
  for (auto l : xset) if (!oset.count(l)) summed.push_back(l);
  for (auto l : wset) if (!xset.count(l)) return false;  // W must not bring new labels
  if (summed.size() != 1 || oset.size() != xset.size() - 1) return false;
  const int32_t r = summed[0];
  if (!wset.count(r)) return false;
  auto find = [](const std::vector<std::pair<int32_t, GemmMode>>& v, int32_t l) -> const GemmMode* {
    for (auto& kv : v) if (kv.first == l) return &kv.second;
    return nullptr;
  };
  // the summed label may sit on either side of the product: C = A*B is also C^T = B^T * A^T,
  // which in the strided formulation is just exchanging the operands and their stride fields
  std::vector<int32_t> ab = cur.in_slots;
  std::vector<std::pair<int32_t, GemmMode>> gm = cur.gm, gn = cur.gn, gk = cur.gk;
  if (!find(gn, r) && find(gm, r)) {
    std::swap(ab[0], ab[1]);
    std::swap(gm, gn);
    for (auto* grp : {&gm, &gn, &gk})
      for (auto& kv : *grp) std::swap(kv.second.sa, kv.second.sb);
  }
  const GemmMode* rmode = find(gn, r);
  if (!rmode || (rmode->extent != 2 && rmode->extent != 4)) return false;
  for (auto l : wl) if (l != r && !find(gm, l)) return false;  // weights vary with rows only
  // strides of W and of the fused output
  const Slot& W = p.slots[nxt.in_slots[wi]];
  std::map<int32_t, int64_t> sw, sc;
  for (size_t a = 0; a < wl.size(); ++a) {
    const GemmMode* wmode = wl[a] == r ? rmode : find(gm, wl[a]);
    if (W.extents[a] != wmode->extent) return false;  // no broadcasting in the fused form
    sw[wl[a]] = W.strides[a];
  }
  const Slot& O = p.slots[nxt.out_slot];
  for (size_t a = 0; a < nxt.out_labels.size(); ++a) sc[nxt.out_labels[a]] = O.strides[a];
  GemmDev g{};
  for (auto& kv : gm) {
    if (g.n_m >= kGemmMaxModes) return false;
    GemmMode m = kv.second;
    m.sc = sc[kv.first];
    m.sw = sw.count(kv.first) ? sw[kv.first] : 0;
    g.m[g.n_m++] = m;
  }
  for (auto& kv : gn) {
    if (kv.first == r) continue;
    if (g.n_n >= kGemmMaxModes) return false;
    GemmMode m = kv.second;
    m.sc = sc[kv.first];
    g.n[g.n_n++] = m;
  }
  for (auto& kv : gk) {
    if (g.n_k >= kGemmMaxModes) return false;
    g.k[g.n_k++] = kv.second;
  }
  fuse_gemm_group(g.m, g.n_m, true);
  fuse_gemm_group(g.n, g.n_n, true);
  fuse_gemm_group(g.k, g.n_k, false);
  g.B = g.M = g.N = g.K = 1;
  for (int i = 0; i < g.n_m; ++i) g.M *= g.m[i].extent;
  for (int i = 0; i < g.n_n; ++i) g.N *= g.n[i].extent;
  for (int i = 0; i < g.n_k; ++i) g.K *= g.k[i].extent;
  if (g.M < 16 || g.N < 16 || g.K < 8 || g.n_k > 4) return false;
  g.fuse_pr = (int32_t)rmode->extent;
  g.fuse_sb = rmode->sb;
  g.fuse_sw = sw[r];
  // merge
  cur.row_labels.clear();
  cur.col_labels.clear();
  cur.batch_labels.clear();
  for (auto& kv : gm) cur.row_labels.push_back(kv.first);
  for (auto& kv : gn) if (kv.first != r) cur.col_labels.push_back(kv.first);
  cur.desc_step = nxt.desc_step;
  cur.overridden = nxt.overridden;
  cur.in_slots = ab;
  cur.in_slots.push_back(nxt.in_slots[wi]);
  cur.out_slot = nxt.out_slot;
  cur.out_numel = nxt.out_numel;
  cur.flops += nxt.flops;
  cur.gemm = g;
  cur.fused_consumer = true;
  return true;
}

static void layout_state(Plan& p) {
  StateLayout& s = p.state;
  size_t n = p.slots.size();
  size_t off = 0;
  s.hdr = off; off += align_up(sizeof(StateHdr), 16);
  s.inv_kappa = off; off += n * sizeof(double);
  s.kappa = off; off += n * sizeof(double);
  s.ptrs = off; off += n * sizeof(void*);
  s.tickets = off; off += align_up((size_t)std::max(1, p.n_tickets) * sizeof(uint32_t), 16);
  s.init_bytes = off;
  s.partials = off; off += 4 * (size_t)kMaxPartials * sizeof(double);
  s.partial_exp = off; off += 2 * (size_t)kMaxPartials * sizeof(int32_t);
  s.total = align_up(off, 256);
}

// One compilation pass.  `ov` (may be null) fixes the physical layout of some intermediates; `want`
// (may be null) receives the layouts this pass would like to see in a second pass: row-index-fastest
// ("M-fast") results for the tcgen05 kernel, whose accumulator rows live in TMEM lanes = threads, so
// that its epilogue stores and the next step's A-operand loads are coalesced with no transposition.
// `violated` is set when a step whose output layout was overridden did not end up on that kernel
// (the other kernels write their output flat in label order): the caller then keeps the first pass.
static int build_plan(const ctn_plan_desc& d, Plan& p, const LayoutOverrides* ov, LayoutOverrides* want, bool* violated) {
  if (d.abi_version != CTN_ABI_VERSION) return CTN_ERR_INVALID;
  if (d.dtype != CTN_F32 && d.dtype != CTN_F64) return CTN_ERR_UNSUPPORTED;
  if (d.n_operands < 1 || d.n_steps < 1 || !d.operand_ranks || !d.steps) return CTN_ERR_INVALID;
  p.dtype = d.dtype;
  p.elem = d.dtype == CTN_F32 ? 4 : 8;
  p.flags = d.flags;
  p.n_operands = d.n_operands;
  const int64_t* ext = d.operand_extents;
  const int64_t* str = d.operand_strides;
  std::vector<int32_t> live;
  for (int i = 0; i < d.n_operands; ++i) {
    Slot s;
    int r = d.operand_ranks[i];
    if (r < 0 || (r > 0 && !ext)) return CTN_ERR_INVALID;
    s.extents.assign(ext, ext + r);
    ext += r;
    for (auto e : s.extents) {
      if (e < 0) return CTN_ERR_INVALID;
      s.numel *= e;
    }
    if (str) {
      s.strides.assign(str, str + r);
      str += r;
      for (auto v : s.strides)
        if (v < 0) return CTN_ERR_UNSUPPORTED;
    } else {
      s.strides.assign(r, 1);
      for (int a = r - 2; a >= 0; --a) s.strides[a] = s.strides[a + 1] * s.extents[a + 1];
    }
    s.operand = i;
    live.push_back(i);
    p.slots.push_back(s);
  }

  // lower every step, fusing Hadamard runs on the way
  std::vector<LoweredStep> lowered;
  const bool fuse = !(d.flags & CTN_PLAN_EXACT_STEPS);
  for (int t = 0; t < d.n_steps; ++t) {
    LoweredStep ls;
    ls.desc_step = t;
    const std::vector<int32_t>* phys = nullptr;
    if (ov && t + 1 < d.n_steps) {
      auto it = ov->find(t);
      if (it != ov->end()) phys = &it->second;
    }
    int rc = lower_step(p, d.steps[t], live, ls, phys);
    if (rc != CTN_OK) return rc;
    if (p.slots[ls.out_slot].numel == 0) return CTN_ERR_UNSUPPORTED;  // empty tensors
    p.flops += ls.flops;
    if (fuse && ls.elementwise && !lowered.empty() && lowered.back().elementwise &&
        !lowered.back().tiny && !ls.tiny) {
      // does this step consume the previous step's output?
      LoweredStep& prev = lowered.back();
      auto it = std::find(ls.in_slots.begin(), ls.in_slots.end(), prev.out_slot);
      size_t merged = prev.in_slots.size() + ls.in_slots.size() - 1;
      if (it != ls.in_slots.end() && merged <= (size_t)kChainMaxIn) {
        for (auto s : ls.in_slots)
          if (s != prev.out_slot) prev.in_slots.push_back(s);
        prev.out_slot = ls.out_slot;
        prev.desc_step = ls.desc_step;
        prev.flops += ls.flops;
        continue;
      }
    }
    lowered.push_back(std::move(ls));
  }
  if (fuse && !(d.flags & CTN_PLAN_NO_GEMM)) {
    for (size_t t = 0; t + 1 < lowered.size(); ++t)
      if (try_fuse_consumer(p, lowered[t], lowered[t + 1])) lowered.erase(lowered.begin() + t + 1);
  }
  if (live.size() != 1) return CTN_ERR_INVALID;  // the list must reduce to one tensor
  p.final_slot = live[0];
  p.slots[p.final_slot].is_final = true;

  // emit device steps + launches, placing intermediates in the arena
  Arena arena;
  std::map<int32_t, int32_t> producer;  // slot -> index of the step that writes it
  std::vector<std::pair<size_t, size_t>> deferred;  // buffers read inside the current interpreter launch
  const bool allow_interp = !(d.flags & CTN_PLAN_NO_INTERPRETER);
  const bool allow_gemm = !(d.flags & CTN_PLAN_NO_GEMM);
  static const bool tf32_v1 = getenv("CTN_TF32_V1") != nullptr;        // round-1 kernel (A/B comparisons)
  static const bool no_mfast = getenv("CTN_TF32_NO_MFAST") != nullptr;  // keep label-order intermediates
  for (size_t t = 0; t < lowered.size(); ++t) {
    LoweredStep& ls = lowered[t];
    StepDev sd{};
    sd.n_in = (int32_t)ls.in_slots.size();
    sd.out_slot = ls.out_slot;
    sd.out_numel = ls.out_numel;
    sd.k_numel = ls.k_numel;
    sd.gemm_off = -1;
    sd.is_final = ls.out_slot == p.final_slot;
    sd.ticket = p.n_tickets++;
    int64_t bytes = p.slots[ls.out_slot].numel;
    for (auto s : ls.in_slots) bytes += p.slots[s].numel;
    bytes *= (int64_t)p.elem;
    const bool big_elementwise = ls.elementwise && !(ls.tiny && allow_interp);
    if (big_elementwise) {
      sd.kind = kChain;
      sd.slot_off = (int32_t)p.slot_list.size();
      p.slot_list.insert(p.slot_list.end(), ls.in_slots.begin(), ls.in_slots.end());
      sd.tracked = sd.n_in > 2;
    } else {
      for (int i = 0; i < sd.n_in; ++i) sd.in_slot[i] = ls.in_slots[i];
      if (ls.fused_consumer) {
        sd.kind = kGemm;
        sd.gemm_off = (int32_t)p.gemms.size();
        const int32_t pr = ls.gemm.fuse_pr;
        const int64_t fsb = ls.gemm.fuse_sb, fsw = ls.gemm.fuse_sw;
        if (tf32_v1) gemm_tf32_v1_configure(p, ls.gemm);
        else gemm_tf32_configure(p, ls.gemm);
        ls.gemm.fuse_pr = pr; ls.gemm.fuse_sb = fsb; ls.gemm.fuse_sw = fsw;
        if (fsb & 3) ls.gemm.lb = 2;  // r-stride breaks 16-byte alignment of B chunks: element-wise staging
        p.gemms.push_back(ls.gemm);
      }
      sd.n_out_modes = (int32_t)ls.out_modes.size();
      sd.n_k_modes = (int32_t)ls.k_modes.size();
      sd.mode_off = (int32_t)p.modes.size();
      p.modes.insert(p.modes.end(), ls.out_modes.begin(), ls.out_modes.end());
      p.modes.insert(p.modes.end(), ls.k_modes.begin(), ls.k_modes.end());
      if (!ls.fused_consumer) sd.kind = (ls.tiny && allow_interp) ? kInterp : kGeneric;
      if (sd.kind == kInterp && sd.n_in == 2) {
        // pairwise step whose offsets fit 32 bits: the interpreter's short path (sd.tracked doubles as its flag)
        int64_t span[2] = {0, 0};
        for (int32_t m = sd.mode_off; m < (int32_t)p.modes.size(); ++m)
          for (int i = 0; i < 2; ++i) span[i] += (p.modes[m].extent - 1) * p.modes[m].stride[i];
        sd.tracked = span[0] < (1ll << 30) && span[1] < (1ll << 30);
      }
      // few outputs, each a long sum (DOT class, <a|b>): grid-wide reduction kernel, not a 1x1 GEMM
      const bool dot_like = sd.kind == kGeneric && sd.n_in <= 4 && ls.out_numel <= kDotMaxOut && ls.k_numel >= kDotMinK;
      if (dot_like) sd.tracked = 1;
      if (sd.kind == kGeneric && !dot_like && allow_gemm && ls.has_gemm) {
        if (int sk = skinny_config(p, sd, ls.gemm)) {
          sd.kind = kGemm;
          sd.gemm_off = (int32_t)p.gemms.size();
          ls.gemm.cfg = sk;
          ls.gemm.splits = 1;
          ls.gemm.la = ls.gemm.lb = 0;
          ls.gemm.scratch_off = 0;
          p.gemms.push_back(ls.gemm);
        } else if (gemm_supported(p, sd, ls.gemm)) {
          sd.kind = kGemm;
          sd.gemm_off = (int32_t)p.gemms.size();
          gemm_configure(p, ls.gemm);
          p.gemms.push_back(ls.gemm);
        } else if (gemm_tf32_supported(p, sd, ls.gemm)) {
          sd.kind = kGemm;
          sd.gemm_off = (int32_t)p.gemms.size();
          if (tf32_v1) gemm_tf32_v1_configure(p, ls.gemm);
          else gemm_tf32_configure(p, ls.gemm);
          p.gemms.push_back(ls.gemm);
        }
      }
    }
    {
      const bool on_tc = sd.kind == kGemm && p.dtype == CTN_F32 && p.gemms[sd.gemm_off].cfg == 8;
      if (ls.overridden && !on_tc && violated) *violated = true;
      if (want && on_tc && !no_mfast && !sd.is_final && !ls.row_labels.empty() && p.gemms[sd.gemm_off].M >= 32) {
        std::vector<int32_t> order = ls.batch_labels;
        order.insert(order.end(), ls.col_labels.begin(), ls.col_labels.end());
        order.insert(order.end(), ls.row_labels.begin(), ls.row_labels.end());
        if (order.size() == p.slots[ls.out_slot].labels.size() && order != p.slots[ls.out_slot].labels)
          (*want)[ls.desc_step] = order;
      }
    }
    // place the output, then release the inputs (never alias within a step)
    Slot& out = p.slots[ls.out_slot];
    if (!out.is_final) out.ws_offset = (int64_t)arena.alloc((size_t)out.numel * p.elem);
    size_t scratch_bytes = 0;
    if (sd.kind == kGemm) {
      scratch_bytes = gemm_scratch_bytes(p.gemms[sd.gemm_off]);
      if (scratch_bytes) {
        sd.scratch_off = (int64_t)arena.alloc(scratch_bytes);
        p.gemms[sd.gemm_off].scratch_off = sd.scratch_off;
      }
    }
    if (sd.kind == kGeneric && sd.tracked) {
      // dot kernel: one partial per (output, CTA)
      scratch_bytes = (size_t)kDotMaxOut * kMaxPartials * sizeof(double);
      sd.scratch_off = (int64_t)arena.alloc(scratch_bytes);
    }
    if (sd.kind == kChain && sd.tracked) {
      // one exponent per tile of the chain output (tile = 1024 f64 / 2048 f32 at most;
      // sized for the smallest tile the kernel variants use)
      scratch_bytes = (size_t)((out.numel + 511) / 512 + 1) * sizeof(int32_t);
      sd.scratch_off = (int64_t)arena.alloc(scratch_bytes);
    }
    // The interpreter runs independent steps of one launch concurrently, so a
    // buffer read inside an interpreter launch must not be handed to a later
    // step of the same launch: hold those releases back until the launch ends.
    if (sd.kind != kInterp) {
      for (auto& r : deferred) arena.release(r.first, r.second);
      deferred.clear();
    }
    for (auto s : ls.in_slots) {
      Slot& in = p.slots[s];
      if (in.operand < 0 && in.ws_offset >= 0) {
        if (sd.kind == kInterp) deferred.emplace_back((size_t)in.ws_offset, (size_t)in.numel * p.elem);
        else arena.release((size_t)in.ws_offset, (size_t)in.numel * p.elem);
      }
    }
    if (scratch_bytes) arena.release((size_t)sd.scratch_off, scratch_bytes);
    int32_t idx = (int32_t)p.steps.size();
    const bool join = sd.kind == kInterp && !p.launches.empty() && p.launches.back().kind == kInterp &&
                      idx - p.launches.back().first_step < kInterpMaxSteps;
    if (sd.kind == kInterp) {
      // dependencies inside the launch: the interpreter runs independent steps on different warps
      const int32_t first = join ? p.launches.back().first_step : idx;
      for (int i = 0; i < kMaxIn; ++i) {
        sd.dep[i] = -1;
        if (i < sd.n_in) {
          auto it = producer.find(sd.in_slot[i]);
          if (it != producer.end() && it->second >= first) sd.dep[i] = it->second - first;
          else if (p.slots[sd.in_slot[i]].operand >= 0) sd.dep[i] = -2;  // caller's operand: never rescaled
        }
      }
    }
    producer[sd.out_slot] = idx;
    p.steps.push_back(sd);
    if (join) {
      p.launches.back().last_step = idx;
      p.launches.back().bytes += bytes;
      p.launches.back().flops += ls.flops;
    } else {
      p.launches.push_back(Launch{sd.kind, idx, idx, bytes, ls.flops, 0});
      if (sd.kind == kChain && sd.tracked)
        p.launches.push_back(Launch{kMaterialize, idx, idx, 2 * out.numel * (int64_t)p.elem, 0.0, 0});
      if (sd.kind == kGemm && p.gemms[sd.gemm_off].splits > 1)
        p.launches.push_back(Launch{kGemmReduce, idx, idx, (int64_t)scratch_bytes, 0.0, 0});
    }
  }
  // the last launch must leave the canonical (T/mean|T|, log-scale) pair:
  // interpreter and tracked-chain launches do it themselves, others need a pass
  const Launch& last = p.launches.back();
  const StepDev& last_step = p.steps[last.last_step];
  (void)last_step;
  bool self_final = last.kind == kInterp || last.kind == kMaterialize;
  p.raw_final = !self_final && (p.flags & CTN_PLAN_RAW_FINAL);
  if (!self_final && !p.raw_final) {
    int64_t bytes = 2 * p.slots[p.final_slot].numel * (int64_t)p.elem;
    p.launches.push_back(Launch{kFinalize, last.last_step, last.last_step, bytes, 0.0, 0});
  }
  // Interpreter launches whose intermediates fit keep them in shared memory (every step's
  // output gets its own place in a per-launch pool: no reuse, so the concurrently running
  // warps cannot clobber one another); outputs are ALSO written to global memory for later launches.
  for (auto& l : p.launches) {
    if (l.kind != kInterp) continue;
    size_t total = 0;
    for (int32_t i = l.first_step; i <= l.last_step; ++i) total += align_up((size_t)p.steps[i].out_numel * p.elem, 16);
    const bool pooled = total <= (size_t)kInterpPoolBytes;
    size_t off = 0;
    for (int32_t i = l.first_step; i <= l.last_step; ++i) {
      StepDev& sdv = p.steps[i];
      sdv.smem_out = pooled ? (int32_t)off : -1;
      off += align_up((size_t)sdv.out_numel * p.elem, 16);
      for (int k = 0; k < kMaxIn; ++k) {
        sdv.smem_in[k] = -1;
        if (pooled && k < sdv.n_in && sdv.dep[k] >= 0) sdv.smem_in[k] = p.steps[l.first_step + sdv.dep[k]].smem_out;
      }
    }
    // Inputs that come from OUTSIDE the launch (operands, earlier intermediates) are staged into the
    // pool once at kernel start when they are small: a dependent chain then never waits for an L2
    // round trip.  The whole address span is copied, so the step's element strides stay valid.
    l.pre_off = (int32_t)p.interp_pre.size();
    if (pooled) {
      std::map<int32_t, int32_t> staged;  // slot -> pool offset (shared operands are staged once)
      for (int32_t i = l.first_step; i <= l.last_step; ++i) {
        StepDev& sdv = p.steps[i];
        for (int k = 0; k < sdv.n_in; ++k) {
          if (sdv.dep[k] >= 0) continue;
          const Slot& sl = p.slots[sdv.in_slot[k]];
          auto it = staged.find(sdv.in_slot[k]);
          if (it != staged.end()) { sdv.smem_in[k] = it->second; continue; }
          if (sl.numel == 0) continue;
          int64_t span = 1;
          for (size_t a = 0; a < sl.extents.size(); ++a) span += (sl.extents[a] - 1) * sl.strides[a];
          const size_t bytes = align_up((size_t)span * p.elem, 16);
          if (span > 512 || total + bytes > (size_t)kInterpPoolBytes) continue;
          staged[sdv.in_slot[k]] = (int32_t)total;
          sdv.smem_in[k] = (int32_t)total;
          p.interp_pre.push_back(InterpPreload{sdv.in_slot[k], (int32_t)total, (int32_t)span, 0});
          total += bytes;
        }
      }
    }
    l.pre_n = (int32_t)p.interp_pre.size() - l.pre_off;
    l.smem_pool = pooled ? (int32_t)std::max<size_t>(total, 16) : 0;
    // packed records: descriptor + modes of a step in one 1 KB block the kernel fetches with cp.async
    static_assert(sizeof(StepDev) <= 256, "StepDev must fit the record header");
    l.rec_off = (int32_t)(p.interp_recs.size() / kInterpRecBytes);
    for (int32_t i = l.first_step; i <= l.last_step; ++i) {
      const StepDev& sdv = p.steps[i];
      size_t at = p.interp_recs.size();
      p.interp_recs.resize(at + kInterpRecBytes, 0);
      memcpy(p.interp_recs.data() + at, &sdv, sizeof(StepDev));
      const int nm = sdv.n_out_modes + sdv.n_k_modes;
      if (nm > 0) memcpy(p.interp_recs.data() + at + 256, &p.modes[sdv.mode_off], (size_t)nm * sizeof(ModeDev));
    }
  }
  p.arena_bytes = arena.top;
  layout_state(p);
  return CTN_OK;
}

static int query_device(Plan& p) {
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    // no device: the plan is still usable for introspection (CPU build checks)
    cudaGetLastError();
    p.device = -1;
    return CTN_OK;
  }
  CTN_CUDA(cudaGetDevice(&p.device));
  int sms = 0;
  CTN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p.device));
  p.sm_count = sms;
  return CTN_OK;
}

static int upload_plan(Plan& p) {
  if (p.device < 0) return CTN_OK;
  auto up = [](void** dst, const void* src, size_t bytes) -> cudaError_t {
    if (bytes == 0) { *dst = nullptr; return cudaSuccess; }
    cudaError_t e = cudaMalloc(dst, bytes);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  CTN_CUDA(up((void**)&p.d_steps, p.steps.data(), p.steps.size() * sizeof(StepDev)));
  CTN_CUDA(up((void**)&p.d_modes, p.modes.data(), p.modes.size() * sizeof(ModeDev)));
  CTN_CUDA(up((void**)&p.d_gemms, p.gemms.data(), p.gemms.size() * sizeof(GemmDev)));
  CTN_CUDA(up((void**)&p.d_slot_list, p.slot_list.data(), p.slot_list.size() * sizeof(int32_t)));
  CTN_CUDA(up((void**)&p.d_interp_recs, p.interp_recs.data(), p.interp_recs.size()));
  CTN_CUDA(up((void**)&p.d_interp_pre, p.interp_pre.data(), p.interp_pre.size() * sizeof(InterpPreload)));
  return CTN_OK;
}

}  // namespace ctn

using namespace ctn;

struct ctn_plan {
  Plan p;
};

extern "C" {

int ctn_abi_version(void) { return CTN_ABI_VERSION; }

int ctn_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ++ok;
  }
  return ok;
}

const char* ctn_status_string(int status) {
  switch (status) {
    case CTN_OK: return "ok";
    case CTN_ERR_INVALID: return "invalid plan description";
    case CTN_ERR_SHAPE: return "extents of a shared index disagree";
    case CTN_ERR_UNSUPPORTED: return "contraction not supported by the B200 engine";
    case CTN_ERR_CUDA: return "CUDA error";
    case CTN_ERR_NOMEM: return "out of memory";
    case CTN_ERR_NO_DEVICE: return "no sm_100 CUDA device available";
    case CTN_ERR_INTERNAL: return "internal kernel error (pipeline time-out)";
    default: return "unknown status";
  }
}

const char* ctn_last_cuda_error(void) { return g_cuda_error.c_str(); }

int ctn_plan_create(const ctn_plan_desc* desc, ctn_plan** plan_out) {
  if (!desc || !plan_out) return CTN_ERR_INVALID;
  ctn_plan* h = new (std::nothrow) ctn_plan();
  if (!h) return CTN_ERR_NOMEM;
  int rc;
  try {
    rc = query_device(h->p);
    LayoutOverrides want;
    if (rc == CTN_OK) rc = build_plan(*desc, h->p, nullptr, &want, nullptr);
    if (rc == CTN_OK && !want.empty()) {
      // second pass with the layouts the tensor-core steps asked for
      ctn_plan* h2 = new (std::nothrow) ctn_plan();
      if (h2) {
        bool violated = false;
        int rc2 = query_device(h2->p);
        if (rc2 == CTN_OK) rc2 = build_plan(*desc, h2->p, &want, nullptr, &violated);
        if (rc2 == CTN_OK && !violated) std::swap(h, h2);
        ctn_plan_destroy(h2);
      }
    }
    if (rc == CTN_OK) rc = upload_plan(h->p);
  } catch (const std::bad_alloc&) {
    rc = CTN_ERR_NOMEM;
  } catch (...) {
    rc = CTN_ERR_INVALID;
  }
  if (rc != CTN_OK) {
    ctn_plan_destroy(h);
    return rc;
  }
  *plan_out = h;
  return CTN_OK;
}

void ctn_plan_destroy(ctn_plan* plan) {
  if (!plan) return;
  Plan& p = plan->p;
  if (p.device >= 0) {
    cudaFree(p.d_steps);
    cudaFree(p.d_modes);
    cudaFree(p.d_gemms);
    cudaFree(p.d_slot_list);
    cudaFree(p.d_interp_recs);
    cudaFree(p.d_interp_pre);
    cudaFree(p.h_ws);
    cudaFree(p.h_out);
    cudaFree(p.h_log);
    cudaFree(p.h_pack);
    std::set<void*> freed;
    for (void* q : p.h_operands)
      if (q && freed.insert(q).second) cudaFree(q);
  }
  delete plan;
}

size_t ctn_plan_workspace_bytes(const ctn_plan* plan) {
  if (!plan) return 0;
  return plan->p.state.total + align_up(plan->p.arena_bytes, 256) + 256;
}

int ctn_plan_output_rank(const ctn_plan* plan) {
  return plan ? (int)plan->p.slots[plan->p.final_slot].extents.size() : -1;
}

int ctn_plan_output_extents(const ctn_plan* plan, int64_t* extents) {
  if (!plan) return CTN_ERR_INVALID;
  const Slot& s = plan->p.slots[plan->p.final_slot];
  if (extents)
    for (size_t i = 0; i < s.extents.size(); ++i) extents[i] = s.extents[i];
  return CTN_OK;
}

int64_t ctn_plan_output_numel(const ctn_plan* plan) {
  return plan ? plan->p.slots[plan->p.final_slot].numel : -1;
}

int64_t ctn_plan_final_kappa_offset(const ctn_plan* plan) {
  if (!plan || !plan->p.raw_final) return -1;
  return (int64_t)(plan->p.state.kappa + (size_t)plan->p.final_slot * sizeof(double));
}

int ctn_plan_num_launches(const ctn_plan* plan) {
  if (!plan) return -1;
  return (int)plan->p.launches.size();
}

double ctn_plan_flops(const ctn_plan* plan) { return plan ? plan->p.flops : 0.0; }

double ctn_plan_step_bytes(const ctn_plan* plan) {
  if (!plan) return 0.0;
  double b = 0;
  for (auto& l : plan->p.launches)
    if (l.kind != kFinalize && l.kind != kMaterialize && l.kind != kGemmReduce) b += (double)l.bytes;
  return b;
}

int ctn_plan_describe(const ctn_plan* plan, char* buf, size_t n) {
  if (!plan || !buf || n == 0) return CTN_ERR_INVALID;
  const Plan& p = plan->p;
  std::string out;
  char line[512];
  for (size_t li = 0; li < p.launches.size(); ++li) {
    const Launch& l = p.launches[li];
    const StepDev& s = p.steps[l.first_step];
    static const char* names[] = {"generic", "chain", "interp", "gemm", "finalize", "materialize", "gemm_reduce"};
    snprintf(line, sizeof line, "%zu %s steps=%d..%d", li, names[l.kind], l.first_step, l.last_step);
    out += line;
    if (l.kind == kGemm) {
      const GemmDev& g = p.gemms[s.gemm_off];
      snprintf(line, sizeof line, " cfg=%d la=%d lb=%d B=%lld M=%lld N=%lld K=%lld fuse_pr=%d cvec=%d", g.cfg, g.la, g.lb,
               (long long)g.B, (long long)g.M, (long long)g.N, (long long)g.K, g.fuse_pr, g.fuse_pad & 1);
      out += line;
    }
    if (l.kind != kFinalize && l.kind != kMaterialize && l.kind != kGemmReduce && l.kind != kInterp) {
      const Slot& o = p.slots[s.out_slot];
      out += " out_strides=";
      for (size_t a = 0; a < o.strides.size(); ++a) {
        snprintf(line, sizeof line, "%s%lld", a ? "," : "", (long long)o.strides[a]);
        out += line;
      }
    }
    out += "\n";
  }
  snprintf(buf, n, "%s", out.c_str());
  return CTN_OK;
}

int ctn_plan_launch_kind(const ctn_plan* plan, int i) {
  if (!plan || i < 0 || i >= (int)plan->p.launches.size()) return -1;
  return plan->p.launches[i].kind;
}

static int execute_impl(const ctn_plan* plan, const void* const* operands, void* out, double* log_scale_out,
                        int32_t* flags_out, void* workspace, int out_mode, void* stream_, cudaEvent_t* events) {
  if (!plan || !operands || !out || !log_scale_out || !workspace) return CTN_ERR_INVALID;
  if (out_mode != CTN_OUT_SPLIT && out_mode != CTN_OUT_DENSE) return CTN_ERR_INVALID;
  const Plan& p = plan->p;
  if (p.device < 0) return CTN_ERR_NO_DEVICE;
  {
    // the plan's descriptor tables live on p.device: launching from another current device would
    // hand kernels on that device pointers they cannot dereference
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != p.device) return CTN_ERR_INVALID;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  char* ws = (char*)workspace;
  if (((uintptr_t)ws & 255) != 0) return CTN_ERR_INVALID;

  // slot -> device pointer table; the state block is (re)initialised by a small kernel
  const StateLayout& st = p.state;
  static thread_local std::vector<const void*> table;
  table.assign(p.slots.size(), nullptr);
  const void** ptrs = table.data();
  char* arena = ws + st.total;
  for (size_t i = 0; i < p.slots.size(); ++i) {
    const Slot& s = p.slots[i];
    if (s.operand >= 0) ptrs[i] = operands[s.operand];
    else if (s.is_final) ptrs[i] = out;
    else if (s.ws_offset >= 0) ptrs[i] = arena + s.ws_offset;
    else ptrs[i] = nullptr;  // fused away
    if (s.operand >= 0 && !ptrs[i]) return CTN_ERR_INVALID;
  }
  {
    cudaError_t e = launch_init_state(p, ws, stream, ptrs);
    if (e != cudaSuccess) {
      set_cuda_error(e, "init_state launch");
      return CTN_ERR_CUDA;
    }
  }

  ExecCtx ctx{&p, ws, stream, out_mode, log_scale_out, flags_out};
  auto chain_vec_ok = [&](const StepDev& s) {
    bool ok = ((uintptr_t)ptrs[s.out_slot] & 15) == 0;
    for (int i = 0; i < s.n_in; ++i) ok = ok && (((uintptr_t)ptrs[p.slot_list[s.slot_off + i]] & 15) == 0);
    return ok;
  };
  for (size_t li = 0; li < p.launches.size(); ++li) {
    const Launch& l = p.launches[li];
    if (events) CTN_CUDA(cudaEventRecord(events[li], stream));
    cudaError_t e = cudaSuccess;
    switch (l.kind) {
      case kGeneric: e = launch_generic(ctx, l); break;
      case kChain: e = launch_chain(ctx, l, chain_vec_ok(p.steps[l.first_step])); break;
      case kMaterialize: e = launch_materialize(ctx, l, chain_vec_ok(p.steps[l.first_step])); break;
      case kInterp: e = launch_interp(ctx, l); break;
      case kGemm: {
        const StepDev& gs = p.steps[l.first_step];
        const bool a_al = ((uintptr_t)ptrs[gs.in_slot[0]] & 15) == 0, b_al = ((uintptr_t)ptrs[gs.in_slot[1]] & 15) == 0;
        const int cfg = p.gemms[gs.gemm_off].cfg;
        if (cfg == 3 || cfg == 4) e = launch_skinny(ctx, l);
        else if (p.dtype != CTN_F32) e = launch_gemm(ctx, l, a_al, b_al);
        else if (cfg == 8) e = launch_gemm_tf32(ctx, l, a_al, b_al, ((uintptr_t)ptrs[gs.out_slot] & 15) == 0, ptrs[gs.in_slot[1]], ptrs[gs.in_slot[0]]);
        else e = launch_gemm_tf32_v1(ctx, l, a_al, b_al);
        break;
      }
      case kGemmReduce: e = launch_gemm_reduce(ctx, l); break;
      case kFinalize: e = launch_finalize(ctx, l); break;
      default: return CTN_ERR_INVALID;
    }
    if (e != cudaSuccess) {
      set_cuda_error(e, "kernel launch");
      return CTN_ERR_CUDA;
    }
  }
  if (events) CTN_CUDA(cudaEventRecord(events[p.launches.size()], stream));
  return CTN_OK;
}

int ctn_plan_execute(const ctn_plan* plan, const void* const* operands, void* out, double* log_scale_out,
                     int32_t* flags_out, void* workspace, int out_mode, void* stream_) {
  return execute_impl(plan, operands, out, log_scale_out, flags_out, workspace, out_mode, stream_, nullptr);
}

int ctn_plan_execute_timed(const ctn_plan* plan, const void* const* operands, void* out, double* log_scale_out,
                           int32_t* flags_out, void* workspace, int out_mode, void* stream_, float* launch_ms) {
  if (!plan || !launch_ms) return CTN_ERR_INVALID;
  const size_t n = plan->p.launches.size();
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) CTN_CUDA(cudaEventCreate(&e));
  int rc = execute_impl(plan, operands, out, log_scale_out, flags_out, workspace, out_mode, stream_, ev.data());
  if (rc == CTN_OK) {
    cudaError_t e = cudaEventSynchronize(ev[n]);
    if (e != cudaSuccess) { set_cuda_error(e, "cudaEventSynchronize"); rc = CTN_ERR_CUDA; }
    for (size_t i = 0; rc == CTN_OK && i < n; ++i) cudaEventElapsedTime(&launch_ms[i], ev[i], ev[i + 1]);
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return rc;
}

int ctn_merge_batch_shards_peer(void* data, int64_t numel, int dtype, const void* workspace, double* const* mailboxes,
                                int world, int rank, uint64_t seq, double* log_scale_out, void* stream) {
  static_assert(CTN_MAILBOX_WORDS == 2 * 4 * kMailboxMaxWorld, "header constant out of date");
  if (!data || !workspace || !mailboxes || !log_scale_out || numel < 0) return CTN_ERR_INVALID;
  if (world < 1 || world > kMailboxMaxWorld || rank < 0 || rank >= world || seq == 0) return CTN_ERR_INVALID;
  if (dtype != CTN_F32 && dtype != CTN_F64) return CTN_ERR_UNSUPPORTED;
  for (int i = 0; i < world; ++i)
    if (!mailboxes[i]) return CTN_ERR_INVALID;
  // the state header sits at the start of every plan's workspace (StateLayout::hdr == 0)
  cudaError_t e = launch_merge_shards_peer(data, numel, dtype, (const StateHdr*)workspace, mailboxes, world, rank,
                                           (unsigned long long)seq, log_scale_out, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    set_cuda_error(e, "merge_shards_peer launch");
    return CTN_ERR_CUDA;
  }
  return CTN_OK;
}

int ctn_reshard_peer(const void* data, int64_t numel, int dtype, const void* workspace, int64_t kappa_offset,
                     double* const* mailboxes, double* const* done_boxes, void* const* recv, uint32_t* ticket,
                     int world, int rank, uint64_t seq, double* log_scale_out, void* stream) {
  if (!data || !workspace || !mailboxes || !done_boxes || !recv || !ticket || !log_scale_out || numel < 0)
    return CTN_ERR_INVALID;
  if (world < 1 || world > kMailboxMaxWorld || rank < 0 || rank >= world || seq == 0 || numel % world) return CTN_ERR_INVALID;
  if (dtype != CTN_F32 && dtype != CTN_F64) return CTN_ERR_UNSUPPORTED;
  for (int i = 0; i < world; ++i)
    if (!mailboxes[i] || !done_boxes[i] || !recv[i]) return CTN_ERR_INVALID;
  const double* kappa_raw = kappa_offset >= 0 ? (const double*)((const char*)workspace + kappa_offset) : nullptr;
  cudaError_t e = launch_reshard_peer(data, numel, dtype, (const StateHdr*)workspace, kappa_raw, mailboxes, done_boxes,
                                      recv, ticket, world, rank, (unsigned long long)seq, log_scale_out,
                                      (cudaStream_t)stream);
  if (e != cudaSuccess) {
    set_cuda_error(e, "reshard_peer launch");
    return CTN_ERR_CUDA;
  }
  return CTN_OK;
}

int ctn_merge_batch_shards(void* data, int64_t numel, int dtype, const double* table, int world, int rank,
                           double* log_scale_out, void* stream) {
  if (!data || !table || !log_scale_out || world < 1 || rank < 0 || rank >= world || numel < 0) return CTN_ERR_INVALID;
  if (dtype != CTN_F32 && dtype != CTN_F64) return CTN_ERR_UNSUPPORTED;
  cudaError_t e = launch_merge_shards(data, numel, dtype, table, world, rank, log_scale_out, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    set_cuda_error(e, "merge_shards launch");
    return CTN_ERR_CUDA;
  }
  return CTN_OK;
}

int ctn_plan_execute_host(ctn_plan* plan, const void* const* operands_host, void* out_host,
                          double* log_scale_host, int32_t* flags_host, int out_mode) {
  if (!plan || !operands_host || !out_host || !log_scale_host) return CTN_ERR_INVALID;
  Plan& p = plan->p;
  if (p.device < 0) return CTN_ERR_NO_DEVICE;
  // the staging buffers below belong to the plan: one host-path execution at a time
  std::lock_guard<std::mutex> guard(p.host_mu);
  struct DeviceGuard {
    int prev = -1;
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
  } dev_guard;
  {
    int cur = -1;
    CTN_CUDA(cudaGetDevice(&cur));
    if (cur != p.device) {
      dev_guard.prev = cur;
      CTN_CUDA(cudaSetDevice(p.device));
    }
  }
  size_t ws_bytes = ctn_plan_workspace_bytes(plan);
  if (!p.h_ws) {
    CTN_CUDA(cudaMalloc(&p.h_ws, ws_bytes));
    p.h_ws_bytes = ws_bytes;
    CTN_CUDA(cudaMalloc(&p.h_out, std::max<size_t>(16, (size_t)p.slots[p.final_slot].numel * p.elem)));
    CTN_CUDA(cudaMalloc((void**)&p.h_log, 16));
    p.h_operands.assign(p.n_operands, nullptr);
    p.h_operand_bytes.assign(p.n_operands, 0);
    for (int i = 0; i < p.n_operands; ++i) {
      const Slot& s = p.slots[i];
      size_t span = 1;
      for (size_t a = 0; a < s.extents.size(); ++a) span += (size_t)(s.extents[a] - 1) * (size_t)s.strides[a];
      p.h_operand_bytes[i] = span * p.elem;
    }
  }
  // Small operands (a 1000-matrix chain has 1001 of them) are gathered on the host and go up
  // in ONE copy; large ones are copied directly.  The same host array passed for several
  // operands (weight sharing) is uploaded once.
  constexpr size_t kPackLimit = 64 * 1024;
  if (p.h_pack_off.empty()) {
    p.h_pack_off.assign(p.n_operands, -1);
    size_t off = 0;
    for (int i = 0; i < p.n_operands; ++i) {
      if (p.h_operand_bytes[i] <= kPackLimit) {
        p.h_pack_off[i] = (int64_t)off;
        off += align_up(std::max<size_t>(p.h_operand_bytes[i], 1), 256);
      }
    }
    p.h_pack_host.assign(off, 0);
    if (off) CTN_CUDA(cudaMalloc(&p.h_pack, off));
  }
  std::map<const void*, int> first_use;
  std::vector<const void*> dev(p.n_operands);
  for (int i = 0; i < p.n_operands; ++i) {
    if (!operands_host[i]) return CTN_ERR_INVALID;
    auto it = first_use.find(operands_host[i]);
    if (it != first_use.end() && p.h_operand_bytes[it->second] >= p.h_operand_bytes[i]) {
      dev[i] = dev[it->second];
      continue;
    }
    first_use[operands_host[i]] = i;
    if (p.h_pack_off[i] >= 0) {
      memcpy(p.h_pack_host.data() + p.h_pack_off[i], operands_host[i], p.h_operand_bytes[i]);
      dev[i] = (const char*)p.h_pack + p.h_pack_off[i];
      continue;
    }
    if (!p.h_operands[i]) CTN_CUDA(cudaMalloc(&p.h_operands[i], std::max<size_t>(16, p.h_operand_bytes[i])));
    CTN_CUDA(cudaMemcpyAsync(p.h_operands[i], operands_host[i], p.h_operand_bytes[i], cudaMemcpyHostToDevice, 0));
    dev[i] = p.h_operands[i];
  }
  if (!p.h_pack_host.empty())
    CTN_CUDA(cudaMemcpyAsync(p.h_pack, p.h_pack_host.data(), p.h_pack_host.size(), cudaMemcpyHostToDevice, 0));
  int rc = ctn_plan_execute(plan, dev.data(), p.h_out, p.h_log, (int32_t*)(p.h_log + 1), p.h_ws, out_mode, nullptr);
  if (rc != CTN_OK) return rc;
  CTN_CUDA(cudaMemcpyAsync(out_host, p.h_out, (size_t)p.slots[p.final_slot].numel * p.elem, cudaMemcpyDeviceToHost, 0));
  double tail[2] = {0, 0};
  CTN_CUDA(cudaMemcpyAsync(tail, p.h_log, 16, cudaMemcpyDeviceToHost, 0));
  CTN_CUDA(cudaStreamSynchronize(0));
  *log_scale_host = tail[0];
  int32_t fl = 0;
  memcpy(&fl, &tail[1], sizeof(int32_t));
  if (flags_host) *flags_host = fl;
  if (fl & CTN_RESULT_INTERNAL) {
    g_cuda_error = "a kernel's internal pipeline timed out (StateHdr diagnostics lost); result invalid";
    return CTN_ERR_INTERNAL;
  }
  return CTN_OK;
}

}  // extern "C"
