// conv.cu -- GraphConv('attr_edge_gp2') forward / backward: the C-ABI entry points and their kernel schedule.
//
// Reference: AttrRelativeEdgeConvGlobalPool2, gcn_lib/sparse/torch_vertex.py:288-341 (forward :319-328,
// message :330-337, aggr='mean' :308), built on gcn_lib.sparse.MLP (torch_nn.py:50-71).
//
// Schedule (training forward):
//   Wpq  = [W1a - W1b ; W1b]                      k_prep_wpq
//   PQ   = x Wpq^T                    [N,2C]      GEMM (node-level; replaces the edge-level Lin1)
//   z1   = P[i] + Q[j] + W1c attr + b1 [E,C]      k_edge_z1  (+ BN1 partial sums)   CSR-by-target slot order
//   z2   = relu(bn1(z1)) W2^T + b2     [E,C]      GEMM with BN+ReLU operand prologue
//   out  = x Wr^T + br                 [N,C]      GEMM
//   out += mean_i relu(bn2(z2))                   k_edge_agg (segmented, no atomics)
//   xn   = relu(bn_n(x_node Wn^T + bn))           GEMM + column stats + apply
// Tape (saved for backward): z1, z2, zn and the three BN statistic blocks.
#include <cstdlib>
#include "common.cuh"

namespace yolat {

struct Gp2Tape {
  float* z1; float* z2; float* zn; float* stat1; float* stat2; float* statn; float* pq;
};

// `lite`: the tape of the recompute backward (edge_bwd.cu) -- no per-edge activation, only node-level tensors:
// the node branch's pre-activation, the three BN statistic blocks and P | Q.
static void gp2_tape_layout(Arena& t, int64_t N, int64_t E, int C, bool lite, Gp2Tape* o) {
  o->z1 = lite ? nullptr : t.take(E * C);
  o->z2 = lite ? nullptr : t.take(E * C);
  o->zn = t.take(N * C);
  o->stat1 = t.take(4 * C);
  o->stat2 = t.take(4 * C);
  o->statn = t.take(4 * C);
  o->pq = lite ? t.take(N * 2 * C) : nullptr;
}

// YOLAT_EDGE=unfused selects the three-kernel edge path (z1 -> GEMM -> aggregate) instead of edge_fused.cu
static bool gemm_fused_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("YOLAT_EDGE");
    v = (e && e[0] == 'u') ? 0 : 1;
  }
  return v == 1;
}

// ---- branch concurrency ------------------------------------------------------------------------------------------
// The edge path, the lin_r product and the node branch of one GraphConv are independent until the very end, and at
// config-2 sizes most of their kernels are too small to fill 148 SMs.  They run on two library-owned side streams,
// forked from / joined back into the caller's stream with events; stream-ordered and capturable, so inside a CUDA
// graph the branches become parallel paths.  YOLAT_BRANCHES=0 runs everything on the caller's stream.
struct Branches {
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
  bool ok = false;
};
// Past this many nodes every kernel of a layer fills the GPU on its own and concurrent branches only fight over SMs and
// L2 (measured at N = 320 000: the P | Q / lin_r / node GEMMs 108 -> 82 us each, pass A of K-EDGE 146 -> 81 us).
constexpr int64_t kBranchMaxNodes = 65536;
static Branches* branches() {
  static thread_local Branches b[16];
  static const bool enabled = !(getenv("YOLAT_BRANCHES") && getenv("YOLAT_BRANCHES")[0] == '0');
  if (!enabled) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  Branches& r = b[dev];
  if (!r.ok) {
    bool good = true;
    for (int i = 0; i < 2; ++i) {
      good = good && cudaStreamCreateWithFlags(&r.side[i], cudaStreamNonBlocking) == cudaSuccess;
      good = good && cudaEventCreateWithFlags(&r.join[i], cudaEventDisableTiming) == cudaSuccess;
    }
    good = good && cudaEventCreateWithFlags(&r.fork, cudaEventDisableTiming) == cudaSuccess;
    if (!good) return nullptr;
    r.ok = true;
  }
  return &r;
}
// side streams start after everything already enqueued on `st`
static void fork_branches(Branches* b, cudaStream_t st) {
  cudaEventRecord(b->fork, st);
  cudaStreamWaitEvent(b->side[0], b->fork, 0);
  cudaStreamWaitEvent(b->side[1], b->fork, 0);
}
// `st` continues after everything enqueued so far on side stream i
static void join_branch(Branches* b, int i, cudaStream_t st) {
  cudaEventRecord(b->join[i], b->side[i]);
  cudaStreamWaitEvent(st, b->join[i], 0);
}

// Training-mode calls with the fused K-EDGE kernels available run tape-free in both directions: the forward keeps
// only statistics + P | Q, the backward recomputes (edge_bwd.cu).  YOLAT_EDGE_BWD=tape selects the round-1 schedule
// (z1 / z2 tape + multi-kernel backward) for A/B comparisons; eval-mode backward and C != 64 always use it.
static bool edge_bwd_recompute_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("YOLAT_EDGE_BWD");
    v = (e && e[0] == 't') ? 0 : 1;
  }
  return v == 1;
}
static bool gp2_lite(int64_t N, int64_t E, int C, int training) {
  return training && E > 0 && edge_fused_supported(C) && edge_fused_fits(N, 3 * C) && gemm_fused_enabled() &&
         edge_bwd_recompute_enabled();
}

// K-EDGE v6 (edge_fused2.cu, channel-major accumulator) for the MMA passes; v5 (edge_fused.cu) with YOLAT_EF=v5 or when
// the aggregation cannot run in place on `out`
static int edge_fused_any(const GraphView& g, int64_t N, int64_t E, int flags, const float* pq, int64_t ldpq, const float* rec,
                          const float* w1, int Cin, const float* b1, const float* stat1, const float* w2, const float* b2,
                          const float* stat2, const float* ew, float* z1, float* z2, float* part, const float* base,
                          int64_t ldb, float* out, int64_t ldo, cudaStream_t st) {
  if (edge_fused2_enabled() && (!(flags & EF_AGG) || base == out))
    return edge_fused2(g, N, E, flags, pq, ldpq, rec, w1, Cin, b1, stat1, w2, b2, stat2, ew, z1, z2, part, base, ldb, out, ldo, st);
  return edge_fused(g, N, E, flags, pq, ldpq, rec, w1, Cin, b1, stat1, w2, b2, stat2, ew, z1, z2, part, base, ldb, out, ldo, st);
}

static bool gp2_channels_ok(int Cin, int Cn, int C) {
  return (C == 32 || C == 64 || C == 128) && Cin >= 1 && Cn >= 1;
}

static int gp2_fwd_impl(const yolat_gp2_params* p, int Cin, int Cn, int C, const float* x, int64_t ldx,
                        const float* x_node, int64_t ldxn, const float* attr, const float* ew,
                        const int32_t* graph, int64_t N, int64_t E, int mode, float* out, int64_t ldo,
                        float* xnode_out, int64_t ldxo, Arena& tape, Arena& ws, cudaStream_t st) {
  const bool dry = ws.dry();
  const int training = (mode & YOLAT_GP2_TRAINING) ? 1 : 0;
  const bool lite = gp2_lite(N, E, C, training);
  const bool no_tape = lite || (mode & YOLAT_GP2_NO_TAPE) != 0;   // z1 / z2 are never written
  Gp2Tape t;
  gp2_tape_layout(tape, N, E, C, lite, &t);
  // Fused path: one GEMM x [Wp; Wq; Wr]^T -> P | Q | lin_r(x) (row stride 3C); the aggregation adds the lin_r column
  // block while writing `out`.  Unfused path (C = 32 / 128): lin_r goes straight to `out`, P | Q has row stride 2C.
  // (Folding lin_r into the P | Q GEMM -- kFusePqr -- was measured slower at the > L2 scale: the third 64-column block
  // wastes half of a 128-wide MMA tile and the wider rows cost the Q gathers their L2 residency.)
  constexpr bool kFusePqr = false;
  const bool fused = E > 0 && edge_fused_supported(C) && edge_fused_fits(N, 3 * C) && gemm_fused_enabled();
  const bool pqr = fused && kFusePqr;
  const int ldpq = pqr ? 3 * C : 2 * C;
  float* wpq = ws.take((int64_t)3 * C * Cin);
  float* bias3 = ws.take(3 * C);
  float* pq = lite ? t.pq : ws.take(N * ldpq);
  const int nparts = edge_z1_nparts(N);
  float* part1 = ws.take((int64_t)nparts * 2 * C);
  if (!dry && (ws.overflow || tape.overflow)) return YOLAT_ERR_WORKSPACE;

  GraphView g;
  if (!dry) graph_layout(N, E, graph, &g);
  Branches* br = (dry || N > kBranchMaxNodes) ? nullptr : branches();
  cudaStream_t st_node = br ? br->side[0] : st, st_linr = br ? br->side[1] : st;
  if (br) fork_branches(br, st);

  // ---- lin_r: out = x Wr^T + br  (torch_vertex.py:325) ------------------------------------------
  if (!pqr) {
    GemmArgs a{};
    a.A = x; a.lda = ldx; a.B = dry ? nullptr : p->wr; a.ldb = Cin; a.C = out; a.ldc = ldo;
    a.M = (int)N; a.N = C; a.K = Cin; a.bias = dry ? nullptr : p->br;
    YOLAT_TRY(gemm(a, GEMM_NT, ws, st_linr));
  }
  // ---- node path: mlp_node(x_node)  (torch_vertex.py:326) -----------------------------------------
  {
    GemmArgs a{};
    a.A = x_node; a.lda = ldxn; a.B = dry ? nullptr : p->wn; a.ldb = Cn; a.C = t.zn; a.ldc = C;
    a.M = (int)N; a.N = C; a.K = Cn; a.bias = dry ? nullptr : p->bnode;
    yolat_bn bnn = dry ? yolat_bn{} : p->bnn;
    YOLAT_TRY(linear_bn_stats(a, ws, &bnn, training, t.statn, st_node));
    if (!dry) YOLAT_TRY(bn_apply(t.zn, C, N, C, t.statn, 1, xnode_out, ldxo, st_node));
  }
  // ---- edge path ---------------------------------------------------------------------------------
  if (E > 0) {
    if (!dry) YOLAT_TRY(edge_prep_wpq(p->w1, Cin, C, wpq, pqr ? p->wr : nullptr, p->br, bias3, st));
    {
      GemmArgs a{};
      a.A = x; a.lda = ldx; a.B = wpq; a.ldb = Cin; a.C = pq; a.ldc = ldpq;
      a.M = (int)N; a.N = ldpq; a.K = Cin; a.bias = pqr ? bias3 : nullptr;
      YOLAT_TRY(gemm(a, GEMM_NT, ws, st));
    }
    if (fused) {
      // K-EDGE fused path (edge_fused.cu): pass A = BN1 statistics (no [E,C] store), then the tcgen05 kernel
      // once for the BN2 statistics and once for the segmented mean; z1 / z2 reach HBM only as backward tape.
      const int ngrid = edge_fused_grid(E), ngrid1 = edge_stats1_grid(E);
      float* part2 = ws.take((int64_t)ngrid * 2 * C);
      float* part1f = ws.take((int64_t)ngrid1 * 2 * C);
      float* rec = ws.take(edge_records_floats(E));
      const float* base = pqr ? pq + 2 * C : out;       // lin_r(x): third column block, or already in `out`
      const int64_t ldbase = pqr ? ldpq : ldo;
      if (!dry) {
        if (ws.overflow) return YOLAT_ERR_WORKSPACE;
        YOLAT_TRY(edge_records(g, E, attr, ldpq, rec, st));
        if (training) YOLAT_TRY(edge_stats1(g, N, E, pq, ldpq, rec, p->w1, Cin, p->b1, part1f, st));
        YOLAT_TRY(bn_finalize_from_partials(part1f, ngrid1, E, C, &p->bn1, training, t.stat1, st));
        if (training) {
          YOLAT_TRY(edge_fused_any(g, N, E, EF_STATS | (no_tape ? 0 : EF_TAPE), pq, ldpq, rec, p->w1, Cin, p->b1, t.stat1,
                               p->w2, p->b2, nullptr, ew, t.z1, t.z2, part2, nullptr, 0, nullptr, 0, st));
          YOLAT_TRY(bn_finalize_from_partials(part2, ngrid, E, C, &p->bn2, 1, t.stat2, st));
          if (br) join_branch(br, 1, st);      // lin_r(x) must be in `out` before the aggregation adds to it
          if (no_tape) {
            YOLAT_TRY(edge_fused_any(g, N, E, EF_AGG, pq, ldpq, rec, p->w1, Cin, p->b1, t.stat1, p->w2, p->b2, t.stat2, ew,
                                 nullptr, nullptr, nullptr, base, ldbase, out, ldo, st));
          } else {
            YOLAT_TRY(edge_agg(g, N, C, t.z2, t.stat2, ew, base, ldbase, out, ldo, st));
          }
        } else {
          YOLAT_TRY(bn_finalize_from_partials(nullptr, 0, E, C, &p->bn2, 0, t.stat2, st));
          if (br) join_branch(br, 1, st);
          YOLAT_TRY(edge_fused_any(g, N, E, EF_AGG | (no_tape ? 0 : EF_TAPE), pq, ldpq, rec, p->w1, Cin, p->b1, t.stat1,
                               p->w2, p->b2, t.stat2, ew, t.z1, t.z2, nullptr, base, ldbase, out, ldo, st));
        }
      }
    } else {
      if (!dry) {
        YOLAT_TRY(edge_z1(g, N, C, pq, attr, p->w1, Cin, p->b1, t.z1, training ? part1 : nullptr, st));
        YOLAT_TRY(bn_finalize_from_partials(part1, nparts, E, C, &p->bn1, training, t.stat1, st));
      }
      {
        GemmArgs a{};
        a.A = t.z1; a.lda = C; a.B = dry ? nullptr : p->w2; a.ldb = C; a.C = t.z2; a.ldc = C;
        a.M = (int)E; a.N = C; a.K = C; a.bias = dry ? nullptr : p->b2;
        a.a_sc = t.stat1; a.a_sh = dry ? nullptr : t.stat1 + C;
        yolat_bn bn2 = dry ? yolat_bn{} : p->bn2;
        YOLAT_TRY(linear_bn_stats(a, ws, &bn2, training, t.stat2, st));
      }
      if (br) join_branch(br, 1, st);
      if (!dry) YOLAT_TRY(edge_agg(g, N, C, t.z2, t.stat2, ew, out, ldo, out, ldo, st));
    }
  } else if (br) {
    join_branch(br, 1, st);
  }
  if (br) join_branch(br, 0, st);
  if (!dry && ws.overflow) return YOLAT_ERR_WORKSPACE;
  return YOLAT_OK;
}

static int zero_opt(float* p, int64_t n, cudaStream_t st) { return p ? fill_zero(p, n, st) : YOLAT_OK; }

static int gp2_bwd_impl(const yolat_gp2_params* p, const yolat_gp2_grads* gr, int Cin, int Cn, int C, const float* x,
                        int64_t ldx, const float* x_node, int64_t ldxn, const float* attr, const float* ew,
                        const int32_t* graph, int64_t N, int64_t E, int training, const float* g_out, int64_t ldgo,
                        const float* g_xnode, int64_t ldgx, float* dx, int64_t lddx, float* dx_node, int64_t lddxn,
                        int accumulate_dx, Arena& tape, Arena& ws, cudaStream_t st) {
  const bool dry = ws.dry();
  const bool lite = gp2_lite(N, E, C, training);
  Gp2Tape t;
  gp2_tape_layout(tape, N, E, C, lite, &t);
  const int ld1 = 2 * Cin + 4;
  yolat_gp2_grads G{};
  if (!dry) G = *gr;
  GraphView g;
  if (!dry) graph_layout(N, E, graph, &g);
  int acc_dx = accumulate_dx, acc_dxn = accumulate_dx;
  Branches* br = (dry || N > kBranchMaxNodes) ? nullptr : branches();
  cudaStream_t st_node = br ? br->side[0] : st, st_linr = br ? br->side[1] : st;
  if (br) fork_branches(br, st);

  // ---- node branch (side stream 0) -----------------------------------------------------------------
  {
    float* dzn = ws.take(N * C);
    BnBwdArgs b{};
    b.gy = g_xnode; b.ldgy = ldgx; b.z = t.zn; b.ldz = C; b.M = N; b.C = C; b.stat = t.statn;
    b.gamma = dry ? nullptr : p->bnn.w; b.relu = 1; b.training = training; b.dz = dzn; b.lddz = C;
    b.dgamma = G.bnn_w; b.dbeta = G.bnn_b; b.dbias = G.bnode;
    YOLAT_TRY(bn_backward(b, ws, st_node));
    if (G.wn || dry) {
      GemmArgs a{};
      a.A = dzn; a.lda = C; a.B = x_node; a.ldb = ldxn; a.C = G.wn; a.ldc = Cn; a.M = C; a.N = Cn; a.K = N;
      YOLAT_TRY(gemm(a, GEMM_TN, ws, st_node));
    }
    if (dx_node || dry) {
      GemmArgs a{};
      a.A = dzn; a.lda = C; a.B = dry ? nullptr : p->wn; a.ldb = Cn; a.C = dx_node; a.ldc = lddxn;
      a.M = (int)N; a.N = Cn; a.K = C; a.accumulate = acc_dxn;
      YOLAT_TRY(gemm(a, GEMM_NN, ws, st_node));
      acc_dxn = 1;
    }
  }

  // ---- lin_r (side stream 1) ------------------------------------------------------------------------
  if (G.wr || dry) {
    GemmArgs a{};
    a.A = g_out; a.lda = ldgo; a.B = x; a.ldb = ldx; a.C = G.wr; a.ldc = Cin; a.M = C; a.N = Cin; a.K = N;
    YOLAT_TRY(gemm(a, GEMM_TN, ws, st_linr));
  }
  if (G.br || dry) YOLAT_TRY(colsum(g_out, ldgo, N, C, G.br, ws, st_linr));
  if (dx || dry) {
    GemmArgs a{};
    a.A = g_out; a.lda = ldgo; a.B = dry ? nullptr : p->wr; a.ldb = Cin; a.C = dx; a.ldc = lddx;
    a.M = (int)N; a.N = Cin; a.K = C; a.accumulate = acc_dx;
    YOLAT_TRY(gemm(a, GEMM_NN, ws, st_linr));
    acc_dx = 1;
  }

  // ---- edge path ------------------------------------------------------------------------------------
  if (lite) {
    // Recompute backward (edge_bwd.cu): three fused passes over the slots, nothing of size [E, C] in HBM.
    EdgeBwdWs bw;
    edge_bwd_layout(ws, N, E, &bw);
    float* dpq = ws.take(N * 2 * C);
    float* wpq = ws.take((int64_t)2 * C * Cin);
    float* dwpq = ws.take((int64_t)2 * C * Cin);
    float* dw1c = ws.take((int64_t)C * 4);
    if (!dry) {
      if (ws.overflow) return YOLAT_ERR_WORKSPACE;
      YOLAT_TRY(edge_bwd_fused(g, N, E, t.pq, 2 * C, attr, p->w1, Cin, p->b1, t.stat1, p->bn1.w, p->w2, p->b2, t.stat2,
                               p->bn2.w, ew, g_out, ldgo, bw, dpq, dw1c, G.w2, G.b1, G.bn1_w, G.bn1_b, G.b2, G.bn2_w,
                               G.bn2_b, st));
    }
    if (G.w1 || dry) {   // dWpq = dPQ^T x ; dW1 = [dWp | dWq - dWp | dW1c]
      GemmArgs a{};
      a.A = dpq; a.lda = 2 * C; a.B = x; a.ldb = ldx; a.C = dwpq; a.ldc = Cin; a.M = 2 * C; a.N = Cin; a.K = N;
      YOLAT_TRY(gemm(a, GEMM_TN, ws, st));
      if (!dry && G.w1) YOLAT_TRY(edge_assemble_dw1(dwpq, dw1c, Cin, C, G.w1, st));
    }
    if (dx || dry) {     // dx += dPQ Wpq  (after lin_r's own contribution to dx)
      if (br) join_branch(br, 1, st);
      if (!dry) YOLAT_TRY(edge_prep_wpq(p->w1, Cin, C, wpq, nullptr, nullptr, nullptr, st));
      GemmArgs a{};
      a.A = dpq; a.lda = 2 * C; a.B = wpq; a.ldb = Cin; a.C = dx; a.ldc = lddx;
      a.M = (int)N; a.N = Cin; a.K = 2 * C; a.accumulate = acc_dx;
      YOLAT_TRY(gemm(a, GEMM_NN, ws, st));
    }
  } else if (E > 0) {
    float* dz2 = ws.take(E * C);
    float* dz1 = ws.take(E * C);
    float* dpq = ws.take(N * 2 * C);
    float* wpq = ws.take((int64_t)2 * C * Cin);
    float* dwpq = ws.take((int64_t)2 * C * Cin);
    float* dw1c = ws.take((int64_t)C * 4);
    const int nparts = edge_z1_nparts(N);
    float* partw = ws.take((int64_t)nparts * C * 4);
    // BN2 + ReLU backward with the mean-aggregation gather fused in:
    //   dy[slot] = g_out[dst[slot]] * deg_inv[dst[slot]] * edge_weight[eid[slot]]
    {
      BnBwdArgs b{};
      b.gy = g_out; b.ldgy = ldgo; b.row_idx = dry ? nullptr : g.dst_t; b.row_scale = dry ? nullptr : g.deg_inv;
      if (ew) { b.slot_idx = g.eid_t; b.slot_scale = ew; }
      b.z = t.z2; b.ldz = C; b.M = E; b.C = C; b.stat = t.stat2; b.gamma = dry ? nullptr : p->bn2.w;
      b.relu = 1; b.training = training; b.dz = dz2; b.lddz = C;
      b.dgamma = G.bn2_w; b.dbeta = G.bn2_b; b.dbias = G.b2;
      YOLAT_TRY(bn_backward(b, ws, st));
    }
    if (G.w2 || dry) {   // dW2 = dz2^T relu(bn1(z1))
      GemmArgs a{};
      a.A = dz2; a.lda = C; a.B = t.z1; a.ldb = C; a.C = G.w2; a.ldc = C; a.M = C; a.N = C; a.K = E;
      a.b_sc = t.stat1; a.b_sh = dry ? nullptr : t.stat1 + C;
      YOLAT_TRY(gemm(a, GEMM_TN, ws, st));
    }
    {                    // da1 = dz2 W2
      GemmArgs a{};
      a.A = dz2; a.lda = C; a.B = dry ? nullptr : p->w2; a.ldb = C; a.C = dz1; a.ldc = C;
      a.M = (int)E; a.N = C; a.K = C;
      YOLAT_TRY(gemm(a, GEMM_NN, ws, st));
    }
    {
      BnBwdArgs b{};
      b.gy = dz1; b.ldgy = C; b.z = t.z1; b.ldz = C; b.M = E; b.C = C; b.stat = t.stat1;
      b.gamma = dry ? nullptr : p->bn1.w; b.relu = 1; b.training = training; b.dz = dz1; b.lddz = C;
      b.dgamma = G.bn1_w; b.dbeta = G.bn1_b; b.dbias = G.b1;
      YOLAT_TRY(bn_backward(b, ws, st));
    }
    if (!dry) YOLAT_TRY(edge_bwd_scatter(g, N, C, dz1, attr, dpq, partw, dw1c, st));
    if (G.w1 || dry) {   // dWpq = dPQ^T x ; dW1 = [dWp | dWq - dWp | dW1c]
      GemmArgs a{};
      a.A = dpq; a.lda = 2 * C; a.B = x; a.ldb = ldx; a.C = dwpq; a.ldc = Cin; a.M = 2 * C; a.N = Cin; a.K = N;
      YOLAT_TRY(gemm(a, GEMM_TN, ws, st));
      if (!dry && G.w1) YOLAT_TRY(edge_assemble_dw1(dwpq, dw1c, Cin, C, G.w1, st));
    }
    if (dx || dry) {     // dx += dPQ Wpq  (after lin_r's own contribution to dx)
      if (br) join_branch(br, 1, st);
      if (!dry) YOLAT_TRY(edge_prep_wpq(p->w1, Cin, C, wpq, nullptr, nullptr, nullptr, st));
      GemmArgs a{};
      a.A = dpq; a.lda = 2 * C; a.B = wpq; a.ldb = Cin; a.C = dx; a.ldc = lddx;
      a.M = (int)N; a.N = Cin; a.K = 2 * C; a.accumulate = acc_dx;
      YOLAT_TRY(gemm(a, GEMM_NN, ws, st));
    }
  } else if (!dry) {
    YOLAT_TRY(zero_opt(G.w1, (int64_t)C * ld1, st)); YOLAT_TRY(zero_opt(G.b1, C, st));
    YOLAT_TRY(zero_opt(G.bn1_w, C, st)); YOLAT_TRY(zero_opt(G.bn1_b, C, st));
    YOLAT_TRY(zero_opt(G.w2, (int64_t)C * C, st)); YOLAT_TRY(zero_opt(G.b2, C, st));
    YOLAT_TRY(zero_opt(G.bn2_w, C, st)); YOLAT_TRY(zero_opt(G.bn2_b, C, st));
  }
  if (br) {
    join_branch(br, 1, st);
    join_branch(br, 0, st);
  }
  if (!dry && ws.overflow) return YOLAT_ERR_WORKSPACE;
  return YOLAT_OK;
}

// ---- one-layer edge convolutions of the same family ----------------------------------------------------------------
// 'edge' / 'attr_edge' / 'attr_edge_gp' of gcn_lib/sparse/torch_vertex.py (:427-484 + :546-557, :219-286 + :560-573,
// :343-425 + :575-590): message = nn(cat(...)) with nn = MLP([F, C], 'relu', 'batch') -- ONE [Linear, BN, ReLU] stage --
// optionally scaled by the edge weight, mean aggregation at the target.  The host embeds the recipe's Linear weight into
// the GP2 column layout [x_i | x_j - x_i | attr] (zero blocks for absent parts), so the node-level P | Q trick applies:
//   z1 = P[i] + Q[j] + W1c attr + b1 ;  out[i] = base[i] + mean_{e -> i} w_e relu(bn(z1_e)).
// Tape: z1 [E, C] in slot order + the BN statistic block (these recipes are not on the timed path).
struct Edge1Tape { float* z1; float* stat1; };
static void edge1_tape_layout(Arena& t, int64_t E, int C, Edge1Tape* o) {
  o->z1 = t.take(E * C);
  o->stat1 = t.take(4 * C);
}

static int edge1_fwd_impl(const float* w1, const float* b1, const yolat_bn* bn, int Cin, int C, const float* x, int64_t ldx,
                          const float* attr, const float* ew, const int32_t* graph, int64_t N, int64_t E, int training,
                          float* out, int64_t ldo, Arena& tape, Arena& ws, cudaStream_t st) {
  const bool dry = ws.dry();
  Edge1Tape t;
  edge1_tape_layout(tape, E, C, &t);
  float* wpq = ws.take((int64_t)2 * C * Cin);
  float* pq = ws.take(N * 2 * C);
  const int nparts = edge_z1_nparts(N);
  float* part1 = ws.take((int64_t)nparts * 2 * C);
  if (!dry && (ws.overflow || tape.overflow)) return YOLAT_ERR_WORKSPACE;
  if (E <= 0) return YOLAT_OK;                 // out keeps `base` (the caller pre-filled it)
  GraphView g;
  if (!dry) {
    graph_layout(N, E, graph, &g);
    YOLAT_TRY(edge_prep_wpq(w1, Cin, C, wpq, nullptr, nullptr, nullptr, st));
  }
  {
    GemmArgs a{};
    a.A = x; a.lda = ldx; a.B = wpq; a.ldb = Cin; a.C = pq; a.ldc = 2 * C; a.M = (int)N; a.N = 2 * C; a.K = Cin;
    YOLAT_TRY(gemm(a, GEMM_NT, ws, st));
  }
  if (!dry) {
    YOLAT_TRY(edge_z1(g, N, C, pq, attr, w1, Cin, b1, t.z1, training ? part1 : nullptr, st));
    YOLAT_TRY(bn_finalize_from_partials(part1, nparts, E, C, bn, training, t.stat1, st));
    YOLAT_TRY(edge_agg(g, N, C, t.z1, t.stat1, ew, out, ldo, out, ldo, st));
  }
  return YOLAT_OK;
}

static int edge1_bwd_impl(const float* w1, const yolat_bn* bn, int Cin, int C, const float* x, int64_t ldx, const float* attr,
                          const float* ew, const int32_t* graph, int64_t N, int64_t E, int training, const float* g_out,
                          int64_t ldgo, float* dx, int64_t lddx, int accumulate_dx, float* dw1, float* db1, float* dgamma,
                          float* dbeta, Arena& tape, Arena& ws, cudaStream_t st) {
  const bool dry = ws.dry();
  Edge1Tape t;
  edge1_tape_layout(tape, E, C, &t);
  const int ld1 = 2 * Cin + 4;
  if (E <= 0) {
    if (!dry) {
      YOLAT_TRY(zero_opt(dw1, (int64_t)C * ld1, st)); YOLAT_TRY(zero_opt(db1, C, st));
      YOLAT_TRY(zero_opt(dgamma, C, st)); YOLAT_TRY(zero_opt(dbeta, C, st));
      if (dx && !accumulate_dx) YOLAT_TRY(fill_zero(dx, N * lddx, st));   // lddx == Cin: contiguous
    }
    return YOLAT_OK;
  }
  GraphView g;
  if (!dry) graph_layout(N, E, graph, &g);
  float* dz1 = ws.take(E * C);
  float* dpq = ws.take(N * 2 * C);
  float* wpq = ws.take((int64_t)2 * C * Cin);
  float* dwpq = ws.take((int64_t)2 * C * Cin);
  float* dw1c = ws.take((int64_t)C * 4);
  const int nparts = edge_z1_nparts(N);
  float* partw = ws.take((int64_t)nparts * C * 4);
  {   // BN + ReLU backward with the mean-aggregation gather fused in
    BnBwdArgs b{};
    b.gy = g_out; b.ldgy = ldgo; b.row_idx = dry ? nullptr : g.dst_t; b.row_scale = dry ? nullptr : g.deg_inv;
    if (ew) { b.slot_idx = g.eid_t; b.slot_scale = ew; }
    b.z = t.z1; b.ldz = C; b.M = E; b.C = C; b.stat = t.stat1; b.gamma = dry ? nullptr : bn->w;
    b.relu = 1; b.training = training; b.dz = dz1; b.lddz = C; b.dgamma = dgamma; b.dbeta = dbeta; b.dbias = db1;
    YOLAT_TRY(bn_backward(b, ws, st));
  }
  if (!dry) YOLAT_TRY(edge_bwd_scatter(g, N, C, dz1, attr, dpq, partw, dw1c, st));
  if (dw1 || dry) {
    GemmArgs a{};
    a.A = dpq; a.lda = 2 * C; a.B = x; a.ldb = ldx; a.C = dwpq; a.ldc = Cin; a.M = 2 * C; a.N = Cin; a.K = N;
    YOLAT_TRY(gemm(a, GEMM_TN, ws, st));
    if (!dry && dw1) YOLAT_TRY(edge_assemble_dw1(dwpq, dw1c, Cin, C, dw1, st));
  }
  if (dx || dry) {
    if (!dry) YOLAT_TRY(edge_prep_wpq(w1, Cin, C, wpq, nullptr, nullptr, nullptr, st));
    GemmArgs a{};
    a.A = dpq; a.lda = 2 * C; a.B = wpq; a.ldb = Cin; a.C = dx; a.ldc = lddx; a.M = (int)N; a.N = Cin; a.K = 2 * C;
    a.accumulate = accumulate_dx;
    YOLAT_TRY(gemm(a, GEMM_NN, ws, st));
  }
  if (!dry && ws.overflow) return YOLAT_ERR_WORKSPACE;
  return YOLAT_OK;
}

}  // namespace yolat

using namespace yolat;

extern "C" {

int64_t yolat_gp2_tape_floats(int64_t N, int64_t E, int Cin, int Cn, int C) {
  return yolat_gp2_tape_floats_mode(N, E, Cin, Cn, C, 0);
}

int64_t yolat_gp2_tape_floats_mode(int64_t N, int64_t E, int Cin, int Cn, int C, int mode) {
  (void)Cin; (void)Cn;
  Arena t(nullptr, 0);
  Gp2Tape o;
  gp2_tape_layout(t, N, E, C, gp2_lite(N, E, C, mode & YOLAT_GP2_TRAINING), &o);
  return t.off;
}

int64_t yolat_gp2_fwd_ws_floats(int64_t N, int64_t E, int Cin, int Cn, int C) {
  if (!gp2_channels_ok(Cin, Cn, C)) return -1;
  int64_t need = 0;
  for (int mode = 0; mode < 2; ++mode) {     // the larger of the eval- and the training-mode plan
    Arena tape(nullptr, 0), ws(nullptr, 0);
    gp2_fwd_impl(nullptr, Cin, Cn, C, nullptr, Cin, nullptr, Cn, nullptr, nullptr, nullptr, N, E, mode, nullptr, C, nullptr,
                 C, tape, ws, nullptr);
    need = ws.off > need ? ws.off : need;
  }
  return need;
}

int64_t yolat_gp2_bwd_ws_floats(int64_t N, int64_t E, int Cin, int Cn, int C) {
  if (!gp2_channels_ok(Cin, Cn, C)) return -1;
  int64_t need = 0;
  for (int training = 0; training < 2; ++training) {
    Arena tape(nullptr, 0), ws(nullptr, 0);
    gp2_bwd_impl(nullptr, nullptr, Cin, Cn, C, nullptr, Cin, nullptr, Cn, nullptr, nullptr, nullptr, N, E, training, nullptr,
                 C, nullptr, C, nullptr, Cin, nullptr, Cn, 0, tape, ws, nullptr);
    need = ws.off > need ? ws.off : need;
  }
  return need;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int yolat_gp2_fwd(const yolat_gp2_params* p, int Cin, int Cn, int C, const float* x, int64_t ldx, const float* x_node,
                  int64_t ldxn, const float* attr, const float* edge_weight, const int32_t* graph, int64_t N, int64_t E,
                  int training, float* out, int64_t ldo, float* xnode_out, int64_t ldxo, float* tape, int64_t tape_floats,
                  float* ws, int64_t ws_floats, void* stream) {
  if (!p || !x || !x_node || !graph || !out || !xnode_out || !tape || !ws || N <= 0 || E < 0) return YOLAT_ERR_INVALID;
  if (E > 0 && (!attr || !aligned16(attr))) return YOLAT_ERR_INVALID;
  if (!gp2_channels_ok(Cin, Cn, C)) return YOLAT_ERR_UNSUPPORTED;
  if (tape_floats < yolat_gp2_tape_floats_mode(N, E, Cin, Cn, C, training)) return YOLAT_ERR_WORKSPACE;
  Arena t(tape, tape_floats), w(ws, ws_floats);
  return gp2_fwd_impl(p, Cin, Cn, C, x, ldx, x_node, ldxn, attr, edge_weight, graph, N, E, training, out, ldo, xnode_out,
                      ldxo, t, w, (cudaStream_t)stream);
}

int yolat_gp2_bwd(const yolat_gp2_params* p, const yolat_gp2_grads* g, int Cin, int Cn, int C, const float* x, int64_t ldx,
                  const float* x_node, int64_t ldxn, const float* attr, const float* edge_weight, const int32_t* graph,
                  int64_t N, int64_t E, int training, const float* g_out, int64_t ldgo, const float* g_xnode, int64_t ldgx,
                  float* dx, int64_t lddx, float* dx_node, int64_t lddxn, int accumulate_dx, const float* tape, float* ws,
                  int64_t ws_floats, void* stream) {
  if (!p || !g || !x || !x_node || !graph || !tape || !ws || !g_out || !g_xnode || N <= 0 || E < 0) return YOLAT_ERR_INVALID;
  if (E > 0 && (!attr || !aligned16(attr))) return YOLAT_ERR_INVALID;
  if (!gp2_channels_ok(Cin, Cn, C)) return YOLAT_ERR_UNSUPPORTED;
  Arena t(const_cast<float*>(tape), yolat_gp2_tape_floats_mode(N, E, Cin, Cn, C, training)), w(ws, ws_floats);
  return gp2_bwd_impl(p, g, Cin, Cn, C, x, ldx, x_node, ldxn, attr, edge_weight, graph, N, E, training, g_out, ldgo,
                      g_xnode, ldgx, dx, lddx, dx_node, lddxn, accumulate_dx, t, w, (cudaStream_t)stream);
}

int64_t yolat_edge1_tape_floats(int64_t N, int64_t E, int Cin, int C) {
  (void)N; (void)Cin;
  Arena t(nullptr, 0);
  Edge1Tape o;
  edge1_tape_layout(t, E, C, &o);
  return t.off;
}

int64_t yolat_edge1_ws_floats(int64_t N, int64_t E, int Cin, int C) {
  if (!gp2_channels_ok(Cin, Cin, C)) return -1;
  Arena t1(nullptr, 0), w1(nullptr, 0), t2(nullptr, 0), w2(nullptr, 0);
  edge1_fwd_impl(nullptr, nullptr, nullptr, Cin, C, nullptr, Cin, nullptr, nullptr, nullptr, N, E > 0 ? E : 1, 1, nullptr, C, t1, w1, nullptr);
  edge1_bwd_impl(nullptr, nullptr, Cin, C, nullptr, Cin, nullptr, nullptr, nullptr, N, E > 0 ? E : 1, 1, nullptr, C, nullptr, Cin, 0,
                 nullptr, nullptr, nullptr, nullptr, t2, w2, nullptr);
  return w1.off > w2.off ? w1.off : w2.off;
}

int yolat_edge1_fwd(const float* w1, const float* b1, const yolat_bn* bn, int Cin, int C, const float* x, int64_t ldx,
                    const float* attr, const float* edge_weight, const int32_t* graph, int64_t N, int64_t E, int training,
                    float* out, int64_t ldo, float* tape, int64_t tape_floats, float* ws, int64_t ws_floats, void* stream) {
  if (!w1 || !bn || !x || !graph || !out || !tape || !ws || N <= 0 || E < 0) return YOLAT_ERR_INVALID;
  if (E > 0 && (!attr || !aligned16(attr))) return YOLAT_ERR_INVALID;
  if (!gp2_channels_ok(Cin, Cin, C)) return YOLAT_ERR_UNSUPPORTED;
  if (tape_floats < yolat_edge1_tape_floats(N, E, Cin, C)) return YOLAT_ERR_WORKSPACE;
  Arena t(tape, tape_floats), w(ws, ws_floats);
  return edge1_fwd_impl(w1, b1, bn, Cin, C, x, ldx, attr, edge_weight, graph, N, E, training, out, ldo, t, w, (cudaStream_t)stream);
}

int yolat_edge1_bwd(const float* w1, const yolat_bn* bn, int Cin, int C, const float* x, int64_t ldx, const float* attr,
                    const float* edge_weight, const int32_t* graph, int64_t N, int64_t E, int training, const float* g_out,
                    int64_t ldgo, float* dx, int64_t lddx, int accumulate_dx, float* dw1, float* db1, float* dgamma,
                    float* dbeta, const float* tape, float* ws, int64_t ws_floats, void* stream) {
  if (!w1 || !bn || !x || !graph || !tape || !ws || !g_out || N <= 0 || E < 0) return YOLAT_ERR_INVALID;
  if (E > 0 && (!attr || !aligned16(attr))) return YOLAT_ERR_INVALID;
  if (!gp2_channels_ok(Cin, Cin, C)) return YOLAT_ERR_UNSUPPORTED;
  Arena t(const_cast<float*>(tape), yolat_edge1_tape_floats(N, E, Cin, C)), w(ws, ws_floats);
  return edge1_bwd_impl(w1, bn, Cin, C, x, ldx, attr, edge_weight, graph, N, E, training, g_out, ldgo, dx, lddx, accumulate_dx,
                        dw1, db1, dgamma, dbeta, t, w, (cudaStream_t)stream);
}

}  // extern "C"
