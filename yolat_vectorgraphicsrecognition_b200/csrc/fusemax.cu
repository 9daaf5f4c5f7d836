// fusemax.cu -- fusion_block + cat + per-proposal scatter-max, fused
// (cad_recognition/architecture3cc_rpn_gp_iter2.py:62-63 and :122):
//     fusion   = relu(bn(feats W^T + b))                  [M,F]     (never written post-BN)
//     out_feat = cat(fusion, feats)                       [M,F+K]   (never built)
//     pooled   = scatter(out_feat, bbox_idx, 'max')       [S,F+K]
// The BN+ReLU is applied on the fly inside the segment-max reader, and the backward uses the fact
// that the post-ReLU gradient has exactly one non-zero row per (proposal, column).
// Tape: z [M,F] pre-BN, stat [4F], arg [S,F+K] int32 (row index of the max, -1 = empty segment).
#include "common.cuh"

namespace yolat {

struct FmTape { float* z; float* stat; int32_t* arg; };

static void fm_tape_layout(Arena& t, int64_t M, int K, int F, int64_t S, FmTape* o) {
  o->z = t.take(M * F);
  o->stat = t.take(4 * F);
  o->arg = t.take<int32_t>(S * (F + K));
}

static int fm_fwd_impl(const float* feats, int64_t ldf, int64_t M, int K, const float* w, const float* b, int F,
                       const yolat_bn* bn, int training, const int32_t* seg, int64_t S, float* pooled, int64_t ldp,
                       Arena& tape, Arena& ws, cudaStream_t st) {
  const bool dry = ws.dry();
  FmTape t;
  fm_tape_layout(tape, M, K, F, S, &t);
  if (!dry && tape.overflow) return YOLAT_ERR_WORKSPACE;
  GemmArgs a{};
  a.A = feats; a.lda = ldf; a.B = w; a.ldb = K; a.C = t.z; a.ldc = F; a.M = (int)M; a.N = F; a.K = K; a.bias = b;
  yolat_bn bb = dry ? yolat_bn{} : *bn;
  YOLAT_TRY(linear_bn_stats(a, ws, &bb, training, t.stat, st));
  if (!dry) {
    SegView sv;
    seg_layout(M, S, seg, &sv);
    YOLAT_TRY(segmax_launch(t.z, F, F, sv, S, t.stat, pooled, ldp, t.arg, F + K, st));
    YOLAT_TRY(segmax_launch(feats, ldf, K, sv, S, nullptr, pooled + F, ldp, t.arg + F, F + K, st));
    if (ws.overflow) return YOLAT_ERR_WORKSPACE;
  }
  return YOLAT_OK;
}

static int fm_bwd_impl(const float* feats, int64_t ldf, int64_t M, int K, const float* w, int F, const yolat_bn* bn,
                       int training, const int32_t* seg, int64_t S, const float* gp, int64_t ldg, float* dfeats,
                       int64_t lddf, int accumulate, float* dw, float* db, float* dgamma, float* dbeta, Arena& tape,
                       Arena& ws, cudaStream_t st) {
  const bool dry = ws.dry();
  FmTape t;
  fm_tape_layout(tape, M, K, F, S, &t);
  const int nparts = fusemax_bwd_nparts(S);
  float* part = ws.take((int64_t)nparts * 2 * F);
  float* bstat = ws.take(2 * F);
  if (!dry) {
    if (ws.overflow) return YOLAT_ERR_WORKSPACE;
    SegView sv;
    seg_layout(M, S, seg, &sv);
    int np = 0;
    YOLAT_TRY(fusemax_bwd_partial_launch(gp, ldg, F, S, t.arg, F + K, t.z, t.stat, part, &np, st));
    YOLAT_TRY(bn_bwd_finalize(part, np, M, F, t.stat, bn->w, training, bstat, dgamma, dbeta, db, st));
    // z <- dz (dense, in place: the tape is consumed)
    YOLAT_TRY(fusemax_bwd_apply_launch(t.z, M, F, sv.seg_of_row, gp, ldg, t.arg, F + K, t.stat, bstat, st));
  }
  if (dw || dry) {
    GemmArgs a{};
    a.A = t.z; a.lda = F; a.B = feats; a.ldb = ldf; a.C = dw; a.ldc = K; a.M = F; a.N = K; a.K = M;
    YOLAT_TRY(gemm(a, GEMM_TN, ws, st));
  }
  if (dfeats || dry) {
    GemmArgs a{};
    a.A = t.z; a.lda = F; a.B = w; a.ldb = K; a.C = dfeats; a.ldc = lddf; a.M = (int)M; a.N = K; a.K = F;
    a.accumulate = accumulate;
    YOLAT_TRY(gemm(a, GEMM_NN, ws, st));
    // pass-through columns: pooled[:, F:F+K] = max feats  ->  one row per (segment, column)
    if (!dry) YOLAT_TRY(segmax_bwd_add_launch(gp + F, ldg, K, S, t.arg + F, F + K, dfeats, lddf, st));
  }
  if (!dry && ws.overflow) return YOLAT_ERR_WORKSPACE;
  return YOLAT_OK;
}

}  // namespace yolat

using namespace yolat;

extern "C" {

int64_t yolat_fusemax_tape_floats(int64_t M, int K, int F, int64_t S) {
  Arena t(nullptr, 0);
  FmTape o;
  fm_tape_layout(t, M, K, F, S, &o);
  return t.off;
}

int64_t yolat_fusemax_ws_floats(int64_t M, int K, int F, int64_t S) {
  Arena t1(nullptr, 0), w1(nullptr, 0), t2(nullptr, 0), w2(nullptr, 0);
  fm_fwd_impl(nullptr, K, M, K, nullptr, nullptr, F, nullptr, 1, nullptr, S, nullptr, F + K, t1, w1, nullptr);
  fm_bwd_impl(nullptr, K, M, K, nullptr, F, nullptr, 1, nullptr, S, nullptr, F + K, nullptr, K, 0, nullptr, nullptr, nullptr,
              nullptr, t2, w2, nullptr);
  return w1.off > w2.off ? w1.off : w2.off;
}

int yolat_fusemax_fwd(const float* feats, int64_t ldf, int64_t M, int K, const float* w, const float* b, int F,
                      const yolat_bn* bn, int training, const int32_t* seg, int64_t S, float* pooled, int64_t ldp,
                      float* tape, int64_t tape_floats, float* ws, int64_t ws_floats, void* stream) {
  if (!feats || !w || !bn || !seg || !pooled || !tape || M <= 0 || K <= 0 || F <= 0 || S <= 0) return YOLAT_ERR_INVALID;
  if (tape_floats < yolat_fusemax_tape_floats(M, K, F, S)) return YOLAT_ERR_WORKSPACE;
  static float dummy;
  Arena t(tape, tape_floats), wsa(ws ? ws : &dummy, ws ? ws_floats : 0);
  return fm_fwd_impl(feats, ldf, M, K, w, b, F, bn, training, seg, S, pooled, ldp, t, wsa, (cudaStream_t)stream);
}

int yolat_fusemax_bwd(const float* feats, int64_t ldf, int64_t M, int K, const float* w, int F, const yolat_bn* bn,
                      int training, const int32_t* seg, int64_t S, const float* g_pooled, int64_t ldg, float* dfeats,
                      int64_t lddf, int accumulate_dfeats, float* dw, float* db, float* dgamma, float* dbeta, float* tape,
                      float* ws, int64_t ws_floats, void* stream) {
  if (!feats || !w || !bn || !seg || !g_pooled || !tape || !ws || M <= 0 || K <= 0 || F <= 0 || S <= 0)
    return YOLAT_ERR_INVALID;
  Arena t(tape, yolat_fusemax_tape_floats(M, K, F, S)), wsa(ws, ws_floats);
  return fm_bwd_impl(feats, ldf, M, K, w, F, bn, training, seg, S, g_pooled, ldg, dfeats, lddf, accumulate_dfeats, dw, db,
                     dgamma, dbeta, t, wsa, (cudaStream_t)stream);
}

}  // extern "C"
