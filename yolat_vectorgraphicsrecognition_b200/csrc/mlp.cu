// mlp.cu -- one [Linear, BatchNorm1d?, ReLU?] stage of gcn_lib.sparse.MLP (gcn_lib/sparse/torch_nn.py:50-71).
//
// Used for mlp_node (inside the conv), fusion_block_super, prediction_cls.{0,1,2}
// (cad_recognition/architecture3cc_rpn_gp_iter2.py:40-41,91-93).  y = act(bn(x W^T + b)).
// Tape: z [M,Nout] (pre-BN / pre-ReLU) when BN or ReLU is present, + the BN statistic block.
#include "common.cuh"

namespace yolat {

struct MlpTape { float* z; float* stat; };

static void mlp_tape_layout(Arena& t, int64_t M, int Nout, int flags, MlpTape* o) {
  o->z = (flags & (YOLAT_MLP_BN | YOLAT_MLP_RELU)) ? t.take(M * Nout) : nullptr;
  o->stat = (flags & YOLAT_MLP_BN) ? t.take(4 * Nout) : nullptr;
}

int mlp_fwd_impl(const float* x, int64_t ldx, int64_t M, int K, const float* w, const float* b, int Nout,
                 const yolat_bn* bn, int flags, float* y, int64_t ldy, Arena& tape, Arena& ws, cudaStream_t st) {
  const bool dry = ws.dry();
  const bool has_bn = flags & YOLAT_MLP_BN, has_relu = flags & YOLAT_MLP_RELU;
  const int training = (flags & YOLAT_MLP_TRAINING) ? 1 : 0;
  MlpTape t;
  mlp_tape_layout(tape, M, Nout, flags, &t);
  if (!dry && tape.overflow) return YOLAT_ERR_WORKSPACE;
  GemmArgs a{};
  a.A = x; a.lda = ldx; a.B = w; a.ldb = K; a.M = (int)M; a.N = Nout; a.K = K; a.bias = b;
  if (t.z || (dry && (has_bn || has_relu))) { a.C = t.z; a.ldc = Nout; } else { a.C = y; a.ldc = ldy; }
  if (has_bn) {
    yolat_bn bb = dry ? yolat_bn{} : *bn;
    YOLAT_TRY(linear_bn_stats(a, ws, &bb, training, t.stat, st));
    if (!dry) YOLAT_TRY(bn_apply(t.z, Nout, M, Nout, t.stat, has_relu, y, ldy, st));
  } else {
    YOLAT_TRY(gemm(a, GEMM_NT, ws, st));
  }
  if (!has_bn && has_relu && !dry) {
    // identity "statistics": sc = 1, sh = 0 are not materialised; copy + relu
    YOLAT_TRY(relu_bwd(t.z, Nout, t.z, Nout, M, Nout, y, ldy, st));   // y = z > 0 ? z : 0
  }
  if (!dry && ws.overflow) return YOLAT_ERR_WORKSPACE;
  return YOLAT_OK;
}

int mlp_bwd_impl(const float* x, int64_t ldx, int64_t M, int K, const float* w, int Nout, const yolat_bn* bn, int flags,
                 const float* gy, int64_t ldgy, float* dx, int64_t lddx, int accumulate_dx, float* dw, float* db,
                 float* dgamma, float* dbeta, Arena& tape, Arena& ws, cudaStream_t st) {
  const bool dry = ws.dry();
  const bool has_bn = flags & YOLAT_MLP_BN, has_relu = flags & YOLAT_MLP_RELU;
  const int training = (flags & YOLAT_MLP_TRAINING) ? 1 : 0;
  MlpTape t;
  mlp_tape_layout(tape, M, Nout, flags, &t);
  const float* dz = gy;
  int64_t lddz = ldgy;
  if (has_bn) {
    float* buf = ws.take(M * Nout);
    BnBwdArgs b{};
    b.gy = gy; b.ldgy = ldgy; b.z = t.z; b.ldz = Nout; b.M = M; b.C = Nout; b.stat = t.stat;
    b.gamma = dry ? nullptr : bn->w; b.relu = has_relu; b.training = training; b.dz = buf; b.lddz = Nout;
    b.dgamma = dgamma; b.dbeta = dbeta; b.dbias = db;
    YOLAT_TRY(bn_backward(b, ws, st));
    dz = buf; lddz = Nout;
  } else {
    if (has_relu) {
      float* buf = ws.take(M * Nout);
      if (!dry) YOLAT_TRY(relu_bwd(gy, ldgy, t.z, Nout, M, Nout, buf, Nout, st));
      dz = buf; lddz = Nout;
    }
    if (db || dry) YOLAT_TRY(colsum(dz, lddz, M, Nout, db, ws, st));
  }
  if (dw || dry) {
    GemmArgs a{};
    a.A = dz; a.lda = lddz; a.B = x; a.ldb = ldx; a.C = dw; a.ldc = K; a.M = Nout; a.N = K; a.K = M;
    YOLAT_TRY(gemm(a, GEMM_TN, ws, st));
  }
  if (dx || dry) {
    GemmArgs a{};
    a.A = dz; a.lda = lddz; a.B = w; a.ldb = K; a.C = dx; a.ldc = lddx; a.M = (int)M; a.N = K; a.K = Nout;
    a.accumulate = accumulate_dx;
    YOLAT_TRY(gemm(a, GEMM_NN, ws, st));
  }
  if (!dry && ws.overflow) return YOLAT_ERR_WORKSPACE;
  return YOLAT_OK;
}

}  // namespace yolat

using namespace yolat;

extern "C" {

int64_t yolat_gemm_ws_floats(int mode, int64_t M, int64_t N, int64_t K) {
  if (mode < 0 || mode > 2) return -1;
  Arena ws(nullptr, 0);
  GemmArgs a{};
  a.M = (int)M; a.N = (int)N; a.K = K;
  gemm(a, (GemmMode)mode, ws, nullptr);
  return ws.off;
}

int yolat_gemm(int mode, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t M,
               int64_t N, int64_t K, const float* bias, int accumulate, float* ws, int64_t ws_floats, void* stream) {
  if (mode < 0 || mode > 2 || !C || M < 0 || N < 0 || K < 0 || (K > 0 && (!A || !B))) return YOLAT_ERR_INVALID;
  if (M >= (1ll << 31) || N >= (1ll << 31)) return YOLAT_ERR_UNSUPPORTED;
  static float dummy;
  Arena wsa(ws ? ws : &dummy, ws ? ws_floats : 0);
  GemmArgs a{};
  a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.C = C; a.ldc = ldc;
  a.M = (int)M; a.N = (int)N; a.K = K; a.bias = bias; a.accumulate = accumulate;
  return gemm(a, (GemmMode)mode, wsa, (cudaStream_t)stream);
}

int64_t yolat_mlp_tape_floats(int64_t M, int K, int Nout, int flags) {
  (void)K;
  Arena t(nullptr, 0);
  MlpTape o;
  mlp_tape_layout(t, M, Nout, flags, &o);
  return t.off;
}

int64_t yolat_mlp_ws_floats(int64_t M, int K, int Nout, int flags) {
  Arena t1(nullptr, 0), w1(nullptr, 0), t2(nullptr, 0), w2(nullptr, 0);
  mlp_fwd_impl(nullptr, K, M, K, nullptr, nullptr, Nout, nullptr, flags, nullptr, Nout, t1, w1, nullptr);
  mlp_bwd_impl(nullptr, K, M, K, nullptr, Nout, nullptr, flags, nullptr, Nout, nullptr, K, 0, nullptr, nullptr, nullptr,
               nullptr, t2, w2, nullptr);
  return w1.off > w2.off ? w1.off : w2.off;
}

int yolat_mlp_fwd(const float* x, int64_t ldx, int64_t M, int K, const float* w, const float* b, int Nout,
                  const yolat_bn* bn, int flags, float* y, int64_t ldy, float* tape, int64_t tape_floats, float* ws,
                  int64_t ws_floats, void* stream) {
  if (!x || !w || !y || M <= 0 || K <= 0 || Nout <= 0) return YOLAT_ERR_INVALID;
  if ((flags & YOLAT_MLP_BN) && !bn) return YOLAT_ERR_INVALID;
  const int64_t need = yolat_mlp_tape_floats(M, K, Nout, flags);
  if (need > 0 && (!tape || tape_floats < need)) return YOLAT_ERR_WORKSPACE;
  static float dummy;
  Arena t(tape ? tape : &dummy, tape_floats), wsa(ws ? ws : &dummy, ws ? ws_floats : 0);
  return mlp_fwd_impl(x, ldx, M, K, w, b, Nout, bn, flags, y, ldy, t, wsa, (cudaStream_t)stream);
}

int yolat_mlp_bwd(const float* x, int64_t ldx, int64_t M, int K, const float* w, int Nout, const yolat_bn* bn, int flags,
                  const float* gy, int64_t ldgy, float* dx, int64_t lddx, int accumulate_dx, float* dw, float* db,
                  float* dgamma, float* dbeta, const float* tape, float* ws, int64_t ws_floats, void* stream) {
  if (!x || !w || !gy || M <= 0 || K <= 0 || Nout <= 0) return YOLAT_ERR_INVALID;
  if ((flags & YOLAT_MLP_BN) && !bn) return YOLAT_ERR_INVALID;
  const int64_t need = yolat_mlp_tape_floats(M, K, Nout, flags);
  if (need > 0 && !tape) return YOLAT_ERR_WORKSPACE;
  static float dummy;
  Arena t(tape ? const_cast<float*>(tape) : &dummy, need), wsa(ws ? ws : &dummy, ws ? ws_floats : 0);
  return mlp_bwd_impl(x, ldx, M, K, w, Nout, bn, flags, gy, ldgy, dx, lddx, accumulate_dx, dw, db, dgamma, dbeta, t, wsa,
                      (cudaStream_t)stream);
}

}  // extern "C"
