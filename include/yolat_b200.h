/*
 * yolat_b200.h -- C ABI of the B200-native engine for YOLaT's Bezier-graph proposal classifier.
 *
 * The reference (microsoft/YOLaT-VectorGraphicsRecognition) has no FFI: its hot path is Python calling
 * PyTorch / torch_scatter / PyG library kernels.  Each entry point below replaces one library-kernel
 * chain at the call site cited next to it (paths relative to the reference tree).  All functions
 *   - take plain device pointers + sizes (no torch types), fp32 values, int64 indices as delivered by
 *     the reference Dataset, int32 internally;
 *   - never allocate, never synchronise, never throw: outputs, tapes (activations saved for backward)
 *     and workspaces are caller-allocated; sizes come from the *_floats / *_ints query functions;
 *   - are stream-ordered on the `stream` argument (a cudaStream_t passed as void*), so they are safe
 *     under CUDA-graph capture;
 *   - return 0 on success or a negative yolat_status.
 * Leading dimensions (ld*) are in elements and let a caller write into a column slice of a wider
 * row-major matrix (this is how torch.cat at architecture3cc_rpn_gp_iter2.py:61,63,66,69,127 disappears).
 */
#ifndef YOLAT_B200_H_
#define YOLAT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  YOLAT_OK = 0,
  YOLAT_ERR_INVALID = -1,     /* bad argument (null pointer, unsupported channel count, ...) */
  YOLAT_ERR_WORKSPACE = -2,   /* workspace / tape smaller than the *_floats query says        */
  YOLAT_ERR_LAUNCH = -3,      /* a kernel launch failed (cudaGetLastError != cudaSuccess)      */
  YOLAT_ERR_UNSUPPORTED = -4
} yolat_status;

int yolat_abi_version(void);
const char* yolat_status_string(int status);
/* last CUDA error string seen by a failing launch on this thread (diagnostics only) */
const char* yolat_last_cuda_error(void);
/* number of kernels this library has launched in this process (bench.py reports the per-step delta) */
int64_t yolat_launch_count(void);

/* Opt-in per-kernel device timing (CUDA events recorded on the launching stream around the kernels named
 * below; bench.py uses it for the roofline of the dominant kernel, which is launched from inside
 * yolat_gp2_fwd).  yolat_prof_enable(1) starts a fresh window, yolat_prof_read synchronises the recorded
 * events and returns the number of launches and their summed duration.  Never enable during graph capture. */
#define YOLAT_PROF_EDGE_STATS1 0       /* ef::k_edge_stats1          (K-EDGE pass A)            */
#define YOLAT_PROF_EDGE_FUSED_STATS 1  /* ef::k_edge_fused<F_STATS>  (K-EDGE pass B)            */
#define YOLAT_PROF_EDGE_FUSED_AGG 2    /* ef::k_edge_fused<F_AGG>    (K-EDGE pass C, the scatter) */
#define YOLAT_PROF_GEMM 3              /* tc::k_tc_gemm (any mode)                              */
#define YOLAT_PROF_EDGE_BWD_D1 4       /* eb::k_edge_bwd<D1>   (K-EDGE backward: BN2 statistics)  */
#define YOLAT_PROF_EDGE_BWD_D2T 5      /* eb::k_edge_bwd<D2T>  (dW2, BN1 statistics, sums by target) */
#define YOLAT_PROF_EDGE_BWD_D2S 6      /* eb::k_edge_bwd<D2S>  (sums by source)                    */
int yolat_prof_enable(int on);
int yolat_prof_read(int id, int64_t* launches, double* total_ms);

/* ------------------------------------------------------------------------------------------------
 * Graph preparation.  Replaces what PyG's MessagePassing.propagate does implicitly on every call
 * (gather by edge_index[0]/[1], scatter by edge_index[1]; gcn_lib/sparse/torch_vertex.py:324) with a
 * once-per-batch CSR build shared by all conv layers and by backward.
 *   edge: int64, element (e, c) at edge[e*stride_e + c*stride_c]; c=0 source j, c=1 target i.
 *         ([E,2] contiguous: stride_e=2, stride_c=1; the transposed [2,E] view of
 *          architecture3cc_rpn_gp_iter2.py:110 has stride_e=1... whatever strides torch reports.)
 *   graph: int32 buffer of yolat_graph_ints(N,E) elements (layout private to the library).
 * Out-of-range indices are counted in the buffer's error slot (read it with yolat_graph_errors after
 * a sync) and the offending edges are dropped.
 * ---------------------------------------------------------------------------------------------- */
int64_t yolat_graph_ints(int64_t N, int64_t E);
int yolat_graph_build(const int64_t* edge, int64_t stride_e, int64_t stride_c, int64_t E, int64_t N,
                      int32_t* graph, void* stream);
/* device pointer to the int32 error counter inside a built graph buffer */
const int32_t* yolat_graph_error_ptr(const int32_t* graph, int64_t N, int64_t E);
/* device pointers to the CSR-by-target arrays (rowptr[N+1], src[E], eid[E]) for inspection/tests */
const int32_t* yolat_graph_rowptr(const int32_t* graph, int64_t N, int64_t E);
const int32_t* yolat_graph_src(const int32_t* graph, int64_t N, int64_t E);
const int32_t* yolat_graph_eid(const int32_t* graph, int64_t N, int64_t E);

/* Segments (proposals) from an index vector (bbox_idx).  Replaces the implicit grouping inside
 * torch_scatter.scatter (architecture3cc_rpn_gp_iter2.py:67,122).  index need not be sorted.
 *   seg: int32 buffer of yolat_segments_ints(M,S): segptr[S+1], perm[M] (rows of segment s are
 *   perm[segptr[s] .. segptr[s+1]) in ascending row order). */
int64_t yolat_segments_ints(int64_t M, int64_t S);
int yolat_segments_build(const int64_t* index, int64_t M, int64_t S, int32_t* seg, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GraphConv('attr_edge_gp2')  ==  AttrRelativeEdgeConvGlobalPool2 (torch_vertex.py:288-341).
 * Parameter block mirrors the module's state dict: nn.0 / nn.1 / nn.3 / nn.4 / lin_r / mlp_node.0 / .1
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* w;        /* gamma  [C] */
  const float* b;        /* beta   [C] */
  float* running_mean;   /* [C], updated in training mode (momentum 0.1) */
  float* running_var;    /* [C], updated with the unbiased variance       */
  int64_t* num_batches_tracked; /* may be NULL */
} yolat_bn;

typedef struct {
  const float* w1; const float* b1; yolat_bn bn1;   /* nn.0 [C, 2Cin+4], nn.1 */
  const float* w2; const float* b2; yolat_bn bn2;   /* nn.3 [C, C],      nn.4 */
  const float* wr; const float* br;                 /* lin_r [C, Cin]         */
  const float* wn; const float* bnode; yolat_bn bnn;/* mlp_node.0 [C, Cn], mlp_node.1 */
} yolat_gp2_params;

typedef struct {         /* gradient outputs, same shapes as the parameters; any pointer may be NULL */
  float* w1; float* b1; float* bn1_w; float* bn1_b;
  float* w2; float* b2; float* bn2_w; float* bn2_b;
  float* wr; float* br;
  float* wn; float* bnode; float* bnn_w; float* bnn_b;
} yolat_gp2_grads;

int64_t yolat_gp2_tape_floats(int64_t N, int64_t E, int Cin, int Cn, int C);
/* Tape size for a given `training` mode word of yolat_gp2_fwd / _bwd.  Training-mode calls at C = 64 run tape-free:
 * the tape holds node-level tensors only (node-branch pre-activation, BN statistic blocks, P | Q) and the backward
 * recomputes every per-edge quantity on chip (csrc/edge_bwd.cu); other calls keep z1 / z2 [E, C] for the backward.
 * yolat_gp2_tape_floats(...) == yolat_gp2_tape_floats_mode(..., 0) is an upper bound valid for every mode. */
int64_t yolat_gp2_tape_floats_mode(int64_t N, int64_t E, int Cin, int Cn, int C, int mode);
int64_t yolat_gp2_fwd_ws_floats(int64_t N, int64_t E, int Cin, int Cn, int C);
int64_t yolat_gp2_bwd_ws_floats(int64_t N, int64_t E, int Cin, int Cn, int C);

/* forward (torch_vertex.py:319-337).  x [N,Cin], x_node [N,Cn], attr [E,4] in ORIGINAL edge order,
 * edge_weight [E] or NULL (`norm`, :337).  out [N,C] = mean-aggregated messages + lin_r(x);
 * xnode_out [N,C] = mlp_node(x_node).
 * training: bit 0 (YOLAT_GP2_TRAINING) = batch statistics + running-stat update; bit 1 (YOLAT_GP2_NO_TAPE) =
 * forward only -- the per-edge activations z1 / z2 are not written to the tape (no [E,C] tensor reaches HBM)
 * and yolat_gp2_bwd must not be called on that tape. */
#define YOLAT_GP2_TRAINING 1
#define YOLAT_GP2_NO_TAPE 2
int yolat_gp2_fwd(const yolat_gp2_params* p, int Cin, int Cn, int C,
                  const float* x, int64_t ldx, const float* x_node, int64_t ldxn,
                  const float* attr, const float* edge_weight,
                  const int32_t* graph, int64_t N, int64_t E, int training,
                  float* out, int64_t ldo, float* xnode_out, int64_t ldxo,
                  float* tape, int64_t tape_floats, float* ws, int64_t ws_floats, void* stream);

/* backward of the above.  g_out [N,C], g_xnode [N,C] (either may be NULL = zero).  dx [N,Cin] /
 * dx_node [N,Cn] may be NULL (head layer: inputs are data).  accumulate_dx != 0 adds into dx / dx_node. */
int yolat_gp2_bwd(const yolat_gp2_params* p, const yolat_gp2_grads* g, int Cin, int Cn, int C,
                  const float* x, int64_t ldx, const float* x_node, int64_t ldxn,
                  const float* attr, const float* edge_weight,
                  const int32_t* graph, int64_t N, int64_t E, int training,
                  const float* g_out, int64_t ldgo, const float* g_xnode, int64_t ldgx,
                  float* dx, int64_t lddx, float* dx_node, int64_t lddxn, int accumulate_dx,
                  const float* tape, float* ws, int64_t ws_floats, void* stream);

/* ------------------------------------------------------------------------------------------------
 * The dense building block under every nn.Linear on the path (gcn_lib/sparse/torch_nn.py:58 and its autograd):
 * C[M,N] (+)= op(A) op(B) (+ bias[n]) on the tcgen05 tensor cores with fp32-accurate 3xTF32 splitting.
 *   mode 0 (NT): A[m*lda+k], B[n*ldb+k]   y  = x W^T + b
 *   mode 1 (NN): A[m*lda+k], B[k*ldb+n]   dx = dy W
 *   mode 2 (TN): A[k*lda+m], B[k*ldb+n]   dW = dy^T x   (K = number of rows reduced over)
 * ---------------------------------------------------------------------------------------------- */
int64_t yolat_gemm_ws_floats(int mode, int64_t M, int64_t N, int64_t K);
int yolat_gemm(int mode, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
               int64_t M, int64_t N, int64_t K, const float* bias, int accumulate,
               float* ws, int64_t ws_floats, void* stream);

/* ------------------------------------------------------------------------------------------------
 * One [Lin, BN?, ReLU?] stage of gcn_lib.sparse.MLP (torch_nn.py:50-71): y = act(bn(x W^T + b)).
 * flags: bit0 = has BatchNorm, bit1 = has ReLU, bit2 = training.
 * tape: z [M,Nout] (pre-BN) + 4*Nout statistics when BN is present.
 * ---------------------------------------------------------------------------------------------- */
#define YOLAT_MLP_BN 1
#define YOLAT_MLP_RELU 2
#define YOLAT_MLP_TRAINING 4
int64_t yolat_mlp_tape_floats(int64_t M, int K, int Nout, int flags);
int64_t yolat_mlp_ws_floats(int64_t M, int K, int Nout, int flags);
int yolat_mlp_fwd(const float* x, int64_t ldx, int64_t M, int K, const float* w, const float* b, int Nout,
                  const yolat_bn* bn, int flags, float* y, int64_t ldy,
                  float* tape, int64_t tape_floats, float* ws, int64_t ws_floats, void* stream);
int yolat_mlp_bwd(const float* x, int64_t ldx, int64_t M, int K, const float* w, int Nout,
                  const yolat_bn* bn, int flags, const float* gy, int64_t ldgy,
                  float* dx, int64_t lddx, int accumulate_dx, float* dw, float* db, float* dgamma, float* dbeta,
                  const float* tape, float* ws, int64_t ws_floats, void* stream);

/* ------------------------------------------------------------------------------------------------
 * torch_scatter.scatter(src, index, dim=0, reduce='mean'|'max')  (architecture3cc_rpn_gp_iter2.py:67,122)
 * over prepared segments.  max: arg [S,C] int32 = row index of the maximum (first occurrence), -1 and
 * value 0 for an empty segment; backward routes to that single row.
 * ---------------------------------------------------------------------------------------------- */
int yolat_segment_mean_fwd(const float* src, int64_t lds, int64_t M, int C, const int32_t* seg, int64_t S,
                           float* out, int64_t ldo, void* stream);
int yolat_segment_mean_bwd(const float* g, int64_t ldg, int64_t M, int C, const int32_t* seg, int64_t S,
                           float* dsrc, int64_t ldd, int accumulate, void* stream);
int yolat_segment_max_fwd(const float* src, int64_t lds, int64_t M, int C, const int32_t* seg, int64_t S,
                          float* out, int64_t ldo, int32_t* arg, void* stream);
int yolat_segment_max_bwd(const float* g, int64_t ldg, int64_t M, int C, int64_t S, const int32_t* arg,
                          float* dsrc, int64_t ldd, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------
 * fusion_block + cat + scatter-max fused (architecture3cc_rpn_gp_iter2.py:62-63,122):
 *   pooled[s, 0:F]   = max_{rows of s} relu(bn(feats W^T + b))      F = 1024
 *   pooled[s, F:F+K] = max_{rows of s} feats                         K = fusion_dims
 * without ever building out_feat [N, F+K].
 * ---------------------------------------------------------------------------------------------- */
int64_t yolat_fusemax_tape_floats(int64_t M, int K, int F, int64_t S);
int64_t yolat_fusemax_ws_floats(int64_t M, int K, int F, int64_t S);
int yolat_fusemax_fwd(const float* feats, int64_t ldf, int64_t M, int K, const float* w, const float* b, int F,
                      const yolat_bn* bn, int training, const int32_t* seg, int64_t S,
                      float* pooled, int64_t ldp, float* tape, int64_t tape_floats,
                      float* ws, int64_t ws_floats, void* stream);
int yolat_fusemax_bwd(const float* feats, int64_t ldf, int64_t M, int K, const float* w, int F,
                      const yolat_bn* bn, int training, const int32_t* seg, int64_t S,
                      const float* g_pooled, int64_t ldg,
                      float* dfeats, int64_t lddf, int accumulate_dfeats,
                      float* dw, float* db, float* dgamma, float* dbeta,
                      float* tape, float* ws, int64_t ws_floats, void* stream);

/* ------------------------------------------------------------------------------------------------
 * CrossEntropyLoss(mean) (architecture3cc_rpn_gp_iter2.py:363,376).  prob [B,ncls] is the tape.
 * bwd: dlogits = (softmax - onehot) * (*g_loss) / B, g_loss a DEVICE scalar (no host sync).
 * Labels outside [0,ncls) contribute 0 to the loss and get a zero gradient row.
 * ---------------------------------------------------------------------------------------------- */
int yolat_softmax_xent_fwd(const float* logits, int64_t ldl, int64_t B, int ncls, const int64_t* labels,
                           float* loss, float* prob, float* ws, int64_t ws_floats, void* stream);
int yolat_softmax_xent_bwd(const float* prob, int64_t B, int ncls, const int64_t* labels, const float* g_loss,
                           float* dlogits, int64_t ldd, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused multi-tensor Adam (replaces torch.optim.Adam(...).step(), cad_recognition/train.py:212,284: one L2-regularised
 * Adam update of every parameter tensor, ~100 eager launches per step in the reference).
 *   table: n_chunks x 4 device addresses (param, grad, exp_avg, exp_avg_sq of the chunk's first element),
 *   count: elements of each chunk (<= yolat_adam_chunk()),
 *   state: 3 doubles on the device: [0] step count t (advanced by the call), [1] lr / (1 - beta1^t),
 *          [2] 1 / sqrt(1 - beta2^t)  -- device-resident, so the call can be captured in a CUDA graph.
 *   grad <- grad * grad_scale + weight_decay * param before the moment updates (torch semantics, not AdamW).
 * ---------------------------------------------------------------------------------------------- */
int yolat_adam_chunk(void);
int yolat_adam_step(const uint64_t* table, const int32_t* count, int64_t n_chunks, double* state, double lr, double beta1,
                    double beta2, double eps, double weight_decay, double grad_scale, void* stream);
/* Same update with the hyper-parameters on the device: state is 9 doubles, [3..8] = lr, beta1, beta2, eps,
 * weight_decay, grad_scale, read by the kernels at execution time -- a captured step follows torch's StepLR
 * (train.py:214) and restored checkpoints without re-capture; the host rewrites state[3..8] when they change. */
int yolat_adam_step_dev(const uint64_t* table, const int32_t* count, int64_t n_chunks, double* state, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Proposal slicing of SparseCADGCN.predict on the device (architecture3cc_rpn_gp_iter2.py:153-234: python range
 * lists, the old->new node dict, the per-edge re-indexing loop, the per-node bbox_idx renumbering loop).
 *   yolat_expand_ranges: out[t] = start[k] + (t - prefix[k]) for the range k holding t; prefix = exclusive prefix
 *                        sums of the K range lengths (prefix[0] = 0), total = their sum.  All device int64.
 *   yolat_slice_graph:   pos_idx [Np] / edge_idx [Ep] = expanded node / edge selections; edge [E_all,2] int64;
 *                        edge_out[e] = (new index of edge[edge_idx[e]][0], ...[1]) (-1 if the endpoint is not selected),
 *                        bbox_idx_out[t] = number of changes of bbox_idx[pos_idx[.]] in (0, t]  (dense renumbering);
 *                        ws: yolat_slice_graph_ints(N_all, Np) int32.
 * ---------------------------------------------------------------------------------------------- */
int yolat_expand_ranges(const int64_t* start, const int64_t* prefix, int64_t K, int64_t total, int64_t* out, void* stream);
int64_t yolat_slice_graph_ints(int64_t N_all, int64_t Np);
int yolat_slice_graph(const int64_t* pos_idx, int64_t Np, const int64_t* edge_idx, int64_t Ep, const int64_t* edge,
                      int64_t E_all, int64_t N_all, const int64_t* bbox_idx, int32_t* ws, int64_t* edge_out,
                      int64_t* bbox_idx_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * One-layer edge convolutions of the same family -- GraphConv(conv = 'edge' | 'attr_edge' | 'attr_edge_gp')
 * (gcn_lib/sparse/torch_vertex.py:738-747; EdgConv :546-557 / WeightedRelativeEdgeConv :427-484, AttrEdgConv :560-573 /
 * AttrRelativeEdgeConv :219-286, EdgConvGlobalPool :575-590 / AttrRelativeEdgeConvGlobalPool :343-425):
 *     out[i] (pre-filled by the caller with lin_r(x) etc.) += mean_{e -> i} w_e relu(bn(W1 [x_i | x_j - x_i | attr_e] + b1))
 * w1 [C, 2Cin+4] is the recipe's Linear weight embedded in that column layout (zero blocks for parts the recipe's
 * concat does not have); 'multilayer_edge' (:593-605, two stages) maps onto yolat_gp2_* the same way.  C in {32,64,128}.
 * Tape: z1 [E, C] + the BN statistic block (yolat_edge1_tape_floats); ws: yolat_edge1_ws_floats.
 * ---------------------------------------------------------------------------------------------- */
int64_t yolat_edge1_tape_floats(int64_t N, int64_t E, int Cin, int C);
int64_t yolat_edge1_ws_floats(int64_t N, int64_t E, int Cin, int C);
int yolat_edge1_fwd(const float* w1, const float* b1, const yolat_bn* bn, int Cin, int C, const float* x, int64_t ldx,
                    const float* attr, const float* edge_weight, const int32_t* graph, int64_t N, int64_t E, int training,
                    float* out, int64_t ldo, float* tape, int64_t tape_floats, float* ws, int64_t ws_floats, void* stream);
int yolat_edge1_bwd(const float* w1, const yolat_bn* bn, int Cin, int C, const float* x, int64_t ldx, const float* attr,
                    const float* edge_weight, const int32_t* graph, int64_t N, int64_t E, int training, const float* g_out,
                    int64_t ldgo, float* dx, int64_t lddx, int accumulate_dx, float* dw1, float* db1, float* dgamma,
                    float* dbeta, const float* tape, float* ws, int64_t ws_floats, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Per-image index offsets of a collated batch (cad_recognition/train.py:238-258: python loop over the batch items on
 * host tensors).  edge [E,2] += node offset of the edge's image, bbox_idx [N] += proposal offset of the node's image,
 * in place, one launch.  tab: [4][G+1] int64 prefix sums (device): edge slices | pos slices | bbox_idx slices |
 * labels slices, as train.collate returns them in `slices` (train.py:123-171).
 * ---------------------------------------------------------------------------------------------- */
int yolat_batch_offsets(int64_t* edge, int64_t E, int64_t* bbox_idx, int64_t N, const int64_t* tab, int64_t G, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Proposal enumeration of the reference Dataset on the device (Datasets/graph_dict3.py:309-789,
 * SESYDFloorPlan._get_proposal with do_mixup off; csrc/proposals.cu).  One image per call.
 *   in:  the image's graph as the Dataset's pickle holds it: pos [n_all,2] f64, is_control / is_super [n_all] u8,
 *        connected components as CSR (cc_ptr [ncc+1], cc_idx [cc_total], original node ids), shape / super edges
 *        [E,2] / [Es,2] int64 (original node ids) with their attribute rows [E,A] / [Es,As] f64, ground-truth boxes
 *        [G,4] f64 + labels [G] int64; sampling_step = bbox_sampling_step, n_classes (label of "no object" =
 *        n_classes - 1), normalize_bbox (graph_dict3.py:53).  All pointers are device pointers.
 *   yolat_proposals_ws_bytes: workspace size in BYTES for these sizes (reads only the sizes); -1 = unsupported.
 *   yolat_proposals_count:    everything up to the proposals' sizes; writes totals[YOLAT_PROP_TOTALS] (device int64):
 *        node / shape-edge / super-edge / proposal counts of the outputs, the number of non-control nodes, an error
 *        mask (YOLAT_PROP_ERR_*; the conditions under which the reference raises) and the first offending component.
 *   yolat_proposals_fill:     writes the outputs (sized from totals by the caller) out of the same workspace.
 *   out: pos [n,2] f64, is_super [n] u8, bbox_idx [n] int64, edge [e,2] / edge_super [es,2] int64, e_attr [e,A] /
 *        e_attr_super [es,As] f64, labels / has_obj [B] int64, bbox / bbox_targets [B,4] f64, stat_feats [B,13] f64,
 *        slice_pos / slice_edge / slice_super / slice_bbox [B+1] int64 (the subcluster_slice_* lists of :371-374),
 *        cc_table [ncc,3] int64 = (first proposal, number of proposals, root proposal) of every component, from which
 *        the caller builds the idxTree roots of :730-768.
 * Proposals of one component are emitted in first-occurrence order of the reference's window walk (the reference's
 * own order is CPython's `set` iteration order); everything else is the reference's output.
 * ---------------------------------------------------------------------------------------------- */
#define YOLAT_PROP_TOTALS 8
#define YOLAT_PROP_T_NODES 0
#define YOLAT_PROP_T_EDGES 1
#define YOLAT_PROP_T_SUPER 2
#define YOLAT_PROP_T_BOXES 3
#define YOLAT_PROP_T_ERR 4
#define YOLAT_PROP_T_ERR_CC 5
#define YOLAT_PROP_T_NODES_IN 6
#define YOLAT_PROP_ERR_ZERO_STEP 1ull    /* a component without x or y extent: np.arange(.., step 0) raises ValueError (:470-476) */
#define YOLAT_PROP_ERR_NO_GT 2ull        /* 'cc has no intersect gt bbox' -> SystemExit (:574-576)              */
#define YOLAT_PROP_ERR_NO_PROPOSAL 4ull  /* a component none of whose sets survives: np.argmax([]) raises (:728) */
#define YOLAT_PROP_ERR_CONTROL_REF 8ull  /* an edge / component names a control point: o2n KeyError (:332-348)    */
#define YOLAT_PROP_ERR_CC_OVERLAP 16ull  /* a node listed in two components (not a partition)                     */
#define YOLAT_PROP_ERR_LIMIT 32ull       /* component > 2^20-2 nodes, > 2^23-1 edges, or too many grid lines       */
#define YOLAT_PROP_ERR_INDEX 64ull       /* node id outside [0, n_all)                                            */
typedef struct {
  const double* pos; const uint8_t* is_control; const uint8_t* is_super; int64_t n_all;
  const int64_t* cc_ptr; const int64_t* cc_idx; int64_t ncc; int64_t cc_total;
  const int64_t* edge; const double* e_attr; int64_t E; int32_t A;
  const int64_t* edge_super; const double* e_attr_super; int64_t Es; int32_t As;
  const double* gt_bbox; const int64_t* gt_labels; int64_t G;
  int32_t sampling_step; int32_t n_classes; int32_t normalize_bbox;
} YolatProposalIn;
typedef struct {
  double* pos; uint8_t* is_super; int64_t* bbox_idx;
  int64_t* edge; int64_t* edge_super; double* e_attr; double* e_attr_super;
  int64_t* labels; int64_t* has_obj; double* bbox; double* bbox_targets; double* stat_feats;
  int64_t* slice_pos; int64_t* slice_edge; int64_t* slice_super; int64_t* slice_bbox; int64_t* cc_table;
} YolatProposalOut;
int64_t yolat_proposals_ws_bytes(const YolatProposalIn* in);
int yolat_proposals_count(const YolatProposalIn* in, void* ws, int64_t ws_bytes, int64_t* totals, void* stream);
int yolat_proposals_fill(const YolatProposalIn* in, void* ws, int64_t ws_bytes, const YolatProposalOut* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YOLAT_B200_H_ */
