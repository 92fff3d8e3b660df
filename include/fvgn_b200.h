/* fvgn_b200 -- C-ABI of the B200-native Gen-FVGN hot path (libfvgn_b200.so).
 *
 * The reference (Litianyu141/Gen-FVGN-steady) is pure Python and has no FFI seam; the seam is the
 * Python module API of src/FVMmodel, which gen_fvgn_steady_b200/FVMmodel mirrors.  This header is
 * the layer directly below that mirror: every entry point replaces the group of torch /
 * torch_scatter / PyG calls cited beside it (paths relative to the reference repo root).
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers; the caller (PyTorch) owns
 *     every buffer including workspaces; functions enqueue on `stream` (a cudaStream_t passed as
 *     void*) and never allocate, synchronise or touch global state.
 *   - float tensors are fp32 row-major; index tensors are int32.
 *   - return value: FVGN_OK or a negative FVGN_ERR_* code (the Python wrapper raises RuntimeError).
 */
#ifndef FVGN_B200_H
#define FVGN_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FVGN_OK 0
#define FVGN_ERR_SHAPE (-1)
#define FVGN_ERR_ALIGN (-2)
#define FVGN_ERR_UNSUPPORTED (-3)
#define FVGN_ERR_LAUNCH (-4)
#define FVGN_ERR_NULL (-5)

/* library / build identification; bit0 of the return = built for sm_100a, bit1 = tcgen05 kernels present */
int fvgn_version(void);

/* ------------------------------------------------------------------ CSR segmented reductions */
#define FVGN_ADJ_ACCUMULATE 1      /* dst += result                               */
#define FVGN_ADJ_DIV_DST_BY_DEG 2  /* result /= max(deg(row),1)   (scatter_mean)   */
#define FVGN_ADJ_DIV_SRC_BY_DEG 4  /* each term /= max(deg(src),1) (its transpose) */
#define FVGN_ADJ_SIMPLE_KERNEL 8   /* testing: use the one-row-per-warp kernel (same bits, slower) */
/* dst[i,:] = sum_{t in [ptr[i],ptr[i+1])} src[nbr[t],:], width in {64,128}.
 * Replaces src/FVMmodel/Models/FVGN/blocks.py:92-99 (x[senders/receivers] + scatter_add) and
 * blocks.py:44-51 (scatter_mean); backward of both is the same call (Adj symmetric). */
int fvgn_adj_reduce(const float* src, const int32_t* ptr, const int32_t* nbr, float* dst, int64_t n_rows,
                    int32_t width, int32_t flags, void* stream);
/* dst[i,0:W] = sum over incidence entries code=(edge*2+role) of src[edge, role*W:(role+1)*W], src is [E,2W].
 * Replaces blocks.py:24-42 (chunk + cat + scatter_add); also the deterministic transpose of the
 * agg[senders]/agg[receivers] gathers of blocks.py:101-107 in backward. */
int fvgn_inc_reduce(const float* src, const int32_t* ptr, const int32_t* code, float* dst, int64_t n_rows,
                    int32_t width, void* stream);
/* Typed variants for the 16-bit tensor-core modes: src / dst element type chosen by FVGN_T_* (accumulation is always
 * fp32 in CSR order). */
#define FVGN_T_F32 0
#define FVGN_T_BF16 1
#define FVGN_T_F16 2
int fvgn_adj_reduce_t(const void* src, int32_t src_type, const int32_t* ptr, const int32_t* nbr, void* dst, int32_t dst_type,
                      int64_t n_rows, int32_t width, int32_t flags, void* stream);
int fvgn_inc_reduce_t(const void* src, int32_t src_type, const int32_t* ptr, const int32_t* code, void* dst,
                      int32_t dst_type, int64_t n_rows, int32_t width, void* stream);

/* ------------------------------------------------------------------ plan-time kernels
 * Stable CSR of a scatter entry list: ptr[n_rows+1], perm[n_entries] = the entries grouped by destination row, inside a
 * row in their original order -- the order torch_scatter / index_add_ visit them (blocks.py:24-51,84-99;
 * FVgrad.py:264-325; Load_mesh/Graph_loader.py builds no CSR: the reference scatters by atomics).  dest: int32 or int64
 * entries in [0, n_rows) (out-of-range entries are skipped).  workspace: fvgn_csr_build_workspace_bytes(n_rows) bytes.
 * Deterministic: counts and fill use integer atomics, then every row orders its own entries by entry number. */
int64_t fvgn_csr_build_workspace_bytes(int64_t n_rows);
int fvgn_csr_build(const void* dest, int32_t dest_is_int64, int64_t n_entries, int64_t n_rows, int32_t* ptr, int32_t* perm,
                   void* workspace, void* stream);
/* *out_u64 += order-sensitive 64-bit content hash of n_words 32-bit words (the key of the topology plan cache: a loader
 * hands a NEW batch object with the SAME index tensors every step, Graph_loader.py:830-1006) */
int fvgn_hash_words(const void* data, int64_t n_words, int64_t seed, void* out_u64, void* stream);
/* dst[i, 0:width] = sum_{t in [ptr[i], ptr[i+1])} w[t] * src[idx[t], 0:width]  (idx NULL: identity, w NULL: 1), entries in
 * CSR order, products rounded before the add (= a sequential index_add_ of src * w); mode 0 sum, 1 mean over the row's
 * entries (count clamped to 1), 2 divided by the row's sum of w.  Deterministic replacement of the scatter / index_add_
 * calls of the stand-alone FV API: utils/utilities.py:16-61, FVInterpolation.py:36-109,218-265, FVgrad.py:183-232. */
int fvgn_csr_weighted_sum(const float* src, int32_t width, int32_t ld, const int32_t* ptr, const int32_t* idx, const float* w,
                          int32_t mode, float* dst, int64_t n_rows, void* stream);
/* the same on fp64 data (setup-time callers: compute_normal_matrix assembles the WLSQ moment matrices in fp64) */
int fvgn_csr_weighted_sum_f64(const void* src, int32_t width, int32_t ld, const int32_t* ptr, const int32_t* idx, const void* w,
                              int32_t mode, void* dst, int64_t n_rows, void* stream);

/* ------------------------------------------------------------------ fused MLP blocks */
#define FVGN_MLP_EDGE 0     /* EdgeBlock  blocks.py:101-111 + EPD.py:170-175,186 : in=[agg[s]|agg[r]|e], K1=384, LN, +e   */
#define FVGN_MLP_NODE 1     /* NodeBlock  blocks.py:54 + EPD.py:163-168,185      : in=[a2|x],           K1=192, LN, +x   */
#define FVGN_MLP_ENC_NODE 2 /* Encoder.nb_encoder EPD.py:118                      : in=x[N,12],          K1=12,  LN       */
#define FVGN_MLP_ENC_EDGE 3 /* importer.py:54-78 + Encoder.eb_encoder EPD.py:119  : in=[xn[s]-xn[r]|dpos|norm], K1=15, LN */
#define FVGN_MLP_DEC 4      /* Decoder EPD.py:215-219                             : in=x[N,128], K1=128, out=3, no LN    */
#define FVGN_MLP_NO_RESIDUAL 1 /* desc.flags */
#define FVGN_MLP_RESIDUAL_FROM_SHADOW 2 /* desc.flags, forward, EDGE / NODE, tensor-core modes ("16-bit latent streams"): the
                                         * residual row is read from the 16-bit shadow in1h instead of the fp32 stream in1 (in1
                                         * may be NULL); out_res may be NULL (only out_resh is written) */
#define FVGN_PREC_FP32 0    /* SIMT fp32 FMA (parity mode, rel 1e-5 vs the reference's CPU fp32)   */
#define FVGN_PREC_BF16 1    /* tcgen05.mma kind::f16 bf16 operands, fp32 accumulate in TMEM (throughput mode) */
#define FVGN_PREC_F16 2     /* tcgen05.mma kind::f16 IEEE-half operands: the 11-bit significand of TF32, the arithmetic the
                               reference's GPU path uses for its Linear layers (src/pre_train_Adam.py:29), at the bytes of
                               bf16; fp32 accumulate; same kernels, same layouts ("bf16" below reads "16-bit") */

typedef struct fvgn_mlp_desc {
  int32_t mode;      /* FVGN_MLP_*  */
  int32_t precision; /* FVGN_PREC_* */
  int64_t rows;      /* edges (EDGE, ENC_EDGE) or nodes */
  const float* in0;  /* EDGE: agg[N,128]  NODE: a2[N,64]  ENC_NODE: xn[N,12]  ENC_EDGE: xn[N,12]  DEC: x[N,128] */
  const float* in1;  /* EDGE: e[E,128]    NODE: x[N,128]  ENC_EDGE: pos[N,2]  else NULL */
  const int32_t* idx_s; /* EDGE / ENC_EDGE: senders   [rows] */
  const int32_t* idx_r; /* EDGE / ENC_EDGE: receivers [rows] */
  /* parameters, reference state_dict layout: net.0.{0,2,4}.{weight,bias}, net.1.{weight,bias} */
  const float* w1; const float* b1; const float* w2; const float* b2; const float* w3; const float* b3;
  const float* ln_g; const float* ln_b; /* NULL for DEC */
  const void* w_bf16; /* FVGN_PREC_BF16 only: packed bf16 operand image made by fvgn_mlp_pack_weights */
  int32_t flags;      /* FVGN_MLP_NO_RESIDUAL: stand-alone EdgeBlock/NodeBlock (no "+ e" / "+ x" in fwd and bwd) */
  int32_t reserved0;
  /* forward outputs */
  float* out;     /* y = MLP(in)  [rows,128] ([rows,3] for DEC); may be NULL for EDGE/NODE if only out_res is wanted */
  float* out_res; /* EDGE: e + y, NODE: x + y; NULL otherwise */
  /* backward inputs */
  const float* d_out;    /* grad wrt out_res (EDGE/NODE) or out (ENC_*, DEC).  EDGE, tensor-core modes: may be NULL when d_gather /
                          * d_gatherh is given (a block whose e' / e + e' only feed its node block: zero upstream gradient) */
  const float* d_gather; /* EDGE: d_a1[N,64]; [d_a1[s]|d_a1[r]] is added to d_out (transpose of blocks.py:24-42) */
  /* backward outputs */
  float* d_in0; /* EDGE: [E,256] = d(agg[s]) | d(agg[r]);  NODE: d_a2[N,64];  DEC: d_x[N,128];  ENC_*: NULL */
  float* d_in1; /* EDGE: d_e[E,128] = d_out + dX[:,256:384];  NODE: d_x[N,128] = d_out + dX[:,64:192] */
  float* partials;     /* [n_partials, fvgn_mlp_param_count(mode)] per-CTA weight-gradient partial sums */
  int32_t n_partials;  /* = grid size of the backward kernel, from fvgn_mlp_bwd_partials() */
  float* d_params;     /* [param_count] flat: w1,b1,w2,b2,w3,b3,ln_g,ln_b (deterministic reduction of partials) */
  void* workspace;     /* backward scratch of fvgn_mlp_bwd_workspace_bytes() bytes (bf16 dZ1 tile images), 1024-B aligned */
  /* FVGN_PREC_BF16 only: bf16 tile images of the first pre-activation Z1 = X W1^T + b1 (one 32 KB pre-swizzled
   * [128 x 128] tile per 128 rows, fvgn_mlp_bwd_workspace_bytes() bytes, 1024-B aligned).  The forward writes it
   * when non-NULL; the backward REQUIRES it (it replaces the recomputation of layer 1 from the block inputs). */
  void* z1_img;
  /* FVGN_PREC_BF16 only: bf16 row-major shadows of the operands / results (2 bytes per element, 16-B aligned rows).
   * The tensor-core kernels read their layer-1 operands from the shadows (no conversion, half the gather bytes) and
   * write the shadow of every latent they produce; fp32 is kept only for the residual streams x / e and their
   * gradients.  in0h: EDGE aggh[N,128]  NODE a2h[N,64]  DEC xh[N,128]  (ENC_*: unused, in0 fp32 is read).
   *             in1h: EDGE eh[E,128]    NODE xh[N,128]. */
  const void* in0h; const void* in1h;
  void* outh;      /* bf16 copy of `out` (EDGE: e' for the node aggregation; ENC_*: shadow of the encoded latent); optional */
  void* out_resh;  /* bf16 shadow of out_res; optional */
  void* d_in0h;    /* backward, 16-bit destination instead of the fp32 d_in0: EDGE [E,256] = d(agg[s]) | d(agg[r]); NODE d_a2[N,64];
                    * DEC d_x[N,128] */
  const void* d_gatherh; /* EDGE backward: bf16 d_a1[N,64] gathered instead of the fp32 d_gather */
  /* backward, optional device scalar: every parameter gradient is multiplied by *grad_unscale when it leaves the
   * deterministic partial reduction (1 / S of the power-of-two gradient pre-scaling the FVGN_PREC_F16 mode applies at the
   * root of the backward pass so that half-precision gradient operands stay in range; exact). */
  const float* grad_unscale;
  /* NODE backward, tensor-core modes, optional: CSR row pointers ptr[N+1] of the node adjacency.  Row i of d_in0h
   * (d_a2, the gradient of the scatter_mean of blocks.py:44-51) leaves divided by max(ptr[i+1] - ptr[i], 1): the
   * transposed mean d_a1 = Adj (D^-1 d_a2) then is a plain adjacency sum (no per-source degree lookups in
   * fvgn_adj_reduce_t; FVGN_ADJ_DIV_SRC_BY_DEG must not be passed as well). */
  const int32_t* d_in0_row_ptr;
  /* EDGE backward, tensor-core modes, optional "node-level layer 1" path (selected by d_aggh != NULL; d_in0 / d_in0h are
   * then not written).  The first 256 input columns of the edge MLP are agg[senders] | agg[receivers]
   * (blocks.py:101-107), so by linearity their part of the layer-1 backward can run per NODE instead of per EDGE:
   *   U_s[i] = sum_{f: s_f = i} dZ1[f],  U_r[i] = sum_{f: r_f = i} dZ1[f]      (incidence CSR order, fp32 sums)
   *   d(agg)[i] = U_s[i] W1[:, 0:128] + U_r[i] W1[:, 128:256]                  -> d_aggh [n_nodes,128] 16-bit
   *   dW1[:, 0:128] = U_s^T agg,  dW1[:, 128:256] = U_r^T agg
   * which removes the [E,256] gradient stream d(agg[s])|d(agg[r]), its incidence reduction and the gathered operand
   * chunks of the edge-level weight gradient.  inc_ptr / inc_code: node incidence CSR (code = edge*2 + role);
   * in0h = aggh is read per node; node_partials: [fvgn_mlp_bwd_node_partials(n_nodes), 128*256] fp32 scratch. */
  const int32_t* inc_ptr; const int32_t* inc_code;
  int64_t n_nodes;
  void* d_aggh;
  float* node_partials;
  int32_t n_node_partials;
  int32_t reserved1;
  void* node_ws;       /* optional, fvgn_mlp_bwd_node_workspace_bytes(n_nodes) bytes, 1024-B aligned: when given, the incidence sums
                        * U_s / U_r are formed by a separate memory-bound kernel as operand tile images (two kernels) */
  /* backward, tensor-core modes, EDGE / NODE / ENC_*, optional "16-bit gradient streams" (FVGN_PREC_F16: the streams carry the
   * power-of-two gradient pre-scaling): d_outh replaces d_out as the upstream gradient [rows,128] (16-bit rows), d_in1h
   * replaces d_in1 as the destination of d_e / d_x = d_out + dX (EDGE / NODE).  Either side may stay fp32. */
  const void* d_outh;
  void* d_in1h;
} fvgn_mlp_desc;

int64_t fvgn_mlp_param_count(int32_t mode);
int32_t fvgn_mlp_bwd_partials(int32_t mode, int32_t precision, int64_t rows); /* number of per-CTA partial buffers to allocate */
int32_t fvgn_mlp_bwd_node_partials(int64_t n_nodes); /* rows of fvgn_mlp_desc.node_partials (node-level layer-1 path) */
int64_t fvgn_mlp_bwd_node_workspace_bytes(int64_t n_nodes); /* bytes of fvgn_mlp_desc.node_ws */
int64_t fvgn_mlp_bwd_workspace_bytes(int32_t mode, int32_t precision, int64_t rows);
int64_t fvgn_mlp_packed_bytes(int32_t mode);
/* fp32 parameters -> 16-bit UMMA operand image in the format of `precision` (FVGN_PREC_BF16 / FVGN_PREC_F16); weights
 * change every optimiser step: call once per step */
int fvgn_mlp_pack_weights(int32_t mode, int32_t precision, const float* w1, const float* w2, const float* w3, void* packed,
                          void* stream);
int fvgn_mlp_forward(const fvgn_mlp_desc* d, void* stream);
int fvgn_mlp_backward(const fvgn_mlp_desc* d, void* stream);

/* ------------------------------------------------------------------ importer prologue / head */
/* Column sums of (x - center[seg])^p over row chunks: chunks[nchunks,3] = (segment, row_begin, row_end);
 * out_partial[nchunks, width].  center may be NULL.  p in {1,2}.  Deterministic replacement of the
 * scatter_mean of importer.py:86-90, Normalizer sums (utils/normalization.py:55-66) and
 * global_add_pool (FVscheme.py:184-188,243-247). */
int fvgn_chunk_colsum(const float* x, int32_t width, int32_t ld, const float* center, int32_t center_ld, int32_t power,
                      const int32_t* chunks, int32_t nchunks, float* out_partial, void* stream);
/* out[seg, :] = sum of out_partial over the chunks of seg (chunk_ptr[nseg+1]) */
int fvgn_chunk_combine(const float* partial, int32_t width, const int32_t* chunk_ptr, int32_t nseg, float* out,
                       void* stream);
/* importer.py:168-176 : uv_old = x[:,0:2]/uvp_dim[batch]; xn[:,0:3]=(x-mean_g)/(std_g+1e-8); xn[:,3:12]=(x-mean)/std */
int fvgn_prologue(const float* x, const int32_t* batch, const float* uvp_dim, const float* gmean, const float* gstd,
                  const float* nmean, const float* nstd, int32_t norm_uvp, float* xn, float* uv_old, int64_t n,
                  void* stream);
/* importer.py:187-201 : uvp = 10 tanh(raw/10); Dirichlet rows <- y, p=0 at PRESS_POINT; uv_hat by integrator
 * (0 explicit, 1 implicit, 2 imex); phi[N,7] = [uvp | uv_hat | uv_old]  (FVscheme.py:643-646) */
int fvgn_head_forward(const float* raw, const float* uv_old, const float* y, const int32_t* node_type, int32_t integrator,
                      float* phi, int64_t n, void* stream);
int fvgn_head_backward(const float* raw, const int32_t* node_type, int32_t integrator, const float* d_phi, float* d_raw,
                       int64_t n, void* stream);

/* ------------------------------------------------------------------ finite-volume loss */
/* Plan-time: fold the fp64 inverse of the per-node 5x5 (2x2 for order 1) moment matrix into per-entry
 * weights q(e) = (A_i^-1 w m(e))[0:nq]; entries are CSR-ordered (row = 'in' node).  moments[nnz,nm].
 * Replaces the run-time row-normalise + torch.linalg.solve of FVgrad.py:335-359. */
int fvgn_wlsq_weights(const float* A, int32_t nm, const int32_t* ptr, const float* moments, int32_t nq, float* q,
                      float* qsum, int64_t n, void* stream);
/* grad[i,c,d] = sum_e q[e,d] (phi[col[e],c] - phi[i,c])   (FVgrad.py:295-325 + :335-359), nc channels, nq in {2,5} */
int fvgn_wlsq_forward(const float* phi, int32_t nc, const int32_t* ptr, const int32_t* col, const float* q, int32_t nq,
                      float* grad, int64_t n, void* stream);
/* d_phi[j,c] (+)= sum_{e in T(j)} sum_d qT[e,d] g[rowT[e],c,d] - sum_d qsum[j,d] g[j,c,d] */
int fvgn_wlsq_backward(const float* g, int32_t nc, const int32_t* tptr, const int32_t* trow, const float* tq,
                       const float* qsum, int32_t nq, float* d_phi, int32_t accumulate, int64_t n, void* stream);

typedef struct fvgn_fv_desc {
  int64_t n_nodes, n_faces, n_cells, n_slots;
  int32_t n_graphs;
  /* fields */
  const float* phi;   /* [N,7]  u,v,p,u_hat,v_hat,u_old,v_old */
  const float* grad;  /* [N,7,2] WLSQ gradient */
  /* geometry / topology (Load_mesh batch layout, int32, slots sorted by cell) */
  const float* pos; const float* y; const int32_t* node_type;
  const int32_t* edge_s; const int32_t* edge_r;
  const float* face_pos; const float* face_area; const int32_t* face_type;
  const float* centroid; const float* cells_area; const int32_t* batch_cell;
  const int32_t* cell_ptr; const int32_t* slot_node; const int32_t* slot_face; const float* slot_unv;
  const int32_t* slot_cell;
  const int32_t* face_slot_ptr; const int32_t* face_slot;   /* face -> slots */
  const int32_t* node_slot_ptr; const int32_t* node_slot;   /* node -> slots */
  const int32_t* inc_ptr; const int32_t* inc_code;          /* node -> (face*2+role) */
  const float* theta; /* [B,9] */ const float* dt; /* [B] */
  /* forward outputs */
  float* res;   /* [C,4] continuity, mom_x, mom_y, sum of squared outlet residuals */
  float* phic;  /* [C,5] cell values of u,v,p,u_old,v_old */
  /* backward */
  const float* coef;  /* [B,4] dL/dres scale: dL/dres[c,k] = coef[b,k]*res[c,k] (k<3), dL/dres[c,3] = coef[b,3] */
  float* d_face;      /* [E,13] scratch: d phi_f[5], d gradphi_f[(0,1,3,4),2] */
  float* d_phi;       /* [N,7]  */
  float* d_grad;      /* [N,7,2] */
  /* residual formulation: 0 = Intergrator.conserved_form (FVscheme.py:50-274, the default), 1 = non_conserved_form
   * (FVscheme.py:276-511: gradient-based continuity, cell-gradient convection / pressure terms, no face BC fix) */
  int32_t form;
  int32_t reserved0;
  float* cell_aux;    /* form 1 only: [C,6] = cell values of u_hat, v_hat and the cell-mean gradient of (u_hat, v_hat);
                         written by fvgn_fv_forward, read by fvgn_fv_backward */
} fvgn_fv_desc;
/* Intergrator.conserved_form FVscheme.py:50-250 (form 0) or non_conserved_form FVscheme.py:276-511 (form 1) with
 * node_to_cell/node_to_face (FVInterpolation.py:36-185) and _fix_face_flux_BC (FVscheme.py:32-48) fused: one pass over cells. */
int fvgn_fv_forward(const fvgn_fv_desc* d, void* stream);
int fvgn_fv_backward(const fvgn_fv_desc* d, void* stream);
/* cell_to_node_2nd_order (FVInterpolation.py:218-265) + BC re-enforcement and re-dimensionalisation
 * (importer.py:223-231): uvp_node[N,3], uvp_cell[C,3] */
int fvgn_fv_outputs(const fvgn_fv_desc* d, const int32_t* batch_node, const float* scale /*[B,3] uvp_dim*sigma*/,
                    int32_t ncn_smooth, float* uvp_node, float* uvp_cell, void* stream);

/* ------------------------------------------------------------------ Transolver_block
 * src/FVMmodel/Models/GraphTransolver/GraphTransolver.py:25-169, heads = 8, dim_head = 16, slice_num = 32 (TransFVGN_v1/v2).
 * The dense projections (in_project_fx/x, to_out, mlp) are fvgn_gemm_tf32 calls (below) in the tensor-core modes and fp32
 * library GEMMs in the parity mode; these entry points replace the
 * broadcast-product + torch_scatter slice / de-slice (:59-90) and the elementwise launches around the GEMMs.
 * chunks[nchunks,3] = (graph id, row begin, row end), rows of a chunk belong to one graph; one CTA per chunk. */
#define FVGN_TS_TOKW 4352   /* per-graph token record: 8*32*16 numerators | 8*32 norms */
#define FVGN_TS_PARAMW 808  /* slice-backward record: dWs[32,16] | dbs[32] | d graph_temperature[8] | colsum dP[256] */
/* number of CTAs (= partial rows) of the row-wise backward kernels for n rows */
int fvgn_ts_row_partials(int64_t n);
/* P[N,256] = [in_project_fx(x) | in_project_x(x)] -> sw[N,256] = softmax(in_project_slice(x_mid)/graph_temperature) (:59-61)
 * and partial[nchunks,4352] = per-chunk sums of sw (x) fx_mid | sw (:62-72); combine with fvgn_chunk_combine. */
int fvgn_ts_slice_forward(const float* P, const float* Ws, const float* bs, const float* temp, const int32_t* chunks,
                          int32_t nchunks, float* sw, float* partial, void* stream);
/* partial[nchunks,4352] = per-chunk sums of sw (x) V | sw for V[N,128] (backward of the de-slice w.r.t. the tokens) */
int fvgn_ts_accumulate(const float* sw, const float* V, const int32_t* chunks, int32_t nchunks, float* partial, void* stream);
/* attention among the slice tokens (:72-81): rec[nb,4352] = token record (numerators | norms) -> tok_out[nb,4096]
 * = softmax(q k^T * scale) v per (graph, head), q / k / v = tok W{q,k,v}^T with tok = num / (norm + 1e-5); wq / wk / wv [16,16] */
int fvgn_ts_token_attention_forward(const float* rec, const float* wq, const float* wk, const float* wv, float scale, int32_t nb,
                                    float* tok_out, void* stream);
/* its autograd: d_rec[nb,4352]; w_partial[nb*8, 768] = per (graph, head) dWq | dWk | dWv, summed by fvgn_chunk_combine */
int fvgn_ts_token_attention_backward(const float* rec, const float* wq, const float* wk, const float* wv, float scale, int32_t nb,
                                     const float* d_tok_out, float* d_rec, float* w_partial, void* stream);
/* out[n, h*16+d] = sum_g sw[n,h,g] tok[graph % tok_mod, h, g, d]  (:83-90); tok rows are tok_ld floats apart */
int fvgn_ts_deslice(const float* sw, const float* tok, int64_t tok_ld, int32_t tok_mod, const int32_t* chunks, int32_t nchunks,
                    float* out, void* stream);
/* autograd of slice + de-slice: d_out = d out_x [N,128], tok_out[.,4096] the attended tokens, d_tok[.,4352] the gradient of
 * the token record (ignored for graphs >= tok_mod: ghost rows of the cell-partition mode) -> dP[N,256], partial[nchunks,808] */
int fvgn_ts_slice_backward(const float* P, const float* sw, const float* d_out, const float* tok_out, const float* d_tok,
                           int32_t tok_mod, const float* Ws, const float* bs, const float* temp, const int32_t* chunks,
                           int32_t nchunks, float* dP, float* partial, void* stream);
/* y = a + bias + res ; z = LayerNorm(y) (to_out bias + residual + ln_2, :163-169); stats[N,2] = (mean, rstd) */
int fvgn_ts_residual_ln_forward(const float* a, const float* bias, const float* res, const float* gamma, const float* beta,
                                float* y, float* z, float* stats, int64_t n, void* stream);
/* d_y = LayerNorm-backward(dz) + d_y_in (nullable) ; partial[fvgn_ts_row_partials(n),512] = dgamma | dbeta | colsum d_y |
 * colsum d_y_in */
int fvgn_ts_residual_ln_backward(const float* dz, const float* y, const float* stats, const float* gamma, const float* d_y_in,
                                 float* d_y, float* partial, int64_t n, void* stream);
/* h[N,256] = GELU(hpre + bias) (MLP.linear_pre, :105,124) ; backward: dhpre = dh * GELU'(hpre + bias),
 * partial[fvgn_ts_row_partials(n),256] = colsum dhpre */
int fvgn_ts_bias_gelu_forward(const float* hpre, const float* bias, float* h, int64_t n, void* stream);
int fvgn_ts_bias_gelu_backward(const float* dh, const float* hpre, const float* bias, float* dhpre, float* partial, int64_t n,
                               void* stream);
/* out[N,128] = a + bias + res (+ 16-bit shadow outh of type outh_type = FVGN_T_BF16 / FVGN_T_F16, nullable)
 * (linear_post bias + residual, :168) */
int fvgn_ts_bias_residual(const float* a, const float* bias, const float* res, float* out, void* outh, int32_t outh_type,
                          int64_t n, void* stream);

/* ------------------------------------------------------------------ dense projections on tcgen05 kind::tf32
 * The Linear layers of the Transolver block (GraphTransolver.py:48-58 in_project_fx / in_project_x, :92-95 to_out, :105-127
 * mlp.linear_pre / linear_post) and their autograd in the arithmetic the reference's GPU scripts select
 * (torch.backends.cuda.matmul.allow_tf32, src/pre_train_Adam.py:29): fp32 operands read by the tensor core as TF32, fp32
 * accumulation.  All matrices fp32 row-major; n, k in {128, 256}.
 *   FVGN_GEMM_NT: C[rows,n] = A[rows,k] B[n,k]^T (+ bias[n]) (+ addend[rows,n])        y  = x W^T + b
 *   FVGN_GEMM_NN: C[rows,n] = A[rows,k] B[k,n]   (+ addend[rows,n])                     dx = dy W (+ residual gradient)
 *   FVGN_GEMM_TN: C[k,n]    = A[rows,k]^T B[rows,n]                                     dW = dy^T x; deterministic: one
 *                 partial [k,n] per CTA over a static row split (partials: [fvgn_gemm_tf32_partials(rows), k*n]), fixed-order sum */
#define FVGN_GEMM_NT 0
#define FVGN_GEMM_NN 1
#define FVGN_GEMM_TN 2
int32_t fvgn_gemm_tf32_partials(int64_t rows);
int fvgn_gemm_tf32(int32_t mode, const float* A, const float* B, const float* bias, const float* addend, float* C, int64_t rows,
                   int32_t n, int32_t k, float* partials, int32_t n_partials, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FVGN_B200_H */
