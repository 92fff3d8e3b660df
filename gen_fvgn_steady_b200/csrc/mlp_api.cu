// C-ABI entry points of the fused MLP blocks: validation + dispatch to the fp32 SIMT kernels
// (mlp_simt.cu) or the bf16 tcgen05 kernels (mlp_tc.cu).
#include "common.cuh"

int fvgn_mlp_forward_simt(const fvgn_mlp_desc* d, void* stream);
int fvgn_mlp_backward_simt(const fvgn_mlp_desc* d, void* stream);
int fvgn_mlp_simt_partials(int64_t rows);
int64_t fvgn_mlp_param_count_impl(int32_t mode);
#ifndef FVGN_EMU
int fvgn_mlp_forward_tc(const fvgn_mlp_desc* d, void* stream);
int fvgn_mlp_backward_tc(const fvgn_mlp_desc* d, void* stream);
int fvgn_mlp_tc_partials(int32_t mode, int64_t rows);
int fvgn_mlp_tc_node_partials(int64_t n_nodes);
int64_t fvgn_mlp_tc_node_ws_bytes(int64_t n_nodes);
int64_t fvgn_mlp_tc_packed_bytes(int32_t mode);
int64_t fvgn_mlp_tc_workspace_bytes(int32_t mode, int64_t rows);
int fvgn_mlp_tc_pack(int32_t mode, int32_t precision, const float* w1, const float* w2, const float* w3, void* packed, void* stream);
#endif

static inline bool is_tc(int32_t precision) { return precision == FVGN_PREC_BF16 || precision == FVGN_PREC_F16; }

extern "C" int fvgn_version(void) {
#ifdef FVGN_EMU
  return 0x0100 << 8;  // emulated build: neither sm_100a nor tcgen05
#else
  return (0x0100 << 8) | 1 | 2;
#endif
}

extern "C" int64_t fvgn_mlp_param_count(int32_t mode) { return fvgn_mlp_param_count_impl(mode); }

extern "C" int32_t fvgn_mlp_bwd_partials(int32_t mode, int32_t precision, int64_t rows) {
#ifndef FVGN_EMU
  if (is_tc(precision)) return fvgn_mlp_tc_partials(mode, rows);
#endif
  (void)mode;
  (void)precision;
  return fvgn_mlp_simt_partials(rows);
}

extern "C" int32_t fvgn_mlp_bwd_node_partials(int64_t n_nodes) {
#ifndef FVGN_EMU
  return fvgn_mlp_tc_node_partials(n_nodes);
#else
  (void)n_nodes;
  return 0;
#endif
}

extern "C" int64_t fvgn_mlp_bwd_node_workspace_bytes(int64_t n_nodes) {
#ifndef FVGN_EMU
  return fvgn_mlp_tc_node_ws_bytes(n_nodes);
#else
  (void)n_nodes;
  return 0;
#endif
}

extern "C" int64_t fvgn_mlp_bwd_workspace_bytes(int32_t mode, int32_t precision, int64_t rows) {
#ifndef FVGN_EMU
  if (is_tc(precision)) return fvgn_mlp_tc_workspace_bytes(mode, rows);
#endif
  (void)mode; (void)precision; (void)rows;
  return 0;
}

extern "C" int64_t fvgn_mlp_packed_bytes(int32_t mode) {
#ifndef FVGN_EMU
  return fvgn_mlp_tc_packed_bytes(mode);
#else
  (void)mode;
  return 0;
#endif
}

extern "C" int fvgn_mlp_pack_weights(int32_t mode, int32_t precision, const float* w1, const float* w2, const float* w3,
                                     void* packed, void* stream) {
#ifndef FVGN_EMU
  if (!is_tc(precision)) return FVGN_ERR_UNSUPPORTED;
  return fvgn_mlp_tc_pack(mode, precision, w1, w2, w3, packed, stream);
#else
  (void)mode; (void)precision; (void)w1; (void)w2; (void)w3; (void)packed; (void)stream;
  return FVGN_ERR_UNSUPPORTED;
#endif
}

static int check_common(const fvgn_mlp_desc* d) {
  if (!d) return FVGN_ERR_NULL;
  if (d->mode < FVGN_MLP_EDGE || d->mode > FVGN_MLP_DEC) return FVGN_ERR_UNSUPPORTED;
  if (d->rows < 0) return FVGN_ERR_SHAPE;
  if (!d->w1 || !d->b1 || !d->w2 || !d->b2 || !d->w3 || !d->b3) return FVGN_ERR_NULL;
  if (d->mode != FVGN_MLP_DEC && (!d->ln_g || !d->ln_b)) return FVGN_ERR_NULL;
  // bf16 mode reads the layer-1 operands of EDGE / NODE / DEC from the bf16 shadows; fp32 in0 / in1 are then optional
  const bool shadow = is_tc(d->precision) &&
                      (d->mode == FVGN_MLP_EDGE || d->mode == FVGN_MLP_NODE || d->mode == FVGN_MLP_DEC);
  if (shadow) {
    if (!d->in0h) return FVGN_ERR_NULL;
    if (d->mode != FVGN_MLP_DEC && !d->in1h) return FVGN_ERR_NULL;
    if (!fvgn_aligned16(d->in0h) || !fvgn_aligned16(d->in1h)) return FVGN_ERR_ALIGN;
  } else {
    if (!d->in0) return FVGN_ERR_NULL;
    if ((d->mode == FVGN_MLP_EDGE || d->mode == FVGN_MLP_ENC_EDGE || d->mode == FVGN_MLP_NODE) && !d->in1) return FVGN_ERR_NULL;
  }
  if ((d->mode == FVGN_MLP_EDGE || d->mode == FVGN_MLP_ENC_EDGE) && (!d->idx_s || !d->idx_r)) return FVGN_ERR_NULL;
  if (!fvgn_aligned16(d->in0) || !fvgn_aligned16(d->in1) || !fvgn_aligned16(d->out) || !fvgn_aligned16(d->out_res) ||
      !fvgn_aligned16(d->d_out) || !fvgn_aligned16(d->d_gather) || !fvgn_aligned16(d->d_in0) ||
      !fvgn_aligned16(d->d_in1) || !fvgn_aligned16(d->ln_g) || !fvgn_aligned16(d->ln_b) || !fvgn_aligned16(d->outh) ||
      !fvgn_aligned16(d->out_resh) || !fvgn_aligned16(d->d_in0h))
    return FVGN_ERR_ALIGN;
  return FVGN_OK;
}

extern "C" int fvgn_mlp_forward(const fvgn_mlp_desc* d, void* stream) {
  if (d && d->rows == 0 && d->mode >= FVGN_MLP_EDGE && d->mode <= FVGN_MLP_DEC) return FVGN_OK;  // empty graph: nothing to do
  int rc = check_common(d);
  if (rc) return rc;
  const bool enc = d->mode == FVGN_MLP_ENC_NODE || d->mode == FVGN_MLP_ENC_EDGE;
  if ((d->mode == FVGN_MLP_EDGE || d->mode == FVGN_MLP_NODE) ? (!d->out_res && !d->out && !d->outh && !d->out_resh)
      : (!d->out && !(enc && is_tc(d->precision) && d->outh))) return FVGN_ERR_NULL;   // encoders: the 16-bit latent alone will do
  const bool res16 = (d->flags & FVGN_MLP_RESIDUAL_FROM_SHADOW) != 0;
  if (res16 && (!is_tc(d->precision) || (d->mode != FVGN_MLP_EDGE && d->mode != FVGN_MLP_NODE) || !d->in1h)) return FVGN_ERR_UNSUPPORTED;
  if (d->out_res && !d->in1 && !res16) return FVGN_ERR_NULL;  // the residual is added from the fp32 stream
  if (d->rows == 0) return FVGN_OK;
  if (d->precision == FVGN_PREC_FP32) return fvgn_mlp_forward_simt(d, stream);
#ifndef FVGN_EMU
  if (is_tc(d->precision)) return fvgn_mlp_forward_tc(d, stream);
#endif
  return FVGN_ERR_UNSUPPORTED;
}

extern "C" int fvgn_mlp_backward(const fvgn_mlp_desc* d, void* stream) {
  if (d && d->rows == 0 && d->mode >= FVGN_MLP_EDGE && d->mode <= FVGN_MLP_DEC) {  // empty graph: zero parameter gradients
    if (!d->d_params) return FVGN_ERR_NULL;
#ifndef FVGN_EMU
    if (cudaMemsetAsync(d->d_params, 0, (size_t)fvgn_mlp_param_count_impl(d->mode) * sizeof(float), (cudaStream_t)stream) != cudaSuccess)
      return FVGN_ERR_LAUNCH;
#else
    for (int64_t i = 0; i < fvgn_mlp_param_count_impl(d->mode); ++i) d->d_params[i] = 0.f;
#endif
    return FVGN_OK;
  }
  int rc = check_common(d);
  if (rc) return rc;
  // d_out == NULL: EDGE, tensor-core modes only -- the block's outputs e' / e + e' have no consumer but the node block
  // (last GnBlock of a processor): the upstream gradient is the gathered d_a1 alone
  const bool grad16 = d->d_outh || d->d_in1h;   // 16-bit gradient streams: tensor-core modes, blocks with LayerNorm
  if (grad16 && (!is_tc(d->precision) || d->mode == FVGN_MLP_DEC)) return FVGN_ERR_UNSUPPORTED;
  if (d->d_in1h && d->mode != FVGN_MLP_EDGE && d->mode != FVGN_MLP_NODE) return FVGN_ERR_UNSUPPORTED;
  if (!fvgn_aligned16(d->d_outh) || !fvgn_aligned16(d->d_in1h)) return FVGN_ERR_ALIGN;
  const bool dead_out = !d->d_out && !d->d_outh && d->mode == FVGN_MLP_EDGE && is_tc(d->precision) && (d->d_gather || d->d_gatherh);
  if ((!d->d_out && !d->d_outh && !dead_out) || !d->partials || !d->d_params || d->n_partials < 1) return FVGN_ERR_NULL;
  const bool node_path = d->d_aggh != nullptr;   // EDGE, tensor-core modes: node-level layer-1 backward
  if (node_path) {
    if (d->mode != FVGN_MLP_EDGE || !is_tc(d->precision)) return FVGN_ERR_UNSUPPORTED;
    if (!d->inc_ptr || !d->inc_code || !d->node_partials || !d->in0h) return FVGN_ERR_NULL;
    if (d->n_nodes < 1 || d->n_node_partials != fvgn_mlp_bwd_node_partials(d->n_nodes)) return FVGN_ERR_SHAPE;
    if (!fvgn_aligned16(d->d_aggh) || ((uintptr_t)d->node_ws & 1023)) return FVGN_ERR_ALIGN;
  }
  if ((d->mode == FVGN_MLP_EDGE || d->mode == FVGN_MLP_NODE) && ((!d->d_in0 && !d->d_in0h && !node_path) || (!d->d_in1 && !d->d_in1h)))
    return FVGN_ERR_NULL;
  if ((d->d_in0h || d->d_gatherh) && !is_tc(d->precision)) return FVGN_ERR_UNSUPPORTED;
  if (d->d_in0h && d->mode != FVGN_MLP_EDGE && d->mode != FVGN_MLP_NODE && d->mode != FVGN_MLP_DEC) return FVGN_ERR_UNSUPPORTED;
  if (!fvgn_aligned16(d->d_gatherh)) return FVGN_ERR_ALIGN;
  if (d->mode == FVGN_MLP_DEC && !d->d_in0 && !d->d_in0h) return FVGN_ERR_NULL;
  if (d->precision == FVGN_PREC_FP32) {
    if (d->n_partials != fvgn_mlp_simt_partials(d->rows)) return FVGN_ERR_SHAPE;
    return fvgn_mlp_backward_simt(d, stream);
  }
#ifndef FVGN_EMU
  if (is_tc(d->precision)) return fvgn_mlp_backward_tc(d, stream);
#endif
  return FVGN_ERR_UNSUPPORTED;
}

#ifdef FVGN_EMU
// tests/emu build: the tcgen05 GEMMs do not exist on the host (the Transolver mirror uses them in tensor-core modes only)
extern "C" int32_t fvgn_gemm_tf32_partials(int64_t) { return 0; }
extern "C" int fvgn_gemm_tf32(int32_t, const float*, const float*, const float*, const float*, float*, int64_t, int32_t, int32_t,
                              float*, int32_t, void*) {
  return FVGN_ERR_UNSUPPORTED;
}
#endif
