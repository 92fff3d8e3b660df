// Differentiable finite-volume PDE loss: WLSQ gradient reconstruction, node->face / node->cell
// interpolation, boundary fix, convective / pressure / viscous face flux, surface-integral to cells,
// and the hand-written backward of all of it.  Everything is a deterministic gather (no atomics).
//
// Reference: FVgrad.py:235-367 (node_based_WLSQ), FVInterpolation.py:36-185,218-265,
// FVscheme.py:32-48 (_fix_face_flux_BC), :50-274 (conserved_form).  Index form in SURVEY.md App. A.
#include "common.cuh"

// ================================================================================ WLSQ plan
// q(e) = (A_i^-1 m_w(e))[0:NQ] in fp64 -> fp32, per CSR entry e of row i.  A is the fp32 moment matrix
// the loader stores (Load_mesh.py:257-269); the inverse is Gauss-Jordan with partial pivoting in fp64.
template <int NM>
__global__ void wlsq_weights_kernel(const float* __restrict__ A, const int32_t* __restrict__ ptr,
                                    const float* __restrict__ moments, int nq, float* __restrict__ q,
                                    float* __restrict__ qsum, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double M[NM][2 * NM];
  for (int r = 0; r < NM; ++r)
    for (int c = 0; c < NM; ++c) {
      M[r][c] = (double)A[(size_t)i * NM * NM + r * NM + c];
      M[r][NM + c] = (r == c) ? 1.0 : 0.0;
    }
  bool singular = false;
  for (int p = 0; p < NM; ++p) {
    int piv = p;
    double best = fabs(M[p][p]);
    for (int r = p + 1; r < NM; ++r)
      if (fabs(M[r][p]) > best) { best = fabs(M[r][p]); piv = r; }
    if (best < 1e-300) { singular = true; break; }
    if (piv != p)
      for (int c = 0; c < 2 * NM; ++c) { const double t = M[p][c]; M[p][c] = M[piv][c]; M[piv][c] = t; }
    const double inv = 1.0 / M[p][p];
    for (int c = 0; c < 2 * NM; ++c) M[p][c] *= inv;
    for (int r = 0; r < NM; ++r) {
      if (r == p) continue;
      const double f = M[r][p];
      if (f != 0.0)
        for (int c = 0; c < 2 * NM; ++c) M[r][c] -= f * M[p][c];
    }
  }
  double s[NM];
  for (int d = 0; d < NM; ++d) s[d] = 0.0;
  for (int e = ptr[i]; e < ptr[i + 1]; ++e) {
    for (int d = 0; d < nq; ++d) {
      double v = 0.0;
      if (!singular)
        for (int k = 0; k < NM; ++k) v += M[d][NM + k] * (double)moments[(size_t)e * NM + k];
      q[(size_t)e * nq + d] = (float)v;
      s[d] += (double)(float)v;
    }
  }
  for (int d = 0; d < nq; ++d) qsum[(size_t)i * nq + d] = (float)s[d];
}

extern "C" int fvgn_wlsq_weights(const float* A, int32_t nm, const int32_t* ptr, const float* moments, int32_t nq,
                                 float* q, float* qsum, int64_t n, void* stream) {
  if (n <= 0) return FVGN_OK;
  if (nq < 1 || nq > nm) return FVGN_ERR_SHAPE;
  const unsigned grid = (unsigned)((n + 127) / 128);
  if (nm == 5) {
    FVGN_LAUNCH_SEQ(wlsq_weights_kernel<5>, grid, 128, 0, stream, A, ptr, moments, nq, q, qsum, n);
  } else if (nm == 2) {
    FVGN_LAUNCH_SEQ(wlsq_weights_kernel<2>, grid, 128, 0, stream, A, ptr, moments, nq, q, qsum, n);
  } else {
    return FVGN_ERR_UNSUPPORTED;  // order 3rd/4th: reference diverges (cond 5e13), not on the live path
  }
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

// ================================================================================ WLSQ apply
// 8 lanes per node, lane c owns channel c: the neighbour's phi row is one contiguous 28-B read and
// the per-entry weights are a broadcast.  Row sums run in CSR (= reference scatter) order.
template <int NQ>
__global__ void __launch_bounds__(256) wlsq_fwd_kernel(const float* __restrict__ phi, int nc, const int32_t* __restrict__ ptr,
                                                       const int32_t* __restrict__ col, const float* __restrict__ q,
                                                       float* __restrict__ grad, int64_t n) {
  const int64_t node = (int64_t)blockIdx.x * 32 + threadIdx.x / 8;
  const int c = threadIdx.x % 8;
  if (node >= n || c >= nc) return;
  const float pi = phi[(size_t)node * nc + c];
  float acc[NQ];
#pragma unroll
  for (int k = 0; k < NQ; ++k) acc[k] = 0.f;
  const int e1 = ptr[node + 1];
  for (int e = ptr[node]; e < e1; ++e) {
    const float d = phi[(size_t)col[e] * nc + c] - pi;
#pragma unroll
    for (int k = 0; k < NQ; ++k) acc[k] = fmaf(q[(size_t)e * NQ + k], d, acc[k]);
  }
#pragma unroll
  for (int k = 0; k < NQ; ++k) grad[((size_t)node * nc + c) * NQ + k] = acc[k];
}

template <int NQ>
__global__ void __launch_bounds__(256) wlsq_bwd_kernel(const float* __restrict__ g, int nc, const int32_t* __restrict__ tptr,
                                                       const int32_t* __restrict__ trow, const float* __restrict__ tq,
                                                       const float* __restrict__ qsum, float* __restrict__ d_phi,
                                                       int accumulate, int64_t n) {
  const int64_t node = (int64_t)blockIdx.x * 32 + threadIdx.x / 8;
  const int c = threadIdx.x % 8;
  if (node >= n || c >= nc) return;
  float s = 0.f;
  const int e1 = tptr[node + 1];
  for (int e = tptr[node]; e < e1; ++e) {
    const float* gi = g + ((size_t)trow[e] * nc + c) * NQ;
#pragma unroll
    for (int k = 0; k < NQ; ++k) s = fmaf(tq[(size_t)e * NQ + k], gi[k], s);
  }
  const float* gj = g + ((size_t)node * nc + c) * NQ;
#pragma unroll
  for (int k = 0; k < NQ; ++k) s = fmaf(-qsum[(size_t)node * NQ + k], gj[k], s);
  float* o = d_phi + (size_t)node * nc + c;
  *o = accumulate ? (*o + s) : s;
}

extern "C" int fvgn_wlsq_forward(const float* phi, int32_t nc, const int32_t* ptr, const int32_t* col, const float* q,
                                 int32_t nq, float* grad, int64_t n, void* stream) {
  if (n <= 0) return FVGN_OK;
  if (nc < 1 || nc > 8) return FVGN_ERR_SHAPE;
  const unsigned grid = (unsigned)((n + 31) / 32);
  if (nq == 2) {
    FVGN_LAUNCH_SEQ(wlsq_fwd_kernel<2>, grid, 256, 0, stream, phi, nc, ptr, col, q, grad, n);
  } else if (nq == 5) {
    FVGN_LAUNCH_SEQ(wlsq_fwd_kernel<5>, grid, 256, 0, stream, phi, nc, ptr, col, q, grad, n);
  } else {
    return FVGN_ERR_UNSUPPORTED;
  }
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

extern "C" int fvgn_wlsq_backward(const float* g, int32_t nc, const int32_t* tptr, const int32_t* trow, const float* tq,
                                  const float* qsum, int32_t nq, float* d_phi, int32_t accumulate, int64_t n,
                                  void* stream) {
  if (n <= 0) return FVGN_OK;
  if (nc < 1 || nc > 8) return FVGN_ERR_SHAPE;
  const unsigned grid = (unsigned)((n + 31) / 32);
  if (nq == 2) {
    FVGN_LAUNCH_SEQ(wlsq_bwd_kernel<2>, grid, 256, 0, stream, g, nc, tptr, trow, tq, qsum, d_phi, accumulate, n);
  } else if (nq == 5) {
    FVGN_LAUNCH_SEQ(wlsq_bwd_kernel<5>, grid, 256, 0, stream, g, nc, tptr, trow, tq, qsum, d_phi, accumulate, n);
  } else {
    return FVGN_ERR_UNSUPPORTED;
  }
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

// ================================================================================ face values
struct FaceVals {
  float uvn[2];    // u,v (new) at the face after the BC fix
  float uvh[2];    // u_hat, v_hat at the face after the BC fix
  float p;         // pressure at the face
  float Gn[2][2];  // face gradient of new u,v      (plain average, FVscheme.py:117-122)
  float Gh[2][2];  // face gradient of u_hat,v_hat
  int type;
};

// node_to_face_2nd_order (FVInterpolation.py:120-185) for channels 0:5 + _fix_face_flux_BC (FVscheme.py:32-48)
__device__ __forceinline__ void face_values(const fvgn_fv_desc& d, int f, FaceVals& o) {
  const int s = d.edge_s[f], r = d.edge_r[f];
  const float fx = d.face_pos[(size_t)f * 2], fy = d.face_pos[(size_t)f * 2 + 1];
  const float sx = fx - d.pos[(size_t)s * 2], sy = fy - d.pos[(size_t)s * 2 + 1];
  const float rx = fx - d.pos[(size_t)r * 2], ry = fy - d.pos[(size_t)r * 2 + 1];
  const float* ps = d.phi + (size_t)s * 7;
  const float* pr = d.phi + (size_t)r * 7;
  const float* gs = d.grad + (size_t)s * 14;
  const float* gr = d.grad + (size_t)r * 14;
  float v[5];
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    const float vs = ps[c] + (gs[c * 2] * sx + gs[c * 2 + 1] * sy);
    const float vr = pr[c] + (gr[c * 2] * rx + gr[c * 2 + 1] * ry);
    v[c] = (vs + vr) / 2.0f;
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      o.Gn[i][k] = (gs[i * 2 + k] + gr[i * 2 + k]) / 2.0f;
      o.Gh[i][k] = (gs[(3 + i) * 2 + k] + gr[(3 + i) * 2 + k]) / 2.0f;
    }
  const int t = d.face_type[f];
  o.type = t;
  o.p = v[2];
  if (t == NT_WALL) {
    o.uvn[0] = o.uvn[1] = o.uvh[0] = o.uvh[1] = 0.f;
  } else if (t == NT_INFLOW) {
    const float y0 = (d.y[(size_t)s * 2] + d.y[(size_t)r * 2]) / 2.0f;
    const float y1 = (d.y[(size_t)s * 2 + 1] + d.y[(size_t)r * 2 + 1]) / 2.0f;
    o.uvn[0] = y0; o.uvn[1] = y1; o.uvh[0] = y0; o.uvh[1] = y1;
  } else {
    o.uvn[0] = v[0]; o.uvn[1] = v[1]; o.uvh[0] = v[3]; o.uvh[1] = v[4];
  }
}

// ================================================================================ forward (one thread per cell)
__global__ void __launch_bounds__(128) fv_cell_fwd_kernel(const fvgn_fv_desc d) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.n_cells) return;
  const int b = d.batch_cell[c];
  const float* th = d.theta + (size_t)b * 9;
  const float th0 = th[0], th2 = th[2], th3 = th[3], th4 = th[4], th5 = th[5];
  const float area = d.cells_area[c];
  const float cx = d.centroid[(size_t)c * 2], cy = d.centroid[(size_t)c * 2 + 1];
  const int k0 = d.cell_ptr[c], k1 = d.cell_ptr[c + 1];
  float cont = 0.f, jx = 0.f, jy = 0.f, psq = 0.f;
  float pc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // u,v,p,u_old,v_old
  for (int k = k0; k < k1; ++k) {
    const int f = d.slot_face[k];
    const float a = d.face_area[f];
    const float Sx = d.slot_unv[(size_t)k * 2] * a, Sy = d.slot_unv[(size_t)k * 2 + 1] * a;
    FaceVals fv;
    face_values(d, f, fv);
    cont += fv.uvn[0] * Sx + fv.uvn[1] * Sy;
    // J_i = sum_j (th2 uh_i uh_j + th3 p delta_ij - th4 d_j uh_i) S_j     FVscheme.py:195-231
    const float uu = fv.uvh[0] * fv.uvh[0], uv = fv.uvh[0] * fv.uvh[1], vv = fv.uvh[1] * fv.uvh[1];
    jx += (uu * th2 + fv.p * th3 - fv.Gh[0][0] * th4) * Sx + (uv * th2 - fv.Gh[0][1] * th4) * Sy;
    jy += (uv * th2 - fv.Gh[1][0] * th4) * Sx + (vv * th2 + fv.p * th3 - fv.Gh[1][1] * th4) * Sy;
    if (fv.type == NT_OUTFLOW) {  // pressure outlet FVscheme.py:145-167
      const float r0 = th4 * (fv.Gn[0][0] * Sx + fv.Gn[0][1] * Sy) - fv.p * Sx;
      const float r1 = th4 * (fv.Gn[1][0] * Sx + fv.Gn[1][1] * Sy) - fv.p * Sy;
      psq += r0 * r0 + r1 * r1;
    }
    // node_to_cell_2nd_order (FVInterpolation.py:56-107)
    const int nd = d.slot_node[k];
    const float rx = cx - d.pos[(size_t)nd * 2], ry = cy - d.pos[(size_t)nd * 2 + 1];
    const float* pn = d.phi + (size_t)nd * 7;
    const float* gn = d.grad + (size_t)nd * 14;
    pc[0] += pn[0] + (gn[0] * rx + gn[1] * ry);
    pc[1] += pn[1] + (gn[2] * rx + gn[3] * ry);
    pc[2] += pn[2] + (gn[4] * rx + gn[5] * ry);
    pc[3] += pn[5] + (gn[10] * rx + gn[11] * ry);
    pc[4] += pn[6] + (gn[12] * rx + gn[13] * ry);
  }
  const float cnt = (float)max(k1 - k0, 1);
#pragma unroll
  for (int i = 0; i < 5; ++i) pc[i] /= cnt;
  const float dt = d.dt[b];
  const float ux = ((pc[0] - pc[3]) / dt) * area, uy = ((pc[1] - pc[4]) / dt) * area;
  const float src = th5 * area;
  float* res = d.res + (size_t)c * 4;
  res[0] = cont;
  res[1] = th0 * ux + (jx - src);
  res[2] = th0 * uy + (jy - src);
  res[3] = psq;
  float* o = d.phic + (size_t)c * 5;
#pragma unroll
  for (int i = 0; i < 5; ++i) o[i] = pc[i];
}

// non_conserved_form (FVscheme.py:276-511, hessian_phi = None): one thread per cell
__global__ void __launch_bounds__(128) fv_cell_fwd_nc_kernel(const fvgn_fv_desc d) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.n_cells) return;
  const int b = d.batch_cell[c];
  const float* th = d.theta + (size_t)b * 9;
  const float th0 = th[0], th2 = th[2], th3 = th[3], th4 = th[4], th5 = th[5];
  const float area = d.cells_area[c];
  const float cx = d.centroid[(size_t)c * 2], cy = d.centroid[(size_t)c * 2 + 1];
  const int k0 = d.cell_ptr[c], k1 = d.cell_ptr[c + 1];
  float vx = 0.f, vy = 0.f, psq = 0.f;
  float pc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // cell values of u,v,p,u_hat,v_hat,u_old,v_old
  float gc[5][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};  // cell-mean gradients of channels 0..4
  for (int k = k0; k < k1; ++k) {
    const int f = d.slot_face[k];
    const float a = d.face_area[f];
    const float Sx = d.slot_unv[(size_t)k * 2] * a, Sy = d.slot_unv[(size_t)k * 2 + 1] * a;
    FaceVals fv;
    face_values(d, f, fv);
    // divergence-form diffusion from the plainly averaged face gradients (:457-470)
    vx += fv.Gh[0][0] * Sx + fv.Gh[0][1] * Sy;
    vy += fv.Gh[1][0] * Sx + fv.Gh[1][1] * Sy;
    if (fv.type == NT_OUTFLOW) {  // pressure outlet (:376-396)
      const float r0 = th4 * (fv.Gn[0][0] * Sx + fv.Gn[0][1] * Sy) - fv.p * Sx;
      const float r1 = th4 * (fv.Gn[1][0] * Sx + fv.Gn[1][1] * Sy) - fv.p * Sy;
      psq += r0 * r0 + r1 * r1;
    }
    const int nd = d.slot_node[k];
    const float rx = cx - d.pos[(size_t)nd * 2], ry = cy - d.pos[(size_t)nd * 2 + 1];
    const float* pn = d.phi + (size_t)nd * 7;
    const float* gn = d.grad + (size_t)nd * 14;
#pragma unroll
    for (int ch = 0; ch < 7; ++ch) pc[ch] += pn[ch] + (gn[ch * 2] * rx + gn[ch * 2 + 1] * ry);
#pragma unroll
    for (int ch = 0; ch < 5; ++ch) { gc[ch][0] += gn[ch * 2]; gc[ch][1] += gn[ch * 2 + 1]; }
  }
  const float cnt = (float)max(k1 - k0, 1);
#pragma unroll
  for (int i = 0; i < 7; ++i) pc[i] /= cnt;
#pragma unroll
  for (int i = 0; i < 5; ++i) { gc[i][0] /= cnt; gc[i][1] /= cnt; }
  const float dt = d.dt[b];
  const float ux = ((pc[0] - pc[5]) / dt) * area, uy = ((pc[1] - pc[6]) / dt) * area;                 // :399
  const float cvx = (gc[3][0] * pc[3] + gc[3][1] * pc[4]) * area, cvy = (gc[4][0] * pc[3] + gc[4][1] * pc[4]) * area;  // :447
  const float gpx = gc[2][0] * area, gpy = gc[2][1] * area;                                           // :454
  const float src = th5 * area;
  float* res = d.res + (size_t)c * 4;
  res[0] = (gc[0][0] + gc[1][1]) * area;                                                              // :403-406
  res[1] = th0 * ux + th2 * cvx + th3 * gpx - th4 * vx - src;                                         // :472-478
  res[2] = th0 * uy + th2 * cvy + th3 * gpy - th4 * vy - src;
  res[3] = psq;
  float* o = d.phic + (size_t)c * 5;
  o[0] = pc[0]; o[1] = pc[1]; o[2] = pc[2]; o[3] = pc[5]; o[4] = pc[6];
  float* ax = d.cell_aux + (size_t)c * 6;
  ax[0] = pc[3]; ax[1] = pc[4]; ax[2] = gc[3][0]; ax[3] = gc[3][1]; ax[4] = gc[4][0]; ax[5] = gc[4][1];
}

// ================================================================================ backward, stage 1 (one thread per face)
// d_face[f] = [d uvn(2), d p, d uvh(2), d Gn(2x2), d Gh(2x2)] summed over the <=2 cells sharing the face
template <int FORM>
__global__ void __launch_bounds__(128) fv_bwd_face_kernel(const fvgn_fv_desc d) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= d.n_faces) return;
  FaceVals fv;
  face_values(d, (int)f, fv);
  const float a = d.face_area[f];
  float dun[2] = {0.f, 0.f}, duh[2] = {0.f, 0.f}, dp = 0.f;
  float dGn[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, dGh[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int t = d.face_slot_ptr[f]; t < d.face_slot_ptr[f + 1]; ++t) {
    const int k = d.face_slot[t];
    const int c = d.slot_cell[k];
    const int b = d.batch_cell[c];
    const float* th = d.theta + (size_t)b * 9;
    const float th2 = th[2], th3 = th[3], th4 = th[4];
    const float* cf = d.coef + (size_t)b * 4;
    const float* res = d.res + (size_t)c * 4;
    const float dcont = cf[0] * res[0], dmx = cf[1] * res[1], dmy = cf[2] * res[2], dps = cf[3];
    const float Sx = d.slot_unv[(size_t)k * 2] * a, Sy = d.slot_unv[(size_t)k * 2 + 1] * a;
    if (FORM == 0) {
      dun[0] += dcont * Sx;
      dun[1] += dcont * Sy;
      const float us = fv.uvh[0] * Sx + fv.uvh[1] * Sy;
      const float dmu = dmx * fv.uvh[0] + dmy * fv.uvh[1];
      duh[0] += th2 * (dmx * us + dmu * Sx);
      duh[1] += th2 * (dmy * us + dmu * Sy);
      dp += th3 * (dmx * Sx + dmy * Sy);
    }  // form 1: the face only carries the diffusion (below) and pressure-outlet terms
    dGh[0][0] -= th4 * dmx * Sx; dGh[0][1] -= th4 * dmx * Sy;
    dGh[1][0] -= th4 * dmy * Sx; dGh[1][1] -= th4 * dmy * Sy;
    if (fv.type == NT_OUTFLOW) {
      const float r0 = th4 * (fv.Gn[0][0] * Sx + fv.Gn[0][1] * Sy) - fv.p * Sx;
      const float r1 = th4 * (fv.Gn[1][0] * Sx + fv.Gn[1][1] * Sy) - fv.p * Sy;
      const float dr0 = 2.0f * r0 * dps, dr1 = 2.0f * r1 * dps;
      dGn[0][0] += th4 * dr0 * Sx; dGn[0][1] += th4 * dr0 * Sy;
      dGn[1][0] += th4 * dr1 * Sx; dGn[1][1] += th4 * dr1 * Sy;
      dp -= dr0 * Sx + dr1 * Sy;
    }
  }
  if (fv.type == NT_WALL || fv.type == NT_INFLOW) { dun[0] = dun[1] = duh[0] = duh[1] = 0.f; }
  float* o = d.d_face + (size_t)f * 13;
  o[0] = dun[0]; o[1] = dun[1]; o[2] = dp; o[3] = duh[0]; o[4] = duh[1];
  o[5] = dGn[0][0]; o[6] = dGn[0][1]; o[7] = dGn[1][0]; o[8] = dGn[1][1];
  o[9] = dGh[0][0]; o[10] = dGh[0][1]; o[11] = dGh[1][0]; o[12] = dGh[1][1];
}

// ================================================================================ backward, stage 2 (one thread per node)
template <int FORM>
__global__ void __launch_bounds__(128) fv_bwd_node_kernel(const fvgn_fv_desc d) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= d.n_nodes) return;
  const float px = d.pos[(size_t)n * 2], py = d.pos[(size_t)n * 2 + 1];
  float dphi[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  float dg[5][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
  for (int t = d.inc_ptr[n]; t < d.inc_ptr[n + 1]; ++t) {
    const int f = d.inc_code[t] >> 1;
    const float* df = d.d_face + (size_t)f * 13;
    const float rx = d.face_pos[(size_t)f * 2] - px, ry = d.face_pos[(size_t)f * 2 + 1] - py;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      const float h = 0.5f * df[c];
      dphi[c] += h;
      dg[c][0] += h * rx;
      dg[c][1] += h * ry;
    }
    dg[0][0] += 0.5f * df[5]; dg[0][1] += 0.5f * df[6]; dg[1][0] += 0.5f * df[7]; dg[1][1] += 0.5f * df[8];
    dg[3][0] += 0.5f * df[9]; dg[3][1] += 0.5f * df[10]; dg[4][0] += 0.5f * df[11]; dg[4][1] += 0.5f * df[12];
  }
  for (int t = d.node_slot_ptr[n]; t < d.node_slot_ptr[n + 1]; ++t) {
    const int k = d.node_slot[t];
    const int c = d.slot_cell[k];
    const int b = d.batch_cell[c];
    const float cnt = (float)max(d.cell_ptr[c + 1] - d.cell_ptr[c], 1);
    const float* cf = d.coef + (size_t)b * 4;
    const float* res = d.res + (size_t)c * 4;
    const float w = d.theta[(size_t)b * 9] * d.cells_area[c] / d.dt[b] / cnt;
    const float dux = cf[1] * res[1] * w, duy = cf[2] * res[2] * w;
    const float rx = d.centroid[(size_t)c * 2] - px, ry = d.centroid[(size_t)c * 2 + 1] - py;
    dphi[0] += dux; dg[0][0] += dux * rx; dg[0][1] += dux * ry;
    dphi[1] += duy; dg[1][0] += duy * rx; dg[1][1] += duy * ry;
    if (FORM == 1) {
      // non_conserved_form: the cell terms depend on the cell values of u_hat and on the cell-mean gradients
      const float* th = d.theta + (size_t)b * 9;
      const float aw = d.cells_area[c] / cnt;
      const float dmx = cf[1] * res[1], dmy = cf[2] * res[2], dcont = cf[0] * res[0];
      const float* ax = d.cell_aux + (size_t)c * 6;  // u_hat_c, v_hat_c, d(u_hat)/dx, /dy, d(v_hat)/dx, /dy
      const float duh = th[2] * aw * (dmx * ax[2] + dmy * ax[4]), dvh = th[2] * aw * (dmx * ax[3] + dmy * ax[5]);
      dphi[3] += duh; dg[3][0] += duh * rx; dg[3][1] += duh * ry;
      dphi[4] += dvh; dg[4][0] += dvh * rx; dg[4][1] += dvh * ry;
      dg[0][0] += dcont * aw;                 // continuity: du/dx + dv/dy
      dg[1][1] += dcont * aw;
      dg[2][0] += th[3] * aw * dmx;           // pressure gradient
      dg[2][1] += th[3] * aw * dmy;
      dg[3][0] += th[2] * aw * dmx * ax[0]; dg[3][1] += th[2] * aw * dmx * ax[1];   // convection: (grad u_hat) . u_hat
      dg[4][0] += th[2] * aw * dmy * ax[0]; dg[4][1] += th[2] * aw * dmy * ax[1];
    }
  }
  float* op = d.d_phi + (size_t)n * 7;
  float* og = d.d_grad + (size_t)n * 14;
#pragma unroll
  for (int c = 0; c < 5; ++c) { op[c] = dphi[c]; og[c * 2] = dg[c][0]; og[c * 2 + 1] = dg[c][1]; }
  op[5] = op[6] = 0.f;
  og[10] = og[11] = og[12] = og[13] = 0.f;
}

static int fv_check(const fvgn_fv_desc* d) {
  if (!d) return FVGN_ERR_NULL;
  if (d->n_nodes < 0 || d->n_faces < 0 || d->n_cells < 0 || d->n_slots < 0 || d->n_graphs < 1) return FVGN_ERR_SHAPE;
  if (!d->phi || !d->grad || !d->res) return FVGN_ERR_NULL;
  return FVGN_OK;
}

extern "C" int fvgn_fv_forward(const fvgn_fv_desc* d, void* stream) {
  int rc = fv_check(d);
  if (rc) return rc;
  if (!d->phic) return FVGN_ERR_NULL;
  if (d->form != 0 && d->form != 1) return FVGN_ERR_UNSUPPORTED;
  if (d->form == 1 && !d->cell_aux) return FVGN_ERR_NULL;
  if (d->n_cells == 0) return FVGN_OK;
  if (d->form == 0) {
    FVGN_LAUNCH_SEQ(fv_cell_fwd_kernel, (unsigned)((d->n_cells + 127) / 128), 128, 0, stream, *d);
  } else {
    FVGN_LAUNCH_SEQ(fv_cell_fwd_nc_kernel, (unsigned)((d->n_cells + 127) / 128), 128, 0, stream, *d);
  }
  FVGN_CHECK_LAUNCH();
  return FVGN_OK;
}

extern "C" int fvgn_fv_backward(const fvgn_fv_desc* d, void* stream) {
  int rc = fv_check(d);
  if (rc) return rc;
  if (!d->coef || !d->d_face || !d->d_phi || !d->d_grad) return FVGN_ERR_NULL;
  if (d->form != 0 && d->form != 1) return FVGN_ERR_UNSUPPORTED;
  if (d->form == 1 && !d->cell_aux) return FVGN_ERR_NULL;
  if (d->n_faces > 0) {
    if (d->form == 0) {
      FVGN_LAUNCH_SEQ(fv_bwd_face_kernel<0>, (unsigned)((d->n_faces + 127) / 128), 128, 0, stream, *d);
    } else {
      FVGN_LAUNCH_SEQ(fv_bwd_face_kernel<1>, (unsigned)((d->n_faces + 127) / 128), 128, 0, stream, *d);
    }
    FVGN_CHECK_LAUNCH();
  }
  if (d->n_nodes > 0) {
    if (d->form == 0) {
      FVGN_LAUNCH_SEQ(fv_bwd_node_kernel<0>, (unsigned)((d->n_nodes + 127) / 128), 128, 0, stream, *d);
    } else {
      FVGN_LAUNCH_SEQ(fv_bwd_node_kernel<1>, (unsigned)((d->n_nodes + 127) / 128), 128, 0, stream, *d);
    }
    FVGN_CHECK_LAUNCH();
  }
  return FVGN_OK;
}

// ================================================================================ outputs (no gradient)
__global__ void fv_out_cell_kernel(const fvgn_fv_desc d, const float* __restrict__ scale, float* __restrict__ uvp_cell) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d.n_cells) return;
  const int b = d.batch_cell[c];
#pragma unroll
  for (int j = 0; j < 3; ++j) uvp_cell[(size_t)c * 3 + j] = d.phic[(size_t)c * 5 + j] * scale[b * 3 + j];
}

// cell_to_node_2nd_order (FVInterpolation.py:241-263) + _enforce_boundary_condition + re-dimensionalise
__global__ void fv_out_node_kernel(const fvgn_fv_desc d, const int32_t* __restrict__ batch_node,
                                   const float* __restrict__ scale, int ncn_smooth, float* __restrict__ uvp_node) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= d.n_nodes) return;
  float v[3];
  if (ncn_smooth) {
    const float px = d.pos[(size_t)n * 2], py = d.pos[(size_t)n * 2 + 1];
    float num[3] = {0.f, 0.f, 0.f}, den = 0.f;
    for (int t = d.node_slot_ptr[n]; t < d.node_slot_ptr[n + 1]; ++t) {
      const int c = d.slot_cell[d.node_slot[t]];
      const float dx = px - d.centroid[(size_t)c * 2], dy = py - d.centroid[(size_t)c * 2 + 1];
      const float w = 1.0f / sqrtf(dx * dx + dy * dy);
#pragma unroll
      for (int j = 0; j < 3; ++j) num[j] += d.phic[(size_t)c * 5 + j] * w;
      den += w;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) v[j] = num[j] / den;
  } else {
#pragma unroll
    for (int j = 0; j < 3; ++j) v[j] = d.phi[(size_t)n * 7 + j];
  }
  const int t = d.node_type[n];
  if (t == NT_WALL || t == NT_INFLOW || t == NT_PRESS_POINT || t == NT_IN_WALL) {
    v[0] = d.y[(size_t)n * 2];
    v[1] = d.y[(size_t)n * 2 + 1];
  }
  if (t == NT_PRESS_POINT) v[2] = 0.f;
  const int b = batch_node[n];
#pragma unroll
  for (int j = 0; j < 3; ++j) uvp_node[(size_t)n * 3 + j] = v[j] * scale[b * 3 + j];
}

extern "C" int fvgn_fv_outputs(const fvgn_fv_desc* d, const int32_t* batch_node, const float* scale, int32_t ncn_smooth,
                               float* uvp_node, float* uvp_cell, void* stream) {
  if (!d || !d->phic || !d->phi) return FVGN_ERR_NULL;
  if (d->n_cells > 0) {
    FVGN_LAUNCH_SEQ(fv_out_cell_kernel, (unsigned)((d->n_cells + 255) / 256), 256, 0, stream, *d, scale, uvp_cell);
    FVGN_CHECK_LAUNCH();
  }
  if (d->n_nodes > 0) {
    FVGN_LAUNCH_SEQ(fv_out_node_kernel, (unsigned)((d->n_nodes + 255) / 256), 256, 0, stream, *d, batch_node, scale,
                    ncn_smooth, uvp_node);
    FVGN_CHECK_LAUNCH();
  }
  return FVGN_OK;
}
