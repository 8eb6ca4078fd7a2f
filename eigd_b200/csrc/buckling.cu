// Linearised-buckling element kernels for the plane-stress Q4 column problem of the reference
// (examples/buckling.py): element stresses of the fundamental path, gather-form assembly of the stress
// (geometric) stiffness matrix G(u, x), and the fused sensitivities of  sum_k w_k^T G(u, x) v_k  with
// respect to the element densities and to the displacement u, plus the Dirichlet reduce / expand maps.
//
// Replaces (paths relative to the reference root):
//   get_stress_stiffness_matrix                       examples/buckling.py:220-255
//   intital_stress_stiffness_matrix_deriv             examples/buckling.py:283-310   (dfds)
//   get_stress_stiffness_matrix_uderiv_tensor         examples/buckling.py:312-322
//   get_stress_stiffness_matrix_xderiv_tensor         examples/buckling.py:324-343
//   reduce_vector / full_vector                       examples/buckling.py:499-518
//   Be, Te, detJ                                      examples/fe_utils.py:58-98
#include "common.cuh"
#include "../../include/eigd_b200.h"

namespace {

struct Geo {
  double Nx[4][4], Ny[4][4], detJ[4];  // [gauss point][node]
};

__device__ __forceinline__ void geometry(const int* __restrict__ conn, const double* __restrict__ xy, int e, int nd[4], Geo& g) {
  const double gp = 0.57735026918962576451;
  double xe[4], ye[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    nd[a] = conn[e * 4 + a];
    xe[a] = xy[2 * nd[a]];
    ye[a] = xy[2 * nd[a] + 1];
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    // index = 2*j + i with xi = pts[i], eta = pts[j] (examples/buckling.py:266-273)
    double xi = (q & 1) ? gp : -gp;
    double eta = (q & 2) ? gp : -gp;
    double Nxi[4] = {-0.25 * (1.0 - eta), 0.25 * (1.0 - eta), 0.25 * (1.0 + eta), -0.25 * (1.0 + eta)};
    double Neta[4] = {-0.25 * (1.0 - xi), -0.25 * (1.0 + xi), 0.25 * (1.0 + xi), 0.25 * (1.0 - xi)};
    double J00 = 0, J10 = 0, J01 = 0, J11 = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      J00 = fma(xe[a], Nxi[a], J00);
      J10 = fma(ye[a], Nxi[a], J10);
      J01 = fma(xe[a], Neta[a], J01);
      J11 = fma(ye[a], Neta[a], J11);
    }
    double det = J00 * J11 - J01 * J10;
    double i00 = J11 / det, i01 = -J01 / det, i10 = -J10 / det, i11 = J00 / det;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      g.Nx[q][a] = i00 * Nxi[a] + i10 * Neta[a];
      g.Ny[q][a] = i01 * Nxi[a] + i11 * Neta[a];
    }
    g.detJ[q] = det;
  }
}

// sdet[e][q][i] = detJ_q * ks[e] * (C0 Be(q) u_e)_i : the weights of Te in the element stress stiffness
__global__ void __launch_bounds__(128)
q4_stress_kernel(int nelems, const int* __restrict__ conn, const double* __restrict__ xy, const double* __restrict__ cmat,
                 const double* __restrict__ ks, const double* __restrict__ u, double* __restrict__ sdet) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nelems) return;
  int nd[4];
  Geo g;
  geometry(conn, xy, e, nd, g);
  double ue[8];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    ue[2 * a] = u[2 * nd[a]];
    ue[2 * a + 1] = u[2 * nd[a] + 1];
  }
  const double c00 = cmat[0], c01 = cmat[1], c02 = cmat[2], c11 = cmat[3], c12 = cmat[4], c22 = cmat[5];
  const double s = ks[e];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    double ex = 0, ey = 0, gxy = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      ex = fma(g.Nx[q][a], ue[2 * a], ex);
      ey = fma(g.Ny[q][a], ue[2 * a + 1], ey);
      gxy = fma(g.Ny[q][a], ue[2 * a], fma(g.Nx[q][a], ue[2 * a + 1], gxy));
    }
    const double w = g.detJ[q] * s;
    sdet[(e * 4 + q) * 3 + 0] = w * (c00 * ex + c01 * ey + c02 * gxy);
    sdet[(e * 4 + q) * 3 + 1] = w * (c01 * ex + c11 * ey + c12 * gxy);
    sdet[(e * 4 + q) * 3 + 2] = w * (c02 * ex + c12 * ey + c22 * gxy);
  }
}

// gather-form assembly of G: non-zero p = sum over its (element, a, b) sources of
// sum_q s0 Nxa Nxb + s1 Nya Nyb + s2 (Nxa Nyb + Nya Nxb), non-zero only for equal displacement components
__global__ void q4_assemble_geometric_kernel(int64_t nnz, const int64_t* __restrict__ src_ptr, const int64_t* __restrict__ src,
                                             const int* __restrict__ conn, const double* __restrict__ xy,
                                             const double* __restrict__ sdet, double* __restrict__ Gv) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  double gv = 0.0;
  for (int64_t qq = src_ptr[p]; qq < src_ptr[p + 1]; ++qq) {
    int64_t s = src[qq];
    int e = (int)(s / 64);
    int ab = (int)(s - (int64_t)e * 64);
    int a = ab >> 3, b = ab & 7;
    if ((a & 1) != (b & 1)) continue;
    int na = a >> 1, nb = b >> 1;
    int nd[4];
    Geo g;
    geometry(conn, xy, e, nd, g);
    double ge = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double* sd = sdet + ((int64_t)e * 4 + q) * 3;
      // select the two nodes without dynamic register indexing
      double nxa = 0, nya = 0, nxb = 0, nyb = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i == na) { nxa = g.Nx[q][i]; nya = g.Ny[q][i]; }
        if (i == nb) { nxb = g.Nx[q][i]; nyb = g.Ny[q][i]; }
      }
      ge += sd[0] * nxa * nxb + sd[1] * nya * nyb + sd[2] * (nxa * nyb + nya * nxb);
    }
    gv += ge;
  }
  Gv[p] = gv;
}

// Per element, with G lanes splitting the modes:
//   dfds[q][i] = detJ_q sum_k sum_d  Te_i(q) : (w_{d,k} v_{d,k}^T)                  (examples/buckling.py:283-310)
//   out_rho[e] += sx * dks[e] * sum_q sum_j (C0 dfds[q])_j (Be(q) u_e)_j             (:324-343)
//   due[e][j8]  = ks[e] * sum_q sum_j Be[j][j8](q) (C0 dfds[q])_j                    (:312-322)
template <int G>
__global__ void __launch_bounds__(256)
q4_gderiv_kernel(int nelems, const int* __restrict__ conn, const double* __restrict__ xy, const double* __restrict__ cmat,
                 const double* __restrict__ W, const double* __restrict__ V, int N, int ldw, const double* __restrict__ ks,
                 const double* __restrict__ dks, const double* __restrict__ u, double sx, double* __restrict__ out_rho,
                 double* __restrict__ due) {
  const int lane = threadIdx.x & (G - 1);
  int e = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G);
  const bool valid = e < nelems;
  if (!valid) e = nelems - 1;
  int nd[4];
  Geo g;
  geometry(conn, xy, e, nd, g);
  double dfds[4][3];
#pragma unroll
  for (int q = 0; q < 4; ++q) dfds[q][0] = dfds[q][1] = dfds[q][2] = 0.0;
  for (int k = lane; k < N; k += G) {
    double w[8], v[8];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      int64_t o = (int64_t)(2 * nd[a]) * ldw + k;
      w[2 * a] = W[o];
      w[2 * a + 1] = W[o + ldw];
      v[2 * a] = V[o];
      v[2 * a + 1] = V[o + ldw];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        double gxw = 0, gyw = 0, gxv = 0, gyv = 0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          gxw = fma(g.Nx[q][a], w[2 * a + d], gxw);
          gyw = fma(g.Ny[q][a], w[2 * a + d], gyw);
          gxv = fma(g.Nx[q][a], v[2 * a + d], gxv);
          gyv = fma(g.Ny[q][a], v[2 * a + d], gyv);
        }
        dfds[q][0] = fma(gxw, gxv, dfds[q][0]);
        dfds[q][1] = fma(gyw, gyv, dfds[q][1]);
        dfds[q][2] += gxw * gyv + gyw * gxv;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
      for (int o = G >> 1; o > 0; o >>= 1) dfds[q][i] += __shfl_xor_sync(0xffffffffu, dfds[q][i], o, G);
    }
  if (!valid || lane != 0) return;
  const double c00 = cmat[0], c01 = cmat[1], c02 = cmat[2], c11 = cmat[3], c12 = cmat[4], c22 = cmat[5];
  double ue[8];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    ue[2 * a] = u[2 * nd[a]];
    ue[2 * a + 1] = u[2 * nd[a] + 1];
  }
  double drho = 0.0;
  double de[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double f0 = g.detJ[q] * dfds[q][0], f1 = g.detJ[q] * dfds[q][1], f2 = g.detJ[q] * dfds[q][2];
    const double t0 = c00 * f0 + c01 * f1 + c02 * f2;
    const double t1 = c01 * f0 + c11 * f1 + c12 * f2;
    const double t2 = c02 * f0 + c12 * f1 + c22 * f2;
    double ex = 0, ey = 0, gxy = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      ex = fma(g.Nx[q][a], ue[2 * a], ex);
      ey = fma(g.Ny[q][a], ue[2 * a + 1], ey);
      gxy = fma(g.Ny[q][a], ue[2 * a], fma(g.Nx[q][a], ue[2 * a + 1], gxy));
      // Be columns: dof u -> [Nx, 0, Ny], dof v -> [0, Ny, Nx]
      de[2 * a] += g.Nx[q][a] * t0 + g.Ny[q][a] * t2;
      de[2 * a + 1] += g.Ny[q][a] * t1 + g.Nx[q][a] * t2;
    }
    drho += t0 * ex + t1 * ey + t2 * gxy;
  }
  if (out_rho) out_rho[e] += sx * dks[e] * drho;
  if (due) {
    const double s = ks[e];
#pragma unroll
    for (int j = 0; j < 8; ++j) due[(int64_t)e * 8 + j] = s * de[j];
  }
}

// dfdu[2 v + d] = sum over the elements e around node v of due[e][2 local(v, e) + d]  (np.add.at, :318-320)
__global__ void q4_dof_gather_kernel(int nnodes, const int* __restrict__ nptr, const int* __restrict__ nelem,
                                     const int* __restrict__ nlocal, const double* __restrict__ due, double* __restrict__ out) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nnodes) return;
  double s0 = 0.0, s1 = 0.0;
  for (int p = nptr[v]; p < nptr[v + 1]; ++p) {
    const double* d = due + (int64_t)nelem[p] * 8 + 2 * nlocal[p];
    s0 += d[0];
    s1 += d[1];
  }
  out[2 * v] = s0;
  out[2 * v + 1] = s1;
}

// full[idx[i], :] = red[i, :]   (full_vector, :506-511; the rest of `full` must be zero)
__global__ void expand_rows_kernel(int64_t nr, int k, const int* __restrict__ idx, const double* __restrict__ red,
                                   double* __restrict__ full) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nr * k) return;
  int64_t i = t / k;
  int c = (int)(t - i * k);
  full[(int64_t)idx[i] * k + c] = red[t];
}

// red[i, :] = full[idx[i], :]   (reduce_vector, :499-503)
__global__ void reduce_rows_kernel(int64_t nr, int k, const int* __restrict__ idx, const double* __restrict__ full,
                                   double* __restrict__ red) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nr * k) return;
  int64_t i = t / k;
  int c = (int)(t - i * k);
  red[t] = full[(int64_t)idx[i] * k + c];
}

}  // namespace

extern "C" int eigd_q4_stress(int nelems, const int* d_conn, const double* d_xy, const double* d_cmat6, const double* d_ks,
                              const double* d_u, double* d_sdet) {
  if (nelems <= 0) return 0;
  EIGD_LAUNCH(q4_stress_kernel, (nelems + 127) / 128, 128, 0, nelems, d_conn, d_xy, d_cmat6, d_ks, d_u, d_sdet);
  EIGD_CHECK_LAUNCH();
  return 0;
}

extern "C" int eigd_q4_assemble_geometric(int nelems, const int* d_conn, const double* d_xy, const double* d_sdet,
                                          const int64_t* d_src_ptr, const int64_t* d_src, int64_t nnz, double* d_Gvals) {
  (void)nelems;
  if (nnz <= 0) return 0;
  EIGD_LAUNCH(q4_assemble_geometric_kernel, (int)((nnz + 127) / 128), 128, 0, nnz, d_src_ptr, d_src, d_conn, d_xy, d_sdet, d_Gvals);
  EIGD_CHECK_LAUNCH();
  return 0;
}

extern "C" int eigd_q4_gderiv(int nelems, const int* d_conn, const double* d_xy, const double* d_cmat6, const double* d_W,
                              const double* d_V, int N, int ldw, const double* d_ks, const double* d_dks, const double* d_u,
                              double sx, double* d_out_rho, double* d_due) {
  if (nelems <= 0 || N <= 0) return 0;
  constexpr int G = 8;
  int grid = (int)(((int64_t)nelems * G + 255) / 256);
  EIGD_LAUNCH((q4_gderiv_kernel<G>), grid, 256, 0, nelems, d_conn, d_xy, d_cmat6, d_W, d_V, N, ldw, d_ks, d_dks, d_u, sx,
              d_out_rho, d_due);
  EIGD_CHECK_LAUNCH();
  return 0;
}

extern "C" int eigd_q4_dof_gather(int nnodes, const int* d_nptr, const int* d_nelem, const int* d_nlocal, const double* d_due,
                                  double* d_out) {
  if (nnodes <= 0) return 0;
  EIGD_LAUNCH(q4_dof_gather_kernel, (nnodes + 255) / 256, 256, 0, nnodes, d_nptr, d_nelem, d_nlocal, d_due, d_out);
  EIGD_CHECK_LAUNCH();
  return 0;
}

extern "C" int eigd_expand_rows(int64_t nr, int k, const int* d_idx, const double* d_red, double* d_full) {
  if (nr <= 0 || k <= 0) return 0;
  EIGD_LAUNCH(expand_rows_kernel, (int)((nr * k + 255) / 256), 256, 0, nr, k, d_idx, d_red, d_full);
  EIGD_CHECK_LAUNCH();
  return 0;
}

extern "C" int eigd_reduce_rows(int64_t nr, int k, const int* d_idx, const double* d_full, double* d_red) {
  if (nr <= 0 || k <= 0) return 0;
  EIGD_LAUNCH(reduce_rows_kernel, (int)((nr * k + 255) / 256), 256, 0, nr, k, d_idx, d_full, d_red);
  EIGD_CHECK_LAUNCH();
  return 0;
}
