// Q4 element kernels: per-element bilinear forms w_e^T (dK_e/ds) v_e, w_e^T (dM_e/dm) v_e summed
// over modes (warp-shuffle segmented reduction, one lane group per element), gather-form
// assembly of K and M through the (element, a, b) -> CSR map, and the element -> node gather.
//
// Replaces the numpy einsum callbacks of the reference examples:
//   thermal      examples/thermal.py:126-148 (K), :150-190 (dK), :192-214 (M), :216-246 (dM)
//   plane stress examples/natural_frequency.py:134-160 (K), :162-203 (dK), :205-236 (M), :238-284 (dM)
//   node scatter examples/thermal.py:612-615, examples/natural_frequency.py:509-512
// Shape functions / Jacobian follow examples/fe_utils.py:4-16, 26-44 (2x2 Gauss, +-1/sqrt(3)).
#include "common.cuh"
#include "../../include/eigd_b200.h"

namespace {

struct Q4Geom {
  double Nx[4][4], Ny[4][4], N[4][4], detJ[4];  // [gauss point][node]
};

__device__ __forceinline__ void q4_geometry(const double xe[4], const double ye[4], Q4Geom& g) {
  const double gp = 0.57735026918962576451;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    // index = 2*j + i with xi = pts[i], eta = pts[j] (thermal.py:117); sums do not depend on it
    double xi = (q & 1) ? gp : -gp;
    double eta = (q & 2) ? gp : -gp;
    double Nxi[4] = {-0.25 * (1.0 - eta), 0.25 * (1.0 - eta), 0.25 * (1.0 + eta), -0.25 * (1.0 + eta)};
    double Neta[4] = {-0.25 * (1.0 - xi), -0.25 * (1.0 + xi), 0.25 * (1.0 + xi), 0.25 * (1.0 - xi)};
    g.N[q][0] = 0.25 * (1.0 - xi) * (1.0 - eta);
    g.N[q][1] = 0.25 * (1.0 + xi) * (1.0 - eta);
    g.N[q][2] = 0.25 * (1.0 + xi) * (1.0 + eta);
    g.N[q][3] = 0.25 * (1.0 - xi) * (1.0 + eta);
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

// G lanes per element; lane l handles modes l, l+G, ...
template <int KIND, int G>
__global__ void __launch_bounds__(256)
q4_quadforms_kernel(int nelems, const int* __restrict__ conn, const double* __restrict__ xy,
                    const double* __restrict__ cmat, const double* __restrict__ WA, const double* __restrict__ WB,
                    const double* __restrict__ V, int N, int ldw, const double* __restrict__ dk,
                    const double* __restrict__ dm, double sA, double sB, double* __restrict__ out) {
  const int lane = threadIdx.x & (G - 1);
  int e = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G);
  const bool valid = e < nelems;
  if (!valid) e = nelems - 1;  // keep the whole warp converged for the shuffles
  int nd[4];
  double xe[4], ye[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    nd[a] = conn[e * 4 + a];
    xe[a] = xy[2 * nd[a]];
    ye[a] = xy[2 * nd[a] + 1];
  }
  Q4Geom g;
  q4_geometry(xe, ye, g);
  double sumK = 0.0, sumM = 0.0;
  for (int k = lane; k < N; k += G) {
    if (KIND == 0) {
      double wa[4], wb[4], v[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        int64_t o = (int64_t)nd[a] * ldw + k;
        v[a] = V[o];
        wa[a] = WA ? WA[o] : 0.0;
        wb[a] = WB ? WB[o] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        double gxw = 0, gyw = 0, gxv = 0, gyv = 0, nw = 0, nv = 0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          gxw = fma(g.Nx[q][a], wa[a], gxw);
          gyw = fma(g.Ny[q][a], wa[a], gyw);
          gxv = fma(g.Nx[q][a], v[a], gxv);
          gyv = fma(g.Ny[q][a], v[a], gyv);
          nw = fma(g.N[q][a], wb[a], nw);
          nv = fma(g.N[q][a], v[a], nv);
        }
        sumK = fma(g.detJ[q], gxw * gxv + gyw * gyv, sumK);
        sumM = fma(g.detJ[q], nw * nv, sumM);
      }
    } else {
      double wa[8], wb[8], v[8];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        int64_t o = (int64_t)(2 * nd[a]) * ldw + k;
        v[2 * a] = V[o];
        v[2 * a + 1] = V[o + ldw];
        wa[2 * a] = WA ? WA[o] : 0.0;
        wa[2 * a + 1] = WA ? WA[o + ldw] : 0.0;
        wb[2 * a] = WB ? WB[o] : 0.0;
        wb[2 * a + 1] = WB ? WB[o + ldw] : 0.0;
      }
      const double c00 = cmat[0], c01 = cmat[1], c02 = cmat[2], c11 = cmat[3], c12 = cmat[4], c22 = cmat[5];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        double ew[3] = {0, 0, 0}, ev[3] = {0, 0, 0}, hw[2] = {0, 0}, hv[2] = {0, 0};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          ew[0] = fma(g.Nx[q][a], wa[2 * a], ew[0]);
          ew[1] = fma(g.Ny[q][a], wa[2 * a + 1], ew[1]);
          ew[2] = fma(g.Ny[q][a], wa[2 * a], fma(g.Nx[q][a], wa[2 * a + 1], ew[2]));
          ev[0] = fma(g.Nx[q][a], v[2 * a], ev[0]);
          ev[1] = fma(g.Ny[q][a], v[2 * a + 1], ev[1]);
          ev[2] = fma(g.Ny[q][a], v[2 * a], fma(g.Nx[q][a], v[2 * a + 1], ev[2]));
          hw[0] = fma(g.N[q][a], wb[2 * a], hw[0]);
          hw[1] = fma(g.N[q][a], wb[2 * a + 1], hw[1]);
          hv[0] = fma(g.N[q][a], v[2 * a], hv[0]);
          hv[1] = fma(g.N[q][a], v[2 * a + 1], hv[1]);
        }
        double s0 = c00 * ev[0] + c01 * ev[1] + c02 * ev[2];
        double s1 = c01 * ev[0] + c11 * ev[1] + c12 * ev[2];
        double s2 = c02 * ev[0] + c12 * ev[1] + c22 * ev[2];
        sumK = fma(g.detJ[q], ew[0] * s0 + ew[1] * s1 + ew[2] * s2, sumK);
        sumM = fma(g.detJ[q], hw[0] * hv[0] + hw[1] * hv[1], sumM);
      }
    }
  }
  // segmented reduction over the G lanes of this element
#pragma unroll
  for (int o = G >> 1; o > 0; o >>= 1) {
    sumK += __shfl_xor_sync(0xffffffffu, sumK, o, G);
    sumM += __shfl_xor_sync(0xffffffffu, sumM, o, G);
  }
  if (valid && lane == 0) {
    double r = 0.0;
    if (WA) r += sA * (dk ? dk[e] : 1.0) * sumK;
    if (WB) r -= sB * (dm ? dm[e] : 1.0) * sumM;
    out[e] += r;
  }
}

// gather-form assembly: value of CSR non-zero p = sum over its (element, a, b) sources
template <int KIND>
__global__ void q4_assemble_kernel(int64_t nnz, const int64_t* __restrict__ src_ptr, const int64_t* __restrict__ src,
                                   const int* __restrict__ conn, const double* __restrict__ xy,
                                   const double* __restrict__ ks, const double* __restrict__ ms,
                                   const double* __restrict__ cmat, double* __restrict__ Kv, double* __restrict__ Mv) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  constexpr int NE = (KIND == 0) ? 4 : 8;
  const double gp = 0.57735026918962576451;
  double kv = 0.0, mv = 0.0;
  for (int64_t q = src_ptr[p]; q < src_ptr[p + 1]; ++q) {
    int64_t s = src[q];
    int e = (int)(s / (NE * NE));
    int ab = (int)(s - (int64_t)e * NE * NE);
    int a = ab / NE, b = ab - a * NE;
    double xe[4], ye[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int nd = conn[e * 4 + i];
      xe[i] = xy[2 * nd];
      ye[i] = xy[2 * nd + 1];
    }
    double ke = 0.0, me = 0.0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      double xi = (g & 1) ? gp : -gp, eta = (g & 2) ? gp : -gp;
      double Nxi[4] = {-0.25 * (1.0 - eta), 0.25 * (1.0 - eta), 0.25 * (1.0 + eta), -0.25 * (1.0 + eta)};
      double Neta[4] = {-0.25 * (1.0 - xi), -0.25 * (1.0 + xi), 0.25 * (1.0 + xi), 0.25 * (1.0 - xi)};
      double Nn[4] = {0.25 * (1.0 - xi) * (1.0 - eta), 0.25 * (1.0 + xi) * (1.0 - eta),
                      0.25 * (1.0 + xi) * (1.0 + eta), 0.25 * (1.0 - xi) * (1.0 + eta)};
      double J00 = 0, J10 = 0, J01 = 0, J11 = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        J00 = fma(xe[i], Nxi[i], J00);
        J10 = fma(ye[i], Nxi[i], J10);
        J01 = fma(xe[i], Neta[i], J01);
        J11 = fma(ye[i], Neta[i], J11);
      }
      double det = J00 * J11 - J01 * J10;
      double i00 = J11 / det, i01 = -J01 / det, i10 = -J10 / det, i11 = J00 / det;
      if (KIND == 0) {
        double nxa = i00 * Nxi[a] + i10 * Neta[a], nya = i01 * Nxi[a] + i11 * Neta[a];
        double nxb = i00 * Nxi[b] + i10 * Neta[b], nyb = i01 * Nxi[b] + i11 * Neta[b];
        ke = fma(det, nxa * nxb + nya * nyb, ke);
        me = fma(det, Nn[a] * Nn[b], me);
      } else {
        int na = a >> 1, ca = a & 1, nb = b >> 1, cb = b & 1;
        double nxa = i00 * Nxi[na] + i10 * Neta[na], nya = i01 * Nxi[na] + i11 * Neta[na];
        double nxb = i00 * Nxi[nb] + i10 * Neta[nb], nyb = i01 * Nxi[nb] + i11 * Neta[nb];
        // columns of Be: dof u -> [Nx, 0, Ny], dof v -> [0, Ny, Nx]
        double ba[3] = {ca ? 0.0 : nxa, ca ? nya : 0.0, ca ? nxa : nya};
        double bb[3] = {cb ? 0.0 : nxb, cb ? nyb : 0.0, cb ? nxb : nyb};
        double t0 = cmat[0] * bb[0] + cmat[1] * bb[1] + cmat[2] * bb[2];
        double t1 = cmat[1] * bb[0] + cmat[3] * bb[1] + cmat[4] * bb[2];
        double t2 = cmat[2] * bb[0] + cmat[4] * bb[1] + cmat[5] * bb[2];
        ke = fma(det, ba[0] * t0 + ba[1] * t1 + ba[2] * t2, ke);
        if (ca == cb) me = fma(det, Nn[na] * Nn[nb], me);
      }
    }
    kv = fma(ks[e], ke, kv);
    mv = fma(ms[e], me, mv);
  }
  if (Kv) Kv[p] = kv;
  if (Mv) Mv[p] = mv;
}

__global__ void node_gather_kernel(int nnodes, const int* __restrict__ nptr, const int* __restrict__ nelem,
                                   const double* __restrict__ ev, double scale, double* __restrict__ out) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nnodes) return;
  double s = 0.0;
  for (int p = nptr[v]; p < nptr[v + 1]; ++p) s += ev[nelem[p]];
  out[v] = scale * s;
}

}  // namespace

extern "C" int eigd_q4_assemble(int kind, int nelems, const int* d_conn, const double* d_xy, const double* d_ks,
                                const double* d_ms, const double* d_cmat6, const int64_t* d_src_ptr,
                                const int64_t* d_src, int64_t nnz, double* d_Kvals, double* d_Mvals) {
  (void)nelems;
  if (nnz <= 0) return 0;
  int grid = (int)((nnz + 127) / 128);
  if (kind == 0) EIGD_LAUNCH(q4_assemble_kernel<0>, grid, 128, 0, nnz, d_src_ptr, d_src, d_conn, d_xy, d_ks, d_ms, d_cmat6, d_Kvals, d_Mvals);
  else if (kind == 1) EIGD_LAUNCH(q4_assemble_kernel<1>, grid, 128, 0, nnz, d_src_ptr, d_src, d_conn, d_xy, d_ks, d_ms, d_cmat6, d_Kvals, d_Mvals);
  else { eigd_set_error("q4_assemble: unknown kind %d", kind); return 1; }
  EIGD_CHECK_LAUNCH();
  return 0;
}

extern "C" int eigd_q4_quadforms(int kind, int nelems, const int* d_conn, const double* d_xy, const double* d_cmat6,
                                 const double* d_WA, const double* d_WB, const double* d_V, int N, int ldw,
                                 const double* d_dk, const double* d_dm, double sA, double sB, double* d_out) {
  if (nelems <= 0 || N <= 0) return 0;
  constexpr int G = 8;
  int grid = (int)(((int64_t)nelems * G + 255) / 256);
  if (kind == 0) EIGD_LAUNCH((q4_quadforms_kernel<0, G>), grid, 256, 0, nelems, d_conn, d_xy, d_cmat6, d_WA, d_WB, d_V, N, ldw, d_dk, d_dm, sA, sB, d_out);
  else if (kind == 1) EIGD_LAUNCH((q4_quadforms_kernel<1, G>), grid, 256, 0, nelems, d_conn, d_xy, d_cmat6, d_WA, d_WB, d_V, N, ldw, d_dk, d_dm, sA, sB, d_out);
  else { eigd_set_error("q4_quadforms: unknown kind %d", kind); return 1; }
  EIGD_CHECK_LAUNCH();
  return 0;
}

extern "C" int eigd_node_gather(int nnodes, const int* d_nptr, const int* d_nelem, const double* d_evals, double scale,
                                double* d_out) {
  if (nnodes <= 0) return 0;
  EIGD_LAUNCH(node_gather_kernel, (nnodes + 255) / 256, 256, 0, nnodes, d_nptr, d_nelem, d_evals, scale, d_out);
  EIGD_CHECK_LAUNCH();
  return 0;
}

// ---- smooth Heaviside projection of the filtered density (examples/node_filter.py:174-181, 193-203) ------------
// forward : out = (tanh(beta eta) + tanh(beta (rho - eta))) / (tanh(beta eta) + tanh(beta (1 - eta)))
// gradient: out = g * (beta / denom) / cosh(beta (rho - eta))^2          (rho = the filtered, unprojected field)
namespace {
__global__ void filter_project_kernel(int n, double beta, double eta, const double* __restrict__ rho,
                                      const double* __restrict__ g, double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double denom = tanh(beta * eta) + tanh(beta * (1.0 - eta));
  const double t = beta * (rho[i] - eta);
  if (g) {
    const double c = cosh(t);
    out[i] = g[i] * (beta / denom) / (c * c);
  } else {
    out[i] = (tanh(beta * eta) + tanh(t)) / denom;
  }
}
}  // namespace

extern "C" int eigd_filter_project(int n, double beta, double eta, const double* d_rho, const double* d_g, double* d_out) {
  if (n <= 0) return 0;
  EIGD_LAUNCH(filter_project_kernel, (n + 255) / 256, 256, 0, n, beta, eta, d_rho, d_g, d_out);
  EIGD_CHECK_LAUNCH();
  return 0;
}

// ---- element density and material interpolation ----------------------------------------------
// rhoE[e] = 1/4 sum_a rho[conn[e, a]]          (examples/thermal.py:348-353, natural_frequency.py:399-404)
// law 0 (thermal.py:132,198,175-188,236-244): ks = k0((1-b) r^p + b), ms = c0((1-b) r + b)
// law 1 (natural_frequency.py:140-143,219-220 "simp"): ks = r^p + r0, ms = c0 r
// law 2 (RAMP, natural_frequency.py:142-143): ks = r/(1+q(1-r)) + r0, ms = c0 r
// par = {p or q, k0, c0, b or r0}
namespace {
__global__ void q4_material_kernel(int law, int nelems, const int* __restrict__ conn, const double* __restrict__ rho,
                                   double p0, double k0, double c0, double b, double* __restrict__ rhoE,
                                   double* __restrict__ ks, double* __restrict__ ms, double* __restrict__ dk,
                                   double* __restrict__ dm) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nelems) return;
  double r;
  if (conn) {
    r = 0.25 * (rho[conn[4 * e]] + rho[conn[4 * e + 1]] + rho[conn[4 * e + 2]] + rho[conn[4 * e + 3]]);
    if (rhoE) rhoE[e] = r;
  } else {
    r = rho[e];
  }
  double vk, vm, gk, gm;
  if (law == 0) {
    vk = k0 * ((1.0 - b) * pow(r, p0) + b);
    gk = (1.0 - b) * k0 * p0 * pow(r, p0 - 1.0);
    vm = c0 * ((1.0 - b) * r + b);
    gm = (1.0 - b) * c0;
  } else if (law == 1) {
    vk = k0 * pow(r, p0) + b;
    gk = k0 * p0 * pow(r, p0 - 1.0);
    vm = c0 * r;
    gm = c0;
  } else {
    double den = 1.0 + p0 * (1.0 - r);
    vk = k0 * r / den + b;
    gk = k0 * (1.0 + p0) / (den * den);
    vm = c0 * r;
    gm = c0;
  }
  if (ks) ks[e] = vk;
  if (ms) ms[e] = vm;
  if (dk) dk[e] = gk;
  if (dm) dm[e] = gm;
}
}  // namespace

extern "C" int eigd_q4_material(int law, int nelems, const int* d_conn, const double* d_rho, const double* par4,
                                double* d_rhoE, double* d_ks, double* d_ms, double* d_dk, double* d_dm) {
  if (nelems <= 0) return 0;
  if (law < 0 || law > 2) { eigd_set_error("q4_material: unknown law %d", law); return 1; }
  EIGD_LAUNCH(q4_material_kernel, (nelems + 255) / 256, 256, 0, law, nelems, d_conn, d_rho, par4[0], par4[1], par4[2],
              par4[3], d_rhoE, d_ks, d_ms, d_dk, d_dm);
  EIGD_CHECK_LAUNCH();
  return 0;
}
