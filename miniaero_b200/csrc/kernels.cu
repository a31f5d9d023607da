// CUDA kernels of the RK4 finite-volume stage for sm_100a.  Compiled twice (see kernels.h):
//   ma_fast   : -fmad=true  (production)
//   ma_strict : -fmad=false -DMA_STRICT (bit-for-bit with the reference's -DCELL_FLUX build)
//
// Two kernels per RK stage, one CTA per tile of cells:
//   grad_limiter_kernel : thread per cell.  Green-Gauss gradient (GreenGauss.h:51-270), stencil
//                         min/max (StencilLimiter.h:56-282) and Venkatakrishnan limiter
//                         (StencilLimiter.h:356-500, VenkatLimiter.h:45-73) gathered over the cell's six
//                         faces in slot order; the reference's per-face scratch arrays (cell_gradient_,
//                         stored_min/max/limiter) never exist.
//   flux_rk_kernel      : phase 1, thread per tile face: Roe (+ viscous) flux or boundary-condition flux,
//                         each face of the tile evaluated once and staged in shared memory;
//                         phase 2, thread per cell: deterministic slot-ordered gather of the six face
//                         fluxes (Flux.h:216-227, the reference's -DCELL_FLUX order), residual, and the
//                         fused RK update (TimeSolverExplicitRK4.h:106-128,355,483).
#include "kernels.h"
#include "physics.cuh"

namespace MA_NS {

using ma::DevMesh;
using ma::StageArgs;
using ma::TileInfoDev;

MA_DEV void load_state(const double *__restrict__ base, int stride, int c, double (&v)[5]) {
#pragma unroll
  for (int k = 0; k < 5; ++k) v[k] = __ldg(base + (size_t)k * stride + c);
}

// Face geometry as the kernels see it.  STRICT keeps the caller's tangent and binormal (the reference
// normalises and uses them, Roe_Flux.h:101-123); FAST needs only the area vector (see roe_flux_normal_only).
#ifdef MA_STRICT
#define MA_GEOM_XF 9
struct FaceGeom {
  double n[3], t[3], b[3];
};
MA_DEV void load_face_geom(const DevMesh &m, size_t NF, int j, FaceGeom &g) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    g.n[d] = __ldg(m.face_geom + (size_t)(0 + d) * NF + j);
    g.t[d] = __ldg(m.face_geom + (size_t)(3 + d) * NF + j);
    g.b[d] = __ldg(m.face_geom + (size_t)(6 + d) * NF + j);
  }
}
MA_DEV void face_roe_flux(const double (&Vl)[5], const double (&Vr)[5], const FaceGeom &g, double (&flux)[5]) {
  roe_flux(Vl, Vr, g.n, g.t, g.b, flux);
}
#else
#define MA_GEOM_XF 3
struct FaceGeom {
  double n[3];
};
MA_DEV void load_face_geom(const DevMesh &m, size_t NF, int j, FaceGeom &g) {
#pragma unroll
  for (int d = 0; d < 3; ++d) g.n[d] = __ldg(m.face_geom + (size_t)d * NF + j);
}
MA_DEV void face_roe_flux(const double (&Vl)[5], const double (&Vr)[5], const FaceGeom &g, double (&flux)[5]) {
  roe_flux_normal_only(Vl, Vr, g.n, flux);
}
#endif

// ------------------------------------------------------------------------------------------------------
// V = primitives (rho, u, v, w, T) of the stage state, [5][stride], ghosts included.
template <bool SECOND>
__global__ void __launch_bounds__(256) grad_limiter_kernel(const DevMesh m, const double *__restrict__ V_,
                                                           double *__restrict__ grad, double *__restrict__ lim,
                                                           int tile_begin) {
  const TileInfoDev T = m.tiles[tile_begin + blockIdx.x];
  const size_t NF = (size_t)m.n_tile_faces;
  for (int lc = threadIdx.x; lc < T.cell_count; lc += blockDim.x) {
    const int c = T.cell_start + lc;
    double V[5];
    load_state(V_, m.stride, c, V);
    const double vol = __ldg(m.cell_vol + c);
    double g[5][3];
    double mn[5], mx[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      g[k][0] = g[k][1] = g[k][2] = 0;
      mn[k] = 1.0e300;   // StencilLimiter.h:227-228
      mx[k] = -1.0e300;
    }
    int fj[6];
#pragma unroll
    for (int s = 0; s < 6; ++s) {  // slot order == the reference's gather order (GreenGauss.h:255-267)
      const unsigned sf = m.slot_face[(size_t)s * m.slot_stride + c];
      const int side = sf >> 15;
      const int j = T.face_start + (int)(sf & 0x7fffu);
      fj[s] = j;
      const int r = __ldg(m.face_right + j);
      double n[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) n[d] = __ldg(m.face_geom + (size_t)d * NF + j);
      if (r >= 0) {
        const int nb = side ? __ldg(m.face_left + j) : r;
        double Vn[5];
        load_state(V_, m.stride, nb, Vn);
#ifdef MA_STRICT
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const double avg = 0.5 * (V[k] + Vn[k]);  // GreenGauss.h:117
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const double q = avg * n[d] / vol;
            g[k][d] += side ? -q : q;  // GreenGauss.h:130-131
          }
        }
#else
        // FAST: the 0.5 and 1/vol factors are applied once after the slot loop
        const double sgn = side ? -1.0 : 1.0;
        const double sn[3] = {sgn * n[0], sgn * n[1], sgn * n[2]};
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const double sum = V[k] + Vn[k];
#pragma unroll
          for (int d = 0; d < 3; ++d) g[k][d] = fma(sum, sn[d], g[k][d]);
        }
#endif
        if (SECOND) {
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            mn[k] = fmin(mn[k], fmin(Vn[k], V[k]));  // StencilLimiter.h:139-140,272-273
            mx[k] = fmax(mx[k], fmax(Vn[k], V[k]));
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
#ifdef MA_STRICT
#pragma unroll
          for (int d = 0; d < 3; ++d) g[k][d] += V[k] * n[d] / vol;  // GreenGauss.h:186-216
#else
          const double two_v = 2.0 * V[k];
#pragma unroll
          for (int d = 0; d < 3; ++d) g[k][d] = fma(two_v, n[d], g[k][d]);
#endif
          if (SECOND) {
            mn[k] = fmin(mn[k], V[k]);
            mx[k] = fmax(mx[k], V[k]);
          }
        }
      }
    }
#ifndef MA_STRICT
    {
      const double half_rvol = 0.5 * rcp(vol);
#pragma unroll
      for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int d = 0; d < 3; ++d) g[k][d] *= half_rvol;
    }
#endif
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
      for (int d = 0; d < 3; ++d) grad[(size_t)(k * 3 + d) * m.stride + c] = g[k][d];

    if (SECOND) {
      double xc[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) xc[d] = __ldg(m.cell_xyz + (size_t)d * m.stride + c);
#ifdef MA_STRICT
      double phi[5] = {1.0, 1.0, 1.0, 1.0, 1.0};  // StencilLimiter.h:308-311
#else
      double pN[5] = {1.0, 1.0, 1.0, 1.0, 1.0}, pD[5] = {1.0, 1.0, 1.0, 1.0, 1.0};
      double dumax[5], dumin[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        dumax[k] = mx[k] - V[k];
        dumin[k] = mn[k] - V[k];
      }
#endif
#pragma unroll
      for (int s = 0; s < 6; ++s) {
        const int j = fj[s];
        double disp[3];
        double dist = 0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          disp[d] = __ldg(m.face_geom + (size_t)(MA_GEOM_XF + d) * NF + j) - xc[d];  // StencilLimiter.h:425-433
          dist += disp[d] * disp[d];
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          double dU = 0;
#pragma unroll
          for (int d = 0; d < 3; ++d) dU += disp[d] * g[k][d];  // StencilLimiter.h:438-446
#ifdef MA_STRICT
          const double dumax = mx[k] - V[k];
          const double dumin = mn[k] - V[k];
          phi[k] = fmin(phi[k], venkat_limit(dumax, dumin, dU, dist));  // StencilLimiter.h:451-455, 345-346
#else
          double N, D;
          venkat_fraction(dumax[k], dumin[k], dU, dist, N, D);
          venkat_fraction_min(N, D, pN[k], pD[k]);
#endif
        }
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) {
#ifdef MA_STRICT
        lim[(size_t)k * m.stride + c] = phi[k];
#else
        lim[(size_t)k * m.stride + c] = quot(pN[k], pD[k]);
#endif
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------
template <bool SECOND, bool VISCOUS>
__global__ void __launch_bounds__(256) flux_rk_kernel(const DevMesh m, const StageArgs a, int tile_begin) {
  extern __shared__ double sflux[];
  const TileInfoDev T = m.tiles[tile_begin + blockIdx.x];
  const size_t NF = (size_t)m.n_tile_faces;
  const int FS = m.flux_smem_stride;
  const double *__restrict__ V_ = a.V;

  // ---- phase 1: one flux per tile face
  for (int e = threadIdx.x; e < T.face_count; e += blockDim.x) {
    const int j = T.face_start + e;
    const int l = __ldg(m.face_left + j);
    const int r = __ldg(m.face_right + j);
    FaceGeom G;
    load_face_geom(m, NF, j, G);
    double Vl[5], flux[5];
    load_state(V_, m.stride, l, Vl);
    if (r >= 0) {
      // interior face: Flux.h:89-160
      double Vr[5];
      load_state(V_, m.stride, r, Vr);
      double gf[5][3];
      if (SECOND) {
        double dl[3], dr[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const double xf = __ldg(m.face_geom + (size_t)(MA_GEOM_XF + d) * NF + j);
          dl[d] = xf - __ldg(m.cell_xyz + (size_t)d * m.stride + l);
          dr[d] = xf - __ldg(m.cell_xyz + (size_t)d * m.stride + r);
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          double tl = 0, tr = 0;
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const double gl = __ldg(a.grad + (size_t)(k * 3 + d) * m.stride + l);
            const double gr = __ldg(a.grad + (size_t)(k * 3 + d) * m.stride + r);
            tl += dl[d] * gl;  // Flux.h:114-121
            tr += dr[d] * gr;
            if (VISCOUS) gf[k][d] = 0.5 * (gl + gr);  // Flux.h:146-149
          }
          Vl[k] += tl * __ldg(a.lim + (size_t)k * m.stride + l);  // Flux.h:124-127
          Vr[k] += tr * __ldg(a.lim + (size_t)k * m.stride + r);
        }
      } else if (VISCOUS) {
#pragma unroll
        for (int k = 0; k < 5; ++k)
#pragma unroll
          for (int d = 0; d < 3; ++d)
            gf[k][d] = 0.5 * (__ldg(a.grad + (size_t)(k * 3 + d) * m.stride + l) +
                              __ldg(a.grad + (size_t)(k * 3 + d) * m.stride + r));
      }
      face_roe_flux(Vl, Vr, G, flux);
      if (VISCOUS) {
        double Vf[5], vflux[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) Vf[k] = 0.5 * (Vl[k] + Vr[k]);  // Flux.h:142-143
        viscous_flux(gf, Vf, G.n, vflux);
#pragma unroll
        for (int k = 0; k < 5; ++k) flux[k] -= vflux[k];
      }
    } else {
      // boundary face, always first order (Extrapolate_BC.h, Tangent_BC.h, Inflow_BC.h, NoSlip_BC.h)
      const int type = -1 - r;
      double Vr[5];
      double area_norm = 0;
      if (type == 0) {  // Extrapolate_BC.h:82-83: Roe(V, V)
#pragma unroll
        for (int k = 0; k < 5; ++k) Vr[k] = Vl[k];
      } else if (type == 2) {  // Inflow_BC.h:84-90
        double Ui[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) Ui[k] = m.inflow[k];
        compute_primitives(Ui, Vr);
      } else {  // Tangent_BC.h:82-101, NoSlip_BC.h:96-112
        mirror_state(Vl, G.n, Vr, area_norm);
      }
      face_roe_flux(Vl, Vr, G, flux);
      if (type == 3) {  // NoSlip_BC.h:114-139 — viscous wall flux regardless of options.viscous
        double xf[3], xc[3], vflux[5];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          xf[d] = __ldg(m.face_geom + (size_t)(MA_GEOM_XF + d) * NF + j);
          xc[d] = __ldg(m.cell_xyz + (size_t)d * m.stride + l);
        }
        noslip_viscous_flux(Vl, G.n, area_norm, xf, xc, vflux);
#pragma unroll
        for (int k = 0; k < 5; ++k) flux[k] -= vflux[k];  // slot = -iflux + vflux == -(iflux - vflux)
      }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) sflux[k * FS + e] = flux[k];
  }
  __syncthreads();

  // ---- phase 2: slot-ordered gather, residual, RK update; the next stage state is stored as primitives
  for (int lc = threadIdx.x; lc < T.cell_count; lc += blockDim.x) {
    const int c = T.cell_start + lc;
#ifdef MA_STRICT
    const double dtv = a.dt / __ldg(m.cell_vol + c);  // Flux.h:224-225: dt_/volume_(i) * flux
#else
    const double dtv = a.dt * rcp(__ldg(m.cell_vol + c));
#endif
    double R[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const unsigned sf = m.slot_face[(size_t)s * m.slot_stride + c];
      const int e = (int)(sf & 0x7fffu);
      const bool right = (sf >> 15) != 0;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const double f = sflux[k * FS + e];
        R[k] = R[k] + dtv * (right ? f : -f);  // Flux.h:172-178: left slot holds -flux, right slot +flux
      }
    }
    double Wn[5];  // conservative state the next stage is evaluated at (or the new solution)
    if (a.kind == 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const double w = a.Un[(size_t)k * m.stride + c];
        a.Acc[(size_t)k * m.stride + c] = w + a.beta * R[k];
        Wn[k] = w + a.alpha_next * R[k];
      }
    } else if (a.kind == 1) {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        a.Acc[(size_t)k * m.stride + c] = a.Acc[(size_t)k * m.stride + c] + a.beta * R[k];
        Wn[k] = a.Un[(size_t)k * m.stride + c] + a.alpha_next * R[k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        Wn[k] = a.Acc[(size_t)k * m.stride + c] + a.beta * R[k];
        a.Un[(size_t)k * m.stride + c] = Wn[k];
      }
    }
    double Vn[5];
    compute_primitives(Wn, Vn);
#pragma unroll
    for (int k = 0; k < 5; ++k) a.Vnext[(size_t)k * m.stride + c] = Vn[k];
  }
}

// U (conservative, owned cells) -> V (primitives): after initial conditions / set_solution
__global__ void primitives_kernel(const DevMesh m, const double *__restrict__ Un, double *__restrict__ V) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m.n_owned) return;
  double U[5], P[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) U[k] = Un[(size_t)k * m.stride + c];
  compute_primitives(U, P);
#pragma unroll
  for (int k = 0; k < 5; ++k) V[(size_t)k * m.stride + c] = P[k];
}

__global__ void initial_conditions_kernel(const DevMesh m, double *__restrict__ Un, int sod, double midx,
                                          double s1_rho, double s1_rhoE, double s2_rho, double s2_rhoE) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m.n_owned) return;
  double u[5];
  if (sod) {  // Initial_Conditions.h:56-93
    const double x = m.cell_xyz[c];
    const bool left = x < midx;
    u[0] = left ? s1_rho : s2_rho;
    u[1] = u[2] = u[3] = 0.0;
    u[4] = left ? s1_rhoE : s2_rhoE;
  } else {  // Initial_Conditions.h:121-131
    for (int k = 0; k < 5; ++k) u[k] = m.inflow[k];
  }
  for (int k = 0; k < 5; ++k) Un[(size_t)k * m.stride + c] = u[k];
}

// ---- device-function probes ---------------------------------------------------------------------------
__global__ void probe_roe_kernel(int n, const double *vl, const double *vr, const double *nn, const double *tt,
                                 const double *bb, double *flux) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double Vl[5], Vr[5], F[5];
  FaceGeom G;
  for (int k = 0; k < 5; ++k) Vl[k] = vl[5 * i + k], Vr[k] = vr[5 * i + k];
  for (int d = 0; d < 3; ++d) {
    G.n[d] = nn[3 * i + d];
#ifdef MA_STRICT
    G.t[d] = tt[3 * i + d], G.b[d] = bb[3 * i + d];
#endif
  }
  face_roe_flux(Vl, Vr, G, F);  // the production face flux of this arithmetic mode
  for (int k = 0; k < 5; ++k) flux[5 * i + k] = F[k];
}
__global__ void probe_viscous_kernel(int n, const double *g, const double *v, const double *a, double *vf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double G[5][3], V[5], A[3], F[5];
  for (int k = 0; k < 5; ++k) {
    V[k] = v[5 * i + k];
    for (int d = 0; d < 3; ++d) G[k][d] = g[15 * i + 3 * k + d];
  }
  for (int d = 0; d < 3; ++d) A[d] = a[3 * i + d];
  viscous_flux(G, V, A, F);
  for (int k = 0; k < 5; ++k) vf[5 * i + k] = F[k];
}
__global__ void probe_primitives_kernel(int n, const double *u, double *v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double U[5], V[5];
  for (int k = 0; k < 5; ++k) U[k] = u[5 * i + k];
  compute_primitives(U, V);
  for (int k = 0; k < 5; ++k) v[5 * i + k] = V[k];
}
__global__ void probe_venkat_kernel(int n, const double *dmax, const double *dmin, const double *du,
                                    const double *dx3, double *phi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
#ifdef MA_STRICT
  phi[i] = venkat_limit(dmax[i], dmin[i], du[i], dx3[i]);
#else
  double N, D;
  venkat_fraction(dmax[i], dmin[i], du[i], dx3[i], N, D);
  phi[i] = quot(N, D);
#endif
}
__global__ void probe_vanalbada_kernel(int n, const double *dmax, const double *dmin, const double *du, double *phi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) phi[i] = vanalbada_limit(dmax[i], dmin[i], du[i]);
}

// ---- launchers ----------------------------------------------------------------------------------------
cudaError_t launch_grad_limiter(const DevMesh &m, const double *V, double *grad, double *lim, bool second,
                                int tile_begin, int ntiles, int threads, cudaStream_t st) {
  if (ntiles <= 0) return cudaSuccess;
  if (second)
    grad_limiter_kernel<true><<<ntiles, threads, 0, st>>>(m, V, grad, lim, tile_begin);
  else
    grad_limiter_kernel<false><<<ntiles, threads, 0, st>>>(m, V, grad, lim, tile_begin);
  return cudaGetLastError();
}

cudaError_t flux_rk_prepare(int smem_bytes) {
  cudaError_t e;
#define MA_SET(K)                                                                          \
  e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);    \
  if (e != cudaSuccess) return e;
  MA_SET((flux_rk_kernel<false, false>))
  MA_SET((flux_rk_kernel<false, true>))
  MA_SET((flux_rk_kernel<true, false>))
  MA_SET((flux_rk_kernel<true, true>))
#undef MA_SET
  return cudaSuccess;
}

cudaError_t launch_flux_rk(const DevMesh &m, const StageArgs &a, bool second, bool viscous, int tile_begin,
                           int ntiles, int threads, cudaStream_t st) {
  if (ntiles <= 0) return cudaSuccess;
  const size_t smem = (size_t)5 * m.flux_smem_stride * sizeof(double);
  if (second && viscous)
    flux_rk_kernel<true, true><<<ntiles, threads, smem, st>>>(m, a, tile_begin);
  else if (second)
    flux_rk_kernel<true, false><<<ntiles, threads, smem, st>>>(m, a, tile_begin);
  else if (viscous)
    flux_rk_kernel<false, true><<<ntiles, threads, smem, st>>>(m, a, tile_begin);
  else
    flux_rk_kernel<false, false><<<ntiles, threads, smem, st>>>(m, a, tile_begin);
  return cudaGetLastError();
}

cudaError_t launch_primitives(const DevMesh &m, const double *Un, double *V, cudaStream_t st) {
  const int threads = 256;
  primitives_kernel<<<(m.n_owned + threads - 1) / threads, threads, 0, st>>>(m, Un, V);
  return cudaGetLastError();
}

cudaError_t launch_initial_conditions(const DevMesh &m, double *Un, int problem_type, double midx, cudaStream_t st) {
  // Initial_Conditions.h:58-71 (host arithmetic, evaluated without FMA contraction)
  const double Rgas = 287.05;
  const double gamma = 1.4;
  const double Cv = Rgas / (gamma - 1.0);
  double P1 = 68947.57, T1 = 288.889, P2 = 6894.757, T2 = 231.11;
  volatile double density1 = P1 / (Rgas * T1);
  volatile double cvt1 = Cv * T1;
  double rhoE1 = density1 * cvt1;
  volatile double density2 = P2 / (Rgas * T2);
  volatile double cvt2 = Cv * T2;
  double rhoE2 = density2 * cvt2;
  const int threads = 256;
  initial_conditions_kernel<<<(m.n_owned + threads - 1) / threads, threads, 0, st>>>(
      m, Un, problem_type == 0 ? 1 : 0, midx, density1, rhoE1, density2, rhoE2);
  return cudaGetLastError();
}

#define MA_PROBE_LAUNCH(kernel, ...)                                      \
  if (n <= 0) return cudaSuccess;                                         \
  kernel<<<(n + 127) / 128, 128, 0, st>>>(n, __VA_ARGS__);                \
  return cudaGetLastError();

cudaError_t probe_roe(int n, const double *vl, const double *vr, const double *nn, const double *tt,
                      const double *bb, double *flux, cudaStream_t st) {
  MA_PROBE_LAUNCH(probe_roe_kernel, vl, vr, nn, tt, bb, flux)
}
cudaError_t probe_viscous(int n, const double *g, const double *v, const double *a, double *vf, cudaStream_t st) {
  MA_PROBE_LAUNCH(probe_viscous_kernel, g, v, a, vf)
}
cudaError_t probe_primitives(int n, const double *u, double *v, cudaStream_t st) {
  MA_PROBE_LAUNCH(probe_primitives_kernel, u, v)
}
cudaError_t probe_venkat(int n, const double *dmax, const double *dmin, const double *du, const double *dx3,
                         double *phi, cudaStream_t st) {
  MA_PROBE_LAUNCH(probe_venkat_kernel, dmax, dmin, du, dx3, phi)
}
cudaError_t probe_vanalbada(int n, const double *dmax, const double *dmin, const double *du, double *phi,
                            cudaStream_t st) {
  MA_PROBE_LAUNCH(probe_vanalbada_kernel, dmax, dmin, du, phi)
}

}  // namespace MA_NS
