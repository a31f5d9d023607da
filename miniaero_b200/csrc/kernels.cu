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

// launch bounds (max threads per CTA, min resident CTAs per SM) of the two stage kernels
#ifndef MA_GRAD_THREADS
#define MA_GRAD_THREADS 256
#define MA_GRAD_MINB 2
#endif
#ifndef MA_FLUX_THREADS
#define MA_FLUX_THREADS 256
#define MA_FLUX_MINB 2
#endif

namespace MA_NS {

using ma::DevMesh;
using ma::StageArgs;
using ma::TileInfoDev;

// 8-byte asynchronous global -> shared copy (LDGSTS): the data lands in shared memory without occupying a
// register or a scoreboard slot of the issuing thread
MA_DEV void cp_async8(double *smem_dst, const double *gmem_src) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gmem_src) : "memory");
}
MA_DEV void cp_async_commit_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

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
__global__ void __launch_bounds__(MA_GRAD_THREADS, MA_GRAD_MINB) grad_limiter_kernel(const DevMesh m, const double *__restrict__ V_,
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
#ifdef MA_STRICT
      mn[k] = 1.0e300;   // StencilLimiter.h:227-228
      mx[k] = -1.0e300;
#else
      mn[k] = mx[k] = V[k];  // every face contributes min/max(V, Vn): start from V, fold the neighbours in
#endif
    }
    int fj[6];
#pragma unroll
    for (int s = 0; s < 6; ++s) {  // slot order == the reference's gather order (GreenGauss.h:255-267)
      const unsigned sf = m.slot_face[(size_t)s * m.slot_stride + c];
      const int side = sf >> 15;
      const int j = T.face_start + (int)(sf & 0x3fffu);
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
#ifdef MA_STRICT
            mn[k] = fmin(mn[k], fmin(Vn[k], V[k]));  // StencilLimiter.h:139-140,272-273
            mx[k] = fmax(mx[k], fmax(Vn[k], V[k]));
#else
            mn[k] = fmin(mn[k], Vn[k]);
            mx[k] = fmax(mx[k], Vn[k]);
#endif
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
#ifdef MA_STRICT
          if (SECOND) {
            mn[k] = fmin(mn[k], V[k]);
            mx[k] = fmax(mx[k], V[k]);
          }
#endif
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
      double dumax[5], ndumin[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        dumax[k] = mx[k] - V[k];
        ndumin[k] = V[k] - mn[k];
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
          // VenkatLimiter.h:45-73 with a = |du|, mm = |dumax| or |dumin| by the sign of du and the common
          // factor du cancelled: phi = (mm^2 + eps2 + 2 a mm) / (mm^2 + eps2 + a (2a + mm)); phi -> 1 as a -> 0
          const double aa = fabs(dU);
          const double mm = dU > 0.0 ? dumax[k] : ndumin[k];
          const double base = fma(mm, mm, dist);
          const double a2 = aa + aa;
          venkat_fraction_min(fma(a2, mm, base), fma(aa, a2 + mm, base), pN[k], pD[k]);
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
__global__ void __launch_bounds__(MA_FLUX_THREADS, MA_FLUX_MINB) flux_rk_kernel(const DevMesh m, const StageArgs a, int tile_begin) {
  extern __shared__ double sflux[];
  const TileInfoDev T = m.tiles[tile_begin + blockIdx.x];
  const size_t NF = (size_t)m.n_tile_faces;
  const int FS = m.flux_smem_stride;
  const int CS = m.rk_smem_stride;
  const double *__restrict__ V_ = a.V;
  double *srk = sflux + 5 * FS;  // [11][CS]: volume, Un[5], Acc[5] of the tile's cells for phase 2

  // ---- phase 0: start the asynchronous copy of the phase-2 operands; it completes behind phase 1
  for (int lc = threadIdx.x; lc < T.cell_count; lc += blockDim.x) {
    const int c = T.cell_start + lc;
    cp_async8(srk + lc, m.cell_vol + c);
    if (a.kind != 2) {
#pragma unroll
      for (int k = 0; k < 5; ++k) cp_async8(srk + (1 + k) * CS + lc, a.Un + (size_t)k * m.stride + c);
    }
    if (a.kind != 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k) cp_async8(srk + (6 + k) * CS + lc, a.Acc + (size_t)k * m.stride + c);
    }
  }

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
#ifdef MA_STRICT
            if (VISCOUS) gf[k][d] = 0.5 * (gl + gr);  // Flux.h:146-149
#else
            if (VISCOUS) gf[k][d] = gl + gr;  // twice the face gradient; the 0.5 is folded into the viscosity
#endif
          }
          Vl[k] += tl * __ldg(a.lim + (size_t)k * m.stride + l);  // Flux.h:124-127
          Vr[k] += tr * __ldg(a.lim + (size_t)k * m.stride + r);
        }
      } else if (VISCOUS) {
#pragma unroll
        for (int k = 0; k < 5; ++k)
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const double gsum = __ldg(a.grad + (size_t)(k * 3 + d) * m.stride + l) +
                                __ldg(a.grad + (size_t)(k * 3 + d) * m.stride + r);
#ifdef MA_STRICT
            gf[k][d] = 0.5 * gsum;
#else
            gf[k][d] = gsum;
#endif
          }
      }
      face_roe_flux(Vl, Vr, G, flux);
      if (VISCOUS) {
#ifdef MA_STRICT
        double Vf[5], vflux[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) Vf[k] = 0.5 * (Vl[k] + Vr[k]);  // Flux.h:142-143
        viscous_flux(gf, Vf, G.n, vflux);
#pragma unroll
        for (int k = 0; k < 5; ++k) flux[k] -= vflux[k];
#else
        // Viscous_Flux.h:65-98 with gf = 2 x face gradient: q_i = sum_j tau_ij(gf) a_j, hh = gf_T . a
        const double third_div = (gf[1][0] + gf[2][1] + gf[3][2]) * (1.0 / 3.0);
        const double txx = gf[1][0] - third_div, tyy = gf[2][1] - third_div, tzz = gf[3][2] - third_div;
        const double txy = 0.5 * (gf[1][1] + gf[2][0]), txz = 0.5 * (gf[1][2] + gf[3][0]);
        const double tyz = 0.5 * (gf[2][2] + gf[3][1]);
        const double q0 = txx * G.n[0] + txy * G.n[1] + txz * G.n[2];
        const double q1 = txy * G.n[0] + tyy * G.n[1] + tyz * G.n[2];
        const double q2 = txz * G.n[0] + tyz * G.n[1] + tzz * G.n[2];
        const double hh = gf[4][0] * G.n[0] + gf[4][1] * G.n[1] + gf[4][2] * G.n[2];
        const double mu = compute_viscosity(0.5 * (Vl[4] + Vr[4]));
        const double uq = (Vl[1] + Vr[1]) * q0 + (Vl[2] + Vr[2]) * q1 + (Vl[3] + Vr[3]) * q2;
        flux[1] -= mu * q0;
        flux[2] -= mu * q1;
        flux[3] -= mu * q2;
        flux[4] -= fma(0.5 * mu, uq, 0.5 * compute_thermal_conductivity(mu) * hh);
#endif
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
  cp_async_commit_wait_all();
  __syncthreads();

  // ---- phase 2: slot-ordered gather, residual, RK update; the next stage state is stored as primitives
  for (int lc = threadIdx.x; lc < T.cell_count; lc += blockDim.x) {
    const int c = T.cell_start + lc;
#ifdef MA_STRICT
    const double dtv = a.dt / srk[lc];  // Flux.h:224-225: dt_/volume_(i) * flux
#else
    const double dtv = a.dt * rcp(srk[lc]);
#endif
    double R[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const unsigned sf = m.slot_face[(size_t)s * m.slot_stride + c];
      const int e = (int)(sf & 0x3fffu);
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
        const double w = srk[(1 + k) * CS + lc];
        a.Acc[(size_t)k * m.stride + c] = w + a.beta * R[k];
        Wn[k] = w + a.alpha_next * R[k];
      }
    } else if (a.kind == 1) {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        a.Acc[(size_t)k * m.stride + c] = srk[(6 + k) * CS + lc] + a.beta * R[k];
        Wn[k] = srk[(1 + k) * CS + lc] + a.alpha_next * R[k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        Wn[k] = srk[(6 + k) * CS + lc] + a.beta * R[k];
        a.Un[(size_t)k * m.stride + c] = Wn[k];
      }
    }
    double Vn[5];
    compute_primitives(Wn, Vn);
#pragma unroll
    for (int k = 0; k < 5; ++k) a.Vnext[(size_t)k * m.stride + c] = Vn[k];
  }
}

#ifndef MA_STRICT
// ======================================================================================================
// FAST tile kernels: cell-block tiles staged in shared memory.
//
// Every global load of cell data is issued by a thread that owns a whole cell (own cells: coalesced SoA
// reads with no index indirection; outside cells of cut faces: one gather per cut face), so the per-face
// gather of 2 x 28 doubles of the reference's compute_face_flux (Flux.h:89-132) never happens.
// ======================================================================================================

// ---- sweep 1: Green-Gauss gradient + stencil min/max + Venkatakrishnan limiter ----------------------
// phase 0: primitives of the tile's cells and of the outside cells of its cut faces -> sV[5][LS]
// phase 1: thread per cell, neighbours read from shared memory through the tile-local face table
template <bool SECOND>
__global__ void __launch_bounds__(MA_GRAD_THREADS, MA_GRAD_MINB)
    grad_limiter_tile_kernel(const DevMesh m, const double *__restrict__ V_, double *__restrict__ grad,
                             double *__restrict__ lim, int tile_begin) {
  extern __shared__ double smem[];
  const TileInfoDev T = m.tiles[tile_begin + blockIdx.x];
  const size_t NF = (size_t)m.n_tile_faces;
  const int LS = m.local_smem_stride;
  const int nc = T.cell_count;
  const int nl = nc + (T.face_count - T.cut_start);
  double *sV = smem;
  for (int i = threadIdx.x; i < nl; i += blockDim.x) {
    const int c = i < nc ? T.cell_start + i : __ldg(m.tile_halo + T.halo_start + (i - nc));
#pragma unroll
    for (int k = 0; k < 5; ++k) sV[k * LS + i] = __ldg(V_ + (size_t)k * m.stride + c);
  }
  __syncthreads();

  for (int lc = threadIdx.x; lc < nc; lc += blockDim.x) {
    const int c = T.cell_start + lc;
    double V[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) V[k] = sV[k * LS + lc];
    const double vol = __ldg(m.cell_vol + c);
    double g[5][3], mn[5], mx[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      g[k][0] = g[k][1] = g[k][2] = 0;
      mn[k] = mx[k] = V[k];  // min/max over {cell, face neighbours} (StencilLimiter.h:139-140, 272-273)
    }
    unsigned sfs[6];
#pragma unroll
    for (int s = 0; s < 6; ++s) sfs[s] = m.slot_face[(size_t)s * m.slot_stride + c];
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const unsigned sf = sfs[s];
      const int side = sf >> 15;
      const int j = T.face_start + (int)(sf & 0x3fffu);
      const double sgn = side ? -1.0 : 1.0;
      double sn[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) sn[d] = sgn * __ldg(m.face_geom + (size_t)d * NF + j);
      if (!(sf & 0x4000u)) {
        const unsigned lr = __ldg(m.face_lr + j);
        const int nb = side ? (int)(lr & 0xffffu) : (int)(lr >> 16);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const double vn = sV[k * LS + nb];
          const double sum = V[k] + vn;  // 2 x GreenGauss.h:117; the 0.5/vol factor is applied after the loop
#pragma unroll
          for (int d = 0; d < 3; ++d) g[k][d] = fma(sum, sn[d], g[k][d]);
          if (SECOND) {
            mn[k] = fmin(mn[k], vn);
            mx[k] = fmax(mx[k], vn);
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const double two_v = 2.0 * V[k];  // GreenGauss.h:186-216
#pragma unroll
          for (int d = 0; d < 3; ++d) g[k][d] = fma(two_v, sn[d], g[k][d]);
        }
      }
    }
    {
      const double half_rvol = 0.5 * rcp(vol);
#pragma unroll
      for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          g[k][d] *= half_rvol;
          grad[(size_t)(k * 3 + d) * m.stride + c] = g[k][d];
        }
    }
    if (SECOND) {
      double xc[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) xc[d] = __ldg(m.cell_xyz + (size_t)d * m.stride + c);
      double pN[5] = {1.0, 1.0, 1.0, 1.0, 1.0}, pD[5] = {1.0, 1.0, 1.0, 1.0, 1.0};
      double dumax[5], ndumin[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        dumax[k] = mx[k] - V[k];
        ndumin[k] = V[k] - mn[k];
      }
#pragma unroll
      for (int s = 0; s < 6; ++s) {
        const int j = T.face_start + (int)(sfs[s] & 0x3fffu);
        double disp[3];
        double dist = 0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          disp[d] = __ldg(m.face_geom + (size_t)(3 + d) * NF + j) - xc[d];  // StencilLimiter.h:425-433
          dist = fma(disp[d], disp[d], dist);
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const double dU = disp[0] * g[k][0] + disp[1] * g[k][1] + disp[2] * g[k][2];  // StencilLimiter.h:438-446
          // VenkatLimiter.h:45-73 with a = |du|, mm = |dumax| or |dumin| by the sign of du and the common factor
          // du cancelled: phi = (mm^2 + eps2 + 2 a mm) / (mm^2 + eps2 + a (2a + mm)); phi -> 1 as a -> 0
          const double a = fabs(dU);
          const double mm = dU > 0.0 ? dumax[k] : ndumin[k];
          const double base = fma(mm, mm, dist);
          const double a2 = a + a;
          const double N = fma(a2, mm, base);
          const double D = fma(a, a2 + mm, base);
          venkat_fraction_min(N, D, pN[k], pD[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) lim[(size_t)k * m.stride + c] = quot(pN[k], pD[k]);
    }
  }
}

// ---- sweep 2: face fluxes + slot-ordered gather + RK stage update -----------------------------------
// Shared memory S[NC][FS] (component-major, one column per tile face):
//   phase A  thread per cell (own cells, then the outside cell of every cut face): loads the cell's record
//            (V 5, grad 15, lim 5, xyz 3) once and, for each of its faces in the tile, stores that side's
//            limited-extrapolated primitives (Flux.h:109-132) and — viscous — its half of the face stress:
//            q_i = sum_j tau_ij(grad) a_j, h = grad(T).a  (Viscous_Flux.h:65-98 is linear in the gradient, and
//            the face gradient is the plain average of the two cell gradients, Flux.h:146-149)
//   phase B  thread per face: Roe flux of the two staged states (+ viscous flux from the staged halves), or
//            the boundary-condition flux; the result overwrites the face's column
//   phase C  thread per cell: gather of the six face fluxes in slot order (Flux.h:216-227), RK update
template <bool SECOND, bool VISCOUS>
__global__ void __launch_bounds__(MA_FLUX_THREADS, MA_FLUX_MINB)
    flux_rk_tile_kernel(const DevMesh m, const StageArgs a, int tile_begin) {
  extern __shared__ double S[];
  constexpr int SIDE = VISCOUS ? 9 : 5;  // components per side: V'[5] (+ q[3], h)
  const TileInfoDev T = m.tiles[tile_begin + blockIdx.x];
  const size_t NF = (size_t)m.n_tile_faces;
  const int FS = m.flux_smem_stride;
  const int nc = T.cell_count;
  const int njobs = nc + (T.face_count - T.cut_start);

  // ---- phase A
  for (int i = threadIdx.x; i < njobs; i += blockDim.x) {
    const bool own = i < nc;
    const int c = own ? T.cell_start + i : __ldg(m.tile_halo + T.halo_start + (i - nc));
    double V[5], lm[5], xc[3], g[5][3];
#pragma unroll
    for (int k = 0; k < 5; ++k) V[k] = __ldg(a.V + (size_t)k * m.stride + c);
    if (SECOND || VISCOUS) {
#pragma unroll
      for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int d = 0; d < 3; ++d) g[k][d] = __ldg(a.grad + (size_t)(k * 3 + d) * m.stride + c);
    }
    if (SECOND) {
#pragma unroll
      for (int k = 0; k < 5; ++k) lm[k] = __ldg(a.lim + (size_t)k * m.stride + c);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) xc[d] = __ldg(m.cell_xyz + (size_t)d * m.stride + c);
    double txx = 0, tyy = 0, tzz = 0, txy = 0, txz = 0, tyz = 0;
    if (VISCOUS) {  // tau_ij = S_ij - delta_ij div/3 (Viscous_Flux.h:80-90)
      const double third_div = (g[1][0] + g[2][1] + g[3][2]) * (1.0 / 3.0);
      txx = g[1][0] - third_div, tyy = g[2][1] - third_div, tzz = g[3][2] - third_div;
      txy = 0.5 * (g[1][1] + g[2][0]), txz = 0.5 * (g[1][2] + g[3][0]), tyz = 0.5 * (g[2][2] + g[3][1]);
    }
    unsigned sfs[6];
    int nslots = 1;
    if (own) {
      nslots = 6;
#pragma unroll
      for (int s = 0; s < 6; ++s) sfs[s] = m.slot_face[(size_t)s * m.slot_stride + c];
    } else {
      const int e = T.cut_start + (i - nc);
      const unsigned lr = __ldg(m.face_lr + T.face_start + e);
      sfs[0] = (unsigned)e | (((lr & 0xffffu) == (unsigned)i) ? 0u : 0x8000u);
    }
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      if (s < nslots) {
        const unsigned sf = sfs[s];
        const int e = (int)(sf & 0x3fffu);
        const int j = T.face_start + e;
        double *col = S + ((sf >> 15) ? SIDE * FS : 0) + e;
        if (sf & 0x4000u) {
          // boundary face: first order (the *_BC.h functors take the cell state as is); the cell centroid is
          // parked in the unused right half for NoSlip_BC.h:114-139
#pragma unroll
          for (int k = 0; k < 5; ++k) col[k * FS] = V[k];
#pragma unroll
          for (int d = 0; d < 3; ++d) col[(SIDE + d) * FS] = xc[d];
        } else {
          double dx[3];
          if (SECOND) {
#pragma unroll
            for (int d = 0; d < 3; ++d) dx[d] = __ldg(m.face_geom + (size_t)(3 + d) * NF + j) - xc[d];
          }
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            double v = V[k];
            if (SECOND) v = fma(lm[k], dx[0] * g[k][0] + dx[1] * g[k][1] + dx[2] * g[k][2], v);  // Flux.h:114-127
            col[k * FS] = v;
          }
          if (VISCOUS) {
            double av[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) av[d] = __ldg(m.face_geom + (size_t)d * NF + j);
            col[5 * FS] = txx * av[0] + txy * av[1] + txz * av[2];
            col[6 * FS] = txy * av[0] + tyy * av[1] + tyz * av[2];
            col[7 * FS] = txz * av[0] + tyz * av[1] + tzz * av[2];
            col[8 * FS] = g[4][0] * av[0] + g[4][1] * av[1] + g[4][2] * av[2];
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- phase B
  for (int e = threadIdx.x; e < T.face_count; e += blockDim.x) {
    const int j = T.face_start + e;
    const unsigned r16 = __ldg(m.face_lr + j) >> 16;
    FaceGeom G;
    load_face_geom(m, NF, j, G);
    double Vl[5], Vr[5], flux[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) Vl[k] = S[k * FS + e];
    if (r16 < 0xFFF0u) {
#pragma unroll
      for (int k = 0; k < 5; ++k) Vr[k] = S[(SIDE + k) * FS + e];
      face_roe_flux(Vl, Vr, G, flux);
      if (VISCOUS) {
        // Viscous_Flux.h:65-98 at the face state 0.5 (Vl' + Vr') (Flux.h:142-143) from the staged halves
        const double Tf = 0.5 * (Vl[4] + Vr[4]);
        const double mu = compute_viscosity(Tf);
        const double kappa_half = 0.5 * compute_thermal_conductivity(mu);
        const double q0 = S[5 * FS + e] + S[(SIDE + 5) * FS + e];
        const double q1 = S[6 * FS + e] + S[(SIDE + 6) * FS + e];
        const double q2 = S[7 * FS + e] + S[(SIDE + 7) * FS + e];
        const double hh = S[8 * FS + e] + S[(SIDE + 8) * FS + e];
        const double half_mu = 0.5 * mu;
        const double uq = (Vl[1] + Vr[1]) * q0 + (Vl[2] + Vr[2]) * q1 + (Vl[3] + Vr[3]) * q2;  // 2 u_face . q
        flux[1] -= mu * q0;
        flux[2] -= mu * q1;
        flux[3] -= mu * q2;
        flux[4] -= fma(half_mu, uq, kappa_half * hh);
      }
    } else {
      const int type = (int)(0xFFFFu - r16);
      double area_norm = 0;
      if (type == 0) {  // Extrapolate_BC.h:82-83
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
      if (type == 3) {  // NoSlip_BC.h:114-139
        double xf[3], xc[3], vflux[5];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          xf[d] = __ldg(m.face_geom + (size_t)(3 + d) * NF + j);
          xc[d] = S[(SIDE + d) * FS + e];
        }
        noslip_viscous_flux(Vl, G.n, area_norm, xf, xc, vflux);
#pragma unroll
        for (int k = 0; k < 5; ++k) flux[k] -= vflux[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) S[k * FS + e] = flux[k];
  }
  __syncthreads();

  // ---- phase C
  for (int lc = threadIdx.x; lc < nc; lc += blockDim.x) {
    const int c = T.cell_start + lc;
    const double dtv = a.dt * rcp(__ldg(m.cell_vol + c));
    double R[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const unsigned sf = m.slot_face[(size_t)s * m.slot_stride + c];
      const int e = (int)(sf & 0x3fffu);
      const double sg = (sf >> 15) ? dtv : -dtv;  // Flux.h:172-178: left slot holds -flux, right slot +flux
#pragma unroll
      for (int k = 0; k < 5; ++k) R[k] = fma(sg, S[k * FS + e], R[k]);
    }
    double Wn[5];
    if (a.kind == 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const double w = a.Un[(size_t)k * m.stride + c];
        a.Acc[(size_t)k * m.stride + c] = fma(a.beta, R[k], w);
        Wn[k] = fma(a.alpha_next, R[k], w);
      }
    } else if (a.kind == 1) {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        a.Acc[(size_t)k * m.stride + c] = fma(a.beta, R[k], a.Acc[(size_t)k * m.stride + c]);
        Wn[k] = fma(a.alpha_next, R[k], a.Un[(size_t)k * m.stride + c]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        Wn[k] = fma(a.beta, R[k], a.Acc[(size_t)k * m.stride + c]);
        a.Un[(size_t)k * m.stride + c] = Wn[k];
      }
    }
    double Vn[5];
    compute_primitives(Wn, Vn);
#pragma unroll
    for (int k = 0; k < 5; ++k) a.Vnext[(size_t)k * m.stride + c] = Vn[k];
  }
}
#endif  // !MA_STRICT

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
static size_t gather_flux_smem(const DevMesh &m) {
  return ((size_t)5 * m.flux_smem_stride + (size_t)11 * m.rk_smem_stride) * sizeof(double);
}
#ifdef MA_STRICT
size_t grad_smem_bytes(const DevMesh &) { return 0; }
size_t flux_smem_bytes(const DevMesh &m, bool, bool) { return gather_flux_smem(m); }
#else
size_t grad_smem_bytes(const DevMesh &m) {
  return m.grad_variant == 1 ? (size_t)5 * m.local_smem_stride * sizeof(double) : 0;
}
size_t flux_smem_bytes(const DevMesh &m, bool, bool viscous) {
  if (m.flux_variant != 1) return gather_flux_smem(m);
  return (size_t)(viscous ? 18 : 10) * m.flux_smem_stride * sizeof(double);
}
#endif

cudaError_t launch_grad_limiter(const DevMesh &m, const double *V, double *grad, double *lim, bool second,
                                int tile_begin, int ntiles, int threads, cudaStream_t st) {
  if (ntiles <= 0) return cudaSuccess;
  const size_t smem = grad_smem_bytes(m);
#ifndef MA_STRICT
  if (m.grad_variant == 1) {
    if (second)
      grad_limiter_tile_kernel<true><<<ntiles, threads, smem, st>>>(m, V, grad, lim, tile_begin);
    else
      grad_limiter_tile_kernel<false><<<ntiles, threads, smem, st>>>(m, V, grad, lim, tile_begin);
    return cudaGetLastError();
  }
#endif
  if (second)
    grad_limiter_kernel<true><<<ntiles, threads, smem, st>>>(m, V, grad, lim, tile_begin);
  else
    grad_limiter_kernel<false><<<ntiles, threads, smem, st>>>(m, V, grad, lim, tile_begin);
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
#ifndef MA_STRICT
  MA_SET((flux_rk_tile_kernel<false, false>))
  MA_SET((flux_rk_tile_kernel<false, true>))
  MA_SET((flux_rk_tile_kernel<true, false>))
  MA_SET((flux_rk_tile_kernel<true, true>))
  MA_SET((grad_limiter_tile_kernel<false>))
  MA_SET((grad_limiter_tile_kernel<true>))
#endif
#undef MA_SET
  return cudaSuccess;
}

cudaError_t launch_flux_rk(const DevMesh &m, const StageArgs &a, bool second, bool viscous, int tile_begin,
                           int ntiles, int threads, cudaStream_t st) {
  if (ntiles <= 0) return cudaSuccess;
  const size_t smem = flux_smem_bytes(m, second, viscous);
#ifndef MA_STRICT
  if (m.flux_variant == 1) {
    if (second && viscous)
      flux_rk_tile_kernel<true, true><<<ntiles, threads, smem, st>>>(m, a, tile_begin);
    else if (second)
      flux_rk_tile_kernel<true, false><<<ntiles, threads, smem, st>>>(m, a, tile_begin);
    else if (viscous)
      flux_rk_tile_kernel<false, true><<<ntiles, threads, smem, st>>>(m, a, tile_begin);
    else
      flux_rk_tile_kernel<false, false><<<ntiles, threads, smem, st>>>(m, a, tile_begin);
    return cudaGetLastError();
  }
#endif
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
