// CUDA kernels of the RK4 finite-volume stage for sm_100a.  Compiled twice (see kernels.h):
//   ma_fast   : -fmad=true  (production)
//   ma_strict : -fmad=false -DMA_STRICT (bit-for-bit with the reference's -DCELL_FLUX build)
//
// Two kernels per RK stage, one CTA per tile of cells (the flux kernel is launched once per pass when cut faces are
// shared between tiles, layout.h):
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
#include <algorithm>

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

using ma::interleaved_tile;

MA_DEV void load_state(const double *__restrict__ base, int stride, int c, double (&v)[5]) {
#pragma unroll
  for (int k = 0; k < 5; ++k) v[k] = __ldg(base + (size_t)k * stride + c);
}

// Face geometry as the kernels see it.  STRICT keeps the caller's tangent and binormal (the reference
// normalises and uses them, Roe_Flux.h:101-123); FAST needs only the area vector (see roe_flux_normal_only).
// Where the geometry of a tile's faces lives: component g of tile face e at base[g * cs + e]
// (STRICT: global SoA [12][n_tile_faces]; FAST: tile-blocked [6][faces rounded up to 16], see layout.h)
struct TileGeom {
  const double *base;
  size_t cs;
};
MA_DEV TileGeom tile_geom(const DevMesh &m, const TileInfoDev &T) {
#ifdef MA_STRICT
  return {m.face_geom + T.face_start, (size_t)m.n_tile_faces};
#else
  return {m.face_geom + (size_t)6 * T.face_start, (size_t)((T.face_count + 15) & ~15)};
#endif
}
#ifdef MA_STRICT
#define MA_GEOM_XF 9
struct FaceGeom {
  double n[3], t[3], b[3];
};
MA_DEV void load_face_geom(const TileGeom &tg, int e, FaceGeom &g) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    g.n[d] = __ldg(tg.base + (size_t)(0 + d) * tg.cs + e);
    g.t[d] = __ldg(tg.base + (size_t)(3 + d) * tg.cs + e);
    g.b[d] = __ldg(tg.base + (size_t)(6 + d) * tg.cs + e);
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
MA_DEV void load_face_geom(const TileGeom &tg, int e, FaceGeom &g) {
#pragma unroll
  for (int d = 0; d < 3; ++d) g.n[d] = __ldg(tg.base + (size_t)d * tg.cs + e);
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
                                                           int tile_begin, int n_first) {
  const TileInfoDev T = m.tiles[tile_begin + interleaved_tile(blockIdx.x, n_first, gridDim.x)];
  const TileGeom tg = tile_geom(m, T);
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
      const int e = (int)(sf & 0x3fffu);
      const int j = T.face_start + e;
      fj[s] = e;
      const int r = __ldg(m.face_right + j);
      double n[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) n[d] = __ldg(tg.base + (size_t)d * tg.cs + e);
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
            mn[k] = dmin(mn[k], Vn[k]);
            mx[k] = dmax(mx[k], Vn[k]);
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
      if (m.limiter == 1) {
        // VanAlbadaLimiter.h:45-65 in place of VenkatLimiter at StencilLimiter.h:455,459 (the alternative the reference
        // ships but never calls): phi = min over the six faces, from 1 (StencilLimiter.h:308-311, 345-346)
        double pva[5] = {1.0, 1.0, 1.0, 1.0, 1.0};
#pragma unroll
        for (int s = 0; s < 6; ++s) {
          const int e = fj[s];
          double disp[3];
#pragma unroll
          for (int d = 0; d < 3; ++d) disp[d] = __ldg(tg.base + (size_t)(MA_GEOM_XF + d) * tg.cs + e) - xc[d];
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            double dU = 0;
#pragma unroll
            for (int d = 0; d < 3; ++d) dU += disp[d] * g[k][d];
            pva[k] = fmin(pva[k], vanalbada_limit(mx[k] - V[k], mn[k] - V[k], dU));
          }
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) lim[(size_t)k * m.stride + c] = pva[k];
        continue;
      }
#pragma unroll
      for (int s = 0; s < 6; ++s) {
        const int e = fj[s];
        double disp[3];
        double dist = 0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          disp[d] = __ldg(tg.base + (size_t)(MA_GEOM_XF + d) * tg.cs + e) - xc[d];  // StencilLimiter.h:425-433
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
  const TileGeom tg = tile_geom(m, T);
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
    load_face_geom(tg, e, G);
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
          const double xf = __ldg(tg.base + (size_t)(MA_GEOM_XF + d) * tg.cs + e);
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
          xf[d] = __ldg(tg.base + (size_t)(MA_GEOM_XF + d) * tg.cs + e);
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
// FAST staged tile kernels: the sm_100a bulk-copy engine (cp.async.bulk, "TMA 1-D") moves every contiguous
// operand run of a tile — geometry of its faces, the SoA records of its own cells, the RK operands — into
// shared memory, the outside cells of its cut faces follow by 8-byte asynchronous gathers (LDGSTS), and ONE
// mbarrier collects both.  No thread holds a register or a scoreboard slot for data in flight, and the
// arithmetic afterwards reads shared memory only (32-bit addressing, immediate offsets).  Several CTAs per
// SM are resident, so one tile's copy phase runs behind another tile's arithmetic.
// ======================================================================================================
MA_DEV unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
MA_DEV void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
MA_DEV void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
MA_DEV void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// the calling thread's earlier cp.async copies arrive on the barrier when they land (counted in the init count)
MA_DEV void mbar_cp_async_arrive(unsigned bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
MA_DEV void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared bulk copy; 16-byte aligned on both sides, bytes a multiple of 16; completes on the mbarrier
MA_DEV void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
MA_DEV void bulk_prefetch_l2(const void *src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
#ifndef MA_HEADER_AHEAD
#define MA_HEADER_AHEAD 512
#endif
#ifndef MA_FLUX_PREFETCH_MIN_BYTES
#define MA_FLUX_PREFETCH_MIN_BYTES 0   // prefetch-ahead only runs of at least this many bytes
#endif
MA_DEV void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// The tile descriptor and outside-cell list of the tile a CTA `ahead` positions later will work on: pulled into L2
// now, so that CTA's first two loads are L2 hits instead of DRAM round trips.
MA_DEV void prefetch_tile_header(const DevMesh &m, int tile, int ntiles_end, int tid) {
  constexpr int ahead = MA_HEADER_AHEAD;  // about the number of CTAs resident on the device
  const int nxt = tile + ahead;
  if (nxt >= ntiles_end) return;
  if (tid == 0) prefetch_l2(m.tiles + nxt);
  if (tid * 32 < m.halo_stride) prefetch_l2(m.tile_halo + (size_t)nxt * m.halo_stride + tid * 32);
}
MA_DEV void cp_async8s(unsigned dst, const double *gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gmem_src) : "memory");
}

// Build-time knobs of the staged flux kernel (defaults = the measured best, profiles/r01e_variants.md;
// tools/build_variants.py sweeps them):
//   MA_C128_FT / MA_C128_FB   threads per CTA / resident CTAs per SM of the 128-cell tile class
//   MA_FLUX_RK_STAGED         1: volume, Un and the RK accumulator of the own cells arrive by bulk copy;
//                             0: phase 2 reads them from global memory (L2-prefetched at CTA start) — 11 KB less
//                                shared memory per CTA, which is what a fourth resident CTA needs
//   MA_FLUX_XC                1: the outside record of a thread's second cut face is staged in shared memory
//   MA_FLUX_PREFETCH_AHEAD    > 0: each CTA bulk-prefetches into L2 the operand runs of the tile that many CTAs ahead
#ifndef MA_C128_FT
#define MA_C128_FT 160
#endif
#ifndef MA_C128_FB
#define MA_C128_FB 4
#endif
#ifndef MA_FLUX_RK_STAGED
#define MA_FLUX_RK_STAGED 0
#endif
#ifndef MA_FLUX_XC
#define MA_FLUX_XC 0
#endif
#ifndef MA_FLUX_PREFETCH_AHEAD
#define MA_FLUX_PREFETCH_AHEAD 0
#endif
// Capacity class of a tile: the staged kernels are compiled for a few (cells, faces, cut faces) capacities so
// that every shared-memory stride is a compile-time constant.
template <int CELLS, int FACES, int HALO, int GRAD_T, int GRAD_B, int GRAD_B1, int FLUX_T, int FLUX_B>
struct TileCap {
  static constexpr int GRAD_MINB1 = GRAD_B1;  // resident CTAs per SM of the one-tile-per-CTA gradient kernel
  static constexpr int NC = CELLS;                          // own cells
  static constexpr int FC = (FACES + 15) / 16 * 16;         // tile faces (the copy length is rounded up to 16)
  static constexpr int HC = HALO;                           // cut faces == staged outside cells
  static constexpr int LS = (CELLS + 2 + HALO + 1) / 2 * 2; // staged cell list: alignment slack + own + outside
  static constexpr int RC = CELLS + 2;                      // staged RK operands
  static constexpr int SC = CELLS + 8;                      // staged slot map (uint16, 16-byte alignment slack)
  static constexpr int GRAD_THREADS = GRAD_T, GRAD_MINB = GRAD_B, FLUX_THREADS = FLUX_T, FLUX_MINB = FLUX_B;
  // cut faces beyond one per flux thread: their outside-cell records are staged in shared memory
  static constexpr int XC = (MA_FLUX_XC && HALO > FLUX_T) ? (HALO - FLUX_T + 1) / 2 * 2 : 0;
  static constexpr bool RK_STAGED = MA_FLUX_RK_STAGED != 0;
};
using Cap64 = TileCap<64, 240, 96, 64, 6, 8, 96, 6>;      // 4x4x4 bricks
#ifndef MA_C128_GB1
#define MA_C128_GB1 4
#endif
#ifndef MA_GRAD_STAGE_CELL
#define MA_GRAD_STAGE_CELL 1
#endif
using Cap128 = TileCap<128, 464, 160, 128, 3, MA_C128_GB1, MA_C128_FT, MA_C128_FB>;  // 4x4x8 / 8x4x4 bricks (flux: 160 threads = one per cut face, 464 faces in three rounds; 20 warps per SM at 94 registers, no spills)
#ifndef MA_C256_FB
#define MA_C256_FB 1
#endif
using Cap256 = TileCap<256, 896, 256, 256, 1, 2, 256, MA_C256_FB>;  // 8x8x4 / 4x8x8 bricks

// ---- sweep 1: Green-Gauss gradient + stencil min/max + Venkatakrishnan limiter ----------------------
// One own cell: neighbours (sV), face normals and centroids (sG) from shared memory; sn[s] = slot_face | slot_nbr << 16
// GLOBAL_G: the geometry is read from the tile's run in global memory (component stride gstride) instead of the staged copy
template <bool SECOND, int LS, bool GLOBAL_G, bool VANALBADA>
MA_DEV void grad_limiter_cell(const double *__restrict__ sG, const int gstride, const double *__restrict__ sV, int pos,
                              const unsigned (&sn)[6], double vol, const double (&xc)[3], int c, int stride,
                              double *__restrict__ grad, double *__restrict__ lim) {
  double V[5], g[5][3], mn[5], mx[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    V[k] = sV[k * LS + pos];
    g[k][0] = g[k][1] = g[k][2] = 0;
    mn[k] = mx[k] = V[k];  // min/max over {cell, face neighbours} (StencilLimiter.h:139-140, 272-273)
  }
#pragma unroll
  for (int s = 0; s < 6; ++s) {
    const int e = (int)(sn[s] & 0x3fffu);
    const bool right = (sn[s] & 0x8000u) != 0;  // the normal points out of elem1: flip it for elem2
    const unsigned nb = sn[s] >> 16;
    double an[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) an[d] = flip_sign_if(GLOBAL_G ? __ldg(sG + d * gstride + e) : sG[d * gstride + e], right);
    if (nb != 0xFFFFu) {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const double vn = sV[k * LS + nb];
        const double sum = V[k] + vn;  // 2 x GreenGauss.h:117; the 0.5/vol factor is applied after the loop
#pragma unroll
        for (int d = 0; d < 3; ++d) g[k][d] = fma(sum, an[d], g[k][d]);
        if (SECOND) {
          mn[k] = dmin(mn[k], vn);
          mx[k] = dmax(mx[k], vn);
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const double two_v = 2.0 * V[k];  // GreenGauss.h:186-216
#pragma unroll
        for (int d = 0; d < 3; ++d) g[k][d] = fma(two_v, an[d], g[k][d]);
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
        grad[(size_t)(k * 3 + d) * stride + c] = g[k][d];
      }
  }
  if (SECOND && VANALBADA) {
    // VanAlbadaLimiter.h:45-65 in place of VenkatLimiter at StencilLimiter.h:455,459 (the alternative the reference
    // ships but never calls): phi = min over the six faces, from 1 (StencilLimiter.h:308-311, 345-346)
    double pva[5] = {1.0, 1.0, 1.0, 1.0, 1.0};
#pragma unroll 1
    for (int s = 0; s < 6; ++s) {
      const int e = (int)(sn[s] & 0x3fffu);
      double disp[3];
#pragma unroll
      for (int d = 0; d < 3; ++d)
        disp[d] = (GLOBAL_G ? __ldg(sG + (3 + d) * gstride + e) : sG[(3 + d) * gstride + e]) - xc[d];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const double dU = fma(disp[2], g[k][2], fma(disp[1], g[k][1], disp[0] * g[k][0]));
        pva[k] = fmin(pva[k], vanalbada_limit(mx[k] - V[k], mn[k] - V[k], dU));
      }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) lim[(size_t)k * stride + c] = pva[k];
  } else if (SECOND) {
    double pN[5] = {1.0, 1.0, 1.0, 1.0, 1.0}, pD[5] = {1.0, 1.0, 1.0, 1.0, 1.0};
    double dumax[5], ndumin[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      dumax[k] = mx[k] - V[k];
      ndumin[k] = V[k] - mn[k];
    }
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const int e = (int)(sn[s] & 0x3fffu);
      double disp[3];
      double dist = 0;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        disp[d] = (GLOBAL_G ? __ldg(sG + (3 + d) * gstride + e) : sG[(3 + d) * gstride + e]) - xc[d];  // StencilLimiter.h:425-433
        dist = fma(disp[d], disp[d], dist);
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const double dU = fma(disp[2], g[k][2], fma(disp[1], g[k][1], disp[0] * g[k][0]));  // StencilLimiter.h:438-446
        // VenkatLimiter.h:45-73 with a = |du|, mm = |dumax| or |dumin| by the sign of du and the common factor
        // du cancelled: phi = (mm^2 + eps2 + 2 a mm) / (mm^2 + eps2 + a (2a + mm)); phi -> 1 as a -> 0
        const double aa = fabs(dU);
        const double mm = dU > 0.0 ? dumax[k] : ndumin[k];
        const double base = fma(mm, mm, dist);
        const double a2 = aa + aa;
        venkat_fraction_min(fma(a2, mm, base), fma(aa, a2 + mm, base), pN[k], pD[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) lim[(size_t)k * stride + c] = quot(pN[k], pD[k]);
  }
}

// Persistent, double-buffered: a CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  While it computes tile i
// out of one shared-memory stage, the bulk copies and the outside-cell gathers of tile i+1 land in the other
// stage, and the tile descriptor / outside-cell ids of tile i+2 are on their way to registers: no load on the
// arithmetic's critical path after the first tile.  Thread per own cell (blockDim >= cells of a tile).
// GDIRECT (persistent only; MINIAERO_GRAD_PERSIST=2): the face geometry is not staged — a stage is the 12 KB of cell
// states, so two stages fit under four resident CTAs — but read from the tile's run in global memory by the cell
// threads, after a bulk L2 prefetch issued one tile ahead.
template <bool SECOND, class CAP, bool PERSIST, bool GDIRECT = false, bool VANALBADA = false>
__global__ void __launch_bounds__(CAP::GRAD_THREADS, (PERSIST && !GDIRECT) ? CAP::GRAD_MINB : CAP::GRAD_MINB1)
    grad_limiter_tma_kernel(const DevMesh m, const double *__restrict__ V_, double *__restrict__ grad,
                            double *__restrict__ lim, int tile_begin, int ntiles, int n_first) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NG = GDIRECT ? 0 : (SECOND ? 6 : 3);  // staged geometry components: normal (+ centroid)
  constexpr int NGC = SECOND ? 6 : 3;                 // geometry components the arithmetic reads
  constexpr int FC = CAP::FC, LS = CAP::LS, RC = CAP::RC, SC = CAP::SC;
  // MA_GRAD_STAGE_CELL (default 1): the per-cell operands (slot maps, volume, centroid) arrive with the bulk copies
  // instead of by sixteen per-thread global loads: sC[1 + NXC][RC] (volume, centroid), sS[12][SC] (slot_face, slot_nbr;
  // 16-bit).  41 KB per CTA instead of 34 (still four CTAs per SM, the register file is the limit), 126 registers and
  // no spill instead of 128 with 24 bytes spilled; 6.47 -> 6.10 ms at 67 M cells (profiles/r02d_variants.md §6)
  constexpr bool SCELL = MA_GRAD_STAGE_CELL != 0;
  constexpr int NXC = SECOND ? 3 : 0;
  constexpr int CELL_D = SCELL ? (1 + NXC) * RC + (12 * SC * 2 + 7) / 8 : 0;
  constexpr int STAGE = NG * FC + 5 * LS + CELL_D;  // doubles per stage: sG[NG][FC], sV[5][LS] (, sC, sS)
  constexpr int HPT = (CAP::HC + CAP::GRAD_THREADS - 1) / CAP::GRAD_THREADS;  // outside cells per thread
  double *sbase = reinterpret_cast<double *>(smem_raw) + 2;  // two mbarriers, then the stages
  const unsigned bar0 = smem_addr(smem_raw);
  const int tid = threadIdx.x;
  const int G = gridDim.x;
  int t = blockIdx.x;
  if (t >= ntiles) return;
  const TileInfoDev *tiles = m.tiles + tile_begin;
  if (tid == 0) {
    mbar_init(bar0, blockDim.x + 1);
    mbar_init(bar0 + 8, blockDim.x + 1);
    mbar_fence_init();
  }
  int ids0[HPT], ids1[HPT];
  // tile descriptors: current, next, the one after
  // (the CTA's tile number t walks the group pair by pair, see interleaved_tile)
  auto tix = [&](int tile) { return interleaved_tile(tile, n_first, ntiles); };
  TileInfoDev T0 = tiles[tix(t)], T1 = T0, T2 = T0;
  if (t + G < ntiles) T1 = tiles[tix(t + G)];
  if (t + 2 * G < ntiles) T2 = tiles[tix(t + 2 * G)];
  // outside cells of this thread's cut faces of the CTA's tile number `tile` (-1: none); the address needs only
  // the tile index, so the loads fly together with the tile descriptor's
  auto load_ids = [&](int tile, int (&ids)[HPT]) {
    const int *p = m.tile_halo + (size_t)(tile_begin + tix(tile)) * m.halo_stride;
#pragma unroll
    for (int j = 0; j < HPT; ++j) {
      const int h = tid + j * CAP::GRAD_THREADS;
      ids[j] = h < m.halo_stride ? __ldg(p + h) : -1;
    }
  };
  // start every copy of tile T into stage st; ids = outside cells of this thread
  auto issue = [&](const TileInfoDev &T, int st, const int (&ids)[HPT]) {
    double *sG = sbase + st * STAGE, *sV = sG + NG * FC;
    const unsigned bar = bar0 + 8 * st;
    const int shift = T.cell_start & 1;
    const int hb = (shift + T.cell_count + 1) & ~1;  // doubles per staged own-cell run == first outside-cell position
    const unsigned fcp = (unsigned)(T.face_count + 15) & ~15u;
    if (tid < 32) {
      const unsigned gbytes = fcp * 8u, vbytes = (unsigned)hb * 8u;
      const int ssh = T.cell_start & 7;
      const unsigned sbytes = (unsigned)((ssh + T.cell_count + 7) & ~7) * 2u;
      if (tid == 0) mbar_arrive_expect_tx(bar, NG * gbytes + 5 * vbytes + (SCELL ? (1 + NXC) * vbytes + 12 * sbytes : 0u));
      __syncwarp();
      // generic-proxy reads of this stage (ordered before by the CTA barrier) precede the copy engine's writes
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (GDIRECT && tid == 31)  // the geometry run of this tile: into L2 now, read by the cell threads one tile later
        bulk_prefetch_l2(m.face_geom + (size_t)6 * T.face_start, (unsigned)NGC * gbytes);
      if (tid < NG)
        bulk_g2s(smem_addr(sG + tid * FC), m.face_geom + (size_t)6 * T.face_start + (size_t)tid * fcp, gbytes, bar);
      else if (tid < NG + 5)
        bulk_g2s(smem_addr(sV + (tid - NG) * LS), V_ + (size_t)(tid - NG) * m.stride + (T.cell_start - shift), vbytes, bar);
      else if (SCELL && tid < NG + 5 + 1 + NXC) {
        const int j = tid - (NG + 5);
        const double *base = j == 0 ? m.cell_vol : m.cell_xyz + (size_t)(j - 1) * m.stride;
        bulk_g2s(smem_addr(sV + 5 * LS + j * RC), base + (T.cell_start - shift), vbytes, bar);
      } else if (SCELL && tid < NG + 5 + 1 + NXC + 12) {
        const int q = tid - (NG + 5 + 1 + NXC);
        const unsigned short *base = q < 6 ? m.slot_face + (size_t)q * m.slot_stride : m.slot_nbr + (size_t)(q - 6) * m.slot_stride;
        unsigned short *sS = reinterpret_cast<unsigned short *>(sV + 5 * LS + (1 + NXC) * RC);
        bulk_g2s(smem_addr(sS + q * SC), base + (T.cell_start - ssh), sbytes, bar);
      }
    }
#pragma unroll
    for (int j = 0; j < HPT; ++j) {
      if (ids[j] >= 0) {
        const int h = tid + j * CAP::GRAD_THREADS;
#pragma unroll
        for (int k = 0; k < 5; ++k) cp_async8s(smem_addr(sV + k * LS + hb + h), V_ + (size_t)k * m.stride + ids[j]);
      }
    }
    mbar_cp_async_arrive(bar);
  };
  // per-cell operands that only this thread needs: registers, loaded one tile ahead
  struct CellRegs {
    unsigned sn[6];  // slot_face | slot_nbr << 16
    double vol, xc[3];
  };
  auto load_cell = [&](const TileInfoDev &T, CellRegs &r) {
    if (SCELL) return;  // read from the stage after the wait
    const int c = T.cell_start + (tid < T.cell_count ? tid : 0);
#pragma unroll
    for (int s = 0; s < 6; ++s)
      r.sn[s] = (unsigned)__ldg(m.slot_face + (size_t)s * m.slot_stride + c) |
                ((unsigned)__ldg(m.slot_nbr + (size_t)s * m.slot_stride + c) << 16);
    r.vol = __ldg(m.cell_vol + c);
#pragma unroll
    for (int d = 0; d < 3; ++d) r.xc[d] = SECOND ? __ldg(m.cell_xyz + (size_t)d * m.stride + c) : 0.0;
  };

  load_ids(t, ids0);
  if (t + G < ntiles) load_ids(t + G, ids1);
  __syncthreads();  // barriers initialised
  issue(T0, 0, ids0);
  CellRegs cur, nxt;
  load_cell(T0, cur);
  nxt = cur;
  for (int i = 0;; ++i) {
    const int st = i & 1;
    const bool has1 = t + G < ntiles, has2 = t + 2 * G < ntiles;
    TileInfoDev T3 = T2;
    if (has1) {
      issue(T1, st ^ 1, ids1);
      load_cell(T1, nxt);
      if (has2) load_ids(t + 2 * G, ids1);                 // consumed at the top of the next iteration
      if (t + 3 * G < ntiles) T3 = tiles[tix(t + 3 * G)];  // consumed two iterations from now
    }
    mbar_wait(bar0 + 8 * st, (unsigned)(i >> 1) & 1u);
    if (SCELL && tid < T0.cell_count) {
      const double *sC = sbase + st * STAGE + NG * FC + 5 * LS;
      const unsigned short *sS = reinterpret_cast<const unsigned short *>(sC + (1 + NXC) * RC);
      const int p = (T0.cell_start & 1) + tid, q = (T0.cell_start & 7) + tid;
#pragma unroll
      for (int s = 0; s < 6; ++s) cur.sn[s] = (unsigned)sS[s * SC + q] | ((unsigned)sS[(6 + s) * SC + q] << 16);
      cur.vol = sC[p];
#pragma unroll
      for (int d = 0; d < 3; ++d) cur.xc[d] = SECOND ? sC[(1 + d) * RC + p] : 0.0;
    }
    if (tid < T0.cell_count) {
      const double *sG = sbase + st * STAGE, *sV = sG + NG * FC;
      if (GDIRECT)
        grad_limiter_cell<SECOND, LS, true, VANALBADA>(m.face_geom + (size_t)6 * T0.face_start, (int)((unsigned)(T0.face_count + 15) & ~15u),
                                            sV, (T0.cell_start & 1) + tid, cur.sn, cur.vol, cur.xc, T0.cell_start + tid,
                                            m.stride, grad, lim);
      else
        grad_limiter_cell<SECOND, LS, false, VANALBADA>(sG, FC, sV, (T0.cell_start & 1) + tid, cur.sn, cur.vol, cur.xc,
                                                        T0.cell_start + tid, m.stride, grad, lim);
    }
    if (!has1) break;
    __syncthreads();  // every read of stage st is done before tile i+2 is copied into it
    T0 = T1, T1 = T2, T2 = T3;
    cur = nxt;
    t += G;
  }
}

// ---- sweep 2: face fluxes + slot-ordered gather + RK stage update -----------------------------------
// Shared memory: sRec[NREC][RC] cell records (V 5, gradient 15, limiter 5, centroid 3) of the tile's own cells;
// sG[6][FC] face geometry, overwritten face by face with the flux; sLR[FC] tile-local face connectivity;
// sSlot[6][SC] slot map; optionally (MA_FLUX_RK_STAGED) sRK[11][RC] volume, Un, Acc of the own cells — by default
// those are read from global memory in phase 2.  All of it arrives by bulk copies.
// The record of the outside cell of a cut face never touches shared memory: the thread that will evaluate that
// face gathers it straight into registers before it waits for the copies, so the gather's latency hides behind
// the copies and the per-CTA footprint (54 KB) stays small enough for four resident CTAs per SM.
//   phase 1  thread per tile face (cut faces first): limited extrapolation of both cell records to the face
//            (Flux.h:109-132), Roe flux (+ viscous flux), or the boundary-condition flux
//   phase 2  thread per own cell: gather of the six face fluxes in slot order (Flux.h:216-227), RK update
template <bool SECOND, bool VISCOUS>
struct FluxRec {
  static constexpr bool GRAD = SECOND || VISCOUS;
  static constexpr int R_G = 5, R_L = R_G + (GRAD ? 15 : 0), R_X = R_L + (SECOND ? 5 : 0), NREC = R_X + (SECOND ? 3 : 0);
  static constexpr int NGEOM = SECOND ? 6 : 3;  // copied geometry components
  static constexpr int NGS = SECOND ? 6 : 5;    // staged columns (the flux needs 5)
};
template <bool SECOND, bool VISCOUS, class CAP>
constexpr size_t flux_tma_smem() {
  using R = FluxRec<SECOND, VISCOUS>;
  return (size_t)(R::NREC * CAP::RC + R::NGS * CAP::FC + (CAP::RK_STAGED ? 11 : 0) * CAP::RC + R::NREC * CAP::XC) * 8 + (size_t)CAP::FC * 4 +
         (size_t)6 * CAP::SC * 2 + 16;
}

// a cell record in shared memory (component stride RC) or in registers
template <bool SECOND, bool VISCOUS, int RC>
struct SmemRecord {
  using R = FluxRec<SECOND, VISCOUS>;
  const double *p;  // sRec + position
  MA_DEV double v(int k) const { return p[k * RC]; }
  MA_DEV double g(int k, int d) const { return p[(R::R_G + k * 3 + d) * RC]; }
  MA_DEV double lim(int k) const { return p[(R::R_L + k) * RC]; }
  MA_DEV double x(int d) const { return p[(R::R_X + d) * RC]; }
};
template <bool SECOND, bool VISCOUS>
struct RegRecord {
  using R = FluxRec<SECOND, VISCOUS>;
  const double (&r)[R::NREC];
  MA_DEV double v(int k) const { return r[k]; }
  MA_DEV double g(int k, int d) const { return r[R::R_G + k * 3 + d]; }
  MA_DEV double lim(int k) const { return r[R::R_L + k]; }
  MA_DEV double x(int d) const { return r[R::R_X + d]; }
};
// one side of an interior face: primitives limited-extrapolated to the face centroid (Flux.h:109-132) and this
// cell's share of the viscous traction.  Viscous_Flux.h:65-98 is linear in the face gradient 0.5 (g_l + g_r)
// (Flux.h:146-149), so each side contributes q = tau(g) . a (momentum rows) and hh = grad T . a on its own: four
// running sums instead of the twelve gradient sums (the density gradient is never used).
template <bool SECOND, bool VISCOUS, bool FIRST, class REC>
MA_DEV void face_side(const REC &rec, const double (&xf)[3], const double (&n)[3], double (&V)[5], double (&q)[4]) {
  double dx[3];
  if (SECOND) {
#pragma unroll
    for (int d = 0; d < 3; ++d) dx[d] = xf[d] - rec.x(d);
  }
  double gv[5][3];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    if (SECOND || (VISCOUS && k > 0)) {
#pragma unroll
      for (int d = 0; d < 3; ++d) gv[k][d] = rec.g(k, d);
    }
    if (SECOND) {
      const double t = fma(dx[2], gv[k][2], fma(dx[1], gv[k][1], dx[0] * gv[k][0]));  // Flux.h:114-121
      V[k] = fma(t, rec.lim(k), rec.v(k));                                         // Flux.h:124-127
    } else {
      V[k] = rec.v(k);
    }
  }
  if (VISCOUS) {
    const double third_div = (gv[1][0] + gv[2][1] + gv[3][2]) * (1.0 / 3.0);
    const double txx = gv[1][0] - third_div, tyy = gv[2][1] - third_div, tzz = gv[3][2] - third_div;
    const double txy = 0.5 * (gv[1][1] + gv[2][0]), txz = 0.5 * (gv[1][2] + gv[3][0]);
    const double tyz = 0.5 * (gv[2][2] + gv[3][1]);
    const double q0 = txx * n[0] + txy * n[1] + txz * n[2];
    const double q1 = txy * n[0] + tyy * n[1] + tyz * n[2];
    const double q2 = txz * n[0] + tyz * n[1] + tzz * n[2];
    const double hh = gv[4][0] * n[0] + gv[4][1] * n[1] + gv[4][2] * n[2];
    q[0] = FIRST ? q0 : q[0] + q0;
    q[1] = FIRST ? q1 : q[1] + q1;
    q[2] = FIRST ? q2 : q[2] + q2;
    q[3] = FIRST ? hh : q[3] + hh;
  }
}
// Roe flux (Roe_Flux.h:49-265) minus the Newtonian viscous flux (Viscous_Flux.h:65-98) at the face state
// 0.5 (Vl + Vr) (Flux.h:142-143); q = (tau . a, grad T . a) of TWICE the face gradient (both sides' shares summed)
MA_DEV void subtract_viscous_flux(const double (&Vl)[5], const double (&Vr)[5], const double (&q)[4], double (&flux)[5]) {
  const double mu = compute_viscosity(0.5 * (Vl[4] + Vr[4]));
  const double uq = (Vl[1] + Vr[1]) * q[0] + (Vl[2] + Vr[2]) * q[1] + (Vl[3] + Vr[3]) * q[2];
  flux[1] -= mu * q[0];
  flux[2] -= mu * q[1];
  flux[3] -= mu * q[2];
  flux[4] -= fma(0.5 * mu, uq, 0.5 * compute_thermal_conductivity(mu) * q[3]);
}
template <bool VISCOUS>
MA_DEV void interior_flux(const double (&Vl)[5], const double (&Vr)[5], const double (&q)[4], const FaceGeom &G,
                          double (&flux)[5]) {
  face_roe_flux(Vl, Vr, G, flux);
  if (VISCOUS) subtract_viscous_flux(Vl, Vr, q, flux);
}

template <bool SECOND, bool VISCOUS, class CAP>
__global__ void __launch_bounds__(CAP::FLUX_THREADS, CAP::FLUX_MINB)
    flux_rk_tma_kernel(const DevMesh m, const StageArgs a, int tile_begin) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using R = FluxRec<SECOND, VISCOUS>;
  constexpr int FC = CAP::FC, RC = CAP::RC, SC = CAP::SC, XC = CAP::XC;
  constexpr int R_G = R::R_G, R_L = R::R_L, R_X = R::R_X, NREC = R::NREC, NGEOM = R::NGEOM;
  using SRec = SmemRecord<SECOND, VISCOUS, RC>;
  using XRec = SmemRecord<SECOND, VISCOUS, (XC ? XC : 1)>;
  using RRec = RegRecord<SECOND, VISCOUS>;
  double *sRec = reinterpret_cast<double *>(smem_raw);                   // [NREC][RC]
  double *sG = sRec + NREC * RC;                                         // [NGS][FC]
  double *sRK = sG + R::NGS * FC;                                        // [11][RC]
  constexpr bool RKS = CAP::RK_STAGED;
  double *sOut = sRK + (RKS ? 11 : 0) * RC;                              // [NREC][XC] outside cells of cut faces >= blockDim
  unsigned *sLR = reinterpret_cast<unsigned *>(sOut + NREC * XC);        // [FC]
  unsigned short *sSlot = reinterpret_cast<unsigned short *>(sLR + FC);  // [6][SC]
  const unsigned bar = smem_addr(sSlot + 6 * SC);
  const int tid = threadIdx.x;
  // outside cell of this thread's cut face (-1: none): its address needs only the tile index, so the load flies
  // together with the tile descriptor's
  const int *halo_ids = m.tile_halo + (size_t)(tile_begin + blockIdx.x) * m.halo_stride;
  const int my_outside = tid < m.halo_stride ? __ldg(halo_ids + tid) : -1;
  // shared cut faces: where this thread's cut face publishes its flux for the tile on the other side (-1: nowhere)
  const int *pub_ids = m.tile_pub ? m.tile_pub + (size_t)(tile_begin + blockIdx.x) * m.halo_stride : nullptr;
  const int my_pub = (pub_ids && tid < m.halo_stride) ? __ldg(pub_ids + tid) : -1;
  const TileInfoDev T = m.tiles[tile_begin + blockIdx.x];
  prefetch_tile_header(m, tile_begin + blockIdx.x, tile_begin + gridDim.x, tid);
  // faces [0, nf) are evaluated here (closed / boundary, then nh cut faces); [nf, T.face_count) arrive as fluxes
  const int nc = T.cell_count, nf = T.n_eval, nimp = T.face_count - T.n_eval;
  const int shift = T.cell_start & 1;
  const int hb = (shift + nc + 1) & ~1;  // doubles per staged own-cell run; positions >= hb are outside cells
  const int nh = nf - T.cut_start;
  const unsigned fcp = (unsigned)(T.face_count + 15) & ~15u;
  const int sshift = T.cell_start & 7;
  // outside cell of this thread's second cut face (tiles with more cut faces than the CTA has threads)
  // (staged in shared memory when XC > 0; otherwise only its lines are pulled into L2 now and the record is gathered
  // when the thread gets to that face)
  constexpr int CUT2 = CAP::HC > CAP::FLUX_THREADS ? CAP::HC - CAP::FLUX_THREADS : 0;
  const int my_outside2 =
      (CUT2 > 0 && tid < CUT2 && CAP::FLUX_THREADS + tid < m.halo_stride) ? __ldg(halo_ids + CAP::FLUX_THREADS + tid) : -1;
  if (tid == 0) {
    mbar_init(bar, XC > 0 ? blockDim.x + 1 : 1);
    mbar_fence_init();
  }
  __syncthreads();

  // ---- copy phase
  // every contiguous operand run of tile TT: emit(shared-memory destination, global source, bytes)
  auto for_each_run = [&](const TileInfoDev &TT, auto &&emit) {
    const int sh = TT.cell_start & 1;
    const int hbb = (sh + TT.cell_count + 1) & ~1;
    const unsigned fcq = (unsigned)(TT.face_count + 15) & ~15u;
    const int ssh = TT.cell_start & 7;
    // geometry of the evaluated faces only: with imports n_eval is even (layout.h), the imported flux columns start there
    const unsigned vbytes = (unsigned)hbb * 8u, gbytes = (TT.imp_area >= 0 ? (unsigned)TT.n_eval : fcq) * 8u, lbytes = fcq * 4u;
    const unsigned sbytes = (unsigned)((ssh + TT.cell_count + 7) & ~7) * 2u;
    const unsigned ibytes = (unsigned)((TT.face_count - TT.n_eval + 1) & ~1) * 8u;
    const size_t c0 = (size_t)(TT.cell_start - sh);
    constexpr int NCOPY = NREC + 11 + NGEOM + 1 + 6 + 5;
    for (int i = tid; i < NCOPY; i += 32) {
      if (i < NREC) {
        const double *base = i < R_G   ? a.V + (size_t)i * m.stride
                             : i < R_L ? a.grad + (size_t)(i - R_G) * m.stride
                             : i < R_X ? a.lim + (size_t)(i - R_L) * m.stride
                                       : m.cell_xyz + (size_t)(i - R_X) * m.stride;
        emit(smem_addr(sRec + i * RC), base + c0, vbytes, true);
      } else if (i < NREC + 11) {
        const int j = i - NREC;
        const double *base = j == 0 ? m.cell_vol : j < 6 ? a.Un + (size_t)(j - 1) * m.stride : a.Acc + (size_t)(j - 6) * m.stride;
        if (j == 0 || (j < 6 ? a.kind != 2 : a.kind != 0)) emit(smem_addr(sRK + j * RC), base + c0, vbytes, RKS);
      } else if (i < NREC + 11 + NGEOM) {
        const int gi = i - NREC - 11;
        emit(smem_addr(sG + gi * FC), m.face_geom + (size_t)6 * TT.face_start + (size_t)gi * fcq, gbytes, true);
      } else if (i == NREC + 11 + NGEOM) {
        emit(smem_addr(sLR), m.face_lr + TT.face_start, lbytes, true);
      } else if (i < NREC + 11 + NGEOM + 1 + 6) {
        const int s = i - (NREC + 11 + NGEOM + 1);
        emit(smem_addr(sSlot + s * SC), m.slot_face + (size_t)s * m.slot_stride + (TT.cell_start - ssh), sbytes, true);
      } else if (TT.imp_area >= 0) {  // fluxes of the imported cut faces, published by the tiles of the earlier launch
        const int k = i - (NREC + 11 + NGEOM + 1 + 6);
        emit(smem_addr(sG + k * FC + TT.n_eval), m.cut_flux + ((size_t)TT.imp_area * 5 + k) * m.import_capacity, ibytes, true);
      }
    }
  };
  if (tid < 32) {
    const unsigned vbytes = (unsigned)hb * 8u, gbytes = (T.imp_area >= 0 ? (unsigned)nf : fcp) * 8u, lbytes = fcp * 4u;
    const unsigned sbytes = (unsigned)((sshift + nc + 7) & ~7) * 2u;
    const unsigned ibytes = T.imp_area >= 0 ? (unsigned)((nimp + 1) & ~1) * 8u : 0u;
    const int nrk = 1 + (a.kind != 2 ? 5 : 0) + (a.kind != 0 ? 5 : 0);
    if (tid == 0)
      mbar_arrive_expect_tx(bar, (NREC + (RKS ? nrk : 0)) * vbytes + NGEOM * gbytes + lbytes + 6 * sbytes + 5 * ibytes);
    __syncwarp();
    for_each_run(T, [&](unsigned dst, const void *src, unsigned bytes, bool staged) {
      if (staged)
        bulk_g2s(dst, src, bytes, bar);
      else  // phase 2 reads this run from global memory: have it in L2 by then
        bulk_prefetch_l2(src, bytes);
    });
  }
  // the outside-cell record of this thread's cut face: registers, in flight together with the copies
  double orec[NREC];
  auto gather_outside = [&](int c) {
#pragma unroll
    for (int k = 0; k < 5; ++k) orec[k] = __ldg(a.V + (size_t)k * m.stride + c);
    if (R::GRAD) {
#pragma unroll
      for (int k = 0; k < 15; ++k) orec[R_G + k] = __ldg(a.grad + (size_t)k * m.stride + c);
    }
    if (SECOND) {
#pragma unroll
      for (int k = 0; k < 5; ++k) orec[R_L + k] = __ldg(a.lim + (size_t)k * m.stride + c);
#pragma unroll
      for (int d = 0; d < 3; ++d) orec[R_X + d] = __ldg(m.cell_xyz + (size_t)d * m.stride + c);
    }
  };
  if (XC > 0) {  // 8-byte asynchronous gathers of the second outside record into shared memory, on the same barrier
    if (my_outside2 >= 0) {
      const unsigned dst = smem_addr(sOut + tid);
      const int c = my_outside2;
#pragma unroll
      for (int k = 0; k < 5; ++k) cp_async8s(dst + (unsigned)(k * XC) * 8u, a.V + (size_t)k * m.stride + c);
      if (R::GRAD) {
#pragma unroll
        for (int k = 0; k < 15; ++k) cp_async8s(dst + (unsigned)((R_G + k) * XC) * 8u, a.grad + (size_t)k * m.stride + c);
      }
      if (SECOND) {
#pragma unroll
        for (int k = 0; k < 5; ++k) cp_async8s(dst + (unsigned)((R_L + k) * XC) * 8u, a.lim + (size_t)k * m.stride + c);
#pragma unroll
        for (int d = 0; d < 3; ++d) cp_async8s(dst + (unsigned)((R_X + d) * XC) * 8u, m.cell_xyz + (size_t)d * m.stride + c);
      }
    }
    mbar_cp_async_arrive(bar);
  }
  if (my_outside >= 0 && tid < nh) gather_outside(my_outside);  // (imported cut faces have an outside cell too: not needed here)
  if (XC == 0 && CUT2 > 0 && my_outside2 >= 0) {
    const int c = my_outside2;
#pragma unroll
    for (int k = 0; k < 5; ++k) prefetch_l2(a.V + (size_t)k * m.stride + c);
    if (R::GRAD) {
#pragma unroll
      for (int k = 0; k < 15; ++k) prefetch_l2(a.grad + (size_t)k * m.stride + c);
    }
    if (SECOND) {
#pragma unroll
      for (int k = 0; k < 5; ++k) prefetch_l2(a.lim + (size_t)k * m.stride + c);
#pragma unroll
      for (int d = 0; d < 3; ++d) prefetch_l2(m.cell_xyz + (size_t)d * m.stride + c);
    }
  }
  if (MA_FLUX_PREFETCH_AHEAD > 0 && tid < 32) {
    // the operand runs of the tile a CTA MA_FLUX_PREFETCH_AHEAD positions later will copy: into L2 now (its descriptor
    // is an L2 hit, prefetch_tile_header), so that CTA's bulk copies are served from L2
    const int nxt = (int)blockIdx.x + MA_FLUX_PREFETCH_AHEAD;
    if (nxt < (int)gridDim.x) {
      const TileInfoDev P = m.tiles[tile_begin + nxt];
      for_each_run(P, [&](unsigned, const void *src, unsigned bytes, bool) {
        if (bytes >= MA_FLUX_PREFETCH_MIN_BYTES) bulk_prefetch_l2(src, bytes);
      });
    }
  }
  mbar_wait(bar, 0);

  // ---- phase 1: one flux per tile face.  Work item w: cut face cut_start + w for w < nh, closed / boundary face
  // w - nh otherwise; thread t takes w = t, t + blockDim, ...
  // a cut face: `outside` is the record of the cell on the other side of the tile boundary
  auto cut_face = [&](int e, auto outside, int pub) {
    const unsigned lr = sLR[e];
    const int pl = (int)(lr & 0xffffu), pr = (int)(lr >> 16);
    FaceGeom G;
    double xf[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      G.n[d] = sG[d * FC + e];
      xf[d] = SECOND ? sG[(3 + d) * FC + e] : 0.0;
    }
    double Vl[5], Vr[5], gs[4], flux[5];
    // the outside record first: its registers are free before the own cell's record is read
    if (pr >= hb) {  // outside cell on the right
      face_side<SECOND, VISCOUS, true>(outside, xf, G.n, Vr, gs);
      face_side<SECOND, VISCOUS, false>(SRec{sRec + pl}, xf, G.n, Vl, gs);
    } else {
      face_side<SECOND, VISCOUS, true>(outside, xf, G.n, Vl, gs);
      face_side<SECOND, VISCOUS, false>(SRec{sRec + pr}, xf, G.n, Vr, gs);
    }
    interior_flux<VISCOUS>(Vl, Vr, gs, G, flux);
#pragma unroll
    for (int k = 0; k < 5; ++k) sG[k * FC + e] = flux[k];  // this thread's own column: geometry is dead
    if (pub >= 0) {  // shared cut faces: the tile on the other side runs in the next launch and imports this flux
#pragma unroll
      for (int k = 0; k < 5; ++k) m.cut_flux[(size_t)pub + (size_t)k * m.import_capacity] = flux[k];
    }
  };
  int w = tid;
  if (w < nh) {  // first cut face of this thread: outside record in registers
    cut_face(T.cut_start + w, RRec{orec}, my_pub);
    w += blockDim.x;
  }
  for (; w < nh; w += blockDim.x) {  // further cut faces: outside record staged in shared memory, else gathered now
    const int pub = pub_ids ? __ldg(pub_ids + w) : -1;
    if (XC > 0 && blockDim.x == CAP::FLUX_THREADS && w - (int)blockDim.x < XC) {
      cut_face(T.cut_start + w, XRec{sOut + (w - (int)blockDim.x)}, pub);
    } else {
      gather_outside(w == tid + CAP::FLUX_THREADS && my_outside2 >= 0 ? my_outside2 : __ldg(halo_ids + w));
      cut_face(T.cut_start + w, RRec{orec}, pub);
    }
  }
  for (; w < nf; w += blockDim.x) {  // closed and boundary faces
    const int e = w - nh;
    const unsigned lr = sLR[e];
    const int pl = (int)(lr & 0xffffu);
    const unsigned pr = lr >> 16;
    FaceGeom G;
#pragma unroll
    for (int d = 0; d < 3; ++d) G.n[d] = sG[d * FC + e];
    double flux[5];
    if (pr < 0xFFF0u) {
      // interior face: Flux.h:89-160
      double xf[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) xf[d] = SECOND ? sG[(3 + d) * FC + e] : 0.0;
      double Vl[5], Vr[5], gs[4];
      face_side<SECOND, VISCOUS, true>(SRec{sRec + pl}, xf, G.n, Vl, gs);
      face_side<SECOND, VISCOUS, false>(SRec{sRec + (int)pr}, xf, G.n, Vr, gs);
      interior_flux<VISCOUS>(Vl, Vr, gs, G, flux);
    } else {
      // boundary face, always first order (Extrapolate_BC.h, Tangent_BC.h, Inflow_BC.h, NoSlip_BC.h)
      const int type = (int)(0xFFFFu - pr);
      double Vl[5], Vr[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) Vl[k] = sRec[k * RC + pl];
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
        const int cg = T.cell_start + (pl - shift);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          xf[d] = SECOND ? sG[(3 + d) * FC + e]
                         : __ldg(m.face_geom + (size_t)6 * T.face_start + (size_t)(3 + d) * fcp + e);
          xc[d] = SECOND ? sRec[(R_X + d) * RC + pl] : __ldg(m.cell_xyz + (size_t)d * m.stride + cg);
        }
        noslip_viscous_flux(Vl, G.n, area_norm, xf, xc, vflux);
#pragma unroll
        for (int k = 0; k < 5; ++k) flux[k] -= vflux[k];  // slot = -iflux + vflux == -(iflux - vflux)
      }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) sG[k * FC + e] = flux[k];
  }
  // RK operands of this thread's (first) cell when they are not staged: requested before the barrier — the face
  // registers are dead here — so that the loads fly while the CTA's slower warps finish their faces
  double pre_vol = 1.0, pre_un[5] = {0, 0, 0, 0, 0}, pre_acc[5] = {0, 0, 0, 0, 0};
  if (!RKS && tid < nc) {
    const int c = T.cell_start + tid;
    pre_vol = __ldg(m.cell_vol + c);
    if (a.kind != 2) {
#pragma unroll
      for (int k = 0; k < 5; ++k) pre_un[k] = __ldg(a.Un + (size_t)k * m.stride + c);
    }
    if (a.kind != 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k) pre_acc[k] = a.Acc[(size_t)k * m.stride + c];
    }
  }
  __syncthreads();

  // ---- phase 2: slot-ordered gather, residual, RK update; the next stage state is stored as primitives
  for (int lc = tid; lc < nc; lc += blockDim.x) {
    const int c = T.cell_start + lc;
    const int p = shift + lc;
    const bool pre = lc == tid;  // first cell of the thread: operands already requested
    const double dtv = a.dt * rcp(RKS ? sRK[p] : pre ? pre_vol : __ldg(m.cell_vol + c));
    double Rs[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const unsigned sf = sSlot[s * SC + sshift + lc];
      const int e = (int)(sf & 0x3fffu);
      const double sg = (sf & 0x8000u) ? dtv : -dtv;  // Flux.h:172-178: left slot holds -flux, right slot +flux
#pragma unroll
      for (int k = 0; k < 5; ++k) Rs[k] = fma(sg, sG[k * FC + e], Rs[k]);
    }
    double Wn[5];  // conservative state the next stage is evaluated at (or the new solution)
    if (a.kind == 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const double w0 = RKS ? sRK[(1 + k) * RC + p] : pre ? pre_un[k] : __ldg(a.Un + (size_t)k * m.stride + c);
        a.Acc[(size_t)k * m.stride + c] = fma(a.beta, Rs[k], w0);
        Wn[k] = fma(a.alpha_next, Rs[k], w0);
      }
    } else if (a.kind == 1) {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const double acc = RKS ? sRK[(6 + k) * RC + p] : pre ? pre_acc[k] : a.Acc[(size_t)k * m.stride + c];
        const double un = RKS ? sRK[(1 + k) * RC + p] : pre ? pre_un[k] : __ldg(a.Un + (size_t)k * m.stride + c);
        a.Acc[(size_t)k * m.stride + c] = fma(a.beta, Rs[k], acc);
        Wn[k] = fma(a.alpha_next, Rs[k], un);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        Wn[k] = fma(a.beta, Rs[k], RKS ? sRK[(6 + k) * RC + p] : pre ? pre_acc[k] : a.Acc[(size_t)k * m.stride + c]);
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

// caller order AoS [n_owned][5] conservative state -> Un (SoA, renumbered) and its primitives V, one thread per cell:
// ma_solver_set_solution / ma_solver_submit (the 40 contiguous bytes of a cell are read by one thread; cells of a
// z-line are neighbours in both numberings, so the five component stores of a warp fall into few sectors)
__global__ void set_state_kernel(const DevMesh m, const double *__restrict__ aos, const int *__restrict__ old2new,
                                 double *__restrict__ Un, double *__restrict__ V) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m.n_owned) return;
  double U[5], P[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) U[k] = aos[(size_t)5 * c + k];
  compute_primitives(U, P);
  const int n = old2new[c];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    Un[(size_t)k * m.stride + n] = U[k];
    V[(size_t)k * m.stride + n] = P[k];
  }
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
#ifdef MA_STRICT
  viscous_flux(G, V, A, F);
#else
  // the production path of the FAST face loop (not the reference-shaped viscous_flux, which FAST only uses for the
  // no-slip wall): the face gradient G = 0.5 (G + G) enters as the two sides' shares of tau.a / gradT.a (face_side),
  // the face state V = 0.5 (V + V) as the two side states, and subtract_viscous_flux turns them into the flux
  double rec[20];
  for (int k = 0; k < 5; ++k) {
    rec[k] = V[k];
    for (int d = 0; d < 3; ++d) rec[5 + 3 * k + d] = G[k][d];
  }
  const double xf[3] = {0.0, 0.0, 0.0};
  double Vs[5], q[4];
  face_side<false, true, true>(RegRecord<false, true>{rec}, xf, A, Vs, q);
  face_side<false, true, false>(RegRecord<false, true>{rec}, xf, A, Vs, q);
  for (int k = 0; k < 5; ++k) F[k] = 0.0;
  subtract_viscous_flux(Vs, Vs, q, F);
  for (int k = 0; k < 5; ++k) F[k] = -F[k];
#endif
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
int pick_tile_class(int, int, int) { return -1; }
int tile_class_threads(int, int) { return 0; }
size_t grad_smem_bytes(const DevMesh &, bool) { return 0; }
size_t flux_smem_bytes(const DevMesh &m, bool, bool) { return gather_flux_smem(m); }
#else
template <class CAP>
static bool cap_fits(int cells, int faces, int halo) {
  return cells <= CAP::NC && faces <= CAP::FC && halo <= CAP::HC;
}
int pick_tile_class(int cells, int faces, int halo) {
  if (cap_fits<Cap64>(cells, faces, halo)) return 0;
  if (cap_fits<Cap128>(cells, faces, halo)) return 1;
  if (cap_fits<Cap256>(cells, faces, halo)) return 2;
  return -1;
}
int tile_class_threads(int cls, int which) {
  switch (cls) {
    case 0: return which ? Cap64::FLUX_THREADS : Cap64::GRAD_THREADS;
    case 1: return which ? Cap128::FLUX_THREADS : Cap128::GRAD_THREADS;
    case 2: return which ? Cap256::FLUX_THREADS : Cap256::GRAD_THREADS;
  }
  return 0;
}
// MINIAERO_GRAD_PERSIST=1: the gradient kernel as a persistent, double-buffered pipeline (tile i+1 is copied while
// tile i is computed).  Measured slower than one tile per CTA (0.77 vs 0.73 ms at 8.4 M cells): the kernel is bound
// by FP64 issue and warp count, not by the copy latency, and the second stage costs one resident CTA per SM.
static int grad_persistent() {  // 0: one tile per CTA, 1: persistent, 2: persistent with the geometry read directly
  const char *e = getenv("MINIAERO_GRAD_PERSIST");
  return e && (e[0] == '1' || e[0] == '2') ? e[0] - '0' : 0;
}
template <class CAP>
static size_t grad_tma_smem(bool second) {  // one or two stages + two mbarriers
  const int mode = grad_persistent();
  const int ng = mode == 2 ? 0 : (second ? 6 : 3);
  const int cell_d = MA_GRAD_STAGE_CELL ? (1 + (second ? 3 : 0)) * CAP::RC + (12 * CAP::SC * 2 + 7) / 8 : 0;
  return (size_t)(mode ? 2 : 1) * (ng * CAP::FC + 5 * CAP::LS + cell_d) * 8 + 16;
}
// grid of a persistent kernel: resident CTAs per SM x SMs of the current device
static int persistent_ctas(int ctas_per_sm) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms * ctas_per_sm;
}
template <class CAP>
static size_t flux_tma_smem_rt(bool second, bool viscous) {
  return second ? (viscous ? flux_tma_smem<true, true, CAP>() : flux_tma_smem<true, false, CAP>())
                : (viscous ? flux_tma_smem<false, true, CAP>() : flux_tma_smem<false, false, CAP>());
}
size_t grad_smem_bytes(const DevMesh &m, bool second) {
  if (m.grad_variant != 1) return 0;
  switch (m.tile_class) {
    case 0: return grad_tma_smem<Cap64>(second);
    case 1: return grad_tma_smem<Cap128>(second);
    case 2: return grad_tma_smem<Cap256>(second);
  }
  return 0;
}
size_t flux_smem_bytes(const DevMesh &m, bool second, bool viscous) {
  if (m.flux_variant != 1) return gather_flux_smem(m);
  switch (m.tile_class) {
    case 0: return flux_tma_smem_rt<Cap64>(second, viscous);
    case 1: return flux_tma_smem_rt<Cap128>(second, viscous);
    case 2: return flux_tma_smem_rt<Cap256>(second, viscous);
  }
  return gather_flux_smem(m);
}
template <class CAP>
static cudaError_t launch_grad_tma(const DevMesh &m, const double *V, double *grad, double *lim, bool second,
                                   int tile_begin, int ntiles, int n_first, cudaStream_t st) {
  const size_t smem = grad_tma_smem<CAP>(second);
  if (second && m.limiter == 1) {  // the alternative limiter: one tile per CTA (the persistent forms are experiments)
    grad_limiter_tma_kernel<true, CAP, false, false, true><<<ntiles, CAP::GRAD_THREADS, smem, st>>>(m, V, grad, lim, tile_begin, ntiles, n_first);
    return cudaGetLastError();
  }
  if (grad_persistent() == 2) {
    const int grid = std::min(ntiles, persistent_ctas(CAP::GRAD_MINB1));
    if (second)
      grad_limiter_tma_kernel<true, CAP, true, true><<<grid, CAP::GRAD_THREADS, smem, st>>>(m, V, grad, lim, tile_begin, ntiles, n_first);
    else
      grad_limiter_tma_kernel<false, CAP, true, true><<<grid, CAP::GRAD_THREADS, smem, st>>>(m, V, grad, lim, tile_begin, ntiles, n_first);
  } else if (grad_persistent()) {
    const int grid = std::min(ntiles, persistent_ctas(CAP::GRAD_MINB));
    if (second)
      grad_limiter_tma_kernel<true, CAP, true><<<grid, CAP::GRAD_THREADS, smem, st>>>(m, V, grad, lim, tile_begin, ntiles, n_first);
    else
      grad_limiter_tma_kernel<false, CAP, true><<<grid, CAP::GRAD_THREADS, smem, st>>>(m, V, grad, lim, tile_begin, ntiles, n_first);
  } else {
    if (second)
      grad_limiter_tma_kernel<true, CAP, false><<<ntiles, CAP::GRAD_THREADS, smem, st>>>(m, V, grad, lim, tile_begin, ntiles, n_first);
    else
      grad_limiter_tma_kernel<false, CAP, false><<<ntiles, CAP::GRAD_THREADS, smem, st>>>(m, V, grad, lim, tile_begin, ntiles, n_first);
  }
  return cudaGetLastError();
}
template <class CAP>
static cudaError_t launch_flux_tma(const DevMesh &m, const StageArgs &a, bool second, bool viscous, int tile_begin,
                                   int ntiles, int threads, cudaStream_t st) {
  if (threads <= 0 || threads > CAP::FLUX_THREADS) threads = CAP::FLUX_THREADS;
  if (second && viscous)
    flux_rk_tma_kernel<true, true, CAP><<<ntiles, threads, flux_tma_smem<true, true, CAP>(), st>>>(m, a, tile_begin);
  else if (second)
    flux_rk_tma_kernel<true, false, CAP><<<ntiles, threads, flux_tma_smem<true, false, CAP>(), st>>>(m, a, tile_begin);
  else if (viscous)
    flux_rk_tma_kernel<false, true, CAP><<<ntiles, threads, flux_tma_smem<false, true, CAP>(), st>>>(m, a, tile_begin);
  else
    flux_rk_tma_kernel<false, false, CAP><<<ntiles, threads, flux_tma_smem<false, false, CAP>(), st>>>(m, a, tile_begin);
  return cudaGetLastError();
}
template <class CAP>
static cudaError_t prepare_tma() {
  cudaError_t e;
#define MA_SET(K, BYTES)                                                                   \
  e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES));  \
  if (e != cudaSuccess) return e;                                                          \
  e = cudaFuncSetAttribute(K, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
  if (e != cudaSuccess) return e;
  MA_SET((grad_limiter_tma_kernel<true, CAP, true, true>), grad_tma_smem<CAP>(true))
  MA_SET((grad_limiter_tma_kernel<false, CAP, true, true>), grad_tma_smem<CAP>(false))
  MA_SET((grad_limiter_tma_kernel<true, CAP, true>), grad_tma_smem<CAP>(true))
  MA_SET((grad_limiter_tma_kernel<false, CAP, true>), grad_tma_smem<CAP>(false))
  MA_SET((grad_limiter_tma_kernel<true, CAP, false>), grad_tma_smem<CAP>(true))
  MA_SET((grad_limiter_tma_kernel<false, CAP, false>), grad_tma_smem<CAP>(false))
  MA_SET((grad_limiter_tma_kernel<true, CAP, false, false, true>), grad_tma_smem<CAP>(true))
  // the direct-geometry gradient variant stages 23 KB per CTA: leave the rest of the SM's array to L1, where the
  // second reader of a face's geometry finds it
  e = cudaFuncSetAttribute((grad_limiter_tma_kernel<true, CAP, true, true>), cudaFuncAttributePreferredSharedMemoryCarveout, 45);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute((grad_limiter_tma_kernel<false, CAP, true, true>), cudaFuncAttributePreferredSharedMemoryCarveout, 45);
  if (e != cudaSuccess) return e;
  MA_SET((flux_rk_tma_kernel<true, true, CAP>), (flux_tma_smem<true, true, CAP>()))
  MA_SET((flux_rk_tma_kernel<true, false, CAP>), (flux_tma_smem<true, false, CAP>()))
  MA_SET((flux_rk_tma_kernel<false, true, CAP>), (flux_tma_smem<false, true, CAP>()))
  MA_SET((flux_rk_tma_kernel<false, false, CAP>), (flux_tma_smem<false, false, CAP>()))
#undef MA_SET
  return cudaSuccess;
}
#endif

cudaError_t launch_grad_limiter(const DevMesh &m, const double *V, double *grad, double *lim, bool second,
                                int tile_begin, int ntiles, int n_first, int threads, cudaStream_t st) {
  if (ntiles <= 0) return cudaSuccess;
#ifndef MA_STRICT
  if (m.grad_variant == 1) {
    switch (m.tile_class) {
      case 0: return launch_grad_tma<Cap64>(m, V, grad, lim, second, tile_begin, ntiles, n_first, st);
      case 1: return launch_grad_tma<Cap128>(m, V, grad, lim, second, tile_begin, ntiles, n_first, st);
      case 2: return launch_grad_tma<Cap256>(m, V, grad, lim, second, tile_begin, ntiles, n_first, st);
    }
  }
#endif
  if (second)
    grad_limiter_kernel<true><<<ntiles, threads, 0, st>>>(m, V, grad, lim, tile_begin, n_first);
  else
    grad_limiter_kernel<false><<<ntiles, threads, 0, st>>>(m, V, grad, lim, tile_begin, n_first);
  return cudaGetLastError();
}

cudaError_t flux_rk_prepare(const DevMesh &m, int smem_bytes) {
  cudaError_t e;
  (void)m;
  // the attribute belongs to the kernel, not to a solver: keep the largest request of any solver of this process,
  // so that an earlier, larger-tiled solver can still launch after a later one was created
  static int high_water = 0;
  if (smem_bytes > high_water) high_water = smem_bytes;
  smem_bytes = high_water;
#define MA_SET(K)                                                                          \
  e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);    \
  if (e != cudaSuccess) return e;
  MA_SET((flux_rk_kernel<false, false>))
  MA_SET((flux_rk_kernel<false, true>))
  MA_SET((flux_rk_kernel<true, false>))
  MA_SET((flux_rk_kernel<true, true>))
#undef MA_SET
#ifndef MA_STRICT
  switch (m.tile_class) {
    case 0: return prepare_tma<Cap64>();
    case 1: return prepare_tma<Cap128>();
    case 2: return prepare_tma<Cap256>();
  }
#endif
  return cudaSuccess;
}

cudaError_t launch_flux_rk(const DevMesh &m, const StageArgs &a, bool second, bool viscous, int tile_begin,
                           int ntiles, int threads, cudaStream_t st) {
  if (ntiles <= 0) return cudaSuccess;
#ifndef MA_STRICT
  if (m.flux_variant == 1) {
    switch (m.tile_class) {
      case 0: return launch_flux_tma<Cap64>(m, a, second, viscous, tile_begin, ntiles, threads, st);
      case 1: return launch_flux_tma<Cap128>(m, a, second, viscous, tile_begin, ntiles, threads, st);
      case 2: return launch_flux_tma<Cap256>(m, a, second, viscous, tile_begin, ntiles, threads, st);
    }
  }
#endif
  const size_t smem = gather_flux_smem(m);
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

cudaError_t launch_set_state(const DevMesh &m, const double *aos, const int *old2new, double *Un, double *V, cudaStream_t st) {
  const int threads = 256;
  set_state_kernel<<<(m.n_owned + threads - 1) / threads, threads, 0, st>>>(m, aos, old2new, Un, V);
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
