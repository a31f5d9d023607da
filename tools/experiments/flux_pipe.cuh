// Pipelined form of the staged flux kernel (FAST arithmetic only; included by kernels.cu inside its namespace).
//
// flux_rk_tma_kernel is one tile per CTA: copy the tile's operand runs, wait, evaluate.  Its timing split
// (profiles/r01e_variants.md §2: copies alone 0.97 ms, arithmetic alone 1.08 ms, together 1.37 ms at 8.4 M cells)
// says a CTA spends a quarter of its life waiting for its own copies.  Here a CTA is persistent and owns TWO
// stages: while its threads evaluate tile i out of one stage, the bulk-copy engine fills the other with tile
// i + gridDim.x, so after the first tile nobody waits for a copy.  The per-SM budget is unchanged — two CTAs of
// 2 x 54 KB and 256 threads instead of four of 54 KB and 128 threads: the same 16 warps at <= 128 registers.
//
//   per tile:  wait for this tile's stage (mbarrier, transaction bytes)
//              phase 1  thread per tile face: limited extrapolation (Flux.h:109-132), Roe (+ viscous) or boundary flux,
//                       the flux overwriting the face's own geometry column (before it, the last warp starts the
//                       copies of the NEXT tile into the other stage)
//              gather the outside-cell record of the NEXT tile's cut face into registers (nothing else is live here;
//              it is consumed first thing in the next phase 1) and request this tile's RK operands
//              CTA barrier
//              phase 2  thread per own cell: slot-ordered gather (Flux.h:216-227), RK update, next stage primitives
//              CTA barrier (the stage may be refilled)
#pragma once

// MA_PIPE_GATHER  0: the next tile's outside record is gathered into registers after phase 1 and carried across the
//                    barriers and phase 2 (latency hidden; the registers are loop-carried)
//                 1: the record's lines are pulled into L2 after phase 1 and the record is gathered right before the
//                    wait for the stage (not loop-carried; an L2 hit's latency is exposed)
#ifndef MA_PIPE_GATHER
#define MA_PIPE_GATHER 0
#endif

template <bool SECOND, bool VISCOUS, class CAP>
struct PipeStage {
  using R = FluxRec<SECOND, VISCOUS>;
  static constexpr int FC = CAP::FC, RC = CAP::RC, SC = CAP::SC;
  static constexpr size_t REC_B = (size_t)R::NREC * RC * 8, G_B = (size_t)R::NGS * FC * 8, LR_B = (size_t)FC * 4;
  static constexpr size_t SLOT_B = (size_t)6 * SC * 2;
  static constexpr size_t BYTES = REC_B + G_B + LR_B + SLOT_B;
  static_assert(REC_B % 16 == 0 && G_B % 16 == 0 && LR_B % 16 == 0 && SLOT_B % 16 == 0, "bulk-copy alignment");
  double *sRec;           // [NREC][RC] records of the own cells
  double *sG;             // [NGS][FC]  face geometry, then the face flux
  unsigned *sLR;          // [FC]       tile-local connectivity
  unsigned short *sSlot;  // [6][SC]    slot map
  MA_DEV explicit PipeStage(unsigned char *p)
      : sRec(reinterpret_cast<double *>(p)),
        sG(reinterpret_cast<double *>(p + REC_B)),
        sLR(reinterpret_cast<unsigned *>(p + REC_B + G_B)),
        sSlot(reinterpret_cast<unsigned short *>(p + REC_B + G_B + LR_B)) {}
};
template <bool SECOND, bool VISCOUS, class CAP>
constexpr size_t flux_pipe_smem() {
  return 2 * PipeStage<SECOND, VISCOUS, CAP>::BYTES + 16 + 64;  // stages, two mbarriers, two tile descriptors
}

template <bool SECOND, bool VISCOUS, class CAP>
__global__ void __launch_bounds__(CAP::PIPE_THREADS, CAP::PIPE_MINB)
    flux_rk_pipe_kernel(const DevMesh m, const StageArgs a, int tile_begin, int ntiles) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using R = FluxRec<SECOND, VISCOUS>;
  using ST = PipeStage<SECOND, VISCOUS, CAP>;
  constexpr int FC = CAP::FC, RC = CAP::RC, SC = CAP::SC, NT = CAP::PIPE_THREADS;
  constexpr int R_G = R::R_G, R_L = R::R_L, R_X = R::R_X, NREC = R::NREC, NGEOM = R::NGEOM;
  static_assert(CAP::HC <= NT, "every cut face of a tile needs a thread of its own");
  using SRec = SmemRecord<SECOND, VISCOUS, RC>;
  using RRec = RegRecord<SECOND, VISCOUS>;
  const unsigned bar0 = smem_addr(smem_raw + 2 * ST::BYTES);
  const int tid = threadIdx.x;
  const int G = gridDim.x;
  int t = blockIdx.x;
  if (t >= ntiles) return;
  // (addresses of the tile table and the outside-cell lists are re-formed from the kernel parameters at every use:
  //  nothing but t, i and the record in flight is carried around the tile loop)
#define MA_PIPE_TILES (m.tiles + tile_begin)
#define MA_PIPE_HALO(tile) (m.tile_halo + ((size_t)tile_begin + (tile)) * m.halo_stride)
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    mbar_fence_init();
  }
  // Per-stage copy of the tile descriptor: warp 0 reads it from global memory when it starts the stage's copies and
  // publishes it here (before its arrive on the stage's barrier); everybody else picks it up after the wait — no
  // descriptor is held in registers across a phase 1.
  TileInfoDev *sTile = reinterpret_cast<TileInfoDev *>(smem_raw + 2 * ST::BYTES + 16);  // [2], 32 bytes apart
  int out_id = tid < m.halo_stride ? __ldg(MA_PIPE_HALO(t) + tid) : -1;

  // start every copy of tile TT into stage st (warp 0)
  // (the LAST warp: its threads evaluate one face per tile where the others evaluate two, so the ~50 copy requests
  //  — serialised lane by lane, each needs its operands in uniform registers — do not make it the warp the CTA waits for)
  const int lane = tid - (NT - 32);
  auto issue = [&](int tile, int st) {
    if (lane < 0) return;
    const TileInfoDev TT = MA_PIPE_TILES[tile];
    const ST S(smem_raw + (size_t)st * ST::BYTES);
    const unsigned bar = bar0 + 8u * st;
    const int sh = TT.cell_start & 1;
    const int hbb = (sh + TT.cell_count + 1) & ~1;
    const unsigned fcq = (unsigned)(TT.face_count + 15) & ~15u;
    const int ssh = TT.cell_start & 7;
    const unsigned vbytes = (unsigned)hbb * 8u, gbytes = fcq * 8u, lbytes = fcq * 4u;
    const unsigned sbytes = (unsigned)((ssh + TT.cell_count + 7) & ~7) * 2u;
    const size_t c0 = (size_t)(TT.cell_start - sh);
    if (lane == 0) {
      *reinterpret_cast<TileInfoDev *>(reinterpret_cast<unsigned char *>(sTile) + 32 * st) = TT;
      mbar_arrive_expect_tx(bar, NREC * vbytes + NGEOM * gbytes + lbytes + 6 * sbytes);
    }
    __syncwarp();
    // generic-proxy accesses to this stage (ordered before by the CTA barrier) precede the copy engine's writes
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    constexpr int NCOPY = NREC + 11 + NGEOM + 1 + 6;
    for (int i = lane; i < NCOPY; i += 32) {
      if (i < NREC) {
        const double *base = i < R_G   ? a.V + (size_t)i * m.stride
                             : i < R_L ? a.grad + (size_t)(i - R_G) * m.stride
                             : i < R_X ? a.lim + (size_t)(i - R_L) * m.stride
                                       : m.cell_xyz + (size_t)(i - R_X) * m.stride;
        bulk_g2s(smem_addr(S.sRec + i * RC), base + c0, vbytes, bar);
      } else if (i < NREC + 11) {  // volume, Un, Acc: read from global memory in phase 2, pulled into L2 now
        const int j = i - NREC;
        const double *base = j == 0 ? m.cell_vol : j < 6 ? a.Un + (size_t)(j - 1) * m.stride : a.Acc + (size_t)(j - 6) * m.stride;
        if (j == 0 || (j < 6 ? a.kind != 2 : a.kind != 0)) bulk_prefetch_l2(base + c0, vbytes);
      } else if (i < NREC + 11 + NGEOM) {
        const int gi = i - NREC - 11;
        bulk_g2s(smem_addr(S.sG + gi * FC), m.face_geom + (size_t)6 * TT.face_start + (size_t)gi * fcq, gbytes, bar);
      } else if (i == NREC + 11 + NGEOM) {
        bulk_g2s(smem_addr(S.sLR), m.face_lr + TT.face_start, lbytes, bar);
      } else {
        const int s = i - (NREC + 11 + NGEOM + 1);
        bulk_g2s(smem_addr(S.sSlot + s * SC), m.slot_face + (size_t)s * m.slot_stride + (TT.cell_start - ssh), sbytes, bar);
      }
    }
  };
  // SoA component stride.  Re-read through an opaque asm at the top of every tile: otherwise the compiler hoists the
  // ~45 loop-invariant component base addresses (V, gradient, limiter, centroid, Un, Acc, Vnext + k * stride) out of
  // the tile loop and keeps them in ~90 registers (196 instead of 128: measured with the launch bound lifted).
  int stride = m.stride;
  // the record of the cell across this thread's cut face: registers
  double orec[NREC];
  auto gather_outside = [&](int c) {
#pragma unroll
    for (int k = 0; k < 5; ++k) orec[k] = __ldg(a.V + (size_t)k * stride + c);
    if (R::GRAD) {
#pragma unroll
      for (int k = 0; k < 15; ++k) orec[R_G + k] = __ldg(a.grad + (size_t)k * stride + c);
    }
    if (SECOND) {
#pragma unroll
      for (int k = 0; k < 5; ++k) orec[R_L + k] = __ldg(a.lim + (size_t)k * stride + c);
#pragma unroll
      for (int d = 0; d < 3; ++d) orec[R_X + d] = __ldg(m.cell_xyz + (size_t)d * stride + c);
    }
  };

  __syncthreads();  // barriers initialised
  issue(t, 0);
  if (MA_PIPE_GATHER == 0 && out_id >= 0) gather_outside(out_id);

  for (int i = 0;; ++i) {
    asm volatile("" : "+r"(stride));
    const int st = i & 1;
    const bool has1 = t + G < ntiles;
    int next_id = -1;
    if (has1) {
      issue(t + G, st ^ 1);
      if (tid < m.halo_stride) next_id = __ldg(MA_PIPE_HALO(t + G) + tid);  // consumed after phase 1
      if (t + 2 * G < ntiles) {  // descriptor and outside-cell list of the tile after next: into L2 now
        if (tid == 0) prefetch_l2(MA_PIPE_TILES + t + 2 * G);
        if (tid * 32 < m.halo_stride) prefetch_l2(MA_PIPE_HALO(t + 2 * G) + tid * 32);
      }
    }
    if (MA_PIPE_GATHER == 1 && out_id >= 0) gather_outside(out_id);
    const ST S(smem_raw + (size_t)st * ST::BYTES);
    double *const sRec = S.sRec, *const sG = S.sG;
    mbar_wait(bar0 + 8u * st, (unsigned)(i >> 1) & 1u);
    const TileInfoDev T = *reinterpret_cast<const TileInfoDev *>(reinterpret_cast<const unsigned char *>(sTile) + 32 * st);
    const int nc = T.cell_count, nf = T.face_count;
    const int shift = T.cell_start & 1;
    const int hb = (shift + nc + 1) & ~1;  // doubles per staged own-cell run; positions >= hb are outside cells
    const int nh = nf - T.cut_start;
    const int sshift = T.cell_start & 7;

    // ---- phase 1: work item w: cut face cut_start + w for w < nh, closed / boundary face w - nh otherwise
    int w = tid;
    if (out_id >= 0) {  // == (w < nh): the tile's outside-cell list has one entry per cut face, -1 beyond
      const int e = T.cut_start + w;
      const unsigned lr = S.sLR[e];
      const int pl = (int)(lr & 0xffffu), pr = (int)(lr >> 16);
      FaceGeom Gm;
      double xf[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        Gm.n[d] = sG[d * FC + e];
        xf[d] = SECOND ? sG[(3 + d) * FC + e] : 0.0;
      }
      double Vl[5], Vr[5], gs[4], flux[5];
      // the outside record first: its registers are free before the own cell's record is read
      if (pr >= hb) {  // outside cell on the right
        face_side<SECOND, VISCOUS, true>(RRec{orec}, xf, Gm.n, Vr, gs);
        face_side<SECOND, VISCOUS, false>(SRec{sRec + pl}, xf, Gm.n, Vl, gs);
      } else {
        face_side<SECOND, VISCOUS, true>(RRec{orec}, xf, Gm.n, Vl, gs);
        face_side<SECOND, VISCOUS, false>(SRec{sRec + pr}, xf, Gm.n, Vr, gs);
      }
      interior_flux<VISCOUS>(Vl, Vr, gs, Gm, flux);
#pragma unroll
      for (int k = 0; k < 5; ++k) sG[k * FC + e] = flux[k];  // this thread's own column: geometry is dead
      w += NT;
    }
    for (; w < nf; w += NT) {  // closed and boundary faces
      const int e = w - nh;
      const unsigned lr = S.sLR[e];
      const int pl = (int)(lr & 0xffffu);
      const unsigned pr = lr >> 16;
      FaceGeom Gm;
#pragma unroll
      for (int d = 0; d < 3; ++d) Gm.n[d] = sG[d * FC + e];
      double flux[5];
      if (pr < 0xFFF0u) {
        // interior face: Flux.h:89-160
        double xf[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) xf[d] = SECOND ? sG[(3 + d) * FC + e] : 0.0;
        double Vl[5], Vr[5], gs[4];
        face_side<SECOND, VISCOUS, true>(SRec{sRec + pl}, xf, Gm.n, Vl, gs);
        face_side<SECOND, VISCOUS, false>(SRec{sRec + (int)pr}, xf, Gm.n, Vr, gs);
        interior_flux<VISCOUS>(Vl, Vr, gs, Gm, flux);
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
          mirror_state(Vl, Gm.n, Vr, area_norm);
        }
        face_roe_flux(Vl, Vr, Gm, flux);
        if (type == 3) {  // NoSlip_BC.h:114-139 — viscous wall flux regardless of options.viscous
          double xf[3], xc[3], vflux[5];
          const int cg = T.cell_start + (pl - shift);
          const unsigned fcp = (unsigned)(nf + 15) & ~15u;
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            xf[d] = SECOND ? sG[(3 + d) * FC + e]
                           : __ldg(m.face_geom + (size_t)6 * T.face_start + (size_t)(3 + d) * fcp + e);
            xc[d] = SECOND ? sRec[(R_X + d) * RC + pl] : __ldg(m.cell_xyz + (size_t)d * stride + cg);
          }
          noslip_viscous_flux(Vl, Gm.n, area_norm, xf, xc, vflux);
#pragma unroll
          for (int k = 0; k < 5; ++k) flux[k] -= vflux[k];  // slot = -iflux + vflux == -(iflux - vflux)
        }
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) sG[k * FC + e] = flux[k];
    }
    // the face registers are dead: request the next tile's outside record and this tile's RK operands now, so the
    // loads fly across the barrier, phase 2 and the wait for the next stage
    if (MA_PIPE_GATHER == 0) {
      if (next_id >= 0) gather_outside(next_id);
    } else {
      if (next_id >= 0) {
        const int c = next_id;
#pragma unroll
        for (int k = 0; k < 5; ++k) prefetch_l2(a.V + (size_t)k * stride + c);
        if (R::GRAD) {
#pragma unroll
          for (int k = 0; k < 15; ++k) prefetch_l2(a.grad + (size_t)k * stride + c);
        }
        if (SECOND) {
#pragma unroll
          for (int k = 0; k < 5; ++k) prefetch_l2(a.lim + (size_t)k * stride + c);
#pragma unroll
          for (int d = 0; d < 3; ++d) prefetch_l2(m.cell_xyz + (size_t)d * stride + c);
        }
      }
    }
    out_id = next_id;
    double pre_vol = 1.0, pre_un[5] = {0, 0, 0, 0, 0}, pre_acc[5] = {0, 0, 0, 0, 0};
    if (tid < nc) {
      const int c = T.cell_start + tid;
      pre_vol = __ldg(m.cell_vol + c);
      if (a.kind != 2) {
#pragma unroll
        for (int k = 0; k < 5; ++k) pre_un[k] = __ldg(a.Un + (size_t)k * stride + c);
      }
      if (a.kind != 0) {
#pragma unroll
        for (int k = 0; k < 5; ++k) pre_acc[k] = a.Acc[(size_t)k * stride + c];
      }
    }
    __syncthreads();

    // ---- phase 2: slot-ordered gather, residual, RK update; the next stage state is stored as primitives
    if (tid < nc) {
      const int c = T.cell_start + tid;
      const double dtv = a.dt * rcp(pre_vol);
      double Rs[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int s = 0; s < 6; ++s) {
        const unsigned sf = S.sSlot[s * SC + sshift + tid];
        const int e = (int)(sf & 0x3fffu);
        const double sg = (sf & 0x8000u) ? dtv : -dtv;  // Flux.h:172-178: left slot holds -flux, right slot +flux
#pragma unroll
        for (int k = 0; k < 5; ++k) Rs[k] = fma(sg, sG[k * FC + e], Rs[k]);
      }
      double Wn[5];  // conservative state the next stage is evaluated at (or the new solution)
      if (a.kind == 0) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          a.Acc[(size_t)k * stride + c] = fma(a.beta, Rs[k], pre_un[k]);
          Wn[k] = fma(a.alpha_next, Rs[k], pre_un[k]);
        }
      } else if (a.kind == 1) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          a.Acc[(size_t)k * stride + c] = fma(a.beta, Rs[k], pre_acc[k]);
          Wn[k] = fma(a.alpha_next, Rs[k], pre_un[k]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          Wn[k] = fma(a.beta, Rs[k], pre_acc[k]);
          a.Un[(size_t)k * stride + c] = Wn[k];
        }
      }
      double Vn[5];
      compute_primitives(Wn, Vn);
#pragma unroll
      for (int k = 0; k < 5; ++k) a.Vnext[(size_t)k * stride + c] = Vn[k];
    }
    if (!has1) break;
    __syncthreads();  // every read of this stage is done before the tile after next is copied into it
    t += G;
  }
}
#undef MA_PIPE_TILES
#undef MA_PIPE_HALO
