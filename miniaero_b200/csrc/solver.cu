// C-ABI implementation of the solver object: the drop-in for TimeSolverExplicitRK4<Device>
// (TimeSolverExplicitRK4.h:160-539).  Owns the device copy of the mesh in the tile-packed SoA layout,
// the four state vectors, gradient and limiter fields, the halo buffers, streams and events.
//
// One RK stage (second order):   grad_limiter_kernel  ->  [halo: gradient+limiter]  ->  flux_rk_kernel
//                                ->  [halo: next stage state]
// versus the reference's ~37 parallel_for launches, 5 host-staged exchanges and ~12 fences per stage
// (TimeSolverExplicitRK4.h:352-486).  With overlap_halo the tiles that touch no ghost cell run while
// the exchanges are in flight on a second stream.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "comm.h"
#include "host_common.h"
#include "geom_kernels.h"
#include "kernels.h"
#include "layout.h"
#include "miniaero_b200.h"
#include "topology_kernels.h"

namespace {

#define MA_CUDA_TRY(expr)                                                                                   \
  do {                                                                                                      \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess)                                                                                  \
      return ma_set_error(MA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                 \
  } while (0)

// ---- layout-independent helper kernels -----------------------------------------------------------------
// device SoA [ncomp][stride] in renumbered order -> caller order AoS [n_owned][ncomp] (any field: get_field)
__global__ void soa_to_caller_kernel(const double *__restrict__ soa, int stride, int ncomp, int n_owned,
                                     const int *__restrict__ old2new, double *__restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)n_owned * ncomp) return;
  const int cell = (int)(i / ncomp), k = (int)(i % ncomp);
  out[i] = soa[(size_t)k * stride + old2new[cell]];
}
// the conservative state, one thread per cell (five strided loads, 40 contiguous bytes stored)
__global__ void state_to_caller_kernel(const double *__restrict__ soa, int stride, int n_owned,
                                       const int *__restrict__ old2new, double *__restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_owned) return;
  const int n = old2new[c];
#pragma unroll
  for (int k = 0; k < 5; ++k) out[(size_t)5 * c + k] = soa[(size_t)k * stride + n];
}
// halo pack / unpack (CopyGhost.h:93-211): buf[i][col0 + k] <-> field[k][ids[i]]
__global__ void pack_kernel(const double *__restrict__ field, int stride, int ncomp, const int *__restrict__ ids,
                            int count, double *__restrict__ buf, int row, int col0) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)count * ncomp) return;
  const int cell = (int)(i / ncomp), k = (int)(i % ncomp);
  buf[(size_t)cell * row + col0 + k] = field[(size_t)k * stride + ids[cell]];
}
__global__ void unpack_kernel(double *__restrict__ field, int stride, int ncomp, const int *__restrict__ ids,
                              int count, const double *__restrict__ buf, int row, int col0) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)count * ncomp) return;
  const int cell = (int)(i / ncomp), k = (int)(i % ncomp);
  field[(size_t)k * stride + ids[cell]] = buf[(size_t)cell * row + col0 + k];
}

template <class T>
int dev_alloc(T **p, size_t n, size_t *tally) {
  *p = nullptr;
  if (n == 0) return MA_OK;
  cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
  if (e != cudaSuccess) (void)cudaGetLastError();  // the failure is reported here: do not leave it as the sticky "last error"
  if (e != cudaSuccess)
    return ma_set_error(e == cudaErrorMemoryAllocation ? MA_ERR_NOMEM : MA_ERR_CUDA,
                        std::string("cudaMalloc of ") + std::to_string(n * sizeof(T)) + " bytes: " + cudaGetErrorString(e));
  if (tally) *tally += n * sizeof(T);
  return MA_OK;
}
template <class T, class A>
int dev_upload(T **p, const std::vector<T, A> &v, size_t *tally) {
  int rc = dev_alloc(p, v.size(), tally);
  if (rc) return rc;
  if (!v.empty()) MA_CUDA_TRY(cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return MA_OK;
}

}  // namespace

struct ma_solver {
  ma_options opt;
  ma_solver_config cfg;
  bool strict = false, second = false, viscous = false, need_grad = false;
  int device = 0;
  // layout (small host pieces kept)
  int n_owned = 0, n_ghost = 0, stride = 0, n_tiles = 0, n_interior_tiles = 0;
  long n_tile_faces_real = 0;
  std::vector<int> peer_rank, peer_send, peer_recv;
  int n_send = 0, n_recv = 0;
  // device
  ma::DevMesh dm;
  ma::TileInfoDev *d_tiles = nullptr;
  double *d_xyz = nullptr, *d_vol = nullptr, *d_geom = nullptr;
  uint16_t *d_slot = nullptr, *d_slot_nbr = nullptr;
  int *d_fl = nullptr, *d_fr = nullptr, *d_old2new = nullptr, *d_send_ids = nullptr, *d_recv_ids = nullptr;
  uint32_t *d_face_lr = nullptr;
  int *d_tile_halo = nullptr, *d_tile_pub = nullptr;
  double *d_cut_flux = nullptr;
  int launch_count[4] = {0, 0, 0, 0};  // tiles per flux launch class (layout.h: interior / boundary x first / second pass)
  size_t n_slot_entries = 0, n_face_entries = 0, n_halo_entries = 0, n_geom_entries = 0;  // array sizes (debug hook)
  bool topology_on_device = false;
  double *d_Un = nullptr, *d_Acc = nullptr, *d_V[2] = {nullptr, nullptr}, *d_grad = nullptr, *d_lim = nullptr;
  int vcur = 0;  // d_V[vcur] holds the primitives of the state the next stage is evaluated at
  double *d_sendbuf = nullptr, *d_recvbuf = nullptr, *d_stage = nullptr;
  size_t stage_elems = 0;
  size_t device_bytes = 0;
  const double *last_stage_prims = nullptr;
  // execution
  cudaStream_t st = nullptr, cs = nullptr;
  bool own_stream = false, own_cs = false;
  cudaEvent_t ev_u = nullptr, ev_a = nullptr, ev_gl = nullptr, ev_b = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  // per-kernel-class timing: event pairs recorded around launches, read back after the step's final sync
  struct ProfPair {
    cudaEvent_t a, b;
    double *acc;
  };
  std::vector<ProfPair> prof_pool;
  size_t prof_used = 0;
  int flux_threads = 256, grad_threads = 128;
  bool profiling = false;
  bool u_pending = false;  // a state exchange is in flight on cs (ev_u marks its end)
  ma_timing tm;
  double sim_time = 0.0;
  long time_it = 0;
  // one RK4 step (4 stages, 8 launches) captured as a CUDA graph: single-domain runs without per-kernel profiling
  // replay it, so a step costs one launch on the host (the reference's own test meshes are launch-bound on a B200).
  // Indexed by the parity of vcur at the start of the step (a step flips it four times).
  cudaGraphExec_t step_graph[2] = {nullptr, nullptr};
  long long step_graph_launches = 0;  // kernel launches one replay stands for
  bool use_graph = true;
  // ma_solver_submit pipeline: two upload and two download staging buffers, one copy stream per direction
  struct Pipe {
    cudaStream_t cin = nullptr, cout = nullptr;
    double *d_in[2] = {nullptr, nullptr}, *d_out[2] = {nullptr, nullptr};
    cudaEvent_t in_ready[2] = {nullptr, nullptr}, in_free[2] = {nullptr, nullptr};
    cudaEvent_t out_ready[2] = {nullptr, nullptr}, out_free[2] = {nullptr, nullptr};
    long submitted = 0;
    bool ready = false;
  } pipe;
};

namespace {

struct Api {
  decltype(&ma_fast::launch_grad_limiter) grad;
  decltype(&ma_fast::launch_flux_rk) flux;
  decltype(&ma_fast::flux_rk_prepare) prepare;
  decltype(&ma_fast::grad_smem_bytes) grad_smem;
  decltype(&ma_fast::flux_smem_bytes) flux_smem;
  decltype(&ma_fast::launch_initial_conditions) ic;
  decltype(&ma_fast::launch_primitives) prims;
  decltype(&ma_fast::launch_set_state) set_state;
};
Api api_of(bool strict) {
  if (strict)
    return {&ma_strict::launch_grad_limiter, &ma_strict::launch_flux_rk, &ma_strict::flux_rk_prepare,
            &ma_strict::grad_smem_bytes, &ma_strict::flux_smem_bytes,
            &ma_strict::launch_initial_conditions, &ma_strict::launch_primitives, &ma_strict::launch_set_state};
  return {&ma_fast::launch_grad_limiter, &ma_fast::launch_flux_rk, &ma_fast::flux_rk_prepare,
          &ma_fast::grad_smem_bytes, &ma_fast::flux_smem_bytes,
          &ma_fast::launch_initial_conditions, &ma_fast::launch_primitives, &ma_fast::launch_set_state};
}

int check_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return ma_set_error(MA_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") +
                                         (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0"));
  if (device < 0 || device >= count) return ma_set_error(MA_ERR_INVALID, "device ordinal out of range");
  MA_CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp p;
  MA_CUDA_TRY(cudaGetDeviceProperties(&p, device));
  if (p.major != 10)
    return ma_set_error(MA_ERR_CUDA, std::string("device '") + p.name + "' is sm_" + std::to_string(p.major) +
                                         std::to_string(p.minor) + "; this library carries sm_100a code only");
  return MA_OK;
}

// Optional per-kernel-class timing (off by default).  An event pair is recorded on the compute stream
// around the launches of one class; nothing is synchronised here, so the timed region is not perturbed.
// collect_profile() reads the pairs back once the step's closing event has completed.
struct ProfScope {
  ma_solver *S;
  ma_solver::ProfPair *p = nullptr;
  cudaStream_t stream;
  ProfScope(ma_solver *s, double *acc, cudaStream_t on = nullptr) : S(s), stream(on ? on : s->st) {
    if (!S->profiling) return;
    if (S->prof_used == S->prof_pool.size()) {
      ma_solver::ProfPair np = {nullptr, nullptr, nullptr};
      if (cudaEventCreate(&np.a) != cudaSuccess || cudaEventCreate(&np.b) != cudaSuccess) return;
      S->prof_pool.push_back(np);
    }
    p = &S->prof_pool[S->prof_used++];
    p->acc = acc;
    cudaEventRecord(p->a, stream);
  }
  ~ProfScope() {
    if (p) cudaEventRecord(p->b, stream);
  }
};
void collect_profile(ma_solver *S) {
  for (size_t i = 0; i < S->prof_used; ++i) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, S->prof_pool[i].a, S->prof_pool[i].b) == cudaSuccess) *S->prof_pool[i].acc += ms * 1e-3;
  }
  S->prof_used = 0;
}

// one halo round on stream `s`: pack the listed fields of the send cells, exchange, unpack into the ghosts
struct FieldRef {
  double *ptr;
  int ncomp;
};
int halo_exchange(ma_solver *S, const FieldRef *fields, int nfields, cudaStream_t s) {
  if (S->n_ghost == 0) return MA_OK;
  ProfScope prof(S, &S->tm.halo_seconds, s);  // pack + exchange + unpack, timed on the stream they run on
  int row = 0;
  for (int f = 0; f < nfields; ++f) row += fields[f].ncomp;
  const int threads = 256;
  int col = 0;
  for (int f = 0; f < nfields; ++f) {
    const long work = (long)S->n_send * fields[f].ncomp;
    if (work)
      pack_kernel<<<(unsigned)((work + threads - 1) / threads), threads, 0, s>>>(
          fields[f].ptr, S->stride, fields[f].ncomp, S->d_send_ids, S->n_send, S->d_sendbuf, row, col);
    col += fields[f].ncomp;
    S->tm.kernel_launches++;
  }
  MA_CUDA_TRY(cudaGetLastError());
  int rc = ma::comm_exchange(S->cfg.comm, S->d_sendbuf, S->d_recvbuf, row, (int)S->peer_rank.size(),
                             S->peer_rank.data(), S->peer_send.data(), S->peer_recv.data(), s);
  if (rc) return rc;
  col = 0;
  for (int f = 0; f < nfields; ++f) {
    const long work = (long)S->n_recv * fields[f].ncomp;
    if (work)
      unpack_kernel<<<(unsigned)((work + threads - 1) / threads), threads, 0, s>>>(
          fields[f].ptr, S->stride, fields[f].ncomp, S->d_recv_ids, S->n_recv, S->d_recvbuf, row, col);
    col += fields[f].ncomp;
    S->tm.kernel_launches++;
  }
  MA_CUDA_TRY(cudaGetLastError());
  return MA_OK;
}

// start the exchange of a freshly written state vector on the comm stream; ev_u marks completion
int start_state_exchange(ma_solver *S, double *state) {
  if (S->n_ghost == 0) return MA_OK;
  MA_CUDA_TRY(cudaEventRecord(S->ev_b, S->st));
  MA_CUDA_TRY(cudaStreamWaitEvent(S->cs, S->ev_b, 0));
  FieldRef f = {state, 5};
  int rc = halo_exchange(S, &f, 1, S->cs);
  if (rc) return rc;
  MA_CUDA_TRY(cudaEventRecord(S->ev_u, S->cs));
  S->u_pending = true;
  return MA_OK;
}
int wait_state_exchange(ma_solver *S) {
  if (S->u_pending) {
    ProfScope exposed(S, &S->tm.halo_wait_seconds);  // a->b spans exactly the stall of the compute stream
    MA_CUDA_TRY(cudaStreamWaitEvent(S->st, S->ev_u, 0));
    S->u_pending = false;
  }
  return MA_OK;
}

// the flux launches of one tile group (0: tiles that touch no ghost cell, 1: the others): one launch, or — shared cut
// faces — the group's two passes in order (the second imports what the first published; stream order is the only
// synchronisation)
int launch_flux_group(ma_solver *S, const Api &K, const ma::StageArgs &a, int group) {
  int begin = group ? S->launch_count[0] + S->launch_count[1] : 0;
  if (!S->n_ghost && group) return MA_OK;  // single domain: every tile is in group 0
  for (int pass = 0; pass < 2; ++pass) {
    const int n = S->launch_count[2 * group + pass];
    if (n > 0) {
      MA_CUDA_TRY(K.flux(S->dm, a, S->second, S->viscous, begin, n, S->flux_threads, S->st));
      S->tm.kernel_launches++;
    }
    begin += n;
  }
  return MA_OK;
}

int run_stage(ma_solver *S, const Api &K, int k) {
  static const double alpha[4] = {0.0, 1.0 / 2.0, 1.0 / 2.0, 1.0};                  // TimeSolverExplicitRK4.h:188-191
  static const double beta[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};       // :192-195
  const double *V = S->d_V[S->vcur];
  double *Vnext = S->d_V[S->vcur ^ 1];
  ma::StageArgs a;
  a.V = V;
  a.Vnext = Vnext;
  a.Un = S->d_Un;
  a.Acc = S->d_Acc;
  a.grad = S->d_grad;
  a.lim = S->d_lim;
  a.dt = S->opt.dt;
  a.alpha_next = (k < 3) ? alpha[k + 1] : 0.0;
  a.beta = beta[k];
  a.kind = (k == 0) ? 0 : (k == 3) ? 2 : 1;
  S->last_stage_prims = V;
  const int nint = S->n_ghost ? S->n_interior_tiles : S->n_tiles;
  const int nbnd = S->n_tiles - nint;

  if (S->need_grad) {
    {
      ProfScope p(S, &S->tm.grad_seconds);
      MA_CUDA_TRY(K.grad(S->dm, V, S->d_grad, S->d_lim, S->second, 0, nint, S->launch_count[0], S->grad_threads, S->st));
      S->tm.kernel_launches += nint > 0;
      int rc = wait_state_exchange(S);
      if (rc) return rc;
      MA_CUDA_TRY(K.grad(S->dm, V, S->d_grad, S->d_lim, S->second, nint, nbnd, S->launch_count[2], S->grad_threads, S->st));
      S->tm.kernel_launches += nbnd > 0;
    }
    if (S->n_ghost) {  // gradient (+ limiter) halo: GreenGauss.h:324-338, StencilLimiter.h:618-631
      MA_CUDA_TRY(cudaEventRecord(S->ev_a, S->st));
      MA_CUDA_TRY(cudaStreamWaitEvent(S->cs, S->ev_a, 0));
      FieldRef f[2] = {{S->d_grad, 15}, {S->d_lim, 5}};
      int rc = halo_exchange(S, f, S->second ? 2 : 1, S->cs);
      if (rc) return rc;
      MA_CUDA_TRY(cudaEventRecord(S->ev_gl, S->cs));
    }
    {
      ProfScope p(S, &S->tm.flux_seconds);
      int rc = launch_flux_group(S, K, a, 0);
      if (rc) return rc;
      if (S->n_ghost) {
        ProfScope exposed(S, &S->tm.halo_wait_seconds);
        MA_CUDA_TRY(cudaStreamWaitEvent(S->st, S->ev_gl, 0));
      }
      rc = launch_flux_group(S, K, a, 1);
      if (rc) return rc;
    }
  } else {
    ProfScope p(S, &S->tm.flux_seconds);
    int rc = launch_flux_group(S, K, a, 0);
    if (rc) return rc;
    rc = wait_state_exchange(S);
    if (rc) return rc;
    rc = launch_flux_group(S, K, a, 1);
    if (rc) return rc;
  }
  S->vcur ^= 1;
  return start_state_exchange(S, Vnext);  // ghosts of the next stage state (TimeSolverExplicitRK4.h:359-375)
}

// One RK4 time step on the solver's stream (TimeSolverExplicitRK4.h:340-491).  Single-domain runs replay a CUDA graph
// of the four stages (captured on first use); runs with a halo exchange (NCCL on a second stream) or with per-kernel
// profiling events launch the stages directly.
int one_step(ma_solver *S, const Api &K) {
  S->sim_time += S->opt.dt;  // TimeSolverExplicitRK4.h:343
  S->time_it++;
  const bool graph_ok = S->use_graph && S->n_ghost == 0 && !S->profiling && !S->u_pending;
  const int par = S->vcur;
  // a step whose launches fail did not happen: the clock and the stage-buffer parity go back to where they were
  auto direct = [&]() {
    int rc = MA_OK;
    for (int k = 0; k < 4 && !rc; ++k) rc = run_stage(S, K, k);
    if (rc) {
      S->sim_time -= S->opt.dt;
      S->time_it--;
      S->vcur = par;
    }
    return rc;
  };
  if (!graph_ok) return direct();
  if (!S->step_graph[par]) {
    const long long before = S->tm.kernel_launches;
    cudaGraph_t g = nullptr;
    // a stream that cannot be captured (the legacy default stream, a stream already capturing) is handled like any
    // other capture failure: direct launches from now on
    const cudaError_t be = cudaStreamBeginCapture(S->st, cudaStreamCaptureModeThreadLocal);
    int rc = MA_OK;
    cudaError_t ce = be;
    if (be == cudaSuccess) {
      for (int k = 0; k < 4 && !rc; ++k) rc = run_stage(S, K, k);
      ce = cudaStreamEndCapture(S->st, &g);
    }
    S->step_graph_launches = S->tm.kernel_launches - before;
    S->tm.kernel_launches = before;
    S->vcur = par;  // nothing ran during the capture: the stage buffers are where they were
    if (rc || ce != cudaSuccess || !g) {  // capture not possible here: fall back to direct launches for good
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
      S->use_graph = false;
      return direct();
    }
    const cudaError_t ie = cudaGraphInstantiate(&S->step_graph[par], g, 0);
    cudaGraphDestroy(g);
    MA_CUDA_TRY(ie);
  }
  MA_CUDA_TRY(cudaGraphLaunch(S->step_graph[par], S->st));
  S->tm.kernel_launches += S->step_graph_launches;
  S->last_stage_prims = S->d_V[par ^ 1];  // stage 3 reads the buffer stage 2 wrote
  return MA_OK;
}

int ensure_staging(ma_solver *S, size_t elems) {
  if (S->stage_elems >= elems) return MA_OK;
  if (S->d_stage) {
    cudaFree(S->d_stage);
    S->device_bytes -= S->stage_elems * sizeof(double);
    S->d_stage = nullptr;
    S->stage_elems = 0;
  }
  int rc = dev_alloc(&S->d_stage, elems, &S->device_bytes);
  if (rc) return rc;
  S->stage_elems = elems;
  return MA_OK;
}

int download_field(ma_solver *S, const double *soa, int ncomp, double *host) {
  MA_CUDA_TRY(cudaSetDevice(S->device));
  int rc = wait_state_exchange(S);
  if (rc) return rc;
  const size_t elems = (size_t)S->n_owned * ncomp;
  rc = ensure_staging(S, elems);
  if (rc) return rc;
  const int threads = 256;
  soa_to_caller_kernel<<<(unsigned)((elems + threads - 1) / threads), threads, 0, S->st>>>(
      soa, S->stride, ncomp, S->n_owned, S->d_old2new, S->d_stage);
  MA_CUDA_TRY(cudaGetLastError());
  MA_CUDA_TRY(cudaMemcpyAsync(host, S->d_stage, elems * sizeof(double), cudaMemcpyDeviceToHost, S->st));
  MA_CUDA_TRY(cudaStreamSynchronize(S->st));
  return MA_OK;
}

}  // namespace

// staging buffers, streams and events of ma_solver_submit; all or nothing (pipe_release undoes a partial set-up)
static int pipe_setup(ma_solver *S, size_t elems) {
  ma_solver::Pipe &P = S->pipe;
  if (!P.cin) MA_CUDA_TRY(cudaStreamCreateWithFlags(&P.cin, cudaStreamNonBlocking));
  if (!P.cout) MA_CUDA_TRY(cudaStreamCreateWithFlags(&P.cout, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    double **bufs[2] = {&P.d_in[i], &P.d_out[i]};
    for (double **b : bufs) {
      if (*b) continue;
      int rc = dev_alloc(b, elems, &S->device_bytes);
      if (rc) return rc;
    }
    cudaEvent_t *evs[4] = {&P.in_ready[i], &P.in_free[i], &P.out_ready[i], &P.out_free[i]};
    for (cudaEvent_t *e : evs)
      if (!*e) MA_CUDA_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }
  P.ready = true;
  return MA_OK;
}
static void pipe_release(ma_solver *S) {
  ma_solver::Pipe &P = S->pipe;
  const size_t bytes = (size_t)S->n_owned * 5 * sizeof(double);
  if (P.cin) cudaStreamSynchronize(P.cin);
  if (P.cout) cudaStreamSynchronize(P.cout);
  for (int i = 0; i < 2; ++i) {
    double **bufs[2] = {&P.d_in[i], &P.d_out[i]};
    for (double **b : bufs)
      if (*b) {
        cudaFree(*b);
        *b = nullptr;
        S->device_bytes -= bytes;
      }
    cudaEvent_t *evs[4] = {&P.in_ready[i], &P.in_free[i], &P.out_ready[i], &P.out_free[i]};
    for (cudaEvent_t *e : evs)
      if (*e) {
        cudaEventDestroy(*e);
        *e = nullptr;
      }
  }
  if (P.cin) cudaStreamDestroy(P.cin);
  if (P.cout) cudaStreamDestroy(P.cout);
  P.cin = P.cout = nullptr;
  P.ready = false;
  P.submitted = 0;
  (void)cudaGetLastError();
}


extern "C" {

void ma_solver_config_default(ma_solver_config *cfg) {
  if (!cfg) return;
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->device = 0;
  cfg->arith = MA_ARITH_FAST;
  cfg->tile_dims[0] = cfg->tile_dims[1] = cfg->tile_dims[2] = 0;
  cfg->block_threads = 0;
  cfg->comm = nullptr;
  cfg->overlap_halo = 1;
  cfg->stream = nullptr;
  cfg->limiter = MA_LIMITER_VENKAT;
  cfg->share_cut_faces = 0;
}

void ma_solver_destroy(ma_solver *S) {
  if (!S) return;
  cudaSetDevice(S->device);
  if (S->st) cudaStreamSynchronize(S->st);
  if (S->cs) cudaStreamSynchronize(S->cs);
  void *ptrs[] = {S->d_tiles, S->d_xyz,  S->d_vol,  S->d_geom, S->d_slot,    S->d_fl,      S->d_fr,   S->d_old2new,
                  S->d_send_ids, S->d_recv_ids, S->d_face_lr, S->d_tile_halo, S->d_tile_pub, S->d_cut_flux, S->d_slot_nbr, S->d_Un, S->d_Acc, S->d_V[0], S->d_V[1], S->d_grad, S->d_lim,
                  S->d_sendbuf, S->d_recvbuf, S->d_stage};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  cudaEvent_t evs[] = {S->ev_u, S->ev_a, S->ev_gl, S->ev_b, S->ev_t0, S->ev_t1};
  for (cudaEvent_t e : evs)
    if (e) cudaEventDestroy(e);
  for (auto &pp : S->prof_pool) {
    if (pp.a) cudaEventDestroy(pp.a);
    if (pp.b) cudaEventDestroy(pp.b);
  }
  for (int i = 0; i < 2; ++i)
    if (S->step_graph[i]) cudaGraphExecDestroy(S->step_graph[i]);
  pipe_release(S);
  if (S->own_cs && S->cs) cudaStreamDestroy(S->cs);
  if (S->own_stream && S->st) cudaStreamDestroy(S->st);
  delete S;
}

// MINIAERO_LAYOUT_TIMING=1: where the set-up time goes (stderr)
struct SetupLap {
  bool on = getenv("MINIAERO_LAYOUT_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void operator()(const char *what) {
    if (!on) return;
    cudaDeviceSynchronize();
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "setup: %-34s %.3f s\n", what, std::chrono::duration<double>(now - t).count());
    t = now;
  }
};

// common front end of the two constructors: configuration checks and the tile size
static int create_prologue(const ma_options *opt, const ma_solver_config *cfg_in, ma_solver_config &cfg, int td[3]) {
  if (cfg_in)
    cfg = *cfg_in;
  else
    ma_solver_config_default(&cfg);
  if (cfg.arith != MA_ARITH_FAST && cfg.arith != MA_ARITH_STRICT)
    return ma_set_error(MA_ERR_INVALID, "ma_solver_create: arith must be MA_ARITH_FAST or MA_ARITH_STRICT");
  if (!(opt->dt > 0.0)) return ma_set_error(MA_ERR_INVALID, "ma_solver_create: dt must be positive");
  if (cfg.limiter != MA_LIMITER_VENKAT && cfg.limiter != MA_LIMITER_VANALBADA)
    return ma_set_error(MA_ERR_INVALID, "ma_solver_create: limiter must be MA_LIMITER_VENKAT or MA_LIMITER_VANALBADA");
  int rc = check_device(cfg.device);
  if (rc) return rc;
  // default tile: 8x8x8 cells for the STRICT kernels (flux staging only), 4x4x8 for the FAST staged kernels, whose
  // shared-memory footprint (28 doubles per own cell + 6 per face) leaves room for four CTAs per SM;
  // z is the fastest cell index inside a tile, so the long z edge makes the cut-face gathers along x and y contiguous
  const int td_default[2][3] = {{4, 4, 8}, {8, 8, 8}};
  for (int d = 0; d < 3; ++d)
    td[d] = cfg.tile_dims[d] > 0 ? cfg.tile_dims[d] : td_default[cfg.arith == MA_ARITH_STRICT ? 1 : 0][d];
  return MA_OK;
}

static int solver_from_layout(ma::HostLayout &L, const ma::StructuredGrid *grid, const ma_options *opt,
                              const ma_solver_config &cfg, ma_solver **out, const ma::TopoPlan *plan = nullptr);

// Shared cut faces (layout.h) pay once a flux launch is many waves of CTAs: two launches per stage instead of one, a
// sixth fewer face evaluations.  Default: on from `kShareMinCells` owned cells (the reference's own test meshes stay
// one launch per stage); MINIAERO_SHARE_CUT_FACES=0|1 overrides the default (developer knob), cfg.share_cut_faces both.
static bool want_shared_cut_faces(const ma_solver_config &cfg, long owned_cells) {
  const long kShareMinCells = 1L << 20;
  if (cfg.arith != MA_ARITH_FAST) return false;
  if (cfg.share_cut_faces > 0) return true;
  if (cfg.share_cut_faces < 0) return false;
  if (const char *e = getenv("MINIAERO_SHARE_CUT_FACES")) return e[0] == '1';
  return owned_cells >= kShareMinCells;
}

int ma_solver_create(const ma_mesh *mesh, const ma_options *opt, const ma_solver_config *cfg_in, ma_solver **out) {
  if (!mesh || !opt || !out) return ma_set_error(MA_ERR_INVALID, "ma_solver_create: null argument");
  *out = nullptr;
  ma_solver_config cfg;
  int td[3];
  int rc = create_prologue(opt, cfg_in, cfg, td);
  if (rc) return rc;
  if (mesh->num_ghosts > 0 && !cfg.comm)
    return ma_set_error(MA_ERR_INVALID, "ma_solver_create: mesh has ghost cells but no communicator was given");
  ma::HostLayout L;
  rc = ma::build_layout(*mesh, td, cfg.arith == MA_ARITH_STRICT, L, want_shared_cut_faces(cfg, mesh->num_owned_cells));
  if (rc) return rc;
  if (cfg.arith == MA_ARITH_FAST && !(L.max_frame_error <= 1e-9))
    return ma_set_error(MA_ERR_INVALID,
                        "MA_ARITH_FAST needs face (normal, tangent, binormal) triples that are orthogonal with unit tangent and "
                        "|binormal| = |normal| (as Face.C:81-96 builds them); use MA_ARITH_STRICT for arbitrary frames");
  return solver_from_layout(L, nullptr, opt, cfg, out);
}

int ma_solver_create_structured(const ma_options *opt, int rank, int num_ranks, const ma_solver_config *cfg_in,
                                ma_solver **out) {
  if (!opt || !out) return ma_set_error(MA_ERR_INVALID, "ma_solver_create_structured: null argument");
  *out = nullptr;
  ma_solver_config cfg;
  int td[3];
  int rc = create_prologue(opt, cfg_in, cfg, td);
  if (rc) return rc;
  if (num_ranks > 1 && !cfg.comm)
    return ma_set_error(MA_ERR_INVALID, "ma_solver_create_structured: more than one rank but no communicator was given");
  // MINIAERO_HOST_GEOMETRY=1 (debugging): evaluate the geometry on the host and upload it, as ma_solver_create does
  const char *hg = getenv("MINIAERO_HOST_GEOMETRY");
  const bool defer = !(hg && hg[0] == '1');
  ma::HostLayout L;
  ma::StructuredGrid grid;
  int nproc[3], block[3], nlocal[3] = {0, 0, 0}, offset[3];
  const bool share = ma_block_decomposition(opt, rank, num_ranks, nproc, block, nlocal, offset) == MA_OK &&
                     want_shared_cut_faces(cfg, (long)nlocal[0] * nlocal[1] * nlocal[2]);
  // Topology on the device (layout.h: TopoPlan; MINIAERO_DEVICE_TOPOLOGY=0 keeps it on the host): the host does the
  // O(tiles) + O(patterns) part, the device stamps the patterns.  Needs the staged kernels (the gather kernels read
  // global face -> cell lists the device builder does not make): a block whose tiles fit no capacity class, the STRICT
  // arithmetic and the gather-kernel knobs go through the host builder.
  {
    const char *dt = getenv("MINIAERO_DEVICE_TOPOLOGY"), *gv = getenv("MINIAERO_GRAD_KERNEL"), *fv = getenv("MINIAERO_FLUX_KERNEL");
    const bool wanted = cfg.arith == MA_ARITH_FAST && defer && !(dt && dt[0] == '0') && !(gv && !strcmp(gv, "gather")) &&
                        !(fv && !strcmp(fv, "gather"));
    if (wanted) {
      ma::TopoPlan plan;
      SetupLap lap;
      rc = ma::build_topology_plan(*opt, rank, num_ranks, td, share, L, &grid, plan);
      if (rc) return rc;
      lap("topology plan (host)");
      if (ma_fast::pick_tile_class(L.max_tile_cells_real, L.max_tile_faces, L.max_tile_halo) >= 0)
        return solver_from_layout(L, &grid, opt, cfg, out, &plan);
    }
  }
  rc = ma::build_layout_structured(*opt, rank, num_ranks, td, cfg.arith == MA_ARITH_STRICT, defer, L, &grid, share);
  if (rc) return rc;
  return solver_from_layout(L, &grid, opt, cfg, out);
}

// Topology on the device (layout.h: TopoPlan): allocates the solver's slot maps, tile-local connectivity, outside-cell
// and publish lists and the caller-order map, uploads the plan's O(tiles) + O(patterns) tables and stamps the patterns
// (topology_kernels.cu).  d_code / d_new2old are temporaries the geometry kernels need next; the caller frees them.
static int build_topology_on_device(ma_solver *S, const ma::HostLayout &L, const ma::StructuredGrid &grid,
                                    const ma::TopoPlan &P, uint32_t **d_code, int **d_new2old) {
  const long n_cells = (long)L.n_owned + L.n_ghost;
  std::vector<int> pat_ext, pat_dummy, pat_cell_off, pat_rank_off, pat_face_off;
  std::vector<uint32_t> cell_abc;
  std::vector<uint16_t> rank_of, face_lc;
  std::vector<uint8_t> face_slot;
  for (const ma::TopoPattern &p : P.patterns) {
    for (int d = 0; d < 3; ++d) pat_ext.push_back(p.ext[d]);
    pat_dummy.push_back(p.dummy_face);
    pat_cell_off.push_back((int)cell_abc.size());
    pat_rank_off.push_back((int)rank_of.size());
    pat_face_off.push_back((int)face_lc.size());
    cell_abc.insert(cell_abc.end(), p.cell_abc.begin(), p.cell_abc.end());
    rank_of.insert(rank_of.end(), p.rank_of.begin(), p.rank_of.end());
    face_lc.insert(face_lc.end(), p.face_lc.begin(), p.face_lc.end());
    face_slot.insert(face_slot.end(), p.face_slot.begin(), p.face_slot.end());
  }
  // scratch: the plan's tables
  int *d_tp = nullptr, *d_to = nullptr, *d_tn = nullptr, *d_pe = nullptr, *d_pd = nullptr, *d_pc = nullptr, *d_pr = nullptr,
      *d_pf = nullptr;
  unsigned char *d_tl = nullptr;
  uint32_t *d_abc = nullptr;
  uint16_t *d_rank = nullptr, *d_flc = nullptr;
  uint8_t *d_fs = nullptr;
  size_t scratch = 0;
  int rc = dev_upload(&d_tp, P.tile_pattern, &scratch);
  if (!rc) rc = dev_upload(&d_to, P.tile_origin, &scratch);
  if (!rc) rc = dev_upload(&d_tn, P.tile_nb, &scratch);
  if (!rc) rc = dev_upload(&d_tl, P.tile_launch, &scratch);
  if (!rc) rc = dev_upload(&d_pe, pat_ext, &scratch);
  if (!rc) rc = dev_upload(&d_pd, pat_dummy, &scratch);
  if (!rc) rc = dev_upload(&d_pc, pat_cell_off, &scratch);
  if (!rc) rc = dev_upload(&d_pr, pat_rank_off, &scratch);
  if (!rc) rc = dev_upload(&d_pf, pat_face_off, &scratch);
  if (!rc) rc = dev_upload(&d_abc, cell_abc, &scratch);
  if (!rc) rc = dev_upload(&d_rank, rank_of, &scratch);
  if (!rc) rc = dev_upload(&d_flc, face_lc, &scratch);
  if (!rc) rc = dev_upload(&d_fs, face_slot, &scratch);
  // persistent: what the stage kernels and the get / set calls read
  const size_t NS = (size_t)6 * L.slot_stride, NF = (size_t)L.n_tile_faces, NH = (size_t)L.n_tiles * L.halo_stride;
  if (!rc) rc = dev_alloc(&S->d_slot, NS, &S->device_bytes);
  if (!rc) rc = dev_alloc(&S->d_slot_nbr, NS, &S->device_bytes);
  if (!rc) rc = dev_alloc(&S->d_face_lr, NF, &S->device_bytes);
  if (!rc) rc = dev_alloc(&S->d_tile_halo, NH, &S->device_bytes);
  if (!rc && L.share_cut_faces) rc = dev_alloc(&S->d_tile_pub, NH, &S->device_bytes);
  if (!rc) rc = dev_alloc(&S->d_old2new, (size_t)n_cells, &S->device_bytes);
  if (!rc) rc = dev_alloc(d_code, NF, &scratch);
  if (!rc) rc = dev_alloc(d_new2old, (size_t)n_cells, &scratch);
  cudaError_t ce = cudaSuccess;
  if (!rc) {
    ce = cudaMemsetAsync(S->d_slot, 0, NS * sizeof(uint16_t), S->st);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(S->d_slot_nbr, 0xFF, NS * sizeof(uint16_t), S->st);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(S->d_face_lr, 0, NF * sizeof(uint32_t), S->st);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(*d_code, 0, NF * sizeof(uint32_t), S->st);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(S->d_tile_halo, 0xFF, NH * sizeof(int), S->st);
    if (ce == cudaSuccess && S->d_tile_pub) ce = cudaMemsetAsync(S->d_tile_pub, 0xFF, NH * sizeof(int), S->st);
    ma::TopoView t;
    t.g = grid.gen;
    t.g.xs = t.g.ys = t.g.zs = nullptr;  // numbering only
    t.n_owned = L.n_owned, t.n_tiles = L.n_tiles, t.slot_stride = L.slot_stride, t.halo_stride = L.halo_stride;
    t.import_capacity = L.import_capacity;
    for (int f = 0; f < 6; ++f) t.bc_of_face[f] = P.bc_of_face[f];
    t.tiles = S->d_tiles, t.tile_pattern = d_tp, t.tile_origin = d_to, t.tile_nb = d_tn, t.tile_launch = d_tl;
    t.pat_ext = d_pe, t.pat_dummy = d_pd, t.pat_cell_off = d_pc, t.pat_rank_off = d_pr, t.pat_face_off = d_pf;
    t.cell_abc = d_abc, t.rank_of = d_rank, t.face_lc = d_flc, t.face_slot = d_fs;
    t.new2old = *d_new2old, t.old2new = S->d_old2new, t.slot_face = S->d_slot, t.slot_nbr = S->d_slot_nbr;
    t.face_lr = S->d_face_lr, t.face_code = *d_code, t.tile_halo = S->d_tile_halo, t.tile_pub = S->d_tile_pub;
    if (ce == cudaSuccess) ce = ma::launch_device_topology(t, n_cells, S->st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(S->st);
  }
  for (void *p : {(void *)d_tp, (void *)d_to, (void *)d_tn, (void *)d_tl, (void *)d_pe, (void *)d_pd, (void *)d_pc, (void *)d_pr,
                  (void *)d_pf, (void *)d_abc, (void *)d_rank, (void *)d_flc, (void *)d_fs})
    if (p) cudaFree(p);
  if (rc) return rc;
  if (ce != cudaSuccess) return ma_set_error(MA_ERR_CUDA, std::string("device topology: ") + cudaGetErrorString(ce));
  return MA_OK;
}

static int solver_from_layout(ma::HostLayout &L, const ma::StructuredGrid *grid, const ma_options *opt,
                              const ma_solver_config &cfg, ma_solver **out, const ma::TopoPlan *plan) {
  int rc = MA_OK;
  SetupLap lap;
  ma_solver *S = new ma_solver();
  std::memset(&S->tm, 0, sizeof(S->tm));
  std::memset(&S->dm, 0, sizeof(S->dm));
  S->opt = *opt;
  S->cfg = cfg;
  S->device = cfg.device;
  S->strict = cfg.arith == MA_ARITH_STRICT;
  S->second = opt->second_order_space != 0;
  S->viscous = opt->viscous != 0;
  S->need_grad = S->second || S->viscous;  // TimeSolverExplicitRK4.h:383
  S->n_owned = L.n_owned, S->n_ghost = L.n_ghost, S->stride = L.stride;
  S->n_tiles = L.n_tiles, S->n_interior_tiles = L.n_interior_tiles;
  S->n_tile_faces_real = L.n_tile_faces_real;
  S->peer_rank = L.peer_rank, S->peer_send = L.peer_send_count, S->peer_recv = L.peer_recv_count;
  S->n_send = (int)L.send_ids.size(), S->n_recv = (int)L.recv_ids.size();
  if (cfg.block_threads > 0) S->flux_threads = S->grad_threads = std::min(256, (cfg.block_threads + 31) / 32 * 32);
  if (const char *ge = getenv("MINIAERO_CUDA_GRAPH")) S->use_graph = ge[0] != '0';  // experiment knob

#define MA_TRY(expr)        \
  do {                      \
    rc = (expr);            \
    if (rc) {               \
      ma_solver_destroy(S); \
      return rc;            \
    }                       \
  } while (0)
#define MA_CU(expr)                                                                                    \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) {                                                                           \
      ma_solver_destroy(S);                                                                            \
      return ma_set_error(MA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));            \
    }                                                                                                  \
  } while (0)

  if (cfg.stream) {
    S->st = (cudaStream_t)cfg.stream;
  } else {
    MA_CU(cudaStreamCreateWithFlags(&S->st, cudaStreamNonBlocking));
    S->own_stream = true;
  }
  if (S->n_ghost && cfg.overlap_halo) {
    // Highest priority: the pack kernel, NCCL's send/recv kernel and the unpack kernel are a few CTAs each, queued
    // behind an interior-tile launch of half a million CTAs.  At equal priority the block scheduler hands them SMs
    // only when the big kernel has nothing left to dispatch — the exchange would START when the work that is meant to
    // hide it ENDS (8 GPUs, round 2: 1.4 ms of exposed wait per stage, 7 % of the step).
    int prio_least = 0, prio_greatest = 0;
    MA_CU(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    MA_CU(cudaStreamCreateWithPriority(&S->cs, cudaStreamNonBlocking, prio_greatest));
    S->own_cs = true;
  } else {
    S->cs = S->st;
  }
  cudaEvent_t *evs[] = {&S->ev_u, &S->ev_a, &S->ev_gl, &S->ev_b};
  for (cudaEvent_t *e : evs) MA_CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  MA_CU(cudaEventCreate(&S->ev_t0));
  MA_CU(cudaEventCreate(&S->ev_t1));

  lap("streams, events");
  // ---- upload the layout
  {
    std::vector<ma::TileInfoDev> tiles(L.tiles.size());
    for (size_t i = 0; i < tiles.size(); ++i)
      tiles[i] = {L.tiles[i].cell_start, L.tiles[i].cell_count, L.tiles[i].face_start, L.tiles[i].face_count,
                  L.tiles[i].cut_start, L.tiles[i].n_eval, L.tiles[i].imp_area, 0};
    MA_TRY(dev_upload(&S->d_tiles, tiles, &S->device_bytes));
  }
  if (L.geometry_deferred) {
    // structured path: face and cell geometry are evaluated on the device (geom_kernels.cu) from the node tables,
    // the per-face codes and the cell permutation; the temporaries go away afterwards
    if (!grid) {
      ma_solver_destroy(S);
      return ma_set_error(MA_ERR_INVALID, "deferred geometry without a structured grid");
    }
    MA_TRY(dev_alloc(&S->d_xyz, (size_t)3 * L.stride, &S->device_bytes));
    MA_TRY(dev_alloc(&S->d_vol, (size_t)L.stride, &S->device_bytes));
    MA_TRY(dev_alloc(&S->d_geom, (size_t)L.geom_components * L.n_tile_faces, &S->device_bytes));
    double *d_xs = nullptr, *d_ys = nullptr, *d_zs = nullptr;
    uint32_t *d_code = nullptr;
    int *d_new2old = nullptr;
    size_t scratch = 0;
    int r2 = MA_OK;
    lap("geometry arrays allocated");
    if (plan) {  // the topology itself is built on the device: face codes and the cell permutation are already there
      r2 = build_topology_on_device(S, L, *grid, *plan, &d_code, &d_new2old);
      lap("topology on the device");
    } else {
      r2 = dev_upload(&d_code, L.face_code, &scratch);
      if (!r2) r2 = dev_upload(&d_new2old, L.new2old, &scratch);
    }
    if (!r2) r2 = dev_upload(&d_xs, grid->tables.xs, &scratch);
    if (!r2) r2 = dev_upload(&d_ys, grid->tables.ys, &scratch);
    if (!r2) r2 = dev_upload(&d_zs, grid->tables.zs, &scratch);
    cudaError_t ge = cudaSuccess;
    if (!r2) {
      ma::GridGen g = grid->gen;
      g.xs = d_xs, g.ys = d_ys, g.zs = d_zs;
      ge = ma::launch_device_geometry(g, S->d_tiles, L.n_tiles, d_code, S->d_geom, L.n_tile_faces, L.geom_components,
                                      d_new2old, (long)L.n_owned + L.n_ghost, L.stride, S->d_xyz, S->d_vol, S->st);
      if (ge == cudaSuccess) ge = cudaStreamSynchronize(S->st);
    }
    for (void *p : {(void *)d_xs, (void *)d_ys, (void *)d_zs, (void *)d_code, (void *)d_new2old})
      if (p) cudaFree(p);
    if (r2) {
      ma_solver_destroy(S);
      return r2;
    }
    MA_CU(ge);
    lap("geometry on the device");
    ma::BigVec<uint32_t>().swap(L.face_code);
  } else {
    MA_TRY(dev_upload(&S->d_xyz, L.cell_xyz, &S->device_bytes));
    MA_TRY(dev_upload(&S->d_vol, L.cell_vol, &S->device_bytes));
    MA_TRY(dev_upload(&S->d_geom, L.face_geom, &S->device_bytes));
    ma::BigVec<double>().swap(L.face_geom);
  }
  if (!plan) {
    MA_TRY(dev_upload(&S->d_slot, L.slot_face, &S->device_bytes));
    if (!S->strict) MA_TRY(dev_upload(&S->d_slot_nbr, L.slot_nbr, &S->device_bytes));
  }
  // kernel variants (FAST: the bulk-copy staged tile kernels when a capacity class holds every tile, else the gather
  // kernels; experiment knobs MINIAERO_GRAD_KERNEL / MINIAERO_FLUX_KERNEL = gather | tma).
  const char *gv = getenv("MINIAERO_GRAD_KERNEL"), *fv = getenv("MINIAERO_FLUX_KERNEL");
  const int tile_class = S->strict ? -1 : ma_fast::pick_tile_class(L.max_tile_cells_real, L.max_tile_faces, L.max_tile_halo);
  const int grad_variant =
      (tile_class >= 0 && !(gv && !strcmp(gv, "gather"))) ? 1 : 0;
  const int flux_variant = (tile_class < 0 || (fv && !strcmp(fv, "gather"))) ? 0 : 1;
  // the global face -> cell lists are read by the gather kernels only (the staged kernels use the 16-bit tile-local
  // connectivity): 29 bytes per cell that the default FAST configuration does not spend
  if (plan && (S->strict || grad_variant == 0 || flux_variant == 0)) {
    ma_solver_destroy(S);
    return ma_set_error(MA_ERR_INVALID, "device topology without the staged kernels (internal error)");
  }
  if (S->strict || grad_variant == 0 || flux_variant == 0) {
    MA_TRY(dev_upload(&S->d_fl, L.face_left, &S->device_bytes));
    MA_TRY(dev_upload(&S->d_fr, L.face_right, &S->device_bytes));
  }
  if (!plan) {
    MA_TRY(dev_upload(&S->d_face_lr, L.face_lr, &S->device_bytes));
    MA_TRY(dev_upload(&S->d_tile_halo, L.tile_halo, &S->device_bytes));
    if (L.share_cut_faces) MA_TRY(dev_upload(&S->d_tile_pub, L.tile_pub, &S->device_bytes));
    MA_TRY(dev_upload(&S->d_old2new, L.old2new, &S->device_bytes));
  }
  for (int i = 0; i < 4; ++i) S->launch_count[i] = L.launch_count[i];
  S->n_slot_entries = (size_t)6 * L.slot_stride, S->n_face_entries = (size_t)L.n_tile_faces;
  S->n_halo_entries = (size_t)L.n_tiles * L.halo_stride, S->n_geom_entries = (size_t)L.geom_components * L.n_tile_faces;
  S->topology_on_device = plan != nullptr;
  if (L.share_cut_faces)
    MA_TRY(dev_alloc(&S->d_cut_flux, (size_t)std::max(1, L.n_import_areas) * 5 * L.import_capacity, &S->device_bytes));
  MA_TRY(dev_upload(&S->d_send_ids, L.send_ids, &S->device_bytes));
  MA_TRY(dev_upload(&S->d_recv_ids, L.recv_ids, &S->device_bytes));
  const size_t sv = (size_t)5 * S->stride;
  double **states[] = {&S->d_Un, &S->d_Acc, &S->d_V[0], &S->d_V[1]};
  for (double **p : states) {
    MA_TRY(dev_alloc(p, sv, &S->device_bytes));
    MA_CU(cudaMemset(*p, 0, sv * sizeof(double)));  // ghosts start at zero like the reference's Views
  }
  if (S->need_grad) {
    MA_TRY(dev_alloc(&S->d_grad, (size_t)15 * S->stride, &S->device_bytes));
    MA_CU(cudaMemset(S->d_grad, 0, (size_t)15 * S->stride * sizeof(double)));
  }
  if (S->second) {
    MA_TRY(dev_alloc(&S->d_lim, sv, &S->device_bytes));
    MA_CU(cudaMemset(S->d_lim, 0, sv * sizeof(double)));
  }
  if (S->n_ghost) {
    MA_TRY(dev_alloc(&S->d_sendbuf, (size_t)20 * S->n_send, &S->device_bytes));
    MA_TRY(dev_alloc(&S->d_recvbuf, (size_t)20 * S->n_recv, &S->device_bytes));
  }

  ma::DevMesh &m = S->dm;
  m.n_owned = S->n_owned;
  m.n_cells = S->n_owned + S->n_ghost;
  m.stride = S->stride;
  m.n_tiles = S->n_tiles;
  m.slot_stride = L.slot_stride;
  m.n_tile_faces = L.n_tile_faces;
  m.flux_smem_stride = (L.max_tile_faces + 15) / 16 * 16 + 1;  // odd stride: conflict-free across components
  m.rk_smem_stride = (L.max_tile_cells + 15) / 16 * 16 + 1;
  m.max_tile_cells = L.max_tile_cells_real, m.max_tile_faces = L.max_tile_faces;
  m.max_tile_halo = L.max_tile_halo, m.max_tile_local = L.max_tile_local;
  m.halo_stride = L.halo_stride;
  m.limiter = cfg.limiter;
  m.tile_class = tile_class;
  m.grad_variant = grad_variant;
  m.flux_variant = flux_variant;
  m.slot_nbr = S->d_slot_nbr;
  m.face_lr = S->d_face_lr;
  m.tile_halo = S->d_tile_halo;
  m.tile_pub = S->d_tile_pub;
  m.cut_flux = S->d_cut_flux;
  m.import_capacity = L.import_capacity;
  m.tiles = S->d_tiles;
  m.cell_xyz = S->d_xyz;
  m.cell_vol = S->d_vol;
  m.slot_face = S->d_slot;
  m.face_geom = S->d_geom;
  m.face_left = S->d_fl;
  m.face_right = S->d_fr;
  // inflow state, TimeSolverExplicitRK4.h:218-223
  m.inflow[0] = 0.5805, m.inflow[1] = 503.96, m.inflow[2] = 0.0, m.inflow[3] = 0.0, m.inflow[4] = 343750.0;
  {
    const bool strict = S->strict;
    const Api K = api_of(strict);
    const size_t gather_smem = ((size_t)5 * m.flux_smem_stride + (size_t)11 * m.rk_smem_stride) * sizeof(double);
    const size_t smem = std::max(std::max(K.flux_smem(m, S->second, S->viscous), K.grad_smem(m, S->second)), gather_smem);
    if (smem > 227 * 1024) {
      ma_solver_destroy(S);
      return ma_set_error(MA_ERR_INVALID, "tile needs more than 227 KB of shared memory; use smaller tile_dims");
    }
    MA_CU(K.prepare(m, (int)gather_smem));
    if (cfg.block_threads <= 0 && !strict) {
      S->flux_threads = m.flux_variant >= 1 ? ma_fast::tile_class_threads(m.tile_class, 1) : 256;
      S->grad_threads = 128;
    }
  }
  lap("remaining uploads, state arrays");
  S->tm.device_bytes = S->device_bytes;
  S->tm.num_tiles = S->n_tiles;
  S->tm.tile_faces_total = (int)std::min<long>(L.n_tile_faces_real, 2147483647L);
  {
    long long ev = 0;
    for (const ma::TileInfo &T : L.tiles) ev += T.n_eval;
    S->tm.faces_evaluated = ev;
  }
  S->tm.num_interior_tiles = S->n_ghost ? S->n_interior_tiles : S->n_tiles;
  S->tm.num_send_cells = S->n_send;
  S->tm.num_recv_cells = S->n_recv;
  MA_CU(cudaDeviceSynchronize());
#undef MA_TRY
#undef MA_CU
  *out = S;
  return MA_OK;
}

int ma_solver_num_cells(const ma_solver *S, int *owned, int *ghosts) {
  if (!S) return ma_set_error(MA_ERR_INVALID, "null solver");
  if (owned) *owned = S->n_owned;
  if (ghosts) *ghosts = S->n_ghost;
  return MA_OK;
}

int ma_solver_initialize(ma_solver *S) {
  if (!S) return ma_set_error(MA_ERR_INVALID, "null solver");
  MA_CUDA_TRY(cudaSetDevice(S->device));
  const Api K = api_of(S->strict);
  const size_t sv = (size_t)5 * S->stride * sizeof(double);
  MA_CUDA_TRY(cudaMemsetAsync(S->d_Un, 0, sv, S->st));
  const double midx = S->opt.lx / 2.0;  // TimeSolverExplicitRK4.h:217
  MA_CUDA_TRY(K.ic(S->dm, S->d_Un, S->opt.problem_type, midx, S->st));
  S->vcur = 0;
  MA_CUDA_TRY(cudaMemsetAsync(S->d_V[0], 0, sv, S->st));
  MA_CUDA_TRY(K.prims(S->dm, S->d_Un, S->d_V[0], S->st));
  S->sim_time = 0.0;
  S->time_it = 0;
  int rc = start_state_exchange(S, S->d_V[0]);
  if (rc) return rc;
  return MA_OK;
}

int ma_solver_step(ma_solver *S, int nsteps) {
  if (!S) return ma_set_error(MA_ERR_INVALID, "null solver");
  if (nsteps < 0) return ma_set_error(MA_ERR_INVALID, "nsteps must be >= 0");
  MA_CUDA_TRY(cudaSetDevice(S->device));
  const Api K = api_of(S->strict);
  MA_CUDA_TRY(cudaEventRecord(S->ev_t0, S->st));
  for (int it = 0; it < nsteps; ++it) {
    int rc = one_step(S, K);
    if (rc) return rc;
  }
  int rc = wait_state_exchange(S);  // the timed region ends with the ghosts of the new solution in place
  if (rc) return rc;
  MA_CUDA_TRY(cudaEventRecord(S->ev_t1, S->st));
  MA_CUDA_TRY(cudaEventSynchronize(S->ev_t1));
  MA_CUDA_TRY(cudaGetLastError());
  float ms = 0;
  MA_CUDA_TRY(cudaEventElapsedTime(&ms, S->ev_t0, S->ev_t1));
  S->tm.step_seconds += ms * 1e-3;
  collect_profile(S);
  S->tm.steps += nsteps;
  S->tm.cell_updates += (long long)nsteps * S->n_owned;
  return MA_OK;
}

int ma_solver_solve(ma_solver *S) {
  if (!S) return ma_set_error(MA_ERR_INVALID, "null solver");
  const auto t0 = std::chrono::steady_clock::now();
  int rc = ma_solver_initialize(S);
  if (rc) return rc;
  const int rank = ma::comm_rank(S->cfg.comm);
  const int freq = S->opt.output_frequency > 0 ? S->opt.output_frequency : S->opt.ntimesteps + 1;
  int done = 0;
  while (done < S->opt.ntimesteps) {  // progress lines as TimeSolverExplicitRK4.h:346-349
    const int chunk = std::min(S->opt.ntimesteps - done, freq - (done % freq));
    rc = ma_solver_step(S, chunk);
    if (rc) return rc;
    done += chunk;
    if (done % freq == 0 && rank == 0)
      fprintf(stdout, "\nTime Step #%i:  Time = %16.9e; dt = %16.9e\n", done, S->sim_time, S->opt.dt);
  }
  rc = ma_solver_synchronize(S);
  if (rc) return rc;
  if (rank == 0) {
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stdout, "\n ... Device Run time: %8.2f seconds ...\n", sec);  // TimeSolverExplicitRK4.h:495
    fflush(stdout);
  }
  return MA_OK;
}

int ma_solver_synchronize(ma_solver *S) {
  if (!S) return ma_set_error(MA_ERR_INVALID, "null solver");
  MA_CUDA_TRY(cudaSetDevice(S->device));
  if (S->pipe.cin) MA_CUDA_TRY(cudaStreamSynchronize(S->pipe.cin));
  MA_CUDA_TRY(cudaStreamSynchronize(S->cs));
  MA_CUDA_TRY(cudaStreamSynchronize(S->st));
  if (S->pipe.cout) MA_CUDA_TRY(cudaStreamSynchronize(S->pipe.cout));
  MA_CUDA_TRY(cudaGetLastError());
  collect_profile(S);  // event pairs of submitted (unsynchronised) steps
  return MA_OK;
}

// Asynchronous ensemble member (see the header).  Three streams: the upload of member i+1 and the download of
// member i-1 run while member i is being stepped; the staging buffers are double-buffered and handed over by events.
int ma_solver_submit(ma_solver *S, const double *state_in, double *state_out, int nsteps) {
  if (!S || !state_in || !state_out) return ma_set_error(MA_ERR_INVALID, "ma_solver_submit: null argument");
  if (nsteps < 0) return ma_set_error(MA_ERR_INVALID, "nsteps must be >= 0");
  MA_CUDA_TRY(cudaSetDevice(S->device));
  ma_solver::Pipe &P = S->pipe;
  const size_t elems = (size_t)S->n_owned * 5;
  if (!P.ready) {
    int rc = pipe_setup(S, elems);
    if (rc) {
      pipe_release(S);  // a failed set-up (typically out of memory) leaves the solver as it was
      return rc;
    }
  }
  const Api K = api_of(S->strict);
  const int slot = (int)(P.submitted & 1);
  const bool reuse = P.submitted >= 2;
  const int threads = 256;
  // upload (copy-in stream)
  if (reuse) MA_CUDA_TRY(cudaStreamWaitEvent(P.cin, P.in_free[slot], 0));
  MA_CUDA_TRY(cudaMemcpyAsync(P.d_in[slot], state_in, elems * sizeof(double), cudaMemcpyHostToDevice, P.cin));
  MA_CUDA_TRY(cudaEventRecord(P.in_ready[slot], P.cin));
  // compute (solver stream)
  int rc = wait_state_exchange(S);
  if (rc) return rc;
  MA_CUDA_TRY(cudaStreamWaitEvent(S->st, P.in_ready[slot], 0));
  MA_CUDA_TRY(K.set_state(S->dm, P.d_in[slot], S->d_old2new, S->d_Un, S->d_V[S->vcur], S->st));
  MA_CUDA_TRY(cudaEventRecord(P.in_free[slot], S->st));
  rc = start_state_exchange(S, S->d_V[S->vcur]);
  if (rc) return rc;
  for (int it = 0; it < nsteps; ++it) {
    rc = one_step(S, K);
    if (rc) return rc;
  }
  rc = wait_state_exchange(S);
  if (rc) return rc;
  if (reuse) MA_CUDA_TRY(cudaStreamWaitEvent(S->st, P.out_free[slot], 0));
  state_to_caller_kernel<<<(unsigned)((S->n_owned + threads - 1) / threads), threads, 0, S->st>>>(
      S->d_Un, S->stride, S->n_owned, S->d_old2new, P.d_out[slot]);
  MA_CUDA_TRY(cudaGetLastError());
  MA_CUDA_TRY(cudaEventRecord(P.out_ready[slot], S->st));
  // download (copy-out stream)
  MA_CUDA_TRY(cudaStreamWaitEvent(P.cout, P.out_ready[slot], 0));
  MA_CUDA_TRY(cudaMemcpyAsync(state_out, P.d_out[slot], elems * sizeof(double), cudaMemcpyDeviceToHost, P.cout));
  MA_CUDA_TRY(cudaEventRecord(P.out_free[slot], P.cout));
  P.submitted++;
  S->tm.steps += nsteps;
  S->tm.cell_updates += (long long)nsteps * S->n_owned;
  S->tm.kernel_launches += 2;
  return MA_OK;
}

int ma_solver_get_solution(ma_solver *S, double *host) {
  if (!S || !host) return ma_set_error(MA_ERR_INVALID, "ma_solver_get_solution: null argument");
  MA_CUDA_TRY(cudaSetDevice(S->device));
  int rc = wait_state_exchange(S);
  if (rc) return rc;
  const size_t elems = (size_t)S->n_owned * 5;
  rc = ensure_staging(S, elems);
  if (rc) return rc;
  const int threads = 256;
  state_to_caller_kernel<<<(unsigned)((S->n_owned + threads - 1) / threads), threads, 0, S->st>>>(
      S->d_Un, S->stride, S->n_owned, S->d_old2new, S->d_stage);
  MA_CUDA_TRY(cudaGetLastError());
  MA_CUDA_TRY(cudaMemcpyAsync(host, S->d_stage, elems * sizeof(double), cudaMemcpyDeviceToHost, S->st));
  MA_CUDA_TRY(cudaStreamSynchronize(S->st));
  return MA_OK;
}

int ma_solver_set_solution(ma_solver *S, const double *host) {
  if (!S || !host) return ma_set_error(MA_ERR_INVALID, "ma_solver_set_solution: null argument");
  MA_CUDA_TRY(cudaSetDevice(S->device));
  int rc = wait_state_exchange(S);
  if (rc) return rc;
  const size_t elems = (size_t)S->n_owned * 5;
  rc = ensure_staging(S, elems);
  if (rc) return rc;
  MA_CUDA_TRY(cudaMemcpyAsync(S->d_stage, host, elems * sizeof(double), cudaMemcpyHostToDevice, S->st));
  MA_CUDA_TRY(api_of(S->strict).set_state(S->dm, S->d_stage, S->d_old2new, S->d_Un, S->d_V[S->vcur], S->st));
  return start_state_exchange(S, S->d_V[S->vcur]);
}

int ma_solver_get_field(ma_solver *S, int field, double *host) {
  if (!S || !host) return ma_set_error(MA_ERR_INVALID, "ma_solver_get_field: null argument");
  switch (field) {
    case MA_FIELD_GRADIENT:
      if (!S->d_grad) return ma_set_error(MA_ERR_INVALID, "gradients are not computed for first-order inviscid runs");
      return download_field(S, S->d_grad, 15, host);
    case MA_FIELD_LIMITER:
      if (!S->d_lim) return ma_set_error(MA_ERR_INVALID, "limiters are only computed for second-order runs");
      return download_field(S, S->d_lim, 5, host);
    case MA_FIELD_STAGE_PRIMITIVES:
      if (!S->last_stage_prims) return ma_set_error(MA_ERR_INVALID, "no RK stage has run yet");
      return download_field(S, S->last_stage_prims, 5, host);
    default:
      return ma_set_error(MA_ERR_INVALID, "unknown field");
  }
}

int ma_solver_debug_array(ma_solver *S, const char *name, void *out, size_t capacity, size_t *size) {
  if (!S || !name) return ma_set_error(MA_ERR_INVALID, "ma_solver_debug_array: null argument");
  MA_CUDA_TRY(cudaSetDevice(S->device));
  const std::string n = name;
  const void *src = nullptr;
  size_t bytes = 0;
  const size_t n_cells = (size_t)S->n_owned + S->n_ghost;
  if (n == "tiles") src = S->d_tiles, bytes = (size_t)S->n_tiles * sizeof(ma::TileInfoDev);
  else if (n == "slot_face") src = S->d_slot, bytes = S->n_slot_entries * sizeof(uint16_t);
  else if (n == "slot_nbr") src = S->d_slot_nbr, bytes = S->d_slot_nbr ? S->n_slot_entries * sizeof(uint16_t) : 0;
  else if (n == "face_lr") src = S->d_face_lr, bytes = S->n_face_entries * sizeof(uint32_t);
  else if (n == "tile_halo") src = S->d_tile_halo, bytes = S->n_halo_entries * sizeof(int);
  else if (n == "tile_pub") src = S->d_tile_pub, bytes = S->d_tile_pub ? S->n_halo_entries * sizeof(int) : 0;
  else if (n == "old2new") src = S->d_old2new, bytes = n_cells * sizeof(int);
  else if (n == "face_geom") src = S->d_geom, bytes = S->n_geom_entries * sizeof(double);
  else if (n == "cell_xyz") src = S->d_xyz, bytes = (size_t)3 * S->stride * sizeof(double);
  else if (n == "cell_vol") src = S->d_vol, bytes = (size_t)S->stride * sizeof(double);
  else if (n == "topology_on_device") bytes = S->topology_on_device ? 1 : 0;
  else return ma_set_error(MA_ERR_INVALID, "ma_solver_debug_array: unknown array '" + n + "'");
  if (size) *size = bytes;
  if (out && src && bytes) {
    MA_CUDA_TRY(cudaStreamSynchronize(S->st));
    MA_CUDA_TRY(cudaMemcpy(out, src, std::min(bytes, capacity), cudaMemcpyDeviceToHost));
  }
  return MA_OK;
}

int ma_solver_get_timing(ma_solver *S, ma_timing *t) {
  if (!S || !t) return ma_set_error(MA_ERR_INVALID, "ma_solver_get_timing: null argument");
  S->tm.device_bytes = S->device_bytes;
  *t = S->tm;
  return MA_OK;
}

int ma_solver_reset_timing(ma_solver *S) {
  if (!S) return ma_set_error(MA_ERR_INVALID, "null solver");
  S->tm.step_seconds = S->tm.grad_seconds = S->tm.flux_seconds = S->tm.halo_seconds = S->tm.halo_wait_seconds = 0.0;
  S->tm.steps = S->tm.cell_updates = S->tm.kernel_launches = 0;
  return MA_OK;
}

int ma_solver_set_profiling(ma_solver *S, int enabled) {
  if (!S) return ma_set_error(MA_ERR_INVALID, "null solver");
  S->profiling = enabled != 0;
  return MA_OK;
}

// ---- probes ----------------------------------------------------------------------------------------------
namespace {
struct DevBuf {
  double *p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  int up(const double *h, size_t n) {
    MA_CUDA_TRY(cudaMalloc((void **)&p, std::max<size_t>(n, 1) * sizeof(double)));
    if (h && n) MA_CUDA_TRY(cudaMemcpy(p, h, n * sizeof(double), cudaMemcpyHostToDevice));
    return MA_OK;
  }
  int down(double *h, size_t n) {
    MA_CUDA_TRY(cudaDeviceSynchronize());
    if (n) MA_CUDA_TRY(cudaMemcpy(h, p, n * sizeof(double), cudaMemcpyDeviceToHost));
    return MA_OK;
  }
};
#define MA_RC(expr)   \
  do {                \
    int _rc = (expr); \
    if (_rc) return _rc; \
  } while (0)
}  // namespace

int ma_probe_roe_flux(int n, const double *vl, const double *vr, const double *nn, const double *tt, const double *bb,
                      double *flux, int arith, int device) {
  if (n < 0 || !vl || !vr || !nn || !tt || !bb || !flux) return ma_set_error(MA_ERR_INVALID, "probe: bad argument");
  MA_RC(check_device(device));
  DevBuf a, b, c, d, e, f;
  const size_t N = (size_t)n;
  MA_RC(a.up(vl, 5 * N));
  MA_RC(b.up(vr, 5 * N));
  MA_RC(c.up(nn, 3 * N));
  MA_RC(d.up(tt, 3 * N));
  MA_RC(e.up(bb, 3 * N));
  MA_RC(f.up(nullptr, 5 * N));
  MA_CUDA_TRY(arith == MA_ARITH_STRICT ? ma_strict::probe_roe(n, a.p, b.p, c.p, d.p, e.p, f.p, 0)
                                       : ma_fast::probe_roe(n, a.p, b.p, c.p, d.p, e.p, f.p, 0));
  return f.down(flux, 5 * N);
}

int ma_probe_viscous_flux(int n, const double *grad, const double *prim, const double *normal, double *vflux,
                          int arith, int device) {
  if (n < 0 || !grad || !prim || !normal || !vflux) return ma_set_error(MA_ERR_INVALID, "probe: bad argument");
  MA_RC(check_device(device));
  DevBuf a, b, c, f;
  const size_t N = (size_t)n;
  MA_RC(a.up(grad, 15 * N));
  MA_RC(b.up(prim, 5 * N));
  MA_RC(c.up(normal, 3 * N));
  MA_RC(f.up(nullptr, 5 * N));
  MA_CUDA_TRY(arith == MA_ARITH_STRICT ? ma_strict::probe_viscous(n, a.p, b.p, c.p, f.p, 0)
                                       : ma_fast::probe_viscous(n, a.p, b.p, c.p, f.p, 0));
  return f.down(vflux, 5 * N);
}

int ma_probe_primitives(int n, const double *cons, double *prim, int arith, int device) {
  if (n < 0 || !cons || !prim) return ma_set_error(MA_ERR_INVALID, "probe: bad argument");
  MA_RC(check_device(device));
  DevBuf a, f;
  const size_t N = (size_t)n;
  MA_RC(a.up(cons, 5 * N));
  MA_RC(f.up(nullptr, 5 * N));
  MA_CUDA_TRY(arith == MA_ARITH_STRICT ? ma_strict::probe_primitives(n, a.p, f.p, 0)
                                       : ma_fast::probe_primitives(n, a.p, f.p, 0));
  return f.down(prim, 5 * N);
}

int ma_probe_venkat(int n, const double *dumax, const double *dumin, const double *du, const double *deltax3,
                    double *phi, int arith, int device) {
  if (n < 0 || !dumax || !dumin || !du || !deltax3 || !phi) return ma_set_error(MA_ERR_INVALID, "probe: bad argument");
  MA_RC(check_device(device));
  DevBuf a, b, c, d, f;
  const size_t N = (size_t)n;
  MA_RC(a.up(dumax, N));
  MA_RC(b.up(dumin, N));
  MA_RC(c.up(du, N));
  MA_RC(d.up(deltax3, N));
  MA_RC(f.up(nullptr, N));
  MA_CUDA_TRY(arith == MA_ARITH_STRICT ? ma_strict::probe_venkat(n, a.p, b.p, c.p, d.p, f.p, 0)
                                       : ma_fast::probe_venkat(n, a.p, b.p, c.p, d.p, f.p, 0));
  return f.down(phi, N);
}

int ma_probe_vanalbada(int n, const double *dumax, const double *dumin, const double *du, double *phi, int arith,
                       int device) {
  if (n < 0 || !dumax || !dumin || !du || !phi) return ma_set_error(MA_ERR_INVALID, "probe: bad argument");
  MA_RC(check_device(device));
  DevBuf a, b, c, f;
  const size_t N = (size_t)n;
  MA_RC(a.up(dumax, N));
  MA_RC(b.up(dumin, N));
  MA_RC(c.up(du, N));
  MA_RC(f.up(nullptr, N));
  MA_CUDA_TRY(arith == MA_ARITH_STRICT ? ma_strict::probe_vanalbada(n, a.p, b.p, c.p, f.p, 0)
                                       : ma_fast::probe_vanalbada(n, a.p, b.p, c.p, f.p, 0));
  return f.down(phi, N);
}

}  // extern "C"
