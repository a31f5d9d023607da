// Reference-side binding: the class a miniAero maintainer adds to call libminiaero_b200.so in place of
// TimeSolverExplicitRK4<Device> (TimeSolverExplicitRK4.h:160-539).  Same constructor and Solve() shape as the
// reference's class at its one call site (Main.C:139-141); it touches only the reference's own containers
// (MeshData.h:43-58, Faces.h:42-71, Cells.h:41-76, Options.h:47-57) and the C ABI of include/miniaero_b200.h.
//
// Compiled and run by the test suite against the UNMODIFIED reference sources (oracle/build_ref.sh builds
// oracle/_ref/miniAero.b200 = the reference's Main.C, mesh generator and this header, with the class name redirected
// by include/reference_binding/use_b200_solver.h; tests/test_reference_binding.py runs the reference's three serial
// integration tests through it on the GPU).
//
// The mesh is read through host mirrors and operator() — no assumption about the View layout of the backend the
// reference was built for (LayoutLeft on Kokkos::Cuda, LayoutRight on the host backends, ViewTypes.h:41-50).
#ifndef MINIAERO_B200_REFERENCE_BINDING_H_
#define MINIAERO_B200_REFERENCE_BINDING_H_

#include <Kokkos_Core.hpp>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "MeshData.h"
#include "Options.h"
#include "miniaero_b200.h"

template <class Device>
class TimeSolverB200 {
  // plain row-major host copies of one Faces<Device> (Faces.h:61-66): what ma_faces points at
  struct HostFaces {
    std::vector<double> xyz, n, t, b;
    std::vector<int> conn, slot;
    explicit HostFaces(const Faces<Device> &f) {
      const int nf = f.nfaces_;
      auto hx = Kokkos::create_mirror_view(f.coordinates_);
      auto hn = Kokkos::create_mirror_view(f.face_normal_);
      auto ht = Kokkos::create_mirror_view(f.face_tangent_);
      auto hb = Kokkos::create_mirror_view(f.face_binormal_);
      auto hc = Kokkos::create_mirror_view(f.face_cell_conn_);
      auto hs = Kokkos::create_mirror_view(f.cell_flux_index_);
      Kokkos::deep_copy(hx, f.coordinates_);
      Kokkos::deep_copy(hn, f.face_normal_);
      Kokkos::deep_copy(ht, f.face_tangent_);
      Kokkos::deep_copy(hb, f.face_binormal_);
      Kokkos::deep_copy(hc, f.face_cell_conn_);
      Kokkos::deep_copy(hs, f.cell_flux_index_);
      xyz.resize((size_t)3 * nf), n.resize((size_t)3 * nf), t.resize((size_t)3 * nf), b.resize((size_t)3 * nf);
      conn.resize((size_t)2 * nf), slot.resize((size_t)2 * nf);
      for (int i = 0; i < nf; ++i) {
        for (int d = 0; d < 3; ++d) {
          xyz[(size_t)3 * i + d] = hx(i, d);
          n[(size_t)3 * i + d] = hn(i, d);
          t[(size_t)3 * i + d] = ht(i, d);
          b[(size_t)3 * i + d] = hb(i, d);
        }
        for (int d = 0; d < 2; ++d) {
          conn[(size_t)2 * i + d] = hc(i, d);
          slot[(size_t)2 * i + d] = hs(i, d);
        }
      }
    }
    ma_faces view() const {
      ma_faces v;
      v.nfaces = (int)(conn.size() / 2);
      v.coordinates = xyz.data(), v.face_normal = n.data(), v.face_tangent = t.data(), v.face_binormal = b.data();
      v.face_cell_conn = conn.data(), v.cell_flux_index = slot.data();
      return v;
    }
  };

  [[noreturn]] static void fail(const char *what) {  // the reference has no error channel: report and stop
    fprintf(stderr, "TimeSolverB200: %s: %s\n", what, ma_last_error());
    exit(1);
  }

 public:
  // TimeSolverExplicitRK4(MeshData<Device>&, const Options&), TimeSolverExplicitRK4.h:180-196.  `comm`: the halo
  // communicator of a block-decomposed run (ma_comm_create with an id broadcast by the caller's MPI), else null.
  TimeSolverB200(struct MeshData<Device> &mesh_data, const Options &options, ma_comm *comm = nullptr, int my_rank = 0,
                 int num_ranks = 1)
      : internal_(mesh_data.internal_faces), my_rank_(my_rank) {
    opt_.problem_type = options.problem_type;
    opt_.lx = options.lx, opt_.ly = options.ly, opt_.lz = options.lz, opt_.angle = options.angle;
    opt_.nx = options.nx, opt_.ny = options.ny, opt_.nz = options.nz;
    opt_.ntimesteps = options.ntimesteps, opt_.dt = options.dt;
    opt_.output_results = options.output_results, opt_.output_frequency = options.output_frequency;
    opt_.second_order_space = options.second_order_space, opt_.viscous = options.viscous;

    const int ncells = mesh_data.num_owned_cells + mesh_data.num_ghosts;
    auto hxyz = Kokkos::create_mirror_view(mesh_data.mesh_cells.coordinates_);  // Cells.h:63
    auto hvol = Kokkos::create_mirror_view(mesh_data.mesh_cells.volumes_);      // Cells.h:64
    Kokkos::deep_copy(hxyz, mesh_data.mesh_cells.coordinates_);
    Kokkos::deep_copy(hvol, mesh_data.mesh_cells.volumes_);
    cell_xyz_.resize((size_t)3 * ncells), cell_vol_.resize((size_t)ncells);
    for (int c = 0; c < ncells; ++c) {
      for (int d = 0; d < 3; ++d) cell_xyz_[(size_t)3 * c + d] = hxyz(c, d);
      cell_vol_[c] = hvol(c);
    }
    mesh_ = ma_mesh();
    mesh_.num_owned_cells = mesh_data.num_owned_cells;  // MeshData.h:46
    mesh_.num_ghosts = mesh_data.num_ghosts;            // MeshData.h:45
    mesh_.cell_coordinates = cell_xyz_.data(), mesh_.cell_volumes = cell_vol_.data();
    mesh_.internal_faces = internal_.view();
    bc_.reserve(mesh_data.boundary_faces.size());
    for (auto &named : mesh_data.boundary_faces) {  // MeshData.h:57; the order is kept (Parallel3DMesh.h:382-396)
      if (mesh_.num_boundary_sets == MA_MAX_BC_SETS) fail("more boundary sets than MA_MAX_BC_SETS");
      const std::string &name = named.first;
      const int type = name == "Extrapolate" ? MA_BC_EXTRAPOLATE : name == "Tangent" ? MA_BC_TANGENT
                       : name == "Inflow"    ? MA_BC_INFLOW      : name == "NoSlip"  ? MA_BC_NOSLIP : -1;
      if (type < 0) fail(("unknown boundary set name '" + name + "'").c_str());
      bc_.emplace_back(named.second);
      mesh_.boundary_type[mesh_.num_boundary_sets] = type;
      mesh_.boundary_faces[mesh_.num_boundary_sets++] = bc_.back().view();
    }
    mesh_.num_ranks = num_ranks, mesh_.my_rank = my_rank;
    if (num_ranks > 1) {  // MeshData.h:50-55: per-rank counts on the host, id lists on the device
      send_count_ = mesh_data.sendCount, recv_count_ = mesh_data.recvCount;
      auto hs = Kokkos::create_mirror_view(mesh_data.send_local_ids);
      auto hr = Kokkos::create_mirror_view(mesh_data.recv_local_ids);
      Kokkos::deep_copy(hs, mesh_data.send_local_ids);
      Kokkos::deep_copy(hr, mesh_data.recv_local_ids);
      size_t ns = 0, nr = 0;
      for (int r = 0; r < num_ranks; ++r) ns += send_count_[r], nr += recv_count_[r];
      send_ids_.resize(ns), recv_ids_.resize(nr);
      for (size_t i = 0; i < ns; ++i) send_ids_[i] = hs((int)i);
      for (size_t i = 0; i < nr; ++i) recv_ids_[i] = hr((int)i);
      mesh_.send_count = send_count_.data(), mesh_.recv_count = recv_count_.data();
      mesh_.send_local_ids = send_ids_.data(), mesh_.recv_local_ids = recv_ids_.data();
    }
    ma_solver_config cfg;
    ma_solver_config_default(&cfg);
    cfg.comm = comm;
    if (const char *a = getenv("MINIAERO_B200_ARITH")) cfg.arith = (a[0] == 's') ? MA_ARITH_STRICT : MA_ARITH_FAST;
    if (ma_solver_create(&mesh_, &opt_, &cfg, &solver_) != MA_OK) fail("ma_solver_create");
  }

  // TimeSolverExplicitRK4::Solve(), TimeSolverExplicitRK4.h:207-539: initial conditions, ntimesteps RK4 steps with the
  // reference's progress lines, then results.<rank> when options.output_results is set (:514-538).
  void Solve() {
    if (ma_solver_solve(solver_) != MA_OK) fail("ma_solver_solve");
    if (opt_.output_results) {
      std::vector<double> sol((size_t)mesh_.num_owned_cells * 5);
      if (ma_solver_get_solution(solver_, sol.data()) != MA_OK) fail("ma_solver_get_solution");
      const std::string name = "results." + std::to_string(my_rank_);
      if (ma_write_results(name.c_str(), &mesh_, sol.data(), 6) != MA_OK) fail("ma_write_results");
    }
  }

  // conserved variables of the owned cells in the caller's cell order, [num_owned_cells][5]
  void solution(double *out) {
    if (ma_solver_get_solution(solver_, out) != MA_OK) fail("ma_solver_get_solution");
  }

  ~TimeSolverB200() { ma_solver_destroy(solver_); }
  TimeSolverB200(const TimeSolverB200 &) = delete;
  TimeSolverB200 &operator=(const TimeSolverB200 &) = delete;

 private:
  ma_options opt_;
  ma_mesh mesh_;
  HostFaces internal_;
  std::vector<HostFaces> bc_;
  std::vector<double> cell_xyz_, cell_vol_;
  std::vector<int> send_count_, recv_count_, send_ids_, recv_ids_;
  int my_rank_ = 0;
  ma_solver *solver_ = nullptr;
};

#endif  // MINIAERO_B200_REFERENCE_BINDING_H_
