// Host driver: the reference's Main.C shape (Main.C:53-152) on top of the C ABI.
//
//   miniaero [--input FILE] [--arith fast|strict] [--limiter venkat|vanalbada] [--tile a,b,c] [--precision N] [--yaml DIR] [--no-yaml]
//
// reads ./miniaero.inp (Options.h:86), generates this rank's block of the hex mesh on the host, hands it to the
// solver (the drop-in for TimeSolverExplicitRK4 at Main.C:139-141), optionally writes results.<rank>
// (TimeSolverExplicitRK4.h:514-538), and finishes with the Mantevo YAML report.
//
// Ranks: one process per GPU, rank / size taken from RANK / WORLD_SIZE (or OMPI_COMM_WORLD_* / PMI_*), device from
// LOCAL_RANK.  The NCCL unique id travels through a file (MINIAERO_RENDEZVOUS, default
// /tmp/miniaero_rdv.<MASTER_PORT or 0>.<launcher pid>[.<run id>]): rank 0 writes it, the others poll, rank 0 removes
// it — the only bootstrap the run needs, standing in for MPI_Init (Main.C:61-63).
#include <fcntl.h>
#include <signal.h>
#include <unistd.h>

#include <cctype>
#include <random>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "miniaero_b200.h"

namespace {

using Clock = std::chrono::steady_clock;
double seconds_since(Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }

int env_int(std::initializer_list<const char *> names, int fallback) {
  for (const char *n : names)
    if (const char *v = getenv(n)) return atoi(v);
  return fallback;
}

[[noreturn]] void die(const char *what) {
  fprintf(stderr, "miniaero: %s: %s\n", what, ma_last_error());
  exit(1);
}

// Rank 0 publishes the NCCL id in a file, the other ranks wait for it.  Two things keep a launch from reading what an
// earlier (possibly aborted) launch left behind under the same name:
//  * the file name carries a per-launch token (MINIAERO_RENDEZVOUS overrides it): the launcher's run id when it exports
//    one, and the parent process id — every rank of one torchrun / mpirun launch on a node is a child of the same agent;
//  * a handshake: rank r > 0 first announces its process id in <file>.hello.<r>; rank 0 waits for an announcement from
//    a LIVE process per rank, then publishes id + the process ids it saw; rank r accepts the file only if it names r's
//    own process id.  A stale id file names dead processes and is ignored, a stale announcement is ignored by rank 0.
// Files are created exclusively under a temporary name (O_EXCL | O_NOFOLLOW: no symlink games in /tmp) and renamed
// into place; rank 0 removes the id file once its communicator exists (ncclCommInitRank returns only after every rank
// has joined, i.e. has read the id), every other rank removes its announcement once it holds the id.
std::string rendezvous_path() {
  if (const char *p = getenv("MINIAERO_RENDEZVOUS")) return p;
  const char *port = getenv("MASTER_PORT");
  const char *run = getenv("TORCHELASTIC_RUN_ID");
  std::string path = std::string("/tmp/miniaero_rdv.") + (port ? port : "0") + "." + std::to_string((long)getppid());
  if (run && *run) {
    path += ".";
    for (const char *c = run; *c; ++c) path += (isalnum((unsigned char)*c) ? *c : '_');
  }
  return path;
}
std::string g_rendezvous_file;  // rank 0: the file to remove (after the communicator exists, or at exit)
void remove_rendezvous_file() {
  if (!g_rendezvous_file.empty()) unlink(g_rendezvous_file.c_str());
  g_rendezvous_file.clear();
}
bool write_file_atomically(const std::string &path, const void *buf, size_t n) {
  const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
  unlink(tmp.c_str());
  const int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW, 0600);
  if (fd < 0) return false;
  const bool ok = write(fd, buf, n) == (ssize_t)n;
  close(fd);
  if (!ok || rename(tmp.c_str(), path.c_str()) != 0) {
    unlink(tmp.c_str());
    return false;
  }
  return true;
}
bool read_file(const std::string &path, void *buf, size_t n) {
  const int fd = open(path.c_str(), O_RDONLY | O_NOFOLLOW);
  if (fd < 0) return false;
  const ssize_t got = read(fd, buf, n);
  close(fd);
  return got == (ssize_t)n;
}
bool exchange_bytes(int rank, int nranks, unsigned char *buf, size_t n, const std::string &path) {
  const int kTries = 6000;  // x 100 ms = 10 minutes: rank 0 may still be generating its mesh
  std::vector<unsigned char> rec(n + sizeof(long long) * (size_t)nranks);
  long long *pids = reinterpret_cast<long long *>(rec.data() + n);
  if (rank == 0) {
    pids[0] = (long long)getpid();
    for (int r = 1; r < nranks; ++r) {
      const std::string hello = path + ".hello." + std::to_string(r);
      long long p = 0;
      int tries = 0;
      while (!(read_file(hello, &p, sizeof(p)) && p > 0 && kill((pid_t)p, 0) == 0)) {
        if (++tries > kTries) return false;
        std::this_thread::sleep_for(std::chrono::milliseconds(100));
      }
      pids[r] = p;
    }
    memcpy(rec.data(), buf, n);
    if (!write_file_atomically(path, rec.data(), rec.size())) return false;
    g_rendezvous_file = path;
    atexit(remove_rendezvous_file);
    return true;
  }
  const std::string hello = path + ".hello." + std::to_string(rank);
  const long long me = (long long)getpid();
  if (!write_file_atomically(hello, &me, sizeof(me))) return false;
  bool ok = false;
  for (int tries = 0; tries < kTries && !ok; ++tries) {
    ok = read_file(path, rec.data(), rec.size()) && pids[rank] == me;
    if (!ok) std::this_thread::sleep_for(std::chrono::milliseconds(100));
  }
  unlink(hello.c_str());
  if (ok) memcpy(buf, rec.data(), n);
  return ok;
}
bool exchange_id(int rank, int nranks, unsigned char id[MA_COMM_ID_BYTES]) {
  if (rank == 0 && ma_comm_get_unique_id(id)) return false;
  return exchange_bytes(rank, nranks, id, MA_COMM_ID_BYTES, rendezvous_path());
}

}  // namespace

int main(int argc, char **argv) {
  const auto t_start = Clock::now();
  std::string input = "miniaero.inp", yaml_dir = ".";
  int arith = MA_ARITH_FAST, precision = 0, tile[3] = {0, 0, 0}, limiter = MA_LIMITER_VENKAT;
  bool yaml = true, rdv_selftest = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> const char * {
      if (i + 1 >= argc) {
        fprintf(stderr, "miniaero: %s needs a value\n", a.c_str());
        exit(2);
      }
      return argv[++i];
    };
    if (a == "--input") input = next();
    else if (a == "--arith") arith = !strcmp(next(), "strict") ? MA_ARITH_STRICT : MA_ARITH_FAST;
    else if (a == "--limiter") limiter = !strcmp(next(), "vanalbada") ? MA_LIMITER_VANALBADA : MA_LIMITER_VENKAT;
    else if (a == "--tile") sscanf(next(), "%d,%d,%d", &tile[0], &tile[1], &tile[2]);
    else if (a == "--precision") precision = atoi(next());
    else if (a == "--yaml") yaml_dir = next();
    else if (a == "--no-yaml") yaml = false;
    else if (a == "--rendezvous-selftest") rdv_selftest = true;
    else {
      fprintf(stderr, "usage: miniaero [--input FILE] [--arith fast|strict] [--limiter venkat|vanalbada] [--tile a,b,c] [--precision N] [--yaml DIR] [--no-yaml]\n");
      return a == "--help" || a == "-h" ? 0 : 2;
    }
  }
  const int num_procs = env_int({"WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE"}, 1);
  const int my_id = env_int({"RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK"}, 0);
  const int device = env_int({"LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK"}, 0);

  if (rdv_selftest) {
    // the rendezvous alone (no GPU, no NCCL): rank 0 publishes a fresh random token, every rank prints the one it got
    unsigned char tok[MA_COMM_ID_BYTES];
    if (my_id == 0) {
      std::random_device rd;
      for (unsigned char &b : tok) b = (unsigned char)rd();
    }
    if (!exchange_bytes(my_id, num_procs, tok, sizeof(tok), rendezvous_path())) {
      fprintf(stderr, "miniaero: rank %d: rendezvous failed\n", my_id);
      return 1;
    }
    unsigned long long h = 1469598103934665603ull;
    for (unsigned char b : tok) h = (h ^ b) * 1099511628211ull;
    fprintf(stdout, "rendezvous rank %d token %016llx\n", my_id, h);
    fflush(stdout);
    if (my_id == 0) {  // stand-in for ncclCommInitRank's rendezvous: wait until the other ranks say they have read it
      const std::string ack = rendezvous_path() + ".ack";
      for (int r = 1; r < num_procs; ++r) {
        bool seen = false;
        for (int tries = 0; tries < 600 && !seen; ++tries) {
          seen = access((ack + std::to_string(r)).c_str(), F_OK) == 0;
          if (!seen) std::this_thread::sleep_for(std::chrono::milliseconds(50));
        }
        unlink((ack + std::to_string(r)).c_str());
      }
      remove_rendezvous_file();
    } else {
      const std::string ack = rendezvous_path() + ".ack" + std::to_string(my_id);
      const int fd = open(ack.c_str(), O_WRONLY | O_CREAT | O_NOFOLLOW, 0600);
      if (fd >= 0) close(fd);
    }
    return 0;
  }

  ma_options opt;
  if (ma_options_read(input.c_str(), &opt)) die("reading the options file");  // Main.C:73-74

  // ---- setup (Main.C:96-129).  The reference-format host mesh is only needed to write results.<rank>; without
  // output the block's layout is built straight from (i, j, k) and its geometry on the device
  // (ma_solver_create_structured): the solver constructor below is then the whole set-up.
  const auto t_setup = Clock::now();
  ma_mesh_storage *mesh = nullptr;
  int nproc[3], block[3], nlocal[3], offset[3];
  if (ma_block_decomposition(&opt, my_id, num_procs, nproc, block, nlocal, offset)) die("block decomposition");
  if (opt.output_results && ma_mesh_generate(&opt, my_id, num_procs, &mesh)) die("mesh generation");
  // ---- communicator and solver (the device layout is part of the set-up)
  ma_comm *comm = nullptr;
  if (num_procs > 1) {
    unsigned char id[MA_COMM_ID_BYTES];
    if (!exchange_id(my_id, num_procs, id)) {
      fprintf(stderr, "miniaero: rank %d could not obtain the NCCL id (%s)\n", my_id, ma_last_error());
      return 1;
    }
    if (ma_comm_create(id, num_procs, my_id, device, &comm)) die("communicator");
    remove_rendezvous_file();  // every rank has read it by now
  }
  ma_solver_config cfg;
  ma_solver_config_default(&cfg);
  cfg.device = device;
  cfg.arith = arith;
  cfg.limiter = limiter;
  cfg.comm = comm;
  for (int d = 0; d < 3; ++d) cfg.tile_dims[d] = tile[d];
  ma_solver *solver = nullptr;
  if (mesh ? ma_solver_create(ma_mesh_view(mesh), &opt, &cfg, &solver)
           : ma_solver_create_structured(&opt, my_id, num_procs, &cfg, &solver))
    die("solver");
  const double setup_s = seconds_since(t_setup);
  if (my_id == 0) fprintf(stdout, "\n ... Setup time: %8.2f seconds ...\n", setup_s);
  int owned_cells = 0;
  ma_solver_num_cells(solver, &owned_cells, nullptr);
  // ---- run on the device (Main.C:131-150)
  const auto t_run = Clock::now();
  if (ma_solver_solve(solver)) die("Solve");  // prints the progress lines and "Device Run time"
  ma_timing tm;
  ma_solver_get_timing(solver, &tm);
  if (my_id == 0)
    fprintf(stdout, " ... %lld cells x %lld steps: %.4e cell-updates/s on this rank (device time %.3f s) ...\n",
            (long long)owned_cells, tm.steps,
            tm.step_seconds > 0 ? (double)tm.cell_updates / tm.step_seconds : 0.0, tm.step_seconds);

  if (opt.output_results) {  // TimeSolverExplicitRK4.h:514-538
    const ma_mesh *mv = ma_mesh_view(mesh);
    std::vector<double> sol((size_t)mv->num_owned_cells * 5);
    if (ma_solver_get_solution(solver, sol.data())) die("reading the solution back");
    const std::string name = "results." + std::to_string(my_id);
    if (ma_write_results(name.c_str(), mv, sol.data(), precision)) die("writing results");
  }
  const double run_s = seconds_since(t_run);
  const double total_s = seconds_since(t_start);

  if (yaml && my_id == 0) {
    ma_report rep;
    memset(&rep, 0, sizeof(rep));
    rep.options = &opt;
    rep.num_ranks = num_procs;
    for (int d = 0; d < 3; ++d) rep.blocks[d] = nproc[d];
    rep.global_cells = (long long)opt.nx * opt.ny * opt.nz;
    rep.timing = &tm;
    rep.setup_seconds = setup_s, rep.run_seconds = run_s, rep.total_seconds = total_s;
    if (const char *pk = getenv("MINIAERO_HBM_PEAK_GBS")) rep.hbm_peak_gbs = atof(pk);
    char path[512];
    if (ma_write_yaml_report(&rep, yaml_dir.c_str(), path, sizeof(path)))
      fprintf(stderr, "miniaero: YAML report not written: %s\n", ma_last_error());
    else
      fprintf(stdout, " ... report: %s ...\n", path);
  }
  ma_solver_destroy(solver);
  if (comm) ma_comm_destroy(comm);
  if (mesh) ma_mesh_free(mesh);
  if (my_id == 0) fprintf(stdout, "\n ... Total elapsed time: %8.2f seconds ...\n", seconds_since(t_start));  // Main.C:90-92
  return 0;
}
