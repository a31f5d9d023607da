// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Stand-in for the nine MPI functions the miniAero reference uses (Main.C:61-88, Parallel3DMesh.C:49-50,
// 316-419, CopyGhost.C:46-72, ElementTopoHexa8.C:88, TimeSolverExplicitRK4.h:285-286), so that the
// UNMODIFIED reference sources can be built with -DWITH_MPI=1 and run as N cooperating processes in a
// container that has no MPI.  oracle/refrun.py starts the N processes with
//     MINIAERO_MPI_RANK, MINIAERO_MPI_SIZE, MINIAERO_MPI_DIR (a private scratch directory).
// Messages are files: a send writes <dir>/m_<src>_<dst>_<tag>_<seq> (tmp + rename, so it appears
// atomically); the matching receive — same (src, dst, tag), same sequence number, which is MPI's
// non-overtaking rule — polls for it in MPI_Waitall.  Slow, but exact, and only used on test-sized meshes.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

typedef int MPI_Comm;
typedef int MPI_Datatype;
struct MPI_Status {
  int MPI_SOURCE, MPI_TAG, MPI_ERROR;
};
#define MPI_COMM_WORLD 0
#define MPI_INT 4
#define MPI_DOUBLE 8
#define MPI_SUCCESS 0
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)

struct MPI_Request {
  int kind;  // 0 none, 1 send (already complete), 2 recv
  void *buf;
  size_t bytes;
  int src, tag;
  long seq;
};

namespace mpi_standin {
struct State {
  int rank = 0, size = 1;
  std::string dir;
  std::map<std::pair<int, int>, long> send_seq, recv_seq;  // (peer, tag) -> next sequence number
  long allgather_seq = 0;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
};
inline State &state() {
  static State s;
  return s;
}
inline std::string msg_path(int src, int dst, int tag, long seq) {
  char b[256];
  std::snprintf(b, sizeof(b), "/m_%d_%d_%d_%ld", src, dst, tag, seq);
  return state().dir + b;
}
inline void write_file(const std::string &path, const void *p, size_t bytes) {
  const std::string tmp = path + ".tmp";
  FILE *f = std::fopen(tmp.c_str(), "wb");
  if (!f) {
    std::perror(tmp.c_str());
    std::abort();
  }
  if (bytes) std::fwrite(p, 1, bytes, f);
  std::fclose(f);
  std::rename(tmp.c_str(), path.c_str());
}
inline void read_file_blocking(const std::string &path, void *p, size_t bytes, bool unlink_after) {
  for (long spins = 0;; ++spins) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (f) {
      const size_t got = bytes ? std::fread(p, 1, bytes, f) : 0;
      std::fclose(f);
      if (got != bytes) {
        std::fprintf(stderr, "mpi stand-in: %s holds %zu bytes, expected %zu\n", path.c_str(), got, bytes);
        std::abort();
      }
      if (unlink_after) std::remove(path.c_str());
      return;
    }
    if (spins > 3000000) {  // ~10 minutes
      std::fprintf(stderr, "mpi stand-in: rank %d timed out waiting for %s\n", state().rank, path.c_str());
      std::abort();
    }
    std::this_thread::sleep_for(std::chrono::microseconds(spins < 200 ? 20 : 200));
  }
}
}  // namespace mpi_standin

inline int MPI_Init(int *, char ***) {
  mpi_standin::State &s = mpi_standin::state();
  const char *r = std::getenv("MINIAERO_MPI_RANK"), *n = std::getenv("MINIAERO_MPI_SIZE"), *d = std::getenv("MINIAERO_MPI_DIR");
  s.rank = r ? std::atoi(r) : 0;
  s.size = n ? std::atoi(n) : 1;
  s.dir = d ? d : ".";
  return MPI_SUCCESS;
}
inline int MPI_Finalize() { return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm, int *n) {
  *n = mpi_standin::state().size;
  return MPI_SUCCESS;
}
inline int MPI_Comm_rank(MPI_Comm, int *r) {
  *r = mpi_standin::state().rank;
  return MPI_SUCCESS;
}
inline double MPI_Wtime() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - mpi_standin::state().t0).count();
}
inline int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm, MPI_Request *req) {
  mpi_standin::State &s = mpi_standin::state();
  const long seq = s.send_seq[std::make_pair(dest, tag)]++;
  mpi_standin::write_file(mpi_standin::msg_path(s.rank, dest, tag, seq), buf, (size_t)count * type);
  *req = MPI_Request{1, nullptr, 0, dest, tag, seq};
  return MPI_SUCCESS;
}
inline int MPI_Irecv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm, MPI_Request *req) {
  mpi_standin::State &s = mpi_standin::state();
  const long seq = s.recv_seq[std::make_pair(source, tag)]++;
  *req = MPI_Request{2, buf, (size_t)count * type, source, tag, seq};
  return MPI_SUCCESS;
}
inline int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *) {
  mpi_standin::State &s = mpi_standin::state();
  for (int i = 0; i < n; ++i) {
    if (reqs[i].kind == 2)
      mpi_standin::read_file_blocking(mpi_standin::msg_path(reqs[i].src, s.rank, reqs[i].tag, reqs[i].seq), reqs[i].buf,
                                      reqs[i].bytes, true);
    reqs[i].kind = 0;
  }
  return MPI_SUCCESS;
}
inline int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf, int recvcount,
                         MPI_Datatype recvtype, MPI_Comm) {
  mpi_standin::State &s = mpi_standin::state();
  const long seq = s.allgather_seq++;
  char b[256];
  std::snprintf(b, sizeof(b), "/ag_%ld_%d", seq, s.rank);
  mpi_standin::write_file(s.dir + b, sendbuf, (size_t)sendcount * sendtype);
  for (int r = 0; r < s.size; ++r) {
    std::snprintf(b, sizeof(b), "/ag_%ld_%d", seq, r);
    mpi_standin::read_file_blocking(s.dir + b, (char *)recvbuf + (size_t)r * recvcount * recvtype,
                                    (size_t)recvcount * recvtype, false);  // every rank reads every file: left in place
  }
  return MPI_SUCCESS;
}
