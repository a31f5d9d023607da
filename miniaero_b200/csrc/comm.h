// Internal interface of the NCCL halo communicator (comm.cpp).
#pragma once
#include <cuda_runtime.h>

struct ma_comm;

namespace ma {
int comm_rank(const ma_comm *c);
int comm_size(const ma_comm *c);
// One grouped send/recv round with every peer: `row` doubles per cell; buffers hold the peers' segments
// back to back in ascending peer rank.  Enqueued on `st`; returns MA_OK or sets the error text.
int comm_exchange(ma_comm *c, const double *sendbuf, double *recvbuf, int row, int npeers, const int *peer_rank,
                  const int *send_count, const int *recv_count, cudaStream_t st);
}  // namespace ma
