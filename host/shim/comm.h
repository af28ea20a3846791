// Minimal process-group rendezvous for the one-process-per-GPU runs of the stand-alone driver
// ([Domain] parallel_mode = FFT_SLAB).  Stands where MOOSE hands Marlin an MPI communicator
// (DomainAction.C:164-189 host-name allgather, postprocessor gatherSum / gatherMin / gatherMax): ranks find each other
// through the torchrun-style environment (RANK, WORLD_SIZE, LOCAL_RANK, MASTER_ADDR, MASTER_PORT) and talk over TCP
// through rank 0.  Only small host-side messages travel here (CUDA IPC handles, postprocessor scalars); field data
// moves GPU to GPU over NVLink inside mrl_dist_*.
#pragma once
#include <cstddef>
#include <string>
#include <vector>

class Comm {
public:
  // world of one process when WORLD_SIZE is unset or 1
  static Comm &world();
  ~Comm();
  int rank() const { return _rank; }
  int size() const { return _size; }
  int localRank() const { return _local_rank; }
  // out: size() * bytes, rank order
  void allgather(const void *in, size_t bytes, void *out);
  enum Op { SUM, MIN, MAX };
  void allreduce(double *values, size_t n, Op op);
  void barrier();

private:
  Comm();
  void sendAll(int fd, const void *p, size_t n);
  void recvAll(int fd, void *p, size_t n);
  int _rank = 0, _size = 1, _local_rank = 0;
  int _hub = -1;                 // ranks > 0: socket to rank 0
  std::vector<int> _peers;       // rank 0: sockets to ranks 1..size-1 (index = rank)
};
