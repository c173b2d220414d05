// comm.cuh -- one-process-per-GPU communication for the split-model path:
// NCCL send/recv halo exchange and small all-gathers for the Krylov / outer scalars.
//
// Replaces, for the path in scope, the reference's distributed backend:
//   MpiRouter (src/Distributed/MpiRouter.f90:238-343) halo of x / ibound for the interface cells
//   PETSc MPIAIJ VecScatter inside MatMult (PetscMatrix.F90:108-113, 455)
//   MPI_Allreduce of Krylov dots / convergence scalars (PetscConvergence.F90:126-158,
//   ParallelSolution.f90:54-235)
// NCCL is loaded with dlopen (libnccl.so.2: the copy torch already mapped, else the system
// one) so that the single-GPU library has no NCCL dependency.
#pragma once
#include "common.cuh"
#include "../../include/mf6gpu.h"

struct mf6gpu_comm {
  int nranks = 1, rank = 0;
  void *nccl = nullptr;  // ncclComm_t
  cudaStream_t stream = 0;
};

namespace mf6 {

// halo pattern of one solution on one rank
struct HaloPlan {
  mf6gpu_comm *comm = nullptr;
  int n_own = 0, n_halo = 0;
  std::vector<int> nbr_rank, send_ptr, recv_ptr;  // per neighbour; *_ptr have size nnbr+1
  DevBuf<int> send_idx;                           // [send_ptr.back()] owned rows (final numbering) to pack
  DevBuf<double> sendbuf;
  bool active() const { return comm != nullptr && comm->nranks > 1; }
  // vec[n_own + recv range of neighbour k] <- neighbour k's owned values
  void exchange(double *vec, cudaStream_t s);
};

// out[rank*count .. ] <- every rank's in[0..count)  (count doubles)
void comm_allgather(mf6gpu_comm *c, const double *in, double *out, size_t count, cudaStream_t s);

}  // namespace mf6
