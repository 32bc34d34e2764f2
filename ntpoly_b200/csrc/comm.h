// Process world + NCCL plumbing. One process per GPU; NCCL is loaded with
// dlopen so that a single-GPU run never needs it. Replaces the MPI communicators
// of the reference's ProcessGrid_t (Source/Fortran/ProcessGridModule.F90:15-56).
#pragma once
#include <cstddef>
#include <cstdint>

namespace ntb {

struct CommHandle;  // opaque (ncclComm_t + size/rank)

struct World {
  int rank = 0;
  int size = 1;
  bool inited = false;
  CommHandle* comm = nullptr;  // nullptr when size == 1
};
World& world();

// Explicit bootstrap (Python host passes a 128-byte ncclUniqueId it broadcast itself).
void world_init_explicit(int rank, int size, const void* nccl_unique_id_128);
// Environment bootstrap: RANK / WORLD_SIZE (+ a rendezvous file keyed on MASTER_PORT);
// single process if the variables are absent.
void world_init_from_env();
void world_get_unique_id(void* out128);

CommHandle* comm_split(CommHandle* parent, int color, int key);  // ncclCommSplit; nullptr if parent null
void comm_free(CommHandle* c);
int comm_size(const CommHandle* c);   // 1 for nullptr
int comm_rank(const CommHandle* c);   // 0 for nullptr
void comm_set_shape(CommHandle* c, int size, int rank);

enum class RedOp { Sum, Max, Min };
// all of these are no-ops on a nullptr / size-1 communicator; buffers are device pointers,
// work is enqueued on the library stream.
void comm_allreduce_f64(CommHandle* c, double* d_buf, size_t count, RedOp op);
void comm_allgather_bytes(CommHandle* c, const void* d_send, void* d_recv, size_t bytes_per_rank);
void comm_group_start();
void comm_group_end();
void comm_broadcast_bytes(CommHandle* c, const void* d_send, void* d_recv, size_t bytes, int root);
void comm_send_bytes(CommHandle* c, const void* d_buf, size_t bytes, int peer);
void comm_recv_bytes(CommHandle* c, void* d_buf, size_t bytes, int peer);

}  // namespace ntb
