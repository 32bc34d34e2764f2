// Peer memory over NVLink / NVSwitch: every rank of the box maps every other rank's slab (CUDA IPC) once, at process
// grid construction. After that the distributed product reads the operand tiles of the other ranks IN PLACE (the
// numeric kernel's bulk copies take peer addresses), and the only thing that still "communicates" per product is one
// small kernel, peer_exchange: a barrier across the ranks that also hands every rank 64 bytes of every other rank
// (descriptors of freshly written tile forms, partial norms, ...). No NCCL call and no copy on the product path.
//
// This replaces, for column-split grids (1 x C x 1), the reference's panel gathers
// (Source/Fortran/comm_includes/ReduceAndComposeMatrix*.f90, MatrixReduceModule.F90:89-286) and the tile halo
// exchange of round 1 (ncclSend/ncclRecv); NCCL remains for bootstrap, the general R x C x S grids and as fallback.
#pragma once
#include "comm.h"
#include "device.cuh"

namespace ntb {

struct PeerPayload { unsigned long long w[8]; };  // 64 bytes per rank and exchange

struct PeerSpace {
  bool ok = false;                               // every rank has mapped every slab
  int n = 1, me = 0;
  size_t slab_bytes = 0;
  unsigned char* base[PEER_MAX] = {};            // this process's mapping of rank p's slab (base[me]: own)
  unsigned long long epoch = 0;                  // exchanges enqueued so far (identical on every rank: SPMD)
  unsigned long long exchanges = 0;
};
PeerSpace& peer();

// Collective over the world communicator (all ranks of the job, one box). Returns peer().ok: false when the GPUs
// cannot map each other, NTB_P2P=0, or there are more than PEER_MAX ranks - then the callers keep to NCCL.
bool peer_setup(CommHandle* world);
void peer_teardown();

// Barrier across all ranks + all-to-all of one payload, enqueued on the library stream (one kernel, no host wait).
// When `out` is given, out[0..n) holds the payload of every rank after the next stream_sync().
// dev8 / n8 / dev_i32_w7: optional device addresses whose contents replace words 0..n8-1 (8-byte values) / word 7 (a
// 32-bit integer, widened) of `mine` when the kernel runs - values that earlier kernels of the stream produce: partial
// norms, a count.
void peer_exchange(const PeerPayload& mine, PeerPayload* out, const void* dev8 = nullptr, int n8 = 0,
                   const int* dev_i32_w7 = nullptr);
inline void peer_barrier() { PeerPayload z{}; peer_exchange(z, nullptr); }

}  // namespace ntb
