#include "comm.h"
#include <fcntl.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <ctime>
#include "device.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include <string>
#include <unistd.h>

namespace ntb {

struct CommHandle {
  ncclComm_t comm = nullptr;
  int size = 1;
  int rank = 0;
};

namespace {
struct NcclApi {
  void* lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommSplit) CommSplit = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
} g_nccl;

#define NCCL_CHECK(expr)                                                              \
  do {                                                                                \
    ncclResult_t _r = (expr);                                                         \
    if (_r != ncclSuccess) {                                                          \
      std::fprintf(stderr, "[ntpoly_b200] NCCL error %s: %s\n", #expr,                \
                   g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?");          \
      NTB_FATAL("NCCL call failed");                                                  \
    }                                                                                 \
  } while (0)

void load_nccl() {
  if (g_nccl.lib) return;
  // Prefer a libnccl already mapped into the process (e.g. the one torch bundles).
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) {
    if (const char* p = std::getenv("NTB_NCCL_LIB")) g_nccl.lib = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
  }
  for (const char* n : names) {
    if (g_nccl.lib) break;
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
  }
  NTB_CHECK(g_nccl.lib != nullptr, "multi-GPU run requested but libnccl.so.2 could not be loaded");
#define LOAD(sym)                                                                     \
  g_nccl.sym = reinterpret_cast<decltype(g_nccl.sym)>(dlsym(g_nccl.lib, "nccl" #sym)); \
  NTB_CHECK(g_nccl.sym != nullptr, "missing NCCL symbol nccl" #sym)
  LOAD(GetUniqueId); LOAD(CommInitRank); LOAD(CommSplit); LOAD(CommDestroy); LOAD(AllReduce);
  LOAD(AllGather); LOAD(Broadcast); LOAD(Send); LOAD(Recv); LOAD(GroupStart); LOAD(GroupEnd);
  LOAD(GetErrorString);
#undef LOAD
}

World g_world;
}  // namespace

World& world() { return g_world; }

void world_get_unique_id(void* out128) {
  load_nccl();
  ncclUniqueId id;
  NCCL_CHECK(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  std::memcpy(out128, &id, 128);
}

void world_init_explicit(int rank, int size, const void* id128) {
  if (g_world.inited) {
    NTB_CHECK(g_world.rank == rank && g_world.size == size, "world already initialised differently");
    return;
  }
  ensure_init();
  g_world.rank = rank;
  g_world.size = size;
  if (size > 1) {
    load_nccl();
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    auto* h = new CommHandle();
    NCCL_CHECK(g_nccl.CommInitRank(&h->comm, size, id, rank));
    h->size = size;
    h->rank = rank;
    g_world.comm = h;
  }
  g_world.inited = true;
}

void world_init_from_env() {
  if (g_world.inited) return;
  const char* r = std::getenv("RANK");
  const char* w = std::getenv("WORLD_SIZE");
  int rank = r ? std::atoi(r) : 0;
  int size = w ? std::atoi(w) : 1;
  if (size <= 1) { world_init_explicit(0, 1, nullptr); return; }
  // file rendezvous for hosts without a Python side (plain C / Fortran callers)
  const char* port = std::getenv("MASTER_PORT");
  std::string path = std::string("/tmp/ntpoly_b200_ncclid_") + (port ? port : "0");
  if (const char* p = std::getenv("NTB_RENDEZVOUS_FILE")) path = p;
  unsigned char id[128];
  // A file left behind by a run that crashed must never be taken for this launch's id (a reader would then hang in
  // ncclCommInitRank): rank 0 removes whatever is there and creates the file exclusively (0600, no symlink followed),
  // readers accept only a file written at most two minutes before their own start (the ranks of one launch start within seconds of each other; what a crashed run left behind is older, or removed by rank 0 first).
  const time_t started = time(nullptr);
  if (rank == 0) {
    world_get_unique_id(id);
    ::unlink(path.c_str());
    std::string tmp = path + ".tmp." + std::to_string((long long)getpid());
    ::unlink(tmp.c_str());
    const int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW, 0600);
    NTB_CHECK(fd >= 0, "cannot create the NCCL rendezvous file");
    NTB_CHECK(::write(fd, id, 128) == 128, "cannot write the NCCL rendezvous file");
    ::close(fd);
    NTB_CHECK(std::rename(tmp.c_str(), path.c_str()) == 0, "cannot publish the NCCL rendezvous file");
  } else {
    for (int tries = 0;; ++tries) {
      struct stat sb;
      const int fd = ::open(path.c_str(), O_RDONLY | O_NOFOLLOW);
      if (fd >= 0) {
        const bool fresh = ::fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && sb.st_mtime + 120 >= started;
        const ssize_t n = fresh ? ::read(fd, id, 128) : 0;
        ::close(fd);
        if (n == 128) break;
      }
      NTB_CHECK(tries < 6000, "timed out waiting for the NCCL rendezvous file");
      usleep(10000);
    }
  }
  world_init_explicit(rank, size, id);
  if (rank == 0) { usleep(200000); std::remove(path.c_str()); }
}

CommHandle* comm_split(CommHandle* parent, int color, int key) {
  if (!parent) return nullptr;
  auto* h = new CommHandle();
  NCCL_CHECK(g_nccl.CommSplit(parent->comm, color, key, &h->comm, nullptr));
  // size/rank of the child follow from colour counting done by the caller
  h->size = -1;
  h->rank = key;
  return h;
}
void comm_free(CommHandle* c) {
  if (!c) return;
  if (c->comm) g_nccl.CommDestroy(c->comm);
  delete c;
}
int comm_size(const CommHandle* c) { return c ? c->size : 1; }
int comm_rank(const CommHandle* c) { return c ? c->rank : 0; }

// exposed for ProcessGrid to finalise size/rank after a split
void comm_set_shape(CommHandle* c, int size, int rank) { if (c) { c->size = size; c->rank = rank; } }

void comm_allreduce_f64(CommHandle* c, double* d_buf, size_t count, RedOp op) {
  readback_flush();
  if (!c || c->size <= 1 || count == 0) return;
  ncclRedOp_t o = op == RedOp::Sum ? ncclSum : (op == RedOp::Max ? ncclMax : ncclMin);
  NCCL_CHECK(g_nccl.AllReduce(d_buf, d_buf, count, ncclDouble, o, c->comm, rt().stream));
}
void comm_allgather_bytes(CommHandle* c, const void* d_send, void* d_recv, size_t bytes) {
  readback_flush();
  if (!c || c->size <= 1) {
    if (d_send != d_recv) CUDA_CHECK(cudaMemcpyAsync(d_recv, d_send, bytes, cudaMemcpyDeviceToDevice, rt().stream));
    return;
  }
  NCCL_CHECK(g_nccl.AllGather(d_send, d_recv, bytes, ncclChar, c->comm, rt().stream));
}
void comm_group_start() {
  readback_flush(); if (g_nccl.lib) NCCL_CHECK(g_nccl.GroupStart()); }
void comm_group_end() { if (g_nccl.lib) NCCL_CHECK(g_nccl.GroupEnd()); }
void comm_broadcast_bytes(CommHandle* c, const void* d_send, void* d_recv, size_t bytes, int root) {
  readback_flush();
  if (!c || c->size <= 1) {
    if (d_send != d_recv && bytes) CUDA_CHECK(cudaMemcpyAsync(d_recv, d_send, bytes, cudaMemcpyDeviceToDevice, rt().stream));
    return;
  }
  if (bytes == 0) return;
  NCCL_CHECK(g_nccl.Broadcast(d_send, d_recv, bytes, ncclChar, root, c->comm, rt().stream));
}
void comm_send_bytes(CommHandle* c, const void* d_buf, size_t bytes, int peer) {
  readback_flush();
  NTB_CHECK(c != nullptr, "send on a null communicator");
  NCCL_CHECK(g_nccl.Send(d_buf, bytes, ncclChar, peer, c->comm, rt().stream));
}
void comm_recv_bytes(CommHandle* c, void* d_buf, size_t bytes, int peer) {
  readback_flush();
  NTB_CHECK(c != nullptr, "recv on a null communicator");
  NCCL_CHECK(g_nccl.Recv(d_buf, bytes, ncclChar, peer, c->comm, rt().stream));
}

}  // namespace ntb
