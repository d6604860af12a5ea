// Patch-gradient exchange of the data-parallel attack: a thin C-ABI wrapper over NCCL.
//
// Replaces torch DDP's reducer on the ONE trainable tensor of the reference (UADA_ddp.py:146,206: every backward()
// all-reduces patch.grad, 30 KB at 50x50).  The collective is launched on the engine's compute stream, so it can be
// recorded into the attack step's CUDA graph between the front-end backward and the patch update.
//
// libnccl.so.2 is resolved with dlopen at first use (inside a PyTorch process this is the copy torch already mapped;
// otherwise the system one), so libvla_b200.so itself has no link-time NCCL dependency and loads on machines without it.
#include <dlfcn.h>
#include <string.h>

#include "../../include/vla_b200.h"
#include "common.cuh"

namespace {

// the part of nccl.h this file needs (stable since NCCL 2.0)
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
enum { kNcclFloat32 = 7, kNcclSum = 0 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  bool tried = false;
} g_nccl;

int load_nccl() {
  if (g_nccl.handle) return 0;
  VLA_REQUIRE(!g_nccl.tried, "NCCL is not available (libnccl.so.2 could not be loaded)");
  g_nccl.tried = true;
  const char* names[] = {getenv("VLA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n || !*n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  VLA_REQUIRE(h != nullptr, "dlopen(libnccl.so.2) failed: %s", dlerror());
#define SYM(field, name)                                              \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name)); \
  VLA_REQUIRE(g_nccl.field != nullptr, "libnccl has no symbol %s", name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(GetErrorString, "ncclGetErrorString");
  SYM(GetVersion, "ncclGetVersion");
#undef SYM
  g_nccl.handle = h;
  return 0;
}

#define VLA_CHECK_NCCL(expr)                                                                   \
  do {                                                                                         \
    int _r = (expr);                                                                           \
    if (_r != 0) {                                                                             \
      vla_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
      return 3;                                                                                \
    }                                                                                          \
  } while (0)

}  // namespace

struct vla_comm {
  NcclComm comm = nullptr;
  int rank = 0, world = 1;
};

static_assert(sizeof(NcclUniqueId) == VLA_COMM_ID_BYTES, "unique id size");

extern "C" int vla_comm_unique_id(void* id_host) {
  VLA_REQUIRE(id_host != nullptr, "vla_comm_unique_id: null argument");
  if (int rc = load_nccl()) return rc;
  NcclUniqueId id;
  VLA_CHECK_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id_host, &id, sizeof(id));
  return 0;
}

extern "C" int vla_comm_create(const void* id_host, int rank, int world, vla_comm** out) {
  VLA_REQUIRE(id_host && out, "vla_comm_create: null argument");
  VLA_REQUIRE(world >= 1 && rank >= 0 && rank < world, "vla_comm_create: bad rank %d of %d", rank, world);
  if (int rc = load_nccl()) return rc;
  NcclUniqueId id;
  memcpy(&id, id_host, sizeof(id));
  vla_comm* c = new vla_comm();
  c->rank = rank;
  c->world = world;
  const int r = g_nccl.CommInitRank(&c->comm, world, id, rank);   // on the calling thread's current device
  if (r != 0) {
    vla_set_error("ncclCommInitRank(rank %d of %d) -> %s", rank, world, g_nccl.GetErrorString(r));
    delete c;
    return 3;
  }
  *out = c;
  return 0;
}

extern "C" void vla_comm_destroy(vla_comm* c) {
  if (!c) return;
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  delete c;
}

extern "C" int vla_comm_world(const vla_comm* c) { return c ? c->world : 1; }

extern "C" int vla_allreduce_patch_grad(vla_comm* c, float* dpatch, int n, void* stream) {
  VLA_REQUIRE(c && c->comm && dpatch && n > 0, "vla_allreduce_patch_grad: null / empty argument");
  VLA_CHECK_NCCL(g_nccl.AllReduce(dpatch, dpatch, static_cast<size_t>(n), kNcclFloat32, kNcclSum, c->comm,
                                  static_cast<cudaStream_t>(stream)));
  return 0;
}
