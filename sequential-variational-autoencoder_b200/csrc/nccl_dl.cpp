// nccl_dl.cpp - see nccl_dl.h
#include "nccl_dl.h"

#include <dlfcn.h>
#include <string.h>

namespace {
struct UniqueId { char internal[128]; };
typedef int (*fn_uid)(UniqueId*);
typedef int (*fn_init)(void**, int, UniqueId, int);
typedef int (*fn_destroy)(void*);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, void*);
typedef const char* (*fn_errstr)(int);
fn_uid p_uid; fn_init p_init; fn_destroy p_destroy; fn_allreduce p_allreduce; fn_errstr p_errstr;
NcclApi g_api;
bool g_loaded = false;
std::string g_load_err;

int w_uid(void* id) { return p_uid(reinterpret_cast<UniqueId*>(id)); }
int w_init(void** comm, int n, const char* id128, int rank) {
  UniqueId id;
  memcpy(id.internal, id128, 128);
  return p_init(comm, n, id, rank);   // ncclUniqueId is passed by value in the NCCL ABI
}
int w_destroy(void* c) { return p_destroy(c); }
int w_allreduce(const void* s, void* r, size_t n, int dt, int op, void* c, void* st) { return p_allreduce(s, r, n, dt, op, c, st); }
const char* w_err(int e) { return p_errstr ? p_errstr(e) : "nccl error"; }
}  // namespace

const std::string& nccl_load_error() { return g_load_err; }

NcclApi* nccl_load(const char* path) {
  if (g_loaded) return &g_api;
  void* lib = nullptr;
  if (path && path[0]) lib = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // already in the process (torch)
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { g_load_err = dlerror() ? dlerror() : "dlopen(libnccl) failed"; return nullptr; }
  p_uid = (fn_uid)dlsym(lib, "ncclGetUniqueId");
  p_init = (fn_init)dlsym(lib, "ncclCommInitRank");
  p_destroy = (fn_destroy)dlsym(lib, "ncclCommDestroy");
  p_allreduce = (fn_allreduce)dlsym(lib, "ncclAllReduce");
  p_errstr = (fn_errstr)dlsym(lib, "ncclGetErrorString");
  if (!p_uid || !p_init || !p_destroy || !p_allreduce) { g_load_err = "libnccl is missing required symbols"; return nullptr; }
  g_api.GetUniqueId = w_uid; g_api.CommInitRank = w_init; g_api.CommDestroy = w_destroy;
  g_api.AllReduce = w_allreduce; g_api.GetErrorString = w_err;
  g_loaded = true;
  return &g_api;
}
