// nccl_dl.h - NCCL reached through dlopen so that libsvae.so has no link-time dependency on it (single-GPU use needs
// no NCCL at all).  The library used is the one already loaded in the process (PyTorch bundles libnccl.so.2) or the
// path the caller passes.
#pragma once
#include <stddef.h>

#include <string>

struct NcclApi {
  int (*GetUniqueId)(void* id);
  int (*CommInitRank)(void** comm, int nranks, const char* id128, int rank);
  int (*CommDestroy)(void* comm);
  int (*AllReduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, void* stream);
  const char* (*GetErrorString)(int);
};
NcclApi* nccl_load(const char* path_or_null);
const std::string& nccl_load_error();
