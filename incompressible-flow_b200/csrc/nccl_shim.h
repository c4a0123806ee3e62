// nccl_shim.h -- minimal run-time binding of NCCL (dlopen), so that libimpes_b200.so has no link-time
// dependency on a particular libnccl and loads on machines without a GPU (symbol-check tests).
// Only the calls the hot path needs: communicator bootstrap, scalar all-reduce (Krylov dot products,
// replaces MPI_Allreduce inside MINRESSolver(MPI_COMM_WORLD) [REF src/par_pressure_solver.cpp:99]),
// and plane send/recv (face-neighbour halo; replaces HYPRE's ParCSR matvec exchange).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstddef>

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclFloat64 = 8 };           // ncclDataType_t
enum { ncclSum = 0, ncclMax = 2 };  // ncclRedOp_t

struct NcclShim {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;

  bool load() {
    if (lib) return true;
    // RTLD_NOLOAD first: reuse the NCCL the process already has (torch's bundled 2.28.x)
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW);
    if (!lib) return false;
#define IMPES_SYM(name) *(void**)(&name) = dlsym(lib, "nccl" #name)
    IMPES_SYM(GetUniqueId); IMPES_SYM(CommInitRank); IMPES_SYM(CommDestroy); IMPES_SYM(AllReduce);
    IMPES_SYM(Send); IMPES_SYM(Recv); IMPES_SYM(GroupStart); IMPES_SYM(GroupEnd);
#undef IMPES_SYM
    return GetUniqueId && CommInitRank && CommDestroy && AllReduce && Send && Recv && GroupStart && GroupEnd;
  }
};
