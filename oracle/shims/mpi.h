// Test-infrastructure shim: single-rank MPI stub (rank 0, size 1) for the reference's lmc/mc sources.
#pragma once
#include <cstdlib>
#include <cstring>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef long MPI_Aint;
typedef void(MPI_User_function)(void *, void *, int *, MPI_Datatype *);
#define MPI_COMM_WORLD 0
#define MPI_DOUBLE 8
#define MPI_INT 4
#define MPI_BYTE 1
#define MPI_SUM 0
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_SUCCESS 0
static inline int lmc_shim_type_size(MPI_Datatype t) { return t > 0 ? t : 1; }
static inline int MPI_Init_thread(int *, char ***, int, int *provided) { if (provided) *provided = 1; return 0; }
static inline int MPI_Finalize() { return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return 0; }
static inline int MPI_Abort(MPI_Comm, int code) { std::abort(); return code; }
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
static inline int MPI_Allreduce(const void *in, void *out, int n, MPI_Datatype t, MPI_Op, MPI_Comm) {
  if (in != out && in != (const void *)-1) std::memcpy(out, in, (size_t)n * lmc_shim_type_size(t));
  return 0;
}
#define MPI_IN_PLACE ((void *)-1)
static inline int MPI_Allgather(const void *in, int n, MPI_Datatype t, void *out, int, MPI_Datatype, MPI_Comm) {
  if (in != out && in != MPI_IN_PLACE) std::memcpy(out, in, (size_t)n * lmc_shim_type_size(t));
  return 0;
}
static inline int MPI_Op_create(MPI_User_function *, int, MPI_Op *op) { *op = 1; return 0; }
static inline int MPI_Op_free(MPI_Op *) { return 0; }
static inline int MPI_Type_create_struct(int n, const int *lens, const MPI_Aint *, const MPI_Datatype *types,
                                         MPI_Datatype *newtype) {
  int sz = 0; for (int i = 0; i < n; ++i) sz += lens[i] * lmc_shim_type_size(types[i]);
  *newtype = sz; return 0;
}
static inline int MPI_Type_commit(MPI_Datatype *) { return 0; }
static inline int MPI_Type_free(MPI_Datatype *) { return 0; }
