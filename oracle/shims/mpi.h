// Test-infrastructure shim: the subset of MPI the reference's lmc/mc sources use, in one process.
//
// By default a thread is rank 0 of a world of size 1 (every collective is a copy).  The harness can also run
// an N-rank job as N THREADS of this process: each thread calls lmc_shim_mpi::attach(world, rank) and then runs
// the unmodified reference driver (mc::KineticMcChainOmpi / KineticMcFirstMpi need exactly 12 ranks).  The
// collectives then exchange through the shared World; reductions are applied in rank order 0..N-1.
#pragma once
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
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
#define MPI_IN_PLACE ((void *)-1)

namespace lmc_shim_mpi {
struct World {
  explicit World(int n) : size(n), slot(static_cast<size_t>(n), nullptr) {}
  int size;
  std::vector<const void *> slot;      // what each rank contributes to the collective in flight
  // sense-reversing spin barrier (yields after a short spin): collectives of an intra-node MPI cost about a microsecond, a
  // condition-variable barrier would add tens of microseconds per collective to the timed reference
  std::atomic<int> waiting{0};
  std::atomic<unsigned long> generation{0};
  void barrier() {
    const unsigned long g = generation.load(std::memory_order_acquire);
    if (waiting.fetch_add(1, std::memory_order_acq_rel) + 1 == size) {
      waiting.store(0, std::memory_order_relaxed);
      generation.store(g + 1, std::memory_order_release);
    } else {
      for (int spin = 0; generation.load(std::memory_order_acquire) == g; ++spin)
        if (spin > 2000) std::this_thread::yield();
    }
  }
};
inline thread_local World *tl_world = nullptr;
inline thread_local int tl_rank = 0;
inline void attach(World *w, int rank) { tl_world = w; tl_rank = rank; }
inline void detach() { tl_world = nullptr; tl_rank = 0; }
// user-defined reduction operators (MPI_Op_create); op 0 is MPI_SUM
inline std::mutex op_mutex;
inline std::vector<MPI_User_function *> op_table{nullptr};
inline int type_size(MPI_Datatype t) { return t > 0 ? t : 1; }
}  // namespace lmc_shim_mpi

static inline int MPI_Init_thread(int *, char ***, int, int *provided) { if (provided) *provided = 1; return 0; }
static inline int MPI_Finalize() { return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = lmc_shim_mpi::tl_rank; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *s) { *s = lmc_shim_mpi::tl_world ? lmc_shim_mpi::tl_world->size : 1; return 0; }
static inline int MPI_Abort(MPI_Comm, int code) { std::abort(); return code; }
static inline int MPI_Bcast(void *buf, int n, MPI_Datatype t, int root, MPI_Comm) {
  using namespace lmc_shim_mpi;
  World *w = tl_world;
  if (!w || w->size == 1) return 0;
  w->slot[static_cast<size_t>(tl_rank)] = buf;
  w->barrier();
  if (tl_rank != root) std::memcpy(buf, w->slot[static_cast<size_t>(root)], static_cast<size_t>(n) * type_size(t));
  w->barrier();
  return 0;
}
static inline int MPI_Allreduce(const void *in, void *out, int n, MPI_Datatype t, MPI_Op op, MPI_Comm) {
  using namespace lmc_shim_mpi;
  World *w = tl_world;
  const size_t bytes = static_cast<size_t>(n) * type_size(t);
  if (!w || w->size == 1) {
    if (in != out && in != MPI_IN_PLACE) std::memcpy(out, in, bytes);
    return 0;
  }
  std::vector<char> mine(bytes);                       // private copy: `in` may alias `out`
  std::memcpy(mine.data(), in == MPI_IN_PLACE ? out : in, bytes);
  w->slot[static_cast<size_t>(tl_rank)] = mine.data();
  w->barrier();
  std::vector<char> acc(bytes);
  std::memcpy(acc.data(), w->slot[0], bytes);
  for (int r = 1; r < w->size; ++r) {                  // rank order 0..N-1
    if (op == MPI_SUM) {
      double *a = reinterpret_cast<double *>(acc.data());
      const double *b = static_cast<const double *>(w->slot[static_cast<size_t>(r)]);
      for (size_t i = 0; i < bytes / sizeof(double); ++i) a[i] += b[i];
    } else {
      MPI_User_function *f;
      { std::lock_guard<std::mutex> g(op_mutex); f = op_table[static_cast<size_t>(op)]; }
      int len = n;
      MPI_Datatype dt = t;
      f(const_cast<void *>(w->slot[static_cast<size_t>(r)]), acc.data(), &len, &dt);
    }
  }
  w->barrier();
  std::memcpy(out, acc.data(), bytes);
  return 0;
}
static inline int MPI_Allgather(const void *in, int n, MPI_Datatype t, void *out, int, MPI_Datatype, MPI_Comm) {
  using namespace lmc_shim_mpi;
  World *w = tl_world;
  const size_t bytes = static_cast<size_t>(n) * type_size(t);
  if (!w || w->size == 1) {
    if (in != out && in != MPI_IN_PLACE) std::memcpy(out, in, bytes);
    return 0;
  }
  std::vector<char> mine(bytes);
  std::memcpy(mine.data(), in == MPI_IN_PLACE ? static_cast<const char *>(out) + bytes * static_cast<size_t>(tl_rank) : in, bytes);
  w->slot[static_cast<size_t>(tl_rank)] = mine.data();
  w->barrier();
  for (int r = 0; r < w->size; ++r)
    std::memcpy(static_cast<char *>(out) + bytes * static_cast<size_t>(r), w->slot[static_cast<size_t>(r)], bytes);
  w->barrier();
  return 0;
}
static inline int MPI_Op_create(MPI_User_function *f, int, MPI_Op *op) {
  std::lock_guard<std::mutex> g(lmc_shim_mpi::op_mutex);
  lmc_shim_mpi::op_table.push_back(f);
  *op = static_cast<int>(lmc_shim_mpi::op_table.size()) - 1;
  return 0;
}
static inline int MPI_Op_free(MPI_Op *) { return 0; }
static inline int MPI_Type_create_struct(int n, const int *lens, const MPI_Aint *, const MPI_Datatype *types,
                                         MPI_Datatype *newtype) {
  int sz = 0;
  for (int i = 0; i < n; ++i) sz += lens[i] * lmc_shim_mpi::type_size(types[i]);
  *newtype = sz;
  return 0;
}
static inline int MPI_Type_commit(MPI_Datatype *) { return 0; }
static inline int MPI_Type_free(MPI_Datatype *) { return 0; }
