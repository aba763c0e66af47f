// slab.cu -- row-slab decomposition of ONE large grid over several GPUs (SURVEY.md 8e case 2, BASELINE.json configs[2]:
// "3600x1801 x 137 levels ... row-slab sharded across 8 GPUs with NVLink halo exchange").
//
// One process (or host thread) per GPU.  A rank OWNS the rows [r0, r1) of every field and holds an EXTENDED slab: its
// rows plus `halo` rows of the neighbouring ranks above and below (none on the outer side of the first / last rank).
// The ordinary single-GPU operator runs on the extended slab as if it were a grid of its own, and the rank keeps the rows
// it owns.  That is exact (mi-fieldcalc_b200/distributed.py states the argument; tests/test_distributed_cpu.py proves it
// against the oracle): the reference's flat-index wrap stays inside slab + halo, rows within `halo` of an artificial slab
// edge are the only ones that differ from the global result and they are exactly the rows that are thrown away, and the
// global border rows belong to the first / last rank, whose slab has no halo there.
//
// What this file adds is the part that needs the interconnect:
//   fcb200_slab_exchange       the halo rows of a BATCH of extended slabs move to rank-1 / rank+1 with ONE grouped
//                              ncclSend/ncclRecv pair per neighbour: a pack kernel gathers the first / last `halo` owned
//                              rows of every field into two contiguous messages, NCCL moves them over NVLink, an unpack
//                              kernel scatters what arrived into the halo rows.  Everything is enqueued on the calling
//                              thread's stream: it orders itself against the operator that produced the rows and the
//                              operator that consumes them, with no host synchronisation.
//   fcb200_slab_combine_flags  the global ValuesDefined of every field from the per-rank flags (ALL iff every rank says
//                              ALL, NONE iff every rank says NONE): one small ncclAllReduce.
// The exchange is needed when a stencil consumes the OUTPUT of a previous sharded operator (shapiro2_filter ->
// thermalFrontParameter, repeated smoothing ...); static inputs are scattered with their halo and need none.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): libfcb200.so has no link-time dependency on it, single-GPU users
// never load it.  In a process that already uses torch.distributed this resolves to the NCCL torch has loaded.
#include "runtime.h"

#include "../../include/fcb200.h"

#include <dlfcn.h>

#include <atomic>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <vector>

namespace fcb200 {
namespace {

// ---- the slice of NCCL's C API used here (nccl.h: stable ABI since 2.0) ------------------------------------------
typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId
{
  char internal[128];
};
enum { NCCL_SUCCESS = 0 };
enum { NCCL_INT32 = 2, NCCL_FLOAT32 = 7 }; // ncclDataType_t
enum { NCCL_MIN = 3 };                      // ncclRedOp_t

struct Nccl
{
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

struct SlabState
{
  std::mutex lock;
  Nccl nccl;
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1, device = -1;
  int* flag_dev = nullptr;  // scratch of the flag all-reduce (grow-only)
  int* flag_host = nullptr; // pinned
  size_t flag_cap = 0;
};

SlabState& slab()
{
  static SlabState s;
  return s;
}
std::atomic<unsigned long long> g_bytes_sent{0};

bool load_nccl(Nccl& n)
{
  if (n.handle)
    return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (n.handle)
      break;
  }
  if (!n.handle) {
    set_error("fcb200: cannot load NCCL (libnccl.so.2): %s", dlerror());
    return false;
  }
#define FCB_SYM(field, name)                                                                                                                         \
  *reinterpret_cast<void**>(&n.field) = dlsym(n.handle, name);                                                                                       \
  if (!n.field) {                                                                                                                                    \
    set_error("fcb200: NCCL symbol %s not found", name);                                                                                             \
    n.handle = nullptr;                                                                                                                              \
    return false;                                                                                                                                    \
  }
  FCB_SYM(GetUniqueId, "ncclGetUniqueId")
  FCB_SYM(CommInitRank, "ncclCommInitRank")
  FCB_SYM(CommDestroy, "ncclCommDestroy")
  FCB_SYM(GroupStart, "ncclGroupStart")
  FCB_SYM(GroupEnd, "ncclGroupEnd")
  FCB_SYM(Send, "ncclSend")
  FCB_SYM(Recv, "ncclRecv")
  FCB_SYM(AllReduce, "ncclAllReduce")
  FCB_SYM(GetErrorString, "ncclGetErrorString")
#undef FCB_SYM
  return true;
}

bool nccl_ok(const Nccl& n, int rc, const char* what)
{
  if (rc == NCCL_SUCCESS)
    return true;
  set_error("fcb200: %s failed: %s", what, n.GetErrorString ? n.GetErrorString(rc) : "NCCL error");
  return false;
}

// the runtime's hook for fcb200_slab_reduce_flags: an in-place MIN over the ranks of {count == 0, count == denom} words
bool slab_reduce_words(int* words, size_t n, cudaStream_t stream)
{
  SlabState& s = slab();
  std::lock_guard<std::mutex> g(s.lock);
  if (!s.comm) {
    set_error("fcb200: fcb200_slab_init has not been called");
    return false;
  }
  return nccl_ok(s.nccl, s.nccl.AllReduce(words, words, n, NCCL_INT32, NCCL_MIN, s.comm, stream), "ncclAllReduce(flag words)");
}

// One thread per float4 (or float) of a halo row.  Message layout: [field][halo row][nx], the same on both sides.
//   PACK:   msg_up   <- owned rows [up, up + halo)              (goes to rank - 1)
//           msg_down <- owned rows [up + rows - halo, up + rows) (goes to rank + 1)
//   UNPACK: rows [0, up)               <- msg_up   (came from rank - 1)
//           rows [up + rows, ext_rows) <- msg_down (came from rank + 1)
template <bool PACK, int W>
__global__ void __launch_bounds__(256) slab_halo_kernel(float* __restrict__ ext, float* __restrict__ msg_up, float* __restrict__ msg_down, int nxw, int ext_rows,
                                                         int halo, int up, int down, int rows, int nfields)
{
  typedef typename std::conditional<W == 4, float4, float>::type V;
  const long long per_side = (long long)nfields * halo * nxw;
  const long long total = ((up ? 1 : 0) + (down ? 1 : 0)) * per_side;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += 256LL * gridDim.x) {
    const bool is_up = up && i < per_side;
    const long long j = is_up ? i : i - (up ? per_side : 0);
    const int x = (int)(j % nxw);
    const int hr = (int)((j / nxw) % halo);
    const int f = (int)(j / ((long long)nxw * halo));
    V* m = reinterpret_cast<V*>(is_up ? msg_up : msg_down) + j;
    int row;
    if (PACK)
      row = is_up ? up + hr : up + rows - halo + hr;
    else
      row = is_up ? hr : up + rows + hr;
    V* e = reinterpret_cast<V*>(ext) + ((long long)f * ext_rows + row) * nxw + x;
    if (PACK)
      *m = *e;
    else
      *e = *m;
  }
}

} // namespace
} // namespace fcb200

using namespace fcb200;

extern "C" {

int fcb200_slab_unique_id(char* id128)
{
  SlabState& s = slab();
  std::lock_guard<std::mutex> g(s.lock);
  if (!load_nccl(s.nccl))
    return -1;
  ncclUniqueId id;
  if (!nccl_ok(s.nccl, s.nccl.GetUniqueId(&id), "ncclGetUniqueId"))
    return -1;
  memcpy(id128, id.internal, 128);
  return 1;
}

int fcb200_slab_init(int rank, int nranks, const char* id128)
{
  SlabState& s = slab();
  std::lock_guard<std::mutex> g(s.lock);
  if (nranks < 1 || rank < 0 || rank >= nranks) {
    set_error("fcb200: invalid slab rank %d of %d", rank, nranks);
    return -1;
  }
  if (s.comm) {
    set_error("fcb200: the slab communicator is already initialised (fcb200_slab_finalize first)");
    return -1;
  }
  s.rank = rank;
  s.nranks = nranks;
  if (!cuda_ok(cudaGetDevice(&s.device), "cudaGetDevice"))
    return -1;
  if (nranks == 1)
    return 1; // nothing to exchange: every call below is a no-op
  if (!load_nccl(s.nccl))
    return -1;
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  if (!nccl_ok(s.nccl, s.nccl.CommInitRank(&s.comm, nranks, id, rank), "ncclCommInitRank"))
    return -1;
  set_slab_reduce(&slab_reduce_words);
  return 1;
}

int fcb200_slab_finalize(void)
{
  SlabState& s = slab();
  std::lock_guard<std::mutex> g(s.lock);
  set_slab_reduce(nullptr);
  if (s.comm)
    s.nccl.CommDestroy(s.comm);
  s.comm = nullptr;
  if (s.flag_dev)
    cudaFree(s.flag_dev);
  if (s.flag_host)
    cudaFreeHost(s.flag_host);
  s.flag_dev = s.flag_host = nullptr;
  s.flag_cap = 0;
  s.rank = 0;
  s.nranks = 1;
  return 1;
}

int fcb200_slab_rank(void) { return slab().rank; }
int fcb200_slab_nranks(void) { return slab().nranks; }
unsigned long long fcb200_slab_bytes_sent(void) { return g_bytes_sent.load(std::memory_order_relaxed); }

int fcb200_slab_partition(int ny, int halo, int rank, int nranks, int* r0, int* r1, int* lo, int* hi)
{
  if (ny <= 0 || halo < 0 || nranks < 1 || rank < 0 || rank >= nranks)
    return 0;
  const int base = ny / nranks, extra = ny % nranks;
  const int b = rank * base + (rank < extra ? rank : extra);
  const int e = b + base + (rank < extra ? 1 : 0);
  if (e - b < (halo > 1 ? halo : 1))
    return 0; // every halo must come from the direct neighbour
  const int l = b - halo > 0 ? b - halo : 0;
  const int h = e + halo < ny ? e + halo : ny;
  if (h - l < 3)
    return 0; // an extended slab must be a valid grid for the operators
  *r0 = b, *r1 = e, *lo = l, *hi = h;
  return 1;
}

int fcb200_slab_exchange(float* ext, int nx, int ext_rows, int nfields, int halo)
{
  SlabState& s = slab();
  if (s.nranks == 1 || halo == 0)
    return 1;
  if (!s.comm) {
    set_error("fcb200: fcb200_slab_init has not been called");
    return -1;
  }
  const int up = s.rank > 0 ? halo : 0, down = s.rank < s.nranks - 1 ? halo : 0;
  const int rows = ext_rows - up - down;
  if (nx <= 0 || nfields <= 0 || halo < 0 || rows < halo) {
    set_error("fcb200: invalid slab (nx=%d ext_rows=%d nfields=%d halo=%d: %d owned rows)", nx, ext_rows, nfields, halo, rows);
    return -1;
  }
  Call call;
  if (!call.ok())
    return -1;
  if (!call.is_device(ext)) {
    set_error("fcb200: fcb200_slab_exchange needs device memory (the slab stays resident between the sharded operators)");
    return -1;
  }
  const size_t msg = (size_t)nfields * halo * nx; // floats per direction
  // [send up | send down | recv up | recv down]
  float* buf = static_cast<float*>(call.scratch(sizeof(float) * 4 * msg));
  if (!call.ok())
    return -1;
  float *send_up = buf, *send_down = buf + msg, *recv_up = buf + 2 * msg, *recv_down = buf + 3 * msg;
  const bool vec = (nx % 4 == 0) && (reinterpret_cast<uintptr_t>(ext) & 15) == 0;
  const int nxw = vec ? nx / 4 : nx;
  const long long work = (long long)((up ? 1 : 0) + (down ? 1 : 0)) * nfields * halo * nxw;
  unsigned blocks = (unsigned)((work + 255) / 256);
  const unsigned cap = (unsigned)sm_count() * 8;
  if (blocks > cap)
    blocks = cap;
  if (blocks == 0)
    blocks = 1;
  cudaStream_t st = call.stream();
  if (vec)
    slab_halo_kernel<true, 4><<<blocks, 256, 0, st>>>(ext, send_up, send_down, nxw, ext_rows, halo, up, down, rows, nfields);
  else
    slab_halo_kernel<true, 1><<<blocks, 256, 0, st>>>(ext, send_up, send_down, nxw, ext_rows, halo, up, down, rows, nfields);
  {
    std::lock_guard<std::mutex> g(s.lock);
    bool ok = nccl_ok(s.nccl, s.nccl.GroupStart(), "ncclGroupStart");
    if (ok && up) {
      ok = nccl_ok(s.nccl, s.nccl.Send(send_up, msg, NCCL_FLOAT32, s.rank - 1, s.comm, st), "ncclSend(up)") &&
           nccl_ok(s.nccl, s.nccl.Recv(recv_up, msg, NCCL_FLOAT32, s.rank - 1, s.comm, st), "ncclRecv(up)");
    }
    if (ok && down) {
      ok = nccl_ok(s.nccl, s.nccl.Send(send_down, msg, NCCL_FLOAT32, s.rank + 1, s.comm, st), "ncclSend(down)") &&
           nccl_ok(s.nccl, s.nccl.Recv(recv_down, msg, NCCL_FLOAT32, s.rank + 1, s.comm, st), "ncclRecv(down)");
    }
    ok = nccl_ok(s.nccl, s.nccl.GroupEnd(), "ncclGroupEnd") && ok;
    if (!ok)
      return -1;
  }
  if (vec)
    slab_halo_kernel<false, 4><<<blocks, 256, 0, st>>>(ext, recv_up, recv_down, nxw, ext_rows, halo, up, down, rows, nfields);
  else
    slab_halo_kernel<false, 1><<<blocks, 256, 0, st>>>(ext, recv_up, recv_down, nxw, ext_rows, halo, up, down, rows, nfields);
  count_launch(2);
  g_bytes_sent.fetch_add((unsigned long long)(sizeof(float) * msg * ((up ? 1 : 0) + (down ? 1 : 0))), std::memory_order_relaxed);
  return call.finish(Finalizer());
}

int fcb200_slab_reduce_flags(void)
{
  if (slab().nranks == 1)
    return 1;
  return reduce_queued_flags();
}

int fcb200_slab_combine_flags(int* fDefined, int nfields)
{
  SlabState& s = slab();
  if (s.nranks == 1 || nfields <= 0)
    return 1;
  if (!s.comm) {
    set_error("fcb200: fcb200_slab_init has not been called");
    return -1;
  }
  // the local flags are final only once the thread's queued work has drained
  if (fcb200_synchronize() < 0)
    return -1;
  std::lock_guard<std::mutex> g(s.lock);
  const size_t need = 2 * (size_t)nfields;
  if (need > s.flag_cap) {
    if (s.flag_dev)
      cudaFree(s.flag_dev);
    if (s.flag_host)
      cudaFreeHost(s.flag_host);
    s.flag_dev = s.flag_host = nullptr;
    s.flag_cap = 0;
    if (!cuda_ok(cudaMalloc((void**)&s.flag_dev, need * sizeof(int)), "cudaMalloc(flags)") ||
        !cuda_ok(cudaMallocHost((void**)&s.flag_host, need * sizeof(int)), "cudaMallocHost(flags)"))
      return -1;
    s.flag_cap = need;
  }
  for (int k = 0; k < nfields; ++k) {
    s.flag_host[2 * k] = fDefined[k] == ALL_DEFINED ? 1 : 0;
    s.flag_host[2 * k + 1] = fDefined[k] == NONE_DEFINED ? 1 : 0;
  }
  Call call;
  if (!call.ok())
    return -1;
  cudaStream_t st = call.stream();
  if (!cuda_ok(cudaMemcpyAsync(s.flag_dev, s.flag_host, need * sizeof(int), cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(flags)"))
    return -1;
  if (!nccl_ok(s.nccl, s.nccl.AllReduce(s.flag_dev, s.flag_dev, need, NCCL_INT32, NCCL_MIN, s.comm, st), "ncclAllReduce(flags)"))
    return -1;
  if (!cuda_ok(cudaMemcpyAsync(s.flag_host, s.flag_dev, need * sizeof(int), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(flags)") ||
      !cuda_ok(cudaStreamSynchronize(st), "cudaStreamSynchronize(flags)"))
    return -1;
  for (int k = 0; k < nfields; ++k)
    fDefined[k] = s.flag_host[2 * k] ? ALL_DEFINED : s.flag_host[2 * k + 1] ? NONE_DEFINED : SOME_DEFINED;
  return 1;
}

} // extern "C"
