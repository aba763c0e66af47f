// runtime.cu -- per-thread CUDA state, pointer residency, staging arena, counters, error text.
// See runtime.h for the contract.  Nothing in here computes field values.
#include "runtime.h"

#include "../../include/fcb200.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace fcb200 {

namespace {

constexpr size_t ALIGN = 256;

inline size_t round_up(size_t v, size_t a)
{
  return (v + a - 1) / a * a;
}

struct Chunk
{
  char* ptr = nullptr;
  size_t size = 0;
  size_t used = 0;
};

// bump allocator over a few large chunks; chunks are merged into one at the next reset
struct Arena
{
  std::vector<Chunk> chunks;
  bool pinned = false;

  void* alloc(size_t bytes)
  {
    bytes = round_up(bytes ? bytes : 1, ALIGN);
    for (auto& c : chunks) {
      if (c.size - c.used >= bytes) {
        void* p = c.ptr + c.used;
        c.used += bytes;
        return p;
      }
    }
    Chunk c;
    c.size = round_up(bytes + bytes / 4, size_t(1) << 20);
    cudaError_t e = pinned ? cudaMallocHost((void**)&c.ptr, c.size) : cudaMalloc((void**)&c.ptr, c.size);
    if (!cuda_ok(e, pinned ? "cudaMallocHost(arena)" : "cudaMalloc(arena)"))
      return nullptr;
    c.used = bytes;
    chunks.push_back(c);
    return c.ptr;
  }

  void release_all()
  {
    for (auto& c : chunks) {
      if (pinned)
        cudaFreeHost(c.ptr);
      else
        cudaFree(c.ptr);
    }
    chunks.clear();
  }

  // called when no work is in flight
  void reset()
  {
    if (chunks.size() > 1) {
      size_t total = 0;
      for (auto& c : chunks)
        total += c.size;
      release_all();
      Chunk c;
      c.size = total;
      cudaError_t e = pinned ? cudaMallocHost((void**)&c.ptr, c.size) : cudaMalloc((void**)&c.ptr, c.size);
      if (e == cudaSuccess)
        chunks.push_back(c);
      else
        cudaGetLastError();
    }
    for (auto& c : chunks)
      c.used = 0;
  }
};

struct Deferred
{
  Finalizer fin;
  const unsigned long long* host_counters;
  // Call::finish_counted: the flag rule in the open (pool_index < 0: an opaque finaliser)
  int* flags = nullptr;
  int nfields = 0;
  unsigned long long denom = 0;
  long long pool_index = -1; // of the rule's first counter in the thread's counter pool
};

std::atomic<SlabReduceFn> g_slab_reduce{nullptr};

// {count == 0, count == denom} per pooled counter (denom 0 = no rule for this counter: {count == 0, 0})
__global__ void flag_words_kernel(const unsigned long long* __restrict__ counters, const unsigned long long* __restrict__ denom, int* __restrict__ words, int n)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const unsigned long long c = counters[i], d = denom[i];
    words[2 * i] = c == 0 ? 1 : 0;
    words[2 * i + 1] = (d != 0 && c == d) ? 1 : 0;
  }
}

thread_local char g_error[512] = "";
thread_local bool* t_capturing_failed = nullptr; // &Graph::failed of the graph this thread is capturing
std::atomic<unsigned long long> g_launches{0};

} // namespace

struct Slot
{
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr; // recorded after the last chunk that used this slot
  bool busy = false;
  Arena dev;
  Arena pin;
  Slot() { pin.pinned = true; }
};

// A captured chain of calls (fcb200_graph_begin .. fcb200_graph_end): the CUDA graph, the arenas its nodes refer to
// (per-field metadata in pinned memory and its device copy, scratch fields, counters) and the finalisers that turn
// the counters into ValuesDefined flags after every launch.
struct Graph
{
  int device = -1;
  Arena dev;
  Arena pin;
  std::vector<Deferred> queue;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  unsigned kernels = 0; // kernel nodes (what one launch adds to fcb200_launch_count)
  bool failed = false;  // a call between begin and end returned an error
  // The counters of all captured calls come from ONE block: one memset node at the head of the graph, one device-to-host copy at
  // its tail (a memset and a copy node per call cost ~5 us each per launch).  Per-field metadata and small tables never change after
  // the capture (scalar arguments and input flags are frozen): they are uploaded once, on `side`, outside the graph.
  static constexpr size_t CBLOCK = 4096;
  unsigned long long* cblock = nullptr;
  unsigned long long* cblock_host = nullptr;
  size_t cused = 0;
  cudaStream_t side = nullptr;
  Graph() { pin.pinned = true; }
  ~Graph()
  {
    if (exec)
      cudaGraphExecDestroy(exec);
    if (graph)
      cudaGraphDestroy(graph);
    if (side)
      cudaStreamDestroy(side);
    dev.release_all();
    pin.release_all();
    cudaGetLastError();
  }
};

struct ThreadState
{
  int device = -1;
  Graph* capturing = nullptr; // between fcb200_graph_begin() and fcb200_graph_end()
  Slot slots[PIPE_SLOTS];
  cudaEvent_t fork_event = nullptr;
  bool pipe_used = false;
  cudaStream_t own_stream = nullptr;
  cudaStream_t user_stream = nullptr;
  bool use_user_stream = false;
  bool deferred = false;
  bool in_flight = false; // deferred work queued since the last synchronisation
  Arena dev;
  Arena pin;
  std::vector<Deferred> queue;
  // Undefined-point counters come from a device pool that is ALWAYS ZERO when handed out: the pool's used
  // part is copied to its pinned mirror once per drain (not once per call) and cleared right after,
  // so a call costs no memset and no device-to-host copy of its own.
  unsigned long long* counter_pool = nullptr;
  unsigned long long* counter_mirror = nullptr; // pinned
  size_t counter_used = 0;
  static constexpr size_t COUNTER_POOL = size_t(1) << 18; // counters (2 MB)

  ThreadState() { pin.pinned = true; }

  void release_counter_pool()
  {
    if (counter_pool)
      cudaFree(counter_pool);
    if (counter_mirror)
      cudaFreeHost(counter_mirror);
    counter_pool = counter_mirror = nullptr;
    counter_used = 0;
  }

  // `count` zeroed counters from the pool (nullptr: pool exhausted or unavailable, the caller falls back)
  unsigned long long* pool_counters(size_t count)
  {
    if (!counter_pool) {
      if (cudaMalloc((void**)&counter_pool, COUNTER_POOL * sizeof(unsigned long long)) != cudaSuccess ||
          cudaMallocHost((void**)&counter_mirror, COUNTER_POOL * sizeof(unsigned long long)) != cudaSuccess ||
          cudaMemsetAsync(counter_pool, 0, COUNTER_POOL * sizeof(unsigned long long), stream()) != cudaSuccess) {
        cudaGetLastError();
        release_counter_pool();
        return nullptr;
      }
      // the pool may be used from the pipeline slots' streams: make the clear visible to all of them
      cudaStreamSynchronize(stream());
    }
    if (counter_used + count > COUNTER_POOL)
      return nullptr;
    unsigned long long* p = counter_pool + counter_used;
    counter_used += count;
    return p;
  }
  ~ThreadState()
  {
    // the context may already be gone at thread/process exit: ignore errors
    if (own_stream)
      cudaStreamDestroy(own_stream);
    dev.release_all();
    pin.release_all();
    release_slots();
    release_counter_pool();
    cudaGetLastError();
  }

  void release_slots()
  {
    for (auto& s : slots) {
      if (s.stream)
        cudaStreamDestroy(s.stream);
      if (s.done)
        cudaEventDestroy(s.done);
      s.stream = nullptr;
      s.done = nullptr;
      s.busy = false;
      s.dev.release_all();
      s.pin.release_all();
    }
    if (fork_event)
      cudaEventDestroy(fork_event);
    fork_event = nullptr;
  }

  bool init_slots()
  {
    for (auto& s : slots) {
      if (!s.stream && !cuda_ok(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "cudaStreamCreate(pipeline)"))
        return false;
      if (!s.done && !cuda_ok(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming), "cudaEventCreate(pipeline)"))
        return false;
    }
    if (!fork_event && !cuda_ok(cudaEventCreateWithFlags(&fork_event, cudaEventDisableTiming), "cudaEventCreate(fork)"))
      return false;
    return true;
  }

  bool init()
  {
    int d = -1;
    if (!cuda_ok(cudaGetDevice(&d), "cudaGetDevice (is a CUDA device visible?)"))
      return false;
    if (d != device) {
      // the calling thread switched device: drop everything that belonged to the old one
      if (device >= 0) {
        cudaSetDevice(device);
        if (own_stream)
          cudaStreamDestroy(own_stream);
        dev.release_all();
        pin.release_all();
        release_slots();
        release_counter_pool();
        own_stream = nullptr;
        cudaSetDevice(d);
      }
      device = d;
    }
    if (!own_stream) {
      if (!cuda_ok(cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking), "cudaStreamCreate"))
        return false;
    }
    return true;
  }

  // A capture always runs on the library's own stream: the caller's stream may be the legacy default stream, which cannot be
  // captured.  The graph is LAUNCHED on the caller's stream like any other work.
  cudaStream_t stream() const { return (use_user_stream && !capturing) ? user_stream : own_stream; }

  // fcb200_slab_reduce_flags: the rules queued so far, their words all-reduced on the stream; read back by drain()
  size_t reduced_counters = 0;       // pooled counters covered by the pending reduction (0 = none pending)
  size_t reduced_queue = 0;          // queue entries covered
  const int* reduced_words = nullptr; // pinned, 2 ints per pooled counter

  int reduce_flags()
  {
    if (!deferred || capturing) {
      set_error("fcb200: fcb200_slab_reduce_flags() belongs between fcb200_begin_deferred() and fcb200_end_deferred()");
      return -1;
    }
    if (reduced_counters) {
      set_error("fcb200: fcb200_slab_reduce_flags() was already called in this deferred region");
      return -1;
    }
    const SlabReduceFn reduce = g_slab_reduce.load();
    if (!reduce || counter_used == 0)
      return 1; // one rank, or nothing to combine
    const size_t n = counter_used;
    unsigned long long* h_denom = static_cast<unsigned long long*>(pin.alloc(n * sizeof(unsigned long long)));
    int* h_words = static_cast<int*>(pin.alloc(2 * n * sizeof(int)));
    unsigned long long* d_denom = static_cast<unsigned long long*>(dev.alloc(n * sizeof(unsigned long long)));
    int* d_words = static_cast<int*>(dev.alloc(2 * n * sizeof(int)));
    if (!h_denom || !h_words || !d_denom || !d_words)
      return -1;
    memset(h_denom, 0, n * sizeof(unsigned long long));
    for (const auto& d : queue)
      if (d.pool_index >= 0)
        for (int k = 0; k < d.nfields; ++k)
          h_denom[d.pool_index + k] = d.denom;
    cudaStream_t st = stream();
    if (!cuda_ok(cudaMemcpyAsync(d_denom, h_denom, n * sizeof(unsigned long long), cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(H2D flag rules)"))
      return -1;
    flag_words_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(counter_pool, d_denom, d_words, (int)n);
    count_launch();
    if (!reduce(d_words, 2 * n, st))
      return -1;
    if (!cuda_ok(cudaMemcpyAsync(h_words, d_words, 2 * n * sizeof(int), cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(D2H flag words)"))
      return -1;
    reduced_counters = n;
    reduced_queue = queue.size();
    reduced_words = h_words;
    return 1;
  }

  bool drain()
  {
    bool ok = true;
    if (counter_used > 0)
      ok = cuda_ok(cudaMemcpyAsync(counter_mirror, counter_pool, counter_used * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream()),
                   "cudaMemcpyAsync(D2H counters)") &&
           cuda_ok(cudaMemsetAsync(counter_pool, 0, counter_used * sizeof(unsigned long long), stream()), "cudaMemsetAsync(counters)");
    counter_used = 0;
    ok = cuda_ok(cudaStreamSynchronize(stream()), "cudaStreamSynchronize") && ok;
    if (ok) {
      for (size_t i = 0; i < queue.size(); ++i) {
        const Deferred& d = queue[i];
        if (i < reduced_queue && d.pool_index >= 0 && (size_t)d.pool_index + (size_t)d.nfields <= reduced_counters) {
          // the global flag: ALL iff no rank counted an undefined point, NONE iff every rank counted all of its points
          for (int k = 0; k < d.nfields; ++k) {
            const int* w = reduced_words + 2 * (d.pool_index + k);
            d.flags[k] = w[0] ? ALL_DEFINED : w[1] ? NONE_DEFINED : SOME_DEFINED;
          }
        } else if (d.fin) {
          d.fin(d.host_counters);
        }
      }
    }
    reduced_counters = reduced_queue = 0;
    reduced_words = nullptr;
    queue.clear();
    in_flight = false;
    dev.reset();
    pin.reset();
    for (auto& s : slots)
      s.busy = false; // the main stream waited for every slot before it drained
    return ok;
  }
};

namespace {
ThreadState& thread_state()
{
  thread_local ThreadState ts;
  return ts;
}
} // namespace

void set_error(const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  if (t_capturing_failed)
    *t_capturing_failed = true; // an error between fcb200_graph_begin() and _end(): the graph is not built
}

bool cuda_ok(cudaError_t e, const char* what)
{
  if (e == cudaSuccess)
    return true;
  set_error("fcb200: %s failed: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  cudaGetLastError(); // clear the sticky-less error so that later calls report their own
  return false;
}

void count_launch(unsigned n)
{
  g_launches.fetch_add(n, std::memory_order_relaxed);
}

int sm_count()
{
  static std::mutex m;
  static int cached[64] = {0};
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64)
    return 148;
  std::lock_guard<std::mutex> lock(m);
  if (!cached[d]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n <= 0)
      n = 148;
    cached[d] = n;
  }
  return cached[d];
}

// ------------------------------------------------------------------------------------------- Call

Call::Call(int slot)
{
  ts_ = &thread_state();
  slot_ = slot;
  ok_ = ts_->init();
  if (!ok_)
    return;
  if (ts_->capturing && slot != 0) {
    set_error("fcb200: host-memory fields cannot be captured into a graph (pass device pointers)");
    ok_ = false;
    return;
  }
  if (slot == 0) {
    stream_ = ts_->stream();
    if (!ts_->in_flight && !ts_->capturing) {
      ts_->dev.reset();
      ts_->pin.reset();
    }
    return;
  }
  // pipeline slot: wait until the chunk that used this slot last is completely done (that is the
  // pipeline's back pressure), then recycle the slot's arenas
  Slot& s = ts_->slots[slot - 1];
  stream_ = s.stream;
  if (s.busy) {
    ok_ = cuda_ok(cudaEventSynchronize(s.done), "cudaEventSynchronize(pipeline slot)");
    s.busy = false;
  }
  for (auto& c : s.dev.chunks)
    c.used = 0;
  for (auto& c : s.pin.chunks)
    c.used = 0;
}

Call::~Call() {}

void* Call::arena_alloc(size_t bytes)
{
  if (!ok_)
    return nullptr;
  void* p = (ts_->capturing ? ts_->capturing->dev : slot_ == 0 ? ts_->dev : ts_->slots[slot_ - 1].dev).alloc(bytes);
  if (!p)
    ok_ = false;
  return p;
}

void* Call::pinned_alloc(size_t bytes)
{
  if (!ok_)
    return nullptr;
  void* p = (ts_->capturing ? ts_->capturing->pin : slot_ == 0 ? ts_->pin : ts_->slots[slot_ - 1].pin).alloc(bytes);
  if (!p)
    ok_ = false;
  return p;
}

bool Call::classify(const void* p, bool* is_dev)
{
  if (!p) {
    set_error("fcb200: null field pointer");
    return false;
  }
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *is_dev = false; // plain host memory on old drivers
    return true;
  }
  *is_dev = (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
  if (attr.type == cudaMemoryTypeDevice && attr.device != ts_->device) {
    set_error("fcb200: field lives on device %d but the calling thread's current device is %d", attr.device, ts_->device);
    return false;
  }
  return true;
}

bool Call::is_device(const void* p)
{
  bool d = true;
  if (ok_ && !classify(p, &d))
    ok_ = false;
  return d;
}

const float* Call::in(const float* p, size_t count)
{
  if (!ok_)
    return nullptr;
  bool dev = false;
  if (!classify(p, &dev)) {
    ok_ = false;
    return nullptr;
  }
  if (dev)
    return p;
  if (ts_->capturing) {
    set_error("fcb200: host-memory fields cannot be captured into a graph (pass device pointers)");
    ok_ = false;
    return nullptr;
  }
  for (const auto& q : pending_)
    if (q.host == p && q.bytes >= count * sizeof(float))
      return static_cast<const float*>(q.dev);
  void* d = arena_alloc(count * sizeof(float));
  if (!d)
    return nullptr;
  if (!cuda_ok(cudaMemcpyAsync(d, p, count * sizeof(float), cudaMemcpyHostToDevice, stream_), "cudaMemcpyAsync(H2D field)")) {
    ok_ = false;
    return nullptr;
  }
  pending_.push_back({p, d, count * sizeof(float), false});
  staged_ += count * sizeof(float);
  return static_cast<const float*>(d);
}

float* Call::out(float* p, size_t count)
{
  if (!ok_)
    return nullptr;
  bool dev = false;
  if (!classify(p, &dev)) {
    ok_ = false;
    return nullptr;
  }
  if (dev)
    return p;
  if (ts_->capturing) {
    set_error("fcb200: host-memory fields cannot be captured into a graph (pass device pointers)");
    ok_ = false;
    return nullptr;
  }
  for (auto& q : pending_) {
    if (q.host == p && q.bytes >= count * sizeof(float)) {
      q.copy_back = true;
      return static_cast<float*>(q.dev);
    }
  }
  void* d = arena_alloc(count * sizeof(float));
  if (!d)
    return nullptr;
  pending_.push_back({p, d, count * sizeof(float), true});
  staged_ += count * sizeof(float);
  return static_cast<float*>(d);
}

float* Call::inout(float* p, size_t count)
{
  const float* d = in(p, count);
  if (!d)
    return nullptr;
  return out(p, count);
}

void* Call::scratch(size_t bytes)
{
  return arena_alloc(bytes);
}

FieldMeta* Call::meta_host(int nfields)
{
  meta_n_ = nfields;
  meta_host_ = static_cast<FieldMeta*>(pinned_alloc(sizeof(FieldMeta) * (size_t)nfields));
  return meta_host_;
}

const FieldMeta* Call::upload_meta()
{
  if (!ok_ || !meta_host_)
    return nullptr;
  return static_cast<const FieldMeta*>(upload_small(nullptr, sizeof(FieldMeta) * (size_t)meta_n_));
}

const void* Call::upload_small(const void* host, size_t bytes)
{
  if (!ok_)
    return nullptr;
  const void* src = host;
  if (host == nullptr) {
    src = meta_host_; // already pinned
  } else {
    void* pin = pinned_alloc(bytes);
    if (!pin)
      return nullptr;
    memcpy(pin, host, bytes);
    src = pin;
  }
  void* d = arena_alloc(bytes);
  if (!d)
    return nullptr;
  if (ts_->capturing) { // frozen at capture time: uploaded now, not a node of the graph
    if (!cuda_ok(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, ts_->capturing->side), "cudaMemcpyAsync(H2D table, capture)") ||
        !cuda_ok(cudaStreamSynchronize(ts_->capturing->side), "cudaStreamSynchronize(capture side stream)")) {
      ok_ = false;
      return nullptr;
    }
    return d;
  }
  if (!cuda_ok(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, stream_), "cudaMemcpyAsync(H2D table)")) {
    ok_ = false;
    return nullptr;
  }
  return d;
}

unsigned long long* Call::counters(int count)
{
  if (!ok_)
    return nullptr;
  counters_n_ = count;
  // (a captured call owns its counters: the graph clears them and copies them back itself at every launch)
  if (Graph* g = ts_->capturing) {
    if (g->cblock && g->cused + (size_t)count <= Graph::CBLOCK) {
      counters_dev_ = g->cblock + g->cused;
      counters_graph_ = true;
      g->cused += (size_t)count;
      return counters_dev_;
    }
  }
  counters_dev_ = ts_->capturing ? nullptr : ts_->pool_counters((size_t)count);
  counters_pooled_ = counters_dev_ != nullptr;
  if (counters_pooled_)
    return counters_dev_;
  counters_dev_ = static_cast<unsigned long long*>(arena_alloc(sizeof(unsigned long long) * (size_t)count));
  if (!counters_dev_)
    return nullptr;
  if (!cuda_ok(cudaMemsetAsync(counters_dev_, 0, sizeof(unsigned long long) * (size_t)count, stream_), "cudaMemsetAsync(counters)")) {
    ok_ = false;
    return nullptr;
  }
  return counters_dev_;
}

int Call::finish(const Finalizer& fin)
{
  if (!ok_)
    return -1;
  finished_ = true;
  if (!cuda_ok(cudaGetLastError(), "kernel launch")) {
    ok_ = false;
    return -1;
  }
  for (const auto& q : pending_) {
    if (q.copy_back) {
      if (!cuda_ok(cudaMemcpyAsync(const_cast<void*>(q.host), q.dev, q.bytes, cudaMemcpyDeviceToHost, stream_), "cudaMemcpyAsync(D2H field)"))
        return -1;
    }
  }
  const unsigned long long* host_counters = nullptr;
  if (counters_n_ > 0 && counters_graph_) {
    host_counters = ts_->capturing->cblock_host + (counters_dev_ - ts_->capturing->cblock); // filled by the graph's last node
  } else if (counters_n_ > 0 && counters_pooled_) {
    host_counters = ts_->counter_mirror + (counters_dev_ - ts_->counter_pool); // filled by drain()
  } else if (counters_n_ > 0) {
    // counters are read by the finaliser when the whole call drains: they live in the main pinned
    // arena, which is not recycled while work is in flight
    void* pin = (ts_->capturing ? ts_->capturing->pin : ts_->pin).alloc(sizeof(unsigned long long) * (size_t)counters_n_);
    if (!pin) {
      ok_ = false;
      return -1;
    }
    if (!cuda_ok(cudaMemcpyAsync(pin, counters_dev_, sizeof(unsigned long long) * (size_t)counters_n_, cudaMemcpyDeviceToHost, stream_),
                 "cudaMemcpyAsync(D2H counters)"))
      return -1;
    host_counters = static_cast<const unsigned long long*>(pin);
  }
  if (ts_->capturing) { // the finaliser runs after every launch of the graph
    ts_->capturing->queue.push_back({fin, host_counters});
    return 1;
  }
  ts_->queue.push_back({fin, host_counters});
  ts_->in_flight = true;
  if (slot_ > 0) { // a chunk of a pipelined call: pipeline_join() drains
    Slot& s = ts_->slots[slot_ - 1];
    if (!cuda_ok(cudaEventRecord(s.done, s.stream), "cudaEventRecord(pipeline slot)"))
      return -1;
    s.busy = true;
    return 1;
  }
  if (ts_->deferred)
    return 1;
  return ts_->drain() ? 1 : -1;
}

int Call::finish_counted(int* fDefined, int nfields, unsigned long long denom, int counter_offset)
{
  const Finalizer fin = [=](const unsigned long long* cnt) {
    for (int k = 0; k < nfields; ++k)
      fDefined[k] = check_defined(cnt[counter_offset + k], denom);
  };
  const bool pooled = counters_pooled_ && slot_ == 0 && !ts_->capturing;
  const size_t before = ts_->queue.size();
  const int rc = finish(fin); // (drains at once unless the thread is in deferred mode)
  if (rc == 1 && pooled && ts_->deferred && ts_->queue.size() == before + 1) {
    Deferred& d = ts_->queue.back();
    d.flags = fDefined;
    d.nfields = nfields;
    d.denom = denom;
    d.pool_index = (counters_dev_ - ts_->counter_pool) + counter_offset;
  }
  return rc;
}

void set_slab_reduce(SlabReduceFn fn)
{
  g_slab_reduce.store(fn);
}

int reduce_queued_flags()
{
  ThreadState& ts = thread_state();
  if (!ts.init())
    return -1;
  return ts.reduce_flags();
}

bool pipeline_fork()
{
  ThreadState& ts = thread_state();
  if (ts.capturing) {
    set_error("fcb200: host-memory fields cannot be captured into a graph (pass device pointers)");
    ts.capturing->failed = true;
    return false;
  }
  if (!ts.init() || !ts.init_slots())
    return false;
  // everything already queued on the main stream happens before the chunks
  if (!cuda_ok(cudaEventRecord(ts.fork_event, ts.stream()), "cudaEventRecord(fork)"))
    return false;
  for (auto& s : ts.slots)
    if (!cuda_ok(cudaStreamWaitEvent(s.stream, ts.fork_event, 0), "cudaStreamWaitEvent(fork)"))
      return false;
  ts.in_flight = true; // keep the main pinned arena (counters) alive until the join
  return true;
}

int pipeline_join()
{
  ThreadState& ts = thread_state();
  for (auto& s : ts.slots) {
    if (s.busy && !cuda_ok(cudaStreamWaitEvent(ts.stream(), s.done, 0), "cudaStreamWaitEvent(join)"))
      return -1;
  }
  if (ts.deferred)
    return 1;
  return ts.drain() ? 1 : -1;
}

} // namespace fcb200

// ------------------------------------------------------------------------------------ C-ABI: runtime

using fcb200::thread_state;

extern "C" {

const char* fcb200_last_error(void)
{
  return fcb200::g_error;
}

const char* fcb200_version(void)
{
  return "fcb200 0.1 (sm_100a)";
}

int fcb200_device_count(void)
{
  int n = 0;
  if (!fcb200::cuda_ok(cudaGetDeviceCount(&n), "cudaGetDeviceCount"))
    return -1;
  return n;
}

int fcb200_set_device(int device)
{
  return fcb200::cuda_ok(cudaSetDevice(device), "cudaSetDevice") ? 1 : -1;
}

int fcb200_set_stream(void* cuda_stream, int use_it)
{
  auto& ts = thread_state();
  if (ts.capturing) {
    fcb200::set_error("fcb200: fcb200_set_stream() inside fcb200_graph_begin() .. fcb200_graph_end()");
    return -1;
  }
  if (ts.in_flight && !ts.drain())
    return -1;
  ts.user_stream = static_cast<cudaStream_t>(cuda_stream);
  ts.use_user_stream = use_it != 0;
  return 1;
}

int fcb200_begin_deferred(void)
{
  auto& ts = thread_state();
  if (ts.capturing) {
    fcb200::set_error("fcb200: fcb200_begin_deferred() inside fcb200_graph_begin() .. fcb200_graph_end()");
    return -1;
  }
  ts.deferred = true;
  return 1;
}

int fcb200_in_deferred(void)
{
  return thread_state().deferred ? 1 : 0;
}

int fcb200_end_deferred(void)
{
  auto& ts = thread_state();
  if (ts.capturing) {
    fcb200::set_error("fcb200: fcb200_end_deferred() inside fcb200_graph_begin() .. fcb200_graph_end()");
    return -1;
  }
  ts.deferred = false;
  if (!ts.in_flight)
    return 1;
  if (!ts.init())
    return -1;
  return ts.drain() ? 1 : -1;
}

int fcb200_synchronize(void)
{
  auto& ts = thread_state();
  if (ts.capturing) {
    fcb200::set_error("fcb200: fcb200_synchronize() inside fcb200_graph_begin() .. fcb200_graph_end()");
    return -1;
  }
  if (!ts.init())
    return -1;
  if (ts.in_flight)
    return ts.drain() ? 1 : -1;
  return fcb200::cuda_ok(cudaStreamSynchronize(ts.stream()), "cudaStreamSynchronize") ? 1 : -1;
}

// ---- graphs: a chain of calls on device-resident fields, captured once and replayed with one launch ----------------------
// Between begin and end every fcb200_* call is recorded instead of run (stream capture): its kernels, the upload of its
// per-field metadata, the clearing and the read-back of its counters.  Everything the nodes refer to that the library owns
// lives in the graph object; the field pointers and the fDefined pointers belong to the caller and must stay valid for as
// long as the graph is launched.  Input flags are read at capture time (as in deferred mode).

int fcb200_graph_begin(void)
{
  auto& ts = thread_state();
  if (!ts.init())
    return -1;
  if (ts.capturing || ts.deferred) {
    fcb200::set_error("fcb200: fcb200_graph_begin() inside %s", ts.capturing ? "another capture" : "deferred mode");
    return -1;
  }
  if (ts.in_flight && !ts.drain())
    return -1;
  auto* g = new fcb200::Graph;
  g->device = ts.device;
  g->cblock = static_cast<unsigned long long*>(g->dev.alloc(fcb200::Graph::CBLOCK * sizeof(unsigned long long)));
  g->cblock_host = static_cast<unsigned long long*>(g->pin.alloc(fcb200::Graph::CBLOCK * sizeof(unsigned long long)));
  if (!g->cblock || !g->cblock_host || !fcb200::cuda_ok(cudaStreamCreateWithFlags(&g->side, cudaStreamNonBlocking), "cudaStreamCreate(capture side stream)")) {
    delete g;
    return -1;
  }
  // relaxed mode: a call may grow the graph's arenas (cudaMalloc) and upload its tables on `side` while the capture is open
  if (!fcb200::cuda_ok(cudaStreamBeginCapture(ts.own_stream, cudaStreamCaptureModeRelaxed), "cudaStreamBeginCapture")) {
    delete g;
    return -1;
  }
  ts.capturing = g;
  ts.deferred = true;
  fcb200::t_capturing_failed = &g->failed;
  if (!fcb200::cuda_ok(cudaMemsetAsync(g->cblock, 0, fcb200::Graph::CBLOCK * sizeof(unsigned long long), ts.stream()), "cudaMemsetAsync(graph counters)"))
    return -1; // (the capture stays open and poisoned: fcb200_graph_end() cleans up)
  return 1;
}

int fcb200_graph_end(void** graph)
{
  auto& ts = thread_state();
  if (graph)
    *graph = nullptr;
  fcb200::Graph* g = ts.capturing;
  if (!g) {
    fcb200::set_error("fcb200: fcb200_graph_end() without fcb200_graph_begin()");
    return -1;
  }
  if (!g->failed && g->cused > 0)
    fcb200::cuda_ok(cudaMemcpyAsync(g->cblock_host, g->cblock, g->cused * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ts.stream()),
                    "cudaMemcpyAsync(D2H graph counters)");
  ts.capturing = nullptr;
  ts.deferred = false;
  fcb200::t_capturing_failed = nullptr;
  const cudaError_t e = cudaStreamEndCapture(ts.own_stream, &g->graph); // (always called: it also closes a broken capture)
  if (g->failed || !graph) {
    cudaGetLastError();
    delete g; // keeps the error text of the call that failed
    if (!graph)
      fcb200::set_error("fcb200: fcb200_graph_end(NULL)");
    return -1;
  }
  if (!fcb200::cuda_ok(e, "cudaStreamEndCapture") || !fcb200::cuda_ok(cudaGraphInstantiate(&g->exec, g->graph, 0), "cudaGraphInstantiate")) {
    delete g;
    return -1;
  }
  size_t n = 0;
  if (cudaGraphGetNodes(g->graph, nullptr, &n) == cudaSuccess && n > 0) {
    std::vector<cudaGraphNode_t> nodes(n);
    if (cudaGraphGetNodes(g->graph, nodes.data(), &n) == cudaSuccess) {
      for (size_t i = 0; i < n; ++i) {
        cudaGraphNodeType t;
        if (cudaGraphNodeGetType(nodes[i], &t) == cudaSuccess && t == cudaGraphNodeTypeKernel)
          ++g->kernels;
      }
    }
  }
  cudaGetLastError();
  *graph = g;
  return 1;
}

int fcb200_graph_launch(void* graph)
{
  auto* g = static_cast<fcb200::Graph*>(graph);
  auto& ts = thread_state();
  if (!g || !g->exec) {
    fcb200::set_error("fcb200: fcb200_graph_launch(): not a graph");
    return -1;
  }
  if (!ts.init())
    return -1;
  if (ts.capturing || ts.device != g->device) {
    fcb200::set_error(ts.capturing ? "fcb200: fcb200_graph_launch() inside a capture" : "fcb200: the graph was captured on device %d", g->device);
    return -1;
  }
  if (!fcb200::cuda_ok(cudaGraphLaunch(g->exec, ts.stream()), "cudaGraphLaunch"))
    return -1;
  fcb200::count_launch(g->kernels);
  for (const auto& d : g->queue)
    ts.queue.push_back(d);
  ts.in_flight = true;
  if (ts.deferred)
    return 1;
  return ts.drain() ? 1 : -1;
}

int fcb200_graph_kernels(void* graph)
{
  auto* g = static_cast<fcb200::Graph*>(graph);
  return g ? (int)g->kernels : -1;
}

int fcb200_graph_destroy(void* graph)
{
  auto* g = static_cast<fcb200::Graph*>(graph);
  if (!g)
    return 1;
  auto& ts = thread_state();
  if (ts.in_flight && !ts.drain()) { // a deferred launch of this graph may still be running
    delete g;
    return -1;
  }
  delete g;
  return 1;
}

unsigned long long fcb200_launch_count(void)
{
  return fcb200::g_launches.load(std::memory_order_relaxed);
}

} // extern "C"
