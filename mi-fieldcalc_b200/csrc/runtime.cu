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
};

thread_local char g_error[512] = "";
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

struct ThreadState
{
  int device = -1;
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

  cudaStream_t stream() const { return use_user_stream ? user_stream : own_stream; }

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
      for (auto& d : queue)
        if (d.fin)
          d.fin(d.host_counters);
    }
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
  if (slot == 0) {
    stream_ = ts_->stream();
    if (!ts_->in_flight) {
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
  void* p = (slot_ == 0 ? ts_->dev : ts_->slots[slot_ - 1].dev).alloc(bytes);
  if (!p)
    ok_ = false;
  return p;
}

void* Call::pinned_alloc(size_t bytes)
{
  if (!ok_)
    return nullptr;
  void* p = (slot_ == 0 ? ts_->pin : ts_->slots[slot_ - 1].pin).alloc(bytes);
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
  counters_dev_ = ts_->pool_counters((size_t)count);
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
  if (!cuda_ok(cudaGetLastError(), "kernel launch"))
    return -1;
  for (const auto& q : pending_) {
    if (q.copy_back) {
      if (!cuda_ok(cudaMemcpyAsync(const_cast<void*>(q.host), q.dev, q.bytes, cudaMemcpyDeviceToHost, stream_), "cudaMemcpyAsync(D2H field)"))
        return -1;
    }
  }
  const unsigned long long* host_counters = nullptr;
  if (counters_n_ > 0 && counters_pooled_) {
    host_counters = ts_->counter_mirror + (counters_dev_ - ts_->counter_pool); // filled by drain()
  } else if (counters_n_ > 0) {
    // counters are read by the finaliser when the whole call drains: they live in the main pinned
    // arena, which is not recycled while work is in flight
    void* pin = ts_->pin.alloc(sizeof(unsigned long long) * (size_t)counters_n_);
    if (!pin)
      return -1;
    if (!cuda_ok(cudaMemcpyAsync(pin, counters_dev_, sizeof(unsigned long long) * (size_t)counters_n_, cudaMemcpyDeviceToHost, stream_),
                 "cudaMemcpyAsync(D2H counters)"))
      return -1;
    host_counters = static_cast<const unsigned long long*>(pin);
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

bool pipeline_fork()
{
  ThreadState& ts = thread_state();
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
  if (ts.in_flight && !ts.drain())
    return -1;
  ts.user_stream = static_cast<cudaStream_t>(cuda_stream);
  ts.use_user_stream = use_it != 0;
  return 1;
}

int fcb200_begin_deferred(void)
{
  auto& ts = thread_state();
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
  if (!ts.init())
    return -1;
  if (ts.in_flight)
    return ts.drain() ? 1 : -1;
  return fcb200::cuda_ok(cudaStreamSynchronize(ts.stream()), "cudaStreamSynchronize") ? 1 : -1;
}

unsigned long long fcb200_launch_count(void)
{
  return fcb200::g_launches.load(std::memory_order_relaxed);
}

} // extern "C"
