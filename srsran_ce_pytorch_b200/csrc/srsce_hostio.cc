// Host half of the host-buffer pipeline (include/srsce.h: srsce_replicate_host).
//
// The dense grid the reference returns repeats one column over the allocated symbols (ce_rule_tensorized.py:346-347):
// 13/14 of its bytes are copies.  When the caller wants that grid in HOST memory, moving the copies over PCIe
// (54 GB/s device->host on this platform) is the bottleneck of the whole call; the box's cores write host memory
// about three times faster than the link delivers it.  So the host-buffer pipeline brings the per-hop columns
// (srsce_estimate_compact) over the link and replicates them here with streaming stores.  This file does NO arithmetic:
// plans whose output carries the per-symbol CFO rotation (:922-929) are refused and take the dense device path.
#include <stdint.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "srsce_internal.h"

namespace srsce {
namespace {

// A small persistent pool: the replication of one chunk lasts a millisecond or two, thread creation would eat it.
class HostPool {
 public:
  static HostPool& get() {
    static HostPool* pool = nullptr;               // never destroyed: the workers sleep until the process ends; a forked
                                                   // child gets a fresh pool (threads do not survive fork)
    static std::mutex m;
    static pid_t owner = 0;
    std::lock_guard<std::mutex> g(m);
    if (pool == nullptr || owner != getpid()) {
      pool = new HostPool();
      owner = getpid();
    }
    return *pool;
  }

  // fn(part) for part in [0, parts), on up to n_threads threads (the caller is one of them); returns when all are done.
  void run(int n_threads, int parts, const std::function<void(int)>& fn) {
    std::lock_guard<std::mutex> serial(run_mutex_);
    grow(n_threads - 1);
    {
      std::lock_guard<std::mutex> g(m_);
      fn_ = &fn;
      parts_ = parts;
      next_.store(0, std::memory_order_relaxed);
      active_ = static_cast<int>(workers_.size()) < n_threads - 1 ? static_cast<int>(workers_.size()) : n_threads - 1;
      pending_ = active_;
      ++generation_;
    }
    wake_.notify_all();
    drain(fn);
    std::unique_lock<std::mutex> g(m_);
    done_.wait(g, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void grow(int n) {
    while (static_cast<int>(workers_.size()) < n) {
      const int id = static_cast<int>(workers_.size());
      workers_.emplace_back([this, id] { loop(id); });
    }
  }
  void drain(const std::function<void(int)>& fn) {
    for (;;) {
      const int p = next_.fetch_add(1, std::memory_order_relaxed);
      if (p >= parts_) break;
      fn(p);
    }
  }
  void loop(int id) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int)>* fn = nullptr;
      {
        std::unique_lock<std::mutex> g(m_);
        wake_.wait(g, [&] { return generation_ != seen; });
        seen = generation_;
        if (id >= active_) continue;               // this job asked for fewer threads
        fn = fn_;
      }
      drain(*fn);
      {
        std::lock_guard<std::mutex> g(m_);
        --pending_;
      }
      done_.notify_one();
    }
  }

  std::mutex run_mutex_, m_;
  std::condition_variable wake_, done_;
  std::vector<std::thread> workers_;
  const std::function<void(int)>* fn_ = nullptr;
  std::atomic<int> next_{0};
  int parts_ = 0, active_ = 0, pending_ = 0;
  uint64_t generation_ = 0;
};

struct Replicate {
  const char* src;          // [n_units][n_hops][n_sc][row_bytes]
  char* dst;                // [n_units][n_sc][n_sym][row_bytes]
  int n_sc, n_sym, n_hops, row_bytes;
  int band2_lo, band2_hi;   // sub-carriers of hop 2's band
  int8_t sel_in[SRSCE_MAX_SYMBOLS];    // source hop of symbol t for a sub-carrier inside hop 2's band (-1: zero)
  int8_t sel_out[SRSCE_MAX_SYMBOLS];   // ... outside it
  bool uniform;             // one hop, every symbol allocated: a line is n_sym copies of one source row
};

// Non-temporal copy of a finished block to its place in the grid: the grid is written once and read by someone else,
// so it should not displace the staging buffers in the cache, and a streaming store does not read the line first.
inline void stream_out(char* dst, const char* src, size_t bytes) {
#if defined(__SSE2__)
  size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
  if (head > bytes) head = bytes;
  memcpy(dst, src, head);
  size_t k = head;
  for (; k + 16 <= bytes; k += 16)
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + k), _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + k)));
  memcpy(dst + k, src + k, bytes - k);
#else
  memcpy(dst, src, bytes);
#endif
}

constexpr int kBlockRows = 32;

void replicate_rows(const Replicate& j, int64_t row0, int64_t row1) {
  const int rb = j.row_bytes, line = j.n_sym * rb;
  const int64_t hop_stride = int64_t(j.n_sc) * rb;
#if defined(__SSE2__)
  // the common shape (one hop, whole slot allocated, 2 or 4 layers, aligned lines): straight from registers
  if (j.uniform && (rb == 16 || rb == 32) && (reinterpret_cast<uintptr_t>(j.dst) & 15u) == 0) {
    for (int64_t r = row0; r < row1; ++r) {
      const char* s = j.src + r * rb;              // n_hops == 1: source rows are consecutive over (unit, sc)
      char* d = j.dst + r * line;
      const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s));
      if (rb == 16) {
        for (int t = 0; t < j.n_sym; ++t) _mm_stream_si128(reinterpret_cast<__m128i*>(d + 16 * t), a);
      } else {
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 16));
        for (int t = 0; t < j.n_sym; ++t) {
          _mm_stream_si128(reinterpret_cast<__m128i*>(d + 32 * t), a);
          _mm_stream_si128(reinterpret_cast<__m128i*>(d + 32 * t + 16), b);
        }
      }
    }
    _mm_sfence();
    return;
  }
#endif
  alignas(64) char block[kBlockRows * SRSCE_MAX_SYMBOLS * SRSCE_MAX_LAYERS * 8];
  for (int64_t b0 = row0; b0 < row1; b0 += kBlockRows) {
    const int64_t b1 = b0 + kBlockRows < row1 ? b0 + kBlockRows : row1;
    for (int64_t r = b0; r < b1; ++r) {
      const int64_t u = r / j.n_sc;
      const int sc = static_cast<int>(r - u * j.n_sc);
      const char* s0 = j.src + (u * j.n_hops * j.n_sc + sc) * int64_t(rb);
      const int8_t* sel = (j.n_hops == 2 && sc >= j.band2_lo && sc < j.band2_hi) ? j.sel_in : j.sel_out;
      char* d = block + (r - b0) * line;
      for (int t = 0; t < j.n_sym; ++t) {
        if (sel[t] < 0) memset(d + t * rb, 0, rb);
        else memcpy(d + t * rb, s0 + sel[t] * hop_stride, rb);
      }
    }
    stream_out(j.dst + b0 * line, block, size_t(b1 - b0) * line);
  }
#if defined(__SSE2__)
  _mm_sfence();
#endif
}

}  // namespace

int replicate_host(const Plan& plan, const void* ch_sc, void* ch_est, int64_t n_units, int n_threads, int flags,
                   std::string& err) {
  const PlanDev& v = plan.dev;
  if (v.cfo_compensate && v.cfo_defined && !(flags & SRSCE_REPLICATE_FILL)) {
    err = "this plan's grid carries the per-symbol CFO rotation: the host side only replicates, use srsce_estimate";
    return SRSCE_ERR_UNSUPPORTED;
  }
  if (n_units == 0) return SRSCE_OK;
  Replicate j;
  j.src = static_cast<const char*>(ch_sc);
  j.dst = static_cast<char*>(ch_est);
  j.n_sc = v.n_sc;
  j.n_sym = v.n_sym;
  j.n_hops = v.n_hops;
  j.row_bytes = 8 * v.n_layers;
  j.band2_lo = v.n_hops == 2 ? v.hop[1].band_sc0 : 0;
  j.band2_hi = v.n_hops == 2 ? v.hop[1].band_sc1 : 0;
  j.uniform = v.n_hops == 1;
  for (int t = 0; t < SRSCE_MAX_SYMBOLS; ++t) {
    const bool in1 = t >= v.hop[0].start_symbol && t < v.hop[0].end_symbol;
    const bool in2 = v.n_hops == 2 && t >= v.hop[1].start_symbol && t < v.hop[1].end_symbol;
    j.sel_out[t] = in1 ? 0 : -1;
    j.sel_in[t] = in2 ? 1 : j.sel_out[t];
    if (t < v.n_sym && !in1) j.uniform = false;
  }
  const int64_t rows = n_units * v.n_sc;
  int hw = static_cast<int>(std::thread::hardware_concurrency());
  if (hw < 1) hw = 1;
  if (n_threads <= 0 || n_threads > hw) n_threads = hw;
  const int64_t min_rows = 2048;                   // below ~1 MB of output per thread the hand-off costs more than it saves
  if (rows / min_rows < n_threads) n_threads = rows / min_rows < 1 ? 1 : static_cast<int>(rows / min_rows);
  if (n_threads == 1) {
    replicate_rows(j, 0, rows);
    return SRSCE_OK;
  }
  const int parts = n_threads * 4;
  HostPool::get().run(n_threads, parts, [&](int p) { replicate_rows(j, rows * p / parts, rows * (p + 1) / parts); });
  return SRSCE_OK;
}

}  // namespace srsce
