// engine.cpp — host side of libbamgpu.so: BGZF scan, the per-batch engine (K1 -> K2 -> K3) and the
// streaming Reader that mirrors bam_tools::Reader.  Compiled by nvcc as CUDA C++ (no kernels here).
//
// Reference code paths replaced (all relative to /root/reference/bam_tools/src):
//   util.rs:18-71          fetch_block / read_block_size / consume_trailer  -> bamgpu_scan_blocks
//   reader/readahead.rs    reader thread + rayon inflate tasks + ordering   -> feeder thread + pinned slots + streams
//   reader.rs:28-55        Reader::new                                      -> bamgpu_reader_new
//   reader.rs:87-98,154-200 read_header                                     -> bamgpu_read_header
//   reader.rs:57-81        read_record / append_record                      -> bamgpu_next_rec / bamgpu_append_record
//   reader.rs:114-152      impl Read                                        -> bamgpu_read
//   reader.rs:101-112      parse_reference_sequences                        -> bamgpu_parse_reference_sequences
#include <cuda_runtime.h>
#include <sched.h>
#include <unistd.h>
#include <cctype>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>
#include <array>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "bamgpu.h"
#include "inflate_core.cuh"
#include "kernels.h"

using namespace bamgpu;

namespace {

constexpr size_t OUT_LEAD_PAD = 64;   // readable bytes before the inflated stream (K1 window loads)
constexpr size_t OUT_TAIL_PAD = 64;   // readable bytes after it (unaligned field loads)
constexpr size_t IN_TAIL_PAD = 64;

inline uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// Growable device / pinned buffers ---------------------------------------------------------------
// cudaFree synchronises the whole device and now and then takes hundreds of milliseconds for a gigabyte; a Reader holds
// ~4 GB of device buffers, and a process that opens Readers one after the other (bam_binary runs, bench.py's e2e loop)
// paid that inside every open/close.  Released device buffers are parked per device and handed to the next taker.
struct DevCache {
  struct Item { void* p; size_t cap; int dev; };
  std::mutex mu;
  std::vector<Item> free_list;
  size_t bytes = 0;
  static constexpr size_t LIMIT = (size_t)48 << 30, ITEM_LIMIT = (size_t)8 << 30;
  void* take(size_t n, int dev, size_t* cap) {
    std::lock_guard<std::mutex> lk(mu);
    size_t best = (size_t)-1;
    for (size_t i = 0; i < free_list.size(); ++i)
      if (free_list[i].dev == dev && free_list[i].cap >= n && (best == (size_t)-1 || free_list[i].cap < free_list[best].cap)) best = i;
    if (best == (size_t)-1 || free_list[best].cap > n + n / 4 + (1 << 20)) return nullptr;
    void* p = free_list[best].p;
    *cap = free_list[best].cap;
    bytes -= *cap;
    free_list.erase(free_list.begin() + best);
    return p;
  }
  bool give(void* p, size_t cap, int dev) {
    std::lock_guard<std::mutex> lk(mu);
    if (cap > ITEM_LIMIT || bytes + cap > LIMIT) return false;
    free_list.push_back({p, cap, dev});
    bytes += cap;
    return true;
  }
  // Out of device memory: hand everything parked for `dev` back to the driver (the caller retries its allocation).
  size_t flush(int dev) {
    std::lock_guard<std::mutex> lk(mu);
    size_t freed = 0;
    for (size_t i = 0; i < free_list.size();) {
      if (free_list[i].dev == dev) { cudaFree(free_list[i].p); freed += free_list[i].cap; bytes -= free_list[i].cap; free_list.erase(free_list.begin() + i); }
      else ++i;
    }
    return freed;
  }
  size_t parked() { std::lock_guard<std::mutex> lk(mu); return bytes; }
};
DevCache& dev_cache() { static DevCache* c = new DevCache(); return *c; }
static size_t dev_cache_bytes() { return dev_cache().parked(); }

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int dev = -1;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    release();
    size_t want = n + n / 8 + 256;
    cudaGetDevice(&dev);
    size_t got = 0;
    if (void* q = dev_cache().take(n, dev, &got)) { p = q; cap = got; return cudaSuccess; }
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaErrorMemoryAllocation && dev_cache().flush(dev)) {   // parked buffers are memory too
      cudaGetLastError();
      e = cudaMalloc(&p, want);
    }
    if (e == cudaSuccess) cap = want; else p = nullptr;
    return e;
  }
  // Callers synchronise the streams that used the buffer before they release it (bamgpu_engine_free, bamgpu_reader_free).
  void release() {
    if (p && !dev_cache().give(p, cap, dev)) cudaFree(p);
    p = nullptr; cap = 0;
  }
  template <class T> T* as() const { return (T*)p; }
};
// Page-locking host memory is slow (about a second per few GB), and a Reader needs several hundred MB of it, so released
// pinned buffers are parked in a small process-wide cache and handed to the next Reader instead of being unpinned.
// Pinned buffers are tagged with the NUMA node they were allocated on (-1: wherever the calling thread ran): the
// mirrors of a GPU's batches should sit in the memory of the socket that GPU hangs off, or eight GPUs' device-to-host
// traffic funnels through the inter-socket link.
struct PinnedCache {
  struct Item { void* p; size_t cap; int node; };
  std::mutex mu;
  std::vector<Item> free_list;
  size_t bytes = 0;
  static constexpr size_t LIMIT = (size_t)32 << 30;
  void* take(size_t n, int node, size_t* cap) {
    std::lock_guard<std::mutex> lk(mu);
    size_t best = (size_t)-1;
    for (size_t i = 0; i < free_list.size(); ++i)
      if (free_list[i].node == node && free_list[i].cap >= n && (best == (size_t)-1 || free_list[i].cap < free_list[best].cap)) best = i;
    if (best == (size_t)-1 || free_list[best].cap > n + n / 4 + (1 << 20)) return nullptr;   // a snug fit only: a big buffer taken for a small need is missed later
    void* p = free_list[best].p;
    *cap = free_list[best].cap;
    bytes -= *cap;
    free_list.erase(free_list.begin() + best);
    return p;
  }
  bool give(void* p, size_t cap, int node) {
    std::lock_guard<std::mutex> lk(mu);
    if (bytes + cap > LIMIT) return false;
    free_list.push_back({p, cap, node});
    bytes += cap;
    return true;
  }
};
PinnedCache& pinned_cache() { static PinnedCache* c = new PinnedCache(); return *c; }

// NUMA node of a CUDA device and the CPUs next to it, from sysfs (best effort: -1 / empty when unknown).
struct NumaInfo { int node = -1; cpu_set_t cpus; bool have_cpus = false; };
const NumaInfo& device_numa(int dev) {
  static std::mutex mu;
  static NumaInfo table[64];
  static bool done[64];
  static NumaInfo none;
  if (dev < 0 || dev >= 64) return none;
  std::lock_guard<std::mutex> lk(mu);
  if (!done[dev]) {
    done[dev] = true;
    NumaInfo& ni = table[dev];
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, dev) == cudaSuccess) {
      for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
      std::string base = std::string("/sys/bus/pci/devices/") + bus;
      if (FILE* f = fopen((base + "/numa_node").c_str(), "r")) { if (fscanf(f, "%d", &ni.node) != 1) ni.node = -1; fclose(f); }
      if (FILE* f = fopen((base + "/local_cpulist").c_str(), "r")) {
        char line[4096] = {0};
        if (fgets(line, sizeof line, f)) {
          CPU_ZERO(&ni.cpus);
          for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
            int a = 0, b = 0;
            int k = sscanf(tok, "%d-%d", &a, &b);
            if (k == 1) b = a;
            if (k >= 1) for (int c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(c, &ni.cpus); ni.have_cpus = true; }
          }
        }
        fclose(f);
      }
    }
    cudaGetLastError();
  }
  return table[dev];
}

struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  int node = -1;
  // near_dev >= 0: allocate on the NUMA node of that GPU (the calling thread is moved next to it for the allocation:
  // cudaMallocHost places the pages where the caller runs)
  cudaError_t reserve(size_t n, int near_dev = -1) {
    if (n <= cap) return cudaSuccess;
    release();
    size_t want = n + n / 8 + 256;
    size_t got = 0;
    const NumaInfo* ni = near_dev >= 0 ? &device_numa(near_dev) : nullptr;
    node = (ni && ni->have_cpus) ? ni->node : -1;
    if (void* q = pinned_cache().take(n, node, &got)) { p = q; cap = got; return cudaSuccess; }
    cpu_set_t old;
    bool moved = false;
    if (node >= 0 && sched_getaffinity(0, sizeof old, &old) == 0 && sched_setaffinity(0, sizeof ni->cpus, &ni->cpus) == 0) moved = true;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;   // page-locking faults the pages in here and now, i.e. on the node the thread is on
    if (moved) sched_setaffinity(0, sizeof old, &old);
    return e;
  }
  void release() {
    if (p && !pinned_cache().give(p, cap, node)) cudaFreeHost(p);
    p = nullptr; cap = 0;
  }
  template <class T> T* as() const { return (T*)p; }
};

}  // namespace

// =================================================================================================
// Engine
// =================================================================================================
struct bamgpu_engine {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  std::string err;
  uint32_t default_flags = 0;
  int inflate_impl = 0;     // bamgpu_opts.inflate_impl: 0 lane pairs (default), 1 one warp per block

  // block table (SoA) staging + device
  PinBuf h_tab;
  DevBuf d_tab;
  DevBuf d_status, d_misc;  // d_misc: work counter (u32 @0), first_bad (u64 @8), ChainResult @64
  DevBuf d_k1ring;
  const uint32_t* d_crc = nullptr;   // the GPU's CRC tables (DeviceShared)
  cudaEvent_t ev_inflated = nullptr; // behind K1 + K2 + the status reduce of the batch in flight
  // inflated data.  The Reader runs with two buffers (pingpong): batch k+1 is inflated into one while batch k is
  // still being copied to the host out of the other.
  DevBuf d_outs[2];
  int cur = 0;
  bool pingpong = false;
  cudaStream_t d2h_stream = nullptr;
  cudaEvent_t ev_k3_done = nullptr;
  // launch/finish split of the inflate stage
  size_t pend_nb = 0;
  uint64_t pend_total = 0, pend_bytes_in = 0;
  bool inflate_launched = false;
  bool inflate_deferred = false;      // BAMGPU_DEFER_CRC: K2 and the status read run on k3_stream; engine_records collects the verdict
  uint64_t data_len = 0;    // bytes valid in d_out (after lead pad), carry included
  // Reader jobs ("lead mode"): a batch is inflated BEFORE the previous batch's record chain is known, so its blocks land at
  // a fixed offset `lead` and the carried-over tail of the previous batch is copied right in front of them afterwards
  // (engine_set_carry: a device-to-device or peer-to-peer copy): valid bytes are [data_start, data_len).
  uint64_t lead = 0;        // 0: classic placement (carry first, blocks behind it)
  uint64_t data_start = 0;  // first valid byte of d_out
  DevBuf d_carry;           // this batch's unconsumed tail, saved for whoever decodes the next batch (maybe another device)
  uint64_t carry_saved = 0;
  cudaEvent_t ev_carry = nullptr;
  bool own_d2h_stream = true;
  cudaStream_t k1_stream = nullptr;   // low priority: K1 of a Reader job
  cudaStream_t k3_stream = nullptr;   // bamgpu_engine_decode: the record kernels run here, next to K2 on the main stream
  uint32_t k1_grid = 0;               // CTAs of the K1 launch in flight (ev[1] fires when they are gone)
  bool numa_aware = false;            // multi-GPU Reader: host mirrors are allocated on this GPU's NUMA node
  cudaEvent_t ev_k1_in = nullptr;
  uint64_t cols_hint = 0;   // records expected in the next batch (sizes the columns without a host round trip)
  // shard-edge exchange (bamgpu_engine_edge_*)
  uint8_t* edge_inbox = nullptr;
  uint8_t* edge_left = nullptr;    // the left-hand neighbour's inbox, mapped here (peer pointer)
  uint8_t* edge_right = nullptr;   // the right-hand neighbour's inbox (we only write its ack flag)
  bool edge_left_ipc = false, edge_right_ipc = false, edge_has_right = false;
  uint32_t* edge_len_scratch = nullptr;
  bool sorted_valid = false;
  uint64_t sort_front = 0, sorted_front_used = 0;   // bytes kept free in front of the sorted output (bamgpu_sort_bam: the BAM header)
  uint32_t sorted_prefix = 0;
  DevBuf d_bgzf, d_bgzf_crc, d_bgzf_tmp;            // BGZF writer output + per-block CRCs + the compressor's slots
  uint64_t halo_len = 0;    // bytes copied behind them from the right-hand shard (multi-GPU)
  uint64_t carry_len = 0;   // bytes at the front of the next batch kept from this one
  uint64_t keep_from = 0;   // where those bytes currently sit
  int32_t n_ref = -1;
  // tiles
  DevBuf d_tiles;
  // One batch's columns on the device, its host mirrors and the bookkeeping of the D2H that fills them.  The Reader
  // (pingpong) alternates between two sets, so that batch k+1's record kernels and D2H can be queued while batch k's
  // D2H is still running and while the caller still holds batch k's pointers.
  struct Mirror {
    DevBuf d_cols;
    uint64_t cols_cap = 0;
    ColumnsDev cols{};
    PinBuf h_data, h_cols;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    bool pending = false;
  };
  Mirror mir[2];
  int msel = 0;
  // variable-length field decode (bamgpu_engine_decode_fields)
  DevBuf d_fscr, d_fout[3];
  PinBuf h_foff[3], h_fout[3];
  FieldsDev fields{};
  uint64_t fields_n = 0, fields_total[3] = {0, 0, 0};
  uint32_t fields_which = 0;
  // sort (bamgpu_engine_sort)
  DevBuf d_sort, d_perm, d_soff, d_sorted;
  PinBuf h_sorted, h_perm;
  uint64_t last_n = 0, last_bad_flags = 0;
  uint32_t last_flags = 0;
  bool last_valid = false;
  PinBuf h_res;
  uint64_t first_record_index = 0;
  // timing
  cudaEvent_t ev[8]{};
  bamgpu_stats stats{};

  uint8_t* out_ptr() const { return d_outs[cur].as<uint8_t>() + OUT_LEAD_PAD; }
};

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static int dbg_level() { const char* e = getenv("BAMGPU_DEBUG"); return e ? atoi(e) : 0; }

#define ECK(call)                                                                                   \
  do {                                                                                              \
    cudaError_t _e = (call);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      E->err = std::string(#call) + ": " + cudaGetErrorString(_e);                                  \
      return BAMGPU_ERR_CUDA;                                                                       \
    }                                                                                               \
  } while (0)

extern "C" const char* bamgpu_status_str(int s) {
  switch (s) {
    case BAMGPU_OK: return "ok";
    case BAMGPU_EOF: return "eof";
    case BAMGPU_ERR_ARG: return "bad argument";
    case BAMGPU_ERR_CUDA: return "CUDA error";
    case BAMGPU_ERR_NOMEM: return "out of memory";
    case BAMGPU_ERR_IO: return "read callback failed";
    case BAMGPU_ERR_BLOCK_TOO_SMALL: return "BGZF block smaller than 26 bytes";
    case BAMGPU_ERR_TRUNCATED_BLOCK: return "truncated BGZF block";
    case BAMGPU_ERR_INFLATE: return "corrupt DEFLATE stream";
    case BAMGPU_ERR_ISIZE: return "inflated size differs from ISIZE";
    case BAMGPU_ERR_CRC: return "CRC32 mismatch";
    case BAMGPU_ERR_BAD_MAGIC: return "invalid BAM header";
    case BAMGPU_ERR_SHORT_HEADER: return "short BAM header";
    case BAMGPU_ERR_SHORT_RECORD: return "record runs past end of stream";
    case BAMGPU_ERR_BUFFER_TOO_SMALL: return "destination buffer too small";
    case BAMGPU_ERR_STATE: return "call out of order";
    case BAMGPU_ERR_SORT_HI_MIXED: return "equal read names with HI on one record only (the reference panics, comparators.rs:90)";
    case BAMGPU_ERR_SORT_BAD_FLAGS: return "flag bits >= 0x1000 (the reference panics, comparators.rs:178)";
    case BAMGPU_ERR_SORT_BAD_TAGS: return "aux tags the reference's HI walk panics on (tags.rs:64-129)";
    case BAMGPU_ERR_TOO_LARGE: return "input does not fit in device memory as one batch";
    default: return "unknown";
  }
}

extern "C" const char* bamgpu_build_info(void) { return "sm_100a;libbamgpu " __DATE__ " " __TIME__; }

// -------------------------------------------------------------------------------------------------
// util.rs:18-71 restated for a memory buffer; keeps ISIZE + CRC32 instead of discarding them.
// -------------------------------------------------------------------------------------------------
extern "C" int bamgpu_scan_blocks(const uint8_t* buf, size_t len, int final, bamgpu_block* blocks, size_t cap,
                                  size_t* n_blocks, size_t* consumed, uint64_t* total_isize) {
  size_t pos = 0, n = 0;
  uint64_t o = 0;
  int rc = BAMGPU_OK;
  for (;;) {
    if (len - pos < 18) {                       // util.rs:58-60: short header read => Ok(0) => EOF
      if (final) rc = BAMGPU_EOF;
      break;
    }
    uint16_t bs = (uint16_t)(rd16(buf + pos + 16) + 1);  // util.rs:65, u16 wrap in release
    if (bs == 0) { rc = BAMGPU_EOF; break; }
    if (bs < 26) { rc = BAMGPU_ERR_BLOCK_TOO_SMALL; break; }  // util.rs:29-38
    if (len - pos < bs) {
      if (final) rc = BAMGPU_ERR_TRUNCATED_BLOCK;  // util.rs:42/44 read_exact fails
      break;
    }
    if (blocks && n < cap) {
      bamgpu_block& b = blocks[n];
      b.in_off = pos + 18;
      b.in_len = (uint32_t)bs - 26;
      b.out_off = o;
      b.isize = rd32(buf + pos + bs - 4);
      b.crc32 = rd32(buf + pos + bs - 8);
      b.reserved = 0;
    } else if (blocks) {
      break;  // table full: caller rescans the rest
    }
    o += rd32(buf + pos + bs - 4);
    ++n;
    pos += bs;
  }
  if (n_blocks) *n_blocks = n;
  if (consumed) *consumed = pos;
  if (total_isize) *total_isize = o;
  return rc;
}

// What every engine of one GPU shares and what costs milliseconds to obtain: the SM count, the stream priority range and
// the CRC tables (one kernel + one wait).  A Reader creates 8 engines per GPU inside the caller's timed Reader::new.
namespace {
struct DeviceShared {
  bool ready = false;
  int num_sms = 148, prio_lo = 0, prio_hi = 0;
  uint32_t* d_crc_tables = nullptr;   // never freed: lives as long as the process
};
DeviceShared* device_shared(int dev) {
  static std::mutex mu;
  static DeviceShared table[64];
  if (dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  DeviceShared& d = table[dev];
  if (!d.ready) {
    if (cudaSetDevice(dev) != cudaSuccess) return nullptr;
    cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetStreamPriorityRange(&d.prio_lo, &d.prio_hi);   // numerically lower = higher priority
    if (cudaMalloc((void**)&d.d_crc_tables, CRC_TABLE_WORDS * 4) != cudaSuccess) return nullptr;
    crc32_init_tables(d.d_crc_tables, 0);
    if (cudaDeviceSynchronize() != cudaSuccess) return nullptr;
    d.ready = true;
  }
  return &d;
}
}  // namespace

extern "C" bamgpu_engine* bamgpu_engine_new(const bamgpu_opts* opts) {
  bamgpu_engine* E = new bamgpu_engine();
  int dev = opts ? opts->device : -1;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    fprintf(stderr, "bamgpu: no CUDA device available; this library has no CPU fallback\n");
    delete E;
    return nullptr;
  }
  if (dev < 0) cudaGetDevice(&dev);
  if (cudaSetDevice(dev) != cudaSuccess) { delete E; return nullptr; }
  E->device = dev;
  DeviceShared* ds = device_shared(dev);
  if (!ds) { delete E; return nullptr; }
  E->num_sms = ds->num_sms;
  E->d_crc = ds->d_crc_tables;
  const int prio_lo = ds->prio_lo, prio_hi = ds->prio_hi;
  if (cudaStreamCreateWithPriority(&E->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { delete E; return nullptr; }
  if (cudaStreamCreateWithPriority(&E->k1_stream, cudaStreamNonBlocking, prio_lo) != cudaSuccess) { delete E; return nullptr; }
  if (cudaStreamCreateWithPriority(&E->k3_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { delete E; return nullptr; }
  cudaEventCreateWithFlags(&E->ev_k1_in, cudaEventDisableTiming);
  if (cudaStreamCreateWithFlags(&E->d2h_stream, cudaStreamNonBlocking) != cudaSuccess) { delete E; return nullptr; }
  cudaEventCreateWithFlags(&E->ev_k3_done, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&E->ev_carry, cudaEventDisableTiming);
  for (auto& M : E->mir) { cudaEventCreate(&M.ev_begin); cudaEventCreate(&M.ev_end); }
  for (auto& e : E->ev) cudaEventCreate(&e);
  E->default_flags = opts ? opts->flags : 0;
  E->inflate_impl = opts ? opts->inflate_impl : 0;
  if (const char* e = getenv("BAMGPU_INFLATE_IMPL")) E->inflate_impl = atoi(e);   // experiments: switch every engine of the process
  if (E->inflate_impl < 0 || E->inflate_impl > 1) { fprintf(stderr, "bamgpu: unknown inflate_impl %d\n", E->inflate_impl); delete E; return nullptr; }
  if (E->d_misc.reserve(8192) != cudaSuccess || E->h_res.reserve(4096) != cudaSuccess) {
    bamgpu_engine_free(E);
    return nullptr;
  }
  if (cudaMemsetAsync(E->d_misc.p, 0, 8192, E->stream) != cudaSuccess) { bamgpu_engine_free(E); return nullptr; }   // ordered before every use
  cudaEventCreateWithFlags(&E->ev_inflated, cudaEventDisableTiming);
  return E;
}

extern "C" void bamgpu_engine_free(bamgpu_engine* E) {
  if (!E) return;
  cudaSetDevice(E->device);
  bamgpu_engine_edge_close(E);   // while the stream still exists
  if (E->k1_stream) { cudaStreamSynchronize(E->k1_stream); cudaStreamDestroy(E->k1_stream); }
  if (E->k3_stream) { cudaStreamSynchronize(E->k3_stream); cudaStreamDestroy(E->k3_stream); }
  if (E->ev_k1_in) cudaEventDestroy(E->ev_k1_in);
  if (E->stream) { cudaStreamSynchronize(E->stream); cudaStreamDestroy(E->stream); }
  if (E->d2h_stream) { cudaStreamSynchronize(E->d2h_stream); if (E->own_d2h_stream) cudaStreamDestroy(E->d2h_stream); }
  if (E->ev_k3_done) cudaEventDestroy(E->ev_k3_done);
  if (E->ev_carry) cudaEventDestroy(E->ev_carry);
  for (auto& M : E->mir) { if (M.ev_begin) cudaEventDestroy(M.ev_begin); if (M.ev_end) cudaEventDestroy(M.ev_end); M.d_cols.release(); M.h_data.release(); M.h_cols.release(); }
  for (auto& e : E->ev) if (e) cudaEventDestroy(e);
  E->h_tab.release(); E->d_tab.release(); E->d_status.release(); E->d_misc.release();
  if (E->ev_inflated) cudaEventDestroy(E->ev_inflated);
  E->d_outs[0].release(); E->d_outs[1].release(); E->d_tiles.release(); E->d_k1ring.release(); E->h_res.release();
  E->d_carry.release();
  // the sort's scratch, permutation, offsets and gathered bodies (round 1 leaked these: ~18.7 GB per 50M-record sort)
  E->d_bgzf.release(); E->d_bgzf_crc.release(); E->d_bgzf_tmp.release();
  E->d_sort.release(); E->d_perm.release(); E->d_soff.release(); E->d_sorted.release(); E->h_sorted.release(); E->h_perm.release();
  E->d_fscr.release();
  for (int k = 0; k < 3; ++k) { E->d_fout[k].release(); E->h_foff[k].release(); E->h_fout[k].release(); }
  delete E;
}

extern "C" const char* bamgpu_engine_last_error(const bamgpu_engine* E) { return E ? E->err.c_str() : "null engine"; }

extern "C" void* bamgpu_device_alloc(bamgpu_engine* E, size_t bytes) {
  if (!E) return nullptr;
  cudaSetDevice(E->device);
  void* p = nullptr;
  if (cudaMalloc(&p, bytes + IN_TAIL_PAD) != cudaSuccess) return nullptr;
  cudaMemset((uint8_t*)p + bytes, 0, IN_TAIL_PAD);
  return p;
}
extern "C" void bamgpu_device_free(bamgpu_engine* E, void* p) { if (E) cudaSetDevice(E->device); if (p) cudaFree(p); }
extern "C" void* bamgpu_pinned_alloc(size_t bytes) { void* p = nullptr; return cudaMallocHost(&p, bytes ? bytes : 1) == cudaSuccess ? p : nullptr; }
extern "C" void bamgpu_pinned_free(void* p) { if (p) cudaFreeHost(p); }
extern "C" int bamgpu_memcpy_h2d(bamgpu_engine* E, void* dst, const void* src, size_t bytes) {
  if (!E) return BAMGPU_ERR_ARG;
  ECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, E->stream));
  ECK(cudaStreamSynchronize(E->stream));
  return BAMGPU_OK;
}
extern "C" int bamgpu_memcpy_d2h(bamgpu_engine* E, void* dst, const void* src, size_t bytes) {
  if (!E) return BAMGPU_ERR_ARG;
  ECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, E->stream));
  ECK(cudaStreamSynchronize(E->stream));
  return BAMGPU_OK;
}

namespace {


// Carves the SoA columns out of one allocation.
void carve_columns(uint8_t* base, uint64_t cap, ColumnsDev& c) {
  size_t o = 0;
  auto take = [&](size_t elem) { void* p = base + o; o += align_up(elem * cap, 256); return p; };
  c.rec_off = (uint64_t*)take(8); c.coord_key = (uint64_t*)take(8);
  c.rec_len = (uint32_t*)take(4); c.ref_id = (int32_t*)take(4); c.pos = (int32_t*)take(4); c.l_seq = (uint32_t*)take(4);
  c.next_ref_id = (int32_t*)take(4); c.next_pos = (int32_t*)take(4); c.tlen = (int32_t*)take(4);
  c.cigar_off = (uint32_t*)take(4); c.seq_off = (uint32_t*)take(4); c.qual_off = (uint32_t*)take(4); c.tags_off = (uint32_t*)take(4);
  c.hi_value = (int32_t*)take(4);
  c.bin = (uint16_t*)take(2); c.n_cigar_op = (uint16_t*)take(2); c.flag = (uint16_t*)take(2);
  c.l_read_name = (uint8_t*)take(1); c.mapq = (uint8_t*)take(1); c.strand = (uint8_t*)take(1); c.hi_present = (uint8_t*)take(1);
}
size_t columns_bytes(uint64_t cap) {
  // 2x8 + 12x4 + 3x2 + 4x1 = 74 bytes per record, each column padded to 256 B
  return (size_t)(74 * cap) + 21 * 256;
}

void fill_public_columns(const ColumnsDev& c, bamgpu_columns& o) {
  o.rec_off = c.rec_off; o.rec_len = c.rec_len; o.ref_id = c.ref_id; o.pos = c.pos; o.l_read_name = c.l_read_name;
  o.mapq = c.mapq; o.bin = c.bin; o.n_cigar_op = c.n_cigar_op; o.flag = c.flag; o.l_seq = c.l_seq;
  o.next_ref_id = c.next_ref_id; o.next_pos = c.next_pos; o.tlen = c.tlen; o.cigar_off = c.cigar_off; o.seq_off = c.seq_off;
  o.qual_off = c.qual_off; o.tags_off = c.tags_off; o.coord_key = c.coord_key; o.strand = c.strand;
  o.hi_present = c.hi_present; o.hi_value = c.hi_value;
}

int map_block_status(uint32_t st) {
  if (st == IST_CRC_MISMATCH) return BAMGPU_ERR_CRC;
  if (st == IST_OUTPUT_OVERRUN || st == IST_SIZE_MISMATCH) return BAMGPU_ERR_ISIZE;
  return BAMGPU_ERR_INFLATE;
}

}  // namespace

// Stage 1: inflate (+CRC) `n_blocks` blocks behind the carried-over bytes.  Split in two so that the Reader can put
// batch k+1's kernels on the stream before it waits for batch k's device-to-host copy:
//   engine_inflate_launch  enqueues carry move, block table upload, K1, K2, status reduce; no host wait
//   engine_inflate_finish  waits, reads the first failing block (if any), books the stage times
static int engine_wait_d2h(bamgpu_engine* E);
static int engine_inflate_finish(bamgpu_engine* E, cudaStream_t s);
static int engine_inflate_launch(bamgpu_engine* E, const uint8_t* d_comp, const bamgpu_block* blocks, size_t n_blocks,
                                 uint32_t flags, cudaStream_t s) {
  const double t_in = now_ms();
  ECK(cudaSetDevice(E->device));
  if (E->inflate_deferred) {   // a deferred verdict nobody asked for (no records call in between): collect it now
    E->inflate_deferred = false;
    ECK(cudaStreamSynchronize(s));
    int w = engine_inflate_finish(E, E->k3_stream);
    if (w < 0) return w;
  }
  // 1. the destination buffer (the other one when ping-ponging) gets the carry at its front; in lead mode the blocks
  //    land behind a fixed gap and the carry is copied into the gap later (engine_set_carry)
  const bool lead_mode = E->lead > 0;
  const uint64_t carry = lead_mode ? E->lead : E->carry_len;
  uint64_t total = 0, bytes_in = 0;
  for (size_t i = 0; i < n_blocks; ++i) { total += blocks[i].isize; bytes_in += blocks[i].in_len + 26; }
  const uint64_t new_len = carry + total;
  const int src = E->cur, dst = (E->pingpong && !lead_mode) ? (E->cur ^ 1) : E->cur;
  const uint8_t* src_ptr = E->d_outs[src].p ? E->d_outs[src].as<uint8_t>() + OUT_LEAD_PAD : nullptr;
  const size_t need = OUT_LEAD_PAD + new_len + BAMGPU_MAX_HALO + OUT_TAIL_PAD;
  if (lead_mode) {
    if (E->d_outs[dst].cap < need) {
      { int w = engine_wait_d2h(E); if (w < 0) return w; }
      ECK(E->d_outs[dst].reserve(need));
      ECK(cudaMemsetAsync(E->d_outs[dst].p, 0, OUT_LEAD_PAD, s));
    }
  } else if (E->d_outs[dst].cap < need) {
    DevBuf nb;
    ECK(nb.reserve(need));
    ECK(cudaMemsetAsync(nb.p, 0, OUT_LEAD_PAD, s));
    if (carry) ECK(cudaMemcpyAsync((uint8_t*)nb.p + OUT_LEAD_PAD, src_ptr + E->keep_from, carry, cudaMemcpyDeviceToDevice, s));
    ECK(cudaStreamSynchronize(s));
    if (dst == src) {
      E->d_outs[dst].release();
    } else {
      { int w = engine_wait_d2h(E); if (w < 0) return w; }  // nobody may still read the buffer we free
      E->d_outs[dst].release();
    }
    E->d_outs[dst] = nb;
  } else if (carry && (dst != src || E->keep_from != 0)) {
    uint8_t* dst_ptr = E->d_outs[dst].as<uint8_t>() + OUT_LEAD_PAD;
    if (dst != src || carry <= E->keep_from) {
      ECK(cudaMemcpyAsync(dst_ptr, src_ptr + E->keep_from, carry, cudaMemcpyDeviceToDevice, s));
    } else {  // same buffer and the ranges overlap: stage through a temporary (rare)
      DevBuf tmp;
      ECK(tmp.reserve(carry));
      ECK(cudaMemcpyAsync(tmp.p, src_ptr + E->keep_from, carry, cudaMemcpyDeviceToDevice, s));
      ECK(cudaMemcpyAsync(dst_ptr, tmp.p, carry, cudaMemcpyDeviceToDevice, s));
      ECK(cudaStreamSynchronize(s));
      tmp.release();
    }
  }
  E->data_start = lead_mode ? E->lead : 0;
  E->cur = dst;
  E->carry_len = 0;
  E->keep_from = 0;
  E->data_len = new_len;
  E->halo_len = 0;
  E->pend_nb = n_blocks;
  E->pend_total = total;
  E->pend_bytes_in = bytes_in;
  E->inflate_launched = true;
  ECK(cudaMemsetAsync(E->out_ptr() + new_len, 0, OUT_TAIL_PAD, s));
  if (n_blocks == 0) return BAMGPU_OK;
  if (n_blocks > 0xfffffff0ull) { E->err = "too many blocks in one batch"; return BAMGPU_ERR_ARG; }

  // 2. block table -> SoA in pinned memory -> device
  const size_t nb = n_blocks;
  const size_t o_in_off = 0, o_out_off = align_up(o_in_off + 8 * nb, 256), o_in_len = align_up(o_out_off + 8 * nb, 256),
               o_isize = align_up(o_in_len + 4 * nb, 256), o_crc = align_up(o_isize + 4 * nb, 256),
               tab_bytes = align_up(o_crc + 4 * nb, 256);
  ECK(E->h_tab.reserve(tab_bytes));
  ECK(E->d_tab.reserve(tab_bytes));
  ECK(E->d_status.reserve(4 * nb));
  {
    uint8_t* h = E->h_tab.as<uint8_t>();
    uint64_t* in_off = (uint64_t*)(h + o_in_off);
    uint64_t* out_off = (uint64_t*)(h + o_out_off);
    uint32_t* in_len = (uint32_t*)(h + o_in_len);
    uint32_t* isize = (uint32_t*)(h + o_isize);
    uint32_t* crc = (uint32_t*)(h + o_crc);
    uint64_t o = carry;
    for (size_t i = 0; i < nb; ++i) {
      in_off[i] = blocks[i].in_off; in_len[i] = blocks[i].in_len; out_off[i] = o; isize[i] = blocks[i].isize; crc[i] = blocks[i].crc32;
      o += blocks[i].isize;
    }
  }
  ECK(cudaMemcpyAsync(E->d_tab.p, E->h_tab.p, tab_bytes, cudaMemcpyHostToDevice, s));
  uint8_t* d = E->d_tab.as<uint8_t>();
  BlockTableDev t{(const uint64_t*)(d + o_in_off), (const uint32_t*)(d + o_in_len), (const uint64_t*)(d + o_out_off),
                  (const uint32_t*)(d + o_isize), (const uint32_t*)(d + o_crc), (uint32_t)nb};
  uint32_t* counter = E->d_misc.as<uint32_t>();
  unsigned long long* first_bad = (unsigned long long*)(E->d_misc.as<uint8_t>() + 8);
  ECK(cudaMemsetAsync(counter, 0, 4, s));
  ECK(cudaMemsetAsync(E->d_status.p, 0xff, 4 * nb, s));
  // Reader jobs: K1 goes onto the engine's LOW-priority stream, everything else stays on the high-priority one.  Several
  // batches' K1 are resident at once (persistent CTAs, ~10 ms each); when one of them ends, the CRC / record kernels
  // of the batch in front get the freed SM slots before the next batch's K1 CTAs do.
  cudaStream_t ks = (lead_mode && E->k1_stream && s == E->stream) ? E->k1_stream : s;
  if (ks != s) {
    ECK(cudaEventRecord(E->ev_k1_in, s));
    ECK(cudaStreamWaitEvent(ks, E->ev_k1_in, 0));
  }
  E->k1_grid = E->inflate_impl == 0 ? inflate_grid((uint32_t)nb, E->num_sms) : 0;
  ECK(cudaEventRecord(E->ev[0], ks));
  if (inflate_ring_bytes(E->num_sms)) ECK(E->d_k1ring.reserve(inflate_ring_bytes(E->num_sms)));
  if (E->inflate_impl == 1) launch_inflate_warp_per_block(d_comp, t, E->out_ptr(), E->d_status.as<uint32_t>(), counter, E->num_sms, ks);
  else launch_inflate(d_comp, t, E->out_ptr(), E->d_status.as<uint32_t>(), counter, E->d_k1ring.p, E->num_sms, ks);
  ECK(cudaEventRecord(E->ev[1], ks));
  if (ks != s) ECK(cudaStreamWaitEvent(s, E->ev[1], 0));
  E->stats.kernel_launches += 1;
  // BAMGPU_DEFER_CRC: everything behind K1 that only concerns the blocks' verdict goes onto the side stream
  cudaStream_t vs = s;
  E->inflate_deferred = false;
  if ((flags & BAMGPU_DEFER_CRC) && !lead_mode && E->k3_stream) {
    vs = E->k3_stream;
    ECK(cudaStreamWaitEvent(vs, E->ev[1], 0));
    E->inflate_deferred = true;
  }
  ECK(cudaEventRecord(E->ev[6], vs));
  if (!(flags & BAMGPU_NO_CRC)) {
    launch_crc32(E->out_ptr(), t, E->d_crc, E->d_status.as<uint32_t>(), E->num_sms, /*small_footprint=*/lead_mode, vs);
    E->stats.kernel_launches += 1;
  }
  ECK(cudaEventRecord(E->ev[2], vs));
  launch_status_reduce(E->d_status.as<uint32_t>(), (uint32_t)nb, first_bad, vs);
  E->stats.kernel_launches += 1;
  launch_copy_small_to_host(E->h_res.p, first_bad, 8, vs);
  ECK(cudaEventRecord(E->ev_inflated, vs));
  if (dbg_level() >= 2) fprintf(stderr, "[bamgpu] inflate_launch host %.2f ms (%zu blocks)\n", now_ms() - t_in, n_blocks);
  return BAMGPU_OK;
}

static int engine_inflate_finish(bamgpu_engine* E, cudaStream_t s) {
  if (!E->inflate_launched) return BAMGPU_OK;
  E->inflate_launched = false;
  const double t_in = now_ms();
  ECK(cudaSetDevice(E->device));
  ECK(cudaStreamSynchronize(s));
  ECK(cudaGetLastError());
  if (dbg_level() >= 2) fprintf(stderr, "[bamgpu] inflate_finish wait %.2f ms\n", now_ms() - t_in);
  if (E->pend_nb == 0) return BAMGPU_OK;
  float ms = 0;
  cudaEventElapsedTime(&ms, E->ev[0], E->ev[1]); E->stats.ms_inflate += ms;
  cudaEventElapsedTime(&ms, E->ev[6], E->ev[2]); E->stats.ms_crc += ms;
  E->stats.blocks += E->pend_nb;
  E->stats.bytes_out += E->pend_total;
  E->stats.bytes_in += E->pend_bytes_in;
  unsigned long long bad = *E->h_res.as<unsigned long long>();
  if (bad != ~0ull) {
    uint32_t idx = (uint32_t)(bad >> 32), st = (uint32_t)bad;
    char msg[160];
    snprintf(msg, sizeof msg, "BGZF block %u of this batch failed (device status %u): %s", idx, st,
             bamgpu_status_str(map_block_status(st)));
    E->err = msg;
    return map_block_status(st);
  }
  return BAMGPU_OK;
}

static int engine_inflate(bamgpu_engine* E, const uint8_t* d_comp, const bamgpu_block* blocks, size_t n_blocks,
                          uint32_t flags, cudaStream_t s) {
  int rc = engine_inflate_launch(E, d_comp, blocks, n_blocks, flags, s);
  if (rc != BAMGPU_OK) return rc;
  return engine_inflate_finish(E, s);
}

// Waits for the device-to-host copies engine_records left running (defer_d2h).
static int engine_wait_mirror(bamgpu_engine* E, int i) {
  bamgpu_engine::Mirror& M = E->mir[i];
  if (M.pending) {
    M.pending = false;
    ECK(cudaEventSynchronize(M.ev_end));
    float ms = 0;
    cudaEventElapsedTime(&ms, M.ev_begin, M.ev_end);
    E->stats.ms_d2h += ms;
  }
  return BAMGPU_OK;
}
static int engine_wait_d2h(bamgpu_engine* E) {
  for (int i = 0; i < 2; ++i) { int w = engine_wait_mirror(E, i); if (w < 0) return w; }
  return BAMGPU_OK;
}

// ---- carry between Reader jobs -------------------------------------------------------------------------------------
// engine_save_carry: copies this batch's unconsumed tail [keep_from, keep_from + carry_len) into the engine's small carry
// buffer (so the big output buffer can be reused at once) and records ev_carry behind the copy.
static int engine_save_carry(bamgpu_engine* E, cudaStream_t s) {
  ECK(cudaSetDevice(E->device));
  E->carry_saved = E->carry_len;
  if (E->carry_len) {
    ECK(E->d_carry.reserve(E->carry_len < 65536 ? 65536 : E->carry_len));
    ECK(cudaMemcpyAsync(E->d_carry.p, E->out_ptr() + E->keep_from, E->carry_len, cudaMemcpyDeviceToDevice, s));
  }
  ECK(cudaEventRecord(E->ev_carry, s));
  return BAMGPU_OK;
}

// engine_set_carry: puts `len` bytes of `from`'s saved tail right in front of this engine's freshly inflated blocks.
// Same device: device-to-device copy; another device: cudaMemcpyPeerAsync (NVLink when peer access is on) - this is
// the "boundary records exchanged peer-to-peer" step of the multi-GPU reader, one partial record per batch edge.
// K1 of this batch must have finished (the caller has waited for it).  If the tail is longer than the gap in front of
// the blocks (a record of more than `lead` bytes), the blocks move to a buffer with a larger gap first.
static int engine_set_carry(bamgpu_engine* E, bamgpu_engine* from, cudaStream_t s) {
  const uint64_t len = from ? from->carry_saved : 0;
  ECK(cudaSetDevice(E->device));
  if (len > E->lead) {
    const uint64_t new_lead = align_up(len + 4096, 1 << 16);
    const uint64_t total = E->data_len - E->lead;
    DevBuf nb;
    ECK(nb.reserve(OUT_LEAD_PAD + new_lead + total + BAMGPU_MAX_HALO + OUT_TAIL_PAD));
    ECK(cudaMemsetAsync(nb.p, 0, OUT_LEAD_PAD, s));
    ECK(cudaMemcpyAsync(nb.as<uint8_t>() + OUT_LEAD_PAD + new_lead, E->out_ptr() + E->lead, total + OUT_TAIL_PAD, cudaMemcpyDeviceToDevice, s));
    ECK(cudaStreamSynchronize(s));
    E->d_outs[E->cur].release();
    E->d_outs[E->cur] = nb;
    E->lead = new_lead;
    E->data_len = new_lead + total;
  }
  E->data_start = E->lead - len;
  if (len) {
    ECK(cudaStreamWaitEvent(s, from->ev_carry, 0));
    uint8_t* dst = E->out_ptr() + E->data_start;
    if (from->device == E->device) ECK(cudaMemcpyAsync(dst, from->d_carry.p, len, cudaMemcpyDeviceToDevice, s));
    else ECK(cudaMemcpyPeerAsync(dst, E->device, from->d_carry.p, from->device, len, s));
    E->stats.p2p_bytes += from->device == E->device ? 0 : len;
  }
  return BAMGPU_OK;
}

// Queues the device-to-host copies `flags` ask for (bodies / offsets / all columns) on the D2H stream, behind whatever
// is on `s` now.  Only what is not mirrored yet is copied, so a later call with more flags adds the missing pieces
// (a caller may switch from next_batch to next_rec in mid-stream).
static int engine_enqueue_mirrors(bamgpu_engine* E, bamgpu_engine::Mirror& M, uint32_t flags, uint64_t n_rec, bamgpu_batch* out,
                                  cudaStream_t s, bool* queued) {
  cudaStream_t cs = E->d2h_stream;
  const bool want_data = (flags & BAMGPU_COPY_DATA) && !out->h_data;
  const bool want_cols = (flags & BAMGPU_COPY_COLUMNS) && n_rec && !out->host.coord_key;
  const bool want_offs = (flags & BAMGPU_COPY_OFFSETS) && !(flags & BAMGPU_COPY_COLUMNS) && n_rec && !out->host.rec_off;
  const bool any_copy = want_data || want_cols || want_offs;
  if (queued) *queued = any_copy;
  if (!any_copy) return BAMGPU_OK;
  ECK(cudaEventRecord(E->ev_k3_done, s));
  ECK(cudaStreamWaitEvent(cs, E->ev_k3_done, 0));
  if (!M.pending) ECK(cudaEventRecord(M.ev_begin, cs));
  if (want_data) {
    ECK(M.h_data.reserve(E->data_len + 64, E->numa_aware ? E->device : -1));
    // lead mode: only [data_start, data_len) holds bytes; the mirror keeps the device offsets (rec_off indexes both)
    const uint64_t from = E->data_start & ~(uint64_t)63;
    if (E->data_len > from) ECK(cudaMemcpyAsync(M.h_data.as<uint8_t>() + from, E->out_ptr() + from, E->data_len - from, cudaMemcpyDeviceToHost, cs));
    out->h_data = M.h_data.as<uint8_t>();
  }
  if (want_offs) {
    // what next_rec() needs besides the bodies: where each one starts and how long it is
    ECK(M.h_cols.reserve(columns_bytes(M.cols_cap), E->numa_aware ? E->device : -1));
    ColumnsDev hc;
    carve_columns(M.h_cols.as<uint8_t>(), M.cols_cap, hc);
    ECK(cudaMemcpyAsync(hc.rec_off, M.cols.rec_off, 8 * n_rec, cudaMemcpyDeviceToHost, cs));
    ECK(cudaMemcpyAsync(hc.rec_len, M.cols.rec_len, 4 * n_rec, cudaMemcpyDeviceToHost, cs));
    out->host.rec_off = hc.rec_off;
    out->host.rec_len = hc.rec_len;
  }
  if (want_cols) {
    size_t bytes = columns_bytes(M.cols_cap);
    ECK(M.h_cols.reserve(bytes, E->numa_aware ? E->device : -1));
    // copy only the used prefix of every column
    ColumnsDev hc;
    carve_columns(M.h_cols.as<uint8_t>(), M.cols_cap, hc);
    auto cp = [&](void* dst, const void* src, size_t elem) {
      return cudaMemcpyAsync(dst, src, elem * n_rec, cudaMemcpyDeviceToHost, cs);
    };
    ECK(cp(hc.rec_off, M.cols.rec_off, 8)); ECK(cp(hc.coord_key, M.cols.coord_key, 8));
    ECK(cp(hc.rec_len, M.cols.rec_len, 4)); ECK(cp(hc.ref_id, M.cols.ref_id, 4)); ECK(cp(hc.pos, M.cols.pos, 4));
    ECK(cp(hc.l_seq, M.cols.l_seq, 4)); ECK(cp(hc.next_ref_id, M.cols.next_ref_id, 4)); ECK(cp(hc.next_pos, M.cols.next_pos, 4));
    ECK(cp(hc.tlen, M.cols.tlen, 4)); ECK(cp(hc.cigar_off, M.cols.cigar_off, 4)); ECK(cp(hc.seq_off, M.cols.seq_off, 4));
    ECK(cp(hc.qual_off, M.cols.qual_off, 4)); ECK(cp(hc.tags_off, M.cols.tags_off, 4)); ECK(cp(hc.hi_value, M.cols.hi_value, 4));
    ECK(cp(hc.bin, M.cols.bin, 2)); ECK(cp(hc.n_cigar_op, M.cols.n_cigar_op, 2)); ECK(cp(hc.flag, M.cols.flag, 2));
    ECK(cp(hc.l_read_name, M.cols.l_read_name, 1)); ECK(cp(hc.mapq, M.cols.mapq, 1)); ECK(cp(hc.strand, M.cols.strand, 1));
    ECK(cp(hc.hi_present, M.cols.hi_present, 1));
    fill_public_columns(hc, out->host);
  }
  ECK(cudaEventRecord(M.ev_end, cs));
  M.pending = true;
  return BAMGPU_OK;
}

// Stage 2: record chain + columns over d_out[0 .. data_len).
static int engine_records(bamgpu_engine* E, uint64_t start, int final, uint32_t flags, cudaStream_t s,
                          bamgpu_batch* out, uint64_t* tail_off_out, bool defer_d2h = false) {
  const double t_in = now_ms();
  ECK(cudaSetDevice(E->device));
  // the mirror set this batch goes to: its previous D2H (two batches ago when ping-ponging) must be over
  if (E->pingpong) E->msel ^= 1;
  bamgpu_engine::Mirror& M = E->mir[E->msel];
  if (E->pingpong) { int w = engine_wait_mirror(E, E->msel); if (w < 0) return w; }
  else { int w = engine_wait_d2h(E); if (w < 0) return w; }
  memset(out, 0, sizeof *out);
  out->data = E->out_ptr();
  out->data_len = E->data_len;
  out->data_start = E->data_start;
  out->first_record_index = E->first_record_index;
  out->chain_start = start;
  ChainResult* d_res = (ChainResult*)(E->d_misc.as<uint8_t>() + 64);
  ChainResult* h_res = (ChainResult*)(E->h_res.as<uint8_t>() + 64);
  uint64_t n_rec = 0, tail_off = start;
  uint32_t tail_kind = EXIT_TAIL;
  uint64_t bad_flags = 0;
  ECK(cudaEventRecord(E->ev[3], s));
  if (!(flags & BAMGPU_NO_RECORDS) && start < E->data_len) {
    RecordParams p{E->out_ptr(), E->data_len + E->halo_len, E->data_len, start, E->n_ref, (flags & BAMGPU_SPECULATIVE_START) ? 0 : 1};
    uint32_t n_tiles = (uint32_t)((E->data_len + TILE_BYTES - 1) / TILE_BYTES);
    // tile arrays: entry u64, exit u64, base u64, count u32, kind u32, nxt u32, jump u32, jump2 u32, reach u8
    size_t nt = n_tiles;
    size_t o_entry = 0, o_exit = align_up(o_entry + 8 * nt, 256), o_base = align_up(o_exit + 8 * nt, 256),
           o_count = align_up(o_base + 8 * nt, 256), o_kind = align_up(o_count + 4 * nt, 256),
           o_nxt = align_up(o_kind + 4 * nt, 256), o_jump = align_up(o_nxt + 4 * nt, 256),
           o_jump2 = align_up(o_jump + 4 * nt, 256), o_reach = align_up(o_jump2 + 4 * nt, 256),
           o_ckpt = align_up(o_reach + nt, 256), tiles_bytes = align_up(o_ckpt + 2 * (size_t)CK_PER_TILE * nt, 256);
    ECK(E->d_tiles.reserve(tiles_bytes));
    uint8_t* tb = E->d_tiles.as<uint8_t>();
    TileArrays t{(uint64_t*)(tb + o_entry), (uint64_t*)(tb + o_exit), (uint32_t*)(tb + o_count), (uint32_t*)(tb + o_kind),
                 (uint32_t*)(tb + o_nxt), (uint32_t*)(tb + o_jump), (uint32_t*)(tb + o_jump2), (uint8_t*)(tb + o_reach),
                 (uint64_t*)(tb + o_base), n_tiles, (uint16_t*)(tb + o_ckpt)};
    // Columns are sized BEFORE the record count is known (from what the previous batch needed, or ~256 bytes per record the
    // first time), so that walk -> resolve -> emit -> extract go onto the stream back to back and the host waits ONCE per
    // batch (round 1: three waits).  If the guess was too small the last two kernels run again with room for all.
    if (M.cols_cap == 0) {
      uint64_t cap = E->cols_hint ? E->cols_hint : (E->data_len - start) / 256 + 4096;
      ECK(M.d_cols.reserve(columns_bytes(cap)));
      M.cols_cap = cap;
      carve_columns(M.d_cols.as<uint8_t>(), cap, M.cols);
    }
    const bool want_hi = (flags & BAMGPU_WANT_HI) != 0;
    launch_tile_walk(p, t, s);
    ECK(launch_chain_resolve(p, t, d_res, (ResolveScratch*)(E->d_misc.as<uint8_t>() + 256), E->num_sms, s));
    launch_tile_emit(p, t, M.cols.rec_off, M.cols.rec_len, M.cols_cap, s);
    launch_extract(p, M.cols, M.cols_cap, want_hi, d_res, /*dev_count=*/true, s);
    E->stats.kernel_launches += 4;
    launch_copy_small_to_host(h_res, d_res, sizeof(ChainResult), s);
    ECK(cudaStreamSynchronize(s));
    ECK(cudaGetLastError());
    if (h_res->broken_tile != 0xffffffffu) { E->err = "internal: record chain left the buffer"; return BAMGPU_ERR_STATE; }
    out->chain_start = h_res->chain_start;
    out->chain_repairs = h_res->rounds;
    n_rec = h_res->n_records;
    tail_off = h_res->tail_off;
    tail_kind = h_res->tail_kind;
    bad_flags = h_res->bad_flags;
    if (n_rec > M.cols_cap) {
      { int w = engine_wait_mirror(E, E->msel); if (w < 0) return w; }
      uint64_t cap = n_rec + n_rec / 8 + 1024;
      ECK(M.d_cols.reserve(columns_bytes(cap)));
      M.cols_cap = cap;
      carve_columns(M.d_cols.as<uint8_t>(), cap, M.cols);
      ECK(cudaMemsetAsync(&d_res->bad_flags, 0, sizeof(unsigned long long), s));
      launch_tile_emit(p, t, M.cols.rec_off, M.cols.rec_len, cap, s);
      launch_extract(p, M.cols, n_rec, want_hi, d_res, /*dev_count=*/false, s);
      E->stats.kernel_launches += 2;
      launch_copy_small_to_host(h_res, d_res, sizeof(ChainResult), s);
      ECK(cudaStreamSynchronize(s));
      ECK(cudaGetLastError());
      bad_flags = h_res->bad_flags;
    }
  } else if (flags & BAMGPU_NO_RECORDS) {
    tail_off = E->data_len;
  }
  ECK(cudaEventRecord(E->ev[4], s));
  // End-of-chain semantics (reader.rs:63-81).
  int rc = BAMGPU_OK;
  if (!(flags & BAMGPU_NO_RECORDS)) {
    if (tail_kind == EXIT_INCOMPLETE && final) {
      E->err = "Failed to read. (record body runs past the end of the stream, reader.rs:69-70)";
      rc = BAMGPU_ERR_SHORT_RECORD;
    }
    if (tail_kind == EXIT_ZERO) rc = BAMGPU_EOF;             // block_size == 0 ends iteration even mid-stream
    if (tail_kind == EXIT_TAIL && final) rc = BAMGPU_EOF;    // 0..3 trailing bytes: silent EOF
  }
  out->n_records = n_rec;
  out->bad_flag_records = bad_flags;
  E->last_n = n_rec;
  E->last_bad_flags = bad_flags;
  E->last_flags = flags;
  E->last_valid = !(flags & BAMGPU_NO_RECORDS);
  E->fields_which = 0;   // decoded fields belong to the previous batch
  E->sorted_valid = false;
  fill_public_columns(M.cols, out->dev);
  // host mirrors: on their own stream, behind the record kernels
  bool any_copy = false;
  { int w = engine_enqueue_mirrors(E, M, flags, n_rec, out, s, &any_copy); if (w < 0) return w; }
  ECK(cudaStreamSynchronize(s));
  ECK(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, E->ev[3], E->ev[4]); E->stats.ms_records += ms;
  if (any_copy && !defer_d2h) { int w = engine_wait_mirror(E, E->msel); if (w < 0) return w; }
  if (dbg_level() >= 2) fprintf(stderr, "[bamgpu] records host wall %.2f ms\n", now_ms() - t_in);
  E->stats.records += n_rec;
  E->stats.batches += 1;
  E->first_record_index += n_rec;
  if (tail_off_out) *tail_off_out = tail_off;
  // what the next batch must see in front of its own bytes
  if (!final && rc == BAMGPU_OK && !(flags & BAMGPU_NO_RECORDS) && E->halo_len == 0) {
    E->keep_from = tail_off;
    E->carry_len = E->data_len - tail_off;
  } else {
    E->keep_from = 0;
    E->carry_len = 0;
  }
  if (E->lead > 0) {   // Reader job: park the tail where the next batch's engine (maybe on another GPU) can fetch it
    int w = engine_save_carry(E, s);
    if (w < 0) return w;
  }
  return rc;
}

extern "C" int bamgpu_engine_decode(bamgpu_engine* E, const uint8_t* d_comp, const bamgpu_block* blocks, size_t n_blocks,
                                    uint64_t stream_start, int final, uint32_t flags, void* cuda_stream,
                                    bamgpu_batch* out, uint64_t* tail_off) {
  if (!E || !out || (n_blocks && (!blocks || !d_comp))) return BAMGPU_ERR_ARG;
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : E->stream;
  flags |= E->default_flags;
  if (n_blocks && !(flags & (BAMGPU_NO_CRC | BAMGPU_NO_RECORDS)) && E->k3_stream && !getenv("BAMGPU_NO_K2_K3_OVERLAP")) {
    // K2 (CRC32 of every block: HBM-bound) and K3 (record chain + columns: latency-bound) both only READ the inflated bytes: K3 goes
    // onto a second stream behind K1 and runs next to K2 (50 M records: 4.8 + 4.8 ms one after the other).  A CRC / inflate error
    // still wins: the status is checked before anything is handed out.
    int rc = engine_inflate_launch(E, d_comp, blocks, n_blocks, flags, s);
    if (rc != BAMGPU_OK) return rc;
    ECK(cudaStreamWaitEvent(E->k3_stream, E->ev[1], 0));          // behind K1
    const int rrc = engine_records(E, stream_start, final, flags, E->k3_stream, out, tail_off);
    rc = engine_inflate_finish(E, s);
    if (rc != BAMGPU_OK) { E->last_valid = false; memset(out, 0, sizeof *out); return rc; }
    return rrc;
  }
  int rc = engine_inflate(E, d_comp, blocks, n_blocks, flags, s);
  if (rc != BAMGPU_OK) return rc;
  return engine_records(E, stream_start, final, flags, s, out, tail_off);
}

extern "C" int bamgpu_engine_inflate(bamgpu_engine* E, const uint8_t* d_comp, const bamgpu_block* blocks, size_t n_blocks,
                                     uint32_t flags, void* cuda_stream) {
  if (!E || (n_blocks && (!blocks || !d_comp))) return BAMGPU_ERR_ARG;
  E->carry_len = 0;
  E->keep_from = 0;
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : E->stream;
  flags |= E->default_flags;
  if (flags & BAMGPU_DEFER_CRC) {
    int rc = engine_inflate_launch(E, d_comp, blocks, n_blocks, flags, s);
    if (rc != BAMGPU_OK || E->inflate_deferred) return rc;   // deferred: bamgpu_engine_records collects the verdict
    return engine_inflate_finish(E, s);
  }
  return engine_inflate(E, d_comp, blocks, n_blocks, flags, s);
}

extern "C" int bamgpu_engine_data(const bamgpu_engine* E, const uint8_t** d_data, uint64_t* len) {
  if (!E || !d_data || !len) return BAMGPU_ERR_ARG;
  *d_data = E->out_ptr();
  *len = E->data_len;
  return BAMGPU_OK;
}

extern "C" int bamgpu_engine_append_halo(bamgpu_engine* E, const uint8_t* d_src, size_t len, void* cuda_stream) {
  if (!E || (len && !d_src) || len > BAMGPU_MAX_HALO) return BAMGPU_ERR_ARG;
  if (!E->d_outs[E->cur].p) return BAMGPU_ERR_STATE;
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : E->stream;
  ECK(cudaSetDevice(E->device));
  if (len) ECK(cudaMemcpyAsync(E->out_ptr() + E->data_len, d_src, len, cudaMemcpyDeviceToDevice, s));
  ECK(cudaMemsetAsync(E->out_ptr() + E->data_len + len, 0, OUT_TAIL_PAD, s));
  E->halo_len = len;
  return BAMGPU_OK;
}

extern "C" int bamgpu_engine_records(bamgpu_engine* E, uint64_t stream_start, int final, uint32_t flags, void* cuda_stream,
                                     bamgpu_batch* out, uint64_t* tail_off) {
  if (!E || !out) return BAMGPU_ERR_ARG;
  const int rc = engine_records(E, stream_start, final, flags | E->default_flags, cuda_stream ? (cudaStream_t)cuda_stream : E->stream, out, tail_off);
  if (E->inflate_deferred) {   // BAMGPU_DEFER_CRC: the blocks' verdict (K2 ran beside the exchange and the record kernels) comes first
    E->inflate_deferred = false;
    const int w = engine_inflate_finish(E, E->k3_stream);
    if (w < 0) { E->last_valid = false; memset(out, 0, sizeof *out); return w; }
  }
  return rc;
}

extern "C" int bamgpu_engine_stats(const bamgpu_engine* E, bamgpu_stats* out) {
  if (!E || !out) return BAMGPU_ERR_ARG;
  *out = E->stats;
  out->ms_total = out->ms_h2d + out->ms_inflate + out->ms_crc + out->ms_records + out->ms_d2h + out->ms_sort + out->ms_gather;
  return BAMGPU_OK;
}
extern "C" void bamgpu_engine_reset_stats(bamgpu_engine* E) { if (E) E->stats = bamgpu_stats{}; }

extern "C" void bamgpu_engine_set_n_ref(bamgpu_engine* E, int32_t n) { if (E) E->n_ref = n; }

// -------------------------------------------------------------------------------------------------
// Sort (sorting/sort.rs:138-178 sort_bam, the part behind the reader): stable sort of the last batch's records by the
// chosen comparator, bodies gathered in that order.
// -------------------------------------------------------------------------------------------------
extern "C" int bamgpu_engine_sort(bamgpu_engine* E, int sort_by, uint32_t flags, void* cuda_stream, bamgpu_sorted* out) {
  if (!E || !out || sort_by < BAMGPU_SORT_NAME || sort_by > BAMGPU_SORT_COORDINATES_AND_STRAND) return BAMGPU_ERR_ARG;
  memset(out, 0, sizeof *out);
  if (!E->last_valid) { E->err = "sort: no decoded batch on this engine"; return BAMGPU_ERR_STATE; }
  if (sort_by == BAMGPU_SORT_NAME_AND_MATCH_MATES && !(E->last_flags & BAMGPU_WANT_HI)) {
    E->err = "sort: NameAndMatchMates needs the batch decoded with BAMGPU_WANT_HI";
    return BAMGPU_ERR_STATE;
  }
  if (sort_by == BAMGPU_SORT_COORDINATES_AND_STRAND && E->last_bad_flags) {
    E->err = "sort: " + std::to_string(E->last_bad_flags) + " record(s) with flag bits >= 0x1000 (Flags::from_bits().unwrap() panics, comparators.rs:178)";
    return BAMGPU_ERR_SORT_BAD_FLAGS;
  }
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : E->stream;
  ECK(cudaSetDevice(E->device));
  { int w = engine_wait_d2h(E); if (w < 0) return w; }
  const uint64_t n = E->last_n;
  if (n >= (1ull << 31)) { E->err = "sort: more than 2^31 records in one batch"; return BAMGPU_ERR_TOO_LARGE; }
  out->n_records = n;
  ECK(E->d_sort.reserve(sort_scratch_bytes(n)));
  ECK(E->d_perm.reserve(4 * (n + 1)));
  ECK(E->d_soff.reserve(8 * (n + 2)));
  SortResult* d_sr = (SortResult*)(E->d_misc.as<uint8_t>() + 192);
  SortResult* h_sr = (SortResult*)(E->h_res.as<uint8_t>() + 192);
  int launches = 0;
  ECK(cudaEventRecord(E->ev[3], s));
  ECK(sort_permutation(E->out_ptr(), E->mir[E->msel].cols, n, sort_by, E->d_sort.p, E->d_perm.as<uint32_t>(), d_sr, &launches, s));
  ECK(cudaMemcpyAsync(h_sr, d_sr, sizeof(SortResult), cudaMemcpyDeviceToHost, s));
  ECK(cudaEventRecord(E->ev[4], s));
  uint64_t total = 0;
  const uint32_t prefix = (flags & BAMGPU_SORT_EMIT_BLOCK_SIZE) ? 4u : 0u;   // (u32 block_size, body)* instead of bodies
  const uint64_t front = E->sort_front;                                      // room in front of the output (a BAM header)
  ECK(sorted_offsets(E->mir[E->msel].cols, E->d_perm.as<uint32_t>(), n, E->d_sort.p, E->d_soff.as<uint64_t>(), &total, prefix, s));   // synchronises
  if (sort_by == BAMGPU_SORT_NAME_AND_MATCH_MATES) {
    if (h_sr->hi_bad_tags) { E->err = "sort: aux tags the reference's HI walk panics on"; return BAMGPU_ERR_SORT_BAD_TAGS; }
    if (h_sr->hi_mixed) {
      E->err = "sort: " + std::to_string(h_sr->hi_mixed) + " pair(s) of equal read names with HI on one record only (Option::unwrap() panics, comparators.rs:90)";
      return BAMGPU_ERR_SORT_HI_MIXED;
    }
  }
  ECK(E->d_sorted.reserve(front + total + 64));
  ECK(gather_sorted_bodies(E->out_ptr(), E->mir[E->msel].cols, E->d_perm.as<uint32_t>(), n, E->d_sort.p, E->d_soff.as<uint64_t>(),
                           E->d_sorted.as<uint8_t>() + front, prefix, s));
  E->sorted_prefix = prefix;
  E->sorted_front_used = front;
  ECK(cudaEventRecord(E->ev[5], s));
  launches += 3;
  out->bodies_len = total;
  out->perm = E->d_perm.as<uint32_t>();
  out->body_off = E->d_soff.as<uint64_t>();
  out->bodies = E->d_sorted.as<uint8_t>() + front;
  E->sorted_valid = true;
  if ((flags & BAMGPU_COPY_COLUMNS) && n) {
    ECK(E->h_perm.reserve(4 * n));
    ECK(cudaMemcpyAsync(E->h_perm.p, E->d_perm.p, 4 * n, cudaMemcpyDeviceToHost, s));
    out->h_perm = E->h_perm.as<uint32_t>();
  }
  if ((flags & BAMGPU_COPY_DATA) && total) {
    ECK(E->h_sorted.reserve(total));
    ECK(cudaMemcpyAsync(E->h_sorted.p, E->d_sorted.as<uint8_t>() + front, total, cudaMemcpyDeviceToHost, s));
    out->h_bodies = E->h_sorted.as<uint8_t>();
  }
  ECK(cudaStreamSynchronize(s));
  ECK(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, E->ev[3], E->ev[4]); E->stats.ms_sort += ms;
  cudaEventElapsedTime(&ms, E->ev[4], E->ev[5]); E->stats.ms_gather += ms;
  E->stats.kernel_launches += launches;
  return BAMGPU_OK;
}


// -------------------------------------------------------------------------------------------------
// Variable-length field decode (csrc/fields.cu; record/bamrawrecord.rs:126-157,208-273).
// -------------------------------------------------------------------------------------------------
extern "C" int bamgpu_engine_decode_fields(bamgpu_engine* E, uint32_t which, uint32_t flags, void* cuda_stream, bamgpu_fields* out) {
  if (!E || !out || !(which & 7u) || (which & ~7u)) return BAMGPU_ERR_ARG;
  memset(out, 0, sizeof *out);
  if (!E->last_valid) { E->err = "decode_fields: no decoded batch on this engine"; return BAMGPU_ERR_STATE; }
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : E->stream;
  ECK(cudaSetDevice(E->device));
  { int w = engine_wait_d2h(E); if (w < 0) return w; }
  const uint64_t n = E->last_n;
  if (n >= (1ull << 31)) { E->err = "decode_fields: more than 2^31 records in one batch"; return BAMGPU_ERR_TOO_LARGE; }
  E->fields_which = 0;
  ECK(E->d_fscr.reserve(fields_scratch_bytes(n)));
  unsigned long long* d_bad = (unsigned long long*)(E->d_misc.as<uint8_t>() + 4096 + 64);
  uint64_t totals[3], bad[3];
  ECK(cudaEventRecord(E->ev[3], s));
  ECK(fields_measure(E->out_ptr(), E->mir[E->msel].cols, n, which, E->d_fscr.p, &E->fields, d_bad, totals, bad, E->num_sms, s));
  ECK(cudaEventRecord(E->ev[4], s));
  for (int k = 0; k < 3; ++k) {
    E->fields_total[k] = totals[k];
    if (which & (1u << k)) { ECK(E->d_fout[k].reserve(totals[k] + 64)); E->fields.out[k] = E->d_fout[k].as<uint8_t>(); }
  }
  launch_fields_decode(E->out_ptr(), E->mir[E->msel].cols, n, which, E->fields, E->num_sms, s);
  ECK(cudaEventRecord(E->ev[5], s));
  bamgpu_field* fo[3] = {&out->cigar, &out->seq, &out->qual};
  for (int k = 0; k < 3; ++k) {
    if (!(which & (1u << k))) continue;
    fo[k]->total = totals[k];
    fo[k]->bad = bad[k];
    fo[k]->off = E->fields.off[k];
    fo[k]->bytes = E->fields.out[k];
    if (flags & BAMGPU_COPY_DATA) {
      ECK(E->h_foff[k].reserve(8 * (n + 1)));
      ECK(cudaMemcpyAsync(E->h_foff[k].p, E->fields.off[k], 8 * (n + 1), cudaMemcpyDeviceToHost, s));
      fo[k]->h_off = E->h_foff[k].as<uint64_t>();
      ECK(E->h_fout[k].reserve(totals[k] + 1));
      if (totals[k]) ECK(cudaMemcpyAsync(E->h_fout[k].p, E->fields.out[k], totals[k], cudaMemcpyDeviceToHost, s));
      fo[k]->h_bytes = E->h_fout[k].as<uint8_t>();
    }
  }
  ECK(cudaStreamSynchronize(s));
  ECK(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, E->ev[3], E->ev[4]); out->ms_measure = ms;
  cudaEventElapsedTime(&ms, E->ev[4], E->ev[5]); out->ms_decode = ms;
  out->n_records = n;
  E->fields_n = n;
  E->fields_which = which;
  E->stats.kernel_launches += n ? 2 + __builtin_popcount(which) : 0;
  return BAMGPU_OK;
}

static int engine_read_digest(bamgpu_engine* E, cudaStream_t s, uint64_t n, bamgpu_digest* out);
extern "C" int bamgpu_engine_field_digest(bamgpu_engine* E, uint32_t field, void* cuda_stream, bamgpu_digest* out) {
  if (!E || !out || (field != 1 && field != 2 && field != 4)) return BAMGPU_ERR_ARG;
  if (!(E->fields_which & field)) { E->err = "field_digest: that field has not been decoded on this engine"; return BAMGPU_ERR_STATE; }
  const int k = field == 1 ? 0 : (field == 2 ? 1 : 2);
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : E->stream;
  ECK(cudaSetDevice(E->device));
  launch_digest_ragged(E->fields.out[k], E->fields.off[k], E->fields_n, (unsigned long long*)(E->d_misc.as<uint8_t>() + 4096), E->num_sms, s);
  E->stats.kernel_launches += E->fields_n ? 1 : 0;
  const int rc = engine_read_digest(E, s, E->fields_n, out);
  out->body_bytes = E->fields_total[k];
  return rc;
}

// -------------------------------------------------------------------------------------------------
// Digests (csrc/digest.cu): size-independent parity checks (DIGEST SPEC).
// -------------------------------------------------------------------------------------------------
static int engine_read_digest(bamgpu_engine* E, cudaStream_t s, uint64_t n, bamgpu_digest* out) {
  unsigned long long* acc = (unsigned long long*)(E->d_misc.as<uint8_t>() + 4096);
  unsigned long long h[4];
  ECK(cudaMemcpyAsync(h, acc, sizeof h, cudaMemcpyDeviceToHost, s));
  ECK(cudaStreamSynchronize(s));
  ECK(cudaGetLastError());
  out->n_records = n;
  out->body_bytes = h[1];
  out->fields = h[2];
  out->bodies = h[3];
  return BAMGPU_OK;
}

extern "C" int bamgpu_engine_digest(bamgpu_engine* E, uint64_t index_base, uint64_t stream_base, void* cuda_stream, bamgpu_digest* out) {
  if (!E || !out) return BAMGPU_ERR_ARG;
  if (!E->last_valid) { E->err = "digest: no decoded batch on this engine"; return BAMGPU_ERR_STATE; }
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : E->stream;
  ECK(cudaSetDevice(E->device));
  launch_digest_records(E->out_ptr(), E->mir[E->msel].cols, E->last_n, index_base, stream_base,
                        (unsigned long long*)(E->d_misc.as<uint8_t>() + 4096), E->num_sms, s);
  E->stats.kernel_launches += E->last_n ? 2 : 0;
  return engine_read_digest(E, s, E->last_n, out);
}

extern "C" int bamgpu_engine_sorted_digest(bamgpu_engine* E, void* cuda_stream, bamgpu_digest* out) {
  if (!E || !out) return BAMGPU_ERR_ARG;
  if (!E->sorted_valid) { E->err = "digest: no sort result on this engine"; return BAMGPU_ERR_STATE; }
  if (E->sorted_prefix) { E->err = "digest: the sort output carries block_size prefixes; the DIGEST SPEC is defined on bodies"; return BAMGPU_ERR_STATE; }
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : E->stream;
  ECK(cudaSetDevice(E->device));
  launch_digest_sorted(E->d_sorted.as<uint8_t>() + E->sorted_front_used, E->d_soff.as<uint64_t>(), E->d_perm.as<uint32_t>(), E->last_n,
                       (unsigned long long*)(E->d_misc.as<uint8_t>() + 4096), E->num_sms, s);
  E->stats.kernel_launches += E->last_n ? 2 : 0;
  return engine_read_digest(E, s, E->last_n, out);
}

// -------------------------------------------------------------------------------------------------
// Shard-edge exchange over peer memory: see include/bamgpu.h and csrc/edge.cu.
// Inbox layout (device memory of the shard that RECEIVES):  halo[2][BAMGPU_MAX_HALO] | flags
//   flags[0] data_step  written by the right-hand neighbour after its halo copy for `step` landed in halo[step & 1]
//   flags[1] ack_step   written by the LEFT-hand neighbour into OUR inbox: it has consumed our halo of that step
//   flags[2..3] halo length of each buffer (written with the data)
// -------------------------------------------------------------------------------------------------
namespace {
constexpr size_t EDGE_FLAGS_OFF = 2 * (size_t)BAMGPU_MAX_HALO;
constexpr size_t EDGE_BYTES = EDGE_FLAGS_OFF + 4096;
}

extern "C" int bamgpu_engine_edge_open(bamgpu_engine* E, bamgpu_edge_handle* mine) {
  if (!E || !mine) return BAMGPU_ERR_ARG;
  ECK(cudaSetDevice(E->device));
  if (!E->edge_inbox) {
    ECK(cudaMalloc(&E->edge_inbox, EDGE_BYTES));   // a plain allocation: cudaIpcGetMemHandle wants the base of one
    ECK(cudaMemset(E->edge_inbox, 0, EDGE_BYTES));
    ECK(cudaMalloc((void**)&E->edge_len_scratch, 64));
    ECK(cudaMemset(E->edge_len_scratch, 0, 64));
  }
  memset(mine, 0, sizeof *mine);
  cudaIpcMemHandle_t h;
  ECK(cudaIpcGetMemHandle(&h, E->edge_inbox));
  static_assert(sizeof h == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(mine->ipc, &h, 64);
  mine->ptr = (uint64_t)(uintptr_t)E->edge_inbox;
  mine->device = E->device;
  mine->pid = (int32_t)getpid();
  return BAMGPU_OK;
}

static int edge_map(bamgpu_engine* E, const bamgpu_edge_handle* h, uint8_t** out, bool* ipc) {
  *out = nullptr;
  *ipc = false;
  if (!h) return BAMGPU_OK;
  if (h->pid == (int32_t)getpid()) {   // same process: the pointer itself, with peer access switched on
    if (h->device != E->device) {
      int can = 0;
      cudaDeviceCanAccessPeer(&can, E->device, h->device);
      if (!can) { E->err = "edge: no peer access between the two GPUs"; return BAMGPU_ERR_CUDA; }
      cudaError_t e = cudaDeviceEnablePeerAccess(h->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { E->err = cudaGetErrorString(e); return BAMGPU_ERR_CUDA; }
      cudaGetLastError();
    }
    *out = (uint8_t*)(uintptr_t)h->ptr;
    return BAMGPU_OK;
  }
  cudaIpcMemHandle_t ih;
  memcpy(&ih, h->ipc, 64);
  void* p = nullptr;
  ECK(cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
  *out = (uint8_t*)p;
  *ipc = true;
  return BAMGPU_OK;
}

extern "C" int bamgpu_engine_edge_connect(bamgpu_engine* E, const bamgpu_edge_handle* left, const bamgpu_edge_handle* right) {
  if (!E) return BAMGPU_ERR_ARG;
  if (!E->edge_inbox) { E->err = "edge_connect before edge_open"; return BAMGPU_ERR_STATE; }
  ECK(cudaSetDevice(E->device));
  int rc = edge_map(E, left, &E->edge_left, &E->edge_left_ipc);
  if (rc < 0) return rc;
  rc = edge_map(E, right, &E->edge_right, &E->edge_right_ipc);
  if (rc < 0) return rc;
  E->edge_has_right = right != nullptr;
  return BAMGPU_OK;
}

extern "C" void bamgpu_engine_edge_close(bamgpu_engine* E) {
  if (!E || !E->edge_inbox) return;
  cudaSetDevice(E->device);
  if (E->stream) cudaStreamSynchronize(E->stream);
  if (E->edge_left && E->edge_left_ipc) cudaIpcCloseMemHandle(E->edge_left);
  if (E->edge_right && E->edge_right_ipc) cudaIpcCloseMemHandle(E->edge_right);
  cudaFree(E->edge_inbox);
  cudaFree(E->edge_len_scratch);
  E->edge_inbox = E->edge_left = E->edge_right = nullptr;
  E->edge_len_scratch = nullptr;
  cudaGetLastError();
}

extern "C" int bamgpu_engine_edge_exchange(bamgpu_engine* E, uint64_t step, void* cuda_stream) {
  if (!E || step == 0) return BAMGPU_ERR_ARG;
  if (!E->edge_inbox) { E->err = "edge_exchange before edge_open"; return BAMGPU_ERR_STATE; }
  if (!E->d_outs[E->cur].p) return BAMGPU_ERR_STATE;
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : E->stream;
  ECK(cudaSetDevice(E->device));
  unsigned long long* my_flags = (unsigned long long*)(E->edge_inbox + EDGE_FLAGS_OFF);
  uint32_t* timed_out = (uint32_t*)(E->d_misc.as<uint8_t>() + 4160);
  const int buf = (int)(step & 1);
  if (E->edge_left) {
    // my head -> the left neighbour's inbox (peer write over NVLink), then publish the step there.
    // Its buffer `buf` held step-2's halo: wait until the neighbour has acknowledged that one.
    unsigned long long* left_flags = (unsigned long long*)(E->edge_left + EDGE_FLAGS_OFF);
    if (step > 2) launch_wait_flag(&my_flags[1], step - 2, timed_out, s);
    const size_t m = (size_t)std::min<uint64_t>(E->data_len, BAMGPU_MAX_HALO);
    if (m) ECK(cudaMemcpyAsync(E->edge_left + (size_t)buf * BAMGPU_MAX_HALO, E->out_ptr(), m, cudaMemcpyDefault, s));
    launch_set_flag(&left_flags[2 + buf], m, s);
    launch_set_flag(&left_flags[0], step, s);
    E->stats.p2p_bytes += m;
    E->stats.kernel_launches += 2 + (step > 2 ? 1 : 0);
  }
  if (E->edge_has_right) {
    // wait (on the device) for the right neighbour's halo of this step, append it, acknowledge
    unsigned long long* right_flags = (unsigned long long*)(E->edge_right + EDGE_FLAGS_OFF);
    launch_wait_flag(&my_flags[0], step, timed_out, s);
    // the halo length travels with the data; it is read on the host only to size the local copy
    ECK(cudaMemcpyAsync(E->h_res.as<uint8_t>() + 512, &my_flags[2 + buf], 8, cudaMemcpyDeviceToHost, s));
    ECK(cudaMemcpyAsync(E->h_res.as<uint8_t>() + 520, timed_out, 4, cudaMemcpyDeviceToHost, s));
    ECK(cudaStreamSynchronize(s));
    if (*(uint32_t*)(E->h_res.as<uint8_t>() + 520)) { E->err = "edge: the right-hand shard never delivered its halo (20 s)"; return BAMGPU_ERR_STATE; }
    const uint64_t m = *(uint64_t*)(E->h_res.as<uint8_t>() + 512);
    if (m > BAMGPU_MAX_HALO) { E->err = "edge: bad halo length"; return BAMGPU_ERR_STATE; }
    int rc = bamgpu_engine_append_halo(E, E->edge_inbox + (size_t)buf * BAMGPU_MAX_HALO, (size_t)m, s);
    if (rc < 0) return rc;
    launch_set_flag(&right_flags[1], step, s);
    E->stats.kernel_launches += 2;
  } else {
    int rc = bamgpu_engine_append_halo(E, nullptr, 0, s);
    if (rc < 0) return rc;
  }
  return BAMGPU_OK;
}


// -------------------------------------------------------------------------------------------------
// BGZF writer with stored blocks (csrc/bgzf_write.cu)
// -------------------------------------------------------------------------------------------------
extern "C" size_t bamgpu_bgzf_bound(size_t len) {
  const size_t full = len / BGZF_STORE_PAYLOAD, rem = len % BGZF_STORE_PAYLOAD;
  return full * BGZF_STORE_BLOCK + (rem ? 18 + 5 + rem + 8 : 0) + 28;
}

extern "C" int bamgpu_engine_bgzf_store(bamgpu_engine* E, const uint8_t* d_src, size_t len, uint8_t* d_out, size_t out_cap,
                                        size_t* out_len, int eof_marker, void* cuda_stream) {
  if (!E || (len && !d_src) || !d_out || !out_len) return BAMGPU_ERR_ARG;
  const size_t need = bamgpu_bgzf_bound(len) - (eof_marker ? 0 : 28);
  if (out_cap < need) { E->err = "bgzf_store: output buffer too small"; return BAMGPU_ERR_BUFFER_TOO_SMALL; }
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : E->stream;
  ECK(cudaSetDevice(E->device));
  const size_t n_blocks = (len + BGZF_STORE_PAYLOAD - 1) / BGZF_STORE_PAYLOAD;
  ECK(E->d_bgzf_crc.reserve(4 * (n_blocks + 1)));
  cudaEvent_t& e0 = E->ev[3];
  cudaEvent_t& e1 = E->ev[4];
  ECK(cudaEventRecord(e0, s));
  launch_crc32_chunks(d_src, len, BGZF_STORE_PAYLOAD, E->d_crc, E->d_bgzf_crc.as<uint32_t>(), E->num_sms, s);
  ECK(bgzf_store_blocks(d_src, len, E->d_bgzf_crc.as<uint32_t>(), d_out, eof_marker != 0, s));
  ECK(cudaEventRecord(e1, s));
  ECK(cudaStreamSynchronize(s));
  ECK(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  E->stats.ms_bgzf_write += ms;
  E->stats.kernel_launches += 2;
  *out_len = need;
  return BAMGPU_OK;
}

// K6 (csrc/deflate.cu): the same framing with COMPRESSED payloads.  level <= 0: stored blocks (bamgpu_engine_bgzf_store).
extern "C" int bamgpu_engine_bgzf_write(bamgpu_engine* E, const uint8_t* d_src, size_t len, uint8_t* d_out, size_t out_cap,
                                        size_t* out_len, int eof_marker, int level, void* cuda_stream) {
  if (level <= 0) return bamgpu_engine_bgzf_store(E, d_src, len, d_out, out_cap, out_len, eof_marker, cuda_stream);
  if (!E || (len && !d_src) || !d_out || !out_len) return BAMGPU_ERR_ARG;
  if (out_cap < bamgpu_bgzf_bound(len) - (eof_marker ? 0 : 28)) { E->err = "bgzf_write: output buffer too small"; return BAMGPU_ERR_BUFFER_TOO_SMALL; }
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : E->stream;
  ECK(cudaSetDevice(E->device));
  const size_t n_blocks = (len + BGZF_STORE_PAYLOAD - 1) / BGZF_STORE_PAYLOAD;
  ECK(E->d_bgzf_crc.reserve(4 * (n_blocks + 1)));
  ECK(E->d_bgzf_tmp.reserve(bgzf_deflate_scratch_bytes(len, E->num_sms)));
  cudaEvent_t& e0 = E->ev[3];
  cudaEvent_t& e1 = E->ev[4];
  ECK(cudaEventRecord(e0, s));
  launch_crc32_chunks(d_src, len, BGZF_STORE_PAYLOAD, E->d_crc, E->d_bgzf_crc.as<uint32_t>(), E->num_sms, s);
  uint64_t total = 0;
  ECK(bgzf_deflate_blocks(d_src, len, E->d_bgzf_crc.as<uint32_t>(), E->d_bgzf_tmp.p, d_out, eof_marker != 0, level, &total, E->num_sms, s));
  ECK(cudaEventRecord(e1, s));
  ECK(cudaStreamSynchronize(s));
  ECK(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  E->stats.ms_bgzf_write += ms;
  E->stats.kernel_launches += 5;
  *out_len = (size_t)total;
  return BAMGPU_OK;
}

// =================================================================================================
// Reader
// =================================================================================================
// The streaming drop-in for bam_tools::Reader.  Replaces readahead.rs:51-143 (reader thread + n-2 inflate workers +
// ordering thread, n Blocks in flight) by:
//   * a FEEDER thread (the reference's reading thread, readahead.rs:88-118): cuts the source into batches of whole BGZF
//     blocks, scans headers AND footers, and starts each batch's host-to-device copy on the copy stream of the GPU the
//     batch goes to (batch k -> GPU k mod D: the file is sharded by BGZF block ranges);
//   * JOBS: every GPU owns J engines (= J batches in flight).  A batch is inflated (K1 + K2) as soon as its bytes are
//     on the device - before the batches in front of it have been parsed, and next to the K1 of other batches, which is
//     what fills the GPU: one 256 MiB batch is 13 k BGZF blocks, a quarter of K1's 56.8 k streams;
//   * the record chain (reader.rs:63-81) is serial across batches: batch k can only be parsed once batch k-1 has told
//     where its last complete record ends.  That partial record is copied in front of batch k's blocks - from whichever
//     GPU decoded batch k-1, peer to peer (engine_set_carry) - and K3 runs; this is the only inter-GPU traffic;
//   * device-to-host copies of one GPU go through ONE stream in batch order; up to D batches beyond the one handed out are
//     staged, so every GPU's link is busy.
namespace {

struct Slot {
  PinBuf host;               // compressed bytes as read from the source (unused when the source itself is pinned memory)
  DevBuf dev;                // same bytes on the device
  size_t len = 0;            // bytes covered by whole blocks
  std::vector<bamgpu_block> blocks;
  bool final = false;        // no more input after this slot
  int status = BAMGPU_OK;    // scan / io status: an error here surfaces AFTER the records of the blocks scanned before it
  uint64_t total_isize = 0;
  cudaEvent_t copied = nullptr;
};

struct DevCtx {
  int dev = 0;
  cudaStream_t h2d = nullptr, d2h = nullptr;
  std::vector<Slot> slots;
  std::deque<int> free_slots;   // guarded by the reader's mutex
};

struct Job {
  bamgpu_engine* E = nullptr;
  int dev_i = 0, slot = -1;
  uint64_t seq = 0;
  enum St { FREE, LAUNCHED, STAGED, OUT } st = FREE;
  bamgpu_batch batch{};
  bool final = false, carry_set = false;
  int input_status = BAMGPU_OK;   // Slot::status
  uint64_t stream_new = 0;        // position in the inflated stream of this batch's first new byte
  uint64_t new_bytes = 0;
};

struct ReadyItem { uint64_t seq; int dev_i, slot; };

}  // namespace

struct bamgpu_reader {
  std::vector<DevCtx> devs;
  std::vector<Job> jobs;            // devs.size() * J; batch seq -> jobs[(seq % D) * J + (seq / D) % J]
  int J = 4;
  bamgpu_read_fn read = nullptr;
  void* ctx = nullptr;
  size_t thread_num = 3;
  uint64_t track_total = 0;
  uint64_t bytes_fed = 0;
  uint32_t flags = 0;
  size_t batch_bytes = 256u << 20;
  size_t first_batch_bytes = 0;     // 0: like every other batch
  std::string err;
  // memory source
  const uint8_t* mem = nullptr;
  size_t mem_len = 0, mem_pos = 0;
  bool mem_pinned = false;          // the caller's buffer is page-locked: H2D straight out of it, no staging copy
  // feeder
  std::thread feeder;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<ReadyItem> ready;
  bool stop = false;
  // consumer state
  uint64_t n_launched = 0, n_staged = 0, n_out = 0;   // batches whose K1 is queued / whose records are parsed / handed out
  uint64_t launched_stream = 0;     // inflated bytes of all launched batches
  uint64_t rec_index = 0;
  bool input_done = false, header_done = false, eof = false, raw_mode = false, no_more = false, records_mode = false;
  int sticky = BAMGPU_OK, sticky_next = BAMGPU_OK;
  std::vector<uint8_t> header;      // without magic
  size_t off_to_refs = 0;
  int32_t n_ref = -1;
  uint64_t next_start_rel = 0;      // chain start of the next staged batch, relative to its data_start (header batch only)
  Job* cur = nullptr;               // the batch the caller holds
  bamgpu_batch batch{};
  bool batch_valid = false;
  uint64_t rec_cursor = 0;          // next record of `batch` to hand out
  uint64_t byte_cursor = 0;         // raw mode: next byte of batch.h_data
  bamgpu_stats base_stats{};

  Job& job_of(uint64_t seq) { size_t D = devs.size(); return jobs[(seq % D) * J + (seq / D) % J]; }
};

static long mem_read(void* ctx, uint8_t* dst, size_t cap) {
  bamgpu_reader* R = (bamgpu_reader*)ctx;
  size_t n = std::min(cap, R->mem_len - R->mem_pos);
  memcpy(dst, R->mem + R->mem_pos, n);
  R->mem_pos += n;
  return (long)n;
}

// Feeder thread: readahead.rs:88-118's reading thread, but it batches whole blocks, scans their headers AND footers,
// and starts the H2D copy on the copy stream of the GPU the batch is for.
static void feeder_main(bamgpu_reader* R) {
  std::vector<uint8_t> carry;  // partial block left over from the previous slot
  bool src_eof = false;
  const size_t D = R->devs.size();
  for (uint64_t seq = 0;; ++seq) {
    const int di = (int)(seq % D);
    DevCtx& dc = R->devs[di];
    int si;
    {
      std::unique_lock<std::mutex> lk(R->mu);
      R->cv.wait(lk, [&] { return R->stop || !dc.free_slots.empty(); });
      if (R->stop) return;
      si = dc.free_slots.front();
      dc.free_slots.pop_front();
    }
    cudaSetDevice(dc.dev);
    Slot& S = dc.slots[si];
    S.blocks.clear();
    S.status = BAMGPU_OK;
    S.final = false;
    S.len = 0;
    S.total_isize = 0;
    // the first batch is small when the caller left the batch size to us: the header and the first records are there
    // after one short copy, and the big batches' copies and K1 start that much earlier
    // ... and the following ones double up to the full size (32, 64, 128, 256 MiB): each is inflated and parsed by the time the
    // device-to-host copy of the one in front ends, so that link - the bound of the whole pass - has work from the first
    // ~13 ms on (with a full-size second batch it idled until that batch's 13 k blocks were through K1 + K3: ~35 ms)
    const size_t bb = (R->first_batch_bytes && seq < 8) ? std::min(R->batch_bytes, R->first_batch_bytes << seq) : R->batch_bytes;
    const size_t cap = R->batch_bytes + 65536 + 64;
    const uint8_t* src = nullptr;   // where the batch's bytes are on the host
    size_t fill = 0;
    if (R->mem && R->mem_pinned) {
      // zero-copy: whole blocks straight out of the caller's page-locked buffer
      src = R->mem + R->mem_pos;
      fill = std::min(bb, R->mem_len - R->mem_pos);
      if (R->mem_pos + fill >= R->mem_len) src_eof = true;
      if (S.dev.reserve(cap + IN_TAIL_PAD) != cudaSuccess) { S.status = BAMGPU_ERR_NOMEM; S.final = true; }
    } else if (S.host.reserve(cap + IN_TAIL_PAD, D > 1 ? dc.dev : -1) != cudaSuccess || S.dev.reserve(cap + IN_TAIL_PAD) != cudaSuccess) {
      S.status = BAMGPU_ERR_NOMEM; S.final = true;
    } else {
      uint8_t* h = S.host.as<uint8_t>();
      src = h;
      fill = carry.size();
      if (fill) memcpy(h, carry.data(), fill);
      carry.clear();
      if (R->mem && !src_eof) {
        // in-memory source: fill the pinned slot with several threads (thread_num, like the reference's inflate pool)
        size_t want = std::min(bb > fill ? bb - fill : 0, R->mem_len - R->mem_pos);
        size_t nt = std::min<size_t>(std::max<size_t>(R->thread_num, 1), 16);
        if (want < (8u << 20)) nt = 1;
        std::vector<std::thread> th;
        size_t per = (want + nt - 1) / nt;
        for (size_t t = 1; t < nt; ++t) {
          size_t o = t * per, n = o < want ? std::min(per, want - o) : 0;
          if (n) th.emplace_back([=] { memcpy(h + fill + o, R->mem + R->mem_pos + o, n); });
        }
        memcpy(h + fill, R->mem + R->mem_pos, std::min(per, want));
        for (auto& x : th) x.join();
        R->mem_pos += want;
        R->bytes_fed += want;
        fill += want;
        if (R->mem_pos >= R->mem_len) src_eof = true;
      }
      while (!R->mem && !src_eof && fill < bb) {
        long n = R->read(R->ctx, h + fill, std::min<size_t>(bb + 65536 - fill, 8u << 20));
        if (n < 0) { S.status = BAMGPU_ERR_IO; src_eof = true; break; }   // the blocks read so far are still decoded
        if (n == 0) { src_eof = true; break; }
        fill += (size_t)n;
        R->bytes_fed += (uint64_t)n;
      }
    }
    if (S.status != BAMGPU_ERR_NOMEM) {
      size_t nb = 0, consumed = 0;
      uint64_t tot = 0;
      // count, then fill.  A scan error (block too small / truncated) keeps the whole blocks in front of it: the
      // reference hands out their records before its reading thread dies on the bad block (readahead.rs:92).
      int rc = bamgpu_scan_blocks(src, fill, src_eof ? 1 : 0, nullptr, 0, &nb, &consumed, &tot);
      S.blocks.resize(nb);
      if (nb) bamgpu_scan_blocks(src, fill, src_eof ? 1 : 0, S.blocks.data(), nb, &nb, &consumed, &tot);
      S.len = consumed;
      S.total_isize = tot;
      if (rc < 0) { if (S.status == BAMGPU_OK) S.status = rc; S.final = true; }
      else if (rc == BAMGPU_EOF || src_eof) S.final = true;   // silent EOF (short header / BSIZE wrap) ends the stream
      else if (R->mem && R->mem_pinned) { /* the next batch starts at the first unconsumed byte */ }
      else carry.assign(src + consumed, src + fill);
      if (R->mem && R->mem_pinned) { R->mem_pos += consumed; R->bytes_fed += consumed; if (consumed == 0 && !S.final) S.final = true; }
      if (S.len) {
        cudaMemcpyAsync(S.dev.p, src, S.len, cudaMemcpyHostToDevice, dc.h2d);
        cudaMemsetAsync(S.dev.as<uint8_t>() + S.len, 0, IN_TAIL_PAD, dc.h2d);
      }
      cudaEventRecord(S.copied, dc.h2d);
    }
    bool fin = S.final;
    {
      std::lock_guard<std::mutex> lk(R->mu);
      R->ready.push_back({seq, di, si});
    }
    R->cv.notify_all();
    if (fin) return;
  }
}

static void reader_destroy(bamgpu_reader* R);

static bamgpu_reader* reader_create(bamgpu_read_fn read, void* ctx, size_t thread_num, const uint64_t* track_progress,
                                    const bamgpu_opts* opts) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    fprintf(stderr, "bamgpu: no CUDA device available; this library has no CPU fallback\n");
    return nullptr;
  }
  bamgpu_reader* R = new bamgpu_reader();
  std::vector<int> devices;
  if (opts && opts->n_devices > 0) {
    for (int i = 0; i < opts->n_devices && i < BAMGPU_MAX_DEVICES; ++i) devices.push_back(opts->devices[i]);
  } else {
    int dev = opts ? opts->device : -1;
    if (dev < 0) cudaGetDevice(&dev);
    devices.push_back(dev);
  }
  for (int d : devices) if (d < 0 || d >= count) { delete R; return nullptr; }
  const size_t D = devices.size();
  // batches in flight per GPU: one GPU needs ~5 K1 launches of 256 MiB in flight to fill its 56.8 k streams; with several
  // GPUs the host link is the limit long before that, and every job costs streams, events and buffers at open / close
  R->J = (opts && opts->jobs_per_device > 0) ? std::min(opts->jobs_per_device, 16) : (D > 1 ? 4 : 8);
  if (R->J < 2) R->J = 2;
  if (const char* e = getenv("BAMGPU_JOBS")) { int v = atoi(e); if (v >= 2 && v <= 16) R->J = v; }
  R->read = read;
  R->ctx = ctx;
  unsigned hc = std::thread::hardware_concurrency();
  if (hc && thread_num > hc) thread_num = hc;      // reader.rs:33-35
  R->thread_num = std::max<size_t>(thread_num, 3);  // readahead.rs:53
  R->track_total = track_progress ? *track_progress : 0;
  R->flags = opts ? opts->flags : 0;
  R->batch_bytes = D > 1 ? ((size_t)96 << 20) : ((size_t)256 << 20);
  if (const char* e = getenv("BAMGPU_BATCH_MIB")) { long v = atol(e); if (v > 0) R->batch_bytes = (size_t)v << 20; }
  if (opts && opts->batch_bytes) R->batch_bytes = (size_t)opts->batch_bytes;
  else R->first_batch_bytes = std::min(R->batch_bytes, (size_t)32 << 20);
  if (R->batch_bytes < 65536) R->batch_bytes = 65536;
  // peer access between every pair of GPUs (the batch-edge records travel GPU to GPU)
  for (size_t a = 0; a < D; ++a)
    for (size_t b = 0; b < D; ++b) {
      if (devices[a] == devices[b]) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
      if (can) { cudaSetDevice(devices[a]); cudaDeviceEnablePeerAccess(devices[b], 0); cudaGetLastError(); }
    }
  R->devs.resize(D);
  R->jobs.resize(D * (size_t)R->J);
  const int n_slots = R->J;   // a batch keeps its input slot until its K1 is over: as many slots as batches in flight
  // per-GPU set-up (streams, events, J engines) runs on one thread per GPU: Reader::new is inside the caller's clock
  std::vector<int> ok(D, 1);
  auto setup = [&](size_t i) {
    DevCtx& dc = R->devs[i];
    dc.dev = devices[i];
    if (cudaSetDevice(dc.dev) != cudaSuccess ||
        cudaStreamCreateWithFlags(&dc.h2d, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&dc.d2h, cudaStreamNonBlocking) != cudaSuccess) { ok[i] = 0; return; }
    dc.slots.resize(n_slots);
    for (int k = 0; k < n_slots; ++k) {
      cudaEventCreateWithFlags(&dc.slots[k].copied, cudaEventDisableTiming);
      dc.free_slots.push_back(k);
    }
    for (int j = 0; j < R->J; ++j) {
      bamgpu_opts eo{};
      eo.device = dc.dev;
      eo.flags = R->flags;
      eo.inflate_impl = opts ? opts->inflate_impl : 0;
      bamgpu_engine* E = bamgpu_engine_new(&eo);
      if (!E) { ok[i] = 0; return; }
      // one D2H stream per GPU: copies leave in batch order
      cudaSetDevice(dc.dev);
      cudaStreamDestroy(E->d2h_stream);
      E->d2h_stream = dc.d2h;
      E->own_d2h_stream = false;
      E->lead = BAMGPU_MAX_HALO;
      E->numa_aware = D > 1;
      Job& jb = R->jobs[i * R->J + j];
      jb.E = E;
      jb.dev_i = (int)i;
    }
  };
  if (D == 1) setup(0);
  else {
    std::vector<std::thread> th;
    for (size_t i = 0; i < D; ++i) th.emplace_back(setup, i);
    for (auto& t : th) t.join();
  }
  for (size_t i = 0; i < D; ++i) if (!ok[i]) { reader_destroy(R); return nullptr; }
  return R;
}

static void reader_destroy(bamgpu_reader* R) {
  const size_t D = R->devs.size();
  auto teardown = [&](size_t i) {
    DevCtx& dc = R->devs[i];
    for (int j = 0; j < R->J; ++j) {
      Job& jb = R->jobs[i * R->J + j];
      if (jb.E) { bamgpu_engine_free(jb.E); jb.E = nullptr; }
    }
    cudaSetDevice(dc.dev);
    if (dc.h2d) { cudaStreamSynchronize(dc.h2d); cudaStreamDestroy(dc.h2d); }
    if (dc.d2h) { cudaStreamSynchronize(dc.d2h); cudaStreamDestroy(dc.d2h); }
    for (auto& s : dc.slots) { if (s.copied) cudaEventDestroy(s.copied); s.host.release(); s.dev.release(); }
  };
  if (D <= 1) { if (D) teardown(0); }
  else {
    std::vector<std::thread> th;
    for (size_t i = 0; i < D; ++i) th.emplace_back(teardown, i);
    for (auto& t : th) t.join();
  }
  delete R;
}

extern "C" bamgpu_reader* bamgpu_reader_new(bamgpu_read_fn read, void* ctx, size_t thread_num, const uint64_t* track_progress,
                                            const bamgpu_opts* opts) {
  if (!read) return nullptr;
  bamgpu_reader* R = reader_create(read, ctx, thread_num, track_progress, opts);
  if (R) R->feeder = std::thread(feeder_main, R);
  return R;
}

extern "C" bamgpu_reader* bamgpu_reader_from_memory(const uint8_t* buf, size_t len, size_t thread_num, const bamgpu_opts* opts) {
  if (!buf && len) return nullptr;
  bamgpu_reader* R = reader_create(mem_read, nullptr, thread_num, nullptr, opts);
  if (!R) return nullptr;
  R->ctx = R;
  R->mem = buf;
  R->mem_len = len;
  if (len) {   // page-locked by the caller (cudaHostAlloc / cudaHostRegister / bamgpu_pinned_alloc)?
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, buf) == cudaSuccess && at.type == cudaMemoryTypeHost) {
      cudaPointerAttributes at2{};
      if (cudaPointerGetAttributes(&at2, buf + len - 1) == cudaSuccess && at2.type == cudaMemoryTypeHost) R->mem_pinned = true;
    }
    cudaGetLastError();
  }
  R->feeder = std::thread(feeder_main, R);
  return R;
}

extern "C" void bamgpu_reader_free(bamgpu_reader* R) {
  if (!R) return;
  {
    std::lock_guard<std::mutex> lk(R->mu);
    R->stop = true;
  }
  R->cv.notify_all();
  if (R->feeder.joinable()) R->feeder.join();
  reader_destroy(R);
}

extern "C" const char* bamgpu_last_error(const bamgpu_reader* R) {
  if (!R) return "null reader (no CUDA device? this library has no CPU fallback)";
  if (!R->err.empty()) return R->err.c_str();
  for (auto& j : R->jobs) if (j.E && !j.E->err.empty()) return j.E->err.c_str();
  return "";
}

extern "C" bamgpu_engine* bamgpu_reader_current_engine(bamgpu_reader* R) { return (R && R->cur) ? R->cur->E : nullptr; }

static void reader_release_slot(bamgpu_reader* R, Job& j) {
  if (j.slot < 0) return;
  {
    std::lock_guard<std::mutex> lk(R->mu);
    R->devs[j.dev_i].free_slots.push_back(j.slot);
  }
  j.slot = -1;
  R->cv.notify_all();
}

// Queues K1 (+K2) for every batch the feeder has ready, as long as a job of its GPU is free.  Waits for the feeder
// while fewer than `need_launched` batches have been launched (0: never waits).  BAMGPU_OK or an error.
static int reader_pump(bamgpu_reader* R, uint32_t flags, uint64_t need_launched) {
  for (;;) {
    if (R->input_done) return BAMGPU_OK;
    Job& j = R->job_of(R->n_launched);
    if (j.st != Job::FREE) return BAMGPU_OK;   // every job of that GPU is busy
    ReadyItem it;
    {
      std::unique_lock<std::mutex> lk(R->mu);
      if (R->ready.empty()) {
        if (R->n_launched >= need_launched) return BAMGPU_OK;
        R->cv.wait(lk, [&] { return !R->ready.empty(); });
      }
      it = R->ready.front();
      // K1's CTAs are persistent (~10 ms) and three of them fill an SM.  If the K1 launches in flight on this GPU filled
      // every slot, the CRC and record kernels of the batch in FRONT would wait for a whole K1 to drain (measured: 37 ms
      // at the start of a pass).  So K1 launches are admitted only up to a CTA budget that leaves some SM slots free.
      {
        const DevCtx& dcx = R->devs[it.dev_i];
        const size_t nb_new = dcx.slots[it.slot].blocks.size();
        bamgpu_engine* E0 = j.E;
        const uint32_t budget = (uint32_t)E0->num_sms * 3u - 64u;
        const uint32_t g_new = E0->inflate_impl == 0 ? inflate_grid((uint32_t)nb_new, E0->num_sms) : 0;
        uint32_t in_flight = 0;
        for (int k = 0; k < R->J; ++k) {
          Job& o = R->jobs[(size_t)it.dev_i * R->J + k];
          if (o.st == Job::LAUNCHED && o.E->inflate_launched && o.E->k1_grid && cudaEventQuery(o.E->ev[1]) == cudaErrorNotReady) in_flight += o.E->k1_grid;
        }
        cudaGetLastError();
        if (in_flight && in_flight + g_new > budget) {
          if (R->n_launched >= need_launched) return BAMGPU_OK;   // try again at the next look
          lk.unlock();
          std::this_thread::sleep_for(std::chrono::microseconds(50));
          continue;
        }
      }
      R->ready.pop_front();
    }
    DevCtx& dc = R->devs[it.dev_i];
    Slot& S = dc.slots[it.slot];
    j.seq = it.seq;
    j.slot = it.slot;
    j.final = S.final;
    j.carry_set = false;
    j.input_status = S.status;
    j.stream_new = R->launched_stream;
    j.new_bytes = S.total_isize;
    R->launched_stream += S.total_isize;
    if (S.final) R->input_done = true;
    bamgpu_engine* E = j.E;
    j.st = Job::LAUNCHED;
    ++R->n_launched;
    if (dbg_level() >= 2) fprintf(stderr, "[bamgpu] @%.2f K1 of batch %llu queued (%zu blocks)\n", now_ms(), (unsigned long long)j.seq, S.blocks.size());
    if (S.status == BAMGPU_ERR_NOMEM) continue;   // nothing on the device; the error surfaces when the batch is staged
    cudaSetDevice(E->device);
    if (cudaStreamWaitEvent(E->stream, S.copied, 0) != cudaSuccess) { R->err = "cudaStreamWaitEvent failed"; return BAMGPU_ERR_CUDA; }
    int rc = engine_inflate_launch(E, S.dev.as<uint8_t>(), S.blocks.data(), S.blocks.size(), flags, E->stream);
    if (rc != BAMGPU_OK) return rc;
  }
}

// Waits for `ev` while keeping the pipeline fed: every batch the feeder gets ready meanwhile has its K1 queued at once,
// not when the consumer next happens to look.
static int reader_wait_event(bamgpu_reader* R, uint32_t flags, cudaEvent_t ev) {
  for (;;) {
    cudaError_t q = cudaEventQuery(ev);
    if (q == cudaSuccess) return BAMGPU_OK;
    if (q != cudaErrorNotReady) { R->err = cudaGetErrorString(q); return BAMGPU_ERR_CUDA; }
    int rc = reader_pump(R, flags, 0);
    if (rc < 0) return rc;
    std::this_thread::sleep_for(std::chrono::microseconds(40));
  }
}

// Waits for batch n_staged's inflate, splices the previous batch's tail in front of it and (unless `raw`) parses it; its
// D2H is queued, not waited for.  BAMGPU_OK: the batch is staged.  BAMGPU_EOF: no more input.
static int reader_stage(bamgpu_reader* R, uint32_t flags, bool raw) {
  for (;;) {
    if (R->no_more) return BAMGPU_EOF;
    if (R->n_staged == R->n_launched) {
      int rc = reader_pump(R, flags, R->n_staged + 1);
      if (rc < 0) return rc;
      if (R->n_staged == R->n_launched) {
        if (R->input_done) { R->no_more = true; return BAMGPU_EOF; }
        R->err = "internal: no free job for the next batch";
        return BAMGPU_ERR_STATE;
      }
    }
    Job& j = R->job_of(R->n_staged);
    bamgpu_engine* E = j.E;
    if (j.input_status == BAMGPU_ERR_NOMEM) { R->err = "out of memory for the input batch"; return BAMGPU_ERR_NOMEM; }
    int rc = E->inflate_launched ? reader_wait_event(R, flags, E->ev_inflated) : BAMGPU_OK;
    if (rc < 0) return rc;
    rc = engine_inflate_finish(E, E->stream);
    reader_release_slot(R, j);
    if (rc < 0) { R->err = E->err; return rc; }
    if (!j.carry_set) {
      bamgpu_engine* prev = (!raw && R->n_staged > 0) ? R->job_of(R->n_staged - 1).E : nullptr;
      rc = engine_set_carry(E, prev, E->stream);
      if (rc < 0) { R->err = E->err; return rc; }
      j.carry_set = true;
    }
    const uint64_t carry = E->lead - E->data_start;
    // a batch that ends the input because of an error is parsed as "more would follow": its complete records are handed
    // out, the error comes afterwards (where the reference's reading thread would have died)
    const bool fin = j.final && j.input_status >= 0;
    E->first_record_index = R->rec_index;
    uint64_t tail = 0;
    if (dbg_level() >= 2) fprintf(stderr, "[bamgpu] @%.2f batch %llu inflated\n", now_ms(), (unsigned long long)j.seq);
    rc = engine_records(E, E->data_start + R->next_start_rel, fin ? 1 : 0, raw ? (flags | BAMGPU_NO_RECORDS | BAMGPU_COPY_DATA) : flags,
                        E->stream, &j.batch, &tail, /*defer_d2h=*/true);
    R->next_start_rel = 0;
    j.batch.stream_offset = j.stream_new - carry;
    if (rc == BAMGPU_ERR_SHORT_RECORD && j.batch.n_records) { R->sticky_next = rc; R->err = E->err; rc = BAMGPU_OK; R->no_more = true; }
    if (rc < 0) { R->err = E->err; return rc; }
    R->rec_index += j.batch.n_records;
    ++R->n_staged;
    j.st = Job::STAGED;
    if (dbg_level() >= 2)
      fprintf(stderr, "[bamgpu] @%.2f staged batch %llu on gpu %d: %llu records, carry %llu B, launched %llu, handed out %llu\n", now_ms(),
              (unsigned long long)j.seq, E->device, (unsigned long long)j.batch.n_records, (unsigned long long)carry,
              (unsigned long long)R->n_launched, (unsigned long long)R->n_out);
    if (rc == BAMGPU_EOF || j.final) R->no_more = true;
    if (j.input_status < 0) { R->sticky_next = j.input_status; R->err = std::string("input: ") + bamgpu_status_str(j.input_status); }
    return BAMGPU_OK;
  }
}

extern "C" int bamgpu_read_header(bamgpu_reader* R, const uint8_t** bytes, size_t* len, size_t* offset_to_ref_seqs) {
  if (!R) return BAMGPU_ERR_ARG;
  if (R->sticky < 0) return R->sticky;
  if (R->header_done || R->raw_mode) { R->err = "read_header called twice or after read()"; return BAMGPU_ERR_STATE; }
  auto fail = [&](int code, const char* what) { R->err = what; R->sticky = code; return code; };
  // The header is read off the first batch on the device (normally a few KB of it).  If a batch ends inside the header,
  // everything inflated so far becomes the "tail" spliced in front of the next batch, exactly like a partial record.
  std::vector<uint8_t> hb;
  Job* hj = nullptr;
  auto next_job = [&]() -> int {
    const uint64_t seq = hj ? hj->seq + 1 : 0;
    int rc = reader_pump(R, R->flags, seq + 1);
    if (rc < 0) return rc;
    if (R->n_launched <= seq) return BAMGPU_ERR_SHORT_HEADER;
    Job& j = R->job_of(seq);
    bamgpu_engine* E = j.E;
    if (j.input_status == BAMGPU_ERR_NOMEM) return BAMGPU_ERR_NOMEM;
    if (E->inflate_launched) { rc = reader_wait_event(R, R->flags, E->ev_inflated); if (rc < 0) return rc; }
    rc = engine_inflate_finish(E, E->stream);
    reader_release_slot(R, j);
    if (rc < 0) { R->err = E->err; return rc; }
    rc = engine_set_carry(E, hj ? hj->E : nullptr, E->stream);
    if (rc < 0) { R->err = E->err; return rc; }
    j.carry_set = true;
    if (hj) { hj->st = Job::FREE; ++R->n_staged; ++R->n_out; }   // its bytes now live in front of this batch
    hj = &j;
    return BAMGPU_OK;
  };
  auto fetch = [&](uint64_t need) -> int {  // make hb hold the first `need` inflated bytes
    for (;;) {
      bamgpu_engine* E = hj->E;
      const uint64_t have = E->data_len - E->data_start;
      if (have >= need) break;
      if (hj->final) return hj->input_status < 0 ? hj->input_status : BAMGPU_ERR_SHORT_HEADER;
      // keep everything inflated so far as the carry of the next batch
      E->keep_from = E->data_start;
      E->carry_len = have;
      int rc = engine_save_carry(E, E->stream);
      if (rc < 0) return rc;
      rc = next_job();
      if (rc < 0) return rc;
    }
    if (hb.size() < need) {
      size_t old = hb.size();
      hb.resize(need);
      bamgpu_engine* E = hj->E;
      int rc = bamgpu_memcpy_d2h(E, hb.data() + old, E->out_ptr() + E->data_start + old, need - old);
      if (rc < 0) return rc;
    }
    return BAMGPU_OK;
  };
  int rc = next_job();
  if (rc == BAMGPU_ERR_SHORT_HEADER) return fail(rc, "short BAM header");
  if (rc < 0) { R->sticky = rc; return rc; }
  if ((rc = fetch(4)) < 0) return fail(rc, "short BAM header");
  if (memcmp(hb.data(), "BAM\1", 4) != 0) return fail(BAMGPU_ERR_BAD_MAGIC, "invalid BAM header");   // reader.rs:90-95
  if ((rc = fetch(8)) < 0) return fail(rc, "short BAM header");
  uint64_t p = 4;
  uint64_t l_text = rd32(hb.data() + p); p += 4;
  if ((rc = fetch(p + l_text + 4)) < 0) return fail(rc, "short BAM header");
  p += l_text;
  R->off_to_refs = (size_t)(p - 4);
  uint32_t n_ref = rd32(hb.data() + p); p += 4;
  for (uint32_t i = 0; i < n_ref; ++i) {
    if ((rc = fetch(p + 4)) < 0) return fail(rc, "short BAM header");
    uint64_t l_name = rd32(hb.data() + p); p += 4;
    if ((rc = fetch(p + l_name + 4)) < 0) return fail(rc, "short BAM header");
    p += l_name + 4;
  }
  R->header.assign(hb.begin() + 4, hb.begin() + p);
  R->header_done = true;
  R->next_start_rel = p;
  R->n_ref = n_ref > 0x7fffffffu ? -1 : (int32_t)n_ref;
  for (auto& j : R->jobs) j.E->n_ref = R->n_ref;
  if (bytes) *bytes = R->header.data();
  if (len) *len = R->header.size();
  if (offset_to_ref_seqs) *offset_to_ref_seqs = R->off_to_refs;
  return BAMGPU_OK;
}

// Produces the next non-empty batch (or EOF).  Per call: the batch the caller held goes back to its GPU; K1 is queued for
// whatever the feeder has ready; up to D batches beyond the one about to be handed out are parsed and their D2H queued
// (one per GPU link); then the call waits for ITS batch's D2H only.
static int reader_next_batch(bamgpu_reader* R, uint32_t extra_flags, bool raw = false) {
  if (R->sticky < 0) return R->sticky;
  if (R->eof) return BAMGPU_EOF;
  const uint32_t flags = R->flags | extra_flags;
  if (R->cur) {
    // the pinned mirrors go back to the process-wide pool: only the batches between "parsed" and "handed out" hold some,
    // however many jobs are inflating
    bamgpu_engine::Mirror& M = R->cur->E->mir[R->cur->E->msel];
    M.h_data.release();
    M.h_cols.release();
    R->cur->st = Job::FREE;
    R->cur = nullptr;
    R->batch_valid = false;
  }
  auto end_of_stream = [&](int rc) {
    if (rc == BAMGPU_EOF) {
      if (R->sticky_next < 0) { R->sticky = R->sticky_next; return R->sticky; }
      R->eof = true;
      return (int)BAMGPU_EOF;
    }
    R->sticky = rc;
    return rc;
  };
  for (;;) {
    if (R->n_out == R->n_staged) {
      int rc = reader_stage(R, flags, raw);
      if (rc != BAMGPU_OK) return end_of_stream(rc);
    }
    Job& j = R->job_of(R->n_out);
    if (j.batch.n_records || raw) break;
    { int w = engine_wait_mirror(j.E, j.E->msel); if (w < 0) return end_of_stream(w); }   // a batch without a record: drop it
    j.st = Job::FREE;
    ++R->n_out;
  }
  {
    int rc = reader_pump(R, flags, 0);
    if (rc < 0) return end_of_stream(rc);
  }
  const double t0 = now_ms();
  // look-ahead: parse the batches behind this one whose K1 is already queued, so that their D2H follows on
  while (!R->no_more && R->sticky_next == BAMGPU_OK && R->n_staged - R->n_out < 1 + R->devs.size() &&
         (R->n_staged < R->n_launched || R->n_staged == R->n_out + 1)) {
    if (R->n_staged == R->n_launched && R->job_of(R->n_launched).st != Job::FREE) break;   // would wait for ourselves
    int rn = reader_stage(R, flags, raw);
    if (rn < 0) { R->sticky_next = rn; R->no_more = true; break; }
    if (rn == BAMGPU_EOF) break;
    int rc = reader_pump(R, flags, 0);
    if (rc < 0) { R->sticky_next = rc; R->no_more = true; break; }
  }
  const double t1 = now_ms();
  Job& j = R->job_of(R->n_out);
  bamgpu_engine* E = j.E;
  // a caller that switched APIs in mid-stream (next_batch, then next_rec) may need mirrors this batch was staged without
  {
    cudaSetDevice(E->device);
    bool q = false;
    int w = engine_enqueue_mirrors(E, E->mir[E->msel], raw ? (flags | BAMGPU_COPY_DATA) : flags, j.batch.n_records, &j.batch, E->stream, &q);
    if (w < 0) { R->sticky = w; R->err = E->err; return w; }
  }
  if (E->mir[E->msel].pending) {
    int pw = reader_wait_event(R, flags, E->mir[E->msel].ev_end);
    if (pw < 0) { R->sticky = pw; return pw; }
  }
  int w = engine_wait_mirror(E, E->msel);
  if (w < 0) { R->sticky = w; R->err = E->err; return w; }
  if (dbg_level() >= 2) fprintf(stderr, "[bamgpu] @%.2f batch %llu handed out: staging ahead took %.2f, waiting for its D2H %.2f ms\n", now_ms(), (unsigned long long)j.seq, t1 - t0, now_ms() - t1);
  j.st = Job::OUT;
  ++R->n_out;
  R->cur = &j;
  R->batch = j.batch;
  R->batch_valid = true;
  R->rec_cursor = 0;
  return BAMGPU_OK;
}

extern "C" int bamgpu_next_batch(bamgpu_reader* R, bamgpu_batch* out) {
  if (!R || !out) return BAMGPU_ERR_ARG;
  if (R->raw_mode) { R->err = "next_batch after read()"; return BAMGPU_ERR_STATE; }
  if (!R->header_done) { R->err = "records requested before read_header()"; return BAMGPU_ERR_STATE; }
  if (R->eof) return R->sticky < 0 ? R->sticky : BAMGPU_EOF;
  R->records_mode = true;
  int rc = reader_next_batch(R, 0);
  if (rc != BAMGPU_OK) return rc;
  *out = R->batch;
  R->rec_cursor = R->batch.n_records;  // the batch is handed out whole
  return BAMGPU_OK;
}

extern "C" int bamgpu_next_rec(bamgpu_reader* R, const uint8_t** body, size_t* len) {
  if (!R || !body || !len) return BAMGPU_ERR_ARG;
  if (R->raw_mode) { R->err = "next_rec after read()"; return BAMGPU_ERR_STATE; }
  if (!R->header_done) { R->err = "records requested before read_header()"; return BAMGPU_ERR_STATE; }
  R->records_mode = true;
  while (!R->batch_valid || R->rec_cursor >= R->batch.n_records) {
    if (R->eof || R->sticky < 0) { *body = nullptr; *len = 0; return R->sticky < 0 ? R->sticky : BAMGPU_EOF; }
    int rc = reader_next_batch(R, BAMGPU_COPY_DATA | BAMGPU_COPY_OFFSETS);
    if (rc != BAMGPU_OK) { *body = nullptr; *len = 0; return rc; }
  }
  if (!R->batch.h_data || !R->batch.host.rec_off) { R->err = "internal: batch has no host mirror"; return BAMGPU_ERR_STATE; }
  uint64_t i = R->rec_cursor++;
  *body = R->batch.h_data + R->batch.host.rec_off[i];
  *len = R->batch.host.rec_len[i];
  // the bodies were written by the GPU's DMA, not by this core: a consumer that looks at every record (bam_binary's md5, main.rs:83-99)
  // misses the cache on each; ask for a record a few calls ahead (its first and last line) while the caller works on this one
  if (i + 6 < R->batch.n_records) {
    const uint8_t* nx = R->batch.h_data + R->batch.host.rec_off[i + 6];
    __builtin_prefetch(nx, 0, 1);
    __builtin_prefetch(nx + R->batch.host.rec_len[i + 6] - 1, 0, 1);
  }
  return BAMGPU_OK;
}

extern "C" int bamgpu_append_record(bamgpu_reader* R, uint8_t* dst, size_t cap, size_t* len) {
  if (!R || !len) return BAMGPU_ERR_ARG;
  const uint8_t* b;
  size_t n;
  int rc = bamgpu_next_rec(R, &b, &n);
  if (rc == BAMGPU_EOF) { *len = 0; return BAMGPU_OK; }   // reader.rs:63-73 returns Ok(0) at EOF
  if (rc < 0) return rc;
  if (n > cap || !dst) {
    R->rec_cursor = R->rec_cursor - 1;  // un-consume (same batch by construction)
    *len = n;
    return BAMGPU_ERR_BUFFER_TOO_SMALL;
  }
  memcpy(dst, b, n);
  *len = n;
  return BAMGPU_OK;
}

// impl Read for Reader (reader.rs:114-152).  Before read_header() it serves the inflated stream from byte 0; after it,
// the stream continues behind the header, as in the reference (read_header consumed those bytes through the same Read).
extern "C" long bamgpu_read(bamgpu_reader* R, uint8_t* dst, size_t cap) {
  if (!R || (!dst && cap)) return BAMGPU_ERR_ARG;
  if (R->sticky < 0) return R->sticky;
  if (R->records_mode) { R->err = "read() after records were requested: use one of the two on a Reader"; return BAMGPU_ERR_STATE; }
  R->raw_mode = true;
  size_t got = 0;
  while (got < cap) {
    if (R->batch_valid && R->byte_cursor < R->batch.data_len) {
      size_t n = (size_t)std::min<uint64_t>(cap - got, R->batch.data_len - R->byte_cursor);
      memcpy(dst + got, R->batch.h_data + R->byte_cursor, n);
      R->byte_cursor += n;
      got += n;
      continue;
    }
    if (R->eof) break;
    const uint64_t skip = R->next_start_rel;   // != 0 only for the batch read_header() stopped in
    int rc = reader_next_batch(R, BAMGPU_NO_RECORDS | BAMGPU_COPY_DATA, /*raw=*/true);
    if (rc == BAMGPU_EOF) break;
    if (rc < 0) return got ? (long)got : rc;
    R->byte_cursor = R->batch.data_start + skip;
  }
  return (long)got;
}

extern "C" int bamgpu_reader_stats(const bamgpu_reader* R, bamgpu_stats* out) {
  if (!R || !out) return BAMGPU_ERR_ARG;
  bamgpu_stats t{};
  for (auto& j : R->jobs) {
    if (!j.E) continue;
    const bamgpu_stats& e = j.E->stats;
    t.blocks += e.blocks; t.bytes_in += e.bytes_in; t.bytes_out += e.bytes_out; t.records += e.records; t.batches += e.batches;
    t.kernel_launches += e.kernel_launches; t.ms_h2d += e.ms_h2d; t.ms_inflate += e.ms_inflate; t.ms_crc += e.ms_crc;
    t.ms_records += e.ms_records; t.ms_d2h += e.ms_d2h; t.ms_sort += e.ms_sort; t.ms_gather += e.ms_gather; t.p2p_bytes += e.p2p_bytes;
    t.ms_bgzf_write += e.ms_bgzf_write;
  }
  t.ms_total = t.ms_h2d + t.ms_inflate + t.ms_crc + t.ms_records + t.ms_d2h + t.ms_sort + t.ms_gather;
  *out = t;
  return BAMGPU_OK;
}

extern "C" int bamgpu_parse_reference_sequences(const uint8_t* b, size_t len, bamgpu_refseq_cb cb, void* user) {
  if (!b || len < 4) return BAMGPU_ERR_SHORT_HEADER;
  uint32_t n = rd32(b);
  size_t p = 4;
  for (uint32_t i = 0; i < n; ++i) {
    if (len - p < 4) return BAMGPU_ERR_SHORT_HEADER;
    size_t l = rd32(b + p); p += 4;
    if (len - p < l) return BAMGPU_ERR_SHORT_HEADER;
    // CStr::from_bytes_with_nul (reader.rs:202-211): exactly one NUL, at the end
    if (l == 0 || b[p + l - 1] != 0 || memchr(b + p, 0, l - 1) != nullptr) return BAMGPU_ERR_SHORT_HEADER;
    const char* name = (const char*)(b + p);
    p += l;
    if (len - p < 4) return BAMGPU_ERR_SHORT_HEADER;
    uint32_t l_ref = rd32(b + p); p += 4;
    if (cb) cb(name, l - 1, l_ref, user);
  }
  return BAMGPU_OK;
}

// =================================================================================================
// sort_bam (sorting/sort.rs:138-178): reader -> records -> stable sort -> bodies to the sink
// =================================================================================================
// -------------------------------------------------------------------------------------------------
// Streamed sort: inputs whose inflated records do not fit in HBM (sort.rs:185-308 read_split_sort_dump_chunks +
// :590-652 merge_sorted_chunks_and_write, i.e. what `mem_limit` is for).  The file goes through the Reader batch by batch;
// of every batch only what the comparator looks at stays on the device — coord_key + strand, or l_read_name / flag / HI and
// the read names packed into a compact store — while the bodies are kept in host memory (the reference writes them to chunk
// files).  One global stable sort of the keys then gives the permutation (the result of the reference's per-chunk sorts +
// N-way merge, ties to the lower chunk = input order), and the bodies are gathered on the host in that order.
// -------------------------------------------------------------------------------------------------
namespace {
struct GrowBuf {   // device array (on the device that is current when it first grows) that keeps its contents when it grows
  DevBuf b;
  size_t used = 0;
  cudaError_t ensure(size_t need, cudaStream_t s) {   // s: a stream of the array's device, with nothing of ours in flight elsewhere
    if (need <= b.cap) return cudaSuccess;
    DevBuf nb;
    cudaError_t e = nb.reserve(std::max(need, b.cap + b.cap / 2));
    if (e != cudaSuccess) return e;
    if (used) {
      e = cudaMemcpyAsync(nb.p, b.p, used, cudaMemcpyDeviceToDevice, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) { nb.release(); return e; }
    }
    b.release();
    b = nb;
    return cudaSuccess;
  }
  // Appends `bytes` from src_dev's memory; home / hs: the array's device and a stream on it, ss: a stream of src_dev.
  cudaError_t append(const void* d_src, size_t bytes, int home, cudaStream_t hs, int src_dev, cudaStream_t ss) {
    cudaSetDevice(home);
    cudaError_t e = ensure(used + bytes + 64, hs);
    cudaSetDevice(src_dev);
    if (e != cudaSuccess) return e;
    if (bytes) e = src_dev == home ? cudaMemcpyAsync((uint8_t*)b.p + used, d_src, bytes, cudaMemcpyDeviceToDevice, ss)
                                   : cudaMemcpyPeerAsync((uint8_t*)b.p + used, home, d_src, src_dev, bytes, ss);
    used += bytes;
    return e;
  }
};
struct HostRun {   // one batch's inflated bytes and where its records are
  uint8_t* mem = nullptr;
  uint64_t first = 0, n = 0;
  std::vector<uint64_t> off;
  std::vector<uint32_t> len;
};
struct MemSrc { const uint8_t* p; size_t n, pos; };
long memsrc_read(void* ctx, uint8_t* dst, size_t cap) {
  MemSrc* m = (MemSrc*)ctx;
  const size_t k = std::min(cap, m->n - m->pos);
  memcpy(dst, m->p + m->pos, k);
  m->pos += k;
  return (long)k;
}
void par_memcpy(uint8_t* dst, const uint8_t* src, size_t n, size_t threads) {
  if (n < ((size_t)8 << 20) || threads < 2) { memcpy(dst, src, n); return; }
  threads = std::min<size_t>(threads, 16);
  std::vector<std::thread> th;
  for (size_t t = 0; t < threads; ++t) {
    const size_t a = n * t / threads, b = n * (t + 1) / threads;
    th.emplace_back([=] { memcpy(dst + a, src + a, b - a); });
  }
  for (auto& x : th) x.join();
}
}  // namespace

static int sort_bam_streamed(bamgpu_read_fn read, void* read_ctx, bamgpu_write_fn write, void* write_ctx, size_t threads, int sort_by,
                             const bamgpu_opts* opts, bamgpu_stats* stats, std::string& msg) {
  bamgpu_opts o{};
  if (opts) o = *opts; else o.device = -1;
  const int out_fmt = o.sort_output;
  if (out_fmt < BAMGPU_OUT_BODIES || out_fmt > BAMGPU_OUT_BAM) { msg = "unknown sort_output"; return BAMGPU_ERR_ARG; }
  o.jobs_per_device = 0;     // (n_devices stays: the batches are decoded on every listed GPU, the keys collect on the first one used)
  if (o.sort_mem_limit) o.batch_bytes = std::min<uint64_t>(std::max<uint64_t>(o.sort_mem_limit / 8, 64 << 10), (uint64_t)256 << 20);
  o.flags &= ~(uint32_t)(BAMGPU_COPY_COLUMNS | BAMGPU_NO_RECORDS | BAMGPU_SPECULATIVE_START);
  o.flags |= BAMGPU_COPY_DATA | BAMGPU_COPY_OFFSETS;
  if (sort_by == BAMGPU_SORT_NAME_AND_MATCH_MATES) o.flags |= BAMGPU_WANT_HI;
  threads = std::max<size_t>(threads, 1);
  bamgpu_reader* R = bamgpu_reader_new(read, read_ctx, threads, nullptr, &o);
  if (!R) { msg = "no usable CUDA device (this library has no CPU fallback)"; return BAMGPU_ERR_CUDA; }
  int rc = BAMGPU_OK;
  std::vector<HostRun> runs;
  GrowBuf g_key, g_strand, g_lname, g_flag, g_hip, g_hiv, g_voff, g_names;
  std::map<int, std::pair<DevBuf, DevBuf>> tmp;   // per source device: packed names / their offsets before they move to the home device
  bamgpu_engine* E = nullptr;      // the engine that holds the current batch
  bamgpu_engine* H = nullptr;      // "home": the first engine used; its device keeps the keys and sorts them
  uint64_t p2p_bytes = 0;
  std::vector<uint8_t> header;
  uint64_t total = 0, bad_flags = 0, out_bytes = 0;
  const bool by_name = sort_by != BAMGPU_SORT_COORDINATES_AND_STRAND;
  do {
    const uint8_t* hb; size_t hl, ho;
    rc = bamgpu_read_header(R, &hb, &hl, &ho);   // read and discarded by the reference (sort.rs:152-153); BAMGPU_OUT_BAM re-emits it
    if (rc != BAMGPU_OK) { msg = bamgpu_last_error(R); break; }
    header.assign(hb, hb + hl);
    // 1. batches: keys stay on the device, bodies go to host memory
    for (;;) {
      bamgpu_batch b;
      rc = bamgpu_next_batch(R, &b);
      if (rc == BAMGPU_EOF) { rc = BAMGPU_OK; break; }
      if (rc != BAMGPU_OK) { msg = bamgpu_last_error(R); break; }
      E = bamgpu_reader_current_engine(R);
      if (!H) H = E;
      const uint64_t nb = b.n_records;
      if (nb == 0) continue;
      if (total + nb >= (1ull << 31)) { rc = BAMGPU_ERR_TOO_LARGE; msg = "sort: more than 2^31 records"; break; }
      const int home = H->device, dev = E->device;
      cudaStream_t hs = H->stream, s = E->stream;
      cudaSetDevice(dev);
      bad_flags += b.bad_flag_records;
      cudaError_t e = cudaSuccess;
      auto app = [&](GrowBuf& g, const void* src, size_t bytes) {
        if (dev != home) p2p_bytes += bytes;
        return g.append(src, bytes, home, hs, dev, s);
      };
      if (!by_name) {
        e = app(g_key, b.dev.coord_key, 8 * nb);
        if (e == cudaSuccess) e = app(g_strand, b.dev.strand, nb);
      } else {
        e = app(g_lname, b.dev.l_read_name, nb);
        if (e == cudaSuccess && sort_by == BAMGPU_SORT_NAME_AND_MATCH_MATES) {
          e = app(g_flag, b.dev.flag, 2 * nb);
          if (e == cudaSuccess) e = app(g_hip, b.dev.hi_present, nb);
          if (e == cudaSuccess) e = app(g_hiv, b.dev.hi_value, 4 * nb);
        }
        // the names: packed on the GPU that decoded the batch, then appended to the store on the home GPU
        uint64_t name_bytes = 0;
        ColumnsDev bc{};
        bc.rec_off = (uint64_t*)b.dev.rec_off;
        bc.l_read_name = (uint8_t*)b.dev.l_read_name;
        auto& t = tmp[dev];
        if (e == cudaSuccess) e = E->d_sort.reserve(sort_scratch_bytes(nb));
        if (e == cudaSuccess) e = names_total(bc, nb, E->d_sort.p, &name_bytes, s);
        if (e == cudaSuccess) e = t.first.reserve(name_bytes + 64);
        if (e == cudaSuccess) e = t.second.reserve(8 * nb + 64);
        if (e == cudaSuccess) e = pack_names(b.data, bc, nb, E->d_sort.p, t.first.as<uint8_t>(), g_names.used, t.second.as<uint64_t>(), s);
        if (e == cudaSuccess) e = app(g_names, t.first.p, name_bytes);
        if (e == cudaSuccess) e = app(g_voff, t.second.p, 8 * nb);
        E->stats.kernel_launches += 3;
      }
      if (e != cudaSuccess) { rc = e == cudaErrorMemoryAllocation ? BAMGPU_ERR_TOO_LARGE : BAMGPU_ERR_CUDA; msg = std::string("streamed sort: ") + cudaGetErrorString(e); break; }
      // the bodies: the batch's valid bytes as they are, plus the records' offsets in them
      HostRun r;
      r.first = total; r.n = nb;
      const uint64_t span = b.data_len - b.data_start;
      r.mem = (uint8_t*)malloc(span + 8);
      if (!r.mem) { rc = BAMGPU_ERR_NOMEM; msg = "streamed sort: out of host memory for the record bodies"; break; }
      par_memcpy(r.mem, b.h_data + b.data_start, span, threads);
      r.off.resize(nb); r.len.assign(b.host.rec_len, b.host.rec_len + nb);
      for (uint64_t i = 0; i < nb; ++i) r.off[i] = b.host.rec_off[i] - b.data_start;
      runs.push_back(std::move(r));
      total += nb;
      if (cudaStreamSynchronize(s) != cudaSuccess) { rc = BAMGPU_ERR_CUDA; msg = "streamed sort: key copy failed"; break; }   // before the batch's buffers are reused
    }
    if (rc != BAMGPU_OK) break;
    const uint32_t prefix = out_fmt == BAMGPU_OUT_BODIES ? 0u : 4u;
    // 2. one stable sort of all keys
    std::vector<uint32_t> perm(total);
    if (total) {
      if (sort_by == BAMGPU_SORT_COORDINATES_AND_STRAND && bad_flags) {
        rc = BAMGPU_ERR_SORT_BAD_FLAGS;
        msg = "sort: " + std::to_string(bad_flags) + " record(s) with flag bits >= 0x1000 (Flags::from_bits().unwrap() panics, comparators.rs:178)";
        break;
      }
      E = H;                       // the keys are on the home device
      cudaSetDevice(E->device);
      cudaStream_t s = E->stream;
      ColumnsDev gc{};
      gc.coord_key = (uint64_t*)g_key.b.p; gc.strand = (uint8_t*)g_strand.b.p; gc.l_read_name = (uint8_t*)g_lname.b.p;
      gc.flag = (uint16_t*)g_flag.b.p; gc.hi_present = (uint8_t*)g_hip.b.p; gc.hi_value = (int32_t*)g_hiv.b.p; gc.rec_off = (uint64_t*)g_voff.b.p;
      SortResult* d_sr = (SortResult*)(E->d_misc.as<uint8_t>() + 192);
      SortResult h_sr{};
      int launches = 0;
      cudaError_t e = E->d_sort.reserve(sort_scratch_bytes(total));
      if (e == cudaSuccess) e = E->d_perm.reserve(4 * (total + 1));
      cudaEvent_t t0, t1;
      cudaEventCreate(&t0); cudaEventCreate(&t1);
      cudaEventRecord(t0, s);
      if (e == cudaSuccess) e = sort_permutation((const uint8_t*)g_names.b.p, gc, total, sort_by, E->d_sort.p, E->d_perm.as<uint32_t>(), d_sr, &launches, s);
      cudaEventRecord(t1, s);
      if (e == cudaSuccess) e = cudaMemcpyAsync(&h_sr, d_sr, sizeof h_sr, cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaMemcpyAsync(perm.data(), E->d_perm.p, 4 * total, cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      float ms = 0;
      cudaEventElapsedTime(&ms, t0, t1);
      cudaEventDestroy(t0); cudaEventDestroy(t1);
      E->stats.ms_sort += ms;
      E->stats.kernel_launches += launches;
      if (e != cudaSuccess) { rc = e == cudaErrorMemoryAllocation ? BAMGPU_ERR_TOO_LARGE : BAMGPU_ERR_CUDA; msg = std::string("streamed sort: ") + cudaGetErrorString(e); break; }
      if (sort_by == BAMGPU_SORT_NAME_AND_MATCH_MATES) {
        if (h_sr.hi_bad_tags) { rc = BAMGPU_ERR_SORT_BAD_TAGS; msg = "sort: aux tags the reference's HI walk panics on"; break; }
        if (h_sr.hi_mixed) {
          rc = BAMGPU_ERR_SORT_HI_MIXED;
          msg = "sort: " + std::to_string(h_sr.hi_mixed) + " pair(s) of equal read names with HI on one record only (Option::unwrap() panics, comparators.rs:90)";
          break;
        }
      }
    }
    // 3. gather on the host in sorted order, in pieces of a whole number of BGZF payloads
    const size_t PIECE = (size_t)1024 * 0xff00;
    std::vector<uint8_t> stage;
    stage.reserve(PIECE + (1 << 20));
    if (out_fmt == BAMGPU_OUT_BAM) { stage.insert(stage.end(), {'B', 'A', 'M', 1}); stage.insert(stage.end(), header.begin(), header.end()); }
    auto flush = [&](size_t nbytes, bool last) -> int {   // the first nbytes of `stage` go out
      if (out_fmt != BAMGPU_OUT_BAM) { out_bytes += nbytes; return nbytes ? write(write_ctx, stage.data(), nbytes) : 0; }
      bamgpu_engine* W = E;
      std::unique_ptr<bamgpu_engine, void (*)(bamgpu_engine*)> own(nullptr, bamgpu_engine_free);
      if (!W) { own.reset(bamgpu_engine_new(&o)); W = own.get(); if (!W) return BAMGPU_ERR_CUDA; }   // a BAM without records: header + EOF marker
      const size_t bound = bamgpu_bgzf_bound(nbytes);
      if (W->d_sorted.reserve(nbytes + 64) != cudaSuccess || W->d_bgzf.reserve(bound + 64) != cudaSuccess) return BAMGPU_ERR_NOMEM;
      std::vector<uint8_t> framed(bound);
      size_t wrote = 0;
      int r = nbytes ? bamgpu_memcpy_h2d(W, W->d_sorted.p, stage.data(), nbytes) : BAMGPU_OK;
      if (r == BAMGPU_OK) r = bamgpu_engine_bgzf_write(W, W->d_sorted.as<uint8_t>(), nbytes, W->d_bgzf.as<uint8_t>(), bound, &wrote, last ? 1 : 0, o.out_compr_level, nullptr);
      if (r == BAMGPU_OK && wrote) r = bamgpu_memcpy_d2h(W, framed.data(), W->d_bgzf.p, wrote);
      if (r != BAMGPU_OK) return r;
      out_bytes += wrote;
      return wrote ? write(write_ctx, framed.data(), wrote) : 0;
    };
    // The records of a piece are looked up (which batch, where, how long), placed by a prefix sum over their lengths and copied
    // by `threads` workers: perm is a random walk through the batches, so both the look-ups and the copies are cache misses that
    // only several threads in flight can hide (one thread: about 2 GB/s).
    const size_t nt = std::min<size_t>(std::max<size_t>(threads, 1), 16);
    std::vector<const uint8_t*> src;
    std::vector<uint32_t> slen;
    std::vector<uint64_t> dpos;
    uint64_t group = 4096;                        // records looked up per round; follows what a piece turns out to hold
    auto parallel_over = [&](uint64_t n, const std::function<void(uint64_t, uint64_t)>& f) {
      const size_t k = n < 2048 ? 1 : nt;
      if (k == 1) { f(0, n); return; }
      std::vector<std::thread> th;
      for (size_t t = 0; t < k; ++t) th.emplace_back([&f, n, t, k] { f(n * t / k, n * (t + 1) / k); });
      for (auto& x : th) x.join();
    };
    for (uint64_t i = 0; i < total && rc == BAMGPU_OK;) {
      const uint64_t cnt = std::min<uint64_t>(group, total - i);
      src.resize(cnt); slen.resize(cnt); dpos.resize(cnt + 1);
      parallel_over(cnt, [&](uint64_t a, uint64_t b) {
        size_t hint = 0;
        for (uint64_t j = a; j < b; ++j) {
          const uint64_t gidx = perm[i + j];
          if (!(hint < runs.size() && gidx >= runs[hint].first && gidx < runs[hint].first + runs[hint].n)) {
            size_t lo = 0, hi = runs.size();
            while (hi - lo > 1) { const size_t m = (lo + hi) / 2; if (runs[m].first <= gidx) lo = m; else hi = m; }
            hint = lo;
          }
          const HostRun& r = runs[hint];
          const uint64_t k = gidx - r.first;
          src[j] = r.mem + r.off[k];
          slen[j] = r.len[k];
        }
      });
      // how many of them fill the piece
      uint64_t take = 0, pos = stage.size();
      while (take < cnt && pos < PIECE) { dpos[take] = pos; pos += prefix + slen[take]; ++take; }
      dpos[take] = pos;
      stage.resize(pos);
      uint8_t* out = stage.data();
      parallel_over(take, [&](uint64_t a, uint64_t b) {
        for (uint64_t j = a; j < b; ++j) {
          uint8_t* d = out + dpos[j];
          if (prefix) memcpy(d, &slen[j], 4);     // little-endian host (x86-64 / aarch64)
          memcpy(d + prefix, src[j], slen[j]);
        }
      });
      i += take;
      group = std::min<uint64_t>(std::max<uint64_t>(2 * take, 4096), (uint64_t)1 << 20);
      if (stage.size() >= PIECE) {
        const size_t whole = stage.size() / 0xff00 * 0xff00;
        const int w = flush(whole, false);
        if (w < 0) { rc = w == BAMGPU_ERR_CUDA || w == BAMGPU_ERR_NOMEM ? w : BAMGPU_ERR_IO; msg = "streamed sort: writing the output failed"; break; }
        stage.erase(stage.begin(), stage.begin() + whole);
      }
    }
    if (rc != BAMGPU_OK) break;
    if (!stage.empty() || out_fmt == BAMGPU_OUT_BAM) {
      const int w = flush(stage.size(), true);
      if (w < 0) { rc = w == BAMGPU_ERR_CUDA || w == BAMGPU_ERR_NOMEM ? w : BAMGPU_ERR_IO; msg = "streamed sort: writing the output failed"; }
    }
  } while (0);
  if (H) H->stats.p2p_bytes += p2p_bytes;
  if (stats) bamgpu_reader_stats(R, stats);
  for (auto& kv : tmp) { cudaSetDevice(kv.first); cudaDeviceSynchronize(); kv.second.first.release(); kv.second.second.release(); }
  if (H) { cudaSetDevice(H->device); cudaStreamSynchronize(H->stream); }
  for (GrowBuf* g : {&g_key, &g_strand, &g_lname, &g_flag, &g_hip, &g_hiv, &g_voff, &g_names}) g->b.release();
  bamgpu_reader_free(R);
  for (auto& r : runs) free(r.mem);
  return rc;
}

// -------------------------------------------------------------------------------------------------
// Sort over several GPUs ("multi-GPU => sample sort"): bamgpu_sort_bam with bamgpu_opts.devices[].
//   1. the file is cut into G contiguous BGZF block ranges of about equal compressed size; every GPU inflates its range
//      (K1 + K2) at the same time; the head of the ranges to the right is copied behind each range (peer copy, <= 1 MiB) so
//      that the record across the cut is whole; the record chain (reader.rs:63-81) then runs range by range — K3 is a few
//      milliseconds, and the start of range g + 1 is simply where range g's last record ended (no guessing);
//   2. a few hundred evenly spaced keys per GPU go to the host, which picks G - 1 splitters; records that compare equal to a
//      splitter all fall into the bucket above it, so a bucket holds whole groups of equal keys;
//   3. every GPU partitions its records by bucket (stable), gathers them as a (u32 block_size, body)* stream and writes
//      bucket d's piece into GPU d's memory (cudaMemcpyPeerAsync, NVLink): pieces sit there in range order;
//   4. GPU d parses what it received (K3 again: keys come from the records themselves) and sorts it with the single-GPU
//      sort — stable, so ties keep range order = input order; the output is bucket 0, bucket 1, ... back to back, byte for
//      byte what the single-GPU sort writes (for the BAM format the 0xff00-byte pieces that cross a bucket edge are
//      completed by copying the odd tail of bucket d in front of bucket d + 1).
// No collective; the exchange is G x G independent peer copies.  Anything unusual (a record longer than the halo, fewer blocks
// than GPUs, a header that does not fit the first range, a scan error) leaves the job to the single-GPU path.
// -------------------------------------------------------------------------------------------------
namespace {
constexpr int SORT_MULTI_NOT_APPLICABLE = 1000;

struct SortShard {
  int dev = 0;
  bamgpu_engine* A = nullptr;   // decodes the block range
  bamgpu_engine* B = nullptr;   // holds and sorts bucket g
  DevBuf d_comp, d_samples, d_split, d_bounds, d_send, d_soff;
  size_t b0 = 0, b1 = 0;
  std::vector<bamgpu_block> blocks;
  uint64_t own_len = 0, n = 0, bad_flags = 0;
  std::vector<SplitKey> samples;
  std::vector<uint32_t> bounds;   // [G + 1] where each bucket starts in the partitioned order
  std::vector<uint64_t> seg;      // [G + 1] ... and in d_send
  uint64_t recv_bytes = 0, recv_records = 0, out_len = 0, front = 0;
  uint8_t* stream = nullptr;      // front bytes + sorted output (device)
  uint64_t framed = 0;            // BAM: bytes of `stream` this GPU frames
  size_t bgzf_len = 0;
  int rc = BAMGPU_OK;
  std::string msg;
  void fail(int code, const std::string& m) { if (rc == BAMGPU_OK) { rc = code; msg = m; } }
};

template <class F> static void for_each_shard(std::vector<SortShard>& sh, F f) {
  std::vector<std::thread> th;
  for (size_t g = 0; g < sh.size(); ++g)
    th.emplace_back([&sh, &f, g] { if (sh[g].rc == BAMGPU_OK) { cudaSetDevice(sh[g].dev); f(sh[g], g); } });
  for (auto& t : th) t.join();
}
static bool shard_failed(const std::vector<SortShard>& sh, int* rc, std::string* msg) {
  for (auto& s : sh) if (s.rc != BAMGPU_OK) { *rc = s.rc; *msg = s.msg; return true; }
  return false;
}

// The inflated bytes of `len` bytes follow: makes the engine's buffer ready to receive them (no inflate).
static int engine_adopt_stream(bamgpu_engine* E, uint64_t len, cudaStream_t s) {
  ECK(cudaSetDevice(E->device));
  { int w = engine_wait_d2h(E); if (w < 0) return w; }
  ECK(E->d_outs[E->cur].reserve(OUT_LEAD_PAD + len + OUT_TAIL_PAD));
  ECK(cudaMemsetAsync(E->d_outs[E->cur].p, 0, OUT_LEAD_PAD, s));
  ECK(cudaMemsetAsync(E->out_ptr() + len, 0, OUT_TAIL_PAD, s));
  E->lead = 0; E->data_start = 0; E->carry_len = 0; E->keep_from = 0; E->data_len = len; E->halo_len = 0;
  E->inflate_launched = false; E->pend_nb = 0;
  return BAMGPU_OK;
}

// reader.rs:87-98,154-200 over inflated bytes that sit in device memory: header bytes without the magic, where the records begin.
static int header_from_device(bamgpu_engine* E, uint64_t avail, std::vector<uint8_t>& header, uint64_t* end, int32_t* n_ref_out) {
  std::vector<uint8_t> hb;
  uint64_t want = std::min<uint64_t>(avail, 1 << 16);
  for (;;) {
    if (hb.size() < want) {
      const size_t old = hb.size();
      hb.resize(want);
      int rc = bamgpu_memcpy_d2h(E, hb.data() + old, E->out_ptr() + old, want - old);
      if (rc < 0) return rc;
    }
    uint64_t need = 0, p = 4;
    do {
      if (hb.size() < 8) { need = 8; break; }
      if (memcmp(hb.data(), "BAM\1", 4) != 0) return BAMGPU_ERR_BAD_MAGIC;
      const uint64_t l_text = rd32(hb.data() + p); p += 4;
      if (hb.size() < p + l_text + 4) { need = p + l_text + 4; break; }
      p += l_text;
      const uint32_t n_ref = rd32(hb.data() + p); p += 4;
      for (uint32_t i = 0; i < n_ref && !need; ++i) {
        if (hb.size() < p + 4) { need = p + 4 + 4096; break; }
        const uint64_t l_name = rd32(hb.data() + p); p += 4;
        if (hb.size() < p + l_name + 4) { need = p + l_name + 4 + 4096; break; }
        p += l_name + 4;
      }
      if (!need) {
        header.assign(hb.begin() + 4, hb.begin() + p);
        *end = p;
        *n_ref_out = n_ref > 0x7fffffffu ? -1 : (int32_t)n_ref;
        return BAMGPU_OK;
      }
    } while (0);
    if (hb.size() >= avail) return BAMGPU_ERR_SHORT_HEADER;
    want = std::min<uint64_t>(avail, std::max<uint64_t>(need, hb.size() * 2));
  }
}

static bool split_less(const SplitKey& a, const SplitKey& b, bool by_name) {
  if (!by_name) return a.key != b.key ? a.key < b.key : a.strand < b.strand;
  const uint32_t l = std::min(a.name_len, b.name_len);
  const int c = memcmp(a.name, b.name, l);
  return c != 0 ? c < 0 : a.name_len < b.name_len;
}

struct OutSeg { int dev; cudaStream_t s; const uint8_t* p; uint64_t len; };
// The segments, one after the other, to the sink through two pinned staging buffers (the copy of piece k + 1 overlaps write(k)).
static int write_segments(const std::vector<OutSeg>& segs, bamgpu_write_fn write, void* write_ctx, uint64_t* out_bytes, double* ms, std::string& msg) {
  const size_t CH = (size_t)64 << 20;
  struct Piece { size_t seg; uint64_t off; size_t len; };
  std::vector<Piece> pieces;
  for (size_t i = 0; i < segs.size(); ++i)
    for (uint64_t o = 0; o < segs[i].len; o += CH) pieces.push_back({i, o, (size_t)std::min<uint64_t>(CH, segs[i].len - o)});
  if (pieces.empty()) return BAMGPU_OK;
  PinBuf st[2];
  if (st[0].reserve(CH) != cudaSuccess || st[1].reserve(CH) != cudaSuccess) { msg = "pinned staging"; return BAMGPU_ERR_NOMEM; }
  std::map<int, std::array<cudaEvent_t, 2>> evs;
  for (auto& sg : segs)
    if (!evs.count(sg.dev)) {
      cudaSetDevice(sg.dev);
      std::array<cudaEvent_t, 2> e{};
      cudaEventCreateWithFlags(&e[0], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&e[1], cudaEventDisableTiming);
      evs[sg.dev] = e;
    }
  const double t0 = now_ms();
  int rc = BAMGPU_OK;
  auto issue = [&](size_t k) {
    const Piece& pc = pieces[k];
    const OutSeg& sg = segs[pc.seg];
    cudaSetDevice(sg.dev);
    cudaMemcpyAsync(st[k & 1].p, sg.p + pc.off, pc.len, cudaMemcpyDeviceToHost, sg.s);
    cudaEventRecord(evs[sg.dev][k & 1], sg.s);
  };
  issue(0);
  for (size_t k = 0; k < pieces.size(); ++k) {
    if (k + 1 < pieces.size()) issue(k + 1);
    const OutSeg& sg = segs[pieces[k].seg];
    if (cudaEventSynchronize(evs[sg.dev][k & 1]) != cudaSuccess) { rc = BAMGPU_ERR_CUDA; msg = "D2H of the sorted output failed"; break; }
    if (write(write_ctx, st[k & 1].as<uint8_t>(), pieces[k].len) < 0) { rc = BAMGPU_ERR_IO; msg = "write callback failed"; break; }
    *out_bytes += pieces[k].len;
  }
  for (auto& sg : segs) { cudaSetDevice(sg.dev); cudaStreamSynchronize(sg.s); }
  *ms += now_ms() - t0;
  for (auto& kv : evs) { cudaSetDevice(kv.first); cudaEventDestroy(kv.second[0]); cudaEventDestroy(kv.second[1]); }
  st[0].release(); st[1].release();
  return rc;
}

static void add_stats(bamgpu_stats& acc, const bamgpu_stats& s) {
  acc.blocks += s.blocks; acc.bytes_in += s.bytes_in; acc.bytes_out += s.bytes_out; acc.records += s.records; acc.batches += s.batches;
  acc.kernel_launches += s.kernel_launches; acc.p2p_bytes += s.p2p_bytes;
  // the GPUs work side by side: the stage times are those of the slowest one
  acc.ms_h2d = std::max(acc.ms_h2d, s.ms_h2d); acc.ms_inflate = std::max(acc.ms_inflate, s.ms_inflate); acc.ms_crc = std::max(acc.ms_crc, s.ms_crc);
  acc.ms_records = std::max(acc.ms_records, s.ms_records); acc.ms_sort = std::max(acc.ms_sort, s.ms_sort);
  acc.ms_gather = std::max(acc.ms_gather, s.ms_gather); acc.ms_bgzf_write = std::max(acc.ms_bgzf_write, s.ms_bgzf_write);
}
}  // namespace

// buf[0..len): the whole BGZF file in host memory.  Returns BAMGPU_OK / an error, or SORT_MULTI_NOT_APPLICABLE (nothing written).
static int sort_bam_multi(const uint8_t* buf, size_t len, bamgpu_write_fn write, void* write_ctx, int sort_by, const bamgpu_opts& opts,
                          bamgpu_stats* stats, std::string& msg) {
  const int G = opts.n_devices;
  const int out_fmt = opts.sort_output;
  if (out_fmt < BAMGPU_OUT_BODIES || out_fmt > BAMGPU_OUT_BAM) { msg = "unknown sort_output"; return BAMGPU_ERR_ARG; }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { msg = "no usable CUDA device (this library has no CPU fallback)"; return BAMGPU_ERR_CUDA; }
  for (int g = 0; g < G; ++g) if (opts.devices[g] < 0 || opts.devices[g] >= count) { msg = "bamgpu_opts.devices: no such device"; return BAMGPU_ERR_ARG; }
  // 0. the block table, cut into G ranges
  size_t nblk = 0, consumed = 0;
  uint64_t isize_total = 0;
  int sc = bamgpu_scan_blocks(buf, len, 1, nullptr, 0, &nblk, &consumed, &isize_total);
  if (sc < 0 || nblk < (size_t)G * 2) return SORT_MULTI_NOT_APPLICABLE;
  std::vector<bamgpu_block> all(nblk);
  sc = bamgpu_scan_blocks(buf, len, 1, all.data(), nblk, &nblk, &consumed, &isize_total);
  if (sc < 0) return SORT_MULTI_NOT_APPLICABLE;
  {   // room on every GPU for its range, what it sends, its bucket and the bucket's sorted output (buckets are about even)?
    std::map<int, uint64_t> need;
    for (int g = 0; g < G; ++g) need[opts.devices[g]] += ((uint64_t)len + 4 * isize_total + isize_total / 2) / G + ((uint64_t)1 << 30);
    for (auto& kv : need) {
      size_t free_b = 0, total_b = 0;
      if (cudaSetDevice(kv.first) != cudaSuccess || cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { msg = "no usable CUDA device (this library has no CPU fallback)"; return BAMGPU_ERR_CUDA; }
      if (kv.second > (uint64_t)free_b + dev_cache_bytes()) return SORT_MULTI_NOT_APPLICABLE;
    }
  }
  for (int a = 0; a < G; ++a)
    for (int b = a + 1; b < G; ++b)
      if (opts.devices[a] != opts.devices[b]) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, opts.devices[a], opts.devices[b]);
        if (can) {
          cudaSetDevice(opts.devices[a]); cudaDeviceEnablePeerAccess(opts.devices[b], 0);
          cudaSetDevice(opts.devices[b]); cudaDeviceEnablePeerAccess(opts.devices[a], 0);
          cudaGetLastError();   // "already enabled" is fine; without peer access the copies are staged by the driver
        }
      }
  std::vector<SortShard> sh(G);
  {
    size_t b = 0;
    for (int g = 0; g < G; ++g) {
      sh[g].dev = opts.devices[g];
      sh[g].b0 = b;
      const uint64_t want = (uint64_t)consumed * (g + 1) / G;
      while (b < nblk && (g == G - 1 || all[b].in_off - 18 < want || b == sh[g].b0)) ++b;
      if (g == G - 1) b = nblk;
      sh[g].b1 = b;
      if (sh[g].b1 == sh[g].b0) return SORT_MULTI_NOT_APPLICABLE;
    }
  }
  const bool by_name = sort_by != BAMGPU_SORT_COORDINATES_AND_STRAND;
  const uint32_t rec_flags = (opts.flags & BAMGPU_NO_CRC) | (sort_by == BAMGPU_SORT_NAME_AND_MATCH_MATES ? BAMGPU_WANT_HI : 0u);
  int rc = BAMGPU_OK;
  double t_phase = now_ms();
  auto phase = [&](const char* what) {
    if (dbg_level() >= 1) { const double t = now_ms(); fprintf(stderr, "[bamgpu] sort over %d GPUs: %-28s %8.2f ms\n", G, what, t - t_phase); t_phase = t; }
  };
  uint64_t p2p_bytes = 0, out_bytes = 0;
  double ms_d2h = 0;
  bamgpu_stats acc{};
  std::vector<uint8_t> header;
  auto cleanup = [&]() {
    for (auto& s : sh) {
      cudaSetDevice(s.dev);
      if (s.A) { cudaStreamSynchronize(s.A->stream); bamgpu_stats st; bamgpu_engine_stats(s.A, &st); add_stats(acc, st); bamgpu_engine_free(s.A); s.A = nullptr; }
      if (s.B) { cudaStreamSynchronize(s.B->stream); bamgpu_stats st; bamgpu_engine_stats(s.B, &st); add_stats(acc, st); bamgpu_engine_free(s.B); s.B = nullptr; }
      for (DevBuf* d : {&s.d_comp, &s.d_samples, &s.d_split, &s.d_bounds, &s.d_send, &s.d_soff}) d->release();
    }
    if (stats) { acc.p2p_bytes += p2p_bytes; acc.ms_d2h += ms_d2h; acc.ms_total = acc.ms_h2d + acc.ms_inflate + acc.ms_crc + acc.ms_records + acc.ms_d2h + acc.ms_sort + acc.ms_gather; *stats = acc; }
  };
#define SHCK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { s.fail(_e == cudaErrorMemoryAllocation ? BAMGPU_ERR_NOMEM : BAMGPU_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); return; } } while (0)
  // 1. every GPU: its range to the device, K1 + K2
  for_each_shard(sh, [&](SortShard& s, size_t) {
    bamgpu_opts eo{};
    eo.device = s.dev;
    eo.flags = opts.flags & BAMGPU_NO_CRC;
    eo.inflate_impl = opts.inflate_impl;
    s.A = bamgpu_engine_new(&eo);
    if (!s.A) { s.fail(BAMGPU_ERR_CUDA, "no usable CUDA device (this library has no CPU fallback)"); return; }
    const uint64_t c0 = all[s.b0].in_off - 18, c1 = all[s.b1 - 1].in_off + all[s.b1 - 1].in_len + 8;
    s.blocks.assign(all.begin() + s.b0, all.begin() + s.b1);
    for (auto& b : s.blocks) b.in_off -= c0;
    SHCK(s.d_comp.reserve(c1 - c0 + IN_TAIL_PAD));
    SHCK(cudaMemsetAsync(s.d_comp.as<uint8_t>() + (c1 - c0), 0, IN_TAIL_PAD, s.A->stream));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, s.A->stream);
    SHCK(cudaMemcpyAsync(s.d_comp.p, buf + c0, c1 - c0, cudaMemcpyHostToDevice, s.A->stream));
    cudaEventRecord(e1, s.A->stream);
    int r = engine_inflate_launch(s.A, s.d_comp.as<uint8_t>(), s.blocks.data(), s.blocks.size(), eo.flags, s.A->stream);
    if (r == BAMGPU_OK) {
      // while K1 runs: the buffer this GPU will gather its (u32 block_size, body)* stream into (a cudaMalloc of several GB takes
      // tens of milliseconds, and the host has nothing else to do now); 4 bytes per record on top of the inflated bytes
      uint64_t isz = 0;
      for (auto& b : s.blocks) isz += b.isize;
      s.d_send.reserve(isz + isz / 16 + (1 << 20));   // (too small for very short records: reserved again later, then)
      r = engine_inflate_finish(s.A, s.A->stream);
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    s.A->stats.ms_h2d += ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (r < 0) { s.fail(r, s.A->err); return; }
    s.own_len = s.A->data_len;
  });
  do {
    if (shard_failed(sh, &rc, &msg)) break;
    phase("upload + K1 + K2");
    // 2. halos: the first bytes to the right of every range, from as many ranges as it takes (<= BAMGPU_MAX_HALO)
    for (int g = 0; g + 1 < G; ++g) {
      SortShard& s = sh[g];
      cudaSetDevice(s.dev);
      uint64_t have = 0;
      for (int h = g + 1; h < G && have < BAMGPU_MAX_HALO; ++h) {
        const uint64_t m = std::min<uint64_t>(sh[h].own_len, BAMGPU_MAX_HALO - have);
        if (m && cudaMemcpyPeerAsync(s.A->out_ptr() + s.own_len + have, s.dev, sh[h].A->out_ptr(), sh[h].dev, m, s.A->stream) != cudaSuccess) {
          rc = BAMGPU_ERR_CUDA; msg = "halo copy failed"; break;
        }
        if (sh[h].dev != s.dev) p2p_bytes += m;
        have += m;
      }
      if (rc != BAMGPU_OK) break;
      cudaMemsetAsync(s.A->out_ptr() + s.own_len + have, 0, OUT_TAIL_PAD, s.A->stream);
      s.A->halo_len = have;
      if (cudaStreamSynchronize(s.A->stream) != cudaSuccess) { rc = BAMGPU_ERR_CUDA; msg = "halo copy failed"; break; }
    }
    if (rc != BAMGPU_OK) break;
    // 3. the header, then the record chain range by range (reader.rs:63-81)
    uint64_t start = 0;
    int32_t n_ref = -1;
    {
      int r = header_from_device(sh[0].A, sh[0].own_len + sh[0].A->halo_len, header, &start, &n_ref);
      if (r == BAMGPU_ERR_BAD_MAGIC) { rc = r; msg = "invalid BAM header"; break; }
      if (r == BAMGPU_ERR_SHORT_HEADER) { rc = SORT_MULTI_NOT_APPLICABLE; break; }
      if (r < 0) { rc = r; msg = sh[0].A->err; break; }
    }
    uint64_t total = 0, bad_flags = 0;
    bool ended = false;
    for (int g = 0; g < G && rc == BAMGPU_OK; ++g) {
      SortShard& s = sh[g];
      if (ended) { s.n = 0; continue; }
      s.A->n_ref = n_ref;
      bamgpu_batch b;
      uint64_t tail = 0;
      const int final = g == G - 1;
      int r = engine_records(s.A, start, final, rec_flags, s.A->stream, &b, &tail);
      if (r < 0) { rc = r; msg = s.A->err; break; }
      s.n = b.n_records;
      bad_flags += b.bad_flag_records;
      total += s.n;
      if (r == BAMGPU_EOF) { ended = true; continue; }
      if (!final) {
        if (tail < s.own_len || tail > s.own_len + s.A->halo_len) { rc = SORT_MULTI_NOT_APPLICABLE; break; }   // a record longer than the halo
        start = tail - s.own_len;
      }
    }
    if (rc != BAMGPU_OK) break;
    phase("halos + header + K3 chain");
    if (total == 0) { rc = SORT_MULTI_NOT_APPLICABLE; break; }
    if (!by_name && bad_flags) {
      rc = BAMGPU_ERR_SORT_BAD_FLAGS;
      msg = "sort: " + std::to_string(bad_flags) + " record(s) with flag bits >= 0x1000 (Flags::from_bits().unwrap() panics, comparators.rs:178)";
      break;
    }
    // 4. splitters from evenly spaced samples
    for_each_shard(sh, [&](SortShard& s, size_t) {
      const uint32_t ns = (uint32_t)std::min<uint64_t>(s.n, 512);
      if (!ns) return;
      SHCK(s.d_samples.reserve(ns * sizeof(SplitKey)));
      SHCK(sample_keys(s.A->out_ptr(), s.A->mir[s.A->msel].cols, s.n, sort_by, ns, s.d_samples.as<SplitKey>(), s.A->stream));
      s.samples.resize(ns);
      SHCK(cudaMemcpyAsync(s.samples.data(), s.d_samples.p, ns * sizeof(SplitKey), cudaMemcpyDeviceToHost, s.A->stream));
      SHCK(cudaStreamSynchronize(s.A->stream));
      s.A->stats.kernel_launches += 1;
    });
    if (shard_failed(sh, &rc, &msg)) break;
    std::vector<SplitKey> split;
    {
      std::vector<SplitKey> pool;
      for (auto& s : sh) pool.insert(pool.end(), s.samples.begin(), s.samples.end());
      std::stable_sort(pool.begin(), pool.end(), [&](const SplitKey& a, const SplitKey& b) { return split_less(a, b, by_name); });
      for (int k = 1; k < G; ++k) split.push_back(pool[std::min(pool.size() - 1, pool.size() * k / G)]);
    }
    // 5. partition by bucket, gather (u32 block_size, body)* in that order
    for_each_shard(sh, [&](SortShard& s, size_t) {
      s.bounds.assign(G + 1, 0);
      s.seg.assign(G + 1, 0);
      if (!s.n) return;
      cudaStream_t st = s.A->stream;
      const ColumnsDev& cols = s.A->mir[s.A->msel].cols;
      SHCK(s.d_split.reserve(sizeof(SplitKey) * (G - 1) + 64));
      SHCK(cudaMemcpyAsync(s.d_split.p, split.data(), sizeof(SplitKey) * (G - 1), cudaMemcpyHostToDevice, st));
      SHCK(s.d_bounds.reserve(4 * (G + 2)));
      SHCK(s.A->d_sort.reserve(sort_scratch_bytes(s.n)));
      SHCK(s.A->d_perm.reserve(4 * (s.n + 1)));
      SHCK(s.d_soff.reserve(8 * (s.n + 2)));
      int launches = 0;
      cudaEventRecord(s.A->ev[3], st);
      SHCK(partition_by_splitters(s.A->out_ptr(), cols, s.n, sort_by, s.d_split.as<SplitKey>(), (uint32_t)(G - 1), s.A->d_sort.p,
                                  s.A->d_perm.as<uint32_t>(), s.d_bounds.as<uint32_t>(), &launches, st));
      cudaEventRecord(s.A->ev[6], st);
      SHCK(cudaMemcpyAsync(s.bounds.data(), s.d_bounds.p, 4 * (G + 1), cudaMemcpyDeviceToHost, st));
      uint64_t total_bytes = 0;
      SHCK(sorted_offsets(cols, s.A->d_perm.as<uint32_t>(), s.n, s.A->d_sort.p, s.d_soff.as<uint64_t>(), &total_bytes, 4, st));   // synchronises
      for (int d = 0; d <= G; ++d) SHCK(cudaMemcpyAsync(&s.seg[d], s.d_soff.as<uint64_t>() + s.bounds[d], 8, cudaMemcpyDeviceToHost, st));
      SHCK(s.d_send.reserve(total_bytes + 64));
      cudaEventRecord(s.A->ev[4], st);
      SHCK(gather_sorted_bodies(s.A->out_ptr(), cols, s.A->d_perm.as<uint32_t>(), s.n, s.A->d_sort.p, s.d_soff.as<uint64_t>(), s.d_send.as<uint8_t>(), 4, st));
      cudaEventRecord(s.A->ev[5], st);
      SHCK(cudaStreamSynchronize(st));
      float ms = 0;
      cudaEventElapsedTime(&ms, s.A->ev[3], s.A->ev[6]); s.A->stats.ms_sort += ms;
      cudaEventElapsedTime(&ms, s.A->ev[4], s.A->ev[5]); s.A->stats.ms_gather += ms;
      s.A->stats.kernel_launches += launches + 3;
    });
    if (shard_failed(sh, &rc, &msg)) break;
    phase("splitters + partition + gather");
    // 6. the exchange: bucket d's piece of every range into GPU d's memory, in range order
    std::vector<std::vector<uint64_t>> roff(G, std::vector<uint64_t>(G, 0));
    for (int d = 0; d < G; ++d) {
      uint64_t o = 0, cnt = 0;
      for (int g = 0; g < G; ++g) { roff[g][d] = o; o += sh[g].seg[d + 1] - sh[g].seg[d]; cnt += sh[g].bounds[d + 1] - sh[g].bounds[d]; }
      sh[d].recv_bytes = o;
      sh[d].recv_records = cnt;
    }
    for_each_shard(sh, [&](SortShard& s, size_t) {
      bamgpu_opts eo{};
      eo.device = s.dev;
      eo.flags = opts.flags & BAMGPU_NO_CRC;
      s.B = bamgpu_engine_new(&eo);
      if (!s.B) { s.fail(BAMGPU_ERR_CUDA, "engine for the bucket"); return; }
      s.B->n_ref = n_ref;
      int r = engine_adopt_stream(s.B, s.recv_bytes, s.B->stream);
      if (r < 0) { s.fail(r, s.B->err); return; }
      SHCK(cudaStreamSynchronize(s.B->stream));
    });
    if (shard_failed(sh, &rc, &msg)) break;
    for_each_shard(sh, [&](SortShard& s, size_t g) {
      for (int d = 0; d < G; ++d) {
        const uint64_t m = s.seg[d + 1] - s.seg[d];
        if (m) SHCK(cudaMemcpyPeerAsync(sh[d].B->out_ptr() + roff[g][d], sh[d].dev, s.d_send.as<uint8_t>() + s.seg[d], s.dev, m, s.A->stream));
      }
      SHCK(cudaStreamSynchronize(s.A->stream));
    });
    if (shard_failed(sh, &rc, &msg)) break;
    for (int g = 0; g < G; ++g)
      for (int d = 0; d < G; ++d) if (sh[g].dev != sh[d].dev) p2p_bytes += sh[g].seg[d + 1] - sh[g].seg[d];
    // the ranges are no longer needed: their memory goes back before the buckets are sorted
    for (auto& s : sh) {
      cudaSetDevice(s.dev);
      bamgpu_stats st; bamgpu_engine_stats(s.A, &st); add_stats(acc, st);
      bamgpu_engine_free(s.A); s.A = nullptr;
      s.d_comp.release(); s.d_send.release(); s.d_soff.release();
    }
    phase("exchange");
    // 7. every GPU parses and sorts its bucket
    const uint32_t prefix = out_fmt == BAMGPU_OUT_BODIES ? 0u : 4u;
    {
      uint64_t carry = out_fmt == BAMGPU_OUT_BAM ? 4 + header.size() : 0;
      for (int d = 0; d < G; ++d) {
        sh[d].front = carry;
        sh[d].out_len = sh[d].recv_bytes - (prefix ? 0 : 4 * sh[d].recv_records);
        const uint64_t L = carry + sh[d].out_len;
        sh[d].framed = (out_fmt == BAMGPU_OUT_BAM && d < G - 1) ? L / BGZF_STORE_PAYLOAD * BGZF_STORE_PAYLOAD : L;
        carry = L - sh[d].framed;
      }
    }
    for_each_shard(sh, [&](SortShard& s, size_t) {
      bamgpu_engine* E = s.B;
      if (s.recv_records == 0) {
        SHCK(E->d_sorted.reserve(s.front + 64));
        s.stream = E->d_sorted.as<uint8_t>();
        return;
      }
      bamgpu_batch b;
      uint64_t tail = 0;
      int r = engine_records(E, 0, 1, rec_flags, E->stream, &b, &tail);
      if (r < 0) { s.fail(r, E->err); return; }
      if (b.n_records != s.recv_records || tail != s.recv_bytes) { s.fail(BAMGPU_ERR_STATE, "internal: the exchanged bucket does not parse back to its records"); return; }
      E->sort_front = s.front;
      bamgpu_sorted srt;
      r = bamgpu_engine_sort(E, sort_by, prefix ? BAMGPU_SORT_EMIT_BLOCK_SIZE : 0, nullptr, &srt);
      E->sort_front = 0;
      if (r < 0) { s.fail(r, E->err); return; }
      if (srt.bodies_len != s.out_len) { s.fail(BAMGPU_ERR_STATE, "internal: bucket output size"); return; }
      s.stream = E->d_sorted.as<uint8_t>();
    });
    if (shard_failed(sh, &rc, &msg)) break;
    phase("K3 + sort of the buckets");
    // 8. output
    std::vector<OutSeg> segs;
    if (out_fmt == BAMGPU_OUT_BAM) {
      {   // "BAM\1" + header in front of bucket 0; the odd tail of bucket d in front of bucket d + 1
        std::vector<uint8_t> hd(4 + header.size());
        memcpy(hd.data(), "BAM\1", 4);
        memcpy(hd.data() + 4, header.data(), header.size());
        cudaSetDevice(sh[0].dev);
        if (cudaMemcpy(sh[0].stream, hd.data(), hd.size(), cudaMemcpyHostToDevice) != cudaSuccess) { rc = BAMGPU_ERR_CUDA; msg = "header upload failed"; break; }
      }
      for (int d = 1; d < G && rc == BAMGPU_OK; ++d) {
        const uint64_t c = sh[d].front;
        if (c && cudaMemcpyPeer(sh[d].stream, sh[d].dev, sh[d - 1].stream + sh[d - 1].framed, sh[d - 1].dev, c) != cudaSuccess) { rc = BAMGPU_ERR_CUDA; msg = "bucket edge copy failed"; }
        if (sh[d].dev != sh[d - 1].dev) p2p_bytes += c;
      }
      if (rc != BAMGPU_OK) break;
      for_each_shard(sh, [&](SortShard& s, size_t g) {
        const bool last = (int)g == G - 1;
        if (s.framed == 0 && !last) return;
        const size_t bound = bamgpu_bgzf_bound(s.framed);
        SHCK(s.B->d_bgzf.reserve(bound + 64));
        int r = bamgpu_engine_bgzf_write(s.B, s.stream, s.framed, s.B->d_bgzf.as<uint8_t>(), bound, &s.bgzf_len, last ? 1 : 0, opts.out_compr_level, nullptr);
        if (r < 0) s.fail(r, s.B->err);
      });
      if (shard_failed(sh, &rc, &msg)) break;
      for (auto& s : sh) if (s.bgzf_len) segs.push_back({s.dev, s.B->stream, s.B->d_bgzf.as<uint8_t>(), s.bgzf_len});
    } else {
      for (auto& s : sh) if (s.out_len) segs.push_back({s.dev, s.B->stream, s.stream, s.out_len});
    }
    rc = write_segments(segs, write, write_ctx, &out_bytes, &ms_d2h, msg);
    phase("output to the sink");
  } while (0);
#undef SHCK
  cleanup();
  return rc;
}

// The whole BGZF file sits in host memory at buf[0..len) (not copied, not freed here).
static int sort_bam_buffer(const uint8_t* buf, size_t len, bamgpu_write_fn write, void* write_ctx, size_t reader_thread_num,
                           int sort_by, const bamgpu_opts* opts, bamgpu_stats* stats, char* err_buf, size_t err_cap) {
  auto fail = [&](int rc, const std::string& msg) {
    if (err_buf && err_cap) snprintf(err_buf, err_cap, "%s", msg.c_str());
    return rc;
  };
  // several GPUs: the sample sort over them (falls through to one GPU when it does not apply)
  if (opts && opts->n_devices > 1) {
    if (opts->n_devices > BAMGPU_MAX_DEVICES) return fail(BAMGPU_ERR_ARG, "too many devices");
    std::string m;
    const int r = sort_bam_multi(buf, len, write, write_ctx, sort_by, *opts, stats, m);
    if (r != SORT_MULTI_NOT_APPLICABLE) return r == BAMGPU_OK ? r : fail(r, m.empty() ? bamgpu_status_str(r) : m);
  }
  {   // does it fit in HBM as one batch?  compressed + inflated + gathered output + columns and sort scratch (~110 B per record)
    uint64_t isize_total = 0;
    size_t nblk = 0, consumed = 0;
    bamgpu_scan_blocks(buf, len, 1, nullptr, 0, &nblk, &consumed, &isize_total);
    size_t free_b = 0, total_b = 0;
    if (opts && opts->device >= 0) cudaSetDevice(opts->device);
    else if (opts && opts->n_devices > 0) cudaSetDevice(opts->devices[0]);
    const bool have_dev = cudaMemGetInfo(&free_b, &total_b) == cudaSuccess;
    const uint64_t need = (uint64_t)len + 2 * isize_total + isize_total / 3 + ((uint64_t)1 << 30);
    if (have_dev && need > (uint64_t)free_b + dev_cache_bytes()) {
      MemSrc src{buf, len, 0};
      std::string m;
      const int r = sort_bam_streamed(memsrc_read, &src, write, write_ctx, reader_thread_num, sort_by, opts, stats, m);
      return r == BAMGPU_OK ? r : fail(r, m.empty() ? bamgpu_status_str(r) : m);
    }
  }
  // 2. one batch = the whole file
  bamgpu_opts o{};
  if (opts) o = *opts; else o.device = -1;
  o.batch_bytes = len + 65536;
  if (o.n_devices > 0 && o.device < 0) o.device = o.devices[0];
  o.n_devices = 0;            // one batch on one device (a device list is served by sort_bam_multi above when it applies)
  o.jobs_per_device = 2;
  o.flags &= ~(uint32_t)(BAMGPU_COPY_DATA | BAMGPU_COPY_COLUMNS | BAMGPU_NO_RECORDS | BAMGPU_SPECULATIVE_START);
  if (sort_by == BAMGPU_SORT_NAME_AND_MATCH_MATES) o.flags |= BAMGPU_WANT_HI;
  reader_thread_num = std::max<size_t>(reader_thread_num, 1);   // sort.rs:150
  bamgpu_reader* R = bamgpu_reader_from_memory(buf, len, reader_thread_num, &o);
  if (!R) return fail(BAMGPU_ERR_CUDA, "no usable CUDA device (this library has no CPU fallback)");
  int rc = BAMGPU_OK;
  std::string msg;
  do {
    const uint8_t* hb; size_t hl, ho;
    rc = bamgpu_read_header(R, &hb, &hl, &ho);   // read and discarded: sort.rs:152-153
    if (rc != BAMGPU_OK) { msg = bamgpu_last_error(R); break; }
    bamgpu_batch b;
    rc = bamgpu_next_batch(R, &b);
    if (rc == BAMGPU_EOF) { rc = BAMGPU_OK; break; }   // no records: empty output
    if (rc != BAMGPU_OK) { msg = bamgpu_last_error(R); break; }
    bamgpu_engine* E = bamgpu_reader_current_engine(R);
    {   // the whole file must have been ONE batch: a second one means it did not fit
      bamgpu_batch b2;
      int r2 = bamgpu_next_batch(R, &b2);   // releases the first batch's job, whose device buffers stay untouched (no input left to launch)
      if (r2 == BAMGPU_OK) { rc = BAMGPU_ERR_TOO_LARGE; msg = "input was split into several batches"; break; }
      if (r2 < 0) { rc = r2; msg = bamgpu_last_error(R); break; }
    }
    // output format (bamgpu_opts.sort_output): the reference's bodies, its spill wire format, or a complete BAM
    const int out_fmt = opts ? opts->sort_output : BAMGPU_OUT_BODIES;
    if (out_fmt < BAMGPU_OUT_BODIES || out_fmt > BAMGPU_OUT_BAM) { rc = BAMGPU_ERR_ARG; msg = "unknown sort_output"; break; }
    const size_t front = out_fmt == BAMGPU_OUT_BAM ? 4 + hl : 0;
    E->sort_front = front;
    bamgpu_sorted srt;
    rc = bamgpu_engine_sort(E, sort_by, out_fmt == BAMGPU_OUT_BODIES ? 0 : BAMGPU_SORT_EMIT_BLOCK_SIZE, nullptr, &srt);
    E->sort_front = 0;
    if (rc != BAMGPU_OK) { msg = E->err; break; }
    const uint8_t* d_result = srt.bodies;
    uint64_t result_len = srt.bodies_len;
    if (out_fmt == BAMGPU_OUT_BAM) {
      // "BAM\1" + the header bytes read_header returned, in front of the sorted records; then BGZF framing
      std::vector<uint8_t> hd(front);
      memcpy(hd.data(), "BAM\1", 4);
      memcpy(hd.data() + 4, hb, hl);
      uint8_t* d_stream = E->d_sorted.as<uint8_t>();
      if (cudaMemcpyAsync(d_stream, hd.data(), front, cudaMemcpyHostToDevice, E->stream) != cudaSuccess ||
          cudaStreamSynchronize(E->stream) != cudaSuccess) { rc = BAMGPU_ERR_CUDA; msg = "header upload failed"; break; }
      const size_t bound = bamgpu_bgzf_bound(front + srt.bodies_len);
      if (E->d_bgzf.reserve(bound + 64) != cudaSuccess) { rc = BAMGPU_ERR_NOMEM; msg = "device memory for the BGZF output"; break; }
      size_t wrote = 0;
      rc = bamgpu_engine_bgzf_write(E, d_stream, front + srt.bodies_len, E->d_bgzf.as<uint8_t>(), bound, &wrote, 1, opts ? opts->out_compr_level : 0, nullptr);
      if (rc != BAMGPU_OK) { msg = E->err; break; }
      d_result = E->d_bgzf.as<uint8_t>();
      result_len = wrote;
    }
    // 3. bodies to the sink through two pinned staging buffers (the copy of chunk k+1 overlaps write(k))
    const size_t CH = (size_t)64 << 20;
    PinBuf st[2];
    cudaEvent_t ev[2];
    if (st[0].reserve(CH) != cudaSuccess || st[1].reserve(CH) != cudaSuccess) { rc = BAMGPU_ERR_NOMEM; msg = "pinned staging"; break; }
    cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming);
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    cudaEventRecord(t0, E->stream);
    const uint64_t total = result_len;
    uint64_t issued = 0, written = 0;
    int k = 0;
    size_t pending[2] = {0, 0};
    auto issue = [&](int slot) {
      size_t m = (size_t)std::min<uint64_t>(CH, total - issued);
      cudaMemcpyAsync(st[slot].p, d_result + issued, m, cudaMemcpyDeviceToHost, E->stream);
      cudaEventRecord(ev[slot], E->stream);
      pending[slot] = m;
      issued += m;
    };
    if (total) issue(0);
    while (written < total) {
      int slot = k & 1;
      if (issued < total) issue(slot ^ 1);
      if (cudaEventSynchronize(ev[slot]) != cudaSuccess) { rc = BAMGPU_ERR_CUDA; msg = "D2H of sorted bodies failed"; break; }
      if (write(write_ctx, st[slot].as<uint8_t>(), pending[slot]) < 0) { rc = BAMGPU_ERR_IO; msg = "write callback failed"; break; }
      written += pending[slot];
      ++k;
    }
    cudaEventRecord(t1, E->stream);
    cudaStreamSynchronize(E->stream);
    float ms = 0;
    cudaEventElapsedTime(&ms, t0, t1);
    E->stats.ms_d2h += ms;
    cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]); cudaEventDestroy(t0); cudaEventDestroy(t1);
    st[0].release(); st[1].release();
  } while (0);
  if (stats) bamgpu_reader_stats(R, stats);
  bamgpu_reader_free(R);
  if (rc != BAMGPU_OK) return fail(rc, msg.empty() ? bamgpu_status_str(rc) : msg);
  return BAMGPU_OK;
}

extern "C" int bamgpu_sort_bam(bamgpu_read_fn read, void* read_ctx, bamgpu_write_fn write, void* write_ctx, size_t reader_thread_num,
                               int sort_by, const bamgpu_opts* opts, bamgpu_stats* stats, char* err_buf, size_t err_cap) {
  auto fail = [&](int rc, const std::string& msg) {
    if (err_buf && err_cap) snprintf(err_buf, err_cap, "%s", msg.c_str());
    return rc;
  };
  if (!read || !write || sort_by < BAMGPU_SORT_NAME || sort_by > BAMGPU_SORT_COORDINATES_AND_STRAND) return fail(BAMGPU_ERR_ARG, "bad argument");
  if (opts && opts->sort_mem_limit) {   // sort.rs:138 mem_limit: the streamed form, straight from the caller's source
    std::string m;
    const int r = sort_bam_streamed(read, read_ctx, write, write_ctx, reader_thread_num, sort_by, opts, stats, m);
    return r == BAMGPU_OK ? r : fail(r, m.empty() ? bamgpu_status_str(r) : m);
  }
  // the whole input (its size decides how it is sorted; the reference's reading thread streams it instead, readahead.rs:88-118)
  size_t cap = (size_t)64 << 20, len = 0;
  uint8_t* buf = (uint8_t*)malloc(cap);
  if (!buf) return fail(BAMGPU_ERR_NOMEM, "out of host memory");
  for (;;) {
    if (len == cap) {
      cap *= 2;
      uint8_t* nb = (uint8_t*)realloc(buf, cap);
      if (!nb) { free(buf); return fail(BAMGPU_ERR_NOMEM, "out of host memory"); }
      buf = nb;
    }
    long got = read(read_ctx, buf + len, std::min(cap - len, (size_t)256 << 20));
    if (got < 0) { free(buf); return fail(BAMGPU_ERR_IO, "read callback failed"); }
    if (got == 0) break;
    len += (size_t)got;
  }
  const int rc = sort_bam_buffer(buf, len, write, write_ctx, reader_thread_num, sort_by, opts, stats, err_buf, err_cap);
  free(buf);
  return rc;
}

extern "C" int bamgpu_sort_bam_from_memory(const uint8_t* buf, size_t len, bamgpu_write_fn write, void* write_ctx, size_t reader_thread_num,
                                           int sort_by, const bamgpu_opts* opts, bamgpu_stats* stats, char* err_buf, size_t err_cap) {
  if ((len && !buf) || !write || sort_by < BAMGPU_SORT_NAME || sort_by > BAMGPU_SORT_COORDINATES_AND_STRAND) {
    if (err_buf && err_cap) snprintf(err_buf, err_cap, "bad argument");
    return BAMGPU_ERR_ARG;
  }
  if (opts && opts->sort_mem_limit) {
    MemSrc src{buf, len, 0};
    std::string m;
    const int r = sort_bam_streamed(memsrc_read, &src, write, write_ctx, reader_thread_num, sort_by, opts, stats, m);
    if (r != BAMGPU_OK && err_buf && err_cap) snprintf(err_buf, err_cap, "%s", m.empty() ? bamgpu_status_str(r) : m.c_str());
    return r;
  }
  return sort_bam_buffer(buf, len, write, write_ctx, reader_thread_num, sort_by, opts, stats, err_buf, err_cap);
}
