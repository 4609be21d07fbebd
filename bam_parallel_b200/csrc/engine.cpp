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

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
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
  static constexpr size_t LIMIT = (size_t)8 << 30, ITEM_LIMIT = (size_t)2 << 30;
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
};
DevCache& dev_cache() { static DevCache* c = new DevCache(); return *c; }

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
    if (void* q = dev_cache().take(want, dev, &got)) { p = q; cap = got; return cudaSuccess; }
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
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
struct PinnedCache {
  std::mutex mu;
  std::vector<std::pair<void*, size_t>> free_list;
  size_t bytes = 0;
  static constexpr size_t LIMIT = (size_t)12 << 30;
  void* take(size_t n, size_t* cap) {
    std::lock_guard<std::mutex> lk(mu);
    size_t best = (size_t)-1;
    for (size_t i = 0; i < free_list.size(); ++i)
      if (free_list[i].second >= n && (best == (size_t)-1 || free_list[i].second < free_list[best].second)) best = i;
    if (best == (size_t)-1 || free_list[best].second > n + n / 4 + (1 << 20)) return nullptr;   // a snug fit only: a big buffer taken for a small need is missed later
    void* p = free_list[best].first;
    *cap = free_list[best].second;
    bytes -= *cap;
    free_list.erase(free_list.begin() + best);
    return p;
  }
  bool give(void* p, size_t cap) {
    std::lock_guard<std::mutex> lk(mu);
    if (bytes + cap > LIMIT) return false;
    free_list.emplace_back(p, cap);
    bytes += cap;
    return true;
  }
};
PinnedCache& pinned_cache() { static PinnedCache* c = new PinnedCache(); return *c; }

struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    release();
    size_t want = n + n / 8 + 256;
    size_t got = 0;
    if (void* q = pinned_cache().take(want, &got)) { p = q; cap = got; return cudaSuccess; }
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p && !pinned_cache().give(p, cap)) cudaFreeHost(p);
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

  // block table (SoA) staging + device
  PinBuf h_tab;
  DevBuf d_tab;
  DevBuf d_status, d_misc;  // d_misc: work counter (u32 @0), first_bad (u64 @8), ChainResult @64
  DevBuf d_crc_tables, d_k1ring;
  bool crc_ready = false;
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
  uint64_t data_len = 0;    // bytes valid in d_out (after lead pad), carry included
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
static int dbg_level() { static int v = -1; if (v < 0) { const char* e = getenv("BAMGPU_DEBUG"); v = e ? atoi(e) : 0; } return v; }

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
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) E->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&E->stream, cudaStreamNonBlocking) != cudaSuccess) { delete E; return nullptr; }
  if (cudaStreamCreateWithFlags(&E->d2h_stream, cudaStreamNonBlocking) != cudaSuccess) { delete E; return nullptr; }
  cudaEventCreateWithFlags(&E->ev_k3_done, cudaEventDisableTiming);
  for (auto& M : E->mir) { cudaEventCreate(&M.ev_begin); cudaEventCreate(&M.ev_end); }
  for (auto& e : E->ev) cudaEventCreate(&e);
  E->default_flags = opts ? opts->flags : 0;
  if (E->d_misc.reserve(8192) != cudaSuccess || E->d_crc_tables.reserve(CRC_TABLE_WORDS * 4) != cudaSuccess ||
      E->h_res.reserve(4096) != cudaSuccess) {
    bamgpu_engine_free(E);
    return nullptr;
  }
  crc32_init_tables(E->d_crc_tables.as<uint32_t>(), E->stream);
  if (cudaStreamSynchronize(E->stream) != cudaSuccess) { bamgpu_engine_free(E); return nullptr; }
  E->crc_ready = true;
  return E;
}

extern "C" void bamgpu_engine_free(bamgpu_engine* E) {
  if (!E) return;
  cudaSetDevice(E->device);
  if (E->stream) { cudaStreamSynchronize(E->stream); cudaStreamDestroy(E->stream); }
  if (E->d2h_stream) { cudaStreamSynchronize(E->d2h_stream); cudaStreamDestroy(E->d2h_stream); }
  if (E->ev_k3_done) cudaEventDestroy(E->ev_k3_done);
  for (auto& M : E->mir) { if (M.ev_begin) cudaEventDestroy(M.ev_begin); if (M.ev_end) cudaEventDestroy(M.ev_end); M.d_cols.release(); M.h_data.release(); M.h_cols.release(); }
  for (auto& e : E->ev) if (e) cudaEventDestroy(e);
  E->h_tab.release(); E->d_tab.release(); E->d_status.release(); E->d_misc.release(); E->d_crc_tables.release();
  E->d_outs[0].release(); E->d_outs[1].release(); E->d_tiles.release(); E->d_k1ring.release(); E->h_res.release();
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

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

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
static int engine_inflate_launch(bamgpu_engine* E, const uint8_t* d_comp, const bamgpu_block* blocks, size_t n_blocks,
                                 uint32_t flags, cudaStream_t s) {
  const double t_in = now_ms();
  ECK(cudaSetDevice(E->device));
  // 1. the destination buffer (the other one when ping-ponging) gets the carry at its front
  const uint64_t carry = E->carry_len;
  uint64_t total = 0, bytes_in = 0;
  for (size_t i = 0; i < n_blocks; ++i) { total += blocks[i].isize; bytes_in += blocks[i].in_len + 26; }
  const uint64_t new_len = carry + total;
  const int src = E->cur, dst = E->pingpong ? (E->cur ^ 1) : E->cur;
  const uint8_t* src_ptr = E->d_outs[src].p ? E->d_outs[src].as<uint8_t>() + OUT_LEAD_PAD : nullptr;
  const size_t need = OUT_LEAD_PAD + new_len + BAMGPU_MAX_HALO + OUT_TAIL_PAD;
  if (E->d_outs[dst].cap < need) {
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
  ECK(cudaEventRecord(E->ev[0], s));
  if (inflate_ring_bytes(E->num_sms)) ECK(E->d_k1ring.reserve(inflate_ring_bytes(E->num_sms)));
  launch_inflate(d_comp, t, E->out_ptr(), E->d_status.as<uint32_t>(), counter, E->d_k1ring.p, E->num_sms, s);
  ECK(cudaEventRecord(E->ev[1], s));
  E->stats.kernel_launches += 1;
  if (!(flags & BAMGPU_NO_CRC)) {
    launch_crc32(E->out_ptr(), t, E->d_crc_tables.as<uint32_t>(), E->d_status.as<uint32_t>(), E->num_sms, s);
    E->stats.kernel_launches += 1;
  }
  ECK(cudaEventRecord(E->ev[2], s));
  launch_status_reduce(E->d_status.as<uint32_t>(), (uint32_t)nb, first_bad, s);
  E->stats.kernel_launches += 1;
  launch_copy_small_to_host(E->h_res.p, first_bad, 8, s);
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
  cudaEventElapsedTime(&ms, E->ev[1], E->ev[2]); E->stats.ms_crc += ms;
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
           tiles_bytes = align_up(o_reach + nt, 256);
    ECK(E->d_tiles.reserve(tiles_bytes));
    uint8_t* tb = E->d_tiles.as<uint8_t>();
    TileArrays t{(uint64_t*)(tb + o_entry), (uint64_t*)(tb + o_exit), (uint32_t*)(tb + o_count), (uint32_t*)(tb + o_kind),
                 (uint32_t*)(tb + o_nxt), (uint32_t*)(tb + o_jump), (uint32_t*)(tb + o_jump2), (uint8_t*)(tb + o_reach),
                 (uint64_t*)(tb + o_base), n_tiles};
    launch_tile_walk(p, t, s);
    ECK(launch_chain_resolve(p, t, d_res, (ResolveScratch*)(E->d_misc.as<uint8_t>() + 256), E->num_sms, s));
    E->stats.kernel_launches += 2;
    launch_copy_small_to_host(h_res, d_res, sizeof(ChainResult), s);
    ECK(cudaStreamSynchronize(s));
    ECK(cudaGetLastError());
    if (h_res->broken_tile != 0xffffffffu) { E->err = "internal: record chain left the buffer"; return BAMGPU_ERR_STATE; }
    out->chain_start = h_res->chain_start;
    out->chain_repairs = h_res->rounds;
    n_rec = h_res->n_records;
    tail_off = h_res->tail_off;
    tail_kind = h_res->tail_kind;
    if (n_rec) {
      if (n_rec > M.cols_cap) {
        uint64_t cap = n_rec + n_rec / 8 + 1024;
        ECK(M.d_cols.reserve(columns_bytes(cap)));
        M.cols_cap = cap;
        carve_columns(M.d_cols.as<uint8_t>(), cap, M.cols);
      }
      launch_tile_emit(p, t, M.cols.rec_off, M.cols.rec_len, s);
      launch_extract(p, M.cols, n_rec, (flags & BAMGPU_WANT_HI) != 0, d_res, s);
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
  fill_public_columns(M.cols, out->dev);
  // host mirrors: on their own stream, behind the record kernels
  cudaStream_t cs = E->d2h_stream;
  const bool any_copy = (flags & BAMGPU_COPY_DATA) || ((flags & (BAMGPU_COPY_COLUMNS | BAMGPU_COPY_OFFSETS)) && n_rec);
  if (any_copy) {
    ECK(cudaEventRecord(E->ev_k3_done, s));
    ECK(cudaStreamWaitEvent(cs, E->ev_k3_done, 0));
    ECK(cudaEventRecord(M.ev_begin, cs));
  }
  if (flags & BAMGPU_COPY_DATA) {
    ECK(M.h_data.reserve(E->data_len + 64));
    if (E->data_len) ECK(cudaMemcpyAsync(M.h_data.p, E->out_ptr(), E->data_len, cudaMemcpyDeviceToHost, cs));
    out->h_data = M.h_data.as<uint8_t>();
  }
  if ((flags & BAMGPU_COPY_OFFSETS) && !(flags & BAMGPU_COPY_COLUMNS) && n_rec) {
    // what next_rec() needs besides the bodies: where each one starts and how long it is
    ECK(M.h_cols.reserve(columns_bytes(M.cols_cap)));
    ColumnsDev hc;
    carve_columns(M.h_cols.as<uint8_t>(), M.cols_cap, hc);
    ECK(cudaMemcpyAsync(hc.rec_off, M.cols.rec_off, 8 * n_rec, cudaMemcpyDeviceToHost, cs));
    ECK(cudaMemcpyAsync(hc.rec_len, M.cols.rec_len, 4 * n_rec, cudaMemcpyDeviceToHost, cs));
    out->host.rec_off = hc.rec_off;
    out->host.rec_len = hc.rec_len;
  }
  if ((flags & BAMGPU_COPY_COLUMNS) && n_rec) {
    size_t bytes = columns_bytes(M.cols_cap);
    ECK(M.h_cols.reserve(bytes));
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
  ECK(cudaStreamSynchronize(s));
  ECK(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, E->ev[3], E->ev[4]); E->stats.ms_records += ms;
  if (any_copy) {
    ECK(cudaEventRecord(M.ev_end, cs));
    M.pending = true;
    if (!defer_d2h) { int w = engine_wait_mirror(E, E->msel); if (w < 0) return w; }
  }
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
  return rc;
}

extern "C" int bamgpu_engine_decode(bamgpu_engine* E, const uint8_t* d_comp, const bamgpu_block* blocks, size_t n_blocks,
                                    uint64_t stream_start, int final, uint32_t flags, void* cuda_stream,
                                    bamgpu_batch* out, uint64_t* tail_off) {
  if (!E || !out || (n_blocks && (!blocks || !d_comp))) return BAMGPU_ERR_ARG;
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : E->stream;
  flags |= E->default_flags;
  int rc = engine_inflate(E, d_comp, blocks, n_blocks, flags, s);
  if (rc != BAMGPU_OK) return rc;
  return engine_records(E, stream_start, final, flags, s, out, tail_off);
}

extern "C" int bamgpu_engine_inflate(bamgpu_engine* E, const uint8_t* d_comp, const bamgpu_block* blocks, size_t n_blocks,
                                     uint32_t flags, void* cuda_stream) {
  if (!E || (n_blocks && (!blocks || !d_comp))) return BAMGPU_ERR_ARG;
  E->carry_len = 0;
  E->keep_from = 0;
  return engine_inflate(E, d_comp, blocks, n_blocks, flags | E->default_flags, cuda_stream ? (cudaStream_t)cuda_stream : E->stream);
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
  return engine_records(E, stream_start, final, flags | E->default_flags, cuda_stream ? (cudaStream_t)cuda_stream : E->stream, out, tail_off);
}

extern "C" int bamgpu_engine_stats(const bamgpu_engine* E, bamgpu_stats* out) {
  if (!E || !out) return BAMGPU_ERR_ARG;
  *out = E->stats;
  out->ms_total = out->ms_h2d + out->ms_inflate + out->ms_crc + out->ms_records + out->ms_d2h + out->ms_sort + out->ms_gather;
  return BAMGPU_OK;
}
extern "C" void bamgpu_engine_reset_stats(bamgpu_engine* E) { if (E) E->stats = bamgpu_stats{}; }

extern "C" void bamgpu_engine_set_n_ref(bamgpu_engine* E, int32_t n) { if (E) E->n_ref = n; }
static void engine_set_n_ref(bamgpu_engine* E, int32_t n) { E->n_ref = n; }

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
  ECK(sorted_offsets(E->mir[E->msel].cols, E->d_perm.as<uint32_t>(), n, E->d_sort.p, E->d_soff.as<uint64_t>(), &total, s));   // synchronises
  if (sort_by == BAMGPU_SORT_NAME_AND_MATCH_MATES) {
    if (h_sr->hi_bad_tags) { E->err = "sort: aux tags the reference's HI walk panics on"; return BAMGPU_ERR_SORT_BAD_TAGS; }
    if (h_sr->hi_mixed) {
      E->err = "sort: " + std::to_string(h_sr->hi_mixed) + " pair(s) of equal read names with HI on one record only (Option::unwrap() panics, comparators.rs:90)";
      return BAMGPU_ERR_SORT_HI_MIXED;
    }
  }
  ECK(E->d_sorted.reserve(total + 64));
  ECK(gather_sorted_bodies(E->out_ptr(), E->mir[E->msel].cols, E->d_perm.as<uint32_t>(), n, E->d_sort.p, E->d_soff.as<uint64_t>(),
                           E->d_sorted.as<uint8_t>(), s));
  ECK(cudaEventRecord(E->ev[5], s));
  launches += 3;
  out->bodies_len = total;
  out->perm = E->d_perm.as<uint32_t>();
  out->body_off = E->d_soff.as<uint64_t>();
  out->bodies = E->d_sorted.as<uint8_t>();
  if ((flags & BAMGPU_COPY_COLUMNS) && n) {
    ECK(E->h_perm.reserve(4 * n));
    ECK(cudaMemcpyAsync(E->h_perm.p, E->d_perm.p, 4 * n, cudaMemcpyDeviceToHost, s));
    out->h_perm = E->h_perm.as<uint32_t>();
  }
  if ((flags & BAMGPU_COPY_DATA) && total) {
    ECK(E->h_sorted.reserve(total));
    ECK(cudaMemcpyAsync(E->h_sorted.p, E->d_sorted.p, total, cudaMemcpyDeviceToHost, s));
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

// =================================================================================================
// Reader
// =================================================================================================
namespace {

struct Slot {
  PinBuf host;               // compressed bytes as read from the source
  DevBuf dev;                // same bytes on the device
  size_t len = 0;            // bytes covered by whole blocks
  std::vector<bamgpu_block> blocks;
  bool final = false;        // no more input after this slot
  int status = BAMGPU_OK;    // scan / io status for this slot
  cudaEvent_t copied = nullptr;
};

}  // namespace

struct bamgpu_reader {
  bamgpu_engine* E = nullptr;
  bamgpu_read_fn read = nullptr;
  void* ctx = nullptr;
  size_t thread_num = 3;
  uint64_t track_total = 0;
  uint64_t bytes_fed = 0;
  uint32_t flags = 0;
  size_t batch_bytes = 256u << 20;
  std::string err;
  // memory source
  const uint8_t* mem = nullptr;
  size_t mem_len = 0, mem_pos = 0;
  // feeder
  static constexpr int N_SLOTS = 3;
  Slot slots[N_SLOTS];
  std::thread feeder;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<int> ready, freeq;
  bool stop = false;
  cudaStream_t copy_stream = nullptr;
  // consumer state
  int cur_slot = -1;
  bool header_done = false, eof = false, raw_mode = false, inflated_pending = false, input_done = false;
  int sticky = BAMGPU_OK;
  std::vector<uint8_t> header;      // without magic
  size_t off_to_refs = 0;
  uint64_t next_start = 0;          // chain start inside the next engine_records call
  bamgpu_batch batch{};
  bool batch_valid = false;
  bamgpu_batch prep{};              // the batch staged behind `batch`: record kernels done, D2H queued
  bool prep_valid = false, no_more = false;
  int prep_sel = 0;
  int sticky_next = BAMGPU_OK;
  uint64_t rec_cursor = 0;          // next record of `batch` to hand out
  uint64_t byte_cursor = 0;         // raw mode: next byte of batch.h_data
  bool pending_final = false;
};

static long mem_read(void* ctx, uint8_t* dst, size_t cap) {
  bamgpu_reader* R = (bamgpu_reader*)ctx;
  size_t n = std::min(cap, R->mem_len - R->mem_pos);
  memcpy(dst, R->mem + R->mem_pos, n);
  R->mem_pos += n;
  return (long)n;
}

// Feeder thread: readahead.rs:88-118's reading thread, but it batches whole blocks into a pinned slot,
// scans their headers AND footers, and starts the H2D copy on its own stream.
static void feeder_main(bamgpu_reader* R) {
  cudaSetDevice(R->E->device);
  std::vector<uint8_t> carry;  // partial block left over from the previous slot
  bool src_eof = false;
  for (;;) {
    int si;
    {
      std::unique_lock<std::mutex> lk(R->mu);
      R->cv.wait(lk, [&] { return R->stop || !R->freeq.empty(); });
      if (R->stop) return;
      si = R->freeq.front();
      R->freeq.pop_front();
    }
    Slot& S = R->slots[si];
    S.blocks.clear();
    S.status = BAMGPU_OK;
    S.final = false;
    // (Smaller first batches were tried to fill the pipeline sooner: every growth of the output buffers then costs a
    // cudaFree, which waits for the D2H in flight - slower overall.)
    const size_t bb = R->batch_bytes;
    size_t cap = R->batch_bytes + 65536 + 64;
    if (S.host.reserve(cap + IN_TAIL_PAD) != cudaSuccess || S.dev.reserve(cap + IN_TAIL_PAD) != cudaSuccess) {
      S.status = BAMGPU_ERR_NOMEM; S.final = true;
    } else {
      uint8_t* h = S.host.as<uint8_t>();
      size_t fill = carry.size();
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
        if (n < 0) { S.status = BAMGPU_ERR_IO; src_eof = true; break; }
        if (n == 0) { src_eof = true; break; }
        fill += (size_t)n;
        R->bytes_fed += (uint64_t)n;
      }
      if (S.status == BAMGPU_OK) {
        size_t nb = 0, consumed = 0;
        uint64_t tot = 0;
        // count, then fill
        int rc = bamgpu_scan_blocks(h, fill, src_eof ? 1 : 0, nullptr, 0, &nb, &consumed, &tot);
        S.blocks.resize(nb);
        if (nb) bamgpu_scan_blocks(h, fill, src_eof ? 1 : 0, S.blocks.data(), nb, &nb, &consumed, &tot);
        S.len = consumed;
        if (rc < 0) { S.status = rc; S.final = true; }
        else if (rc == BAMGPU_EOF || src_eof) S.final = true;   // silent EOF (short header / BSIZE wrap) ends the stream
        else carry.assign(h + consumed, h + fill);
        if (S.len) {
          memset(h + S.len, 0, IN_TAIL_PAD);
          cudaMemcpyAsync(S.dev.p, h, S.len + IN_TAIL_PAD, cudaMemcpyHostToDevice, R->copy_stream);
        }
        cudaEventRecord(S.copied, R->copy_stream);
      } else {
        S.final = true;
      }
    }
    bool fin = S.final;
    {
      std::lock_guard<std::mutex> lk(R->mu);
      R->ready.push_back(si);
    }
    R->cv.notify_all();
    if (fin) return;
  }
}

static bamgpu_reader* reader_create(bamgpu_read_fn read, void* ctx, size_t thread_num, const uint64_t* track_progress,
                                    const bamgpu_opts* opts) {
  bamgpu_engine* E = bamgpu_engine_new(opts);
  if (!E) return nullptr;
  bamgpu_reader* R = new bamgpu_reader();
  R->E = E;
  E->pingpong = true;
  R->read = read;
  R->ctx = ctx;
  unsigned hc = std::thread::hardware_concurrency();
  if (hc && thread_num > hc) thread_num = hc;      // reader.rs:33-35
  R->thread_num = std::max<size_t>(thread_num, 3);  // readahead.rs:53
  R->track_total = track_progress ? *track_progress : 0;
  R->flags = opts ? opts->flags : 0;
  if (opts && opts->batch_bytes) R->batch_bytes = (size_t)opts->batch_bytes;
  if (R->batch_bytes < 65536) R->batch_bytes = 65536;
  cudaStreamCreateWithFlags(&R->copy_stream, cudaStreamNonBlocking);
  for (int i = 0; i < bamgpu_reader::N_SLOTS; ++i) {
    cudaEventCreateWithFlags(&R->slots[i].copied, cudaEventDisableTiming);
    R->freeq.push_back(i);
  }
  return R;
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
  cudaSetDevice(R->E->device);
  if (R->copy_stream) { cudaStreamSynchronize(R->copy_stream); cudaStreamDestroy(R->copy_stream); }
  for (auto& s : R->slots) { if (s.copied) cudaEventDestroy(s.copied); s.host.release(); s.dev.release(); }
  bamgpu_engine_free(R->E);
  delete R;
}

extern "C" const char* bamgpu_last_error(const bamgpu_reader* R) {
  if (!R) return "null reader (no CUDA device? this library has no CPU fallback)";
  return R->err.empty() ? R->E->err.c_str() : R->err.c_str();
}

// Returns the current slot to the feeder and takes the next one; puts its inflate on the stream (no host wait).
// BAMGPU_EOF when input is exhausted.  With `nonblocking` it returns BAMGPU_EOF+1 (= 2) instead of waiting when the
// feeder has no slot ready yet.
static int reader_inflate_launch_next(bamgpu_reader* R, uint32_t flags, bool nonblocking) {
  if (R->input_done) return BAMGPU_EOF;
  int si;
  {
    std::unique_lock<std::mutex> lk(R->mu);
    if (nonblocking && R->ready.empty()) return 2;
    if (R->cur_slot >= 0) {  // its kernels have finished (every caller synchronises the previous inflate first)
      R->freeq.push_back(R->cur_slot);
      R->cur_slot = -1;
      R->cv.notify_all();
    }
    R->cv.wait(lk, [&] { return !R->ready.empty(); });
    si = R->ready.front();
    R->ready.pop_front();
  }
  R->cur_slot = si;
  Slot& S = R->slots[si];
  if (S.final) R->input_done = true;
  R->pending_final = S.final;
  if (S.status < 0) {
    R->err = std::string("input: ") + bamgpu_status_str(S.status);
    return S.status;
  }
  bamgpu_engine* E = R->E;
  cudaSetDevice(E->device);
  if (cudaStreamWaitEvent(E->stream, S.copied, 0) != cudaSuccess) { R->err = "cudaStreamWaitEvent failed"; return BAMGPU_ERR_CUDA; }
  int rc = engine_inflate_launch(E, S.dev.as<uint8_t>(), S.blocks.data(), S.blocks.size(), flags, E->stream);
  if (rc != BAMGPU_OK) return rc;
  R->inflated_pending = true;
  return BAMGPU_OK;
}

static int reader_inflate_next(bamgpu_reader* R, uint32_t flags) {
  int rc = reader_inflate_launch_next(R, flags, false);
  if (rc != BAMGPU_OK) return rc;
  return engine_inflate_finish(R->E, R->E->stream);
}

extern "C" int bamgpu_read_header(bamgpu_reader* R, const uint8_t** bytes, size_t* len, size_t* offset_to_ref_seqs) {
  if (!R) return BAMGPU_ERR_ARG;
  if (R->sticky < 0) return R->sticky;
  if (R->header_done || R->raw_mode) { R->err = "read_header called twice or after read()"; return BAMGPU_ERR_STATE; }
  bamgpu_engine* E = R->E;
  // Inflate batches until the whole header is on the device (normally the first batch).
  std::vector<uint8_t> hb;
  auto fetch = [&](uint64_t need) -> int {  // make hb hold the first `need` inflated bytes
    while (E->data_len < need) {
      if (R->input_done) return BAMGPU_ERR_SHORT_HEADER;
      // keep everything inflated so far as carry for the next inflate
      E->carry_len = E->data_len; E->keep_from = 0;
      int rc = reader_inflate_next(R, R->flags);
      if (rc == BAMGPU_EOF) return BAMGPU_ERR_SHORT_HEADER;
      if (rc < 0) return rc;
    }
    if (hb.size() < need) {
      size_t old = hb.size();
      hb.resize(need);
      int rc = bamgpu_memcpy_d2h(E, hb.data() + old, E->out_ptr() + old, need - old);
      if (rc < 0) return rc;
    }
    return BAMGPU_OK;
  };
  int rc = reader_inflate_next(R, R->flags);
  if (rc < 0) { R->sticky = rc; return rc; }
  auto fail = [&](int code, const char* what) { R->err = what; R->sticky = code; return code; };
  if ((rc = fetch(4)) < 0) return fail(rc == BAMGPU_ERR_SHORT_HEADER ? BAMGPU_ERR_SHORT_HEADER : rc, "short BAM header");
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
  R->next_start = p;
  engine_set_n_ref(E, n_ref > 0x7fffffffu ? -1 : (int32_t)n_ref);
  if (bytes) *bytes = R->header.data();
  if (len) *len = R->header.size();
  if (offset_to_ref_seqs) *offset_to_ref_seqs = R->off_to_refs;
  return BAMGPU_OK;
}

// Stages the next non-empty batch: waits for its inflate (launching it first if need be), runs the record kernels and
// queues its D2H (not waited for) -> R->prep / R->prep_sel.  BAMGPU_OK, BAMGPU_EOF (no more input) or an error.
static int reader_stage_next(bamgpu_reader* R, uint32_t flags) {
  bamgpu_engine* E = R->E;
  for (;;) {
    if (R->no_more) return BAMGPU_EOF;
    int rc;
    if (!R->inflated_pending) {
      rc = reader_inflate_launch_next(R, flags, false);
      if (rc == BAMGPU_EOF) { R->no_more = true; return BAMGPU_EOF; }   // whatever carry was left has been judged with final=1
      if (rc < 0) return rc;
    }
    const double t0 = now_ms();
    rc = engine_inflate_finish(E, E->stream);
    if (rc < 0) return rc;
    R->inflated_pending = false;
    const bool fin = R->pending_final;
    uint64_t tail = 0;
    const double t1 = now_ms();
    rc = engine_records(E, R->next_start, fin ? 1 : 0, flags, E->stream, &R->prep, &tail, /*defer_d2h=*/true);
    R->next_start = 0;
    if (rc < 0) return rc;
    if (rc == BAMGPU_EOF || fin) R->no_more = true;
    R->prep_sel = E->msel;
    if (dbg_level() >= 2)
      fprintf(stderr, "[bamgpu] stage @%.1f: wait inflate %.2f, records %.2f ms (%llu records)\n", t0, t1 - t0, now_ms() - t1,
              (unsigned long long)R->prep.n_records);
    if (R->prep.n_records) return BAMGPU_OK;
    { int w = engine_wait_mirror(E, E->msel); if (w < 0) return w; }   // an empty batch: drop it
  }
}

// Produces the next non-empty batch (or EOF). Uses the inflated-but-unparsed batch left by read_header first.
// Pipeline per call k:  [batch k was staged by the previous call: its D2H is running]  ->  launch inflate(k+1) into the
// device buffer the caller has just released, wait for it, K3(k+1), queue D2H(k+1) behind D2H(k) (other mirror set)  ->
// wait D2H(k)  ->  hand out batch k.  So the copy engine goes from one batch's D2H straight into the next one's, K1/K3 of
// batch k+1 run under D2H(k), the feeder thread's H2D of batch k+2 under both, and batch k's device and host pointers
// stay valid until the next call.
static int reader_next_batch(bamgpu_reader* R, uint32_t extra_flags) {
  if (R->sticky < 0) return R->sticky;
  if (R->eof) return BAMGPU_EOF;
  if (!R->header_done) { R->err = "records requested before read_header()"; return BAMGPU_ERR_STATE; }
  bamgpu_engine* E = R->E;
  const uint32_t flags = R->flags | extra_flags;
  if (!R->prep_valid) {
    if (R->sticky_next < 0) { R->sticky = R->sticky_next; return R->sticky; }
    int rc = reader_stage_next(R, flags);
    if (rc == BAMGPU_EOF) { R->eof = true; return BAMGPU_EOF; }
    if (rc < 0) { R->sticky = rc; return rc; }
    R->prep_valid = true;
  }
  const bamgpu_batch cur = R->prep;
  const int cur_sel = R->prep_sel;
  R->prep_valid = false;
  const double t0 = now_ms();
  const int rn = reader_stage_next(R, flags);
  if (rn == BAMGPU_OK) R->prep_valid = true;
  else if (rn < 0) R->sticky_next = rn;      // surfaces on the next call; this batch is still good
  const double t1 = now_ms();
  int w = engine_wait_mirror(E, cur_sel);
  if (w < 0) { R->sticky = w; return w; }
  if (dbg_level() >= 2) fprintf(stderr, "[bamgpu] next_batch @%.1f: stage next %.2f, wait d2h %.2f ms\n", t0, t1 - t0, now_ms() - t1);
  R->batch = cur;
  R->batch_valid = true;
  R->rec_cursor = 0;
  return BAMGPU_OK;
}

extern "C" int bamgpu_next_batch(bamgpu_reader* R, bamgpu_batch* out) {
  if (!R || !out) return BAMGPU_ERR_ARG;
  if (R->raw_mode) { R->err = "next_batch after read()"; return BAMGPU_ERR_STATE; }
  if (R->eof) return R->sticky < 0 ? R->sticky : BAMGPU_EOF;
  int rc = reader_next_batch(R, 0);
  if (rc != BAMGPU_OK) return rc;
  *out = R->batch;
  R->rec_cursor = R->batch.n_records;  // the batch is handed out whole
  return BAMGPU_OK;
}

extern "C" int bamgpu_next_rec(bamgpu_reader* R, const uint8_t** body, size_t* len) {
  if (!R || !body || !len) return BAMGPU_ERR_ARG;
  if (R->raw_mode) { R->err = "next_rec after read()"; return BAMGPU_ERR_STATE; }
  while (!R->batch_valid || R->rec_cursor >= R->batch.n_records) {
    if (R->eof) { *body = nullptr; *len = 0; return R->sticky < 0 ? R->sticky : BAMGPU_EOF; }
    int rc = reader_next_batch(R, BAMGPU_COPY_DATA | BAMGPU_COPY_OFFSETS);
    if (rc != BAMGPU_OK) { *body = nullptr; *len = 0; return rc; }
  }
  if (!R->batch.h_data || !R->batch.host.rec_off) { R->err = "internal: batch has no host mirror"; return BAMGPU_ERR_STATE; }
  uint64_t i = R->rec_cursor++;
  *body = R->batch.h_data + R->batch.host.rec_off[i];
  *len = R->batch.host.rec_len[i];
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

extern "C" long bamgpu_read(bamgpu_reader* R, uint8_t* dst, size_t cap) {
  if (!R || (!dst && cap)) return BAMGPU_ERR_ARG;
  if (R->sticky < 0) return R->sticky;
  if (R->header_done && !R->raw_mode) { R->err = "read() after read_header(): use the record calls"; return BAMGPU_ERR_STATE; }
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
    int rc = reader_inflate_next(R, R->flags | BAMGPU_NO_RECORDS);
    if (rc == BAMGPU_EOF) { R->eof = true; break; }
    if (rc < 0) { R->sticky = rc; return got ? (long)got : rc; }
    uint64_t tail;
    rc = engine_records(R->E, 0, R->pending_final, R->flags | BAMGPU_NO_RECORDS | BAMGPU_COPY_DATA, R->E->stream, &R->batch, &tail);
    if (rc < 0) { R->sticky = rc; return got ? (long)got : rc; }
    R->inflated_pending = false;
    R->batch_valid = true;
    R->byte_cursor = 0;
    if (R->pending_final) R->eof = true;
  }
  return (long)got;
}

extern "C" int bamgpu_reader_stats(const bamgpu_reader* R, bamgpu_stats* out) {
  if (!R) return BAMGPU_ERR_ARG;
  return bamgpu_engine_stats(R->E, out);
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
extern "C" int bamgpu_sort_bam(bamgpu_read_fn read, void* read_ctx, bamgpu_write_fn write, void* write_ctx, size_t reader_thread_num,
                               int sort_by, const bamgpu_opts* opts, bamgpu_stats* stats, char* err_buf, size_t err_cap) {
  auto fail = [&](int rc, const std::string& msg) {
    if (err_buf && err_cap) snprintf(err_buf, err_cap, "%s", msg.c_str());
    return rc;
  };
  if (!read || !write || sort_by < BAMGPU_SORT_NAME || sort_by > BAMGPU_SORT_COORDINATES_AND_STRAND) return fail(BAMGPU_ERR_ARG, "bad argument");
  // 1. the whole input (its size decides the batch size; the reference's reading thread streams it instead, readahead.rs:88-118)
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
  // 2. one batch = the whole file
  bamgpu_opts o{};
  if (opts) o = *opts; else o.device = -1;
  o.batch_bytes = len + 65536;
  o.flags &= ~(uint32_t)(BAMGPU_COPY_DATA | BAMGPU_COPY_COLUMNS | BAMGPU_NO_RECORDS | BAMGPU_SPECULATIVE_START);
  if (sort_by == BAMGPU_SORT_NAME_AND_MATCH_MATES) o.flags |= BAMGPU_WANT_HI;
  reader_thread_num = std::max<size_t>(reader_thread_num, 1);   // sort.rs:150
  bamgpu_reader* R = bamgpu_reader_from_memory(buf, len, reader_thread_num, &o);
  if (!R) { free(buf); return fail(BAMGPU_ERR_CUDA, "no usable CUDA device (this library has no CPU fallback)"); }
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
    if (R->prep_valid || !R->no_more) { rc = BAMGPU_ERR_TOO_LARGE; msg = "input was split into several batches"; break; }
    bamgpu_engine* E = R->E;
    bamgpu_sorted srt;
    rc = bamgpu_engine_sort(E, sort_by, 0, nullptr, &srt);
    if (rc != BAMGPU_OK) { msg = E->err; break; }
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
    const uint64_t total = srt.bodies_len;
    uint64_t issued = 0, written = 0;
    int k = 0;
    size_t pending[2] = {0, 0};
    auto issue = [&](int slot) {
      size_t m = (size_t)std::min<uint64_t>(CH, total - issued);
      cudaMemcpyAsync(st[slot].p, srt.bodies + issued, m, cudaMemcpyDeviceToHost, E->stream);
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
  free(buf);
  if (rc != BAMGPU_OK) return fail(rc, msg.empty() ? bamgpu_status_str(rc) : msg);
  return BAMGPU_OK;
}
