// sort.cu — GPU sort of decoded records + gather of their bodies: reproduces the OUTPUT BYTES of the reference's
// external merge sort (BASELINE.json configs[3], configs[4]).
//
// Replaces bam_tools/src/sorting/sort.rs:138-178 (sort_bam): read_split_sort_dump_chunks (:185-308; per chunk
// extract_key + rayon par_sort_by, stable, :351-371) and merge_sorted_chunks_and_write (:590-652; N-way heap merge, ties
// to the lower chunk index :487-510) — i.e. ONE globally stable sort of all records by the comparator of the chosen
// SortBy (sorting/comparators.rs:61-147), whose output is the concatenation of the record BODIES (no block_size, no
// header, no BGZF; sort.rs:643).  Everything is resident in HBM, so there are no chunks, no temp files and no merge.
//
//   CoordinatesAndStrand (comparators.rs:117-147): key = (refID == -1 ? i32::MAX : refID, pos, reverse strand).
//     K3 already wrote coord_key (both i32 biased to unsigned, refID in the high word) and strand; the order is a
//     stable LSD radix sort: 1 bit of strand, then the 64 bits of coord_key.
//   Name (comparators.rs:61-65): str::cmp of the read names INCLUDING their NUL (bamrawrecord.rs:118-123) = bytewise
//     lexicographic, a proper prefix first.  Variable-length keys: stable LSD passes from the last bytes to the first over
//     the names zero-padded past l_read_name - packed so that only the bits that differ between names are sorted
//     (k_name_masks / k_name_group_key).  Zero padding alone would tie "A\0" with "A\0\0"; l_read_name as the LEAST
//     significant key (stable sorts keep it as the order among names that tie on every byte) makes the result exact for
//     arbitrary name bytes.
//   NameAndMatchMates (comparators.rs:78-93): name, then HI (Option<i32>; None == None; None vs Some panics in the
//     reference -> reported in SortResult.hi_mixed), then the full u16 flag: the same, with flag and HI sorted first.
// The radix sort primitive is CUB's DeviceRadixSort (stable, onesweep); key construction, group refinement and the
// body gather — the HBM-heavy part — are the kernels below.
#include <cub/cub.cuh>

#include <utility>
#include <vector>

#include "kernels.h"

namespace bamgpu {
namespace {

__global__ void k_iota(uint32_t* v, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}

template <typename T, typename K>
__global__ void k_gather_key(const T* __restrict__ col, const uint32_t* __restrict__ idx, K* __restrict__ key, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) key[i] = (K)col[idx[i]];
}

// HI as an unsigned key: Some(v) -> v biased; None -> 0 (a name group mixing None and Some is an error anyway).
__global__ void k_hi_key(const uint8_t* __restrict__ present, const int32_t* __restrict__ value, const uint32_t* __restrict__ idx,
                         uint32_t* __restrict__ key, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    uint32_t r = idx[i];
    key[i] = present[r] == 1 ? ((uint32_t)value[r] ^ 0x80000000u) : 0u;
  }
}

// ---- name keys without their constant bits ----------------------------------------------------------------------
// Byte j of a name, zero past l_read_name ("virtual byte": what the zero-padded comparison sees).
__device__ __forceinline__ uint32_t name_vbyte(const uint8_t* __restrict__ nm, uint32_t len, uint32_t j) { return j < len ? nm[j] : 0u; }

// masks: byte j = OR over all records of (virtual byte j XOR record 0's virtual byte j), j < max_len: the bits of position j that
// are not the same in every name.  Real read names are mostly digits and fixed separators behind a common prefix: 3 - 4 varying
// bits per byte, none at all in many positions (the whole instrument : run : flowcell prefix).  One thread per record and tile of
// 64 positions (16 packed words), OR-reduced per warp; names up to 64 bytes take ONE pass over the records, and no separate
// common-prefix scan is needed.
constexpr uint32_t NAME_TILE = 64;
__global__ void k_name_masks(const uint8_t* __restrict__ data, const uint64_t* __restrict__ rec_off, const uint8_t* __restrict__ l_name,
                             uint32_t n, uint32_t* __restrict__ masks /* 4 positions per word */) {
  const uint32_t tile = blockIdx.y, j0 = tile * NAME_TILE;
  const uint8_t* ref = data + rec_off[0] + 32;
  const uint32_t lref = l_name[0];
  uint32_t rw[NAME_TILE / 4], m[NAME_TILE / 4];
#pragma unroll
  for (uint32_t w = 0; w < NAME_TILE / 4; ++w) {
    rw[w] = 0; m[w] = 0;
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k) rw[w] |= name_vbyte(ref, lref, j0 + 4 * w + k) << (8 * k);
  }
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint8_t* nm = data + rec_off[i] + 32;
    const uint32_t len = l_name[i];
#pragma unroll
    for (uint32_t w = 0; w < NAME_TILE / 4; ++w) {
      uint32_t v = 0;
      if (j0 + 4 * w < len) {
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) v |= name_vbyte(nm, len, j0 + 4 * w + k) << (8 * k);
      }
      m[w] |= v ^ rw[w];
    }
  }
#pragma unroll
  for (uint32_t w = 0; w < NAME_TILE / 4; ++w) {
    const uint32_t r = __reduce_or_sync(0xffffffffu, m[w]);
    if ((threadIdx.x & 31) == 0 && r) atomicOr(&masks[j0 / 4 + w], r);
  }
}

// One LSD pass sorts by a GROUP of positions: the low width[e] bits of the virtual bytes at pos[e], most significant position
// first, optionally l_read_name in the low 8 bits (the last tie-break of the zero-padded comparison: "A\0" before "A\0\0").
struct NameGroup {
  uint8_t pos[64], width[64];
  uint32_t count, with_len;
};
template <typename K>
__global__ void k_name_group_key(const uint8_t* __restrict__ data, const uint64_t* __restrict__ rec_off, const uint8_t* __restrict__ l_name,
                                 const uint32_t* __restrict__ idx, NameGroup g, K* __restrict__ key, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t rec = idx[i];
  const uint8_t* nm = data + rec_off[rec] + 32;
  const uint32_t len = l_name[rec];
  uint64_t k = 0;
  for (uint32_t e = 0; e < g.count; ++e) {
    const uint32_t w = g.width[e];
    k = (k << w) | (name_vbyte(nm, len, g.pos[e]) & ((1u << w) - 1u));
  }
  if (g.with_len) k = (k << 8) | len;
  key[i] = (K)k;
}

__global__ void k_max_u8(const uint8_t* __restrict__ v, uint32_t n, uint32_t* out) {
  uint32_t m = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, (uint32_t)v[i]);
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// NameAndMatchMates: neighbours in the sorted order whose names are equal byte for byte must agree on HI being present
// (comparators.rs:90 unwrap).
__global__ void k_check_hi_names(const uint8_t* __restrict__ data, const uint64_t* __restrict__ rec_off, const uint32_t* __restrict__ idx,
                                 const uint8_t* __restrict__ l_name, const uint8_t* __restrict__ present, uint32_t n, SortResult* res) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const uint32_t b = idx[i];
    if (present[b] == 2) atomicAdd(&res->hi_bad_tags, 1ull);
    if (i) {
      const uint32_t a = idx[i - 1];
      if (l_name[a] == l_name[b] && (present[a] == 1) != (present[b] == 1)) {
        const uint8_t* pa = data + rec_off[a] + 32;
        const uint8_t* pb = data + rec_off[b] + 32;
        const uint32_t len = l_name[a];
        uint32_t k = 0;
        while (k < len && pa[k] == pb[k]) ++k;
        if (k == len) atomicAdd(&res->hi_mixed, 1ull);
      }
    }
  }
}

__global__ void k_perm_len(const uint32_t* __restrict__ rec_len, const uint32_t* __restrict__ perm, uint64_t* __restrict__ len64, uint32_t n, uint32_t prefix) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) len64[i] = (uint64_t)rec_len[perm[i]] + prefix;
}

// One warp per record: body bytes data[rec_off .. +rec_len) -> out[out_off ..).  Destination-aligned 16-byte stores;
// each is assembled from the two aligned source vectors that cover it (the second is the next iteration's first, so
// every source byte is loaded once per warp).  HBM-bound: reads and writes every body byte once.
__global__ void __launch_bounds__(256) k_gather_bodies(const uint8_t* __restrict__ data, const uint64_t* __restrict__ rec_off,
                                                       const uint32_t* __restrict__ rec_len, const uint32_t* __restrict__ perm,
                                                       const uint64_t* __restrict__ out_off, uint8_t* __restrict__ out, uint32_t n,
                                                       uint32_t prefix) {
  // prefix = 4: the record's block_size field (the 4 bytes in front of the body in the inflated stream) is copied with it:
  // (u32 block_size, body)*, the reference's spill wire format (sort.rs:340-347) and a BAM record stream
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const uint32_t rec = perm[i];
    const uint8_t* src = data + rec_off[rec] - prefix;
    uint8_t* dst = out + out_off[i];
    const uint32_t len = rec_len[rec] + prefix;
    // head: up to 15 bytes until dst is 16-byte aligned
    uint32_t headn = (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15);
    if (headn > len) headn = len;
    if (lane < headn) dst[lane] = src[lane];
    const uint32_t body = (len - headn) & ~15u;
    const uint8_t* s = src + headn;
    uint8_t* d = dst + headn;
    const uint32_t sh = (uint32_t)((uintptr_t)s & 3) * 8;
    const uint32_t* sw = (const uint32_t*)((uintptr_t)s & ~(uintptr_t)3);   // readable: the stream has pads on both sides
    for (uint32_t o = lane * 16; o < body; o += 32 * 16) {
      const uint32_t* w = sw + (o >> 2);
      uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = sh ? w[4] : 0u;
      uint4 v;
      v.x = __funnelshift_r(w0, w1, sh);
      v.y = __funnelshift_r(w1, w2, sh);
      v.z = __funnelshift_r(w2, w3, sh);
      v.w = __funnelshift_r(w3, w4, sh);
      *(uint4*)(d + o) = v;
    }
    const uint32_t tail = len - headn - body;   // < 16
    if (lane < tail) d[body + lane] = s[body + lane];
  }
}

// ---- streamed sort (inputs whose bodies do not fit in HBM): only the keys stay on the device ----
__global__ void k_name_len64(const uint8_t* __restrict__ l_name, uint32_t n, uint64_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n) out[i] = i < n ? l_name[i] : 0;
}
// The names of a batch, back to back in dst (which becomes store[store_base ..) — on this or, after a peer copy, on another
// device); voff[i] = the "record offset" that makes store + voff + 32 the name.
__global__ void k_pack_names(const uint8_t* __restrict__ data, const uint64_t* __restrict__ rec_off, const uint8_t* __restrict__ l_name,
                             const uint64_t* __restrict__ local_off, uint32_t n, uint8_t* __restrict__ dst, uint64_t store_base,
                             uint64_t* __restrict__ voff) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const uint64_t o = local_off[i];
    const uint8_t* src = data + rec_off[i] + 32;
    const uint32_t len = l_name[i];
    for (uint32_t k = 0; k < len; ++k) dst[o + k] = src[k];
    voff[i] = store_base + o - 32;
  }
}


// ---- sample sort over several GPUs (bamgpu_sort_bam with a device list): splitters and the stable partition by them ----
// a <= b in the order the buckets are cut by: (coord_key, strand), or the read names bytewise (a proper prefix first).
// Records that compare equal here land in ONE bucket, so HI / flag / input order are settled by the bucket's own stable sort.
__device__ __forceinline__ bool split_le_name(const uint8_t* __restrict__ a, uint32_t la, const uint8_t* __restrict__ b, uint32_t lb) {
  const uint32_t l = la < lb ? la : lb;
  for (uint32_t k = 0; k < l; ++k) {
    const uint32_t x = a[k], y = b[k];
    if (x != y) return x < y;
  }
  return la <= lb;
}
__global__ void k_sample_keys(const uint8_t* __restrict__ data, const uint64_t* __restrict__ rec_off, const uint8_t* __restrict__ l_name,
                              const uint64_t* __restrict__ coord_key, const uint8_t* __restrict__ strand, uint32_t n, uint32_t n_samples,
                              int by_name, SplitKey* __restrict__ out) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_samples) return;
  const uint32_t i = (uint32_t)(((uint64_t)j * n + n / 2) / n_samples);   // evenly spaced over the (unsorted) input
  SplitKey& o = out[j];
  o.key = by_name ? 0 : coord_key[i];
  o.strand = by_name ? 0 : strand[i];
  o.name_len = by_name ? l_name[i] : 0;
  if (by_name) {
    const uint8_t* src = data + rec_off[i] + 32;
    for (uint32_t k = 0; k < o.name_len; ++k) o.name[k] = src[k];
  }
}
__global__ void k_bucket_of(const uint8_t* __restrict__ data, const uint64_t* __restrict__ rec_off, const uint8_t* __restrict__ l_name,
                            const uint64_t* __restrict__ coord_key, const uint8_t* __restrict__ strand, uint32_t n, int by_name,
                            const SplitKey* __restrict__ sp, uint32_t n_split, uint32_t* __restrict__ bucket) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t b = 0;
  if (by_name) {
    const uint8_t* nm = data + rec_off[i] + 32;
    const uint32_t ln = l_name[i];
    for (uint32_t d = 0; d < n_split; ++d) b += split_le_name(sp[d].name, sp[d].name_len, nm, ln) ? 1u : 0u;
  } else {
    const uint64_t k = coord_key[i];
    const uint32_t st = strand[i];
    for (uint32_t d = 0; d < n_split; ++d) b += (sp[d].key < k || (sp[d].key == k && sp[d].strand <= st)) ? 1u : 0u;
  }
  bucket[i] = b;   // splitters are sorted, so this is the index of the bucket
}
// bounds[b] = first position of the partitioned order whose bucket is >= b, for b = 0 .. n_buckets (bounds[n_buckets] = n)
__global__ void k_bucket_bounds(const uint32_t* __restrict__ sorted_bucket, uint32_t n, uint32_t n_buckets, uint32_t* __restrict__ bounds) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  const uint32_t prev = i ? sorted_bucket[i - 1] : 0u, cur = i < n ? sorted_bucket[i] : n_buckets;
  if (i == 0) bounds[0] = 0;
  for (uint32_t b = prev + 1; b <= cur; ++b) bounds[b] = i;
}

inline uint32_t blocks_for(uint32_t n, uint32_t threads = 256) { return (n + threads - 1) / threads; }
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct Carve {
  uint8_t* p;
  size_t used = 0;
  template <class T> T* take(size_t count) {
    T* r = (T*)(p + used);
    used = align256(used + count * sizeof(T));
    return r;
  }
};

size_t cub_temp_bytes(uint32_t n) {
  size_t a = 0, b = 0, c = 0, d = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 64);
  cub::DeviceRadixSort::SortPairs(nullptr, b, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 32);
  cub::DeviceScan::InclusiveSum(nullptr, c, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
  cub::DeviceScan::ExclusiveSum(nullptr, d, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int)n + 1);
  size_t m = a > b ? a : b;
  m = m > c ? m : c;
  m = m > d ? m : d;
  return align256(m + 256);
}

}  // namespace

size_t sort_scratch_bytes(uint64_t n) {
  const uint32_t m = (uint32_t)n + 1;
  // key64 A/B, val A/B, key32 A/B, seg_elem, flags/scan, cub temp
  return align256(8ull * m) * 2 + align256(4ull * m) * 6 + cub_temp_bytes((uint32_t)n) + 4096;
}

// Writes perm[0..n): perm[i] = index of the i-th record of the sorted order.  `scratch` >= sort_scratch_bytes(n).
cudaError_t sort_permutation(const uint8_t* data, const ColumnsDev& c, uint64_t n64, int sort_by, void* scratch, uint32_t* perm,
                             SortResult* d_res, int* launches, cudaStream_t s) {
  const uint32_t n = (uint32_t)n64;
  cudaMemsetAsync(d_res, 0, sizeof(SortResult), s);
  if (n == 0) return cudaGetLastError();
  Carve cv{(uint8_t*)scratch};
  uint64_t* k64a = cv.take<uint64_t>(n + 1);
  uint64_t* k64b = cv.take<uint64_t>(n + 1);
  uint32_t* va = cv.take<uint32_t>(n + 1);
  uint32_t* vb = cv.take<uint32_t>(n + 1);
  uint32_t* k32a = cv.take<uint32_t>(n + 1);
  uint32_t* k32b = cv.take<uint32_t>(n + 1);
  uint32_t* seg_elem = cv.take<uint32_t>(n + 1);
  uint32_t* flags = cv.take<uint32_t>(n + 1);
  void* tmp = cv.p + cv.used;
  size_t tmp_bytes = cub_temp_bytes(n);
  const uint32_t g = blocks_for(n);
  int nl = 0;
  auto sort32 = [&](int bits) {   // (k32a, va) -> (k32b, vb), then swap names
    cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const uint32_t*)k32a, k32b, (const uint32_t*)va, vb, (int)n, 0, bits, s);
    std::swap(va, vb);
    std::swap(k32a, k32b);
    nl += 3;
  };
  auto sort64 = [&](int bits = 64) {
    cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const uint64_t*)k64a, k64b, (const uint32_t*)va, vb, (int)n, 0, bits, s);
    std::swap(va, vb);
    std::swap(k64a, k64b);
    nl += 5;
  };
  k_iota<<<g, 256, 0, s>>>(va, n);
  ++nl;
  if (sort_by == SORT_COORDINATES_AND_STRAND) {
    k_gather_key<uint8_t, uint32_t><<<g, 256, 0, s>>>(c.strand, va, k32a, n);
    sort32(1);
    k_gather_key<uint64_t, uint64_t><<<g, 256, 0, s>>>(c.coord_key, va, k64a, n);
    sort64();
    nl += 2;
  } else {
    if (sort_by == SORT_NAME_AND_MATCH_MATES) {
      k_gather_key<uint16_t, uint32_t><<<g, 256, 0, s>>>(c.flag, va, k32a, n);
      sort32(16);
      k_hi_key<<<g, 256, 0, s>>>(c.hi_present, c.hi_value, va, k32a, n);
      sort32(32);
      nl += 2;
    }
    // The comparison is bytewise over the names zero-padded to the longest, then by l_read_name.  Stable LSD passes, least
    // significant first - but only over the bits that differ between names: one pass over all names ORs, per byte position, the
    // XOR against the first name (k_name_masks); positions that are the same everywhere (the instrument : run : flowcell prefix,
    // separators) drop out, a digit position keeps 4 bits.  The remaining bits are packed, most significant position first,
    // into groups of <= 64 bits; every group is one key gather + one radix sort over just its bits, and l_read_name rides in the
    // low 8 bits of the least significant group.  50 M Illumina-style names: ~80 varying bits = 2 groups, 11 radix passes and 2
    // gathers, where whole 8-byte chunks took 3 groups, 25 passes and 3 gathers (20.0 ms).
    uint32_t* d_max = (uint32_t*)&d_res->max_name_len;
    k_max_u8<<<592, 256, 0, s>>>(c.l_read_name, n, d_max);
    uint32_t* d_masks = (uint32_t*)((uint8_t*)scratch + sort_scratch_bytes(n64) - 2048);   // the slack at the end of the scratch: 256 positions, 4 per word
    cudaMemsetAsync(d_masks, 0, 256, s);
    ++nl;
    uint32_t max_len = 0;
    cudaMemcpyAsync(&max_len, d_max, 4, cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return e;
    uint8_t masks[256] = {0};
    const uint32_t lcp = 0;
    if (max_len) {
      const uint32_t tiles = (max_len + NAME_TILE - 1) / NAME_TILE;
      uint32_t gx = (n + 255) / 256;
      const uint32_t cap = 592u * 4 / tiles + 1;
      if (gx > cap) gx = cap;
      k_name_masks<<<dim3(gx, tiles), 256, 0, s>>>(data, c.rec_off, c.l_read_name, n, d_masks);
      ++nl;
      cudaMemcpyAsync(masks, d_masks, 256, cudaMemcpyDeviceToHost, s);   // little-endian host: byte j of the array = position j
      e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) return e;
    }
    // groups, least significant (highest positions) first
    std::vector<NameGroup> groups;
    {
      NameGroup cur{};
      uint32_t bits = 8;                 // the first group carries l_read_name
      cur.with_len = 1;
      std::vector<std::pair<uint8_t, uint8_t>> rev;   // (pos, width) of the current group, highest position first
      auto close = [&]() {
        cur.count = (uint32_t)rev.size();
        for (uint32_t k2 = 0; k2 < cur.count; ++k2) { cur.pos[k2] = rev[cur.count - 1 - k2].first; cur.width[k2] = rev[cur.count - 1 - k2].second; }
        groups.push_back(cur);
        cur = NameGroup{};
        rev.clear();
        bits = 0;
      };
      for (uint32_t j = max_len; j-- > lcp;) {
        const uint32_t m = masks[j];
        if (!m) continue;
        const uint32_t w = 32u - (uint32_t)__builtin_clz(m);
        if (bits + w > 64 || rev.size() == 64) close();
        rev.push_back({(uint8_t)j, (uint8_t)w});
        bits += w;
      }
      close();                           // (also when no position varies: l_read_name alone)
    }
    for (const NameGroup& grp : groups) {
      uint32_t bits = grp.with_len ? 8u : 0u;
      for (uint32_t k2 = 0; k2 < grp.count; ++k2) bits += grp.width[k2];
      if (bits == 0) continue;
      if (bits <= 32) {
        k_name_group_key<uint32_t><<<g, 256, 0, s>>>(data, c.rec_off, c.l_read_name, va, grp, k32a, n);
        sort32((int)bits);
      } else {
        k_name_group_key<uint64_t><<<g, 256, 0, s>>>(data, c.rec_off, c.l_read_name, va, grp, k64a, n);
        sort64((int)bits);
      }
      ++nl;
    }
    if (sort_by == SORT_NAME_AND_MATCH_MATES) {
      k_check_hi_names<<<g, 256, 0, s>>>(data, c.rec_off, va, c.l_read_name, c.hi_present, n, d_res);
      ++nl;
    }
  }
  cudaMemcpyAsync(perm, va, 4ull * n, cudaMemcpyDeviceToDevice, s);
  if (launches) *launches += nl;
  return cudaGetLastError();
}

// Streamed sort: appends the read names of `n` records (a decoded batch) to the compact name store and writes the virtual
// record offsets sort_permutation needs to find them there.  `scratch` >= sort_scratch_bytes(n).  *name_bytes (host) gets
// the bytes appended; synchronises.  The caller makes sure the store holds store_base + 255 * n more bytes... or calls
// names_total first (cheap) — here the store must hold store_base + *name_bytes, checked by the caller through names_total.
cudaError_t names_total(const ColumnsDev& c, uint64_t n64, void* scratch, uint64_t* name_bytes, cudaStream_t s) {
  const uint32_t n = (uint32_t)n64;
  *name_bytes = 0;
  if (n == 0) return cudaSuccess;
  Carve cv{(uint8_t*)scratch};
  uint64_t* len64 = cv.take<uint64_t>(n + 1);
  cv.take<uint64_t>(n + 1);
  for (int k = 0; k < 6; ++k) cv.take<uint32_t>(n + 1);
  void* tmp = cv.p + cv.used;
  size_t tmp_bytes = cub_temp_bytes(n);
  k_name_len64<<<blocks_for(n + 1), 256, 0, s>>>(c.l_read_name, n, len64);
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, (const uint64_t*)len64, len64, (int)n + 1, s);
  cudaMemcpyAsync(name_bytes, len64 + n, 8, cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  return e != cudaSuccess ? e : cudaGetLastError();
}
// After names_total on the same scratch: copies the names to dst[0 .. name_bytes) and writes voff[0..n) for a store in which
// they will sit at store_base.
cudaError_t pack_names(const uint8_t* data, const ColumnsDev& c, uint64_t n64, void* scratch, uint8_t* dst, uint64_t store_base,
                       uint64_t* voff, cudaStream_t s) {
  const uint32_t n = (uint32_t)n64;
  if (n == 0) return cudaSuccess;
  k_pack_names<<<blocks_for(n), 256, 0, s>>>(data, c.rec_off, c.l_read_name, (const uint64_t*)scratch, n, dst, store_base, voff);
  return cudaGetLastError();
}


// Sample sort over several GPUs, step 1: n_samples keys of this shard, evenly spaced, to `d_out` (device).
cudaError_t sample_keys(const uint8_t* data, const ColumnsDev& c, uint64_t n64, int sort_by, uint32_t n_samples, SplitKey* d_out, cudaStream_t s) {
  const uint32_t n = (uint32_t)n64;
  if (n == 0 || n_samples == 0) return cudaSuccess;
  k_sample_keys<<<blocks_for(n_samples), 256, 0, s>>>(data, c.rec_off, c.l_read_name, c.coord_key, c.strand, n, n_samples,
                                                     sort_by != SORT_COORDINATES_AND_STRAND, d_out);
  return cudaGetLastError();
}
// Step 2: perm[0..n) = the shard's records grouped by bucket (bucket b = records with splitter[b-1] <= key < splitter[b]), input
// order kept inside a bucket; d_bounds[0..n_split+1] = where each bucket starts in perm.  `scratch` >= sort_scratch_bytes(n).
cudaError_t partition_by_splitters(const uint8_t* data, const ColumnsDev& c, uint64_t n64, int sort_by, const SplitKey* d_split, uint32_t n_split,
                                   void* scratch, uint32_t* perm, uint32_t* d_bounds, int* launches, cudaStream_t s) {
  const uint32_t n = (uint32_t)n64, n_buckets = n_split + 1;
  if (n == 0) return cudaMemsetAsync(d_bounds, 0, 4 * (n_buckets + 1), s);
  Carve cv{(uint8_t*)scratch};
  cv.take<uint64_t>(n + 1);
  cv.take<uint64_t>(n + 1);
  uint32_t* va = cv.take<uint32_t>(n + 1);
  uint32_t* vb = cv.take<uint32_t>(n + 1);
  uint32_t* k32a = cv.take<uint32_t>(n + 1);
  uint32_t* k32b = cv.take<uint32_t>(n + 1);
  cv.take<uint32_t>(n + 1);
  cv.take<uint32_t>(n + 1);
  void* tmp = cv.p + cv.used;
  size_t tmp_bytes = cub_temp_bytes(n);
  const uint32_t g = blocks_for(n);
  k_iota<<<g, 256, 0, s>>>(va, n);
  k_bucket_of<<<g, 256, 0, s>>>(data, c.rec_off, c.l_read_name, c.coord_key, c.strand, n, sort_by != SORT_COORDINATES_AND_STRAND, d_split,
                               n_split, k32a);
  int bits = 1;
  while ((1u << bits) < n_buckets) ++bits;
  cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const uint32_t*)k32a, k32b, (const uint32_t*)va, vb, (int)n, 0, bits, s);
  k_bucket_bounds<<<blocks_for(n + 1), 256, 0, s>>>(k32b, n, n_buckets, d_bounds);
  cudaMemcpyAsync(perm, vb, 4ull * n, cudaMemcpyDeviceToDevice, s);
  if (launches) *launches += 6;
  return cudaGetLastError();
}

// out_off[0..n] = exclusive scan of the body lengths in sorted order; bodies copied to out back to back.
// `scratch` as above (only the front is used).  *total_bytes is read back (synchronises).
cudaError_t gather_sorted_bodies(const uint8_t* data, const ColumnsDev& c, const uint32_t* perm, uint64_t n64, void* scratch,
                                 uint64_t* out_off, uint8_t* out, uint32_t prefix, cudaStream_t s) {
  const uint32_t n = (uint32_t)n64;
  if (n == 0) return cudaSuccess;
  k_gather_bodies<<<148 * 8, 256, 0, s>>>(data, c.rec_off, c.rec_len, perm, out_off, out, n, prefix);
  return cudaGetLastError();
}

cudaError_t sorted_offsets(const ColumnsDev& c, const uint32_t* perm, uint64_t n64, void* scratch, uint64_t* out_off, uint64_t* total,
                           uint32_t prefix, cudaStream_t s) {
  const uint32_t n = (uint32_t)n64;
  *total = 0;
  if (n == 0) return cudaMemsetAsync(out_off, 0, 8, s);
  Carve cv{(uint8_t*)scratch};
  uint64_t* len64 = cv.take<uint64_t>(n + 1);
  cv.take<uint64_t>(n + 1);
  for (int k = 0; k < 6; ++k) cv.take<uint32_t>(n + 1);
  void* tmp = cv.p + cv.used;
  size_t tmp_bytes = cub_temp_bytes(n);
  k_perm_len<<<blocks_for(n), 256, 0, s>>>(c.rec_len, perm, len64, n, prefix);
  cudaMemsetAsync(len64 + n, 0, 8, s);
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, (const uint64_t*)len64, out_off, (int)n + 1, s);
  cudaMemcpyAsync(total, out_off + n, 8, cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace bamgpu
