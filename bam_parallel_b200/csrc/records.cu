// records.cu — K3: BAM record boundaries, fixed fields, variable-field offsets and sort keys.
//
// Replaces the single-threaded consumer loop of the reference:
//   Reader::read_block_size / append_record        bam_tools/src/reader.rs:57-81
//   Records::next_rec                              reader/records.rs:23-29
// and materialises what the sort later pulls lazily out of every record:
//   fixed-field offsets                            record/bamrawrecord.rs:86-97
//   variable-field offsets (usize form)            record/bamrawrecord.rs:61-77
//   KeyTuple::CoordinatesAndStrand / Name / HI     sorting/comparators.rs:41-59, record/tags.rs:113-129
//
// The chain  next = p + 4 + u32[p]  is serial by definition and the reference walks it blindly
// (no validation).  We get the identical chain in parallel:
//   1. k3_tile_walk: the inflated stream is cut into 64 KiB tiles; one thread per tile guesses the
//      first record start in its tile with a plausibility filter (spec invariants only, so a real
//      record start of a valid BAM always passes; a candidate must be followed by three more
//      look-alike records along its block_size links) and walks the chain to the tile end
//      -> (entry, exit, count).  The start tile begins at the KNOWN start (end of the header).
//   2. k3_resolve (one CTA): tile k points at the tile holding its exit iff that tile's guess equals
//      the exit; pointer DOUBLING from the start tile marks the tiles on the true chain in log2(n)
//      rounds.  A guess that does not match (or a tile a long record jumped into) is never trusted:
//      the chain stops there, thread 0 re-walks serially from the true entry until it
//      re-synchronises, and the resolve repeats.  The result is therefore exactly the blind walk,
//      never the filter's opinion.  An exclusive scan then gives every tile its first record index.
//   3. k3_tile_emit re-walks and writes rec_off / rec_len, k3_extract (one thread per record)
//      writes the SoA columns + keys.
#include <cooperative_groups.h>

#include "kernels.h"
#include "tags.cuh"

namespace cg = cooperative_groups;

namespace bamgpu {
namespace {

constexpr uint64_t NONE64 = ~0ull;

__device__ __forceinline__ uint32_t ld32(const uint8_t* data, uint64_t off) {
  const uint32_t* a = (const uint32_t*)(data + (off & ~3ull));
  uint32_t sh = (uint32_t)(off & 3) * 8;
  uint32_t lo = a[0];
  return sh ? __funnelshift_r(lo, a[1], sh) : lo;
}

// Spec invariants of a record start (SAM spec §1.4, §4.2).  Used ONLY to guess; never to decide.
// `bs` is the already loaded block_size at q.
__device__ __forceinline__ bool plausible_fields(const RecordParams& p, uint64_t q, uint32_t bs) {
  if (p.data_len - q < 36) return false;
  if (bs < 33u || bs > (1u << 28)) return false;
  int32_t ref = (int32_t)ld32(p.data, q + 4);
  if (ref < -1 || (p.n_ref >= 0 && ref >= p.n_ref)) return false;
  if ((int32_t)ld32(p.data, q + 8) < -1) return false;
  uint32_t w = ld32(p.data, q + 12);
  uint32_t l_name = w & 0xff;
  if (l_name == 0) return false;
  uint32_t n_cig = ld32(p.data, q + 16) & 0xffff;
  uint32_t l_seq = ld32(p.data, q + 20);
  if (l_seq > bs) return false;
  int32_t nref = (int32_t)ld32(p.data, q + 24);
  if (nref < -1 || (p.n_ref >= 0 && nref >= p.n_ref)) return false;
  if ((int32_t)ld32(p.data, q + 28) < -1) return false;
  uint64_t need = 32ull + l_name + 4ull * n_cig + (uint64_t)((l_seq + 1u) >> 1) + l_seq;
  if (need > bs) return false;
  uint64_t nul = q + 4 + 32 + l_name - 1;
  if (nul < p.data_len && p.data[nul] != 0) return false;
  if (l_name > 1) {  // read names are [!-?A-~]{1,254}
    uint32_t c0 = p.data[q + 36];
    if (c0 < 0x21 || c0 > 0x7e) return false;
  }
  return true;
}

// A candidate is accepted when it and the next VERIFY_DEPTH records along its chain all look like
// records, or the chain ends at (or just past) the end of the data first.  One look-alike inside
// sequence or quality bytes is common; four in a row along block_size links is not.
constexpr int VERIFY_DEPTH = 3;
__device__ bool plausible_chain(const RecordParams& p, uint64_t q, uint32_t bs) {
  if (!plausible_fields(p, q, bs)) return false;
  for (int i = 0; i < VERIFY_DEPTH; ++i) {
    q += 4ull + bs;
    if (q > p.data_len) return q - p.data_len <= WILD_JUMP_BYTES;  // last record of a batch may be cut; a wild jump is not
    if (p.data_len - q < 36) return true;                     // tail: nothing left to check
    bs = ld32(p.data, q);
    if (!plausible_fields(p, q, bs)) return false;
  }
  return true;
}

// First accepted candidate in [from, te), or NONE64.  Bytes are scanned through a rolling 64-bit
// window (one aligned load per 4 candidates); the cheap block_size range test rejects almost all.
__device__ uint64_t find_entry(const RecordParams& p, uint64_t from, uint64_t te) {
  if (from >= te) return NONE64;
  uint64_t a = from & ~3ull;
  const uint32_t* w = (const uint32_t*)(p.data + a);
  uint32_t lo = w[0];
  for (; a < te; a += 4) {
    uint32_t hi = *(const uint32_t*)(p.data + a + 4);  // data is readable 64 bytes past data_len
#pragma unroll
    for (uint32_t b = 0; b < 4; ++b) {
      uint64_t q = a + b;
      uint32_t bs = b ? __funnelshift_r(lo, hi, 8 * b) : lo;
      if (bs - 33u <= (1u << 28) - 33u && q >= from && q < te && plausible_chain(p, q, bs)) return q;
    }
    lo = hi;
  }
  return NONE64;
}

struct WalkOut { uint64_t exit_off; uint32_t count; uint32_t kind; };

// reader.rs:57-81 restricted to records STARTING before tile_end.  When EMIT, writes offsets.
// `ck` (nullable): every CK_STRIDE-th record start of the tile is noted there, relative to `ck_base` (the tile's first
// byte; a record counted here starts inside the tile, so 16 bits suffice).  k3_tile_emit restarts from these notes with
// one lane per note instead of re-walking the tile's ~200 records as one dependent chain.
template <bool EMIT>
__device__ WalkOut walk(const RecordParams& p, uint64_t q, uint64_t tile_end, uint64_t* rec_off, uint32_t* rec_len,
                        uint64_t base, uint64_t cap = ~0ull, uint16_t* ck = nullptr, uint64_t ck_base = 0) {
  WalkOut o;
  uint32_t cnt = 0;
  for (;;) {
    if (p.data_len - q < 4) { o.kind = EXIT_TAIL; break; }           // :78 UnexpectedEof => 0
    if (q >= tile_end) { o.kind = q >= p.own_len ? EXIT_LIMIT : EXIT_NORMAL; break; }
    uint32_t bs = ld32(p.data, q);
    if (bs == 0) { o.kind = EXIT_ZERO; break; }                       // :64-72, records.rs:25
    if (p.data_len - (q + 4) < bs) { o.kind = EXIT_INCOMPLETE; break; }  // :69-70 would panic if final
    if (EMIT && base + cnt < cap) { rec_off[base + cnt] = q + 4; rec_len[base + cnt] = bs; }
    if (ck && (cnt % CK_STRIDE) == 0) ck[cnt / CK_STRIDE] = (uint16_t)(q - ck_base);
    ++cnt;
    q += 4ull + bs;
  }
  o.exit_off = q;
  o.count = cnt;
  return o;
}

__global__ void __launch_bounds__(128) k3_tile_walk(RecordParams p, TileArrays t) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= t.n_tiles) return;
  const uint64_t tb = (uint64_t)k * TILE_BYTES;
  uint64_t te = tb + TILE_BYTES;
  if (te > p.own_len) te = p.own_len;
  const uint32_t k0 = (uint32_t)(p.start / TILE_BYTES);
  uint64_t e = NONE64;
  if (k == k0) e = p.start_known ? p.start : find_entry(p, p.start, te);
  else if (k > k0) e = find_entry(p, tb, te);
  t.entry[k] = e;
  if (e == NONE64) { t.exit_off[k] = 0; t.count[k] = 0; t.exit_kind[k] = EXIT_NONE; return; }
  WalkOut o = walk<false>(p, e, te, nullptr, nullptr, 0, ~0ull, t.ckpt + (size_t)k * CK_PER_TILE, tb);
  t.exit_off[k] = o.exit_off;
  t.count[k] = o.count;
  t.exit_kind[k] = o.kind;
}

// One cooperative launch (one CTA per SM, grid-wide barriers) resolves the whole tile graph:
//   link:   tile k points at the tile holding its exit iff that tile's guess equals the exit;
//   reach:  pointer doubling from the start tile marks the tiles on the true chain;
//   repair: if the chain stops at a tile whose guess was wrong (or that a long record jumped into),
//           one thread re-walks serially from the true entry until it re-synchronises with a guess,
//           and the resolve repeats.  The result is therefore exactly the blind walk of reader.rs:63-81;
//   scan:   exclusive scan of the counts of reachable tiles = first record index of every tile.
// Arrays written by one CTA and read by another between barriers are read with ld.cg (L2, the
// coherence point); grid.sync() orders the phases.
constexpr int RESOLVE_THREADS = 512;

__device__ __forceinline__ uint32_t ldcg32(const uint32_t* p) { return __ldcg(p); }
__device__ __forceinline__ uint64_t ldcg64(const void* p) { return (uint64_t)__ldcg((const unsigned long long*)p); }
__device__ __forceinline__ uint8_t ldcg8(const uint8_t* p) { return __ldcg(p); }

__global__ void __launch_bounds__(RESOLVE_THREADS) k3_resolve(RecordParams p, TileArrays t, ChainResult* res,
                                                              ResolveScratch* g) {
  cg::grid_group grid = cg::this_grid();
  const uint32_t n = t.n_tiles;
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  if (tid == 0) { g->first = p.start_known ? (uint32_t)(p.start / TILE_BYTES) : 0xffffffffu; g->repairs = 0; }
  grid.sync();
  if (!p.start_known) {
    // shard-edge mode: the chain begins at the first tile (from the start tile on) that found a guess
    uint32_t best = 0xffffffffu;
    for (uint32_t k = (uint32_t)(p.start / TILE_BYTES) + tid; k < n; k += nthreads)
      if (t.entry[k] != NONE64) { best = k; break; }
    if (best != 0xffffffffu) atomicMin(&g->first, best);
    grid.sync();
  }
  const uint32_t k0 = ldcg32(&g->first);
  if (k0 >= n) {  // nothing that looks like a record anywhere
    for (uint32_t k = tid; k < n; k += nthreads) { t.reach[k] = 0; t.base[k] = 0; }
    if (tid == 0) {
      res->n_records = 0; res->tail_off = p.start; res->tail_kind = EXIT_NONE; res->broken_tile = 0xffffffffu;
      res->broken_entry = 0; res->bad_flags = 0; res->rounds = 0; res->chain_start = NONE64;
    }
    return;
  }
  for (;;) {
    if (tid == 0) { g->broken_tile = 0xffffffffu; g->tail_off = p.start; g->tail_kind = EXIT_TAIL; }
    for (uint32_t k = tid; k < n; k += nthreads) {
      uint32_t nx = k;
      uint64_t ex = ldcg64(&t.exit_off[k]);
      if (ldcg64(&t.entry[k]) != NONE64 && ldcg32(&t.exit_kind[k]) == EXIT_NORMAL) {
        uint32_t j = (uint32_t)(ex / TILE_BYTES);
        if (j < n && ldcg64(&t.entry[j]) == ex) nx = j;
      }
      t.nxt[k] = nx;
      t.jump[k] = nx;
      t.reach[k] = (k == k0) ? 1 : 0;
    }
    grid.sync();
    uint32_t* a = t.jump;
    uint32_t* b = t.jump2;
    for (uint64_t span = 1; span < (uint64_t)n * 2; span <<= 1) {
      for (uint32_t k = tid; k < n; k += nthreads) {
        uint32_t j = ldcg32(&a[k]);
        if (ldcg8(&t.reach[k])) t.reach[j] = 1;
        b[k] = ldcg32(&a[j]);
      }
      grid.sync();
      uint32_t* tmp = a; a = b; b = tmp;
    }
    // the chain's last tile: reachable and its own successor (exactly one such tile)
    for (uint32_t k = tid; k < n; k += nthreads) {
      if (ldcg8(&t.reach[k]) && t.nxt[k] == k) {
        uint32_t kind = ldcg32(&t.exit_kind[k]);
        uint64_t ex = ldcg64(&t.exit_off[k]);
        if (ldcg64(&t.entry[k]) == NONE64) {          // only the start tile can be reachable without an entry
          g->tail_off = p.start; g->tail_kind = EXIT_NONE;
        } else if (kind == EXIT_NORMAL) {            // successor's guess mismatched (or successor out of range)
          g->broken_entry = ex;
          g->broken_tile = (uint32_t)(ex / TILE_BYTES);
        } else {
          g->tail_off = ex;
          g->tail_kind = kind;
        }
      }
    }
    grid.sync();
    uint32_t broken = ldcg32(&g->broken_tile);
    if (broken == 0xffffffffu || broken >= n) break;
    if (tid == 0) {
      uint32_t tile = broken;
      uint64_t entry = ldcg64(&g->broken_entry);
      uint32_t reps = 0;
      for (;;) {
        uint64_t te = (uint64_t)(tile + 1) * TILE_BYTES;
        if (te > p.own_len) te = p.own_len;
        WalkOut o = walk<false>(p, entry, te, nullptr, nullptr, 0, ~0ull, t.ckpt + (size_t)tile * CK_PER_TILE, (uint64_t)tile * TILE_BYTES);
        t.entry[tile] = entry;
        t.exit_off[tile] = o.exit_off;
        t.count[tile] = o.count;
        t.exit_kind[tile] = o.kind;
        ++reps;
        if (o.kind != EXIT_NORMAL) break;
        uint32_t j = (uint32_t)(o.exit_off / TILE_BYTES);
        if (j >= n || ldcg64(&t.entry[j]) == o.exit_off) break;  // re-synchronised with the guess
        tile = j;
        entry = o.exit_off;
      }
      g->repairs += reps;
    }
    grid.sync();
  }
  // ---- exclusive scan of (reach ? count : 0): every CTA owns a contiguous chunk of tiles ----
  __shared__ unsigned long long warp_sums[RESOLVE_THREADS / 32];
  __shared__ unsigned long long s_carry;
  const uint32_t chunk = (n + gridDim.x - 1) / gridDim.x;
  const uint32_t c0 = blockIdx.x * chunk, c1 = (c0 + chunk < n) ? c0 + chunk : n;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int pass = 0; pass < 2; ++pass) {
    if (threadIdx.x == 0) {
      unsigned long long basev = 0;
      if (pass == 1)
        for (uint32_t i = 0; i < blockIdx.x; ++i) basev += ldcg64(&g->chunk_sum[i]);
      s_carry = basev;
    }
    __syncthreads();
    for (uint32_t base = c0; base < c1; base += RESOLVE_THREADS) {
      uint32_t k = base + threadIdx.x;
      bool in = k < c1;
      unsigned long long v = (in && ldcg8(&t.reach[k])) ? ldcg32(&t.count[k]) : 0;
      unsigned long long x = v;
      for (int o = 1; o < 32; o <<= 1) {
        unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      if (lane == 31) warp_sums[wid] = x;
      __syncthreads();
      if (wid == 0) {
        unsigned long long s = lane < RESOLVE_THREADS / 32 ? warp_sums[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) {
          unsigned long long y = __shfl_up_sync(0xffffffffu, s, o);
          if (lane >= o) s += y;
        }
        if (lane < RESOLVE_THREADS / 32) warp_sums[lane] = s;  // inclusive
      }
      __syncthreads();
      unsigned long long off = s_carry + (wid ? warp_sums[wid - 1] : 0) + (x - v);
      if (pass == 1 && in) t.base[k] = off;
      __syncthreads();
      if (threadIdx.x == RESOLVE_THREADS - 1) s_carry = off + v;
      __syncthreads();
    }
    if (pass == 0) {
      if (threadIdx.x == 0) g->chunk_sum[blockIdx.x] = s_carry;
      grid.sync();
    }
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    res->n_records = s_carry;
    res->tail_off = ldcg64(&g->tail_off);
    res->tail_kind = ldcg32(&g->tail_kind);
    res->broken_tile = ldcg32(&g->broken_tile) >= n ? ldcg32(&g->broken_tile) : 0xffffffffu;
    res->broken_entry = 0;
    res->bad_flags = 0;
    res->rounds = ldcg32(&g->repairs);
    res->chain_start = ldcg64(&t.entry[k0]);
  }
}

// One WARP per tile: lane j restarts at the tile's j-th note (record CK_STRIDE * j of the tile) and writes the offsets and
// lengths of its CK_STRIDE records - chains of 8 dependent loads side by side instead of one chain of ~200 (round 1:
// 3.07 ms at 1.9 % issue utilisation for 50 M records).
// `cap`: capacity of rec_off / rec_len.  The host sizes the columns from the previous batch and learns the true count
// only after this kernel is already queued; records beyond cap are dropped here and the host re-runs with more room.
__global__ void __launch_bounds__(256) k3_tile_emit(RecordParams p, TileArrays t, uint64_t* rec_off, uint32_t* rec_len, uint64_t cap) {
  const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (k >= t.n_tiles || !t.reach[k] || t.entry[k] == NONE64) return;
  const uint32_t count = t.count[k];
  const uint64_t base = t.base[k], tb = (uint64_t)k * TILE_BYTES;
  const uint16_t* ck = t.ckpt + (size_t)k * CK_PER_TILE;
  for (uint32_t j = lane; j * CK_STRIDE < count; j += 32) {
    uint64_t q = tb + ck[j];
    const uint32_t first = j * CK_STRIDE, m = count - first < CK_STRIDE ? count - first : CK_STRIDE;
    for (uint32_t i = 0; i < m; ++i) {
      const uint32_t bs = ld32(p.data, q);
      const uint64_t idx = base + first + i;
      if (idx < cap) { rec_off[idx] = q + 4; rec_len[idx] = bs; }
      q += 4ull + bs;
    }
  }
}

// ---- HI tag: record/tags.rs:113-129 (walk: tags.cuh) through bamrawrecord.rs:80-104 (u8-wrapping offset) -------------
// returns 0 none, 1 found (value in *val), 2 = the reference would panic
__device__ int hit_count(const uint8_t* body, uint32_t len, int32_t* val) {
  uint32_t l_name = body[8];
  uint32_t cig = (32u + l_name) & 0xffu;  // bamrawrecord.rs:81 `32 + self.l_read_name()` in u8 (release wrap)
  uint32_t n_cig = body[12] | (body[13] << 8);
  uint32_t l_seq = body[16] | (body[17] << 8) | (body[18] << 16) | ((uint32_t)body[19] << 24);
  uint64_t tags = (uint64_t)cig + 4ull * n_cig + (uint64_t)((l_seq + 1u) / 2u) + l_seq;
  if (tags > len) return 2;
  uint64_t vo, vl;
  uint8_t ty;
  const int rc = tag_find(body + tags, len - tags, 'H', 'I', &vo, &vl, &ty);
  if (rc != 1) return rc;
  const uint8_t* v = body + tags + vo;
  switch (ty) {
    case 'A': case 'C': *val = (int32_t)v[0]; return 1;
    case 'c': *val = (int32_t)(int8_t)v[0]; return 1;
    case 's': *val = (int32_t)(int16_t)(v[0] | (v[1] << 8)); return 1;
    case 'S': *val = (int32_t)(v[0] | (v[1] << 8)); return 1;
    case 'i': case 'I': *val = (int32_t)(v[0] | (v[1] << 8) | (v[2] << 16) | ((uint32_t)v[3] << 24)); return 1;
    default: return 2;  // tags.rs:127 panic
  }
}

// `n` is the record count when the host knows it; with dev_count the kernel is launched for `n` = the columns' capacity
// and takes the count the resolve kernel left in res->n_records (no host round trip between the two).
__global__ void __launch_bounds__(256) k3_extract(RecordParams p, ColumnsDev c, uint64_t n, bool want_hi, ChainResult* res, bool dev_count) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (dev_count) { const uint64_t m = res->n_records; if (m < n) n = m; }
  bool bad = false;
  if (i < n) {
    const uint64_t off = c.rec_off[i];
    const uint32_t len = c.rec_len[i];
    // a body shorter than 32 bytes makes the reference's get_slice panic; the columns then hold
    // whatever bytes follow (still inside the buffer padding) — callers see rec_len < 32.
    uint32_t w0 = ld32(p.data, off), w1 = ld32(p.data, off + 4), w2 = ld32(p.data, off + 8), w3 = ld32(p.data, off + 12);
    uint32_t w4 = ld32(p.data, off + 16), w5 = ld32(p.data, off + 20), w6 = ld32(p.data, off + 24), w7 = ld32(p.data, off + 28);
    int32_t ref = (int32_t)w0, pos = (int32_t)w1;
    uint32_t l_name = w2 & 0xff, n_cig = w3 & 0xffff, flag = w3 >> 16, l_seq = w4;
    c.ref_id[i] = ref; c.pos[i] = pos;
    c.l_read_name[i] = (uint8_t)l_name; c.mapq[i] = (uint8_t)(w2 >> 8); c.bin[i] = (uint16_t)(w2 >> 16);
    c.n_cigar_op[i] = (uint16_t)n_cig; c.flag[i] = (uint16_t)flag; c.l_seq[i] = l_seq;
    c.next_ref_id[i] = (int32_t)w5; c.next_pos[i] = (int32_t)w6; c.tlen[i] = (int32_t)w7;
    uint32_t cig = 32u + l_name, seq = cig + 4u * n_cig, qual = seq + (l_seq + 1u) / 2u, tags = qual + l_seq;
    c.cigar_off[i] = cig; c.seq_off[i] = seq; c.qual_off[i] = qual; c.tags_off[i] = tags;
    uint32_t r = (ref == -1) ? 0x7fffffffu : (uint32_t)ref;               // comparators.rs:51-54
    c.coord_key[i] = ((uint64_t)(r ^ 0x80000000u) << 32) | (uint64_t)((uint32_t)pos ^ 0x80000000u);
    c.strand[i] = (flag & 0x10u) ? 1 : 0;                                  // comparators.rs:173-180
    bad = (flag & 0xF000u) != 0;                                           // Flags::from_bits(..).unwrap()
    uint8_t hp = 0; int32_t hv = 0;
    if (want_hi && len >= 20) { int rc = hit_count(p.data + off, len, &hv); hp = (uint8_t)rc; if (rc != 1) hv = 0; }
    c.hi_present[i] = hp; c.hi_value[i] = hv;
  }
  unsigned m = __ballot_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(&res->bad_flags, (unsigned long long)__popc(m));
}

}  // namespace

static inline uint32_t grid_for(uint64_t n, uint32_t block) { return (uint32_t)((n + block - 1) / block); }

void launch_tile_walk(const RecordParams& p, const TileArrays& t, cudaStream_t s) {
  if (t.n_tiles) k3_tile_walk<<<grid_for(t.n_tiles, 128), 128, 0, s>>>(p, t);
}

cudaError_t launch_chain_resolve(const RecordParams& p, const TileArrays& t, ChainResult* d_result, ResolveScratch* d_scratch,
                                 int num_sms, cudaStream_t s) {
  if (t.n_tiles == 0) return cudaSuccess;
  int grid = num_sms < RESOLVE_MAX_CTAS ? num_sms : RESOLVE_MAX_CTAS;
  int need = (int)((t.n_tiles + RESOLVE_THREADS - 1) / RESOLVE_THREADS);
  if (grid > need) grid = need;
  RecordParams pp = p;
  TileArrays tt = t;
  void* args[] = {&pp, &tt, &d_result, &d_scratch};
  return cudaLaunchCooperativeKernel((const void*)k3_resolve, dim3(grid), dim3(RESOLVE_THREADS), args, 0, s);
}

void launch_tile_emit(const RecordParams& p, const TileArrays& t, uint64_t* rec_off, uint32_t* rec_len, uint64_t cap, cudaStream_t s) {
  if (t.n_tiles) k3_tile_emit<<<grid_for((uint64_t)t.n_tiles * 32, 256), 256, 0, s>>>(p, t, rec_off, rec_len, cap);
}

void launch_extract(const RecordParams& p, const ColumnsDev& c, uint64_t n_records, bool want_hi, ChainResult* d_result,
                    bool dev_count, cudaStream_t s) {
  if (n_records) k3_extract<<<grid_for(n_records, 256), 256, 0, s>>>(p, c, n_records, want_hi, d_result, dev_count);
}

}  // namespace bamgpu
