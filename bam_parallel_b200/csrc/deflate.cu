// deflate.cu — K6: BGZF writer with COMPRESSED blocks ("optional BGZF re-compression to emit a real sorted
// BAM ... a GPU deflate writer is the natural follow-on").
//
// The reference's sort_bam takes an `_out_compr_level` it never uses (sorting/sort.rs:143) and dumps bodies (sort.rs:643);
// a sorted .bam that samtools / htslib / bam_tools::Reader can read needs BGZF blocks (gz.rs:9-20) whose payload is a
// DEFLATE stream (RFC 1951).  bgzf_write.cu frames STORED blocks at copy speed; this file compresses: every 0xff00-byte
// payload becomes one DEFLATE block with the FIXED Huffman code (BTYPE = 01), or a stored block when that would be larger.
//
// Mapping: one CTA per payload (BGZF blocks are independent, exactly as for K1), 512 threads, the payload resident in
// shared memory.  Per 2048-position chunk:
//   A. candidates: a single-probe hash table of 4-byte sequences (u16 positions, shared memory).  Positions are looked up
//      and inserted 256 at a time, so a candidate is at least the previous sub-step old — the adjacent BAM record (~330 B
//      back: read-name prefixes, tag layouts, flags) is visible; runs (binned qualities) are caught by a distance-1 probe;
//   B. every position verifies its candidates and extends the match 4 bytes per step (<= 258, never past the payload);
//   C. the greedy parse  next(p) = p + max(1, len(p))  is a linked list through the chunk: the positions on the chain that
//      starts where the previous chunk's last token ended are marked by pointer doubling (11 rounds), not by one thread
//      walking ~350 tokens;
//   D. marked positions turn into fixed-Huffman bit strings (<= 31 bits: literal, or length code + extra + distance code +
//      extra), a block-wide exclusive scan of the bit lengths places them, atomicOr into a shared-memory word buffer packs
//      them, and the whole words go to the payload's slot in global memory with coalesced stores.
// A second kernel lays the blocks out back to back: header (BSIZE), DEFLATE bytes (or the stored form), CRC32 + ISIZE.
// CRC32 comes from K2's warp CRC (crc32.cu).  Integer / byte work, HBM- and shared-memory-bound; no tensor-core content.
#include <cub/cub.cuh>

#include <cstdio>

#include "kernels.h"

namespace bamgpu {
namespace {

#ifndef DF_WALK
#define DF_WALK 1
#endif
#ifndef DF_LAZY
#define DF_LAZY 0
#endif
#ifndef DF_PREFETCH
#define DF_PREFETCH 0
#endif
constexpr int DF_THREADS = 512;
constexpr int DF_CHUNK = 2048;                 // positions per parse round
#ifndef DF_SUB_N
#define DF_SUB_N 256
#endif
constexpr int DF_SUB = DF_SUB_N;            // positions looked up / inserted together
#ifndef DF_TAB32
#define DF_TAB32 1
#endif
constexpr int DF_HASH_BITS = DF_TAB32 ? 12 : 13;   // 16 KB of shared memory either way: 4096 x u32 or 8192 x u16
constexpr int DF_ITEMS = DF_CHUNK / DF_THREADS;
constexpr uint32_t DF_MAX_MATCH = 258, DF_MIN_MATCH = 4;
constexpr uint32_t DF_RAW_BYTES = BGZF_STORE_PAYLOAD + 48;   // payload + alignment slack in front + zero pad behind

struct DfShared {
  uint32_t raw[DF_RAW_BYTES / 4];
  uint16_t htab[8192];                         // DF_TAB32: used as 4096 x u32
  uint16_t mlen[DF_CHUNK];                     // phase A leaves the candidate (position + 1) here, phase B replaces it by the match length
  uint16_t mdist[DF_CHUNK];
  uint16_t jmp[2][DF_CHUNK];
  uint8_t mark[DF_CHUNK];
  uint32_t outw[DF_CHUNK + 8];                 // <= 31 bits per position + the carried word
  uint32_t exit_rel, overflow;
};

__device__ __forceinline__ uint32_t ld32u(const uint32_t* __restrict__ raw, uint32_t byte_off) {
  const uint32_t w = byte_off >> 2;
  return __funnelshift_r(raw[w], raw[w + 1], (byte_off & 3u) * 8u);
}

// bytes a.. and b.. agree for the first 4; how far do they agree (<= maxl)?
__device__ __forceinline__ uint32_t extend_match(const uint32_t* __restrict__ raw, uint32_t a, uint32_t b, uint32_t maxl) {
  uint32_t l = 4;
  while (l < maxl) {
    const uint32_t x = ld32u(raw, a + l) ^ ld32u(raw, b + l);
    if (x) { l += (uint32_t)(__ffs((int)x) - 1) >> 3; break; }
    l += 4;
  }
  return l < maxl ? l : maxl;
}

// Fixed-Huffman bit string of one token, LSB first (Huffman codes bit-reversed, extra bits as they are): RFC 1951 3.2.5/3.2.6.
__device__ __forceinline__ void token_bits(uint32_t lit, uint32_t len, uint32_t dist, uint32_t& bits, uint32_t& nbits) {
  if (len == 0) {
    const uint32_t c = lit < 144 ? 0x30u + lit : 0x190u + (lit - 144u);
    nbits = lit < 144 ? 8u : 9u;
    bits = __brev(c) >> (32 - nbits);
    return;
  }
  const uint32_t y = len - 3;
  uint32_t sym, eb, ev;
  if (y < 8) { sym = 257 + y; eb = 0; ev = 0; }
  else if (len == 258) { sym = 285; eb = 0; ev = 0; }
  else {
    const uint32_t nb = 31u - (uint32_t)__clz((int)y);
    eb = nb - 2;
    sym = 257 + 4 * (nb - 1) + ((y >> eb) & 3u);
    ev = y & ((1u << eb) - 1u);
  }
  uint32_t cb = sym < 280 ? 7u : 8u;
  const uint32_t c = sym < 280 ? sym - 256u : 0xC0u + (sym - 280u);
  uint32_t v = __brev(c) >> (32 - cb);
  v |= ev << cb; cb += eb;
  const uint32_t x = dist - 1;
  uint32_t dsym, deb, dev;
  if (x < 4) { dsym = x; deb = 0; dev = 0; }
  else {
    const uint32_t nb = 31u - (uint32_t)__clz((int)x);
    deb = nb - 1;
    dsym = 2 * nb + ((x >> deb) & 1u);
    dev = x & ((1u << deb) - 1u);
  }
  v |= (__brev(dsym) >> 27) << cb; cb += 5;
  v |= dev << cb; cb += deb;
  bits = v;
  nbits = cb;   // <= 8 + 5 + 5 + 13 = 31
}

// Phases A - C for the chunk [c0, c0 + clen) of an n-byte payload at raw + mis: on return mark[i] != 0 for the positions of the
// greedy parse, mlen / mdist hold their matches (mlen 0: a literal).  `entry`: where the chain enters the chunk.
using DfScan = cub::BlockScan<uint32_t, DF_THREADS>;
#ifdef DF_PROFILE   // diagnostic build: cycles of CTA 0's thread 0 per phase
__device__ unsigned long long df_prof[8];
#define DF_TICK(k) do { __syncthreads(); if (threadIdx.x == 0 && blockIdx.x == 0) { const long long t_ = clock64(); df_prof[k] += (unsigned long long)(t_ - t_prof); t_prof = t_; } } while (0)
#else
#define DF_TICK(k) do { } while (0)
#endif
__device__ __forceinline__ void parse_chunk(DfShared& S, DfScan::TempStorage& scan_tmp, uint32_t mis, uint32_t c0, uint32_t clen, uint32_t n,
                                            uint32_t entry, uint32_t tid) {
  // runs: nz[i] = the first chunk position q >= i whose byte differs from the byte in front of it (clen: none), by a reverse
  // min-scan.  The distance-1 match at i is then nz[i] - i bytes long: no position of a long run (binned base qualities) loops
  // over it - with a compare loop per position a warp waited for its longest run at every one of them (6x the time).
  uint16_t* nz = S.jmp[1];
#ifdef DF_PROFILE
  long long t_prof = clock64();
#endif
  {
    const uint8_t* rb = reinterpret_cast<const uint8_t*>(S.raw) + mis + c0;
    uint32_t v[DF_ITEMS];
#pragma unroll
    for (int k = 0; k < DF_ITEMS; ++k) {
      const uint32_t r = tid * DF_ITEMS + k;
      v[k] = clen;
      if (r < clen) {
        const uint32_t i = clen - 1 - r;
        const bool same = c0 + i > 0 && rb[i] == rb[(int)i - 1];
        v[k] = same ? clen : i;
      }
    }
    DfScan(scan_tmp).InclusiveScan(v, v, cub::Min());
#pragma unroll
    for (int k = 0; k < DF_ITEMS; ++k) {
      const uint32_t r = tid * DF_ITEMS + k;
      if (r < clen) nz[clen - 1 - r] = (uint16_t)v[k];
    }
  }
  __syncthreads();
  DF_TICK(0);
  // A. candidates
  for (uint32_t s0 = 0; s0 < clen; s0 += DF_SUB) {
    uint32_t h = 0;
    bool live = false;
    const uint32_t i = s0 + tid;
    if (tid < DF_SUB && i < clen && c0 + i + DF_MIN_MATCH <= n) {
      live = true;
      h = (ld32u(S.raw, mis + c0 + i) * 2654435761u) >> (32 - DF_HASH_BITS);
      S.mlen[i] = DF_TAB32 ? (uint16_t)reinterpret_cast<const uint32_t*>(S.htab)[h] : S.htab[h];
    } else if (tid < DF_SUB && i < clen) {
      S.mlen[i] = 0;
    }
    __syncthreads();
    // Positions inside a run (the same 4 bytes as one position earlier) are not inserted: the distance-1 probe covers them, and
    // hundreds of equal hashes fighting for one table word made this phase 70 % of the kernel.  Among the lanes of a warp that
    // still share a hash, only the highest position goes to the table.
    if (live && nz[i] - i >= 4) live = false;
#if DF_TAB32
    // one 32-bit word per entry (position + 1, 0: empty): atomicMax lets the HIGHEST position of the sub-step win whatever the
    // thread order - the same bytes on every run - in one instruction
    if (live) atomicMax(reinterpret_cast<uint32_t*>(S.htab) + h, c0 + i + 1);
#else
    if (tid < DF_SUB) {
      const uint32_t peers = __match_any_sync(0xffffffffu, live ? h : 0xffffffffu - (tid & 31u));
      if (live && (peers >> (tid & 31u)) > 1u) live = false;      // a higher lane (= a higher position) has the same hash
    }
    if (live) {
      // entry = position + 1 (0: empty); the HIGHEST position of the sub-step wins, whatever the thread order: the output
      // is the same bytes on every run.  Two entries share a 32-bit word: compare-and-swap on the word.
      uint32_t* w = reinterpret_cast<uint32_t*>(S.htab) + (h >> 1);
      const uint32_t sh = (h & 1u) * 16u, mine = c0 + i + 1;
      uint32_t old = *w;
      for (;;) {
        if (((old >> sh) & 0xffffu) >= mine) break;
        const uint32_t seen = atomicCAS(w, old, (old & ~(0xffffu << sh)) | (mine << sh));
        if (seen == old) break;
        old = seen;
      }
    }
#endif
    __syncthreads();
  }
  DF_TICK(1);
#if DF_LAZY
  // B + C in one: every warp walks its own 128 positions, 32 at a time.  The lanes only VERIFY their candidates (4 bytes) and
  // look up their run lengths; the first lane with a match is the next token, and its match is extended by the whole warp (128
  // bytes per round).  Positions a match jumps over are never looked at.  Same tokens as the two-phase form below.
  {
    const uint32_t lane = tid & 31u, per = DF_CHUNK / (DF_THREADS / 32), w0 = (tid >> 5) * per;
    const uint32_t end = w0 + per < clen ? w0 + per : clen;
    for (uint32_t q = w0 + lane; q < end; q += 32) S.mark[q] = 0;
    __syncwarp();
    for (uint32_t c = w0; c < end;) {
      const uint32_t q = c + lane, pos = c0 + q;
      uint32_t l1 = 0, c1 = 0, maxl = 0;
      bool okh = false;
      if (q < end && pos + DF_MIN_MATCH <= n) {
        maxl = n - pos < DF_MAX_MATCH ? n - pos : DF_MAX_MATCH;
        l1 = nz[q] - q;
        if (l1 && q + l1 == clen && l1 < maxl) {
          const uint8_t* rb = reinterpret_cast<const uint8_t*>(S.raw) + mis + pos;
          while (l1 < maxl && rb[l1] == rb[l1 - 1]) ++l1;
        }
        l1 = l1 < maxl ? l1 : maxl;
        c1 = S.mlen[q];
        okh = l1 < 64 && c1 && pos - (c1 - 1) <= 32768u && ld32u(S.raw, mis + c1 - 1) == ld32u(S.raw, mis + pos);
      }
      const uint32_t any = __ballot_sync(0xffffffffu, okh || l1 >= DF_MIN_MATCH);
      if (!any) {
        if (q < end) { S.mlen[q] = 0; S.mark[q] = 1; }
        c += 32;
        __syncwarp();
        continue;
      }
      const uint32_t f = (uint32_t)__ffs((int)any) - 1u, qf = c + f, posf = c0 + qf;
      if (lane < f) { S.mlen[q] = 0; S.mark[q] = 1; }
      const bool hf = __shfl_sync(0xffffffffu, (int)okh, (int)f) != 0;
      const uint32_t l1f = __shfl_sync(0xffffffffu, l1, (int)f), c1f = __shfl_sync(0xffffffffu, c1, (int)f);
      const uint32_t maxlf = __shfl_sync(0xffffffffu, maxl, (int)f);
      uint32_t best = l1f >= DF_MIN_MATCH ? l1f : 0u, dist = 1;
      if (hf) {
        const uint32_t a = mis + c1f - 1, b = mis + posf;
        uint32_t lh = maxlf;
        for (uint32_t base = 0; base < maxlf; base += 128) {
          const uint32_t off = base + 4 * lane;
          const uint32_t x = off < maxlf ? (ld32u(S.raw, a + off) ^ ld32u(S.raw, b + off)) : 1u;
          const uint32_t mm = __ballot_sync(0xffffffffu, x != 0);
          if (mm) {
            const uint32_t m = (uint32_t)__ffs((int)mm) - 1u;
            const uint32_t xm = __shfl_sync(0xffffffffu, x, (int)m);
            lh = base + 4 * m + ((uint32_t)(__ffs((int)xm) - 1) >> 3);
            break;
          }
        }
        lh = lh < maxlf ? lh : maxlf;
        if (lh > best) { best = lh; dist = posf - (c1f - 1); }
      }
      uint32_t L = best;
      if (qf + L > end) { L = end - qf; if (L < 3) L = 0; }
      if (lane == f) { S.mlen[qf] = (uint16_t)L; S.mdist[qf] = (uint16_t)dist; S.mark[qf] = 1; }
      c = qf + (L ? L : 1u);
      __syncwarp();
    }
  }
  __syncthreads();
  DF_TICK(3);
}
#else
  // B. verify + extend.  (Giving a thread DF_ITEMS consecutive positions and deriving the match inside a long match from the position
  // in front of it - no compare loop there - was measured SLOWER, 34.1 vs 32.1 ms: the loops are not what this phase waits for.)
#if DF_PREFETCH
  // the loads every position needs - run length, candidate, its four bytes and the candidate's - for all DF_ITEMS positions of the
  // thread first: four chains of dependent shared-memory loads overlap instead of following each other (the compare loops below keep
  // the compiler from doing it)
  uint32_t pf_l1[DF_ITEMS], pf_c1[DF_ITEMS], pf_eq[DF_ITEMS];
#pragma unroll
  for (int k = 0; k < DF_ITEMS; ++k) {
    const uint32_t i = tid + k * DF_THREADS;
    pf_l1[k] = 0; pf_c1[k] = 0;
    if (i < clen) { pf_l1[k] = nz[i] - i; pf_c1[k] = S.mlen[i]; }
  }
#pragma unroll
  for (int k = 0; k < DF_ITEMS; ++k) {
    const uint32_t i = tid + k * DF_THREADS, pos = c0 + i;
    pf_eq[k] = 0;
    if (i < clen && pos + DF_MIN_MATCH <= n && pf_c1[k] && pos - (pf_c1[k] - 1) <= 32768u)
      pf_eq[k] = ld32u(S.raw, mis + pf_c1[k] - 1) == ld32u(S.raw, mis + pos) ? 1u : 0u;
  }
#endif
#pragma unroll
  for (int k = 0; k < DF_ITEMS; ++k) {
    const uint32_t i = tid + k * DF_THREADS;
    if (i < clen) {
      const uint32_t pos = c0 + i;
      uint32_t best = 0, bdist = 0;
      if (pos + DF_MIN_MATCH <= n) {
        const uint32_t maxl = n - pos < DF_MAX_MATCH ? n - pos : DF_MAX_MATCH;
#if DF_PREFETCH
        uint32_t l1 = pf_l1[k];
#else
        uint32_t l1 = nz[i] - i;              // bytes from pos on that repeat the byte in front of them
#endif
        if (l1 && i + l1 == clen && l1 < maxl) {   // the run reaches the end of the chunk: follow it (few positions per chunk)
          const uint8_t* rb = reinterpret_cast<const uint8_t*>(S.raw) + mis + pos;
          while (l1 < maxl && rb[l1] == rb[l1 - 1]) ++l1;
        }
        l1 = l1 < maxl ? l1 : maxl;
#if DF_PREFETCH
        const uint32_t c1 = pf_c1[k];
        if (l1 < 64 && pf_eq[k]) {
#else
        const uint32_t c1 = S.mlen[i];        // candidate position + 1 (same thread overwrites it below)
        if (l1 < 64 && c1 && pos - (c1 - 1) <= 32768u && ld32u(S.raw, mis + c1 - 1) == ld32u(S.raw, mis + pos)) {   // the DEFLATE window is 32 KiB
#endif
#ifdef DF_EXT_CAP   // experiment: how much of this phase is the tail of long compare loops?
          best = extend_match(S.raw, mis + c1 - 1, mis + pos, maxl < DF_EXT_CAP ? maxl : DF_EXT_CAP);
#else
          best = extend_match(S.raw, mis + c1 - 1, mis + pos, maxl);
#endif
          bdist = pos - (c1 - 1);
        }
        if (l1 >= DF_MIN_MATCH && l1 >= best) { best = l1; bdist = 1; }
      }
      S.mlen[i] = (uint16_t)best;
      S.mdist[i] = (uint16_t)bdist;
#if DF_WALK
      S.mark[i] = 0;
#else
      const uint32_t nx = i + (best ? best : 1u);
      S.jmp[0][i] = (uint16_t)(nx < clen ? nx : DF_CHUNK);
      S.mark[i] = i == entry ? 1 : 0;
#endif
    }
  }
  __syncthreads();
  DF_TICK(2);
#if DF_WALK
  // C. the greedy parse  next(p) = p + max(1, len(p)): every warp walks its own 128 positions, 32 at a time - the first lane with a
  // match decides how far the cursor jumps, the lanes in front of it are literals.  Matches are cut at the end of the warp's range
  // (so every range, and every chunk, starts a token: no chain from range to range, no block-wide rounds); a cut shorter than 3
  // bytes becomes a literal.  (Marking the chain of the whole chunk by pointer doubling - 11 block-wide rounds - was 24 % of the
  // kernel's instructions and most of its barriers; -DDF_WALK=0 builds it.)
  {
    const uint32_t lane = tid & 31u, w0 = (tid >> 5) * (DF_CHUNK / (DF_THREADS / 32));
    const uint32_t end = w0 + DF_CHUNK / (DF_THREADS / 32) < clen ? w0 + DF_CHUNK / (DF_THREADS / 32) : clen;
    for (uint32_t c = w0; c < end;) {
      const uint32_t q = c + lane;
      const uint32_t l = q < end ? S.mlen[q] : 0u;
      const uint32_t any = __ballot_sync(0xffffffffu, l != 0);
      if (!any) {
        if (q < end) S.mark[q] = 1;
        c += 32;
        continue;
      }
      const uint32_t f = (uint32_t)__ffs((int)any) - 1u;
      uint32_t L = __shfl_sync(0xffffffffu, l, (int)f);
      const uint32_t qf = c + f;
      if (qf + L > end) {
        L = end - qf;
        if (L < 3) L = 0;
        if (lane == f) S.mlen[qf] = (uint16_t)L;
      }
      if (lane <= f) S.mark[q] = 1;
      c = qf + (L ? L : 1u);
      __syncwarp();
    }
  }
  __syncthreads();
  DF_TICK(3);
}
#else
  // C. mark the greedy chain by pointer doubling
  int cur = 0;
  for (uint32_t span = 1; span < clen; span <<= 1) {
#pragma unroll
    for (int k = 0; k < DF_ITEMS; ++k) {
      const uint32_t i = tid + k * DF_THREADS;
      if (i < clen) {
        const uint32_t j = S.jmp[cur][i];
        if (j < DF_CHUNK && S.mark[i]) S.mark[j] = 1;
        S.jmp[cur ^ 1][i] = (uint16_t)(j < DF_CHUNK ? S.jmp[cur][j] : DF_CHUNK);
      }
    }
    cur ^= 1;
    __syncthreads();
  }
}
#endif
#endif

// One CTA per payload.  slots: n_blocks x DF_SLOT bytes; sizes[b] = DEFLATE bytes, or 0 = "store it" (no gain / overflow).
constexpr uint32_t DF_SLOT = 1u << 16;
__global__ void __launch_bounds__(DF_THREADS, 2)
k6_deflate_fixed(const uint8_t* __restrict__ src, unsigned long long len, uint8_t* __restrict__ slots, uint32_t* __restrict__ sizes,
                 uint32_t n_blocks) {
  extern __shared__ __align__(16) uint8_t df_smem[];
  DfShared& S = *reinterpret_cast<DfShared*>(df_smem);
  using Scan = cub::BlockScan<uint32_t, DF_THREADS>;
  __shared__ typename Scan::TempStorage scan_tmp;
  const uint32_t tid = threadIdx.x;
  for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const unsigned long long o = (unsigned long long)blk * BGZF_STORE_PAYLOAD;
    const uint32_t n = (uint32_t)(len - o < BGZF_STORE_PAYLOAD ? len - o : BGZF_STORE_PAYLOAD);
    const uint8_t* p = src + o;
    const uint32_t mis = (uint32_t)((uintptr_t)p & 15u);     // the payload sits at raw + mis: aligned 16-byte loads either side
    {
      const uint4* g = reinterpret_cast<const uint4*>(p - mis);
      uint4* d = reinterpret_cast<uint4*>(S.raw);
      const uint32_t vecs = (mis + n) / 16;                      // whole vectors; the bytes behind them one by one (nothing is read
      for (uint32_t i = tid; i < vecs; i += DF_THREADS) d[i] = g[i];   // past src + len)
      for (uint32_t i = vecs * 16 + tid; i < mis + n; i += DF_THREADS) if (i >= mis) reinterpret_cast<uint8_t*>(S.raw)[i] = (p - mis)[i];
      for (uint32_t i = tid; i < 8192; i += DF_THREADS) S.htab[i] = 0;
      if (tid == 0) S.overflow = 0;
    }
    __syncthreads();
    {   // zero what follows the payload (match extension may look up to 7 bytes past it)
      uint8_t* rb = reinterpret_cast<uint8_t*>(S.raw);
      for (uint32_t i = mis + n + tid; i < ((mis + n + 15) / 16) * 16 + 16 && i < DF_RAW_BYTES; i += DF_THREADS) rb[i] = 0;
    }
    __syncthreads();
    uint32_t* gout = reinterpret_cast<uint32_t*>(slots + (size_t)blk * DF_SLOT);
    uint32_t wpos = 0;              // whole words already in gout
    uint32_t carry_word = 3u;       // BFINAL = 1, BTYPE = 01
    uint32_t carry_bits = 3;
    uint32_t entry = 0;             // where the chain enters the chunk (relative)
    for (uint32_t c0 = 0; c0 < n; c0 += DF_CHUNK) {
      const uint32_t clen = n - c0 < DF_CHUNK ? n - c0 : DF_CHUNK;
      parse_chunk(S, scan_tmp, mis, c0, clen, n, entry, tid);
#ifdef DF_PROFILE
      long long t_prof = clock64();
#endif
      // D. tokens -> bits
      uint32_t bits[DF_ITEMS], nb[DF_ITEMS], off[DF_ITEMS];
      const uint8_t* rb = reinterpret_cast<const uint8_t*>(S.raw) + mis + c0;
#pragma unroll
      for (int k = 0; k < DF_ITEMS; ++k) {
        const uint32_t i = tid * DF_ITEMS + k;   // blocked arrangement for the scan
        bits[k] = 0; nb[k] = 0;
        if (i < clen && S.mark[i]) {
          const uint32_t l = S.mlen[i];
          token_bits(rb[i], l, S.mdist[i], bits[k], nb[k]);
          const uint32_t nx = i + (l ? l : 1u);
          if (nx >= clen) S.exit_rel = nx - clen;   // exactly one marked position leaves the chunk
        }
      }
      if (entry >= clen && tid == 0) S.exit_rel = entry - clen;   // a match from the previous chunk covers this (short, last) one
      uint32_t total = 0;
      Scan(scan_tmp).ExclusiveSum(nb, off, total);
      for (uint32_t i = tid; i < (carry_bits + total + 31) / 32 + 1; i += DF_THREADS) S.outw[i] = i == 0 ? carry_word : 0u;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < DF_ITEMS; ++k) {
        if (nb[k]) {
          const uint32_t bo = carry_bits + off[k], w = bo >> 5, sh = bo & 31u;
          atomicOr(&S.outw[w], bits[k] << sh);
          if (sh + nb[k] > 32) atomicOr(&S.outw[w + 1], bits[k] >> (32 - sh));
        }
      }
      __syncthreads();
      const uint32_t tb = carry_bits + total, full = tb >> 5;
      if ((wpos + full + 2) * 4 > DF_SLOT) { if (tid == 0) S.overflow = 1; }
      else for (uint32_t i = tid; i < full; i += DF_THREADS) gout[wpos + i] = S.outw[i];
      carry_word = S.outw[full];
      carry_bits = tb & 31u;
      entry = DF_WALK ? 0u : S.exit_rel;
      __syncthreads();
      if (!S.overflow) wpos += full;
      DF_TICK(4);
    }
    // end of block: symbol 256 = seven zero bits
    if (tid == 0) {
      uint32_t tb = carry_bits + 7;
      uint32_t bytes = 0;
      if (!S.overflow) {
        gout[wpos] = carry_word;
        if (tb > 32) gout[wpos + 1] = 0;
        bytes = wpos * 4 + (tb + 7) / 8;
      }
      sizes[blk] = (S.overflow || bytes >= n + 5) ? 0u : bytes;
    }
    __syncthreads();
  }
}

// ---- dynamic Huffman codes (BTYPE = 10, RFC 1951 3.2.7): level >= 2 ---------------------------------------------------
// Pass 1 is the parse above; instead of emitting bits it compacts the tokens of every chunk into a per-CTA buffer in global
// memory (it stays in L2) and counts symbols.  Then the CTA builds the two codes - sort by frequency (bitonic, all threads),
// two-queue Huffman merge (one thread: ~300 dependent steps), leaf depths by walking up from every leaf in parallel, canonical
// codes by counting - writes the header with the code lengths sent one by one through the code-length code (no repeat
// symbols: ~0.5 % larger, a third of the code), and pass 2 turns the tokens into bits with table look-ups.  If the fixed code
// would be smaller (tiny payloads) its tables are loaded instead; if neither beats a stored block, the block is stored.
struct DfBuild {            // overlays DfShared::htab (dead after pass 1)
  uint32_t key[512];        // freq << 9 | symbol, ascending; 0xffffffff = unused
  uint32_t weight[576];
  uint16_t parent[576];
  uint16_t lcode[288], dcode[32], ccode[20];
  uint8_t llen[288], dlen[32], cllen[20];
  uint32_t bl_count[16], next_code[16];
  uint32_t cfreq[20];
  uint32_t maxd, hlit, hdist, hclen, cost_dyn, cost_fix, use_fixed;
};
static_assert(sizeof(DfBuild) <= sizeof(uint16_t) * 8192, "DfBuild must fit the hash table it overlays");
struct DfDyn {
  DfShared P;
  uint32_t lfreq[288], dfreq[32];
  uint32_t ntok_chunk;
};
constexpr uint32_t DF_TOK_BATCH = 1024;        // tokens per emit round (2 per thread)
__constant__ uint8_t DF_CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// lens[0..nsym) = code lengths (<= maxbits) of a Huffman code for freq[0..nsym); at least two symbols get a code.  All threads.
__device__ void huff_lengths(const uint32_t* __restrict__ freq, uint32_t nsym, uint32_t maxbits, uint8_t* __restrict__ lens, DfBuild& B, uint32_t tid) {
  for (uint32_t scale = 0;; ++scale) {
    uint32_t f = tid < nsym ? freq[tid] : 0u;
    if (f && scale) { f >>= scale; f = f ? f : 1u; }
    B.key[tid] = f ? (f << 9 | tid) : 0xffffffffu;
    if (tid < nsym) lens[tid] = 0;
    if (tid == 0) B.maxd = 0;
    uint32_t n = (uint32_t)__syncthreads_count(f != 0);
    for (uint32_t k = 2; k <= 512; k <<= 1)
      for (uint32_t j = k >> 1; j > 0; j >>= 1) {
        const uint32_t o = tid ^ j;
        if (o > tid) {
          const uint32_t a = B.key[tid], b = B.key[o];
          if ((a > b) == ((tid & k) == 0)) { B.key[tid] = b; B.key[o] = a; }
        }
        __syncthreads();
      }
    if (tid == 0) {   // DEFLATE wants a complete code: at least two symbols (zlib build_tree does the same)
      if (n == 0) { B.key[0] = (1u << 9) | 0u; B.key[1] = (1u << 9) | 1u; }
      else if (n == 1) {
        const uint32_t only = B.key[0], dummy = (1u << 9) | ((only & 511u) ? 0u : 1u);
        if (dummy < only) { B.key[1] = only; B.key[0] = dummy; } else B.key[1] = dummy;
      }
    }
    if (n < 2) n = 2;
    __syncthreads();
    if (tid < n) B.weight[tid] = B.key[tid] >> 9;
    __syncthreads();
    if (tid == 0) {   // two queues: leaves in ascending order, internal nodes in the order they are made (ascending too)
      uint32_t li = 0, ni = n, nn = n;
      uint32_t wl = B.weight[0], wn = 0xffffffffu;
      for (uint32_t m = 0; m + 1 < n; ++m) {
        uint32_t pick[2], w = 0;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (li < n && (ni >= nn || wl <= wn)) { pick[q] = li; w += wl; ++li; wl = li < n ? B.weight[li] : 0xffffffffu; }
          else { pick[q] = ni; w += wn; ++ni; wn = ni < nn ? B.weight[ni] : 0xffffffffu; }
        }
        B.weight[nn] = w;
        B.parent[pick[0]] = (uint16_t)nn;
        B.parent[pick[1]] = (uint16_t)nn;
        if (ni == nn) wn = w;       // the new node is the head of the internal queue
        ++nn;
      }
    }
    __syncthreads();
    if (tid < n) {
      const uint32_t root = 2 * n - 2;
      uint32_t d = 0, k = tid;
      while (k != root) { k = B.parent[k]; ++d; }
      lens[B.key[tid] & 511u] = (uint8_t)d;
      atomicMax(&B.maxd, d);
    }
    __syncthreads();
    if (B.maxd <= maxbits) break;     // else: flatter frequencies, again (ends at all-ones = a balanced tree)
    __syncthreads();
  }
}

// codes[s] = the canonical code of symbol s, bit-reversed (DEFLATE packs Huffman codes most significant bit first).  All threads.
__device__ void canonical_codes(const uint8_t* __restrict__ lens, uint32_t nsym, uint16_t* __restrict__ codes, DfBuild& B, uint32_t tid) {
  if (tid < 16) B.bl_count[tid] = 0;
  __syncthreads();
  if (tid < nsym && lens[tid]) atomicAdd(&B.bl_count[lens[tid]], 1u);
  __syncthreads();
  if (tid == 0) {
    uint32_t code = 0;
    B.next_code[0] = 0;
    for (uint32_t L = 1; L <= 15; ++L) { code = (code + B.bl_count[L - 1]) << 1; B.next_code[L] = code; }
  }
  __syncthreads();
  if (tid < nsym) {
    const uint32_t L = lens[tid];
    uint32_t c = 0;
    if (L) {
      uint32_t r = 0;
      for (uint32_t t = 0; t < tid; ++t) r += lens[t] == L ? 1u : 0u;
      c = __brev(B.next_code[L] + r) >> (32 - L);
    }
    codes[tid] = (uint16_t)c;
  }
  __syncthreads();
}

__device__ __forceinline__ void len_symbol(uint32_t len, uint32_t& sym, uint32_t& eb, uint32_t& ev) {
  const uint32_t y = len - 3;
  if (y < 8) { sym = 257 + y; eb = 0; ev = 0; }
  else if (len == 258) { sym = 285; eb = 0; ev = 0; }
  else {
    const uint32_t nb = 31u - (uint32_t)__clz((int)y);
    eb = nb - 2;
    sym = 257 + 4 * (nb - 1) + ((y >> eb) & 3u);
    ev = y & ((1u << eb) - 1u);
  }
}
__device__ __forceinline__ void dist_symbol(uint32_t dist, uint32_t& sym, uint32_t& eb, uint32_t& ev) {
  const uint32_t x = dist - 1;
  if (x < 4) { sym = x; eb = 0; ev = 0; }
  else {
    const uint32_t nb = 31u - (uint32_t)__clz((int)x);
    eb = nb - 1;
    sym = 2 * nb + ((x >> eb) & 1u);
    ev = x & ((1u << eb) - 1u);
  }
}

struct DfOut { uint32_t* gout; uint32_t wpos, carry_word, carry_bits; };

// Packs the items (blocked order: thread t holds items 2 t, 2 t + 1; nb = 0: none) behind the bits written so far.  All threads.
template <class ScanT>
__device__ __forceinline__ void emit_batch(const uint64_t (&bits)[2], uint32_t (&nb)[2], DfOut& O, DfShared& S, typename ScanT::TempStorage& tmp, uint32_t tid) {
  uint32_t off[2], total = 0;
  ScanT(tmp).ExclusiveSum(nb, off, total);
  for (uint32_t i = tid; i < (O.carry_bits + total + 31) / 32 + 2; i += DF_THREADS) S.outw[i] = i == 0 ? O.carry_word : 0u;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (nb[k]) {
      const uint32_t bo = O.carry_bits + off[k], w = bo >> 5, sh = bo & 31u;
      atomicOr(&S.outw[w], (uint32_t)(bits[k] << sh));
      const uint64_t rest = sh ? bits[k] >> (32 - sh) : bits[k] >> 32;
      if ((uint32_t)rest) atomicOr(&S.outw[w + 1], (uint32_t)rest);
      if (rest >> 32) atomicOr(&S.outw[w + 2], (uint32_t)(rest >> 32));
    }
  }
  __syncthreads();
  const uint32_t tb = O.carry_bits + total, full = tb >> 5;
  const bool room = (O.wpos + full + 2) * 4 <= DF_SLOT;
  if (!room) { if (tid == 0) S.overflow = 1; }
  else for (uint32_t i = tid; i < full; i += DF_THREADS) O.gout[O.wpos + i] = S.outw[i];
  O.carry_word = S.outw[full];
  O.carry_bits = tb & 31u;
  __syncthreads();
  if (!S.overflow) O.wpos += full;
}

__global__ void __launch_bounds__(DF_THREADS, 2)
k6_deflate_dynamic(const uint8_t* __restrict__ src, unsigned long long len, uint8_t* __restrict__ slots, uint32_t* __restrict__ sizes,
                   uint32_t* __restrict__ tokbuf, uint32_t n_blocks) {
  extern __shared__ __align__(16) uint8_t df_smem[];
  DfDyn& D = *reinterpret_cast<DfDyn*>(df_smem);
  DfShared& S = D.P;
  DfBuild& B = *reinterpret_cast<DfBuild*>(S.htab);
  uint32_t* stage = reinterpret_cast<uint32_t*>(S.jmp);      // 2048 tokens: the jump arrays are dead once the chain is marked
  using Scan = cub::BlockScan<uint32_t, DF_THREADS>;
  __shared__ typename Scan::TempStorage scan_tmp;
  const uint32_t tid = threadIdx.x;
  uint32_t* gtok = tokbuf + (size_t)blockIdx.x * BGZF_STORE_PAYLOAD;
  for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const unsigned long long o = (unsigned long long)blk * BGZF_STORE_PAYLOAD;
    const uint32_t n = (uint32_t)(len - o < BGZF_STORE_PAYLOAD ? len - o : BGZF_STORE_PAYLOAD);
    const uint8_t* p = src + o;
    const uint32_t mis = (uint32_t)((uintptr_t)p & 15u);
    {
      const uint4* g = reinterpret_cast<const uint4*>(p - mis);
      uint4* d = reinterpret_cast<uint4*>(S.raw);
      const uint32_t vecs = (mis + n) / 16;                      // whole vectors; the bytes behind them one by one (nothing is read
      for (uint32_t i = tid; i < vecs; i += DF_THREADS) d[i] = g[i];   // past src + len)
      for (uint32_t i = vecs * 16 + tid; i < mis + n; i += DF_THREADS) if (i >= mis) reinterpret_cast<uint8_t*>(S.raw)[i] = (p - mis)[i];
      for (uint32_t i = tid; i < 8192; i += DF_THREADS) S.htab[i] = 0;
      if (tid < 288) D.lfreq[tid] = tid == 256 ? 1u : 0u;     // one end-of-block symbol
      if (tid < 32) D.dfreq[tid] = 0;
      if (tid == 0) S.overflow = 0;
    }
    __syncthreads();
    {
      uint8_t* rb = reinterpret_cast<uint8_t*>(S.raw);
      for (uint32_t i = mis + n + tid; i < ((mis + n + 15) / 16) * 16 + 16 && i < DF_RAW_BYTES; i += DF_THREADS) rb[i] = 0;
    }
    __syncthreads();
    // ---- pass 1: parse, count, keep the tokens ----
    uint32_t ntok = 0, entry = 0;
    for (uint32_t c0 = 0; c0 < n; c0 += DF_CHUNK) {
      const uint32_t clen = n - c0 < DF_CHUNK ? n - c0 : DF_CHUNK;
      parse_chunk(S, scan_tmp, mis, c0, clen, n, entry, tid);
      const uint8_t* rb = reinterpret_cast<const uint8_t*>(S.raw) + mis + c0;
      uint32_t tok[DF_ITEMS], flag[DF_ITEMS], idx[DF_ITEMS], cnt = 0;
#pragma unroll
      for (int k = 0; k < DF_ITEMS; ++k) {
        const uint32_t i = tid * DF_ITEMS + k;
        tok[k] = 0; flag[k] = 0;
        if (i < clen && S.mark[i]) {
          flag[k] = 1;
          const uint32_t l = S.mlen[i];
          if (l) {
            const uint32_t dist = S.mdist[i];
            uint32_t sym, eb, ev;
            len_symbol(l, sym, eb, ev);
            atomicAdd(&D.lfreq[sym], 1u);
            dist_symbol(dist, sym, eb, ev);
            atomicAdd(&D.dfreq[sym], 1u);
            tok[k] = 0x80000000u | ((l - 3) << 15) | (dist - 1);
          } else {
            tok[k] = rb[i];
            atomicAdd(&D.lfreq[tok[k]], 1u);
          }
          const uint32_t nx = i + (l ? l : 1u);
          if (nx >= clen) S.exit_rel = nx - clen;
        }
      }
      if (entry >= clen && tid == 0) S.exit_rel = entry - clen;
      Scan(scan_tmp).ExclusiveSum(flag, idx, cnt);
      __syncthreads();     // (the jump arrays are no longer read: stage may be written)
#pragma unroll
      for (int k = 0; k < DF_ITEMS; ++k) if (flag[k]) stage[idx[k]] = tok[k];
      __syncthreads();
      for (uint32_t i = tid; i < cnt; i += DF_THREADS) gtok[ntok + i] = stage[i];
      ntok += cnt;
      entry = DF_WALK ? 0u : S.exit_rel;
      __syncthreads();
    }
    // ---- the two codes ----
    huff_lengths(D.lfreq, 286, 15, B.llen, B, tid);
    if (tid >= 286 && tid < 288) B.llen[tid] = 0;
    __syncthreads();
    canonical_codes(B.llen, 286, B.lcode, B, tid);
    huff_lengths(D.dfreq, 30, 15, B.dlen, B, tid);
    if (tid >= 30 && tid < 32) B.dlen[tid] = 0;
    __syncthreads();
    canonical_codes(B.dlen, 30, B.dcode, B, tid);
    // header fields and the code-length code
    if (tid < 20) B.cfreq[tid] = 0;
    if (tid == 0) { B.hlit = 257; B.hdist = 1; B.cost_dyn = 0; B.cost_fix = 0; }
    __syncthreads();
    if (tid < 286 && B.llen[tid]) atomicMax(&B.hlit, tid + 1);
    if (tid < 30 && B.dlen[tid]) atomicMax(&B.hdist, tid + 1);
    __syncthreads();
    const uint32_t hlit = B.hlit, hdist = B.hdist;
    if (tid < hlit) atomicAdd(&B.cfreq[B.llen[tid]], 1u);
    if (tid < hdist) atomicAdd(&B.cfreq[B.dlen[tid]], 1u);
    __syncthreads();
    huff_lengths(B.cfreq, 19, 7, B.cllen, B, tid);
    canonical_codes(B.cllen, 19, B.ccode, B, tid);
    if (tid == 0) {
      uint32_t h = 4;
      for (uint32_t i = 0; i < 19; ++i) if (B.cllen[DF_CL_ORDER[i]]) h = i + 1 > h ? i + 1 : h;
      B.hclen = h;
    }
    // which is smaller: these codes + their header, or the fixed code?  (extra bits cost the same either way)
    {
      uint32_t dyn = 0, fix = 0;
      if (tid < 286) { const uint32_t f = D.lfreq[tid]; dyn += f * B.llen[tid]; fix += f * (tid < 144 ? 8u : tid < 256 ? 9u : tid < 280 ? 7u : 8u); }
      if (tid < 30) { const uint32_t f = D.dfreq[tid]; dyn += f * B.dlen[tid]; fix += f * 5u; }
      if (tid < 19) dyn += B.cfreq[tid] * B.cllen[tid];
      if (dyn) atomicAdd(&B.cost_dyn, dyn);
      if (fix) atomicAdd(&B.cost_fix, fix);
    }
    __syncthreads();
    const bool use_fixed = B.cost_fix + 3 <= B.cost_dyn + 17 + 3 * B.hclen;
    const uint32_t hclen = B.hclen;
    __syncthreads();
    if (use_fixed) {
      if (tid < 288) {
        const uint32_t L = tid < 144 ? 8u : tid < 256 ? 9u : tid < 280 ? 7u : 8u;
        const uint32_t c = tid < 144 ? 0x30u + tid : tid < 256 ? 0x190u + (tid - 144u) : tid < 280 ? tid - 256u : 0xC0u + (tid - 280u);
        B.llen[tid] = (uint8_t)L;
        B.lcode[tid] = (uint16_t)(__brev(c) >> (32 - L));
      }
      if (tid < 32) { B.dlen[tid] = 5; B.dcode[tid] = (uint16_t)(__brev(tid) >> 27); }
      __syncthreads();
    }
    // ---- the header ----
    DfOut O{reinterpret_cast<uint32_t*>(slots + (size_t)blk * DF_SLOT), 0u, 0u, 0u};
    {
      uint64_t bits[2] = {0, 0};
      uint32_t nb[2] = {0, 0};
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const uint32_t j = tid * 2 + k;
        if (use_fixed) {
          if (j == 0) { bits[k] = 3; nb[k] = 3; }
        } else if (j == 0) {
          bits[k] = 5u | ((hlit - 257u) << 3) | ((hdist - 1u) << 8) | ((hclen - 4u) << 13);   // BFINAL = 1, BTYPE = 10, HLIT, HDIST, HCLEN
          nb[k] = 17;
        } else if (j <= hclen) {
          bits[k] = B.cllen[DF_CL_ORDER[j - 1]];
          nb[k] = 3;
        } else if (j - 1 - hclen < hlit + hdist) {
          const uint32_t q = j - 1 - hclen;
          const uint32_t v = q < hlit ? B.llen[q] : B.dlen[q - hlit];
          bits[k] = B.ccode[v];
          nb[k] = B.cllen[v];
        }
      }
      emit_batch<Scan>(bits, nb, O, S, scan_tmp, tid);
    }
    // ---- pass 2: tokens -> bits ----
    for (uint32_t t0 = 0; t0 <= ntok; t0 += DF_TOK_BATCH) {
      uint64_t bits[2] = {0, 0};
      uint32_t nb[2] = {0, 0};
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const uint32_t t = t0 + tid * 2 + k;
        if (t < ntok) {
          const uint32_t tk = gtok[t];
          if (tk & 0x80000000u) {
            uint32_t sym, eb, ev;
            len_symbol(((tk >> 15) & 0xffu) + 3u, sym, eb, ev);
            uint64_t v = B.lcode[sym];
            uint32_t cb = B.llen[sym];
            v |= (uint64_t)ev << cb; cb += eb;
            dist_symbol((tk & 0x7fffu) + 1u, sym, eb, ev);
            v |= (uint64_t)B.dcode[sym] << cb; cb += B.dlen[sym];
            v |= (uint64_t)ev << cb; cb += eb;
            bits[k] = v; nb[k] = cb;
          } else {
            bits[k] = B.lcode[tk]; nb[k] = B.llen[tk];
          }
        } else if (t == ntok) {
          bits[k] = B.lcode[256]; nb[k] = B.llen[256];     // end of block
        }
      }
      emit_batch<Scan>(bits, nb, O, S, scan_tmp, tid);
    }
    if (tid == 0) {
      uint32_t bytes = 0;
      if (!S.overflow) {
        O.gout[O.wpos] = O.carry_word;
        bytes = O.wpos * 4 + (O.carry_bits + 7) / 8;
      }
      sizes[blk] = (S.overflow || bytes >= n + 5) ? 0u : bytes;
    }
    __syncthreads();
  }
}

// sizes[b] (DEFLATE bytes, 0 = stored) -> bytes of BGZF block b on the wire
__global__ void k6_block_bytes(const uint32_t* __restrict__ sizes, unsigned long long len, uint32_t n_blocks, unsigned long long* __restrict__ wire) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > n_blocks) return;
  if (b == n_blocks) { wire[b] = 0; return; }
  const unsigned long long o = (unsigned long long)b * BGZF_STORE_PAYLOAD;
  const uint32_t n = (uint32_t)(len - o < BGZF_STORE_PAYLOAD ? len - o : BGZF_STORE_PAYLOAD);
  wire[b] = 18 + (sizes[b] ? sizes[b] : 5 + n) + 8;
}

__constant__ uint8_t DF_BGZF_HDR[16] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0};
__constant__ uint8_t DF_EOF_MARKER[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43,
                                          0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};

// One warp per block: header, DEFLATE bytes out of the slot (or the stored form out of src), CRC32 + ISIZE.
__global__ void __launch_bounds__(256) k6_assemble(const uint8_t* __restrict__ src, unsigned long long len, const uint8_t* __restrict__ slots,
                                                   const uint32_t* __restrict__ sizes, const unsigned long long* __restrict__ woff,
                                                   const uint32_t* __restrict__ crc, uint8_t* __restrict__ out, uint32_t n_blocks, bool eof_marker) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < n_blocks; b += warps) {
    const unsigned long long o = (unsigned long long)b * BGZF_STORE_PAYLOAD;
    const uint32_t n = (uint32_t)(len - o < BGZF_STORE_PAYLOAD ? len - o : BGZF_STORE_PAYLOAD);
    const uint32_t sz = sizes[b];
    const uint32_t body = sz ? sz : 5 + n;
    uint8_t* w = out + woff[b];
    if (lane < 16) w[lane] = DF_BGZF_HDR[lane];
    const uint32_t bsize = 18 + body + 8 - 1;
    if (lane == 16) w[16] = (uint8_t)bsize;
    if (lane == 17) w[17] = (uint8_t)(bsize >> 8);
    uint8_t* d = w + 18;
    if (sz) {
      const uint8_t* sl = slots + (size_t)b * DF_SLOT;
      for (uint32_t i = lane; i < sz; i += 32) d[i] = sl[i];
    } else {
      if (lane == 0) { d[0] = 0x01; d[1] = (uint8_t)n; d[2] = (uint8_t)(n >> 8); d[3] = (uint8_t)~n; d[4] = (uint8_t)(~n >> 8); }
      for (uint32_t i = lane; i < n; i += 32) d[5 + i] = src[o + i];
    }
    if (lane < 8) {
      const uint32_t v = lane < 4 ? crc[b] : n;
      d[body + lane] = (uint8_t)(v >> (8 * (lane & 3u)));
    }
  }
  if (eof_marker && blockIdx.x == 0 && threadIdx.x < 28) out[woff[n_blocks] + threadIdx.x] = DF_EOF_MARKER[threadIdx.x];
}

}  // namespace

size_t bgzf_deflate_scratch_bytes(uint64_t len, int num_sms) {
  const uint64_t nb = (len + BGZF_STORE_PAYLOAD - 1) / BGZF_STORE_PAYLOAD;
  size_t cub_tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_tmp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)nb + 1);
  // slots | sizes u32 | wire u64 | woff u64 | cub temp | per-CTA token buffers (level >= 2)
  return align_up(nb * DF_SLOT, 256) + align_up(4 * (nb + 1), 256) + 2 * align_up(8 * (nb + 2), 256) + align_up(cub_tmp, 256) +
         align_up((size_t)num_sms * 2 * BGZF_STORE_PAYLOAD * 4, 256) + 256;
}

// Compresses src[0..len) into BGZF blocks at out (capacity >= the stored bound).  crc: per-payload CRC32 (launch_crc32_chunks).
// *out_len (host) is read back: synchronises.
cudaError_t bgzf_deflate_blocks(const uint8_t* src, uint64_t len, const uint32_t* crc, void* scratch, uint8_t* out, bool eof_marker, int level,
                                uint64_t* out_len, int num_sms, cudaStream_t s) {
  const uint64_t nb64 = (len + BGZF_STORE_PAYLOAD - 1) / BGZF_STORE_PAYLOAD;
  const uint32_t nb = (uint32_t)nb64;
  size_t cub_tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_tmp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)nb + 1);
  uint8_t* p = (uint8_t*)scratch;
  uint8_t* slots = p; p += align_up((size_t)nb * DF_SLOT, 256);
  uint32_t* sizes = (uint32_t*)p; p += align_up(4 * ((size_t)nb + 1), 256);
  unsigned long long* wire = (unsigned long long*)p; p += align_up(8 * ((size_t)nb + 2), 256);
  unsigned long long* woff = (unsigned long long*)p; p += align_up(8 * ((size_t)nb + 2), 256);
  void* tmp = p; p += align_up(cub_tmp, 256);
  uint32_t* tokbuf = (uint32_t*)p;
  if (nb) {
    const uint32_t grid = nb < (uint32_t)num_sms * 2 ? nb : (uint32_t)num_sms * 2;
    if (level >= 2) {
      const size_t smem = sizeof(DfDyn);
      cudaFuncSetAttribute(k6_deflate_dynamic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k6_deflate_dynamic<<<grid, DF_THREADS, smem, s>>>(src, len, slots, sizes, tokbuf, nb);
    } else {
      const size_t smem = sizeof(DfShared);
      cudaFuncSetAttribute(k6_deflate_fixed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k6_deflate_fixed<<<grid, DF_THREADS, smem, s>>>(src, len, slots, sizes, nb);
    }
  }
  k6_block_bytes<<<(nb + 1 + 255) / 256, 256, 0, s>>>(sizes, len, nb, wire);
  cub::DeviceScan::ExclusiveSum(tmp, cub_tmp, (const unsigned long long*)wire, woff, (int)nb + 1, s);
  k6_assemble<<<nb ? (nb < 148u * 8 ? (nb + 7) / 8 : 148 * 8) : 1, 256, 0, s>>>(src, len, slots, sizes, woff, crc, out, nb, eof_marker);
  unsigned long long total = 0;
  cudaMemcpyAsync(&total, woff + nb, 8, cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return e;
#ifdef DF_PROFILE
  {
    unsigned long long h[8] = {0}, z[8] = {0};
    cudaMemcpyFromSymbol(h, df_prof, sizeof h);
    cudaMemcpyToSymbol(df_prof, z, sizeof z);
    fprintf(stderr, "[k6 profile, CTA 0, cycles] runs %llu  candidates %llu  verify+extend %llu  parse %llu  bits %llu\n", h[0], h[1], h[2], h[3], h[4]);
  }
#endif
  *out_len = total + (eof_marker ? 28 : 0);
  return cudaGetLastError();
}

}  // namespace bamgpu
