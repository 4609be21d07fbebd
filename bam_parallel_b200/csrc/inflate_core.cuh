// inflate_core.cuh — per-stream raw-DEFLATE (RFC 1951) decoder used by kernel K1.
//
// Replaces the reference's `inflate_data` (bam_tools/src/util.rs:10-16, i.e. flate2's
// DeflateDecoder::read_to_end).  Every BGZF block is one independent stream, and a stream is worked on by a PAIR of
// lanes in two different warps of one CTA:
//   * the DECODER lane (Decoder<>) owns the bit reader and the Huffman tables, turns the bitstream into tokens
//     (up to 3 literals + one match, or 4 literals) and validates them (distances, output size);
//   * the COPIER lane (Copier) owns the output position, writes the literals and performs the LZ77 copies.
// They talk through a small token ring in shared memory (single producer / single consumer, no blocking waits: a lane
// whose ring is full or empty just skips the trip).  32 streams advance side by side inside a warp either way, so
// every issued instruction does useful work for up to 32 blocks, but the two halves are separate instruction streams:
// the scheduler overlaps the copiers' window-load latency with the decoders' ALU/shared-memory chains, each half
// needs roughly half the registers (twice the warps per SM), and each loop body is half as divergent.
// Design notes (see DESIGN.md "K1"):
//   * Huffman decode is canonical and LUT-free: the 15 left-aligned code-length limits of the
//     literal/length and distance codes live in REGISTERS, the code length is found by a 4-level
//     select search over them (no divergence between lanes, no memory access), and only one
//     index offset and the symbol come from shared memory.  Tables are 384 B per stream
//     (byte-wide sorted symbols; canonical order puts the symbols >= 256 of each code length
//     behind its literals, so one threshold per length replaces a per-symbol flag).
//   * Shared memory is interleaved by lane ([word][lane]) so any per-lane index pattern is
//     bank-conflict free.
//   * Input is fetched 16 bytes at a time with cp.async (LDGSTS) into two shared-memory slots
//     per lane, two chunks ahead of use, so neither the load latency nor a register scoreboard
//     sits on the decode dependency chain.
//   * Output goes through a 4-byte staging register and is stored as aligned 32-bit words.
//   * LZ77 copies are software-pipelined: a copy is only REGISTERED when its token is taken; its source words are
//     requested (aligned loads, no use) at the end of that trip and shifted into place and stored at the start of
//     the next one, at most 16 (distance < 35) or 32 bytes per trip in straight-line code; what is left is
//     requested again.  The window read is an L2/DRAM access (tens of thousands of streams x 32 KB of history do
//     not fit in L1), so this takes its latency off the critical path, and no lane ever loops over a load inside a
//     trip while the other 31 wait.  Distances < 8 (the dist=1 runs of BAM quality strings) are synthesised in
//     registers from the last 8 output bytes and continue as an ordinary copy at a multiple of the period.
//   * The DEFLATE block header / code-table build is a cold out-of-line function working on a
//     copy of the bit reader, so the hot loop's state stays in registers.  (It also has to be
//     out of line: with ptxas 12.9 -O1..-O3 the fully inlined form miscompiled on sm_100a —
//     the harness that reproduces it, tools/debug/k1_debug2.cu, is in the history at commit 95cc6bd.)
//
// Everything here is __host__ __device__ so tests/host_emul.cpp can run the identical logic on
// the CPU build box (no GPU there).  That emulation is a TEST of this code, not a fallback:
// the product library only ever launches the kernel.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define BG_HD __host__ __device__ __forceinline__
#define BG_HD_COLD __host__ __device__ __noinline__
#else
#define BG_HD inline
#define BG_HD_COLD inline
#endif

namespace bamgpu {

// Per-block status written by K1 (0 = ok).
enum : uint32_t {
  IST_OK = 0,
  IST_BAD_BTYPE = 1,
  IST_BAD_STORED = 2,
  IST_BAD_COUNTS = 3,     // HLIT > 286 or HDIST > 30
  IST_BAD_CODELEN = 4,    // code-length code: invalid symbol / repeat without previous / overflow
  IST_OVERSUBSCRIBED = 5,
  IST_NO_EOB = 6,
  IST_BAD_SYMBOL = 7,     // unassigned Huffman code, litlen 286/287, dist 30/31
  IST_BAD_DISTANCE = 8,   // distance reaches before the start of the block
  IST_OUTPUT_OVERRUN = 9, // more bytes than ISIZE
  IST_INPUT_OVERRUN = 10, // stream runs past its payload
  IST_SIZE_MISMATCH = 11, // final EOB reached with fewer bytes than ISIZE
};

// Shared-memory words per lane (each lane's words are interleaved with the other 31 lanes' words).
constexpr int LIT_SYM_WORDS = 72;   // 288 x u8: low 8 bits of the sorted literal/length symbols
constexpr int DIST_SYM_WORDS = 8;   // 32 x u8
// META, per code length L: bits 0-8 (first lit/len index - first code) mod 512, bits 9-13 the same for the distance
// code mod 32, bits 14-22 the index of the first lit/len symbol >= 256 of that length.  With -DBG_META_REGS=1 the 15
// words live in the decoder's registers and are picked by the same select network as the code-length limits: 64 B less
// shared memory per stream and one shared-memory round trip less per symbol, for 15 more selects per look-up.  Measured
// (8-token rings): 65.6 ms at 3 CTAs per SM against 63.8 with META in shared memory, 74.1 ms at 4 CTAs -> off by default.
#ifndef BG_META_REGS
#define BG_META_REGS 0
#endif
constexpr int META_WORDS = BG_META_REGS ? 0 : 16;
constexpr int LANE_WORDS = LIT_SYM_WORDS + DIST_SYM_WORDS + META_WORDS;  // 80 words = 320 B (96 = 384 B with META in shared memory)
constexpr int SM_LIT_SYM = 0;
constexpr int SM_DIST_SYM = LIT_SYM_WORDS;
constexpr int SM_META = SM_DIST_SYM + DIST_SYM_WORDS;

// Decoder lane states.  ST_BEGIN/ST_END/ST_FINISH/ST_STORED wait for room in the token ring to push their token.
enum LaneState : uint32_t { ST_NEXT = 0, ST_BEGIN, ST_HEADER, ST_DECODE, ST_STORED, ST_END, ST_FINISH, ST_DONE };

struct W4 { uint32_t x, y, z, w; };

#if defined(__CUDA_ARCH__)
// Copies a value through an opaque move so the compiler cannot keep re-reading its memory home (the limits
// come back from the out-of-line header parser through local memory and must live in registers afterwards).
BG_HD uint32_t bg_pin(uint32_t x) {
  uint32_t y;
  asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
BG_HD uint32_t bg_brev(uint32_t x) { return __brev(x); }
BG_HD uint32_t bg_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
BG_HD uint32_t bg_funnel_l(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_l(lo, hi, sh); }
BG_HD W4 bg_ld_in16(const void* p) {
  uint4 v = __ldg((const uint4*)p);
  return W4{v.x, v.y, v.z, v.w};
}
#else
inline uint32_t bg_pin(uint32_t x) { return x; }
inline uint32_t bg_brev(uint32_t x) {
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
  x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
  return (x >> 16) | (x << 16);
}
inline uint32_t bg_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
  sh &= 31;
  return sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
}
inline uint32_t bg_funnel_l(uint32_t lo, uint32_t hi, uint32_t sh) {
  sh &= 31;
  return sh ? ((hi << sh) | (lo >> (32 - sh))) : hi;
}
inline W4 bg_ld_in16(const void* p) {
  W4 v;
  memcpy(&v, p, 16);
  return v;
}
#endif

// Work description shared by all lanes.
struct InflateJob {
  const uint8_t* comp;       // compressed bytes, 16-byte aligned; readable 64 bytes past the last payload
  const uint64_t* in_off;    // per block: payload offset in comp (>= 16)
  const uint32_t* in_len;    // per block: payload bytes
  const uint64_t* out_off;   // per block: offset in out
  const uint32_t* isize;     // per block: expected inflated bytes
  uint8_t* out;              // inflated stream, 4-byte aligned; readable 16 bytes BEFORE out[0] and 4 past the end
  uint32_t* status;          // per block: IST_*
  uint32_t n_blocks;
};

// ---- input bit reader: 64-bit bit buffer fed from 16-byte chunks ------------------------------------
// On the GPU the chunks after the current one are staged in shared memory by cp.async (LDGSTS): two
// 16-byte slots per lane hold chunks k+1 and k+2 while chunk k is consumed from registers, so the
// global-memory latency of the input never sits on the decode dependency chain and no register
// scoreboard is tied up by a load in flight.  On the host (tests) chunks are read directly.
constexpr int STAGE_SLOT_BYTES = 32 * 16;              // one slot of one warp: 32 lanes x 16 bytes
constexpr int STAGE_WARP_BYTES = 2 * STAGE_SLOT_BYTES; // two slots

struct BitReader {
  const uint8_t* qp;         // next 16-byte chunk to request
  uint32_t q0, q1, q2, q3;   // current chunk; q0 is the next word to feed
  uint32_t qn;               // words left in q
  uint32_t stage;            // GPU: shared-memory byte address of the slot holding the next chunk
  uint32_t stage_x;          // slot 0 address XOR slot 1 address: stage ^= stage_x switches slots
  uint64_t bb;
  uint32_t nbits;
  uint32_t fed;              // payload bits moved into bb so far
  uint32_t in_bits;          // payload length in bits

#if defined(__CUDA_ARCH__)
  BG_HD void request(uint32_t slot_addr) {   // chunk at qp -> shared slot, asynchronously
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n\tcp.async.commit_group;" ::"r"(slot_addr), "l"(qp) : "memory");
    qp += 16;
  }
  BG_HD void next_chunk() {
    asm volatile("cp.async.wait_group 1;" ::: "memory");   // everything but the newest request has landed
    uint32_t x, y, z, w;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(stage) : "memory");
    q0 = x; q1 = y; q2 = z; q3 = w;
    request(stage);                       // the slot just read takes chunk k+3
    stage ^= stage_x;                     // the other slot holds the next chunk
  }
  BG_HD void begin_chunks(W4& first) {
    first = bg_ld_in16(qp);
    qp += 16;
    request(stage);
    request(stage ^ stage_x);
  }
#else
  BG_HD void next_chunk() {
    W4 v = bg_ld_in16(qp);
    qp += 16;
    q0 = v.x; q1 = v.y; q2 = v.z; q3 = v.w;
  }
  BG_HD void begin_chunks(W4& first) {
    first = bg_ld_in16(qp);
    qp += 16;
  }
#endif

  // `stage0`: shared-memory address of this lane's slot 0 (16-byte aligned); ignored on the host.
  BG_HD void begin(const uint8_t* p, uint32_t len, uint32_t stage0) {
    uintptr_t a = (uintptr_t)p;
    qp = (const uint8_t*)(a & ~(uintptr_t)15);
    stage = stage0;
    stage_x = stage0 ^ (stage0 + (uint32_t)STAGE_SLOT_BYTES);
    W4 q;
    begin_chunks(q);
    uint32_t lead = (uint32_t)(a & 15);
    uint32_t skipw = lead >> 2;
    q0 = q.x; q1 = q.y; q2 = q.z; q3 = q.w;
    if (skipw >= 2) { q0 = q2; q1 = q3; }
    if (skipw & 1) { q0 = q1; q1 = q2; q2 = q3; }
    qn = 4 - skipw;
    uint32_t sb = (lead & 3) * 8;
    bb = (uint64_t)(q0 >> sb);
    nbits = 32 - sb;
    fed = nbits;
    q0 = q1; q1 = q2; q2 = q3;
    --qn;
    in_bits = len * 8;
  }
  // After refill() at least 33 bits are valid in bb.
  BG_HD void refill() {
    if (nbits <= 32) {
      if (qn == 0) {
        next_chunk();
        qn = 4;
      }
      bb |= (uint64_t)q0 << nbits;
      nbits += 32;
      fed += 32;
      q0 = q1; q1 = q2; q2 = q3;
      --qn;
    }
  }
  BG_HD void consume(uint32_t n) { bb >>= n; nbits -= n; }
  BG_HD uint32_t bits_consumed() const { return fed - nbits; }
  // Call when done with a stream: no request may stay in flight into a slot the next stream will reuse.
  BG_HD void drain() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
  }
};

// ---- cold path: DEFLATE block header + code tables ----------------------------------------------
struct HeaderResult {
  uint32_t llim[16], dlim[16];  // [1..15]: exclusive limits, left-aligned to 15 bits
  uint32_t meta[16];            // [1..15]: the META words
  uint32_t state, err, bfinal, stored_len;
};

template <int STRIDE>
struct TableBuilder {
  uint32_t* sm;
  BG_HD uint32_t& smw(int w) { return sm[w * STRIDE]; }
  BG_HD void set_byte(int base, uint32_t i, uint32_t v) {
    uint32_t& w = smw(base + (int)(i >> 2));
    uint32_t sh = (i & 3) * 8;
    w = (w & ~(0xffu << sh)) | (v << sh);
  }
  // Builds the sorted-symbol table and the limits for `n` symbols whose code lengths are lens[0..n);
  // a[L] gets (index of the first symbol of length L) - (first code of length L), mod 2^16, so that
  // symbol index = a[L] + code.
  // Returns 0 or IST_*.
  BG_HD uint32_t build(const uint8_t* lens, uint32_t n, bool is_lit, uint32_t* lim, uint32_t* a, uint32_t* thr) {
    uint32_t cnt[16], cur[16];   // counts, then scatter cursors (cold path: local memory is fine)
    for (int L = 0; L < 16; ++L) cnt[L] = 0;
    for (uint32_t s = 0; s < n; ++s) cnt[lens[s]] += 1;
    uint32_t code = 0, offs = 0, bad = 0;
#pragma unroll
    for (int L = 1; L <= 15; ++L) {
      code <<= 1;
      if (code + cnt[L] > (1u << L)) bad = 1;
      lim[L] = (code + cnt[L]) << (15 - L);
      a[L] = (offs - code) & 0xffffu;
      cur[L] = offs;  // scatter cursor
      code += cnt[L];
      offs += cnt[L];
    }
    if (bad) return IST_OVERSUBSCRIBED;
    // Canonical order sorts by (length, symbol), so within one length every symbol >= 256 (end of block, length
    // codes) comes after every literal: one threshold index per length tells them apart.
    if (is_lit) {
      uint32_t nlit[16];
#pragma unroll
      for (int L = 0; L < 16; ++L) nlit[L] = 0;
      for (uint32_t s = 0; s < n && s < 256; ++s) {
        uint32_t l = lens[s];
#pragma unroll
        for (int L = 1; L <= 15; ++L) nlit[L] += (l == (uint32_t)L) ? 1u : 0u;
      }
#pragma unroll
      for (int L = 1; L <= 15; ++L) thr[L] = cur[L] + nlit[L];   // cursor == first index of length L here
    }
    for (uint32_t s = 0; s < n; ++s) {
      uint32_t l = lens[s];
      if (l) {
        uint32_t pos = cur[l];
        cur[l] = pos + 1;
        if (is_lit) {
          set_byte(SM_LIT_SYM, pos, s & 0xffu);          // >= 256: 0 = end of block, 1..29 = length codes 257..285
        } else {
          set_byte(SM_DIST_SYM, pos, s);
        }
      }
    }
    return IST_OK;
  }
};

// Parses one DEFLATE block header (and its code tables) from *brp into shared memory / *hr.
template <int STRIDE>
BG_HD_COLD void parse_header(BitReader* brp, uint32_t* sm, HeaderResult* hr) {
  BitReader br = *brp;
  uint8_t lens[320];
  hr->err = 0;
  hr->stored_len = 0;
  hr->bfinal = 0;
  for (int k = 0; k < 16; ++k) hr->llim[k] = hr->dlim[k] = hr->meta[k] = 0;
#define BG_FAIL(code) do { hr->err = (code); hr->state = ST_NEXT; *brp = br; return; } while (0)
  br.refill();
  if (br.bits_consumed() + 3 > br.in_bits) BG_FAIL(IST_INPUT_OVERRUN);
  hr->bfinal = (uint32_t)br.bb & 1u;
  uint32_t btype = ((uint32_t)br.bb >> 1) & 3u;
  br.consume(3);
  if (btype == 0) {
    br.consume(br.nbits & 7);  // to the next byte boundary
    br.refill();
    uint32_t len = (uint32_t)br.bb & 0xffffu, nlen = ((uint32_t)br.bb >> 16) & 0xffffu;
    br.consume(32);
    if (len != (nlen ^ 0xffffu)) BG_FAIL(IST_BAD_STORED);
    if (br.bits_consumed() > br.in_bits) BG_FAIL(IST_INPUT_OVERRUN);
    hr->stored_len = len;
    hr->state = len ? ST_STORED : (hr->bfinal ? ST_NEXT : ST_HEADER);
    *brp = br;
    return;
  }
  if (btype == 3) BG_FAIL(IST_BAD_BTYPE);
  uint32_t hlit, hdist;
  if (btype == 1) {
    hlit = 288; hdist = 30;
    for (int i = 0; i < 144; ++i) lens[i] = 8;
    for (int i = 144; i < 256; ++i) lens[i] = 9;
    for (int i = 256; i < 280; ++i) lens[i] = 7;
    for (int i = 280; i < 288; ++i) lens[i] = 8;
    for (int i = 288; i < 318; ++i) lens[i] = 5;
  } else {
    br.refill();
    hlit = ((uint32_t)br.bb & 31u) + 257;
    hdist = (((uint32_t)br.bb >> 5) & 31u) + 1;
    uint32_t hclen = (((uint32_t)br.bb >> 10) & 15u) + 4;
    br.consume(14);
    if (hlit > 286 || hdist > 30) BG_FAIL(IST_BAD_COUNTS);
    // code-length code lengths, 3 bits each, sent in the RFC's permuted order
    // 16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15; packed here 3 bits per symbol.
    uint64_t cl = 0;
    for (uint32_t i = 0; i < hclen; ++i) {
      br.refill();
      uint32_t l = (uint32_t)br.bb & 7u;
      br.consume(3);
      uint32_t sym;
      if (i < 3) sym = 16 + i;
      else {
        uint32_t j = (i - 3) >> 1;                       // odd i: 0,7,6,..,1   even i: 8,9,..,15
        sym = (i & 1) ? (j == 0 ? 0u : 8u - j) : 8u + j;
      }
      cl |= (uint64_t)l << (3 * sym);
    }
    // canonical code for the 19-symbol alphabet (max length 7), kept in registers
    uint32_t ccnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t s = 0; s < 19; ++s) {
      uint32_t l = (uint32_t)(cl >> (3 * s)) & 7u;
#pragma unroll
      for (int L = 1; L <= 7; ++L) ccnt[L] += (l == (uint32_t)L) ? 1u : 0u;
    }
    uint32_t clim[8], cbase[8], coffs[8];
    {
      uint32_t code = 0, offs = 0, bad = 0;
#pragma unroll
      for (int L = 1; L <= 7; ++L) {
        code <<= 1;
        if (code + ccnt[L] > (1u << L)) bad = 1;
        cbase[L] = code << (7 - L);
        clim[L] = (code + ccnt[L]) << (7 - L);
        coffs[L] = offs;
        code += ccnt[L];
        offs += ccnt[L];
      }
      if (bad) BG_FAIL(IST_OVERSUBSCRIBED);
    }
    // sorted symbols of the code-length code, 5 bits each, in a 128-bit register pair
    uint64_t cs_lo = 0, cs_hi = 0;
    {
      uint32_t cur[8];
#pragma unroll
      for (int L = 1; L <= 7; ++L) cur[L] = coffs[L];
      for (uint32_t s = 0; s < 19; ++s) {
        uint32_t l = (uint32_t)(cl >> (3 * s)) & 7u;
        if (l) {
          uint32_t pos = 0;
#pragma unroll
          for (int L = 1; L <= 7; ++L)
            if (l == (uint32_t)L) { pos = cur[L]; cur[L] = pos + 1; }
          if (pos < 12) cs_lo |= (uint64_t)s << (5 * pos);
          else cs_hi |= (uint64_t)s << (5 * (pos - 12));
        }
      }
    }
    uint32_t total = hlit + hdist, i = 0, prev = 0;
    while (i < total) {
      br.refill();
      uint32_t v7 = bg_brev((uint32_t)br.bb) >> 25;
      uint32_t L = 1;
#pragma unroll
      for (int k = 1; k <= 6; ++k) L += (v7 >= clim[k]) ? 1u : 0u;
      if (v7 >= clim[7]) BG_FAIL(IST_BAD_CODELEN);
      uint32_t base = 0, offs = 0;
#pragma unroll
      for (int k = 1; k <= 7; ++k)
        if (L == (uint32_t)k) { base = cbase[k]; offs = coffs[k]; }
      uint32_t idx = offs + ((v7 - base) >> (7 - L));
      uint32_t sym = idx < 12 ? (uint32_t)(cs_lo >> (5 * idx)) & 31u : (uint32_t)(cs_hi >> (5 * (idx - 12))) & 31u;
      br.consume(L);
      if (sym < 16) {
        lens[i++] = (uint8_t)sym;
        prev = sym;
      } else {
        uint32_t rep, val;
        if (sym == 16) {
          if (i == 0) BG_FAIL(IST_BAD_CODELEN);
          rep = 3 + ((uint32_t)br.bb & 3u); br.consume(2); val = prev;
        } else if (sym == 17) {
          rep = 3 + ((uint32_t)br.bb & 7u); br.consume(3); val = 0; prev = 0;
        } else {
          rep = 11 + ((uint32_t)br.bb & 127u); br.consume(7); val = 0; prev = 0;
        }
        if (i + rep > total) BG_FAIL(IST_BAD_CODELEN);
        for (uint32_t r = 0; r < rep; ++r) lens[i++] = (uint8_t)val;
      }
    }
    if (lens[256] == 0) BG_FAIL(IST_NO_EOB);
    if (br.bits_consumed() > br.in_bits) BG_FAIL(IST_INPUT_OVERRUN);
  }
  TableBuilder<STRIDE> tb{sm};
  uint32_t al[16], ad[16], thr[16];
  uint32_t e = tb.build(lens + hlit, hdist, false, hr->dlim, ad, thr);
  if (!e) e = tb.build(lens, hlit, true, hr->llim, al, thr);
  if (e) BG_FAIL(e);
  hr->meta[0] = 0;
  for (int L = 1; L <= 15; ++L) {
    hr->meta[L] = (al[L] & 0x1ffu) | ((ad[L] & 31u) << 9) | ((thr[L] & 0x1ffu) << 14);
    if (!BG_META_REGS) sm[(SM_META + L) * STRIDE] = hr->meta[L];
  }
  hr->state = ST_DECODE;
  *brp = br;
#undef BG_FAIL
}

// ---- token ring between a decoder lane and its copier lane ------------------------------------------
// A token is two 32-bit words, written and read with ONE 64-bit shared-memory access:
//   w0: bits 0-23 the first three literal bytes, bits 24-26 the literal count (0..4), bits 27-29 the kind, bit 31 the phase
//   w1: TK_LITS / TK_MATCH  (len - 3) | (dist - 1) << 8 | fourth literal << 23
//       TK_BEGIN block index     TK_END final IST_* status of the block
//       TK_RAW   len | offset in the payload << 16 (stored block: bytes copied straight from the input)
// Single producer, single consumer, nobody ever blocks on it.  Token i lives in slot i mod RING_TOKENS and carries
// phase (i / RING_TOKENS) & 1, so the consumer sees a token appear by polling the slot itself: there is no head
// counter whose store would have to be ordered behind the slot's (st.release = MEMBAR.ALL.CTA, which was measured
// to wait for the copier's global stores every trip).  The consumer publishes its count with a plain store after
// the read; shared-memory accesses of one warp are performed in issue order.
enum : uint32_t { TK_LITS = 0, TK_MATCH = 1, TK_BEGIN = 2, TK_END = 3, TK_RAW = 4, TK_DONE = 5 };
#ifndef BG_RING_SHIFT
#define BG_RING_SHIFT 3
#endif
constexpr uint32_t RING_SHIFT = BG_RING_SHIFT;
constexpr uint32_t RING_TOKENS = 1u << RING_SHIFT;        // per stream
// The ring lives in shared memory ([slot][lane], 8 B per lane) or, with -DBG_RING_GLOBAL=1, in global memory (128 B per
// stream: 8-byte slots, then the tail counter; volatile accesses, i.e. served by L2).  The global form frees 68 B of
// shared memory per stream - the next smaller carve-out, 32 KB more L1 per SM for the copiers' window loads - for
// ~300 cycles of latency per poll, which only the copiers (not the bottleneck) wait for: the decoder reads the tail
// one trip ahead of using it.
#ifndef BG_RING_GLOBAL
#define BG_RING_GLOBAL 0
#endif
constexpr int RING_ROW_BYTES = 32 * 8;                    // shared form: one slot of every lane of a warp
constexpr int RING_WARP_BYTES = BG_RING_GLOBAL ? 0 : (int)(RING_TOKENS * RING_ROW_BYTES + 32 * 4);   // slots + the row of tail counters
constexpr int RING_GLOBAL_STREAM_BYTES = RING_TOKENS * 8 + 8 <= 128 ? 128 : (RING_TOKENS * 8 + 8 <= 256 ? 256 : (RING_TOKENS * 8 + 8 <= 512 ? 512 : 1024));   // global form: whole lines per stream

struct TokRing {
  BG_HD static uint32_t phase(uint32_t idx) { return ((idx >> RING_SHIFT) & 1u) << 31; }
#if defined(__CUDACC__)
#if BG_RING_GLOBAL
  uint8_t* g;       // this stream's 128-byte line: RING_TOKENS 8-byte slots, then the tail counter
  __device__ __forceinline__ void init_global(uint8_t* stream_line) { g = stream_line; }
#else
  uint32_t slots;   // shared-memory byte address of this lane's slot 0 (slot stride RING_ROW_BYTES)
  uint32_t ctr;     // ... of this lane's tail counter
  __device__ __forceinline__ void init(uint32_t warp_base, uint32_t lane) {
    slots = warp_base + lane * 8;
    ctr = warp_base + RING_TOKENS * RING_ROW_BYTES + lane * 4;
  }
#endif
#if defined(__CUDA_ARCH__)
#if BG_RING_GLOBAL
  BG_HD void reset() {
    for (uint32_t i = 0; i < RING_TOKENS; ++i)
      asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(g + 8 * i), "r"(0x80000000u), "r"(0u));
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(g + 8 * RING_TOKENS), "r"(0u));
  }
  BG_HD uint32_t tail() const { uint32_t v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(g + 8 * RING_TOKENS)); return v; }
  BG_HD void push(uint32_t idx, uint32_t w0, uint32_t w1) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(g + 8 * (idx & (RING_TOKENS - 1))), "r"(w0 | phase(idx)), "r"(w1));
  }
  BG_HD bool peek(uint32_t idx, uint32_t& w0, uint32_t& w1) const {
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(w0), "=r"(w1) : "l"(g + 8 * (idx & (RING_TOKENS - 1))));
    return ((w0 ^ phase(idx)) >> 31) == 0;
  }
  BG_HD void set_tail(uint32_t v) { asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(g + 8 * RING_TOKENS), "r"(v)); }
#else
  BG_HD void reset() {   // every slot holds the phase the first lap does NOT use
    for (uint32_t i = 0; i < RING_TOKENS; ++i)
      asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(slots + i * RING_ROW_BYTES), "r"(0x80000000u), "r"(0u));
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(ctr), "r"(0u));
  }
  BG_HD uint32_t tail() const { uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(ctr)); return v; }
  BG_HD void push(uint32_t idx, uint32_t w0, uint32_t w1) {   // idx = tokens pushed so far
    asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(slots + (idx & (RING_TOKENS - 1)) * RING_ROW_BYTES), "r"(w0 | phase(idx)), "r"(w1));
  }
  // token number idx if it has been pushed
  BG_HD bool peek(uint32_t idx, uint32_t& w0, uint32_t& w1) const {
    asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w0), "=r"(w1) : "r"(slots + (idx & (RING_TOKENS - 1)) * RING_ROW_BYTES));
    return ((w0 ^ phase(idx)) >> 31) == 0;
  }
  BG_HD void set_tail(uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(ctr), "r"(v)); }
#endif
#else   // host pass of nvcc: never called
  BG_HD void reset() {}
  BG_HD uint32_t tail() const { return 0; }
  BG_HD void push(uint32_t, uint32_t, uint32_t) {}
  BG_HD bool peek(uint32_t, uint32_t&, uint32_t&) const { return false; }
  BG_HD void set_tail(uint32_t) {}
#endif
#else
  uint32_t* mem;    // RING_TOKENS * 2 words of slots, then the tail counter
  BG_HD void reset() {
    for (uint32_t i = 0; i < RING_TOKENS; ++i) { mem[2 * i] = 0x80000000u; mem[2 * i + 1] = 0; }
    mem[2 * RING_TOKENS] = 0;
  }
  BG_HD uint32_t tail() const { return mem[2 * RING_TOKENS]; }
  BG_HD void push(uint32_t idx, uint32_t w0, uint32_t w1) {
    mem[2 * (idx & (RING_TOKENS - 1))] = w0 | phase(idx); mem[2 * (idx & (RING_TOKENS - 1)) + 1] = w1;
  }
  BG_HD bool peek(uint32_t idx, uint32_t& w0, uint32_t& w1) const {
    w0 = mem[2 * (idx & (RING_TOKENS - 1))]; w1 = mem[2 * (idx & (RING_TOKENS - 1)) + 1];
    return ((w0 ^ phase(idx)) >> 31) == 0;
  }
  BG_HD void set_tail(uint32_t v) { mem[2 * RING_TOKENS] = v; }
#endif
};

// ---- the decoder lane --------------------------------------------------------------------------------
template <int STRIDE>  // shared-memory word stride between consecutive table entries of one lane (32 on GPU)
struct Decoder {
  uint32_t* sm;          // this lane's shared-memory table column (generic pointer: the cold table builder)
  uint32_t sma;          // GPU: the same as a shared-window byte address (the hot look-ups)
  BitReader br;
  TokRing ring;
  uint32_t head;         // tokens pushed so far
  uint32_t tail_seen;    // BG_RING_GLOBAL: the copier's count as of the previous trip
  uint32_t state, bfinal, err, blk, cp_rem;
  uint32_t produced, isize;     // bytes the tokens pushed so far stand for / ISIZE of the block
  uint32_t lpk[8], dpk[8];   // code-length limits of the literal/length and distance codes, packed in pairs (code_len_packed); registers
  uint32_t mreg[16];         // BG_META_REGS: the META words, [1..15] (+ [0] for an invalid code); registers

  // ---------------- shared-memory accessors ----------------
#if defined(__CUDA_ARCH__)
  // 32-bit shared-window addresses: one multiply-add per access instead of 64-bit generic pointer arithmetic.
  // Ordinary loads (the compiler folds the cvta into LDS), so they are ordered against the table builder's stores
  // and free to be scheduled early; volatile asm loads were measured to triple the decoders' shared-memory stalls.
  BG_HD uint32_t smr(uint32_t w) const {
    return *(const uint32_t*)__cvta_shared_to_generic(sma + w * (4u * STRIDE));
  }
  BG_HD uint32_t smb(uint32_t base, uint32_t i) const {   // byte i of the byte array starting at word `base`
    return *(const uint8_t*)__cvta_shared_to_generic(sma + (base + (i >> 2)) * (4u * STRIDE) + (i & 3u));
  }
#else
  BG_HD uint32_t smr(uint32_t w) const { return sm[w * STRIDE]; }
  BG_HD uint32_t smb(uint32_t base, uint32_t i) const {
    return ((const uint8_t*)(sm + (base + (i >> 2)) * STRIDE))[i & 3];
  }
#endif

  // ---------------- canonical Huffman ----------------
  // 1 + number of k in 1..15 with v >= lim[k] (16 = v is beyond the last code: invalid).  The limits are exclusive
  // upper bounds of the codes of length <= k, left-aligned, hence non-decreasing in k: the count is found by a
  // 4-level binary search whose candidates are picked with selects as soon as each comparison is known
  // (4 compares + 11 selects, against 14 compare-and-add pairs for the linear form).
  BG_HD static uint32_t code_len(const uint32_t* lim, uint32_t v) {
    const bool p3 = v >= lim[8];
    const uint32_t a = p3 ? lim[12] : lim[4];
    const uint32_t b2 = p3 ? lim[10] : lim[2], b6 = p3 ? lim[14] : lim[6];
    const uint32_t c1 = p3 ? lim[9] : lim[1], c3 = p3 ? lim[11] : lim[3], c5 = p3 ? lim[13] : lim[5], c7 = p3 ? lim[15] : lim[7];
    const bool p2 = v >= a;
    const uint32_t b = p2 ? b6 : b2;
    const uint32_t d1 = p2 ? c5 : c1, d3 = p2 ? c7 : c3;
    const bool p1 = v >= b;
    const uint32_t d = p1 ? d3 : d1;
    const bool p0 = v >= d;
    return 1u + (p3 ? 8u : 0u) + (p2 ? 4u : 0u) + (p1 ? 2u : 0u) + (p0 ? 1u : 0u);
  }

  // The same search over limits packed two to a register (lim[k] | lim[k + 8] << 16 in pk[k], k = 1..7; pk[0] =
  // lim[8]): the first comparison picks the halves with one byte permute each.  Used for the distance code, which is
  // decoded once per match: 8 registers instead of 15 for one more instruction.
  // `mr` ([1..15], BG_META_REGS): the META word of the found length is picked by the same four comparisons — 8 selects on
  // the first one (known long before the last), then 4, 2, 1 — and returned in m; [0] serves the invalid "length 16".
  BG_HD static uint32_t code_len_packed(const uint32_t* pk, const uint32_t* mr, uint32_t v, uint32_t& m) {
    const bool p3 = v >= pk[0];
#if defined(__CUDA_ARCH__)
    const uint32_t sel = p3 ? 0x4432u : 0x4410u;
#define BG_HALF(x) __byte_perm((x), 0u, sel)
#else
#define BG_HALF(x) (p3 ? (x) >> 16 : (x) & 0xffffu)
#endif
    const uint32_t a = BG_HALF(pk[4]);
    const uint32_t b2 = BG_HALF(pk[2]), b6 = BG_HALF(pk[6]);
    const uint32_t c1 = BG_HALF(pk[1]), c3 = BG_HALF(pk[3]), c5 = BG_HALF(pk[5]), c7 = BG_HALF(pk[7]);
#undef BG_HALF
    const bool p2 = v >= a;
    const uint32_t b = p2 ? b6 : b2;
    const uint32_t d1 = p2 ? c5 : c1, d3 = p2 ? c7 : c3;
    const bool p1 = v >= b;
    const uint32_t d = p1 ? d3 : d1;
    const bool p0 = v >= d;
#if BG_META_REGS
    // length = 1 + 8 p3 + 4 p2 + 2 p1 + p0: candidates by p3 first
    const uint32_t e1 = p3 ? mr[9] : mr[1], e2 = p3 ? mr[10] : mr[2], e3 = p3 ? mr[11] : mr[3], e4 = p3 ? mr[12] : mr[4];
    const uint32_t e5 = p3 ? mr[13] : mr[5], e6 = p3 ? mr[14] : mr[6], e7 = p3 ? mr[15] : mr[7], e8 = p3 ? mr[0] : mr[8];
    const uint32_t f1 = p2 ? e5 : e1, f2 = p2 ? e6 : e2, f3 = p2 ? e7 : e3, f4 = p2 ? e8 : e4;
    const uint32_t g1 = p1 ? f3 : f1, g2 = p1 ? f4 : f2;
    m = p0 ? g2 : g1;
#else
    (void)mr;
    m = 0;
#endif
    return 1u + (p3 ? 8u : 0u) + (p2 ? 4u : 0u) + (p1 ? 2u : 0u) + (p0 ? 1u : 0u);
  }

  // A "token" is up to LITS_PER_TOKEN - 1 leading literals plus one closing symbol (literal, match, end of block or
  // an invalid code).  Decoding several literals per trip amortises the sparse match code over more output: most
  // lanes start with a literal, and a lane that sees a length stops there.
#ifndef BG_LITS_PER_TOKEN
#define BG_LITS_PER_TOKEN 4
#endif
  static constexpr int LITS_PER_TOKEN = BG_LITS_PER_TOKEN;
  struct Tok { uint32_t lits, nlit, len, dist, kind; };

  // One literal/length symbol.  Needs 15 valid bits in br.bb; consumes the code.  Returns 0 literal (value in e),
  // 1 length code (e = code + 1), 2 end of block, 3 invalid.
  BG_HD uint32_t decode_litlen(uint32_t& e) {
    uint32_t c = bg_brev((uint32_t)br.bb) >> 17;
    uint32_t m;
    uint32_t L = code_len_packed(lpk, mreg, c, m);
    const bool bad = L > 15u;          // c >= llim[15]: no such code
    L &= 15u;                          // (0 then: nothing consumed, the table reads stay in range)
#if !BG_META_REGS
    m = smr(SM_META + L);
#endif
    uint32_t idx = (m + (c >> (15 - L))) & 0x1ffu;
    uint32_t spc = idx >= ((m >> 14) & 0x1ffu) ? 1u : 0u;
    idx = idx < 287u ? idx : 287u;
    e = smb(SM_LIT_SYM, idx);
    br.consume(L);
    uint32_t kind = spc ? (e == 0 ? 2u : 1u) : 0u;
    if (bad) kind = 3;
    return kind;
  }

  // Huffman-decodes the next token: touches only the bitstream and the tables.
  //   t.nlit literals packed in t.lits, then t.kind 1: match (t.len, t.dist)   2: end of block   3: invalid code
  //   4: nothing (the trip ended on a literal, already packed)
  BG_HD void decode_token(Tok& t) {
    br.refill();   // >= 33 bits: enough for two codes (<= 15 bits each) without another refill
    t.lits = 0;
    t.nlit = 0;
    t.len = 0;
    t.dist = 0;
    uint32_t e;
    uint32_t kind = decode_litlen(e);
#pragma unroll
    for (int i = 1; i < LITS_PER_TOKEN; ++i) {
      if (kind == 0) {
        if (i == 2) br.refill();   // two codes per refill; the fourth symbol still has >= 18 bits
        t.lits |= e << (8 * t.nlit);
        t.nlit += 1;
        kind = decode_litlen(e);
      }
    }
    if (kind == 0) {   // a closing literal simply joins the leading ones: one token for up to LITS_PER_TOKEN bytes
      t.lits |= e << (8 * t.nlit);
      t.nlit += 1;
      kind = 4;
    }
    t.kind = kind;
    if (kind == 1) {
      uint32_t lc = e - 1;   // length code 0..28 (29, 30: symbols 286/287, invalid)
      if (lc > 28) t.kind = 3;
      uint32_t eb = lc < 8 ? 0u : (lc == 28 ? 0u : (lc - 4) >> 2);
      uint32_t base = lc < 8 ? 3u + lc : (lc == 28 ? 258u : 3u + ((4u + (lc & 3u)) << eb));
      eb &= 7u;
      br.refill();
      t.len = base + ((uint32_t)br.bb & ((1u << eb) - 1u));
      br.consume(eb);
      uint32_t c = bg_brev((uint32_t)br.bb) >> 17;   // >= 28 bits left: the distance code and its <= 13 extra bits fit
      uint32_t m;
      uint32_t L = code_len_packed(dpk, mreg, c, m);
      if (L > 15u) t.kind = 3;
      L &= 15u;
#if !BG_META_REGS
      m = smr(SM_META + L);
#endif
      uint32_t di = ((m >> 9) + (c >> (15 - L))) & 31u;
      uint32_t ds = smb(SM_DIST_SYM, di);
      br.consume(L);
      if (ds >= 30) t.kind = 3;
      uint32_t deb = ds < 2 ? 0u : (ds >> 1) - 1u;
      deb &= 15u;
      t.dist = (ds < 2 ? ds + 1u : 1u + ((2u | (ds & 1u)) << deb)) + ((uint32_t)br.bb & ((1u << deb) - 1u));
      br.consume(deb);
    }
  }

  BG_HD void push(uint32_t w0, uint32_t w1) { ring.push(head, w0, w1); head += 1; }

  // Everything a trip does outside the Huffman decode: the tokens that open and close a block, stored blocks.
  BG_HD void cold_trip(const InflateJob& job, uint32_t stage0) {
    if (state == ST_BEGIN) {
      push(TK_BEGIN << 27, blk);
      state = ST_HEADER;
    } else if (state == ST_STORED) {
      // stored block: the copier takes the bytes straight from the input; skip them here
      const uint32_t off = pay_skipped + (br.bits_consumed() >> 3), len = cp_rem, in_len = job.in_len[blk];
      if (off + len > in_len) { err = IST_INPUT_OVERRUN; state = ST_END; return; }
      if (len > isize - produced) { err = IST_OUTPUT_OVERRUN; state = ST_END; return; }
      push(TK_RAW << 27, len | (off << 16));
      produced += len;
      br.drain();
      br.begin(job.comp + job.in_off[blk] + off + len, in_len - (off + len), stage0);
      pay_skipped = off + len;
      state = bfinal ? ST_END : ST_HEADER;
    } else if (state == ST_END) {
      if (!err) {
        if (produced != isize) err = IST_SIZE_MISMATCH;
        else if (br.bits_consumed() > br.in_bits) err = IST_INPUT_OVERRUN;
      }
      br.drain();
      push(TK_END << 27, err);
      state = ST_NEXT;
    } else if (state == ST_FINISH) {
      push(TK_DONE << 27, 0);
      state = ST_DONE;
    }
  }
  uint32_t pay_skipped;   // payload bytes before the position the bit reader was last (re)started at

  // One trip: at most one token.  A lane whose ring is full just skips it.  Returns false if the lane had nothing
  // to do (ring full, or waiting for the warp-level part of the loop: next block, header).
  BG_HD bool trip(const InflateJob& job, uint32_t stage0) {
    if (state == ST_HEADER || state == ST_NEXT || state == ST_DONE) return false;
    // (Re-reading the copier's count only when the ring looks full was measured 18 % SLOWER: the extra branch splits
    // the warp and the decode below then runs once per half.)
#if BG_RING_GLOBAL
    {   // the copier's count as read one trip ago (its L2 latency is off the chain); stale = conservative
      const uint32_t seen = tail_seen;
      tail_seen = ring.tail();
      if (head - seen >= RING_TOKENS) return false;
    }
#else
    if (head - ring.tail() >= RING_TOKENS) return false;
#endif
    if (state == ST_DECODE) {
      Tok t;
#if BG_DIAG_NULL_DECODER   // diagnostic only (wrong output): a fixed token mix without any Huffman work -> the copiers' own pace
      {
        const uint32_t left = isize - produced, k = head & 7u;
        t.lits = 0x41424344u; t.nlit = (k & 1u) ? 1u : ((k & 2u) ? 0u : 3u); t.len = 0; t.dist = 0; t.kind = 4;
        if (k != 5u) { t.kind = 1; t.len = 3u + ((head * 7u) & 15u); t.dist = 8u + (((head * 2654435761u) >> 17) & 0x7ff0u); }
        if (t.dist > produced + t.nlit) { t.kind = 4; t.nlit = 3; }
        if (t.nlit + (t.kind == 1 ? t.len : 0u) >= left) { t.kind = 2; t.nlit = left > 4u ? 4u : left; if (t.nlit < left) t.kind = 4; }
      }
#else
      decode_token(t);
#endif
      uint32_t w0 = (t.lits & 0xffffffu) | (t.nlit << 24), w1 = (t.lits >> 24) << 23;
      uint32_t np = produced + t.nlit, bad = 0;
      bool any = t.nlit != 0;
      if (t.kind == 1) {
        if (t.dist > np) bad = IST_BAD_DISTANCE;
        np += t.len;
        w0 |= TK_MATCH << 27;
        w1 |= (t.len - 3u) | ((t.dist - 1u) << 8);
        any = true;
      }
      if (t.kind == 3) bad = IST_BAD_SYMBOL;
      if (!bad && np > isize) bad = IST_OUTPUT_OVERRUN;
      if (bad) {
        err = bad;
        state = ST_END;
      } else {
        produced = np;
        if (t.kind == 2) state = bfinal ? ST_END : ST_HEADER;
        if (any) push(w0, w1);
      }
    } else {
      cold_trip(job, stage0);
    }
    return true;
  }

  BG_HD void header() {
    HeaderResult hr;
    BitReader tmp = br;   // only this copy's address escapes: the lane itself stays in registers
    parse_header<STRIDE>(&tmp, sm, &hr);
    br = tmp;
    bfinal = hr.bfinal;
    cp_rem = hr.stored_len;
    state = hr.state == ST_NEXT ? (uint32_t)ST_END : hr.state;   // parse_header: ST_NEXT = this was the last block (or an error)
    if (hr.err) { err = hr.err; state = ST_END; }
    {   // unconditional: the limits must not be live across the call above (ptxas then keeps them in local memory)
      lpk[0] = bg_pin(hr.llim[8]);
      dpk[0] = bg_pin(hr.dlim[8]);
#pragma unroll
      for (int k = 1; k <= 7; ++k) {
        lpk[k] = bg_pin(hr.llim[k] | (hr.llim[k + 8] << 16));
        dpk[k] = bg_pin(hr.dlim[k] | (hr.dlim[k + 8] << 16));
      }
#if BG_META_REGS
#pragma unroll
      for (int k = 0; k <= 15; ++k) mreg[k] = bg_pin(hr.meta[k]);
#endif
    }
  }

  // `stage0`: shared-memory byte address of this lane's input staging slot 0 (GPU only).
  BG_HD void begin_block(const InflateJob& job, uint32_t b, uint32_t stage0) {
    blk = b;
    err = 0;
    bfinal = 0;
    produced = 0;
    isize = job.isize[b];
    pay_skipped = 0;
    br.begin(job.comp + job.in_off[b], job.in_len[b], stage0);
    state = ST_BEGIN;
  }
};

// ---- the copier lane ---------------------------------------------------------------------------------
struct Copier {
  TokRing ring;
  uint32_t tail;         // tokens taken so far
  uint32_t done;         // TK_DONE seen
  uint32_t blk;
  // ---- output: byte offsets relative to ob (the 4-byte aligned address at or below the block start) ----
  uint8_t* ob;
  uint32_t head;         // first own byte (0..3)
  uint32_t opos, oend;   // next byte to write / one past the last own byte
  uint32_t sreg;         // bytes of the word at opos & ~3 not yet stored
  uint32_t full_lo, full_span;  // aligned words a with (a - full_lo) < full_span lie wholly inside the block
  uint32_t fast_span;           // ... and with (a - full_lo) < fast_span so do the three words after a
  // ---- copy in flight: its source words were requested one trip ago ----
  uint32_t p_len, p_dist, p_n, p_ssh;   // bytes left, distance (0: from the input, stored block), bytes of the next step, bit shift of its source words
  const uint8_t* p_raw;                  // p_dist == 0: where the next input byte is
  uint32_t pw0, pw1, pw2, pw3, pw4, pw5, pw6, pw7, pw8;

  BG_HD const uint32_t* gptr(uint32_t a /*aligned*/) const { return (const uint32_t*)(ob + a); }
  // ---------------- output ----------------
  // Slow, always-correct store of one aligned word: touches only this block's own bytes.
  BG_HD void store_word(uint32_t a /*aligned*/, uint32_t w) {
    if (a - full_lo < full_span) {
      *(uint32_t*)(ob + a) = w;
    } else {  // head / tail word shared with a neighbouring block
      for (uint32_t i = 0; i < 4; ++i)
        if (a + i >= head && a + i < oend) ob[a + i] = (uint8_t)(w >> (8 * i));
    }
  }
  // v holds n (1..4) valid low bytes, the rest zero.
  BG_HD void emit(uint32_t v, uint32_t n) {
    uint32_t k = opos & 3u;
    uint32_t sh = k * 8;
    sreg |= v << sh;
    if (k + n >= 4) {
      store_word(opos & ~3u, sreg);
      sreg = sh ? (v >> (32 - sh)) : 0u;
    }
    opos += n;
  }
  // n (1..16) bytes held in v0..v3 (bytes past n may be garbage); opos + n <= oend (the decoder checked it).
  BG_HD void emit_n16(uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3, uint32_t n) {
    uint32_t a = opos & ~3u;
    uint32_t k = opos & 3u, sh = k * 8;
    uint32_t o0 = sreg | (v0 << sh);
    uint32_t o1 = bg_funnel_l(v0, v1, sh);
    uint32_t o2 = bg_funnel_l(v1, v2, sh);
    uint32_t o3 = bg_funnel_l(v2, v3, sh);
    uint32_t o4 = bg_funnel_l(v3, 0u, sh);
    uint32_t tot = k + n;          // 1..19
    uint32_t nfull = tot >> 2;     // whole words completed: 0..4
    if (a - full_lo < fast_span) {
      // all four words lie wholly inside the block: straight predicated stores
      uint32_t* w = (uint32_t*)(ob + a);
      if (nfull > 0) w[0] = o0;
      if (nfull > 1) w[1] = o1;
      if (nfull > 2) w[2] = o2;
      if (nfull > 3) w[3] = o3;
    } else {
      if (nfull > 0) store_word(a, o0);
      if (nfull > 1) store_word(a + 4, o1);
      if (nfull > 2) store_word(a + 8, o2);
      if (nfull > 3) store_word(a + 12, o3);
    }
    uint32_t r = o0;
    if (nfull == 1) r = o1;
    if (nfull == 2) r = o2;
    if (nfull == 3) r = o3;
    if (nfull == 4) r = o4;
    uint32_t rb = tot & 3u;
    sreg = r & ((1u << (8 * rb)) - 1u);
    opos += n;
  }
  // Make every byte below opos visible in memory (bytes of the partially filled word).
  BG_HD void flush_partial() {
    uint32_t k = opos & 3u;
    uint32_t a = opos - k;
    for (uint32_t i = 0; i < k; ++i)
      if (a + i >= head) ob[a + i] = (uint8_t)(sreg >> (8 * i));
  }

  // ---------------- copies ----------------
  // Every copy is only REGISTERED (p_len, p_dist; request_copy asks for its source words) and performed by
  // complete_copy one trip later, in straight-line code, at most 16 (dist < 35) or 32 bytes per trip; what is left
  // is requested again and the lane takes no new token until the copy is done.  So no lane ever loops over a window
  // load inside a trip while the other 31 wait for its latency.
  //   dist >= 8: the first 20 (36 for dist >= 35 and more than 16 bytes) source bytes are requested as aligned
  //     words, no use; the L2/DRAM latency of the window read overlaps the rest of the trip and the ring hand-over.
  //     A step uses at most dist - 3 of them: those lie below opos - 3, i.e. are in memory (only the <= 3 bytes of
  //     sreg are not), and the following steps see what this one stored (byte-serial semantics for overlapping copies).
  //   dist < 8 (the dist=1 runs of BAM quality strings, ~2 % of matches): the two words before the partially filled
  //     one are requested; complete_copy builds 16 bytes of the periodic pattern from them and the staging register,
  //     and the rest continues as an ordinary copy at a multiple of the period (17..21, so 14..16 bytes per step),
  //     which by then lies wholly inside the periodic region.
  //   dist == 0: a stored block; the source is the compressed input at p_raw.
  // One request site, one instruction sequence for every kind of copy: a second set of loads into the same
  // registers elsewhere in the loop would have to wait for these to land (register scoreboard).
  BG_HD void request_copy() {
    const uint32_t len = p_len, dist = p_dist;
    const bool raw = dist == 0, near = dist < 8 && !raw;
    const uint32_t src = opos - dist;
    const uint32_t lim = (raw || near) ? 16u : (dist - 3u < 16u ? dist - 3u : 16u);
    p_n = len < lim ? len : lim;
    const uint32_t sa = raw ? (uint32_t)(uintptr_t)p_raw : (near ? opos : src);
    p_ssh = (sa & 3u) * 8;
    // far: the aligned words holding src..; near: the two words before the partially filled one (the buffer is
    // readable 16 bytes before out[0]); raw: the aligned words holding p_raw..
    const uint32_t* w = raw ? (const uint32_t*)((uintptr_t)p_raw & ~(uintptr_t)3) : gptr(sa & ~3u) - (near ? 2 : 0);
    pw0 = w[0]; pw1 = w[1]; pw2 = w[2]; pw3 = w[3]; pw4 = w[4];
    if ((dist >= 35 || raw) && len > 16) { pw5 = w[5]; pw6 = w[6]; pw7 = w[7]; pw8 = w[8]; }
  }
  BG_HD void complete_copy() {
    if (p_len) {
      uint32_t dist = p_dist, ssh = p_ssh;
      if (dist - 1u < 7u) {
        // last 8 bytes before opos = two stored words + the staging register -> 16 bytes of the periodic pattern
        const uint64_t X = (uint64_t)bg_funnel_r(pw0, pw1, ssh) | ((uint64_t)bg_funnel_r(pw1, sreg, ssh) << 32);
        uint64_t W = X >> (8 * (8 - dist));  // low `dist` bytes = the period
        for (uint32_t k = dist; k < 8; k <<= 1) W |= W << (8 * k);
        const uint32_t deff = dist < 3 ? 4u : (dist == 3 ? 6u : dist);   // a multiple of the period in 4..7
        const uint32_t wsh = 8 * (8 - deff);
        pw0 = (uint32_t)W; W = (W >> 32) | ((W >> wsh) << 32);
        pw1 = (uint32_t)W; W = (W >> 32) | ((W >> wsh) << 32);
        pw2 = (uint32_t)W; W = (W >> 32) | ((W >> wsh) << 32);
        pw3 = (uint32_t)W;
        pw4 = 0;
        ssh = 0;
        dist = (uint32_t)(0x5654A4A51ull >> (5 * (dist - 1))) & 31u;   // dist * floor((16 + dist) / dist): 17,18,18,20,20,18,21
        p_dist = dist;
      }
      const bool two = dist >= 35 || dist == 0;
      uint32_t rem = p_len, n = p_n;
      emit_n16(bg_funnel_r(pw0, pw1, ssh), bg_funnel_r(pw1, pw2, ssh), bg_funnel_r(pw2, pw3, ssh),
               bg_funnel_r(pw3, pw4, ssh), n);
      rem -= n;
      if (rem && two) {
        n = rem < 16 ? rem : 16;
        emit_n16(bg_funnel_r(pw4, pw5, ssh), bg_funnel_r(pw5, pw6, ssh), bg_funnel_r(pw6, pw7, ssh),
                 bg_funnel_r(pw7, pw8, ssh), n);
        rem -= n;
      }
      if (dist == 0) p_raw += p_len - rem;
      p_len = rem;   // != 0: the copy goes on next trip at the same distance
    }
  }

  BG_HD void begin_block(const InflateJob& job, uint32_t b) {
    blk = b;
    uint64_t o = job.out_off[b];
    ob = job.out + (o & ~(uint64_t)3);
    head = (uint32_t)(o & 3);
    opos = head;
    oend = head + job.isize[b];
    sreg = 0;
    full_lo = head ? 4u : 0u;
    uint32_t full_hi = oend & ~3u;
    full_span = full_hi > full_lo ? full_hi - full_lo : 0u;
    fast_span = full_span >= 16u ? full_span - 12u : 0u;
  }

  // Tokens that open / close a block or carry a stored block.
  BG_HD void cold_token(const InflateJob& job, uint32_t kind, uint32_t w1) {
    if (kind == TK_BEGIN) {
      begin_block(job, w1);
    } else if (kind == TK_END) {
      flush_partial();
      job.status[blk] = w1;
    } else if (kind == TK_RAW) {
      p_len = w1 & 0xffffu;
      p_dist = 0;
      p_raw = job.comp + job.in_off[blk] + (w1 >> 16);
    } else if (kind == TK_DONE) {
      done = 1;
    }
  }

  // One trip: finish the pending copy step; take the token (w0, w1) the caller peeked (`have`; never after TK_DONE)
  // unless the copy is still not done; request the source words of what is pending.
  BG_HD void trip(const InflateJob& job, bool have, uint32_t w0, uint32_t w1) {
#if BG_DIAG_NULL_COPIER   // diagnostic only (wrong output): take tokens, move no bytes -> the decoders' own pace
    if (have) {
      tail += 1;
      ring.set_tail(tail);
      const uint32_t kind = (w0 >> 27) & 7u;
      if (kind == TK_BEGIN) blk = w1;
      else if (kind == TK_END) job.status[blk] = w1;
      else if (kind == TK_DONE) done = 1;
    }
    return;
#endif
    complete_copy();
    if (have && p_len == 0) {
      tail += 1;
      ring.set_tail(tail);
      const uint32_t nlit = (w0 >> 24) & 7u, kind = (w0 >> 27) & 7u;
      if (nlit) emit((w0 & 0xffffffu) | (((w1 >> 23) & 0xffu) << 24), nlit);
      if (kind == TK_MATCH) {
        p_len = (w1 & 0xffu) + 3u;
        p_dist = ((w1 >> 8) & 0x7fffu) + 1u;
      } else if (kind != TK_LITS) {
        cold_token(job, kind, w1);
      }
    }
    if (p_len) request_copy();
  }
};

}  // namespace bamgpu
