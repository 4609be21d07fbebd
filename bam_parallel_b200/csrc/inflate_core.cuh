// inflate_core.cuh — per-lane raw-DEFLATE (RFC 1951) decoder used by kernel K1.
//
// Replaces the reference's `inflate_data` (bam_tools/src/util.rs:10-16, i.e. flate2's
// DeflateDecoder::read_to_end) for one BGZF block per LANE: 32 independent streams advance in
// lock-step inside a warp.  Design notes (see DESIGN.md "K1"):
//   * Huffman decode is canonical and LUT-free: the 15 left-aligned code-length limits of the
//     literal/length and distance codes live in REGISTERS, the code length is found with 14
//     independent compares (no divergence between lanes, no memory access), and only the
//     (base,offset) pair and the symbol come from shared memory.  Tables are 736 B per stream
//     so 288 streams fit in one SM's 227 KB.
//   * Shared memory is interleaved by lane ([word][lane]) so any per-lane index pattern is
//     bank-conflict free.
//   * Output goes through a 4-byte staging register and is stored as aligned 32-bit words;
//     LZ77 copies stream 4 bytes per step either from a register-resident periodic window
//     (distance < 8, covers the dist=1 runs of BAM quality strings) or from global memory
//     with one aligned load per step (distance >= 8).
//   * A lane is a small state machine (DECODE / COPY / STORED / HEADER / NEXT) so that a long
//     match or a table build in one lane does not stall the other 31.
//
// Everything here is __host__ __device__ so tests/host_emul.cpp can run the identical logic on
// the CPU build box (no GPU there).  That emulation is a TEST of this code, not a fallback:
// the product library only ever launches the kernel.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BG_HD __host__ __device__ __forceinline__
#if defined(BG_INLINE_HEADER)
#define BG_HD_COLD __host__ __device__ __forceinline__
#else
#define BG_HD_COLD __host__ __device__ __noinline__
#endif
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

// Shared-memory words per lane.
constexpr int LIT_SYM_WORDS = 144;  // 288 x u16
constexpr int DIST_SYM_WORDS = 8;   // 32 x u8
constexpr int META_WORDS = 16;      // per code length: base (15 bit, left aligned) | offset << 16
constexpr int LANE_WORDS = LIT_SYM_WORDS + DIST_SYM_WORDS + 2 * META_WORDS;  // 184 words = 736 B
constexpr int SM_LIT_SYM = 0;
constexpr int SM_DIST_SYM = LIT_SYM_WORDS;
constexpr int SM_LIT_META = SM_DIST_SYM + DIST_SYM_WORDS;
constexpr int SM_DIST_META = SM_LIT_META + META_WORDS;

enum LaneState : uint32_t { ST_NEXT = 0, ST_HEADER, ST_DECODE, ST_COPY_WIN, ST_COPY_MEM, ST_STORED, ST_DONE };

#if defined(__CUDA_ARCH__)
BG_HD uint32_t bg_brev(uint32_t x) { return __brev(x); }
BG_HD uint32_t bg_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
BG_HD uint32_t bg_ld_in(const uint32_t* p) { return __ldg(p); }
#else
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
inline uint32_t bg_ld_in(const uint32_t* p) { return *p; }
#endif

// Work description shared by all lanes.
struct InflateJob {
  const uint8_t* comp;       // compressed bytes; readable 16 bytes past the last payload
  const uint64_t* in_off;    // per block: payload offset in comp
  const uint32_t* in_len;    // per block: payload bytes
  const uint64_t* out_off;   // per block: offset in out
  const uint32_t* isize;     // per block: expected inflated bytes
  uint8_t* out;              // inflated stream; readable 16 bytes BEFORE out[0] and 4 past the end
  uint32_t* status;          // per block: IST_*
  uint32_t n_blocks;
};

template <int STRIDE>  // shared-memory word stride between consecutive entries of one lane (32 on GPU)
struct Lane {
  // ---- tables (shared memory column of this lane) ----
  uint32_t* sm;
  // ---- input bit reader ----
  const uint32_t* wp;      // next aligned word to load
  const uint32_t* w0;      // first aligned word of the payload
  const uint32_t* wend;    // first aligned word entirely past the payload (never loaded)
  uint64_t bb;
  uint32_t nbits;
  uint32_t skip_bits;      // bits of *w0 that precede the payload
  uint32_t in_bits;        // payload length in bits
  // ---- output ----
  uint8_t* out;
  uint64_t opos, ostart, oend;
  uint32_t sreg;
  // ---- decode state ----
  uint32_t state, bfinal, err, blk;
  uint32_t llim[16], dlim[16];  // [1..15] used; limits are exclusive, left-aligned to 15 bits
  // ---- pending copy ----
  uint32_t cp_rem;
  uint64_t cp_a;           // window mode: the 8-byte periodic window; memory mode: aligned source byte offset
  uint32_t cp_b, cp_c;     // window mode: cp_b = 8*(8-d_eff); memory mode: cp_b = previous word, cp_c = shift

  // ---------------- shared-memory accessors ----------------
  BG_HD uint32_t& smw(int w) { return sm[w * STRIDE]; }
  BG_HD uint32_t lit_sym(uint32_t i) { return (smw(SM_LIT_SYM + (i >> 1)) >> ((i & 1) * 16)) & 0xffffu; }
  BG_HD void set_lit_sym(uint32_t i, uint32_t v) {
    uint32_t& w = smw(SM_LIT_SYM + (i >> 1));
    uint32_t sh = (i & 1) * 16;
    w = (w & ~(0xffffu << sh)) | (v << sh);
  }
  BG_HD uint32_t dist_sym(uint32_t i) { return (smw(SM_DIST_SYM + (i >> 2)) >> ((i & 3) * 8)) & 0xffu; }
  BG_HD void set_dist_sym(uint32_t i, uint32_t v) {
    uint32_t& w = smw(SM_DIST_SYM + (i >> 2));
    uint32_t sh = (i & 3) * 8;
    w = (w & ~(0xffu << sh)) | (v << sh);
  }

  // ---------------- bit reader ----------------
  BG_HD void refill() {
    if (nbits <= 32) {
      uint32_t w = wp < wend ? bg_ld_in(wp) : 0u;  // past the payload: zeros; the overrun is reported later
      bb |= (uint64_t)w << nbits;
      ++wp;
      nbits += 32;
    }
  }
  BG_HD void consume(uint32_t n) { bb >>= n; nbits -= n; }
  BG_HD uint32_t bits_consumed() const { return (uint32_t)(wp - w0) * 32u - skip_bits - nbits; }

  BG_HD void begin_input(const uint8_t* p, uint32_t len) {
    uintptr_t a = (uintptr_t)p;
    w0 = (const uint32_t*)(a & ~(uintptr_t)3);
    wend = (const uint32_t*)((a + len + 3) & ~(uintptr_t)3);
    skip_bits = (uint32_t)(a & 3) * 8;
    wp = w0 + 1;
    bb = (uint64_t)(bg_ld_in(w0) >> skip_bits);
    nbits = 32 - skip_bits;
    in_bits = len * 8;
  }

  // ---------------- output ----------------
  BG_HD void store_word(uint64_t a /*aligned*/, uint32_t w) {
    if (a >= ostart && a + 4 <= oend) {
      *(uint32_t*)(out + a) = w;
    } else {  // head / tail word shared with a neighbouring block: touch only our own bytes
      for (int i = 0; i < 4; ++i)
        if (a + i >= ostart && a + i < oend) out[a + i] = (uint8_t)(w >> (8 * i));
    }
  }
  // v holds n (1..4) valid low bytes, the rest zero.
  BG_HD void emit(uint32_t v, uint32_t n) {
    uint32_t k = (uint32_t)opos & 3u;
    uint32_t sh = k * 8;
    sreg |= v << sh;
    if (k + n >= 4) {
      store_word(opos & ~(uint64_t)3, sreg);
      sreg = sh ? (v >> (32 - sh)) : 0u;
    }
    opos += n;
  }
  // Make every byte below opos visible in memory (bytes of the partially filled word).
  BG_HD void flush_partial() {
    uint32_t k = (uint32_t)opos & 3u;
    uint64_t a = opos - k;
    for (uint32_t i = 0; i < k; ++i)
      if (a + i >= ostart) out[a + i] = (uint8_t)(sreg >> (8 * i));
  }

  // ---------------- canonical Huffman ----------------
  // Number of k in 1..14 with v >= lim[k], plus 1.
  BG_HD static uint32_t code_len(const uint32_t* lim, uint32_t v) {
    uint32_t L = 1;
#pragma unroll
    for (int k = 1; k <= 14; ++k) L += (v >= lim[k]) ? 1u : 0u;
    return L;
  }

  // Builds sorted-symbol table + per-length meta + register limits for `n` symbols whose code
  // lengths are lens[0..n).  is_lit selects the table and the symbol encoding.  Returns 0 or IST_*.
  BG_HD uint32_t build_table(const uint8_t* lens, uint32_t n, bool is_lit) {
    const int meta = is_lit ? SM_LIT_META : SM_DIST_META;
    uint32_t* lim = is_lit ? llim : dlim;
    for (int L = 0; L < 16; ++L) smw(meta + L) = 0;
    for (uint32_t s = 0; s < n; ++s) smw(meta + lens[s]) += 1;
    uint32_t cnt[16];
#pragma unroll
    for (int L = 1; L <= 15; ++L) cnt[L] = smw(meta + L);
    uint32_t code = 0, offs = 0, bad = 0;
#pragma unroll
    for (int L = 1; L <= 15; ++L) {
      code <<= 1;
      if (code + cnt[L] > (1u << L)) bad = 1;
      lim[L] = (code + cnt[L]) << (15 - L);
      smw(meta + L) = offs;  // scatter cursor
      code += cnt[L];
      offs += cnt[L];
    }
    if (bad) return IST_OVERSUBSCRIBED;
    for (uint32_t s = 0; s < n; ++s) {
      uint32_t l = lens[s];
      if (l) {
        uint32_t pos = smw(meta + l);
        smw(meta + l) = pos + 1;
        if (is_lit) {
          uint32_t e;
          if (s < 256) e = s;
          else if (s == 256) e = 0x8000u;
          else if (s > 285) e = 0x8000u | 0x1ffu;  // 286/287: not valid in a stream
          else {
            uint32_t c = s - 257, eb, base;
            if (c < 8) { eb = 0; base = 3 + c; }
            else if (c == 28) { eb = 0; base = 258; }
            else { eb = (c - 4) >> 2; base = 3 + ((4 + (c & 3)) << eb); }
            e = 0x8000u | (eb << 12) | base;
          }
          set_lit_sym(pos, e);
        } else {
          set_dist_sym(pos, s);
        }
      }
    }
    code = 0; offs = 0;
#pragma unroll
    for (int L = 1; L <= 15; ++L) {
      code <<= 1;
      smw(meta + L) = (code << (15 - L)) | (offs << 16);
      code += cnt[L];
      offs += cnt[L];
    }
    return IST_OK;
  }

  // Parses one DEFLATE block header (and its code tables).  Sets state.
  BG_HD_COLD void do_header(uint8_t* lens /* >= 320 bytes of private scratch */) {
    refill();
    if (bits_consumed() + 3 > in_bits) { err = IST_INPUT_OVERRUN; state = ST_NEXT; return; }
    bfinal = (uint32_t)bb & 1u;
    uint32_t btype = ((uint32_t)bb >> 1) & 3u;
    consume(3);
    if (btype == 0) {
      consume(nbits & 7);  // to the next byte boundary
      refill();
      uint32_t len = (uint32_t)bb & 0xffffu, nlen = ((uint32_t)bb >> 16) & 0xffffu;
      consume(32);
      if (len != (nlen ^ 0xffffu)) { err = IST_BAD_STORED; state = ST_NEXT; return; }
      cp_rem = len;
      state = len ? ST_STORED : (bfinal ? ST_NEXT : ST_HEADER);
      return;
    }
    if (btype == 3) { err = IST_BAD_BTYPE; state = ST_NEXT; return; }
    uint32_t hlit, hdist;
    if (btype == 1) {
      hlit = 288; hdist = 30;
      for (int i = 0; i < 144; ++i) lens[i] = 8;
      for (int i = 144; i < 256; ++i) lens[i] = 9;
      for (int i = 256; i < 280; ++i) lens[i] = 7;
      for (int i = 280; i < 288; ++i) lens[i] = 8;
      for (int i = 288; i < 318; ++i) lens[i] = 5;
    } else {
      refill();
      hlit = ((uint32_t)bb & 31u) + 257;
      hdist = (((uint32_t)bb >> 5) & 31u) + 1;
      uint32_t hclen = (((uint32_t)bb >> 10) & 15u) + 4;
      consume(14);
      if (hlit > 286 || hdist > 30) { err = IST_BAD_COUNTS; state = ST_NEXT; return; }
      // code-length code lengths, 3 bits each, sent in the RFC's permuted order
      // 16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15; packed here 3 bits per symbol.
      uint64_t cl = 0;
      for (uint32_t i = 0; i < hclen; ++i) {
        refill();
        uint32_t l = (uint32_t)bb & 7u;
        consume(3);
        uint32_t sym;
        if (i < 3) sym = 16 + i;
        else {
          uint32_t j = (i - 3) >> 1;                       // odd i: 0,7,6,..,1   even i: 8,9,..,15
          sym = (i & 1) ? (j == 0 ? 0u : 8u - j) : 8u + j;
        }
        cl |= (uint64_t)l << (3 * sym);
      }
      // canonical code for the 19-symbol alphabet (max length 7), kept in registers + lit_meta/dist_meta scratch
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
        if (bad) { err = IST_OVERSUBSCRIBED; state = ST_NEXT; return; }
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
        refill();
        uint32_t v7 = bg_brev((uint32_t)bb) >> 25;
        uint32_t L = 1;
#pragma unroll
        for (int k = 1; k <= 6; ++k) L += (v7 >= clim[k]) ? 1u : 0u;
        if (v7 >= clim[7]) { err = IST_BAD_CODELEN; state = ST_NEXT; return; }
        uint32_t base = 0, offs = 0;
#pragma unroll
        for (int k = 1; k <= 7; ++k)
          if (L == (uint32_t)k) { base = cbase[k]; offs = coffs[k]; }
        uint32_t idx = offs + ((v7 - base) >> (7 - L));
        uint32_t sym = idx < 12 ? (uint32_t)(cs_lo >> (5 * idx)) & 31u : (uint32_t)(cs_hi >> (5 * (idx - 12))) & 31u;
        consume(L);
        if (sym < 16) {
          lens[i++] = (uint8_t)sym;
          prev = sym;
        } else {
          uint32_t rep, val;
          if (sym == 16) {
            if (i == 0) { err = IST_BAD_CODELEN; state = ST_NEXT; return; }
            rep = 3 + ((uint32_t)bb & 3u); consume(2); val = prev;
          } else if (sym == 17) {
            rep = 3 + ((uint32_t)bb & 7u); consume(3); val = 0; prev = 0;
          } else {
            rep = 11 + ((uint32_t)bb & 127u); consume(7); val = 0; prev = 0;
          }
          if (i + rep > total) { err = IST_BAD_CODELEN; state = ST_NEXT; return; }
          for (uint32_t r = 0; r < rep; ++r) lens[i++] = (uint8_t)val;
        }
      }
      if (lens[256] == 0) { err = IST_NO_EOB; state = ST_NEXT; return; }
      if (bits_consumed() > in_bits) { err = IST_INPUT_OVERRUN; state = ST_NEXT; return; }
    }
    uint32_t e = build_table(lens, hlit, true);
    if (!e) e = build_table(lens + hlit, hdist, false);
    if (e) { err = e; state = ST_NEXT; return; }
    state = ST_DECODE;
  }

  // One literal/length (+ distance) token.  Returns emitted literal in v/n (n = 0 or 1).
  BG_HD void do_decode(uint32_t& v, uint32_t& n) {
    refill();
    uint32_t c = bg_brev((uint32_t)bb) >> 17;
    uint32_t L = code_len(llim, c);
    if (c >= llim[15]) { err = IST_BAD_SYMBOL; state = ST_NEXT; return; }
    uint32_t m = smw(SM_LIT_META + L);
    uint32_t e = lit_sym((m >> 16) + ((c - (m & 0xffffu)) >> (15 - L)));
    consume(L);
    if (!(e & 0x8000u)) {
      if (opos >= oend) { err = IST_OUTPUT_OVERRUN; state = ST_NEXT; return; }
      v = e; n = 1;
      return;
    }
    uint32_t blen = e & 0x1ffu;
    if (blen == 0) {  // end of block
      state = bfinal ? ST_NEXT : ST_HEADER;
      return;
    }
    if (blen == 0x1ffu) { err = IST_BAD_SYMBOL; state = ST_NEXT; return; }
    uint32_t eb = (e >> 12) & 7u;
    uint32_t len = blen + ((uint32_t)bb & ((1u << eb) - 1u));
    consume(eb);
    refill();
    c = bg_brev((uint32_t)bb) >> 17;
    L = code_len(dlim, c);
    if (c >= dlim[15]) { err = IST_BAD_SYMBOL; state = ST_NEXT; return; }
    m = smw(SM_DIST_META + L);
    uint32_t ds = dist_sym((m >> 16) + ((c - (m & 0xffffu)) >> (15 - L)));
    consume(L);
    if (ds >= 30) { err = IST_BAD_SYMBOL; state = ST_NEXT; return; }
    uint32_t deb = ds < 2 ? 0u : (ds >> 1) - 1u;
    uint32_t dist = (ds < 2 ? ds + 1u : 1u + ((2u | (ds & 1u)) << deb)) + ((uint32_t)bb & ((1u << deb) - 1u));
    consume(deb);
    if ((uint64_t)dist > opos - ostart) { err = IST_BAD_DISTANCE; state = ST_NEXT; return; }
    if (opos + len > oend) { err = IST_OUTPUT_OVERRUN; state = ST_NEXT; return; }
    flush_partial();
    cp_rem = len;
    if (dist < 8) {
      // last 8 bytes before opos (the buffer is readable 16 bytes before out[0])
      uint64_t q = opos - 8;
      const uint32_t* a = (const uint32_t*)(out + (q & ~(uint64_t)3));
      uint32_t sh = (uint32_t)(q & 3) * 8;
      uint32_t x0 = a[0], x1 = a[1], x2 = a[2];
      uint64_t X = (uint64_t)bg_funnel_r(x0, x1, sh) | ((uint64_t)bg_funnel_r(x1, x2, sh) << 32);
      uint64_t W = X >> (8 * (8 - dist));  // low `dist` bytes = the period
      for (uint32_t k = dist; k < 8; k <<= 1) W |= W << (8 * k);
      uint32_t deff = dist < 3 ? 4u : (dist == 3 ? 6u : dist);
      cp_a = W;
      cp_b = 8 * (8 - deff);
      state = ST_COPY_WIN;
    } else {
      uint64_t src = opos - dist;
      uint64_t sa = src & ~(uint64_t)3;
      cp_a = sa + 4;
      cp_b = *(const uint32_t*)(out + sa);
      cp_c = (uint32_t)(src & 3) * 8;
      state = ST_COPY_MEM;
    }
  }

  BG_HD void do_copy_win(uint32_t& v, uint32_t& n) {
    n = cp_rem < 4 ? cp_rem : 4;
    v = (uint32_t)cp_a;
    if (n < 4) v &= (1u << (8 * n)) - 1u;
    cp_a = (cp_a >> 32) | ((cp_a >> cp_b) << 32);
    cp_rem -= n;
    if (cp_rem == 0) state = ST_DECODE;
  }

  BG_HD void do_copy_mem(uint32_t& v, uint32_t& n) {
    n = cp_rem < 4 ? cp_rem : 4;
    uint32_t nx = *(const uint32_t*)(out + cp_a);
    v = bg_funnel_r(cp_b, nx, cp_c);
    if (n < 4) v &= (1u << (8 * n)) - 1u;
    cp_b = nx;
    cp_a += 4;
    cp_rem -= n;
    if (cp_rem == 0) state = ST_DECODE;
  }

  BG_HD void do_stored(uint32_t& v, uint32_t& n) {
    refill();
    n = cp_rem < 4 ? cp_rem : 4;
    if (opos + n > oend) { err = IST_OUTPUT_OVERRUN; state = ST_NEXT; n = 0; return; }
    v = (uint32_t)bb;
    if (n < 4) v &= (1u << (8 * n)) - 1u;
    consume(8 * n);
    cp_rem -= n;
    if (cp_rem == 0) state = bfinal ? ST_NEXT : ST_HEADER;
  }

  // Finishes the current block (if any): flush, verify sizes, write status.
  BG_HD void finish_block(const InflateJob& job) {
    if (blk != 0xffffffffu) {
      flush_partial();
      if (!err) {
        if (opos != oend) err = IST_SIZE_MISMATCH;
        else if (bits_consumed() > in_bits) err = IST_INPUT_OVERRUN;
      }
      job.status[blk] = err;
    }
  }

  BG_HD void begin_block(const InflateJob& job, uint32_t b) {
    blk = b;
    err = 0;
    bfinal = 0;
    ostart = job.out_off[b];
    oend = ostart + job.isize[b];
    opos = ostart;
    sreg = 0;
    out = job.out;
    begin_input(job.comp + job.in_off[b], job.in_len[b]);
    state = ST_HEADER;
  }

  // One step of the lane state machine (everything except ST_NEXT, which needs the work counter).
  BG_HD void step(uint8_t* lens) {
    uint32_t v = 0, n = 0;
    if (state == ST_DECODE) do_decode(v, n);
    else if (state == ST_COPY_MEM) do_copy_mem(v, n);
    else if (state == ST_COPY_WIN) do_copy_win(v, n);
    else if (state == ST_STORED) do_stored(v, n);
    else if (state == ST_HEADER) do_header(lens);
    if (n) emit(v, n);
  }
};

}  // namespace bamgpu
