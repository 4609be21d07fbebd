// inflate_warp.cu — K1 variant 1: ONE WARP PER BGZF BLOCK, the mapping BASELINE.json's north-star prescribes
// ("one warp per independent <=64 KB BGZF block: dynamic Huffman tables in shared memory, decode with warp-level
// primitives, LZ77 back-references copied cooperatively across the warp, coalesced stores").
//
// Same job as the default kernel (inflate.cu: one block per lane PAIR), same inputs, outputs and status words; selected
// with bamgpu_opts.inflate_impl = 1.  It exists so that the choice between the two mappings is a measurement
// (profiles/r2_k1_variants.txt, DESIGN.md "K1 variants"), not an assertion.
//
// Mapping: the 32 lanes of a warp hold ONE bit reader in lockstep (every lane has the same bit buffer, loads the same
// input word - a broadcast, one sector - and looks up the same table entry), so the Huffman chain costs each symbol's
// instructions once per warp and nothing has to be shuffled around; the lanes only part ways to move bytes:
//   * tables: a 10-bit first-level look-up table for the literal/length code and a 9-bit one for the distance code
//     (entry = symbol | code length), canonical limits + sorted symbols for the few longer codes; 3.9 KB per warp, built
//     cooperatively (code lengths are parsed by the bit reader; ranks inside one code length come from __match_any_sync
//     over 32 symbols at a time; every lane scatters its symbol's 2^(10-L) entries);
//   * LZ77 copy: lane l copies bytes l, l+32, ... of the match; a source byte is (i mod dist) bytes behind the match
//     start, i.e. ALWAYS in front of the match, so overlapping copies (dist < len, the dist=1 runs of quality strings)
//     need no ordering inside the copy - one __syncwarp() in front of it makes the earlier stores visible;
//   * literals: lane 0 stores the byte.
// Loads and stores of a copy are 32 consecutive bytes per iteration (one sector).
#include <cstdio>
#include <cstdlib>

#include "inflate_core.cuh"
#include "kernels.h"

namespace bamgpu {
namespace {

constexpr int WB_WARPS = 8;                 // warps (= blocks in flight) per CTA
constexpr int WB_THREADS = WB_WARPS * 32;
constexpr int LIT_BITS = 10, DIST_BITS = 9;

__constant__ uint8_t CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};   // RFC 1951 3.2.7

struct __align__(16) WarpTables {
  uint16_t lit_lut[1 << LIT_BITS];     // bits 0-8 symbol, bits 9-12 code length (0: code longer than LIT_BITS)
  uint16_t dist_lut[1 << DIST_BITS];   // bits 0-4 symbol, bits 5-8 code length
  uint16_t lit_sorted[288];            // symbols in canonical order
  uint8_t dist_sorted[32];
  uint32_t lit_lim[16], dist_lim[16];  // [L]: exclusive limit of the codes of length <= L, left-aligned to 15 bits
  uint16_t lit_off[16], dist_off[16];  // [L]: index of the first symbol of length L minus its first code
  uint16_t cursor[16];                 // scatter cursors while building
  uint8_t lens[320];
};

struct Bits {   // identical in every lane
  const uint8_t* p;   // next aligned word to load
  uint64_t bb;
  uint32_t nbits, fed, in_bits;
  __device__ __forceinline__ void begin(const uint8_t* src, uint32_t len) {
    uintptr_t a = (uintptr_t)src;
    p = (const uint8_t*)(a & ~(uintptr_t)3);
    uint32_t sb = (uint32_t)(a & 3) * 8;
    bb = (uint64_t)(__ldg((const uint32_t*)p) >> sb);
    p += 4;
    nbits = 32 - sb;
    fed = nbits;
    in_bits = len * 8;
  }
  // >= 33 valid bits afterwards.  Never loads beyond the word that holds the payload's last bit (+ one more: the buffer is
  // padded): a corrupt stream runs on zeros until one of the checks stops it.
  __device__ __forceinline__ void refill() {
    if (nbits <= 32) {
      const uint32_t w = fed <= in_bits + 32 ? __ldg((const uint32_t*)p) : 0u;
      bb |= (uint64_t)w << nbits;
      p += 4;
      nbits += 32;
      fed += 32;
    }
  }
  __device__ __forceinline__ uint32_t peek(uint32_t n) const { return (uint32_t)bb & ((1u << n) - 1u); }
  __device__ __forceinline__ void consume(uint32_t n) { bb >>= n; nbits -= n; }
  __device__ __forceinline__ uint32_t take(uint32_t n) { uint32_t v = peek(n); consume(n); return v; }
  __device__ __forceinline__ uint32_t consumed() const { return fed - nbits; }
};

// Canonical tables for `n` symbols with code lengths lens[0..n): sorted symbols, limits, offsets and the first-level
// look-up table.  Warp-cooperative; returns 0 or IST_OVERSUBSCRIBED (uniform).
template <int LUT_BITS, bool LIT>
__device__ uint32_t build_code(WarpTables& T, const uint8_t* lens, uint32_t n, int lane) {
  uint16_t* lut = LIT ? T.lit_lut : T.dist_lut;
  uint32_t* lim = LIT ? T.lit_lim : T.dist_lim;
  uint16_t* off = LIT ? T.lit_off : T.dist_off;
  // counts per length: every lane counts its symbols, then a warp reduction per length
  uint32_t cnt_lo = 0, cnt_hi = 0;   // 16 x 4-bit... symbols per lane <= 10, so nibbles suffice
  for (uint32_t s = lane; s < n; s += 32) {
    uint32_t l = lens[s];
    if (l < 8) cnt_lo += 1u << (4 * l); else cnt_hi += 1u << (4 * (l - 8));
  }
  uint32_t code = 0, offs = 0, bad = 0;
#pragma unroll
  for (int L = 1; L <= 15; ++L) {
    uint32_t c = (L < 8 ? cnt_lo >> (4 * L) : cnt_hi >> (4 * (L - 8))) & 15u;
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    code <<= 1;
    if (code + c > (1u << L)) bad = 1;
    if (lane == 0) {
      lim[L] = (code + c) << (15 - L);
      off[L] = (uint16_t)((offs - code) & 0xffffu);
      T.cursor[L] = (uint16_t)offs;
    }
    code += c;
    offs += c;
  }
  if (bad) return IST_OVERSUBSCRIBED;
  for (uint32_t i = lane; i < (1u << LUT_BITS); i += 32) lut[i] = 0;
  __syncwarp();
  // canonical rank: symbols in increasing order, 32 at a time; lanes with the same length form a group whose members
  // take consecutive slots behind the group's cursor
  for (uint32_t base = 0; base < n; base += 32) {
    const uint32_t s = base + lane;
    const uint32_t l = s < n ? lens[s] : 0;
    const uint32_t grp = __match_any_sync(0xffffffffu, l);
    const uint32_t rank = __popc(grp & ((1u << lane) - 1u));
    uint32_t pos = 0;
    if (l) pos = T.cursor[l] + rank;
    __syncwarp();
    if (l && rank == 0) T.cursor[l] = (uint16_t)(T.cursor[l] + __popc(grp));
    __syncwarp();
    if (l) {
      if (LIT) T.lit_sorted[pos] = (uint16_t)s; else T.dist_sorted[pos] = (uint8_t)s;
      if (l <= (uint32_t)LUT_BITS) {
        const uint32_t c = (pos - off[l]) & ((1u << l) - 1u);          // the code, l bits, MSB first
        const uint32_t r = __brev(c) >> (32 - l);                        // as it appears in the LSB-first bit buffer
        const uint16_t e = (uint16_t)(s | (l << (LIT ? 9 : 5)));
        for (uint32_t k = r; k < (1u << LUT_BITS); k += 1u << l) lut[k] = e;
      }
    }
  }
  __syncwarp();
  return IST_OK;
}

// One symbol of a code: first-level table, canonical search for the rare longer codes.  Needs 15 valid bits.
template <int LUT_BITS, bool LIT>
__device__ __forceinline__ uint32_t decode_sym(const WarpTables& T, Bits& br, bool& bad) {
  const uint16_t* lut = LIT ? T.lit_lut : T.dist_lut;
  const uint32_t e = lut[br.peek(LUT_BITS)];
  uint32_t L = e >> (LIT ? 9 : 5);
  if (L) { br.consume(L); return e & (LIT ? 511u : 31u); }
  const uint32_t* lim = LIT ? T.lit_lim : T.dist_lim;
  const uint16_t* off = LIT ? T.lit_off : T.dist_off;
  const uint32_t c = __brev((uint32_t)br.bb) >> 17;
  L = LUT_BITS + 1;
  while (L <= 15 && c >= lim[L]) ++L;
  if (L > 15) { bad = true; return 0; }
  const uint32_t idx = (off[L] + (c >> (15 - L))) & 0xffffu;
  br.consume(L);
  if (LIT) { if (idx >= 288) { bad = true; return 0; } return T.lit_sorted[idx]; }
  if (idx >= 32) { bad = true; return 0; }
  return T.dist_sorted[idx];
}

// DEFLATE block header + tables (RFC 1951 3.2.3-3.2.7).  Uniform across the warp.  Returns IST_*; *btype_out 0 = stored.
__device__ uint32_t read_block_header(WarpTables& T, Bits& br, int lane, uint32_t& bfinal, uint32_t& btype_out, uint32_t& stored_len) {
  br.refill();
  if (br.consumed() + 3 > br.in_bits) return IST_INPUT_OVERRUN;
  bfinal = br.take(1);
  const uint32_t btype = br.take(2);
  btype_out = btype;
  if (btype == 3) return IST_BAD_BTYPE;
  if (btype == 0) {
    br.consume(br.nbits & 7);
    br.refill();
    const uint32_t len = br.take(16), nlen = br.take(16);
    if (len != (nlen ^ 0xffffu)) return IST_BAD_STORED;
    if (br.consumed() > br.in_bits) return IST_INPUT_OVERRUN;
    stored_len = len;
    return IST_OK;
  }
  uint32_t hlit, hdist;
  if (btype == 1) {
    hlit = 288; hdist = 30;
    for (int i = lane; i < 318; i += 32) T.lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : i < 288 ? 8 : 5;
  } else {
    br.refill();
    hlit = br.take(5) + 257;
    hdist = br.take(5) + 1;
    const uint32_t hclen = br.take(4) + 4;
    if (hlit > 286 || hdist > 30) return IST_BAD_COUNTS;
    // the code-length code: 19 symbols, 3-bit lengths in the RFC's permuted order, max length 7 -> a 128-entry table
    // kept in the (not yet needed) distance look-up table
    uint64_t cl = 0;
    for (uint32_t i = 0; i < hclen; ++i) {
      br.refill();
      const uint32_t l = br.take(3);
      cl |= (uint64_t)l << (3 * CL_ORDER[i]);
    }
    uint32_t ccnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t s = 0; s < 19; ++s) ccnt[(cl >> (3 * s)) & 7u] += 1;
    uint32_t next_code[8], code = 0;
    ccnt[0] = 0;
    for (int L = 1; L <= 7; ++L) { code = (code + ccnt[L - 1]) << 1; next_code[L] = code; if (code + ccnt[L] > (1u << L)) return IST_OVERSUBSCRIBED; }
    uint16_t* cl_lut = T.dist_lut;   // 128 entries: symbol | length << 5
    for (int i = lane; i < 128; i += 32) cl_lut[i] = 0;
    __syncwarp();
    if (lane == 0) {
      for (uint32_t s = 0; s < 19; ++s) {
        const uint32_t l = (uint32_t)(cl >> (3 * s)) & 7u;
        if (!l) continue;
        const uint32_t c = next_code[l]++;
        const uint32_t r = __brev(c) >> (32 - l);
        for (uint32_t k = r; k < 128; k += 1u << l) cl_lut[k] = (uint16_t)(s | (l << 5));
      }
    }
    __syncwarp();
    const uint32_t total = hlit + hdist;
    uint32_t i = 0, prev = 0;
    while (i < total) {
      br.refill();
      const uint32_t e = cl_lut[br.peek(7)];
      const uint32_t L = e >> 5, sym = e & 31u;
      if (!L) return IST_BAD_CODELEN;
      br.consume(L);
      if (sym < 16) {
        if (lane == 0) T.lens[i] = (uint8_t)sym;
        ++i;
        prev = sym;
      } else {
        uint32_t rep, val;
        if (sym == 16) { if (i == 0) return IST_BAD_CODELEN; rep = 3 + br.take(2); val = prev; }
        else if (sym == 17) { rep = 3 + br.take(3); val = 0; prev = 0; }
        else { rep = 11 + br.take(7); val = 0; prev = 0; }
        if (i + rep > total) return IST_BAD_CODELEN;
        for (uint32_t r = lane; r < rep; r += 32) T.lens[i + r] = (uint8_t)val;
        i += rep;
      }
    }
    __syncwarp();
    if (T.lens[256] == 0) return IST_NO_EOB;
    if (br.consumed() > br.in_bits) return IST_INPUT_OVERRUN;
  }
  __syncwarp();
  uint32_t e = build_code<DIST_BITS, false>(T, T.lens + hlit, hdist, lane);
  if (!e) e = build_code<LIT_BITS, true>(T, T.lens, hlit, lane);
  return e;
}

__global__ void __launch_bounds__(WB_THREADS) k1_inflate_warp_per_block(InflateJob job, uint32_t* __restrict__ work_counter) {
  __shared__ WarpTables tables[WB_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpTables& T = tables[warp];
  for (;;) {
    uint32_t b = 0;
    if (lane == 0) b = atomicAdd(work_counter, 1u);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= job.n_blocks) break;
    const uint32_t isize = job.isize[b], in_len = job.in_len[b];
    const uint8_t* in = job.comp + job.in_off[b];   // where the bit reader was last (re)started, and what is left from there
    uint32_t in_left = in_len;
    uint8_t* out = job.out + job.out_off[b];
    Bits br;
    br.begin(in, in_left);
    uint32_t opos = 0, err = IST_OK, bfinal = 0;
    while (!err && !bfinal) {
      uint32_t btype = 0, stored = 0;
      err = read_block_header(T, br, lane, bfinal, btype, stored);
      if (err) break;
      if (btype == 0) {   // stored: bytes straight from the input
        const uint32_t at = br.consumed() >> 3;
        if (at + stored > in_left) { err = IST_INPUT_OVERRUN; break; }
        if (stored > isize - opos) { err = IST_OUTPUT_OVERRUN; break; }
        for (uint32_t i = lane; i < stored; i += 32) out[opos + i] = in[at + i];
        opos += stored;
        in += at + stored;
        in_left -= at + stored;
        br.begin(in, in_left);
        continue;
      }
      for (;;) {
        br.refill();
        bool bad = false;
        const uint32_t sym = decode_sym<LIT_BITS, true>(T, br, bad);
        if (bad) { err = IST_BAD_SYMBOL; break; }
        if (sym < 256) {
          if (opos >= isize) { err = IST_OUTPUT_OVERRUN; break; }
          if (lane == 0) out[opos] = (uint8_t)sym;
          ++opos;
          continue;
        }
        if (sym == 256) break;
        const uint32_t lc = sym - 257;
        if (lc > 28) { err = IST_BAD_SYMBOL; break; }
        const uint32_t eb = lc < 8 ? 0u : (lc == 28 ? 0u : (lc - 4) >> 2);
        const uint32_t lbase = lc < 8 ? 3u + lc : (lc == 28 ? 258u : 3u + ((4u + (lc & 3u)) << eb));
        br.refill();
        const uint32_t len = lbase + br.take(eb);
        const uint32_t ds = decode_sym<DIST_BITS, false>(T, br, bad);
        if (bad || ds >= 30) { err = IST_BAD_SYMBOL; break; }
        const uint32_t deb = ds < 2 ? 0u : (ds >> 1) - 1u;
        br.refill();
        const uint32_t dist = (ds < 2 ? ds + 1u : 1u + ((2u | (ds & 1u)) << deb)) + br.take(deb);
        if (dist > opos) { err = IST_BAD_DISTANCE; break; }
        if (len > isize - opos) { err = IST_OUTPUT_OVERRUN; break; }
        __syncwarp();   // the bytes this match reads were stored by other lanes
        const uint8_t* src = out + opos - dist;
        if (dist >= len) {
          for (uint32_t i = lane; i < len; i += 32) out[opos + i] = src[i];
        } else {
          for (uint32_t i = lane; i < len; i += 32) out[opos + i] = src[i % dist];
        }
        opos += len;
      }
    }
    if (!err) {
      if (opos != isize) err = IST_SIZE_MISMATCH;
      else if (br.consumed() > br.in_bits) err = IST_INPUT_OVERRUN;
    }
    if (lane == 0) job.status[b] = err;
    __syncwarp();
  }
}

}  // namespace

void launch_inflate_warp_per_block(const uint8_t* comp, const BlockTableDev& t, uint8_t* out, uint32_t* status,
                                   uint32_t* work_counter, int num_sms, cudaStream_t s) {
  if (t.n_blocks == 0) return;
  InflateJob job{comp, t.in_off, t.in_len, t.out_off, t.isize, out, status, t.n_blocks};
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1_inflate_warp_per_block, WB_THREADS, 0);
  if (per_sm < 1) per_sm = 1;
  uint32_t max_ctas = (uint32_t)num_sms * (uint32_t)per_sm;
  uint32_t need = (t.n_blocks + WB_WARPS - 1) / WB_WARPS;
  uint32_t grid = need < max_ctas ? need : max_ctas;
  static bool once = false;
  if (!once && getenv("BAMGPU_DEBUG")) {
    once = true;
    fprintf(stderr, "[bamgpu] K1 variant 1 (warp per block): %d threads/CTA, %zu B static smem/CTA, %d CTAs/SM\n", WB_THREADS,
            sizeof(WarpTables) * WB_WARPS, per_sm);
  }
  k1_inflate_warp_per_block<<<grid, WB_THREADS, 0, s>>>(job, work_counter);
}

}  // namespace bamgpu
