// fields.cu — K5: the variable-length fields of every record of a decoded batch, decoded the way the reference's
// accessors do it:
//   CIGAR   decode_cigar(get_bytes(RawCigar))      record/bamrawrecord.rs:223-236 over get_cigar :126-157 — "<len><op>"
//           per operation; an unmapped record (refID < 0 or pos < 0) or n_cigar_op == 0 gives ""; the long-CIGAR
//           convention (two ops, the first `<x>S`) is resolved through the CG tag, with the reference's quirk that x
//           is compared with (l_seq + 1) / 2 — the BYTE length of the packed sequence — not with l_seq;
//   SEQ     decode_seq(get_bytes(RawSequence))     :259-270 — "=ACMGRSVTWYHKDBN"[nibble], and the reference drops EVERY
//           low nibble that is 0 (not only the padding nibble of an odd-length read);
//   QUAL    get_bytes(RawQual)                     :99 — the l_seq raw phred bytes.
// All three go through the get_bytes closures (:80-84), whose `32 + l_read_name` is u8 arithmetic: for l_read_name >= 224 the
// slices start 256 bytes early in a release build.  Mirrored here.  Where the reference panics (a slice beyond the record,
// a CIGAR byte count that is not a multiple of 4, an operation code > 8, a malformed tag in front of CG) the record's field
// is empty and counted in bad[] — an error the caller can see instead of an abort.
//
// Two passes: k5_sizes (a thread per record) measures every field, a CUB exclusive scan turns the lengths into offsets, then
// k5_decode_cigar (a thread per record: a few characters each) and k5_decode_flat (a CTA per 256 records, threads over the
// aligned words of the CTA's contiguous output range) write the strings.  Byte work bound by HBM: per record it reads the
// packed fields once per pass and writes 1 + 2 + 1 bytes per base.
#include <cub/device/device_scan.cuh>

#include "kernels.h"
#include "tags.cuh"

namespace bamgpu {
namespace {

__device__ __forceinline__ uint32_t ldu32(const uint8_t* p) {   // any alignment; reads only the four bytes
  const uintptr_t a = (uintptr_t)p;
  const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(a & 3) * 8;
  return sh ? __funnelshift_r(w[0], w[1], sh) : w[0];
}
__device__ __forceinline__ uint32_t dec_digits(uint32_t x) {    // op_len < 2^28: at most 9 digits
  return 1u + (x >= 10u) + (x >= 100u) + (x >= 1000u) + (x >= 10000u) + (x >= 100000u) + (x >= 1000000u) + (x >= 10000000u) +
         (x >= 100000000u);
}
__device__ __forceinline__ uint64_t warp_sum64(uint64_t v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Offsets of the variable-length fields as get_bytes computes them (bamrawrecord.rs:80-84).
struct Geom {
  int32_t ref, pos;
  uint32_t n_cig, l_seq, nseq;
  uint64_t cig, seq, qual, tags;
};
__device__ __forceinline__ Geom geometry(const uint8_t* body) {
  Geom g;
  g.ref = (int32_t)ldu32(body);
  g.pos = (int32_t)ldu32(body + 4);
  const uint32_t w2 = ldu32(body + 8), w3 = ldu32(body + 12);
  g.n_cig = w3 & 0xffffu;
  g.l_seq = ldu32(body + 16);
  g.nseq = (g.l_seq + 1u) / 2u;                 // u32 arithmetic like the reference (wraps for l_seq = 2^32 - 1)
  g.cig = (32u + (w2 & 0xffu)) & 0xffu;         // `32 + self.l_read_name()` in u8
  g.seq = g.cig + 4ull * g.n_cig;
  g.qual = g.seq + g.nseq;
  g.tags = g.qual + g.l_seq;
  return g;
}

// get_cigar (bamrawrecord.rs:126-157): where the CIGAR bytes are.  Returns false if the reference would panic.
__device__ inline bool cigar_source(const uint8_t* body, uint32_t len, const Geom& g, uint64_t* src, uint64_t* nbytes) {
  *src = 0; *nbytes = 0;
  if (g.ref < 0 || g.pos < 0 || g.n_cig == 0) return true;          // &[]
  const uint64_t fb = 4ull * g.n_cig;
  if (g.cig + fb > len) return false;                                // get_slice panics
  *src = g.cig; *nbytes = fb;
  const uint32_t first = ldu32(body + g.cig);
  if ((first & 0xfu) != 4u || (first >> 4) != g.nseq || g.n_cig != 2u) return true;
  if (g.tags > len) return false;                                    // `self.0.len() - get_tags_offset()` underflows
  uint64_t vo, vl;
  uint8_t ty;
  const int rc = tag_find(body + g.tags, len - g.tags, 'C', 'G', &vo, &vl, &ty);
  if (rc == 2) return false;
  if (rc == 1) { *src = g.tags + vo; *nbytes = vl; }                  // whatever the tag's type is, its value bytes
  return true;
}

constexpr int K5_THREADS = 256;
constexpr uint32_t K5_SEQ_IRREGULAR = 8;   // flags bit: a low nibble 0 somewhere else than in the last byte (decode_seq drops it)

// Number of bytes of w (4 packed sequence bytes) whose LOW nibble is 0.
__device__ __forceinline__ uint32_t zero_low_nibbles(uint32_t w) {
  const uint32_t t = w & 0x0f0f0f0fu;                       // every byte 0..15: bit 7 of t + 0x7f is set iff the byte is not 0
  return 4u - __popc((t + 0x7f7f7f7fu) & 0x80808080u);
}

// Pass 1, one thread per record: field lengths, where the CIGAR bytes are, panic flags.
__global__ void __launch_bounds__(K5_THREADS) k5_sizes(const uint8_t* __restrict__ data, const uint64_t* __restrict__ rec_off,
                                                       const uint32_t* __restrict__ rec_len, uint64_t n, uint32_t which,
                                                       FieldsDev f, unsigned long long* bad) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t flags = 0;
  if (i < n) {
    const uint8_t* body = data + rec_off[i];
    const uint32_t len = rec_len[i];
    uint64_t cl = 0, sl = 0, ql = 0, csrc = 0, cnb = 0;
    if (len < 32) {
      flags = which & 7u;                       // every fixed-field get_slice panics
    } else {
      const Geom g = geometry(body);
      if (which & BAMGPU_FIELD_SEQ_BIT) {
        if (g.seq + g.nseq > len) flags |= BAMGPU_FIELD_SEQ_BIT;
        else if (g.nseq) {
          const uint8_t* q = body + g.seq;
          uint32_t z = 0, k = 0;
          for (; k + 4 <= g.nseq; k += 4) z += zero_low_nibbles(ldu32(q + k));
          for (; k < g.nseq; ++k) z += (q[k] & 15u) == 0u;
          const uint32_t zlast = (q[g.nseq - 1] & 15u) == 0u;
          if (z != zlast) flags |= K5_SEQ_IRREGULAR;
          sl = 2ull * g.nseq - z;
        }
      }
      if (which & BAMGPU_FIELD_QUAL_BIT) {
        if (g.qual + g.l_seq > len) flags |= BAMGPU_FIELD_QUAL_BIT;
        else ql = g.l_seq;
      }
      if (which & BAMGPU_FIELD_CIGAR_BIT) {
        bool ok = cigar_source(body, len, g, &csrc, &cnb);
        if (ok && (cnb & 3)) ok = false;        // chunks(4): read_u32 on the short last chunk -> unwrap panic
        if (ok) {
          uint32_t badop = 0;
          for (uint64_t k = 0; k < cnb; k += 4) {
            const uint32_t v = ldu32(body + csrc + k);
            badop |= (v & 15u) > 8u;            // get_cig_op panics
            cl += dec_digits(v >> 4) + 1u;
          }
          if (badop) ok = false;
        }
        if (!ok) { flags |= BAMGPU_FIELD_CIGAR_BIT; cl = 0; cnb = 0; }
      }
    }
    if (which & BAMGPU_FIELD_CIGAR_BIT) { f.off[0][i] = cl; f.cig_src[i] = (uint32_t)csrc; f.cig_nbytes[i] = (uint32_t)cnb; }
    if (which & BAMGPU_FIELD_SEQ_BIT) f.off[1][i] = sl;
    if (which & BAMGPU_FIELD_QUAL_BIT) f.off[2][i] = ql;
    f.flags[i] = (uint8_t)flags;
  }
  for (int k = 0; k < 3; ++k) {
    const uint32_t m = __ballot_sync(0xffffffffu, (flags >> k) & 1u);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(&bad[k], (unsigned long long)__popc(m));
  }
}

// Pass 2a, one thread per record: the CIGAR strings (a handful of characters per record).
__global__ void __launch_bounds__(K5_THREADS) k5_decode_cigar(const uint8_t* __restrict__ data, const uint64_t* __restrict__ rec_off,
                                                              uint64_t n, FieldsDev f) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t nb = f.cig_nbytes[i];
  if (!nb) return;
  const uint8_t* src = data + rec_off[i] + f.cig_src[i];
  uint8_t* o = f.out[0] + f.off[0][i];
  for (uint32_t k = 0; k < nb; k += 4) {
    const uint32_t v = ldu32(src + k);
    uint32_t x = v >> 4;
    const uint32_t d = dec_digits(x);
    for (int j = (int)d - 1; j >= 0; --j) { o[j] = (uint8_t)('0' + x % 10u); x /= 10u; }
    o[d] = (uint8_t)"MIDNSHP=X???????"[v & 15u];           // get_cig_op, bamrawrecord.rs:208-221
    o += d + 1;
  }
}

// 16 bytes at any address: two aligned 128-bit loads and a byte shift (reads up to 31 bytes past q & ~15).
__device__ __forceinline__ uint4 ld16u(const uint8_t* q) {
  const uintptr_t a = (uintptr_t)q;
  const uint4 A = *(const uint4*)(a & ~(uintptr_t)15), B = *(const uint4*)((a & ~(uintptr_t)15) + 16);
  const uint32_t ws = (uint32_t)(a >> 2) & 3u, bs = (uint32_t)(a & 3) * 8;
  uint32_t w0 = A.x, w1 = A.y, w2 = A.z, w3 = A.w, w4 = B.x, w5 = B.y, w6 = B.z, w7 = B.w;
  if (ws & 2) { w0 = w2; w1 = w3; w2 = w4; w3 = w5; w4 = w6; w5 = w7; }
  if (ws & 1) { w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; }
  return make_uint4(__funnelshift_r(w0, w1, bs), __funnelshift_r(w1, w2, bs), __funnelshift_r(w2, w3, bs), __funnelshift_r(w3, w4, bs));
}

// Pass 2b: sequence strings and raw qualities.  A CTA takes K5_REC consecutive records — whose outputs are one contiguous
// range of each field's byte array — and its threads walk that range in aligned 16-byte pieces.  A coarse index in shared
// memory (record at every 2^s-th output byte, built with K5_COARSE binary searches per CTA) and a short linear scan over the
// records' output offsets find the piece's record; if the piece lies inside one record (nearly always) its source bytes come
// from one unaligned 16-byte load — a sequence piece through a 256-entry "packed byte -> two characters" table in shared
// memory — and it is stored with one 128-bit store.  So loads and stores are coalesced however short or long the reads are,
// and the cost per byte does not depend on the read length.  Both fields of a chunk are written by the same CTA, one after
// the other.
#ifndef K5_FLAT_RECORDS
#define K5_FLAT_RECORDS 256
#endif
#ifndef K5_FLAT_THREADS_N
#define K5_FLAT_THREADS_N 256
#endif
constexpr int K5_REC = K5_FLAT_RECORDS;        // records per CTA (measured 11.5 ms at 256 x 256 threads, 11.7 at 64 x 128, 15.4 at 16 x 64)
constexpr int K5_FT = K5_FLAT_THREADS_N;       // threads per CTA
constexpr int K5_COARSE = K5_FT;
static_assert(K5_REC <= K5_FT && K5_FT >= 32 && K5_FT % 32 == 0, "one thread loads one record's geometry");
struct FlatShared {
  uint64_t dst[K5_REC + 1];   // output offset of each record's field (absolute)
  uint64_t src[K5_REC];       // offset in `data` of the field's packed bytes
  uint16_t coarse[K5_COARSE + 1];
  uint8_t flag[K5_REC];
};

template <int FIELD>   // 1 sequence, 2 quality
__device__ __forceinline__ void k5_flat_field(const uint8_t* __restrict__ data, const uint64_t* __restrict__ rec_off,
                                              const uint32_t* __restrict__ rec_len, uint64_t r0, uint32_t R, const FieldsDev& f,
                                              FlatShared& S, const uint16_t* s_pair, const uint32_t* s_bytemask) {
  const uint64_t* off = f.off[FIELD];
  uint8_t* out = f.out[FIELD];
  __syncthreads();                               // the previous field is done with S
  if (threadIdx.x < R) {
    const uint64_t i = r0 + threadIdx.x;
    const uint64_t ro = rec_off[i];
    uint64_t src = ro;
    if (rec_len[i] >= 32) { const Geom g = geometry(data + ro); src = ro + (FIELD == 1 ? g.seq : g.qual); }
    S.src[threadIdx.x] = src;
    S.dst[threadIdx.x] = off[i];
    S.flag[threadIdx.x] = f.flags[i];
  }
  if (threadIdx.x == 0) S.dst[R] = off[r0 + R];
  __syncthreads();
  const uint64_t begin = S.dst[0], end = S.dst[R];
  if (begin == end) return;                      // (uniform)
  // coarse index: entry t = the record holding output byte (begin & ~15) + (t << sh)
  const uint64_t base = begin & ~15ull;
  uint32_t sh = 9;
  while (((end - base) >> sh) >= (uint64_t)K5_COARSE) ++sh;
  if (threadIdx.x <= (uint32_t)((end - base) >> sh)) {
    uint64_t q = base + ((uint64_t)threadIdx.x << sh);
    if (q < begin) q = begin;
    uint32_t a = 0, b = R;                       // invariant: dst[a] <= q < dst[b]; the last such a (empty fields share offsets)
    while (b - a > 1) { const uint32_t m = (a + b) >> 1; if (S.dst[m] <= q) a = m; else b = m; }
    S.coarse[threadIdx.x] = (uint16_t)a;
  }
  __syncthreads();
  for (uint64_t p = base + 16ull * threadIdx.x; p < end; p += 16ull * K5_FT) {
    const uint64_t lo = p < begin ? begin : p, hi = p + 16 < end ? p + 16 : end;   // bytes [lo, hi) of this piece are the chunk's
    uint32_t a = S.coarse[(p - base) >> sh];
    while (S.dst[a + 1] <= lo) ++a;              // record of byte lo (skips empty fields)
    uint32_t v0 = 0, v1 = 0, v2 = 0, v3 = 0, have = 0;   // the piece and the bitmap of its bytes produced so far
    for (;;) {                                   // every record with bytes in [lo, hi): one, sometimes two, rarely more
      const uint64_t da = S.dst[a], db = S.dst[a + 1];
      const uint64_t slo = da > lo ? da : lo, shi = db < hi ? db : hi;
      if (shi > slo && !(FIELD == 1 && (S.flag[a] & K5_SEQ_IRREGULAR))) {   // (irregular sequences: the warp loop below)
        // the 16 bytes the piece would hold if it lay wholly inside record a's field (it may begin before the field)
        const int64_t c = (int64_t)(p - da);
        uint4 v;
        if (FIELD == 2) {
          v = ld16u(data + S.src[a] + c);
        } else {
          // characters c .. c + 15 = nibbles c .. c + 15: packed bytes floor(c / 2) .. + 8
          const uint4 x = ld16u(data + S.src[a] + (c >> 1));
          uint32_t w[5];
#pragma unroll
          for (int t = 0; t < 5; ++t) {          // w[t] = the characters of packed bytes 2t, 2t + 1
            const uint32_t xx = t < 2 ? x.x : (t < 4 ? x.y : x.z), s16 = (uint32_t)(t & 1) * 16;
            w[t] = (uint32_t)s_pair[(xx >> s16) & 0xffu] | ((uint32_t)s_pair[(xx >> (s16 + 8)) & 0xffu] << 16);
          }
          const uint32_t odd = (uint32_t)(c & 1) * 8;   // an odd c starts at the low nibble of the first byte: one character later
          v = make_uint4(__funnelshift_r(w[0], w[1], odd), __funnelshift_r(w[1], w[2], odd), __funnelshift_r(w[2], w[3], odd),
                         __funnelshift_r(w[3], w[4], odd));
        }
        const uint32_t bits = ((1u << (uint32_t)(shi - p)) - 1u) & ~((1u << (uint32_t)(slo - p)) - 1u);   // bytes [slo, shi) of the piece
        if (bits == 0xffffu) { v0 = v.x; v1 = v.y; v2 = v.z; v3 = v.w; }
        else {
          const uint32_t m0 = s_bytemask[bits & 15u], m1 = s_bytemask[(bits >> 4) & 15u], m2 = s_bytemask[(bits >> 8) & 15u], m3 = s_bytemask[bits >> 12];
          v0 = (v0 & ~m0) | (v.x & m0); v1 = (v1 & ~m1) | (v.y & m1); v2 = (v2 & ~m2) | (v.z & m2); v3 = (v3 & ~m3) | (v.w & m3);
        }
        have |= bits;
      }
      if (db >= hi) break;
      ++a;
    }
    if (have == 0xffffu) { *(uint4*)(out + p) = make_uint4(v0, v1, v2, v3); continue; }
    // the ends of the chunk, and pieces that share bytes with an irregular sequence: only the bytes produced here
    for (uint32_t k = 0; k < 16; ++k)
      if ((have >> k) & 1u) { const uint32_t wv = k < 4 ? v0 : (k < 8 ? v1 : (k < 12 ? v2 : v3)); out[p + k] = (uint8_t)(wv >> (8 * (k & 3))); }
  }
  if (FIELD == 1) {
    // sequences with a low nibble 0 in the middle: character positions depend on the data -> one warp per such record
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    for (uint32_t j = warp; j < R; j += K5_FT / 32) {
      if (!(S.flag[j] & K5_SEQ_IRREGULAR)) continue;
      const uint8_t* q = data + S.src[j];
      const uint32_t nseq = (ldu32(data + rec_off[r0 + j] + 16) + 1u) / 2u;
      uint8_t* o = out + S.dst[j];
      uint64_t basec = 0;
      for (uint32_t k0 = 0; k0 < nseq; k0 += 32) {
        const uint32_t k = k0 + lane;
        const bool valid = k < nseq;
        const uint32_t by = valid ? q[k] : 0u;
        const bool two = valid && (by & 15u) != 0u;
        const uint32_t mv = __ballot_sync(0xffffffffu, valid), m2 = __ballot_sync(0xffffffffu, two);
        if (valid) {
          uint8_t* w = o + basec + __popc(mv & lt) + __popc(m2 & lt);
          const uint32_t pr = s_pair[by];
          w[0] = (uint8_t)pr;
          if (two) w[1] = (uint8_t)(pr >> 8);
        }
        basec += __popc(mv) + __popc(m2);
      }
    }
  }
}

__global__ void __launch_bounds__(K5_FT) k5_decode_flat(const uint8_t* __restrict__ data, const uint64_t* __restrict__ rec_off,
                                                             const uint32_t* __restrict__ rec_len, uint64_t n, uint32_t which, FieldsDev f) {
  __shared__ FlatShared S;
  __shared__ uint16_t s_pair[256];   // packed sequence byte -> its two characters (high nibble first): get_seq_base, :238-258
  __shared__ uint32_t s_bytemask[16];  // 4-bit byte bitmap -> 32-bit byte mask
  const uint64_t r0 = (uint64_t)blockIdx.x * K5_REC;
  const uint32_t R = (uint32_t)(n - r0 < K5_REC ? n - r0 : K5_REC);
  for (uint32_t b = threadIdx.x; b < 256; b += K5_FT) {
    s_pair[b] = (uint16_t)((uint8_t)"=ACMGRSVTWYHKDBN"[b >> 4] | ((uint8_t)"=ACMGRSVTWYHKDBN"[b & 15] << 8));
    if (b < 16) s_bytemask[b] = ((b & 1) ? 0xffu : 0u) | ((b & 2) ? 0xff00u : 0u) | ((b & 4) ? 0xff0000u : 0u) | ((b & 8) ? 0xff000000u : 0u);
  }
  if (which & BAMGPU_FIELD_SEQ_BIT) k5_flat_field<1>(data, rec_off, rec_len, r0, R, f, S, s_pair, s_bytemask);
  if (which & BAMGPU_FIELD_QUAL_BIT) k5_flat_field<2>(data, rec_off, rec_len, r0, R, f, S, s_pair, s_bytemask);
}

}  // namespace

size_t fields_scratch_bytes(uint64_t n) {
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int)(n + 1));
  return 3 * align_up(8 * (n + 1), 256) + 2 * align_up(4 * n + 4, 256) + align_up(n + 1, 256) + align_up(tmp, 256) + 256;
}

// Carves the scratch, measures the requested fields and turns the lengths into offsets (n + 1 each).  totals / bad (host,
// 3 words each) are filled after a synchronise: the caller sizes the output buffers from them.
cudaError_t fields_measure(const uint8_t* data, const ColumnsDev& c, uint64_t n, uint32_t which, void* scratch, FieldsDev* f,
                           unsigned long long* d_bad, uint64_t* totals, uint64_t* bad, int num_sms, cudaStream_t s) {
  uint8_t* p = (uint8_t*)scratch;
  for (int k = 0; k < 3; ++k) { f->off[k] = (uint64_t*)p; p += align_up(8 * (n + 1), 256); f->out[k] = nullptr; }
  f->cig_src = (uint32_t*)p; p += align_up(4 * n + 4, 256);
  f->cig_nbytes = (uint32_t*)p; p += align_up(4 * n + 4, 256);
  f->flags = p; p += align_up(n + 1, 256);
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, (const uint64_t*)nullptr, (uint64_t*)nullptr, (int)(n + 1));
  cudaMemsetAsync(d_bad, 0, 24, s);
  for (int k = 0; k < 3; ++k) { totals[k] = 0; bad[k] = 0; cudaMemsetAsync(f->off[k] + n, 0, 8, s); }
  if (n == 0) { for (int k = 0; k < 3; ++k) cudaMemsetAsync(f->off[k], 0, 8, s); return cudaStreamSynchronize(s); }
  k5_sizes<<<(unsigned)((n + K5_THREADS - 1) / K5_THREADS), K5_THREADS, 0, s>>>(data, c.rec_off, c.rec_len, n, which, *f, d_bad);
  for (int k = 0; k < 3; ++k)
    if (which & (1u << k)) {
      cub::DeviceScan::ExclusiveSum(p, tmp, (const uint64_t*)f->off[k], f->off[k], (int)(n + 1), s);
      cudaMemcpyAsync(&totals[k], f->off[k] + n, 8, cudaMemcpyDeviceToHost, s);
    }
  cudaMemcpyAsync(bad, d_bad, 24, cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  return e != cudaSuccess ? e : cudaGetLastError();
}

void launch_fields_decode(const uint8_t* data, const ColumnsDev& c, uint64_t n, uint32_t which, const FieldsDev& f, int num_sms,
                          cudaStream_t s) {
  if (!n) return;
  const unsigned grid = (unsigned)((n + K5_THREADS - 1) / K5_THREADS);
  if (which & BAMGPU_FIELD_CIGAR_BIT) k5_decode_cigar<<<grid, K5_THREADS, 0, s>>>(data, c.rec_off, n, f);
  if (which & (BAMGPU_FIELD_SEQ_BIT | BAMGPU_FIELD_QUAL_BIT)) k5_decode_flat<<<(unsigned)((n + K5_REC - 1) / K5_REC), K5_FT, 0, s>>>(data, c.rec_off, c.rec_len, n, which, f);
}

}  // namespace bamgpu
