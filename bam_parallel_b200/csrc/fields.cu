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
// Two passes, one warp per record: k5_sizes measures every field (the lanes share a record's operations / bytes), a CUB
// exclusive scan turns the lengths into offsets, k5_decode writes the strings.  Byte work bound by HBM: per record it reads
// the packed fields once per pass and writes 1 + 2 + 1 bytes per base.
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

__global__ void __launch_bounds__(K5_THREADS) k5_sizes(const uint8_t* __restrict__ data, const uint64_t* __restrict__ rec_off,
                                                       const uint32_t* __restrict__ rec_len, uint64_t n, uint32_t which,
                                                       FieldsDev f, unsigned long long* bad) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  uint32_t nbad_c = 0, nbad_s = 0, nbad_q = 0;
  for (uint64_t i = warp; i < n; i += nwarps) {
    const uint8_t* body = data + rec_off[i];
    const uint32_t len = rec_len[i];
    uint32_t flags = 0;
    uint64_t cl = 0, sl = 0, ql = 0, csrc = 0, cnb = 0;
    if (len < 32) {
      flags = which & 7u;                       // every fixed-field get_slice panics
    } else {
      const Geom g = geometry(body);
      if (which & BAMGPU_FIELD_SEQ_BIT) {
        if (g.seq + g.nseq > len) flags |= BAMGPU_FIELD_SEQ_BIT;
        else {
          uint32_t z = 0;
          for (uint32_t k = lane; k < g.nseq; k += 32) z += (body[g.seq + k] & 15u) == 0u;
          sl = 2ull * g.nseq - warp_sum64(z);
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
          uint32_t c = 0, badop = 0;
          for (uint64_t k = lane; k < cnb / 4; k += 32) {
            const uint32_t v = ldu32(body + csrc + 4 * k);
            badop |= (v & 15u) > 8u;
            c += dec_digits(v >> 4) + 1u;
          }
          cl = warp_sum64(c);
          if (__any_sync(0xffffffffu, badop)) ok = false;   // get_cig_op panics
        }
        if (!ok) { flags |= BAMGPU_FIELD_CIGAR_BIT; cl = 0; cnb = 0; }
      }
    }
    if (lane == 0) {
      if (which & BAMGPU_FIELD_CIGAR_BIT) { f.off[0][i] = cl; f.cig_src[i] = (uint32_t)csrc; f.cig_nbytes[i] = (uint32_t)cnb; }
      if (which & BAMGPU_FIELD_SEQ_BIT) f.off[1][i] = sl;
      if (which & BAMGPU_FIELD_QUAL_BIT) f.off[2][i] = ql;
      f.flags[i] = (uint8_t)flags;
      nbad_c += (flags & BAMGPU_FIELD_CIGAR_BIT) != 0; nbad_s += (flags & BAMGPU_FIELD_SEQ_BIT) != 0; nbad_q += (flags & BAMGPU_FIELD_QUAL_BIT) != 0;
    }
  }
  if (lane == 0) {
    if (nbad_c) atomicAdd(&bad[0], (unsigned long long)nbad_c);
    if (nbad_s) atomicAdd(&bad[1], (unsigned long long)nbad_s);
    if (nbad_q) atomicAdd(&bad[2], (unsigned long long)nbad_q);
  }
}

__global__ void __launch_bounds__(K5_THREADS) k5_decode(const uint8_t* __restrict__ data, const uint64_t* __restrict__ rec_off,
                                                        const uint32_t* __restrict__ rec_len, uint64_t n, uint32_t which, FieldsDev f) {
  __shared__ uint8_t base_tab[16], op_tab[16];
  if (threadIdx.x < 16) {
    base_tab[threadIdx.x] = (uint8_t)"=ACMGRSVTWYHKDBN"[threadIdx.x];   // get_seq_base, bamrawrecord.rs:238-258
    op_tab[threadIdx.x] = (uint8_t)"MIDNSHP=X???????"[threadIdx.x];     // get_cig_op, :208-221
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t i = warp; i < n; i += nwarps) {
    const uint8_t* body = data + rec_off[i];
    const uint32_t len = rec_len[i], flags = f.flags[i];
    if (len < 32) continue;
    const Geom g = geometry(body);
    if ((which & BAMGPU_FIELD_CIGAR_BIT) && !(flags & BAMGPU_FIELD_CIGAR_BIT)) {
      const uint64_t src = f.cig_src[i], nops = f.cig_nbytes[i] / 4;
      uint8_t* o = f.out[0] + f.off[0][i];
      uint32_t base = 0;
      for (uint64_t k0 = 0; k0 < nops; k0 += 32) {
        const uint64_t k = k0 + lane;
        const bool valid = k < nops;
        const uint32_t v = valid ? ldu32(body + src + 4 * k) : 0u;
        const uint32_t L = valid ? dec_digits(v >> 4) + 1u : 0u;
        uint32_t incl = L;
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        if (valid) {
          uint8_t* q = o + base + (incl - L);
          q[L - 1] = op_tab[v & 15u];
          uint32_t x = v >> 4;
          for (int j = (int)L - 2; j >= 0; --j) { q[j] = (uint8_t)('0' + x % 10u); x /= 10u; }
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    if ((which & BAMGPU_FIELD_SEQ_BIT) && !(flags & BAMGPU_FIELD_SEQ_BIT)) {
      uint8_t* o = f.out[1] + f.off[1][i];
      uint64_t base = 0;
      for (uint32_t k0 = 0; k0 < g.nseq; k0 += 32) {
        const uint32_t k = k0 + lane;
        const bool valid = k < g.nseq;
        const uint32_t b = valid ? body[g.seq + k] : 0u;
        const bool two = valid && (b & 15u) != 0u;
        const uint32_t mv = __ballot_sync(0xffffffffu, valid), m2 = __ballot_sync(0xffffffffu, two);
        if (valid) {
          uint8_t* q = o + base + __popc(mv & lt) + __popc(m2 & lt);
          q[0] = base_tab[b >> 4];
          if (two) q[1] = base_tab[b & 15u];
        }
        base += __popc(mv) + __popc(m2);
      }
    }
    if ((which & BAMGPU_FIELD_QUAL_BIT) && !(flags & BAMGPU_FIELD_QUAL_BIT)) {
      uint8_t* o = f.out[2] + f.off[2][i];
      const uint8_t* q = body + g.qual;
      for (uint32_t k = lane; k < g.l_seq; k += 32) o[k] = q[k];
    }
  }
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
  k5_sizes<<<num_sms * 8, K5_THREADS, 0, s>>>(data, c.rec_off, c.rec_len, n, which, *f, d_bad);
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
  if (n) k5_decode<<<num_sms * 8, K5_THREADS, 0, s>>>(data, c.rec_off, c.rec_len, n, which, f);
}

}  // namespace bamgpu
