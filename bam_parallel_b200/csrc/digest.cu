// digest.cu — 64-bit digests of what the decode path produced, computed on the device from ITS outputs.
//
// A 50M-record decode yields 16.9 GB of record bodies and 3.7 GB of columns; comparing them with a CPU restatement of
// the reference byte for byte means moving all of it over PCIe.  Instead both sides fold every value into a few words
// with the same published function (the DIGEST SPEC) and compare those:
//   mix64(z)      = splitmix64 finaliser of z + 0x9E3779B97F4A7C15
//   h_fields(i)   = fold h = mix64(h ^ v) over [record index, stream offset of the body, block_size, the 11 fixed fields
//                   (record/bamrawrecord.rs:86-97), the 4 variable-field offsets (:61-77), coord_key, strand
//                   (sorting/comparators.rs:41-59,173-180), HI present / value (record/tags.rs:113-129)]
//   h_body(bytes) = sum over the body's 32-bit little-endian words w_k (zero padded) of mix64(w_k | k << 32)
//   fields = sum_i h_fields(i)      bodies = sum_i mix64(h_body(body_i) + i * 0x9E3779B97F4A7C15)        (mod 2^64)
//   sort result:  perm = sum_i mix64(perm[i] | i << 32), bodies as above over the gathered OUTPUT bytes (sort.rs:643)
// The sums are additive over shards (index and stream offset are global), so N GPUs add their digests.
// Inputs here are the columns K3 wrote and the bytes K1 / the gather wrote — nothing is re-derived from the input file.
#include "kernels.h"

namespace bamgpu {
namespace {

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One thread per record (grid-stride); per-warp partial sums, one atomic per warp at the end.
__global__ void __launch_bounds__(256) k_digest_fields(ColumnsDev c, unsigned long long n, unsigned long long index_base,
                                                       unsigned long long stream_base, unsigned long long* acc) {
  unsigned long long sum = 0, bytes = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    unsigned long long h = 0;
#define FOLD(v) h = mix64(h ^ (unsigned long long)(v))
    FOLD(index_base + i); FOLD(stream_base + c.rec_off[i]); FOLD(c.rec_len[i]);
    FOLD((uint32_t)c.ref_id[i]); FOLD((uint32_t)c.pos[i]); FOLD(c.l_read_name[i]); FOLD(c.mapq[i]); FOLD(c.bin[i]);
    FOLD(c.n_cigar_op[i]); FOLD(c.flag[i]); FOLD(c.l_seq[i]); FOLD((uint32_t)c.next_ref_id[i]); FOLD((uint32_t)c.next_pos[i]);
    FOLD((uint32_t)c.tlen[i]); FOLD(c.cigar_off[i]); FOLD(c.seq_off[i]); FOLD(c.qual_off[i]); FOLD(c.tags_off[i]);
    FOLD(c.coord_key[i]); FOLD(c.strand[i]);
    const bool hp = c.hi_present[i] == 1;
    FOLD(hp ? 1u : 0u); FOLD((uint32_t)(hp ? c.hi_value[i] : 0));
#undef FOLD
    sum += h;
    bytes += c.rec_len[i];
  }
  sum = warp_sum(sum);
  bytes = warp_sum(bytes);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&acc[2], sum); atomicAdd(&acc[1], bytes); }
}

// One warp per record (grid-stride over records): lanes take the body's words round robin.
// off[i] = body start; len = lens ? lens[i] : off[i + 1] - off[i] (sorted output: n + 1 offsets).
// `data` is readable a few bytes past the last body (both the inflated stream and the gather buffer are padded).
__global__ void __launch_bounds__(256) k_digest_bodies(const uint8_t* __restrict__ data, const uint64_t* __restrict__ off,
                                                       const uint32_t* __restrict__ lens, unsigned long long n,
                                                       unsigned long long index_base, unsigned long long* acc) {
  const int lane = threadIdx.x & 31;
  const unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned long long nwarps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  unsigned long long total = 0;
  for (unsigned long long i = warp; i < n; i += nwarps) {
    const uint64_t o = off[i];
    const uint32_t len = lens ? lens[i] : (uint32_t)(off[i + 1] - o);
    const uint32_t nw = (len + 3) >> 2;
    unsigned long long h = 0;
    for (uint32_t k = lane; k < nw; k += 32) {
      const uint64_t a = o + 4ull * k;
      const uint32_t* p = (const uint32_t*)(data + (a & ~3ull));
      const uint32_t sh = (uint32_t)(a & 3) * 8;
      uint32_t w = p[0];
      if (sh) w = __funnelshift_r(w, p[1], sh);
      const uint32_t left = len - 4 * k;                 // bytes of this word that belong to the body
      if (left < 4) w &= (1u << (8 * left)) - 1u;
      h += mix64((unsigned long long)w | ((unsigned long long)k << 32));
    }
    h = warp_sum(h);
    total += mix64(h + (index_base + i) * 0x9E3779B97F4A7C15ull);   // identical in every lane
  }
  if (lane == 0 && total) atomicAdd(&acc[3], total);
}

__global__ void __launch_bounds__(256) k_digest_perm(const uint32_t* __restrict__ perm, const uint64_t* __restrict__ off,
                                                     unsigned long long n, unsigned long long* acc) {
  unsigned long long sum = 0, bytes = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    sum += mix64((unsigned long long)perm[i] | (i << 32));
    bytes += off[i + 1] - off[i];
  }
  sum = warp_sum(sum);
  bytes = warp_sum(bytes);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&acc[2], sum); atomicAdd(&acc[1], bytes); }
}

}  // namespace

// acc: 4 x u64 in device memory {n_records, body_bytes, fields | perm, bodies}; zeroed here.
void launch_digest_records(const uint8_t* data, const ColumnsDev& c, uint64_t n, uint64_t index_base, uint64_t stream_base,
                           unsigned long long* acc, int num_sms, cudaStream_t s) {
  cudaMemsetAsync(acc, 0, 32, s);
  if (n == 0) return;
  k_digest_fields<<<num_sms * 4, 256, 0, s>>>(c, n, index_base, stream_base, acc);
  k_digest_bodies<<<num_sms * 8, 256, 0, s>>>(data, c.rec_off, c.rec_len, n, index_base, acc);
}

void launch_digest_sorted(const uint8_t* bodies, const uint64_t* body_off, const uint32_t* perm, uint64_t n,
                          unsigned long long* acc, int num_sms, cudaStream_t s) {
  cudaMemsetAsync(acc, 0, 32, s);
  if (n == 0) return;
  k_digest_perm<<<num_sms * 4, 256, 0, s>>>(perm, body_off, n, acc);
  k_digest_bodies<<<num_sms * 8, 256, 0, s>>>(bodies, body_off, nullptr, n, 0, acc);
}

void launch_digest_ragged(const uint8_t* bytes, const uint64_t* off, uint64_t n, unsigned long long* acc, int num_sms, cudaStream_t s) {
  cudaMemsetAsync(acc, 0, 32, s);
  if (n == 0) return;
  k_digest_bodies<<<num_sms * 8, 256, 0, s>>>(bytes, off, nullptr, n, 0, acc);
}

}  // namespace bamgpu
