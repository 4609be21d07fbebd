// crc32.cu — K2: CRC-32 (IEEE, reflected, as in RFC 1952) of every inflated BGZF block, checked against
// the block's footer.  The reference throws the footer away (bam_tools/src/util.rs:44,68-71); this
// path verifies it, which is a documented, stricter divergence.
//
// One warp per block.  Rows of 128 bytes are loaded fully coalesced (lane i takes word i of the row)
// and each lane keeps its own partial remainder a_i = sum_r Z_128^(R-1-r)(w[r][i]) using 4 table
// look-ups per word ("advance by 128 zero bytes" tables, replicated per lane in shared memory so that the
// data-dependent look-ups never collide on a bank).  The 32 remainders are then folded with a
// Horner chain of "advance by 4 zero bytes" steps, and the unaligned head / tail bytes are handled
// bytewise by lane 0.  CRC is GF(2)-linear, so this equals the serial CRC bit for bit.
#include "kernels.h"

namespace bamgpu {

namespace {

constexpr uint32_t POLY = 0xEDB88320u;

// tables[0..1023]: Z4 slice tables  T4[k][b] = state b<<(8k)... in the usual slicing-by-4 layout
// tables[1024..2047]: Z128 tables   T128[k][b] = advance (b << 8k) by 128 zero bytes
__global__ void k_crc_tables(uint32_t* tables) {
  uint32_t b = threadIdx.x;  // 256 threads
  // byte table
  uint32_t c = b;
  for (int i = 0; i < 8; ++i) c = (c >> 1) ^ ((c & 1) ? POLY : 0);
  __shared__ uint32_t t0[256];
  t0[b] = c;
  __syncthreads();
  // advance a 32-bit state by one zero byte: s' = t0[s & 0xff] ^ (s >> 8)
  auto adv = [&](uint32_t s, int nbytes) {
    for (int i = 0; i < nbytes; ++i) s = t0[s & 0xff] ^ (s >> 8);
    return s;
  };
  for (int k = 0; k < 4; ++k) {
    uint32_t v = b << (8 * k);
    tables[k * 256 + b] = adv(v, 4);
    tables[1024 + k * 256 + b] = adv(v, 128);
  }
}

__device__ __forceinline__ uint32_t z_apply(const uint32_t* t, uint32_t s) {
  return t[s & 0xff] ^ t[256 + ((s >> 8) & 0xff)] ^ t[512 + ((s >> 16) & 0xff)] ^ t[768 + (s >> 24)];
}
// Same on a table replicated REPL times ([entry][copy], lane l uses copy l mod REPL).  REPL = 32: lane l only ever touches
// bank l, so the four look-ups of a warp are conflict free whatever the data.
template <int REPL>
__device__ __forceinline__ uint32_t z_apply_lane(const uint32_t* t /* + copy */, uint32_t s) {
  return t[(s & 0xff) * REPL] ^ t[(256 + ((s >> 8) & 0xff)) * REPL] ^ t[(512 + ((s >> 16) & 0xff)) * REPL] ^ t[(768 + (s >> 24)) * REPL];
}

// Two shapes of the same kernel:
//   <32, 32>  one 32-warp CTA per SM, the table replicated per lane (132 KB): the whole file in one launch (bench.py's
//             device-resident step), 3.5 TB/s;
//   <4, 8>    four warps, 8 copies of the table (36 KB), <= 32 registers: small enough to sit NEXT to three resident K1
//             CTAs (186 KB, 61 k registers) - the Reader runs the CRC of one batch while the K1 of the next batches holds
//             every SM, and the big shape would wait for an SM to drain completely.
// CRC-32 of n bytes at p, computed by one warp (the value is returned in every lane).
template <int REPL>
__device__ __forceinline__ uint32_t warp_crc32(const uint8_t* p, uint32_t n, const uint32_t* Z128 /* + copy */, const uint32_t* Z4, int lane) {
  uint32_t head = (uint32_t)((4 - ((uintptr_t)p & 3)) & 3);
  if (head > n) head = n;
  uint32_t s = 0xffffffffu;
  // head bytes up to 4-byte alignment (all lanes compute the same value).  The classic byte table
  // T0[x] equals Z4 applied to x << 24, i.e. Z4's fourth slice.
  for (uint32_t i = 0; i < head; ++i) s = Z4[768 + ((s ^ p[i]) & 0xff)] ^ (s >> 8);
  const uint32_t rows = (n - head) >> 7;
  const uint32_t* w = (const uint32_t*)(p + head) + lane;
  uint32_t a = 0;
  if (rows) {
    uint32_t x = w[0];
    if (lane == 0) x ^= s;
    a = x;
    uint32_t r = 1;
    for (; r + 4 <= rows; r += 4) {   // four rows in flight: the loads do not wait for the table walk
      uint32_t x0 = w[r * 32], x1 = w[(r + 1) * 32], x2 = w[(r + 2) * 32], x3 = w[(r + 3) * 32];
      a = z_apply_lane<REPL>(Z128, a) ^ x0;
      a = z_apply_lane<REPL>(Z128, a) ^ x1;
      a = z_apply_lane<REPL>(Z128, a) ^ x2;
      a = z_apply_lane<REPL>(Z128, a) ^ x3;
    }
    for (; r < rows; ++r) a = z_apply_lane<REPL>(Z128, a) ^ w[r * 32];
    // fold the 32 lanes: total = Z4(...Z4(Z4(a0) ^ a1) ...) ^ a31, then one more Z4
    uint32_t acc = 0;
    for (int i = 0; i < 32; ++i) {
      uint32_t ai = __shfl_sync(0xffffffffu, a, i);
      acc = z_apply(Z4, acc ^ ai);
    }
    s = acc;
  }
  // tail bytes
  const uint8_t* q = p + head + (rows << 7);
  const uint32_t tail = n - head - (rows << 7);
  for (uint32_t i = 0; i < tail; ++i) s = Z4[768 + ((s ^ q[i]) & 0xff)] ^ (s >> 8);
  return ~s;
}

template <int WARPS, int REPL>
__device__ __forceinline__ void load_crc_tables(uint32_t* smem, const uint32_t* __restrict__ g_tables) {
  for (int i = threadIdx.x; i < 1024 * REPL; i += blockDim.x) smem[i] = g_tables[1024 + i / REPL];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) smem[1024 * REPL + i] = g_tables[i];
  __syncthreads();
}

// Two shapes of the same kernel:
//   <32, 32>  one 32-warp CTA per SM, the table replicated per lane (132 KB): the whole file in one launch (bench.py's
//             device-resident step), 3.5 TB/s;
//   <4, 8>    four warps, 8 copies of the table (36 KB), <= 32 registers: small enough to sit NEXT to three resident K1
//             CTAs (186 KB, 61 k registers) - the Reader runs the CRC of one batch while the K1 of the next batches holds
//             every SM, and the big shape would wait for an SM to drain completely.
template <int WARPS, int REPL>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 32 ? 1 : 16)
k2_crc32(const uint8_t* __restrict__ out, BlockTableDev t, const uint32_t* __restrict__ g_tables,
         uint32_t* __restrict__ status) {
  extern __shared__ __align__(16) uint32_t k2_smem[];
  load_crc_tables<WARPS, REPL>(k2_smem, g_tables);
  const int lane = threadIdx.x & 31;
  const uint32_t* Z128 = k2_smem + (lane % REPL);   // [1024][REPL]
  const uint32_t* Z4 = k2_smem + 1024 * REPL;       // [1024]
  const uint32_t warps_total = gridDim.x * WARPS;
  for (uint32_t b = blockIdx.x * WARPS + (threadIdx.x >> 5); b < t.n_blocks; b += warps_total) {
    if (status[b] != 0) continue;  // warp-uniform
    const uint32_t s = warp_crc32<REPL>(out + t.out_off[b], t.isize[b], Z128, Z4, lane);
    if (lane == 0 && s != t.crc[b]) status[b] = IST_CRC_MISMATCH;
  }
}

// The same warp CRC over consecutive `chunk`-byte pieces of one buffer: the footers of the BGZF blocks the writer emits.
__global__ void __launch_bounds__(32 * 32, 1)
k_crc32_chunks(const uint8_t* __restrict__ src, unsigned long long len, uint32_t chunk, const uint32_t* __restrict__ g_tables,
               uint32_t* __restrict__ crc_out) {
  extern __shared__ __align__(16) uint32_t k2_smem[];
  load_crc_tables<32, 32>(k2_smem, g_tables);
  const int lane = threadIdx.x & 31;
  const uint32_t* Z128 = k2_smem + lane;
  const uint32_t* Z4 = k2_smem + 1024 * 32;
  const unsigned long long n_chunks = (len + chunk - 1) / chunk;
  for (unsigned long long b = (unsigned long long)blockIdx.x * 32 + (threadIdx.x >> 5); b < n_chunks; b += (unsigned long long)gridDim.x * 32) {
    const unsigned long long o = b * chunk;
    const uint32_t n = (uint32_t)(len - o < chunk ? len - o : chunk);
    const uint32_t s = warp_crc32<32>(src + o, n, Z128, Z4, lane);
    if (lane == 0) crc_out[b] = s;
  }
}

}  // namespace

void crc32_init_tables(uint32_t* d_tables, cudaStream_t s) { k_crc_tables<<<1, 256, 0, s>>>(d_tables); }

void launch_crc32(const uint8_t* out, const BlockTableDev& t, const uint32_t* d_tables, uint32_t* status,
                  int num_sms, bool small_footprint, cudaStream_t s) {
  if (t.n_blocks == 0) return;
  if (small_footprint) {
    constexpr int W = 4, R = 8, SMEM = 1024 * R * 4 + 1024 * 4;
    uint32_t grid = (t.n_blocks + W - 1) / W;
    if (grid > (uint32_t)num_sms * 2) grid = (uint32_t)num_sms * 2;
    k2_crc32<W, R><<<grid, W * 32, SMEM, s>>>(out, t, d_tables, status);
    return;
  }
  constexpr int W = 32, R = 32, SMEM = 1024 * R * 4 + 1024 * 4;   // Z128 per lane (128 KB) + Z4 (4 KB)
  uint32_t grid = (t.n_blocks + W - 1) / W;
  if (grid > (uint32_t)num_sms) grid = (uint32_t)num_sms;   // one CTA per SM: the per-lane table takes 132 KB of shared memory
  cudaFuncSetAttribute(k2_crc32<W, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  k2_crc32<W, R><<<grid, W * 32, SMEM, s>>>(out, t, d_tables, status);
}

void launch_crc32_chunks(const uint8_t* src, uint64_t len, uint32_t chunk, const uint32_t* d_tables, uint32_t* crc_out, int num_sms, cudaStream_t s) {
  if (len == 0) return;
  constexpr int SMEM = 1024 * 32 * 4 + 1024 * 4;
  uint64_t n_chunks = (len + chunk - 1) / chunk;
  uint32_t grid = (uint32_t)((n_chunks + 31) / 32);
  if (grid > (uint32_t)num_sms) grid = (uint32_t)num_sms;
  cudaFuncSetAttribute(k_crc32_chunks, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  k_crc32_chunks<<<grid, 32 * 32, SMEM, s>>>(src, len, chunk, d_tables, crc_out);
}

}  // namespace bamgpu
