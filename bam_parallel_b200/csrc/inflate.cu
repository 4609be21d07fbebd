// inflate.cu — K1: raw-DEFLATE inflate of BGZF blocks on sm_100a.
//
// Replaces bam_tools/src/util.rs:10-16 (inflate_data) + reader/readahead.rs:145-150
// (decompress_block) and the n-2 rayon inflate workers of readahead.rs:88-118.
//
// Mapping: a persistent grid of 8-warp CTAs, 3 CTAs per SM.  Every BGZF block is one stream, worked on by a pair of
// lanes: lane l of decoder warp w (warps 0-3: bit reader + Huffman tables -> tokens) and lane l of copier warp w + 4
// (tokens -> literals and LZ77 copies in the output), talking through an 8-token ring in shared memory.  So a CTA
// advances 128 streams, an SM 384, with 24 warps resident: half as many again as the one-lane-per-stream
// form (which was pinned to 16 warps per SM by its 128 registers and 416 B of shared memory per lane) to hide the
// Huffman dependency chains, the shared-memory look-ups and the window-load latency behind one another.  A decoder
// lane pulls its next block index from a global counter, so block sizes need not balance.  The per-lane logic is in
// inflate_core.cuh.  Output offsets come from the host's exclusive scan of ISIZE (bamgpu_scan_blocks), so the
// inflated stream is contiguous and records may straddle blocks.
#include "inflate_core.cuh"
#include "kernels.h"
#include <cstdio>
#include <cstdlib>

namespace bamgpu {

#ifndef K1_PAIR_WARPS_N
#define K1_PAIR_WARPS_N 4
#endif
constexpr int K1_PAIR_WARPS = K1_PAIR_WARPS_N;                // decoder warps per CTA (= copier warps per CTA); a multiple of 4 (setmaxnreg works on warpgroups)
constexpr int K1_THREADS = 2 * K1_PAIR_WARPS * 32;
#ifndef K1_CTAS
#define K1_CTAS 3
#endif
constexpr int K1_CTAS_PER_SM = K1_CTAS;
constexpr int K1_TABLE_BYTES = K1_PAIR_WARPS * 32 * LANE_WORDS * 4;                // 49152: Huffman tables
constexpr int K1_STAGE_BYTES = K1_PAIR_WARPS * STAGE_WARP_BYTES;                   //  4096: input staging slots
constexpr int K1_RING_BYTES = K1_PAIR_WARPS * RING_WARP_BYTES;                     //  8704: token rings (8 tokens)
constexpr int K1_SMEM_BYTES = K1_TABLE_BYTES + K1_STAGE_BYTES + K1_RING_BYTES;   // 61952: 3 CTAs fit the 196 KB carve-out, leaving 60 KB of L1

// The kernel is compiled for 80 registers per thread (3 CTAs x 256 threads per SM); the copier warpgroup gives 32 of
// them back (setmaxnreg.dec) and the decoder warpgroup, whose Huffman state lives in registers, may take them.
constexpr int K1_DECODER_REGS = K1_CTAS * K1_THREADS > 768 ? 80 : 112, K1_COPIER_REGS = 48;

// Copier scheduling.  A copier trip costs the same issue slots whether 3 or 30 of its lanes have something to do, so
// the warp sleeps until K1_MIN_READY lanes have a token or a copy in flight, some ring holds K1_URGENT tokens (never
// hold a decoder up), or it has slept K1_WAIT_LIMIT times (the tail of the job, where few lanes are live).  Measured
// on the 50 M-record file every setting between "no batching" and (28, 14, 256) lands within 2 % (71.8-73.4 ms): the
// decoders' dependency chains, not the copiers' issue slots, bound K1.  tools/debug/k1_sweep.sh reruns it.
#ifndef K1_URGENT
#define K1_URGENT (RING_TOKENS >= 8 ? RING_TOKENS / 2 : RING_TOKENS)
#endif
#ifndef K1_MIN_READY
#define K1_MIN_READY 20
#endif
#ifndef K1_IDLE_NS
#define K1_IDLE_NS 100     // sleep when no lane has anything to do
#endif
#ifndef K1_SLEEP_NS
#define K1_SLEEP_NS 100    // sleep while waiting for more lanes
#endif
#ifndef K1_WAIT_LIMIT
#define K1_WAIT_LIMIT 16
#endif
constexpr uint32_t K1_URGENT_FILL = K1_URGENT;
constexpr int K1_COPIER_MIN_READY = K1_MIN_READY;
constexpr uint32_t K1_COPIER_WAIT_LIMIT = K1_WAIT_LIMIT;
#ifndef K1_ROUND_TRIPS
#define K1_ROUND_TRIPS 128
#endif
constexpr uint32_t K1_TRIPS_PER_ROUND = K1_ROUND_TRIPS;  // trips a decoder lane makes between two looks at the warp-wide state

__global__ void __launch_bounds__(K1_THREADS, K1_CTAS_PER_SM)
k1_inflate_pairs(InflateJob job, uint32_t* __restrict__ work_counter, uint8_t* __restrict__ ring_mem) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pw = warp % K1_PAIR_WARPS;
  const bool is_decoder = warp < K1_PAIR_WARPS;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
  TokRing ring;
#if BG_RING_GLOBAL
  ring.init_global(ring_mem + ((size_t)(blockIdx.x * K1_PAIR_WARPS + pw) * 32 + lane) * RING_GLOBAL_STREAM_BYTES);
#else
  ring.init(smem_base + K1_TABLE_BYTES + K1_STAGE_BYTES + pw * RING_WARP_BYTES, lane);
#endif
  if (is_decoder) ring.reset();
  __syncthreads();
  if (is_decoder) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(K1_DECODER_REGS));
    Decoder<32> D;
    D.sm = smem + pw * (32 * LANE_WORDS) + lane;
    D.sma = smem_base + (pw * (32 * LANE_WORDS) + lane) * 4;
    D.ring = ring;
    D.head = 0;
    D.tail_seen = 0;
    D.state = ST_NEXT;
    D.blk = 0xffffffffu;
    D.err = 0;
    const uint32_t stage0 = smem_base + K1_TABLE_BYTES + pw * STAGE_WARP_BYTES + lane * 16;
    // (Assigning the first block of every stream statically, interleaved over the CTAs so that every CTA gets the same share, was
    // measured 27 % SLOWER on the 50 M-record file - 82.7 vs 65.1 ms: the lanes of a warp then work 29 MB apart instead of on
    // neighbouring blocks.  The counter hands neighbouring blocks to neighbouring lanes.)
    for (;;) {
      if (D.state == ST_NEXT) {
        uint32_t b = atomicAdd(work_counter, 1u);
        if (b < job.n_blocks) D.begin_block(job, b, stage0);
        else D.state = ST_FINISH;
      }
      if (__all_sync(0xffffffffu, D.state == ST_DONE)) break;
      if (D.state == ST_HEADER) D.header();
      __syncwarp();
      for (uint32_t i = 0; i < K1_TRIPS_PER_ROUND; ++i) {
        const bool worked = D.trip(job, stage0);
        if (!__any_sync(0xffffffffu, worked)) {
          // every lane is waiting (ring full, or for the round to end): leave the issue slots to the copiers
          if (__all_sync(0xffffffffu, D.state == ST_HEADER || D.state == ST_NEXT || D.state == ST_DONE)) break;
          __nanosleep(100);
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(K1_COPIER_REGS));
    Copier C;
    C.ring = ring;
    C.tail = 0;
    C.done = 0;
    C.p_len = 0;
    C.blk = 0xffffffffu;
    uint32_t waited = 0;
    for (;;) {
      uint32_t w0, w1, u0, u1;
      const bool have = !C.done && C.ring.peek(C.tail, w0, w1);
      const uint32_t ready = __ballot_sync(0xffffffffu, C.p_len != 0 || have);
      if (ready == 0) {
        if (__all_sync(0xffffffffu, C.done != 0)) break;
        __nanosleep(K1_IDLE_NS);
        continue;
      }
      if (__popc(ready) < K1_COPIER_MIN_READY && ++waited < K1_COPIER_WAIT_LIMIT) {
        // is any ring at least half full?
        const bool urgent = have && C.ring.peek(C.tail + K1_URGENT_FILL - 1, u0, u1);
        if (!__any_sync(0xffffffffu, urgent)) {
          __nanosleep(K1_SLEEP_NS);
          continue;
        }
      }
      waited = 0;
      C.trip(job, have, w0, w1);
    }
  }
}

size_t inflate_smem_bytes() { return K1_SMEM_BYTES; }
// CTAs for n_blocks: just enough streams for every block, at most the resident 3 per SM.  (Rounding up to a whole number of CTAs per
// SM changes nothing - 9.6 / 11.0 / 64.9 ms for 14 k / 32 k / 259 k blocks either way: the block counter gives the first CTAs full
// loads whatever the grid, and the extra ones leave at once.)
uint32_t inflate_grid(uint32_t n_blocks, int num_sms) {
  const uint32_t lanes_per_cta = K1_PAIR_WARPS * 32, max_ctas = (uint32_t)num_sms * K1_CTAS_PER_SM;
  const uint32_t need = (n_blocks + lanes_per_cta - 1) / lanes_per_cta;
  return need < max_ctas ? need : max_ctas;
}
size_t inflate_ring_bytes(int num_sms) { return BG_RING_GLOBAL ? (size_t)num_sms * K1_CTAS_PER_SM * K1_PAIR_WARPS * 32 * RING_GLOBAL_STREAM_BYTES : 0; }

void launch_inflate(const uint8_t* comp, const BlockTableDev& t, uint8_t* out, uint32_t* status,
                    uint32_t* work_counter, void* ring_mem, int num_sms, cudaStream_t s) {
  if (t.n_blocks == 0) return;
  // per-device attribute; cheap enough to set on every launch (keeps multi-device processes correct)
  cudaFuncSetAttribute(k1_inflate_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM_BYTES);
#ifdef K1_CARVEOUT
  cudaFuncSetAttribute(k1_inflate_pairs, cudaFuncAttributePreferredSharedMemoryCarveout, K1_CARVEOUT);   // experiments: percent of the 228 KB
#endif
  InflateJob job{comp, t.in_off, t.in_len, t.out_off, t.isize, out, status, t.n_blocks};
  const uint32_t grid = inflate_grid(t.n_blocks, num_sms);
  static bool once = false;
  if (!once && getenv("BAMGPU_DEBUG")) {
    once = true;
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k1_inflate_pairs, K1_THREADS, K1_SMEM_BYTES);
    fprintf(stderr, "[bamgpu] K1: %d threads/CTA, %d B smem/CTA, %d CTAs/SM resident (wanted %d), ring %u tokens\n", K1_THREADS,
            K1_SMEM_BYTES, nb, K1_CTAS_PER_SM, RING_TOKENS);
  }
  k1_inflate_pairs<<<grid, K1_THREADS, K1_SMEM_BYTES, s>>>(job, work_counter, (uint8_t*)ring_mem);
}

// ---- first failing block ----------------------------------------------------------------------
__global__ void k_status_reduce(const uint32_t* __restrict__ status, uint32_t n, unsigned long long* first_bad) {
  unsigned long long best = ~0ull;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t st = status[i];
    if (st) { unsigned long long v = ((unsigned long long)i << 32) | st; if (v < best) best = v; }
  }
  for (int o = 16; o; o >>= 1) {
    unsigned long long x = __shfl_xor_sync(0xffffffffu, best, o);
    if (x < best) best = x;
  }
  if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(first_bad, best);
}

__global__ void k_copy_small(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint32_t words) {
  for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
}
void launch_copy_small_to_host(void* h_pinned, const void* d_src, uint32_t bytes, cudaStream_t s) {
  k_copy_small<<<1, 64, 0, s>>>((uint32_t*)h_pinned, (const uint32_t*)d_src, bytes / 4);
}

void launch_status_reduce(const uint32_t* status, uint32_t n_blocks, unsigned long long* first_bad, cudaStream_t s) {
  cudaMemsetAsync(first_bad, 0xff, sizeof(unsigned long long), s);
  if (n_blocks == 0) return;
  uint32_t grid = (n_blocks + 255) / 256;
  if (grid > 592) grid = 592;
  k_status_reduce<<<grid, 256, 0, s>>>(status, n_blocks, first_bad);
}

}  // namespace bamgpu
