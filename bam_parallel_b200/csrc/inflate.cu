// inflate.cu — K1: raw-DEFLATE inflate of BGZF blocks on sm_100a.
//
// Replaces bam_tools/src/util.rs:10-16 (inflate_data) + reader/readahead.rs:145-150
// (decompress_block) and the n-2 rayon inflate workers of readahead.rs:88-118.
//
// Mapping: a persistent grid of 4-warp CTAs, 4 CTAs per SM (16 warps = 512 lanes per SM, the
// most that fit in 227 KB of shared memory at 384 B of Huffman tables + 32 B of input staging per lane).  Every LANE
// owns one BGZF block at a time and pulls the next block index from a global counter, so a
// warp advances 32 independent DEFLATE streams side by side.  The per-lane logic is in
// inflate_core.cuh.  Output offsets come from the host's exclusive scan of ISIZE
// (bamgpu_scan_blocks), so the inflated stream is contiguous and records may straddle blocks.
#include "inflate_core.cuh"
#include "kernels.h"

namespace bamgpu {

constexpr int K1_WARPS_PER_CTA = 4;
#ifndef K1_CTAS
#define K1_CTAS 4
#endif
constexpr int K1_CTAS_PER_SM = K1_CTAS;
constexpr int K1_TABLE_BYTES = K1_WARPS_PER_CTA * 32 * LANE_WORDS * 4;            // 40320: Huffman tables
constexpr int K1_SMEM_BYTES = K1_TABLE_BYTES + K1_WARPS_PER_CTA * STAGE_WARP_BYTES;   // + 3072: input staging slots

constexpr uint32_t K1_TOKENS_PER_ROUND = 128;  // tokens a lane decodes between two looks at the warp-wide state

__global__ void __launch_bounds__(K1_WARPS_PER_CTA * 32, K1_CTAS_PER_SM)
k1_inflate_lane_per_block(InflateJob job, uint32_t* __restrict__ work_counter) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Lane<32> L;
  L.sm = smem + warp * (32 * LANE_WORDS) + lane;
  const uint32_t stage0 = (uint32_t)__cvta_generic_to_shared(smem) + K1_TABLE_BYTES + warp * STAGE_WARP_BYTES + lane * 16;
  L.state = ST_NEXT;
  L.blk = 0xffffffffu;
  L.err = 0;
  for (;;) {
    if (L.state == ST_NEXT) {
      L.finish_block(job);
      uint32_t b = atomicAdd(work_counter, 1u);
      if (b < job.n_blocks) L.begin_block(job, b, stage0);
      else { L.state = ST_DONE; L.blk = 0xffffffffu; }
    }
    if (__all_sync(0xffffffffu, L.state == ST_DONE)) break;
    L.round(K1_TOKENS_PER_ROUND);
    __syncwarp();
  }
}

size_t inflate_smem_bytes() { return K1_SMEM_BYTES; }

void launch_inflate(const uint8_t* comp, const BlockTableDev& t, uint8_t* out, uint32_t* status,
                    uint32_t* work_counter, int num_sms, cudaStream_t s) {
  if (t.n_blocks == 0) return;
  // per-device attribute; cheap enough to set on every launch (keeps multi-device processes correct)
  cudaFuncSetAttribute(k1_inflate_lane_per_block, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM_BYTES);
  InflateJob job{comp, t.in_off, t.in_len, t.out_off, t.isize, out, status, t.n_blocks};
  uint32_t lanes_per_cta = K1_WARPS_PER_CTA * 32;
  uint32_t max_ctas = (uint32_t)num_sms * K1_CTAS_PER_SM;
  uint32_t need = (t.n_blocks + lanes_per_cta - 1) / lanes_per_cta;
  uint32_t grid = need < max_ctas ? need : max_ctas;
  k1_inflate_lane_per_block<<<grid, lanes_per_cta, K1_SMEM_BYTES, s>>>(job, work_counter);
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

void launch_status_reduce(const uint32_t* status, uint32_t n_blocks, unsigned long long* first_bad, cudaStream_t s) {
  cudaMemsetAsync(first_bad, 0xff, sizeof(unsigned long long), s);
  if (n_blocks == 0) return;
  uint32_t grid = (n_blocks + 255) / 256;
  if (grid > 592) grid = 592;
  k_status_reduce<<<grid, 256, 0, s>>>(status, n_blocks, first_bad);
}

}  // namespace bamgpu
