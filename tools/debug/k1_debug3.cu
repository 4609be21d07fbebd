#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <zlib.h>
#include "../../bam_parallel_b200/csrc/inflate_core.cuh"
using namespace bamgpu;

template <int V>
__global__ void __launch_bounds__(96, 3) kv(InflateJob job, uint32_t* __restrict__ work_counter) {
  extern __shared__ uint32_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Lane<32> L;
  L.sm = smem + warp * (32 * LANE_WORDS) + lane;
  L.state = ST_NEXT;
  L.blk = 0xffffffffu;
  L.err = 0;
  uint8_t lens[320];
  if (V == 1) for (int i = 0; i < 320; ++i) lens[i] = 0xAA;
  for (;;) {
    if (L.state == ST_NEXT) {
      L.finish_block(job);
      uint32_t b;
      if (V == 4) { b = (L.blk == 0xffffffffu) ? blockIdx.x * blockDim.x + threadIdx.x : 0xfffffff0u; }
      else b = atomicAdd(work_counter, 1u);
      if (b < job.n_blocks) L.begin_block(job, b);
      else { L.state = ST_DONE; L.blk = 0xffffffffu; }
    }
    if (__all_sync(0xffffffffu, L.state == ST_DONE)) break;
    if (V == 7) __syncwarp();
    if (V == 3) { if (L.state != ST_DONE) L.step(lens); }
    else L.step(lens);
    if (V == 5) __syncwarp();
    if (V == 9) asm volatile("" ::: "memory");
  }
}

template <int V>
__global__ void kv_nolb(InflateJob job, uint32_t* __restrict__ work_counter) {
  extern __shared__ uint32_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Lane<32> L;
  L.sm = smem + warp * (32 * LANE_WORDS) + lane;
  L.state = ST_NEXT;
  L.blk = 0xffffffffu;
  L.err = 0;
  uint8_t lens[320];
  for (;;) {
    if (L.state == ST_NEXT) {
      L.finish_block(job);
      uint32_t b = atomicAdd(work_counter, 1u);
      if (b < job.n_blocks) L.begin_block(job, b);
      else { L.state = ST_DONE; L.blk = 0xffffffffu; }
    }
    if (__all_sync(0xffffffffu, L.state == ST_DONE)) break;
    L.step(lens);
  }
}

int main(int argc, char** argv) {
  int NB = argc > 1 ? atoi(argv[1]) : 64;
  std::vector<uint8_t> src((size_t)NB * 65280);
  srand(1);
  for (size_t i = 0; i < src.size(); ++i) src[i] = "ACGT!FFF:,#read"[rand() % 15];
  std::vector<uint8_t> comp;
  std::vector<uint64_t> in_off(NB), out_off(NB);
  std::vector<uint32_t> in_len(NB), isize(NB);
  z_stream zs; memset(&zs, 0, sizeof zs);
  deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
  for (int b = 0; b < NB; ++b) {
    size_t base = comp.size();
    comp.resize(base + 18 + 70000);
    deflateReset(&zs);
    zs.next_in = src.data() + (size_t)b * 65280; zs.avail_in = 65280; zs.next_out = comp.data() + base + 18; zs.avail_out = 70000;
    deflate(&zs, Z_FINISH);
    uint32_t clen = 70000 - zs.avail_out;
    in_off[b] = base + 18; in_len[b] = clen; out_off[b] = (uint64_t)b * 65280; isize[b] = 65280;
    comp.resize(base + 18 + clen + 8);
  }
  comp.resize(comp.size() + 64);
  uint8_t *dcomp, *dout; uint64_t *dio, *doo; uint32_t *dil, *dis, *dst, *dcnt;
  cudaMalloc(&dcomp, comp.size()); cudaMemcpy(dcomp, comp.data(), comp.size(), cudaMemcpyHostToDevice);
  cudaMalloc(&dout, src.size() + 128);
  cudaMalloc(&dio, 8 * NB); cudaMalloc(&doo, 8 * NB); cudaMalloc(&dil, 4 * NB); cudaMalloc(&dis, 4 * NB); cudaMalloc(&dst, 4 * NB); cudaMalloc(&dcnt, 4);
  cudaMemcpy(dio, in_off.data(), 8 * NB, cudaMemcpyHostToDevice); cudaMemcpy(doo, out_off.data(), 8 * NB, cudaMemcpyHostToDevice);
  cudaMemcpy(dil, in_len.data(), 4 * NB, cudaMemcpyHostToDevice); cudaMemcpy(dis, isize.data(), 4 * NB, cudaMemcpyHostToDevice);
  InflateJob job{dcomp, dio, dil, doo, dis, dout + 64, dst, (uint32_t)NB};
  const int SM = 3 * 32 * LANE_WORDS * 4;
  uint32_t grid = (NB + 95) / 96;
  for (int v = 0; v <= 10; ++v) {
    cudaMemset(dout, 0xEE, src.size() + 128); cudaMemset(dst, 0xff, 4 * NB); cudaMemset(dcnt, 0, 4);
    switch (v) {
      case 0: cudaFuncSetAttribute(kv<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM); kv<0><<<grid, 96, SM>>>(job, dcnt); break;
      case 1: cudaFuncSetAttribute(kv<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM); kv<1><<<grid, 96, SM>>>(job, dcnt); break;
      case 2: cudaFuncSetAttribute(kv_nolb<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM); kv_nolb<0><<<grid, 96, SM>>>(job, dcnt); break;
      case 3: cudaFuncSetAttribute(kv<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM); kv<3><<<grid, 96, SM>>>(job, dcnt); break;
      case 4: cudaFuncSetAttribute(kv<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM); kv<4><<<grid, 96, SM>>>(job, dcnt); break;
      case 5: cudaFuncSetAttribute(kv<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM); kv<5><<<grid, 96, SM>>>(job, dcnt); break;
      case 7: cudaFuncSetAttribute(kv<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM); kv<7><<<grid, 96, SM>>>(job, dcnt); break;
      case 9: cudaFuncSetAttribute(kv<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM); kv<9><<<grid, 96, SM>>>(job, dcnt); break;
      case 10: cudaFuncSetAttribute(kv<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM); kv<0><<<NB, 1, SM>>>(job, dcnt); break;
      case 6: cudaFuncSetAttribute(kv<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM); kv<0><<<grid, 32, SM>>>(job, dcnt); break;
    }
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint32_t> st(NB); cudaMemcpy(st.data(), dst, 4 * NB, cudaMemcpyDeviceToHost);
    std::vector<uint8_t> out(src.size()); cudaMemcpy(out.data(), dout + 64, out.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int b = 0; b < NB; ++b) {
      bool ok = memcmp(out.data() + (size_t)b * 65280, src.data() + (size_t)b * 65280, 65280) == 0;
      if (st[b] != 0 || !ok) ++bad;
    }
    printf("variant %d: %s, %d / %d blocks bad, status[0]=%u\n", v, cudaGetErrorString(e), bad, NB, st[0]);
  }
  return 0;
}
