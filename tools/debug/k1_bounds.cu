// Debug harness: the real K1 kernel compiled with BG_BOUNDS on a BAM file given on the command line.
#define BG_BOUNDS 1
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../bam_parallel_b200/csrc/inflate.cu"
using namespace bamgpu;
int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); size_t n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> d(n + 64, 0); if (fread(d.data(), 1, n, f) != n) return 1;
  std::vector<uint64_t> in_off, out_off; std::vector<uint32_t> in_len, isize;
  size_t pos = 0; uint64_t o = 0;
  while (pos + 18 <= n) {
    uint32_t bs = (d[pos + 16] | d[pos + 17] << 8) + 1;
    uint32_t isz = d[pos + bs - 4] | d[pos + bs - 3] << 8 | d[pos + bs - 2] << 16 | (uint32_t)d[pos + bs - 1] << 24;
    in_off.push_back(pos + 18); in_len.push_back(bs - 26); out_off.push_back(o); isize.push_back(isz);
    o += isz; pos += bs;
  }
  uint32_t NB = in_off.size();
  printf("%u blocks, %llu bytes out\n", NB, (unsigned long long)o);
  uint8_t *dcomp, *dout; uint64_t *dio, *doo; uint32_t *dil, *dis, *dst, *dcnt;
  cudaMalloc(&dcomp, d.size()); cudaMemcpy(dcomp, d.data(), d.size(), cudaMemcpyHostToDevice);
  cudaMalloc(&dout, o + 128); cudaMemset(dout, 0xEE, o + 128);
  cudaMalloc(&dio, 8 * NB); cudaMalloc(&doo, 8 * NB); cudaMalloc(&dil, 4 * NB); cudaMalloc(&dis, 4 * NB); cudaMalloc(&dst, 4 * NB); cudaMalloc(&dcnt, 4);
  cudaMemcpy(dio, in_off.data(), 8 * NB, cudaMemcpyHostToDevice); cudaMemcpy(doo, out_off.data(), 8 * NB, cudaMemcpyHostToDevice);
  cudaMemcpy(dil, in_len.data(), 4 * NB, cudaMemcpyHostToDevice); cudaMemcpy(dis, isize.data(), 4 * NB, cudaMemcpyHostToDevice);
  cudaMemset(dst, 0xff, 4 * NB); cudaMemset(dcnt, 0, 4);
  cudaFuncSetAttribute(k1_inflate_lane_per_block, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM_BYTES);
  InflateJob job{dcomp, dio, dil, doo, dis, dout + 64, dst, NB, o, n};
  uint32_t grid = (NB + 95) / 96; if (grid > 148 * K1_CTAS_PER_SM) grid = 148 * K1_CTAS_PER_SM;
  k1_inflate_lane_per_block<<<grid, 96, K1_SMEM_BYTES>>>(job, dcnt);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<uint32_t> st(NB); cudaMemcpy(st.data(), dst, 4 * NB, cudaMemcpyDeviceToHost);
    int bad = 0; for (uint32_t b = 0; b < NB; ++b) if (st[b]) { bad++; printf("%u:%u ", b % 96, st[b]); }
    printf("%d bad\n", bad);
  }
  return 0;
}
