// Debug harness 2: real K1 kernel + launch config on N zlib-made blocks, per-block status and compare.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <zlib.h>
#include "../../bam_parallel_b200/csrc/inflate.cu"
using namespace bamgpu;

int main(int argc, char** argv) {
  int NB = argc > 1 ? atoi(argv[1]) : 64;
  std::vector<uint8_t> src((size_t)NB * 65280);
  srand(1);
  for (size_t i = 0; i < src.size(); ++i) src[i] = "ACGT!FFF:,#read"[rand() % 15];
  std::vector<uint8_t> comp;
  std::vector<uint64_t> in_off(NB), out_off(NB);
  std::vector<uint32_t> in_len(NB), isize(NB), crc(NB);
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
  cudaMalloc(&dout, src.size() + 128); cudaMemset(dout, 0xEE, src.size() + 128);
  cudaMalloc(&dio, 8 * NB); cudaMalloc(&doo, 8 * NB); cudaMalloc(&dil, 4 * NB); cudaMalloc(&dis, 4 * NB); cudaMalloc(&dst, 4 * NB); cudaMalloc(&dcnt, 4);
  cudaMemcpy(dio, in_off.data(), 8 * NB, cudaMemcpyHostToDevice); cudaMemcpy(doo, out_off.data(), 8 * NB, cudaMemcpyHostToDevice);
  cudaMemcpy(dil, in_len.data(), 4 * NB, cudaMemcpyHostToDevice); cudaMemcpy(dis, isize.data(), 4 * NB, cudaMemcpyHostToDevice);
  cudaMemset(dst, 0xff, 4 * NB); cudaMemset(dcnt, 0, 4);
  BlockTableDev t{dio, dil, doo, dis, nullptr, (uint32_t)NB};
  launch_inflate(dcomp, t, dout + 64, dst, dcnt, 148, 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<uint32_t> st(NB); cudaMemcpy(st.data(), dst, 4 * NB, cudaMemcpyDeviceToHost);
  std::vector<uint8_t> out(src.size()); cudaMemcpy(out.data(), dout + 64, out.size(), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int b = 0; b < NB; ++b) {
    bool ok = memcmp(out.data() + (size_t)b * 65280, src.data() + (size_t)b * 65280, 65280) == 0;
    if (st[b] != 0 || !ok) { if (bad++ < 40) printf("block %d status %u ok %d\n", b, st[b], ok); }
  }
  printf("%d / %d blocks bad\n", bad, NB);
  return 0;
}
