// Debug harness: runs the K1 lane logic for ONE block on device (1 thread) and on host, dumps lens/state.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <zlib.h>
#include "../../bam_parallel_b200/csrc/inflate_core.cuh"
using namespace bamgpu;

struct Dump { uint8_t lens[320]; uint32_t state, err, hl; uint32_t llim[16], dlim[16]; uint32_t sm[LANE_WORDS]; uint32_t steps; uint64_t opos; uint32_t bits; };

template <int STRIDE>
BG_HD void run(InflateJob job, uint32_t* smbase, Dump* d, uint32_t max_steps) {
  Lane<STRIDE> L;
  L.sm = smbase;
  L.state = ST_NEXT; L.blk = 0xffffffffu; L.err = 0;
  uint8_t lens[320];
  for (int i = 0; i < 320; ++i) lens[i] = 0xAA;
  L.begin_block(job, 0);
  uint32_t steps = 0;
  L.step(lens);  // header
  for (int i = 0; i < 320; ++i) d->lens[i] = lens[i];
  d->state = L.state; d->err = L.err;
  for (int i = 0; i < 16; ++i) { d->llim[i] = L.llim[i]; d->dlim[i] = L.dlim[i]; }
  for (int i = 0; i < LANE_WORDS; ++i) d->sm[i] = L.smw(i);
  while (L.state != ST_NEXT && steps < max_steps) { L.step(lens); ++steps; }
  L.finish_block(job);
  d->steps = steps; d->opos = L.opos; d->bits = L.bits_consumed();
  d[1].err = L.err; d[1].state = L.state;
}

__global__ void k(InflateJob job, Dump* d, uint32_t max_steps) {
  extern __shared__ uint32_t smem[];
  if (threadIdx.x == 0) run<32>(job, smem, d, max_steps);
}

int main() {
  // make a BAM-ish payload
  std::vector<uint8_t> src(65280);
  srand(1);
  for (size_t i = 0; i < src.size(); ++i) src[i] = "ACGT!FFF:,#read"[rand() % 15];
  std::vector<uint8_t> comp(70000 + 64, 0);
  z_stream zs; memset(&zs, 0, sizeof zs);
  deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
  zs.next_in = src.data(); zs.avail_in = src.size(); zs.next_out = comp.data() + 18; zs.avail_out = 70000;
  deflate(&zs, Z_FINISH);
  uint32_t clen = 70000 - zs.avail_out;
  printf("clen=%u\n", clen);
  uint64_t in_off = 18, out_off = 0; uint32_t in_len = clen, isize = src.size();
  // host
  std::vector<uint8_t> hout(64 + 65536 + 64, 0xEE);
  uint32_t hstatus = 0xffffffff;
  InflateJob hj{comp.data(), &in_off, &in_len, &out_off, &isize, hout.data() + 64, &hstatus, 1};
  std::vector<uint32_t> hsm(LANE_WORDS, 0);
  Dump hd[2]; memset(hd, 0, sizeof hd);
  run<1>(hj, hsm.data(), hd, 1u << 30);
  printf("host: state=%u err=%u steps=%u opos=%llu bits=%u status=%u ok=%d\n", hd[0].state, hd[0].err, hd[0].steps, (unsigned long long)hd[0].opos, hd[0].bits, hstatus, memcmp(hout.data() + 64, src.data(), src.size()) == 0);
  // device
  uint8_t *dcomp, *dout; uint64_t *dio, *doo; uint32_t *dil, *dis, *dst; Dump* dd;
  cudaMalloc(&dcomp, comp.size()); cudaMemcpy(dcomp, comp.data(), comp.size(), cudaMemcpyHostToDevice);
  cudaMalloc(&dout, hout.size()); cudaMemset(dout, 0xEE, hout.size());
  cudaMalloc(&dio, 8); cudaMalloc(&doo, 8); cudaMalloc(&dil, 4); cudaMalloc(&dis, 4); cudaMalloc(&dst, 4); cudaMalloc(&dd, 2 * sizeof(Dump));
  cudaMemcpy(dio, &in_off, 8, cudaMemcpyHostToDevice); cudaMemcpy(doo, &out_off, 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dil, &in_len, 4, cudaMemcpyHostToDevice); cudaMemcpy(dis, &isize, 4, cudaMemcpyHostToDevice);
  cudaMemset(dst, 0xff, 4); cudaMemset(dd, 0, 2 * sizeof(Dump));
  InflateJob dj{dcomp, dio, dil, doo, dis, dout + 64, dst, 1};
  k<<<1, 32, 32 * LANE_WORDS * 4>>>(dj, dd, 1u << 30);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  Dump gd[2]; cudaMemcpy(gd, dd, sizeof gd, cudaMemcpyDeviceToHost);
  uint32_t gstatus; cudaMemcpy(&gstatus, dst, 4, cudaMemcpyDeviceToHost);
  std::vector<uint8_t> gout(hout.size()); cudaMemcpy(gout.data(), dout, gout.size(), cudaMemcpyDeviceToHost);
  printf("dev : state=%u err=%u steps=%u opos=%llu bits=%u status=%u ok=%d\n", gd[0].state, gd[0].err, gd[0].steps, (unsigned long long)gd[0].opos, gd[0].bits, gstatus, memcmp(gout.data() + 64, src.data(), src.size()) == 0);
  int nd = 0;
  for (int i = 0; i < 320; ++i) if (gd[0].lens[i] != hd[0].lens[i]) { if (nd++ < 20) printf("lens[%d] host %u dev %u\n", i, hd[0].lens[i], gd[0].lens[i]); }
  for (int i = 0; i < 16; ++i) if (gd[0].llim[i] != hd[0].llim[i] || gd[0].dlim[i] != hd[0].dlim[i]) printf("lim[%d] host %u/%u dev %u/%u\n", i, hd[0].llim[i], hd[0].dlim[i], gd[0].llim[i], gd[0].dlim[i]);
  int ns = 0;
  for (int i = 0; i < LANE_WORDS; ++i) if (gd[0].sm[i] != hd[0].sm[i]) { if (ns++ < 20) printf("sm[%d] host %08x dev %08x\n", i, hd[0].sm[i], gd[0].sm[i]); }
  printf("lens diffs %d, sm diffs %d\n", nd, ns);
  size_t fd = 0; for (; fd < src.size(); ++fd) if (gout[64 + fd] != src[fd]) break;
  printf("first output diff at %zu\n", fd);
  return 0;
}
