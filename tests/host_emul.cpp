// host_emul.cpp — runs the K1 per-stream DEFLATE logic (bam_parallel_b200/csrc/inflate_core.cuh: Decoder + Copier
// + the token ring between them) on the CPU, so the decoder can be unit-tested on the GPU-less build box.
// TEST ONLY: the product library never links this; there is no CPU decode path in the product.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../bam_parallel_b200/csrc/inflate_core.cuh"

using namespace bamgpu;

extern "C" int host_emul_inflate(const uint8_t* comp, const uint64_t* in_off, const uint32_t* in_len,
                                 const uint64_t* out_off, const uint32_t* isize, uint8_t* out /* 16B lead pad before */,
                                 uint32_t* status, uint32_t n_blocks, int lanes) {
  InflateJob job{comp, in_off, in_len, out_off, isize, out, status, n_blocks};
  if (lanes < 1) lanes = 1;
  std::vector<std::vector<uint32_t>> sm(lanes, std::vector<uint32_t>(LANE_WORDS, 0xDEADBEEFu));
  std::vector<std::vector<uint32_t>> rings(lanes, std::vector<uint32_t>(2 * RING_TOKENS + 1, 0));
  std::vector<Decoder<1>> D(lanes);
  std::vector<Copier> C(lanes);
  uint32_t counter = 0;
  for (int i = 0; i < lanes; ++i) {
    D[i].sm = sm[i].data();
    D[i].sma = 0;
    D[i].ring.mem = rings[i].data();
    D[i].ring.reset();
    D[i].head = 0;
    D[i].tail_seen = 0;
    D[i].state = ST_NEXT;
    D[i].blk = 0xffffffffu;
    D[i].err = 0;
    C[i].ring.mem = rings[i].data();
    C[i].tail = 0;
    C[i].done = 0;
    C[i].p_len = 0;
    C[i].blk = 0xffffffffu;
  }
  // Decoder and copier lanes take turns with odd, different trip counts per turn, so that the ring is seen full,
  // empty and everything in between, and every resume point is exercised.
  for (uint64_t round = 0;; ++round) {
    bool all_done = true;
    for (int i = 0; i < lanes; ++i) {
      if (D[i].state == ST_NEXT) {
        uint32_t b = counter++;
        if (b < n_blocks) D[i].begin_block(job, b, 0);
        else D[i].state = ST_FINISH;
      }
      if (D[i].state == ST_HEADER) D[i].header();
      const int dt = 1 + (int)((round * 7 + i) % 23);
      for (int k = 0; k < dt; ++k) D[i].trip(job, 0);
      const int ct = 1 + (int)((round * 5 + i * 3) % 19);
      for (int k = 0; k < ct; ++k) {
        uint32_t w0 = 0, w1 = 0;
        const bool have = !C[i].done && C[i].ring.peek(C[i].tail, w0, w1);   // taken only if no copy is left pending
        C[i].trip(job, have, w0, w1);
      }
      if (!C[i].done) all_done = false;
    }
    if (all_done) break;
    if (round > (1ull << 40)) return -1;
  }
  return 0;
}
