// host_emul.cpp — runs the K1 per-lane DEFLATE logic (bam_parallel_b200/csrc/inflate_core.cuh) on the
// CPU, one lane at a time, so the decoder can be unit-tested on the GPU-less build box.
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
  const uint32_t budget = 7;  // odd and small: exercises every resume point
  InflateJob job{comp, in_off, in_len, out_off, isize, out, status, n_blocks};
  if (lanes < 1) lanes = 1;
  std::vector<std::vector<uint32_t>> sm(lanes, std::vector<uint32_t>(LANE_WORDS, 0xDEADBEEFu));
  std::vector<Lane<1>> L(lanes);
  uint32_t counter = 0;
  for (int i = 0; i < lanes; ++i) {
    L[i].sm = sm[i].data();
    L[i].state = ST_NEXT;
    L[i].blk = 0xffffffffu;
    L[i].err = 0;
  }
  // round-robin over lanes, exactly like the kernel's warp loop
  for (;;) {
    bool all_done = true;
    for (int i = 0; i < lanes; ++i) {
      if (L[i].state == ST_NEXT) {
        L[i].finish_block(job);
        uint32_t b = counter++;
        if (b < n_blocks) L[i].begin_block(job, b, 0);
        else { L[i].state = ST_DONE; L[i].blk = 0xffffffffu; }
      }
      if (L[i].state != ST_DONE) all_done = false;
    }
    if (all_done) break;
    for (int i = 0; i < lanes; ++i) L[i].round(budget);
  }
  return 0;
}
