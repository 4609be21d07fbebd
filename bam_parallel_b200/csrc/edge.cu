// edge.cu — flag kernels of the shard-edge exchange: one producer GPU, one consumer GPU, NVLink peer memory.
//
// Shard g needs the first bytes of shard g+1's inflated stream so that the record straddling the edge is whole
// (reader.rs:63-81 walks straight through BGZF block boundaries).  Shard g+1 WRITES that head into an inbox in shard g's
// memory (cudaMemcpyAsync to a peer pointer: NVLink, no NCCL, no host) and then publishes the step number with
// k_set_flag; shard g's stream holds k_wait_flag until the number arrives.  Flags are 64-bit words that only ever
// grow, so a late or repeated look at them is harmless.
#include "kernels.h"

namespace bamgpu {
namespace {

__global__ void k_wait_flag(const unsigned long long* flag, unsigned long long want, uint32_t* timed_out) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= want) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > 20000000000ull) { *timed_out = 1; break; }   // 20 s: the neighbour is gone; never hang the GPU
    __nanosleep(500);
  }
  __threadfence_system();   // acquire side: the halo bytes written before the flag are visible to what follows
}

__global__ void k_set_flag(unsigned long long* flag, unsigned long long value) {
  __threadfence_system();   // release side: the peer copy queued before this kernel has completed (stream order)
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(value) : "memory");
}

}  // namespace

void launch_wait_flag(const unsigned long long* flag, unsigned long long want, uint32_t* timed_out, cudaStream_t s) {
  k_wait_flag<<<1, 1, 0, s>>>(flag, want, timed_out);
}
void launch_set_flag(unsigned long long* flag, unsigned long long value, cudaStream_t s) { k_set_flag<<<1, 1, 0, s>>>(flag, value); }

}  // namespace bamgpu
