/* next_rec_harness.c — what a Rust `records().next_rec()` loop costs on top of the batch pipeline.
 *
 * The reference's consumer (bam_binary hash mode, main.rs:83-99) calls Records::next_rec() once per record
 * (records.rs:23-29); the Rust shim (rust/bamgpu-sys) forwards each call to bamgpu_next_rec.  This C loop does the
 * same 50M calls without an interpreter in the way, touches every body (first, last byte and length) and reports
 * the count.  Built into bam_parallel_b200/lib/libbamharness.so; only bench.py's `e2e_next_rec` uses it. */
#include <stdint.h>
#include <time.h>

#include "bamgpu.h"

int harness_next_rec_all(bamgpu_reader* r, uint64_t* n_records, uint64_t* body_bytes, uint64_t* checksum, double* seconds) {
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  uint64_t n = 0, bytes = 0, sum = 0;
  int rc;
  for (;;) {
    const uint8_t* body;
    size_t len;
    rc = bamgpu_next_rec(r, &body, &len);
    if (rc != BAMGPU_OK) break;
    sum += (uint64_t)body[0] + body[len - 1] + len;
    bytes += len;
    ++n;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  *n_records = n;
  *body_bytes = bytes;
  *checksum = sum;
  *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  return rc == BAMGPU_EOF ? BAMGPU_OK : rc;
}
