// bgzf_write.cu — BGZF writer with stored DEFLATE blocks.
//
// The reference's sort writes record bodies back to back (sorting/sort.rs:643; README.md:5 "just dumps" them) and spills
// chunks as (u32 len, body)* (sort.rs:340-347).  This writer turns a BAM byte stream that is already in device memory -
// "BAM\1" + header + (u32 block_size, body)* in sorted order - into a complete .bam: BGZF blocks (gz.rs:9-20: 18-byte
// header with the BC extra field, CRC32 + ISIZE trailer) whose payload is ONE stored DEFLATE block (RFC 1951 3.2.4:
// BFINAL=1 BTYPE=00, LEN, ~LEN), 0xff00 bytes per block like htslib, then the 28-byte EOF marker.  Stored blocks keep the
// writer at copy speed (what `samtools view -u` emits); every BAM reader, this repository's included, reads the result.
// The payload bytes move with one strided device-to-device copy, the CRCs come from K2's warp CRC (crc32.cu), and the
// kernel here writes the 31 framing bytes of every block.
#include "kernels.h"

namespace bamgpu {
namespace {

__constant__ uint8_t BGZF_HDR[16] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0};
__constant__ uint8_t BGZF_EOF_MARKER[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43,
                                            0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};

// One thread per block: header in front of the payload, footer behind it.  The last block may be short; its payload is
// copied here too (the strided copy only covers the full blocks).
__global__ void k_bgzf_frames(const uint8_t* __restrict__ src, unsigned long long len, const uint32_t* __restrict__ crc,
                              uint8_t* __restrict__ out, bool eof_marker) {
  const unsigned long long n_blocks = (len + BGZF_STORE_PAYLOAD - 1) / BGZF_STORE_PAYLOAD;
  const unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < n_blocks) {
    const unsigned long long o = b * BGZF_STORE_PAYLOAD;
    const uint32_t n = (uint32_t)(len - o < BGZF_STORE_PAYLOAD ? len - o : BGZF_STORE_PAYLOAD);
    uint8_t* w = out + b * BGZF_STORE_BLOCK;
    for (int i = 0; i < 16; ++i) w[i] = BGZF_HDR[i];
    const uint32_t bsize = 18 + 5 + n + 8 - 1;            // BSIZE = total block size - 1 (util.rs:55-66 adds the 1 back)
    w[16] = (uint8_t)bsize; w[17] = (uint8_t)(bsize >> 8);
    w[18] = 0x01;                                           // BFINAL = 1, BTYPE = 00 (stored), then to the byte boundary
    w[19] = (uint8_t)n; w[20] = (uint8_t)(n >> 8);
    w[21] = (uint8_t)~n; w[22] = (uint8_t)(~n >> 8);
    if (n < BGZF_STORE_PAYLOAD) for (uint32_t i = 0; i < n; ++i) w[23 + i] = src[o + i];
    uint8_t* f = w + 23 + n;
    const uint32_t c = crc[b];
    f[0] = (uint8_t)c; f[1] = (uint8_t)(c >> 8); f[2] = (uint8_t)(c >> 16); f[3] = (uint8_t)(c >> 24);
    f[4] = (uint8_t)n; f[5] = (uint8_t)(n >> 8); f[6] = (uint8_t)(n >> 16); f[7] = (uint8_t)(n >> 24);
  }
  if (b == 0 && eof_marker) {
    const unsigned long long full = len / BGZF_STORE_PAYLOAD, rem = len % BGZF_STORE_PAYLOAD;
    uint8_t* e = out + full * BGZF_STORE_BLOCK + (rem ? 18 + 5 + rem + 8 : 0);
    for (int i = 0; i < 28; ++i) e[i] = BGZF_EOF_MARKER[i];
  }
}

}  // namespace

cudaError_t bgzf_store_blocks(const uint8_t* src, uint64_t len, const uint32_t* crc, uint8_t* out, bool eof_marker, cudaStream_t s) {
  const uint64_t full = len / BGZF_STORE_PAYLOAD;
  if (full) {
    cudaError_t e = cudaMemcpy2DAsync(out + 23, BGZF_STORE_BLOCK, src, BGZF_STORE_PAYLOAD, BGZF_STORE_PAYLOAD, full, cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return e;
  }
  const uint64_t n_blocks = (len + BGZF_STORE_PAYLOAD - 1) / BGZF_STORE_PAYLOAD;
  const uint64_t threads = n_blocks ? n_blocks : 1;
  k_bgzf_frames<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(src, len, crc, out, eof_marker);
  return cudaGetLastError();
}

}  // namespace bamgpu
