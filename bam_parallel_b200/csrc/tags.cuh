// tags.cuh — the reference's aux-tag walk (record/tags.rs:62-111 get_tag_data + get_tag) as one device function, shared by
// K3's HI look-up (records.cu, tags.rs:113-129) and K5's CG look-up (fields.cu, bamrawrecord.rs:126-157).
#pragma once
#include <stdint.h>

namespace bamgpu {

__device__ __forceinline__ int tag_size(uint8_t c) {        // tags.rs:51-58
  switch (c) {
    case 'C': case 'c': case 'A': return 1;
    case 'S': case 's': return 2;
    case 'I': case 'i': case 'f': return 4;
    default: return -1;
  }
}
__device__ __forceinline__ bool tag_type_ok(uint8_t c) {    // tags.rs:33-49: any other byte panics
  switch (c) {
    case 'A': case 'B': case 'C': case 'c': case 'f': case 'H': case 'i': case 'I': case 'S': case 's': case 'Z': return true;
    default: return false;
  }
}

// Walks data[0, n) like get_tag does: the value of EVERY tag in front of the wanted one is parsed first (and may panic).
// Returns 0 not found, 1 found (*val = offset of the value bytes in data, *vlen = their count — array bytes for B, the
// string without its NUL for Z / H, the item for scalars —, *type = the type byte), 2 = the reference would panic
// (unknown type byte, slice out of range).
__device__ inline int tag_find(const uint8_t* data, uint64_t n, uint8_t t0, uint8_t t1, uint64_t* val, uint64_t* vlen, uint8_t* type) {
  uint64_t idx = 0;
  while (idx < n) {
    if (idx + 2 > n) return 2;
    const uint8_t* d = data + idx + 2;
    const uint64_t dl = n - idx - 2;
    if (dl < 1) return 2;
    const uint8_t ty = d[0];
    if (!tag_type_ok(ty)) return 2;
    uint64_t v, used, len;
    if (ty == 'B') {
      if (dl < 2 || !tag_type_ok(d[1])) return 2;
      const int isz = tag_size(d[1]);
      if (isz < 0 || dl < 6) return 2;
      const uint64_t cnt = d[2] | (d[3] << 8) | (d[4] << 16) | ((uint64_t)d[5] << 24);
      len = cnt * (uint64_t)isz;
      if (dl < 6 + len) return 2;
      v = 6; used = 6 + len;
    } else if (ty == 'Z' || ty == 'H') {
      uint64_t i = 1;
      for (;;) { if (i >= dl) return 2; if (d[i] == 0) break; ++i; }
      v = 1; len = i - 1; used = i + 1;
    } else {
      const int isz = tag_size(ty);
      if (dl < 1 + (uint64_t)isz) return 2;
      v = 1; len = (uint64_t)isz; used = 1 + len;
    }
    if (data[idx] == t0 && data[idx + 1] == t1) {
      *val = idx + 2 + v; *vlen = len; *type = ty;
      return 1;
    }
    idx += 2 + used;
  }
  return 0;
}

}  // namespace bamgpu
