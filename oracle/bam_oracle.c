/*
 * bam_oracle.c — CPU restatement of the reference's BGZF→BAM-record hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  Nothing under
 * bam_parallel_b200/ links or calls it; the product path fails loudly without its
 * CUDA library instead of falling back to this.
 *
 * PARITY UNPINNED: the reference (NickRoz1/bam_parallel) is Rust and this image has
 * no Rust toolchain, its only golden values are md5s over two BAM files that need
 * network access (tests_list.py:6-48), and its DEFLATE decoder is the un-vendored
 * crate flate2 ^1.0.1 (bam_tools/Cargo.toml:11; Cargo.lock is git-ignored).  The
 * one reference-held known-answer vector on this path — BAMRawRecord::default()
 * (bam_tools/src/record/bamrawrecord.rs:180-197) — is checked in
 * tests/test_oracle.py.  DEFLATE decoding is deterministic (RFC 1951), so system
 * zlib raw inflate (windowBits -15) yields the same bytes as flate2 on every valid
 * stream.
 *
 * Each function cites the reference file:line (relative to bam_tools/src/) it restates.
 * All integers are little-endian, as byteorder::LittleEndian in the reference.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <zlib.h>

/* gz.rs:9-20 */
#define HEADER_SIZE 10
#define TRAILER_SIZE 8
#define BGZF_HEADER_SIZE 18

enum {
  ORC_OK = 0,
  ORC_ERR_BLOCK_TOO_SMALL = -1, /* util.rs:29-38 InvalidData "expected clen >= 26" */
  ORC_ERR_TRUNCATED_BLOCK = -2, /* util.rs:42,44 read_exact failure -> unwrap panic readahead.rs:92 */
  ORC_ERR_INFLATE = -3,         /* readahead.rs:148 expect("Decompression failed.") */
  ORC_ERR_NOMEM = -4,
  ORC_ERR_BAD_MAGIC = -5,       /* reader.rs:90-95 "invalid BAM header" */
  ORC_ERR_SHORT_HEADER = -6,    /* reader.rs:172-190 unwrap()/? on short header */
  ORC_ERR_SHORT_RECORD = -7,    /* reader.rs:69-70 expect("Failed to read.") */
  ORC_ERR_BAD_FLAGS = -8,       /* comparators.rs:178 Flags::from_bits(..).unwrap() */
  ORC_ERR_HI_MIXED = -9,        /* comparators.rs:90 None.unwrap() */
  ORC_ERR_BAD_TAG = -10,        /* tags.rs:49 / :127 panics */
};

static inline uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline uint32_t rd32(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* util.rs:55-66 read_block_size: 18-byte header, short read => 0, BSIZE+1 wraps as u16
 * (release build), NO magic / CM / FLG / XLEN / SI checks. */
static uint16_t bgzf_block_size(const uint8_t* p, size_t remaining) {
  if (remaining < BGZF_HEADER_SIZE) return 0;
  return (uint16_t)(rd16(p + 16) + 1);
}

/* One BGZF block located in the file. */
typedef struct {
  uint64_t coff;  /* offset of the 18-byte header in the file */
  uint32_t bsize; /* total block size (header + cdata + trailer) */
} orc_block;

/* util.rs:18-53 fetch_block applied over the whole file: returns number of blocks or <0.
 * Stops at the first header that cannot be read in full or whose (BSIZE+1) wraps to 0. */
long orc_scan_blocks(const uint8_t* file, size_t n, orc_block* out, size_t cap) {
  size_t pos = 0;
  long count = 0;
  for (;;) {
    uint16_t bs = bgzf_block_size(file + pos, n - pos);
    if (bs == 0) break;
    if (bs < BGZF_HEADER_SIZE + TRAILER_SIZE) return ORC_ERR_BLOCK_TOO_SMALL;
    if (n - pos < (size_t)bs) return ORC_ERR_TRUNCATED_BLOCK; /* cdata or trailer read_exact fails */
    if (out && (size_t)count < cap) {
      out[count].coff = pos;
      out[count].bsize = bs;
    }
    ++count;
    pos += bs;
  }
  return count;
}

/* util.rs:10-16 inflate_data: raw DEFLATE, read_to_end. Output is unbounded in the
 * reference (Vec grows); here `cap` must be large enough (callers pass >= 1 MiB or
 * grow).  Trailing bytes after the final DEFLATE block are ignored, as flate2 does.
 * CRC32 and ISIZE in the trailer are never looked at (util.rs:44,68-71). */
static long inflate_raw(z_stream* zs, const uint8_t* cdata, size_t clen, uint8_t* dst, size_t cap) {
  inflateReset(zs);
  zs->next_in = (Bytef*)cdata;
  zs->avail_in = (uInt)clen;
  zs->next_out = dst;
  zs->avail_out = (uInt)cap;
  int rc = inflate(zs, Z_FINISH);
  if (rc != Z_STREAM_END) return ORC_ERR_INFLATE; /* includes Z_BUF_ERROR: truncated stream or cap hit */
  return (long)(cap - zs->avail_out);
}

/* Whole-file decompression: concatenation of every block's inflated bytes, which is the
 * byte stream `impl Read for Reader` serves (reader.rs:114-152).  *ustream is malloc'ed.
 * block_ulen (optional, cap entries) receives each block's inflated size. */
long orc_inflate_all(const uint8_t* file, size_t n, uint8_t** ustream, uint64_t* ulen, uint32_t* block_ulen,
                     size_t cap_blocks) {
  long nb = orc_scan_blocks(file, n, NULL, 0);
  if (nb < 0) return nb;
  orc_block* blocks = (orc_block*)malloc(sizeof(orc_block) * (size_t)(nb ? nb : 1));
  if (!blocks) return ORC_ERR_NOMEM;
  orc_scan_blocks(file, n, blocks, (size_t)nb);
  size_t cap = (size_t)nb * 65536 + (1u << 20);
  uint8_t* out = (uint8_t*)malloc(cap);
  if (!out) { free(blocks); return ORC_ERR_NOMEM; }
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit2(&zs, -15) != Z_OK) { free(blocks); free(out); return ORC_ERR_NOMEM; }
  size_t o = 0;
  long rc = nb;
  for (long i = 0; i < nb; ++i) {
    const uint8_t* c = file + blocks[i].coff + BGZF_HEADER_SIZE;
    size_t clen = blocks[i].bsize - BGZF_HEADER_SIZE - TRAILER_SIZE;
    if (cap - o < (1u << 20)) { /* keep 1 MiB of headroom: a block may inflate past 64 KiB */
      cap = cap * 2;
      uint8_t* t = (uint8_t*)realloc(out, cap);
      if (!t) { rc = ORC_ERR_NOMEM; break; }
      out = t;
    }
    long u = inflate_raw(&zs, c, clen, out + o, cap - o);
    if (u < 0) { rc = u; break; }
    if (block_ulen && (size_t)i < cap_blocks) block_ulen[i] = (uint32_t)u;
    o += (size_t)u;
  }
  inflateEnd(&zs);
  free(blocks);
  if (rc < 0) { free(out); return rc; }
  *ustream = out;
  *ulen = o;
  return nb;
}

void orc_free(void* p) { free(p); }

/* reader.rs:87-98 + :154-200 read_header: checks "BAM\1", returns the header bytes
 * WITHOUT the magic (l_text | text | n_ref | (l_name | name | l_ref)*), the offset of
 * n_ref inside that byte string (4 + l_text), and how many stream bytes were consumed
 * (4 + header_len) so the record chain can start there. */
int orc_read_header(const uint8_t* u, uint64_t ulen, uint64_t* hdr_off, uint64_t* hdr_len, uint64_t* off_to_refs,
                    uint64_t* consumed) {
  if (ulen < 4) return ORC_ERR_SHORT_HEADER;
  if (memcmp(u, "BAM\1", 4) != 0) return ORC_ERR_BAD_MAGIC;
  uint64_t p = 4;
  if (ulen - p < 4) return ORC_ERR_SHORT_HEADER;
  uint64_t l_text = rd32(u + p);
  p += 4;
  if (ulen - p < l_text) return ORC_ERR_SHORT_HEADER;
  p += l_text;
  *off_to_refs = p - 4;
  if (ulen - p < 4) return ORC_ERR_SHORT_HEADER;
  uint32_t n_ref = rd32(u + p);
  p += 4;
  for (uint32_t i = 0; i < n_ref; ++i) {
    if (ulen - p < 4) return ORC_ERR_SHORT_HEADER;
    uint64_t l_name = rd32(u + p);
    p += 4;
    if (ulen - p < l_name) return ORC_ERR_SHORT_HEADER;
    p += l_name;
    if (ulen - p < 4) return ORC_ERR_SHORT_HEADER;
    p += 4;
  }
  *hdr_off = 4;
  *hdr_len = p - 4;
  *consumed = p;
  return ORC_OK;
}

/* reader.rs:57-81 + records.rs:23-29: the blind block_size chain.  From `start`, read
 * u32 block_size (fewer than 4 bytes left => UnexpectedEof => 0 => end); block_size==0
 * ends iteration; a body that runs past the stream end is the reference's
 * expect("Failed to read.") panic => ORC_ERR_SHORT_RECORD.  Writes body offsets/lengths
 * (offset of the first body byte, i.e. after block_size) when out arrays are non-NULL.
 * Returns record count or <0. */
long orc_walk_records(const uint8_t* u, uint64_t ulen, uint64_t start, uint64_t* rec_off, uint32_t* rec_len,
                      size_t cap) {
  uint64_t p = start;
  long n = 0;
  for (;;) {
    if (ulen - p < 4) break;
    uint32_t bs = rd32(u + p);
    if (bs == 0) break;
    if (ulen - (p + 4) < bs) return ORC_ERR_SHORT_RECORD;
    if (rec_off && (size_t)n < cap) {
      rec_off[n] = p + 4;
      rec_len[n] = bs;
    }
    ++n;
    p += 4 + (uint64_t)bs;
  }
  return n;
}

/* Fixed-field SoA columns: record/bamrawrecord.rs:86-97 (body-relative offsets). */
typedef struct {
  int32_t* ref_id;      /* @0  */
  int32_t* pos;         /* @4  */
  uint8_t* l_read_name; /* @8  */
  uint8_t* mapq;        /* @9  */
  uint16_t* bin;        /* @10 */
  uint16_t* n_cigar_op; /* @12 */
  uint16_t* flag;       /* @14 */
  uint32_t* l_seq;      /* @16 */
  int32_t* next_ref_id; /* @20 */
  int32_t* next_pos;    /* @24 */
  int32_t* tlen;        /* @28 */
  /* variable-length field offsets, body-relative: bamrawrecord.rs:61-77 (usize form) */
  uint32_t* cigar_off;  /* 32 + l_read_name */
  uint32_t* seq_off;    /* + 4*n_cigar_op */
  uint32_t* qual_off;   /* + (l_seq+1)/2 */
  uint32_t* tags_off;   /* + l_seq */
  /* sort keys: sorting/comparators.rs:41-59 */
  uint64_t* coord_key;  /* (refID' ^ 2^31) << 32 | (pos ^ 2^31); strand kept separately (see coord_key()) */
  uint8_t* hi_present;  /* tags.rs:113-129 get_hit_count -> Option<i32> */
  int32_t* hi_value;
} orc_columns;

/* tags.rs:33-49 get_tag_type: any other byte panics. */
static int tag_type_ok(uint8_t c) { return strchr("ABCcfHiISsZ", c) != NULL && c != 0; }
/* tags.rs:51-58 tag_size */
static int tag_size(uint8_t c) {
  switch (c) {
    case 'C': case 'c': case 'A': return 1;
    case 'S': case 's': return 2;
    case 'I': case 'i': case 'f': return 4;
    default: return -1;
  }
}

/* tags.rs:62-111 get_tag_data + get_tag, restated over an explicit [data, data+len) slice.
 * Returns 1 found / 0 not found / <0 for the reference's panics (bad type byte, slice
 * index out of range).  On found: *val points at the value bytes, *type is the type byte. */
static int get_tag(const uint8_t* data, size_t len, const char tag[2], const uint8_t** val, uint8_t* type) {
  size_t idx = 0;
  while (idx < len) {
    /* get_tag_data(&data[idx + 2..]) is evaluated BEFORE the name comparison (tags.rs:97-99) */
    if (idx + 2 > len) return ORC_ERR_BAD_TAG;      /* &data[idx+2..] slice panic */
    const uint8_t* d = data + idx + 2;
    size_t dl = len - idx - 2;
    if (dl < 1) return ORC_ERR_BAD_TAG;             /* data[0] index panic */
    uint8_t t = d[0];
    if (!tag_type_ok(t)) return ORC_ERR_BAD_TAG;
    const uint8_t* v;
    size_t used;
    if (t == 'B') {
      if (dl < 2) return ORC_ERR_BAD_TAG;
      if (!tag_type_ok(d[1])) return ORC_ERR_BAD_TAG;
      int isz = tag_size(d[1]);
      if (isz < 0) return ORC_ERR_BAD_TAG;          /* tag_size(..).unwrap() on None */
      if (dl < 6) return ORC_ERR_BAD_TAG;
      size_t cnt = rd32(d + 2);
      size_t bytes = cnt * (size_t)isz;
      if (dl < 6 + bytes) return ORC_ERR_BAD_TAG;
      v = d + 6;
      used = 6 + bytes;
    } else if (t == 'Z' || t == 'H') {
      size_t i = 1;
      for (;;) {
        if (i >= dl) return ORC_ERR_BAD_TAG;        /* data[idx] index panic */
        if (d[i] == 0) break;
        ++i;
      }
      v = d + 1;
      used = i + 1;
    } else {
      int isz = tag_size(t);
      if (dl < 1 + (size_t)isz) return ORC_ERR_BAD_TAG;
      v = d + 1;
      used = 1 + (size_t)isz;
    }
    if (data[idx] == (uint8_t)tag[0] && data[idx + 1] == (uint8_t)tag[1]) {
      *val = v;
      *type = t;
      return 1;
    }
    idx += 2 + used;
  }
  return 0;
}

/* tags.rs:113-129 get_hit_count over the RawTags slice as get_bytes(RawTags) computes it
 * (bamrawrecord.rs:80-104): NOTE the closure does `32 + self.l_read_name()` in u8, so the
 * tags offset is computed from ((32 + l_read_name) mod 256) in a release build; for
 * l_read_name >= 224 the slice starts 256 bytes early.  Mirrored here. */
static int hit_count(const uint8_t* body, uint32_t len, uint8_t* present, int32_t* value) {
  uint8_t l_name = body[8];
  uint32_t cig = (uint8_t)(32u + l_name); /* u8 wrap */
  uint32_t n_cig = rd16(body + 12);
  uint32_t l_seq = rd32(body + 16);
  uint64_t tags = (uint64_t)cig + 4ull * n_cig + (uint64_t)((l_seq + 1u) / 2u) + l_seq; /* (l_seq+1)/2 in u32, wraps like the reference */
  *present = 0;
  *value = 0;
  if (tags > len) return ORC_ERR_BAD_TAG; /* get_slice range panic */
  const uint8_t* v;
  uint8_t t;
  int rc = get_tag(body + tags, (size_t)(len - tags), "HI", &v, &t);
  if (rc <= 0) return rc;
  *present = 1;
  switch (t) {
    case 'A': case 'C': *value = (int32_t)v[0]; break;
    case 'c': *value = (int32_t)(int8_t)v[0]; break;
    case 's': *value = (int32_t)(int16_t)rd16(v); break;
    case 'S': *value = (int32_t)rd16(v); break;
    case 'i': case 'I': *value = (int32_t)rd32(v); break;
    default: return ORC_ERR_BAD_TAG; /* tags.rs:127 panic */
  }
  return ORC_OK;
}

/* Order-isomorphic packing of the (refID', pos) part of KeyTuple::CoordinatesAndStrand(i32, i32, bool)
 * (comparators.rs:51-57 + :117-147): refID' = (refID == -1 ? i32::MAX : refID); both i32 are biased
 * by 2^31 so unsigned comparison of the u64 equals signed lexicographic comparison of (refID', pos).
 * The third component (false < true on flag & 0x10) does not fit and is kept in the `strand` column:
 * full order = (coord_key, strand). */
static uint64_t coord_key(int32_t ref_id, int32_t pos) {
  if (ref_id == -1) ref_id = INT32_MAX;
  uint64_t r = (uint32_t)ref_id ^ 0x80000000u;
  uint64_t p = (uint32_t)pos ^ 0x80000000u;
  return (r << 32) | p;
}

/* Extract every column for `n` records.  `want_hi` enables the HI tag walk (only the
 * name+mates sort uses it).  `strand` gets flag&0x10 (comparators.rs:173-180) and
 * *bad_flags counts records whose flag has bits >= 0x1000 (reference would panic in the
 * coordinate sort, flags.rs:2-31 has 12 flags). */
int orc_extract(const uint8_t* u, const uint64_t* rec_off, const uint32_t* rec_len, size_t n, orc_columns* c,
                uint8_t* strand, int want_hi, uint64_t* bad_flags, uint64_t* bad_tags) {
  uint64_t bf = 0, bt = 0;
  for (size_t i = 0; i < n; ++i) {
    const uint8_t* b = u + rec_off[i];
    /* Records shorter than 32 bytes would panic in get_slice; callers only pass bodies >= 32. */
    c->ref_id[i] = (int32_t)rd32(b + 0);
    c->pos[i] = (int32_t)rd32(b + 4);
    c->l_read_name[i] = b[8];
    c->mapq[i] = b[9];
    c->bin[i] = rd16(b + 10);
    c->n_cigar_op[i] = rd16(b + 12);
    c->flag[i] = rd16(b + 14);
    c->l_seq[i] = rd32(b + 16);
    c->next_ref_id[i] = (int32_t)rd32(b + 20);
    c->next_pos[i] = (int32_t)rd32(b + 24);
    c->tlen[i] = (int32_t)rd32(b + 28);
    uint32_t cig = 32u + b[8];
    uint32_t seq = cig + 4u * c->n_cigar_op[i];
    uint32_t qual = seq + (c->l_seq[i] + 1u) / 2u;
    uint32_t tags = qual + c->l_seq[i];
    c->cigar_off[i] = cig;
    c->seq_off[i] = seq;
    c->qual_off[i] = qual;
    c->tags_off[i] = tags;
    c->coord_key[i] = coord_key(c->ref_id[i], c->pos[i]);
    if (strand) strand[i] = (c->flag[i] & 0x10) ? 1 : 0;
    if (c->flag[i] & 0xF000) ++bf;
    if (want_hi) {
      int rc = hit_count(b, rec_len[i], &c->hi_present[i], &c->hi_value[i]);
      if (rc < 0) ++bt;
    } else if (c->hi_present) {
      c->hi_present[i] = 0;
      c->hi_value[i] = 0;
    }
  }
  if (bad_flags) *bad_flags = bf;
  if (bad_tags) *bad_tags = bt;
  return ORC_OK;
}

/* ---- sort orders (sorting/comparators.rs, sort.rs:351-371,487-510) ------------------------------
 * The reference sorts each chunk with a stable par_sort_by and merges chunks with a
 * lower-chunk-index tie-break, i.e. the global order is the STABLE sort of the input
 * by the comparator.  We sort a permutation with (comparator, input index). */
typedef struct {
  const uint8_t* u;
  const uint64_t* rec_off;
  const orc_columns* c;
  const uint8_t* strand;
  int mode; /* 0 coord+strand, 1 name, 2 name+mates */
  int err;
} sort_ctx;

/* comparators.rs:61-65: str::cmp == bytewise lexicographic over the name INCLUDING its
 * trailing NUL (get_range gives 32..32+l_read_name, bamrawrecord.rs:118-123). */
static int cmp_names(const sort_ctx* x, uint32_t a, uint32_t b) {
  const uint8_t* na = x->u + x->rec_off[a] + 32;
  const uint8_t* nb = x->u + x->rec_off[b] + 32;
  size_t la = x->c->l_read_name[a], lb = x->c->l_read_name[b];
  size_t m = la < lb ? la : lb;
  int r = memcmp(na, nb, m);
  if (r) return r;
  return (la > lb) - (la < lb);
}

static int cmp_perm(const void* pa, const void* pb, void* vx) {
  sort_ctx* x = (sort_ctx*)vx;
  uint32_t a = *(const uint32_t*)pa, b = *(const uint32_t*)pb;
  const orc_columns* c = x->c;
  int r = 0;
  if (x->mode == 0) { /* comparators.rs:117-147 */
    uint64_t ka = c->coord_key[a], kb = c->coord_key[b];
    if (ka != kb) r = ka < kb ? -1 : 1;
    else r = (int)x->strand[a] - (int)x->strand[b];
  } else {
    r = cmp_names(x, a, b);
    if (r == 0 && x->mode == 2) { /* comparators.rs:78-93 */
      uint8_t pa_ = c->hi_present[a], pb_ = c->hi_present[b];
      int same = (pa_ == pb_) && (!pa_ || c->hi_value[a] == c->hi_value[b]);
      if (same) {
        r = (int)c->flag[a] - (int)c->flag[b];
      } else {
        if (!pa_ || !pb_) { x->err = ORC_ERR_HI_MIXED; r = 0; }
        else r = c->hi_value[a] < c->hi_value[b] ? -1 : 1;
      }
    }
  }
  if (r) return r;
  return (a > b) - (a < b); /* stability: (comparator, input index) is a total order */
}

/* Since (comparator, input index) is a total order, ANY correct sorting algorithm yields the one stable order the
 * reference's par_sort_by + chunk-index tie-break produce (sort.rs:351-371,487-510).  Large inputs are sorted as
 * `threads` runs (qsort_r each, in parallel) that are then merged pairwise, level by level, also in parallel. */
typedef struct { sort_ctx ctx; uint32_t* a; uint32_t* b; size_t lo, mid, hi; } sort_job;

static void* sort_run_thread(void* arg) {
  sort_job* j = (sort_job*)arg;
  qsort_r(j->a + j->lo, j->hi - j->lo, sizeof(uint32_t), cmp_perm, &j->ctx);
  return NULL;
}
static void* sort_merge_thread(void* arg) {
  sort_job* j = (sort_job*)arg;
  const uint32_t* a = j->a;
  uint32_t* o = j->b;
  size_t i = j->lo, k = j->mid, w = j->lo;
  while (i < j->mid && k < j->hi) {
    if (cmp_perm(&a[i], &a[k], &j->ctx) <= 0) o[w++] = a[i++];
    else o[w++] = a[k++];
  }
  while (i < j->mid) o[w++] = a[i++];
  while (k < j->hi) o[w++] = a[k++];
  return NULL;
}

int orc_sort_perm_mt(const uint8_t* u, const uint64_t* rec_off, const orc_columns* c, const uint8_t* strand, size_t n,
                     int mode, uint32_t* perm, int threads) {
  sort_ctx ctx = {u, rec_off, c, strand, mode, 0};
  for (size_t i = 0; i < n; ++i) perm[i] = (uint32_t)i;
  if (threads < 1) threads = 1;
  if (n < 65536) threads = 1;
  int runs = 1;
  while (runs < threads) runs <<= 1;
  if (runs == 1) {
    qsort_r(perm, n, sizeof(uint32_t), cmp_perm, &ctx);
    return ctx.err;
  }
  uint32_t* tmp = (uint32_t*)malloc(sizeof(uint32_t) * n);
  if (!tmp) return ORC_ERR_NOMEM;
  sort_job* jobs = (sort_job*)calloc((size_t)runs, sizeof(sort_job));
  pthread_t* th = (pthread_t*)calloc((size_t)runs, sizeof(pthread_t));
  for (int r = 0; r < runs; ++r) {
    jobs[r].ctx = ctx; jobs[r].a = perm; jobs[r].lo = n * (size_t)r / (size_t)runs; jobs[r].hi = n * (size_t)(r + 1) / (size_t)runs;
    pthread_create(&th[r], NULL, sort_run_thread, &jobs[r]);
  }
  int err = 0;
  for (int r = 0; r < runs; ++r) { pthread_join(th[r], NULL); if (jobs[r].ctx.err) err = jobs[r].ctx.err; }
  uint32_t *src = perm, *dst = tmp;
  for (int width = 1; width < runs; width <<= 1) {
    int nj = 0;
    for (int r = 0; r < runs; r += 2 * width) {
      sort_job* j = &jobs[nj];
      j->ctx = ctx; j->a = src; j->b = dst;
      j->lo = n * (size_t)r / (size_t)runs; j->mid = n * (size_t)(r + width) / (size_t)runs; j->hi = n * (size_t)(r + 2 * width) / (size_t)runs;
      pthread_create(&th[nj], NULL, sort_merge_thread, j);
      ++nj;
    }
    for (int k = 0; k < nj; ++k) { pthread_join(th[k], NULL); if (jobs[k].ctx.err) err = jobs[k].ctx.err; }
    uint32_t* t = src; src = dst; dst = t;
  }
  if (src != perm) memcpy(perm, src, sizeof(uint32_t) * n);
  free(tmp); free(jobs); free(th);
  return err;
}

int orc_sort_perm(const uint8_t* u, const uint64_t* rec_off, const orc_columns* c, const uint8_t* strand, size_t n,
                  int mode, uint32_t* perm) {
  return orc_sort_perm_mt(u, rec_off, c, strand, n, mode, perm, 1);
}

/* ---- CPU baseline: the reference reader's thread topology, timed -----------------------------
 * readahead.rs:51-123: one reading thread walks block headers and hands each block to an
 * inflate task; n-2 workers inflate (util.rs:10-16); completed blocks are re-sequenced by
 * index (the reference uses an ordering thread with a min-heap, :68-86; a ring indexed by
 * block number has the same effect); only n Blocks circulate (:58-61), so the reader stalls
 * until the consumer returns one (:90).  The consumer is single-threaded and, per record,
 * reads a u32 then copies the body out of the current block(s) into a reused buffer
 * (reader.rs:57-73, :114-152), exactly two reads + one resize per record.
 * Inflate is zlib 1.3 here vs miniz_oxide in the reference: same bytes, different speed. */
/* Signalling: the reference hands Blocks around over flume channels (readahead.rs:54-57), i.e. every hand-over wakes
 * exactly one peer.  Here every ring slot has its own mutex + condition variable (reader <-> consumer hand-over of
 * that slot, worker -> consumer completion) and the inflate workers share one queue whose producer wakes ONE worker
 * (pthread_cond_signal).  (Round 1 used one mutex and pthread_cond_broadcast for everything: every state change woke
 * every thread, and the port got SLOWER with more cores.) */
typedef struct {
  pthread_mutex_t mu;
  pthread_cond_t cv;
  int state;           /* 0 free, 1 queued, 2 inflating, 3 done, 4 eof, 5 error */
  uint8_t* ubuf;       /* 1 MiB */
  uint32_t ulen;
  const uint8_t* cptr;
  uint32_t clen;
} cpu_slot;

typedef struct {
  const uint8_t* file;
  size_t n;
  int n_slots;
  int n_workers;
  cpu_slot* slots;
  /* work queue: block indices [head_work, head_fill) wait for a worker */
  pthread_mutex_t q_mu;
  pthread_cond_t q_cv;
  uint64_t head_fill;  /* next block index the reader will fill */
  uint64_t head_work;  /* next block index a worker will take */
  int reader_done;
  volatile int stop;   /* consumer finished early (bounded sample): everybody exits */
  int error;
} cpu_reader;

static void* cpu_reader_thread(void* arg) {
  cpu_reader* r = (cpu_reader*)arg;
  size_t pos = 0;
  uint64_t k = 0;
  for (;;) {
    cpu_slot* s = &r->slots[k % (uint64_t)r->n_slots];
    /* readahead.rs:90: wait until the consumer has returned this Block */
    pthread_mutex_lock(&s->mu);
    while (s->state != 0 && !r->stop) pthread_cond_wait(&s->cv, &s->mu);
    pthread_mutex_unlock(&s->mu);
    int st = 1;
    if (!r->stop) {
      uint16_t bs = bgzf_block_size(r->file + pos, r->n - pos);
      if (bs == 0) st = 4;
      else if (bs < BGZF_HEADER_SIZE + TRAILER_SIZE || r->n - pos < bs) st = 5;
      else {
        s->cptr = r->file + pos + BGZF_HEADER_SIZE;
        s->clen = (uint32_t)bs - BGZF_HEADER_SIZE - TRAILER_SIZE;
        pos += bs;
      }
      pthread_mutex_lock(&s->mu);
      s->state = st;
      if (st != 1) pthread_cond_broadcast(&s->cv);   /* eof / error sentinel: the consumer may be waiting for it */
      pthread_mutex_unlock(&s->mu);
    }
    pthread_mutex_lock(&r->q_mu);
    if (r->stop || st != 1) {
      r->reader_done = 1;
      pthread_cond_broadcast(&r->q_cv);           /* once, at the end: every worker must see it */
      pthread_mutex_unlock(&r->q_mu);
      return NULL;
    }
    r->head_fill = ++k;
    pthread_cond_signal(&r->q_cv);                /* one task, one worker */
    pthread_mutex_unlock(&r->q_mu);
  }
}

static void* cpu_worker_thread(void* arg) {
  cpu_reader* r = (cpu_reader*)arg;
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  inflateInit2(&zs, -15);
  for (;;) {
    pthread_mutex_lock(&r->q_mu);
    while (r->head_work >= r->head_fill && !r->reader_done && !r->stop) pthread_cond_wait(&r->q_cv, &r->q_mu);
    if (r->stop || r->head_work >= r->head_fill) { pthread_mutex_unlock(&r->q_mu); break; }
    uint64_t k = r->head_work++;
    pthread_mutex_unlock(&r->q_mu);
    cpu_slot* s = &r->slots[k % (uint64_t)r->n_slots];
    long u = inflate_raw(&zs, s->cptr, s->clen, s->ubuf, 1u << 20);
    pthread_mutex_lock(&s->mu);
    if (u < 0) s->state = 5; else { s->ulen = (uint32_t)u; s->state = 3; }
    pthread_cond_broadcast(&s->cv);               /* <= 2 waiters on a slot: the consumer (for this block) and, when the ring is full, the reader (for the slot to come back) */
    pthread_mutex_unlock(&s->mu);
  }
  inflateEnd(&zs);
  return NULL;
}

typedef struct {
  cpu_reader* r;
  uint64_t next_block;
  int slot;           /* current slot or -1 */
  uint32_t cur;       /* cursor in current block */
  int eof;
  uint64_t ubytes;
} cpu_consumer;

/* readahead.rs:128-142 get_block: give the spent Block back, wait for the next in order. */
static int consumer_next_block(cpu_consumer* c) {
  cpu_reader* r = c->r;
  if (c->slot >= 0) {
    cpu_slot* o = &r->slots[c->slot];
    pthread_mutex_lock(&o->mu);
    o->state = 0;
    pthread_cond_broadcast(&o->cv);               /* the reader may be waiting for exactly this slot */
    pthread_mutex_unlock(&o->mu);
  }
  int slot = (int)(c->next_block % (uint64_t)r->n_slots);
  cpu_slot* s = &r->slots[slot];
  pthread_mutex_lock(&s->mu);
  while (s->state < 3) pthread_cond_wait(&s->cv, &s->mu);
  int st = s->state;
  pthread_mutex_unlock(&s->mu);
  if (st == 4) { c->eof = 1; c->slot = -1; return 0; }
  if (st == 5) { c->eof = 1; c->slot = -1; r->error = 1; return -1; }
  c->slot = slot;
  c->cur = 0;
  c->next_block++;
  c->ubytes += s->ulen;
  return 1;
}

/* impl Read for Reader (reader.rs:114-152) driven by read_exact. Returns bytes copied (< want only at EOF). */
static size_t consumer_read_exact(cpu_consumer* c, uint8_t* dst, size_t want) {
  size_t got = 0;
  while (got < want) {
    if (c->eof) break;
    if (c->slot < 0 || c->cur == c->r->slots[c->slot].ulen) {
      if (consumer_next_block(c) <= 0) break;
      continue; /* ErrorKind::Interrupted -> read_exact retries */
    }
    size_t avail = c->r->slots[c->slot].ulen - c->cur;
    size_t k = want - got < avail ? want - got : avail;
    memcpy(dst + got, c->r->slots[c->slot].ubuf + c->cur, k);
    c->cur += (uint32_t)k;
    got += k;
  }
  return got;
}

/* Runs the whole reader over an in-memory BAM with `threads` total threads (clamped to >=3
 * as readahead.rs:53).  Returns 0 and fills records / ubytes / seconds / a checksum (sum of
 * all body bytes, so the copy cannot be optimised away), or <0 on the reference's panics. */
int orc_cpu_reader_run(const uint8_t* file, size_t n, int threads, uint64_t max_records, uint64_t* records,
                       uint64_t* ubytes, double* seconds, uint64_t* checksum) {
  if (threads < 3) threads = 3;
  cpu_reader r;
  memset(&r, 0, sizeof r);
  r.file = file;
  r.n = n;
  r.n_slots = threads;
  r.n_workers = threads - 2;
  r.slots = (cpu_slot*)calloc((size_t)threads, sizeof(cpu_slot));
  for (int i = 0; i < threads; ++i) {
    r.slots[i].ubuf = (uint8_t*)malloc(1u << 20);
    pthread_mutex_init(&r.slots[i].mu, NULL);
    pthread_cond_init(&r.slots[i].cv, NULL);
  }
  pthread_mutex_init(&r.q_mu, NULL);
  pthread_cond_init(&r.q_cv, NULL);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  pthread_t rt;
  pthread_t* wt = (pthread_t*)calloc((size_t)r.n_workers, sizeof(pthread_t));
  pthread_create(&rt, NULL, cpu_reader_thread, &r);
  for (int i = 0; i < r.n_workers; ++i) pthread_create(&wt[i], NULL, cpu_worker_thread, &r);

  cpu_consumer c;
  memset(&c, 0, sizeof c);
  c.r = &r;
  c.slot = -1;
  int rc = ORC_OK;
  uint64_t nrec = 0, sum = 0;
  size_t cap = 1 << 16;
  uint8_t* rec = (uint8_t*)malloc(cap);
  /* read_header (reader.rs:87-98,154-200) */
  uint8_t w[4];
  do {
    if (consumer_read_exact(&c, w, 4) < 4 || memcmp(w, "BAM\1", 4)) { rc = ORC_ERR_BAD_MAGIC; break; }
    if (consumer_read_exact(&c, w, 4) < 4) { rc = ORC_ERR_SHORT_HEADER; break; }
    uint32_t l_text = rd32(w);
    if (l_text > cap) { cap = l_text; rec = (uint8_t*)realloc(rec, cap); }
    if (consumer_read_exact(&c, rec, l_text) < l_text) { rc = ORC_ERR_SHORT_HEADER; break; }
    if (consumer_read_exact(&c, w, 4) < 4) { rc = ORC_ERR_SHORT_HEADER; break; }
    uint32_t n_ref = rd32(w);
    for (uint32_t i = 0; i < n_ref && rc == ORC_OK; ++i) {
      if (consumer_read_exact(&c, w, 4) < 4) { rc = ORC_ERR_SHORT_HEADER; break; }
      uint32_t l = rd32(w);
      if (l > cap) { cap = l; rec = (uint8_t*)realloc(rec, cap); }
      if (consumer_read_exact(&c, rec, l) < l) { rc = ORC_ERR_SHORT_HEADER; break; }
      if (consumer_read_exact(&c, w, 4) < 4) { rc = ORC_ERR_SHORT_HEADER; break; }
    }
  } while (0);
  /* records().next_rec() loop (records.rs:23-29 -> reader.rs:57-73) */
  while (rc == ORC_OK && nrec < max_records) {
    if (consumer_read_exact(&c, w, 4) < 4) break; /* UnexpectedEof => 0 */
    uint32_t bs = rd32(w);
    if (bs == 0) break;
    if (bs > cap) { cap = bs; rec = (uint8_t*)realloc(rec, cap); }
    if (consumer_read_exact(&c, rec, bs) < bs) { rc = ORC_ERR_SHORT_RECORD; break; }
    sum += rec[0] + rec[bs - 1] + bs;
    ++nrec;
  }
  if (r.error) rc = ORC_ERR_INFLATE;
  clock_gettime(CLOCK_MONOTONIC, &t1);
  /* shut down: tell the reader and the workers to exit wherever they are */
  pthread_mutex_lock(&r.q_mu);
  r.stop = 1;
  pthread_cond_broadcast(&r.q_cv);
  pthread_mutex_unlock(&r.q_mu);
  for (int i = 0; i < threads; ++i) {
    pthread_mutex_lock(&r.slots[i].mu);
    pthread_cond_broadcast(&r.slots[i].cv);
    pthread_mutex_unlock(&r.slots[i].mu);
  }
  pthread_join(rt, NULL);
  for (int i = 0; i < r.n_workers; ++i) pthread_join(wt[i], NULL);
  for (int i = 0; i < threads; ++i) { free(r.slots[i].ubuf); pthread_mutex_destroy(&r.slots[i].mu); pthread_cond_destroy(&r.slots[i].cv); }
  free(r.slots); free(wt); free(rec);
  if (records) *records = nrec;
  if (ubytes) *ubytes = c.ubytes;
  if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  if (checksum) *checksum = sum;
  return rc;
}

/* zlib's CRC-32 (what a BGZF trailer holds), exposed so tests can check the GPU CRC kernel. */
uint32_t orc_crc32(const uint8_t* p, size_t n) { return (uint32_t)crc32(crc32(0, NULL, 0), p, (uInt)n); }

/* =====================================================================================================
 * Variable-length field decode: record/bamrawrecord.rs:80-104 get_bytes, :126-157 get_cigar,
 * :208-236 get_cig_op / decode_cigar, :238-270 get_seq_base / decode_seq.  Test infrastructure like the rest of this file.
 * A "panic" of the reference (slice out of range, unwrap on a short chunk, unsupported op) is ORC_ERR_PANIC for that record.
 * ===================================================================================================== */
#define ORC_ERR_PANIC (-40)
typedef struct { int32_t ref, pos; uint32_t n_cig, l_seq, nseq; uint64_t cig, seq, qual, tags; } var_geom;

/* the closures of get_bytes (:80-84): `32 + self.l_read_name()` is u8 + u8 (wraps in a release build), the rest usize;
 * ((l_seq + 1) / 2) is u32 arithmetic */
static void var_geometry(const uint8_t* body, var_geom* g) {
  g->ref = (int32_t)rd32(body); g->pos = (int32_t)rd32(body + 4);
  g->n_cig = rd16(body + 12); g->l_seq = rd32(body + 16);
  g->nseq = (uint32_t)(g->l_seq + 1u) / 2u;
  g->cig = (uint8_t)(32u + body[8]);
  g->seq = g->cig + 4ull * g->n_cig;
  g->qual = g->seq + g->nseq;
  g->tags = g->qual + g->l_seq;
}

/* get_cigar (:126-157) */
static int get_cigar(const uint8_t* body, uint32_t len, const var_geom* g, const uint8_t** src, size_t* nb) {
  *src = body; *nb = 0;
  if (g->ref < 0 || g->pos < 0 || g->n_cig == 0) return ORC_OK;               /* return &[] */
  size_t fb = 4 * (size_t)g->n_cig;
  if (g->cig + fb > len) return ORC_ERR_PANIC;                                 /* get_slice */
  *src = body + g->cig; *nb = fb;
  uint32_t first = rd32(body + g->cig);                                        /* &cigar_field_data[..4] */
  if ((first & 0xf) != 4 || (size_t)(first >> 4) != (size_t)g->nseq || g->n_cig != 2) return ORC_OK;
  if (g->tags > len) return ORC_ERR_PANIC;                                     /* self.0.len() - get_tags_offset() */
  const uint8_t* v; uint8_t t;
  size_t tl = (size_t)(len - g->tags);
  int rc = get_tag(body + g->tags, tl, "CG", &v, &t);
  if (rc < 0) return ORC_ERR_PANIC;
  if (rc == 1) {   /* tag_val.0: the value slice get_tag_data returns (tags.rs:62-93) */
    size_t vl;
    if (t == 'B') vl = (size_t)rd32(v - 4) * (size_t)tag_size(*(v - 5));
    else if (t == 'Z' || t == 'H') { vl = 0; while (v[vl] != 0) ++vl; }
    else vl = (size_t)tag_size(t);
    *src = v; *nb = vl;
  }
  return ORC_OK;
}

/* decode_cigar (:223-236); dst NULL = length only */
static long decode_cigar(const uint8_t* src, size_t nb, uint8_t* dst) {
  static const char OPS[] = "MIDNSHP=X";
  long o = 0;
  for (size_t i = 0; i < nb; i += 4) {
    if (nb - i < 4) return ORC_ERR_PANIC;              /* slice.read_u32().unwrap() on a short chunk */
    uint32_t cig = rd32(src + i), op = cig & 0xf, n = cig >> 4;
    char tmp[16];
    int d = snprintf(tmp, sizeof tmp, "%u", n);        /* op_len.to_string() */
    if (op > 8) return ORC_ERR_PANIC;                  /* get_cig_op */
    if (dst) { memcpy(dst + o, tmp, (size_t)d); dst[o + d] = (uint8_t)OPS[op]; }
    o += d + 1;
  }
  return o;
}

/* decode_seq (:259-270) */
static long decode_seq(const uint8_t* src, size_t nb, uint8_t* dst) {
  static const char BASES[] = "=ACMGRSVTWYHKDBN";
  long o = 0;
  for (size_t i = 0; i < nb; ++i) {
    uint32_t first = src[i] >> 4, second = src[i] & 0xf;
    if (dst) dst[o] = (uint8_t)BASES[first];
    ++o;
    if (second != 0) { if (dst) dst[o] = (uint8_t)BASES[second]; ++o; }
  }
  return o;
}

/* One record's field: 1 CIGAR string, 2 sequence string, 4 raw quality bytes.  Returns its length or ORC_ERR_PANIC. */
static long decode_one_field(const uint8_t* body, uint32_t len, int field, uint8_t* dst) {
  if (len < 32) return ORC_ERR_PANIC;                   /* get_slice of a fixed field */
  var_geom g;
  var_geometry(body, &g);
  if (field == 1) {
    const uint8_t* src; size_t nb;
    int rc = get_cigar(body, len, &g, &src, &nb);
    if (rc < 0) return rc;
    return decode_cigar(src, nb, dst);
  }
  if (field == 2) {
    if (g.seq + g.nseq > len) return ORC_ERR_PANIC;
    return decode_seq(body + g.seq, g.nseq, dst);
  }
  if (g.qual + g.l_seq > len) return ORC_ERR_PANIC;
  if (dst) memcpy(dst, body + g.qual, g.l_seq);
  return (long)g.l_seq;
}

/* All records.  Call with out == NULL first: off[0..n] gets the offsets (a panicking record contributes nothing and is
 * counted in *bad); then again with `out` of off[n] bytes. */
int orc_decode_field(const uint8_t* u, const uint64_t* rec_off, const uint32_t* rec_len, size_t n, int field, uint64_t* off,
                     uint8_t* out, uint64_t* bad) {
  uint64_t o = 0, nb = 0;
  for (size_t i = 0; i < n; ++i) {
    if (!out) off[i] = o;
    long l = decode_one_field(u + rec_off[i], rec_len[i], field, NULL);
    if (l < 0) { ++nb; continue; }
    if (out) decode_one_field(u + rec_off[i], rec_len[i], field, out + off[i]);
    o += (uint64_t)l;
  }
  if (!out) off[n] = o;
  if (bad) *bad = nb;
  return ORC_OK;
}

/* =====================================================================================================
 * Full-size parity helpers (tests at BASELINE.json's 50M-record sizes).  Everything below is the SAME
 * restatement as above, run on several threads, plus a digest: 64-bit hashes of every value the reader /
 * record accessors / sort produce, folded into a few words, so a GPU run over 50M records can be compared
 * with the oracle without moving 20 GB of columns around.  The digest's definition (DIGEST SPEC):
 *   mix64(z)       = splitmix64 finaliser of z + 0x9E3779B97F4A7C15
 *   h_fields(i)    = fold of [index i, stream offset of the body, block_size, refID, pos, l_read_name, mapq, bin,
 *                    n_cigar_op, flag, l_seq, next_refID, next_pos, tlen, cigar/seq/qual/tags offsets, coord_key,
 *                    strand, HI present, HI value] with h = mix64(h ^ value), starting from h = 0
 *   h_body(bytes)  = sum over 32-bit little-endian words w_k of the body (zero padded) of mix64(w_k | k << 32)
 *   fields = sum_i h_fields(i);  bodies = sum_i mix64(h_body(body_i) + i * 0x9E3779B97F4A7C15)      (mod 2^64)
 *   sorted: perm = sum_i mix64(perm[i] | i << 32);  bodies as above over the OUTPUT order
 * The product computes the same function on the GPU from ITS columns and buffers (csrc/digest.cu).
 * ===================================================================================================== */
static inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

static uint64_t body_hash(const uint8_t* b, uint32_t len) {
  uint64_t h = 0;
  uint32_t k = 0, i = 0;
  for (; i + 4 <= len; i += 4, ++k) h += mix64((uint64_t)rd32(b + i) | ((uint64_t)k << 32));
  if (i < len) {
    uint32_t w = 0;
    for (uint32_t j = 0; i + j < len; ++j) w |= (uint32_t)b[i + j] << (8 * j);
    h += mix64((uint64_t)w | ((uint64_t)k << 32));
  }
  return h;
}

typedef struct { uint64_t n_records, body_bytes, fields, bodies; } orc_digest_t;

typedef struct {
  /* parallel-for description */
  int kind;           /* 0 inflate, 1 extract, 2 digest, 3 sorted digest, 4 field digest */
  int field;
  size_t lo, hi;
  /* inflate */
  const uint8_t* file; const orc_block* blocks; const uint64_t* uoff; uint8_t* out; int err;
  /* records */
  const uint8_t* u; const uint64_t* rec_off; const uint32_t* rec_len; const orc_columns* c; uint8_t* strand; int want_hi;
  const uint32_t* perm; uint64_t index_base, stream_base;
  uint64_t bad_flags, bad_tags;
  orc_digest_t d;
} par_job;

static void shift_columns(const orc_columns* c, size_t o, orc_columns* s) {
  s->ref_id = c->ref_id + o; s->pos = c->pos + o; s->l_read_name = c->l_read_name + o; s->mapq = c->mapq + o;
  s->bin = c->bin + o; s->n_cigar_op = c->n_cigar_op + o; s->flag = c->flag + o; s->l_seq = c->l_seq + o;
  s->next_ref_id = c->next_ref_id + o; s->next_pos = c->next_pos + o; s->tlen = c->tlen + o;
  s->cigar_off = c->cigar_off + o; s->seq_off = c->seq_off + o; s->qual_off = c->qual_off + o; s->tags_off = c->tags_off + o;
  s->coord_key = c->coord_key + o; s->hi_present = c->hi_present ? c->hi_present + o : NULL; s->hi_value = c->hi_value ? c->hi_value + o : NULL;
}

static void* par_thread(void* arg) {
  par_job* j = (par_job*)arg;
  if (j->kind == 0) {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) { j->err = ORC_ERR_NOMEM; return NULL; }
    for (size_t b = j->lo; b < j->hi; ++b) {
      const uint8_t* cdat = j->file + j->blocks[b].coff + BGZF_HEADER_SIZE;
      size_t clen = j->blocks[b].bsize - BGZF_HEADER_SIZE - TRAILER_SIZE;
      size_t cap = (size_t)(j->uoff[b + 1] - j->uoff[b]);
      /* one spare byte would let zlib report "more output than ISIZE"; Z_FINISH with an exact-size buffer returns
       * Z_STREAM_END only if the stream ends exactly there, which is what we verify */
      long got = inflate_raw(&zs, cdat, clen, j->out + j->uoff[b], cap);
      if (got < 0 || (size_t)got != cap) { j->err = ORC_ERR_INFLATE; break; }
    }
    inflateEnd(&zs);
  } else if (j->kind == 1) {
    orc_columns sc;
    shift_columns(j->c, j->lo, &sc);
    orc_extract(j->u, j->rec_off + j->lo, j->rec_len + j->lo, j->hi - j->lo, &sc, j->strand ? j->strand + j->lo : NULL, j->want_hi,
                &j->bad_flags, &j->bad_tags);
  } else if (j->kind == 2) {
    const orc_columns* c = j->c;
    orc_digest_t d = {0, 0, 0, 0};
    for (size_t i = j->lo; i < j->hi; ++i) {
      uint64_t idx = j->index_base + i, h = 0;
#define FOLD(v) h = mix64(h ^ (uint64_t)(v))
      FOLD(idx); FOLD(j->stream_base + j->rec_off[i]); FOLD(j->rec_len[i]);
      FOLD((uint32_t)c->ref_id[i]); FOLD((uint32_t)c->pos[i]); FOLD(c->l_read_name[i]); FOLD(c->mapq[i]); FOLD(c->bin[i]);
      FOLD(c->n_cigar_op[i]); FOLD(c->flag[i]); FOLD(c->l_seq[i]); FOLD((uint32_t)c->next_ref_id[i]); FOLD((uint32_t)c->next_pos[i]);
      FOLD((uint32_t)c->tlen[i]); FOLD(c->cigar_off[i]); FOLD(c->seq_off[i]); FOLD(c->qual_off[i]); FOLD(c->tags_off[i]);
      FOLD(c->coord_key[i]); FOLD(j->strand[i]); FOLD(c->hi_present[i] == 1 ? 1u : 0u); FOLD((uint32_t)(c->hi_present[i] == 1 ? c->hi_value[i] : 0));
#undef FOLD
      d.fields += h;
      d.bodies += mix64(body_hash(j->u + j->rec_off[i], j->rec_len[i]) + idx * 0x9E3779B97F4A7C15ull);
      d.body_bytes += j->rec_len[i];
    }
    d.n_records = j->hi - j->lo;
    j->d = d;
  } else if (j->kind == 4) {
    orc_digest_t d = {0, 0, 0, 0};
    size_t cap = 1 << 16;
    uint8_t* buf = (uint8_t*)malloc(cap);
    for (size_t i = j->lo; i < j->hi; ++i) {
      const uint8_t* body = j->u + j->rec_off[i];
      long l = decode_one_field(body, j->rec_len[i], j->field, NULL);
      if (l < 0) { l = 0; d.fields += 1; }               /* `fields` counts the panicking records here */
      if ((size_t)l + 8 > cap) { cap = (size_t)l * 2 + 8; buf = (uint8_t*)realloc(buf, cap); }
      if (l) decode_one_field(body, j->rec_len[i], j->field, buf);
      d.bodies += mix64(body_hash(buf, (uint32_t)l) + (j->index_base + i) * 0x9E3779B97F4A7C15ull);
      d.body_bytes += (uint64_t)l;
    }
    free(buf);
    d.n_records = j->hi - j->lo;
    j->d = d;
  } else {
    orc_digest_t d = {0, 0, 0, 0};
    for (size_t i = j->lo; i < j->hi; ++i) {
      uint32_t r = j->perm[i];
      d.fields += mix64((uint64_t)r | ((uint64_t)i << 32));
      d.bodies += mix64(body_hash(j->u + j->rec_off[r], j->rec_len[r]) + (uint64_t)i * 0x9E3779B97F4A7C15ull);
      d.body_bytes += j->rec_len[r];
    }
    d.n_records = j->hi - j->lo;
    j->d = d;
  }
  return NULL;
}

static int par_run(par_job* proto, size_t n, int threads, par_job** jobs_out) {
  if (threads < 1) threads = 1;
  if ((size_t)threads > n) threads = n ? (int)n : 1;
  par_job* jobs = (par_job*)calloc((size_t)threads, sizeof(par_job));
  pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
  for (int t = 0; t < threads; ++t) {
    jobs[t] = *proto;
    jobs[t].lo = n * (size_t)t / (size_t)threads;
    jobs[t].hi = n * (size_t)(t + 1) / (size_t)threads;
    pthread_create(&th[t], NULL, par_thread, &jobs[t]);
  }
  for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
  free(th);
  *jobs_out = jobs;
  return threads;
}

/* util.rs:18-71 + :10-16 over the whole file on `threads` threads.  Output placement uses the ISIZE footers (which the
 * reference ignores) ONLY to know where each block's bytes go; every block is then required to inflate to exactly
 * that many bytes, otherwise ORC_ERR_INFLATE (use the serial orc_inflate_all for files with lying footers). */
long orc_inflate_all_mt(const uint8_t* file, size_t n, uint8_t** ustream, uint64_t* ulen, int threads) {
  long nb = orc_scan_blocks(file, n, NULL, 0);
  if (nb < 0) return nb;
  orc_block* blocks = (orc_block*)malloc(sizeof(orc_block) * (size_t)(nb ? nb : 1));
  uint64_t* uoff = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(nb + 1));
  if (!blocks || !uoff) return ORC_ERR_NOMEM;
  orc_scan_blocks(file, n, blocks, (size_t)nb);
  uint64_t tot = 0;
  for (long i = 0; i < nb; ++i) { uoff[i] = tot; tot += rd32(file + blocks[i].coff + blocks[i].bsize - 4); }
  uoff[nb] = tot;
  uint8_t* out = (uint8_t*)malloc(tot + 64);
  if (!out) { free(blocks); free(uoff); return ORC_ERR_NOMEM; }
  memset(out + tot, 0, 64);
  par_job proto;
  memset(&proto, 0, sizeof proto);
  proto.kind = 0; proto.file = file; proto.blocks = blocks; proto.uoff = uoff; proto.out = out;
  par_job* jobs;
  int nt = par_run(&proto, (size_t)nb, threads, &jobs);
  long rc = nb;
  for (int t = 0; t < nt; ++t) if (jobs[t].err) rc = jobs[t].err;
  free(jobs); free(blocks); free(uoff);
  if (rc < 0) { free(out); return rc; }
  *ustream = out;
  *ulen = tot;
  return rc;
}

int orc_extract_mt(const uint8_t* u, const uint64_t* rec_off, const uint32_t* rec_len, size_t n, orc_columns* c,
                   uint8_t* strand, int want_hi, uint64_t* bad_flags, uint64_t* bad_tags, int threads) {
  par_job proto;
  memset(&proto, 0, sizeof proto);
  proto.kind = 1; proto.u = u; proto.rec_off = rec_off; proto.rec_len = rec_len; proto.c = c; proto.strand = strand; proto.want_hi = want_hi;
  par_job* jobs;
  int nt = par_run(&proto, n, threads, &jobs);
  uint64_t bf = 0, bt = 0;
  for (int t = 0; t < nt; ++t) { bf += jobs[t].bad_flags; bt += jobs[t].bad_tags; }
  free(jobs);
  if (bad_flags) *bad_flags = bf;
  if (bad_tags) *bad_tags = bt;
  return ORC_OK;
}

/* DIGEST SPEC over the oracle's own walk + columns.  stream_base is added to rec_off (records of a shard), index_base
 * to the record index. */
int orc_digest(const uint8_t* u, const uint64_t* rec_off, const uint32_t* rec_len, const orc_columns* c, const uint8_t* strand,
               size_t n, uint64_t index_base, uint64_t stream_base, int threads, orc_digest_t* out) {
  par_job proto;
  memset(&proto, 0, sizeof proto);
  proto.kind = 2; proto.u = u; proto.rec_off = rec_off; proto.rec_len = rec_len; proto.c = c; proto.strand = (uint8_t*)strand;
  proto.index_base = index_base; proto.stream_base = stream_base;
  par_job* jobs;
  int nt = par_run(&proto, n, threads, &jobs);
  orc_digest_t d = {0, 0, 0, 0};
  for (int t = 0; t < nt; ++t) { d.n_records += jobs[t].d.n_records; d.body_bytes += jobs[t].d.body_bytes; d.fields += jobs[t].d.fields; d.bodies += jobs[t].d.bodies; }
  if (n == 0) d.n_records = 0;
  free(jobs);
  *out = d;
  return ORC_OK;
}

/* DIGEST SPEC of a sort result: `fields` holds the permutation digest, `bodies` the digest of the output bytes
 * (sort.rs:643: bodies written back to back in sorted order). */
int orc_sorted_digest(const uint8_t* u, const uint64_t* rec_off, const uint32_t* rec_len, const uint32_t* perm, size_t n,
                      int threads, orc_digest_t* out) {
  par_job proto;
  memset(&proto, 0, sizeof proto);
  proto.kind = 3; proto.u = u; proto.rec_off = rec_off; proto.rec_len = rec_len; proto.perm = perm;
  par_job* jobs;
  int nt = par_run(&proto, n, threads, &jobs);
  orc_digest_t d = {0, 0, 0, 0};
  for (int t = 0; t < nt; ++t) { d.n_records += jobs[t].d.n_records; d.body_bytes += jobs[t].d.body_bytes; d.fields += jobs[t].d.fields; d.bodies += jobs[t].d.bodies; }
  if (n == 0) d.n_records = 0;
  free(jobs);
  *out = d;
  return ORC_OK;
}

/* bodies-digest (DIGEST SPEC) of one decoded field without materialising it: body_bytes = total length,
 * bodies = sum_i mix64(h_body(field_i) + i * K), fields = number of records the reference would panic on. */
int orc_field_digest(const uint8_t* u, const uint64_t* rec_off, const uint32_t* rec_len, size_t n, int field, int threads,
                     orc_digest_t* out) {
  par_job proto;
  memset(&proto, 0, sizeof proto);
  proto.kind = 4; proto.field = field; proto.u = u; proto.rec_off = rec_off; proto.rec_len = rec_len;
  par_job* jobs;
  int nt = par_run(&proto, n, threads, &jobs);
  orc_digest_t d = {0, 0, 0, 0};
  for (int t = 0; t < nt; ++t) { d.n_records += jobs[t].d.n_records; d.body_bytes += jobs[t].d.body_bytes; d.fields += jobs[t].d.fields; d.bodies += jobs[t].d.bodies; }
  if (n == 0) d.n_records = 0;
  free(jobs);
  *out = d;
  return ORC_OK;
}
