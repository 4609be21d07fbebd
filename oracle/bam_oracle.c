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
#include <pthread.h>
#include <stdint.h>
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

static sort_ctx* g_ctx; /* qsort has no context argument; oracle sorts are single-threaded */

/* comparators.rs:61-65: str::cmp == bytewise lexicographic over the name INCLUDING its
 * trailing NUL (get_range gives 32..32+l_read_name, bamrawrecord.rs:118-123). */
static int cmp_names(uint32_t a, uint32_t b) {
  const uint8_t* na = g_ctx->u + g_ctx->rec_off[a] + 32;
  const uint8_t* nb = g_ctx->u + g_ctx->rec_off[b] + 32;
  size_t la = g_ctx->c->l_read_name[a], lb = g_ctx->c->l_read_name[b];
  size_t m = la < lb ? la : lb;
  int r = memcmp(na, nb, m);
  if (r) return r;
  return (la > lb) - (la < lb);
}

static int cmp_perm(const void* pa, const void* pb) {
  uint32_t a = *(const uint32_t*)pa, b = *(const uint32_t*)pb;
  const orc_columns* c = g_ctx->c;
  int r = 0;
  if (g_ctx->mode == 0) { /* comparators.rs:117-147 */
    uint64_t ka = c->coord_key[a], kb = c->coord_key[b];
    if (ka != kb) r = ka < kb ? -1 : 1;
    else r = (int)g_ctx->strand[a] - (int)g_ctx->strand[b];
  } else {
    r = cmp_names(a, b);
    if (r == 0 && g_ctx->mode == 2) { /* comparators.rs:78-93 */
      uint8_t pa_ = c->hi_present[a], pb_ = c->hi_present[b];
      int same = (pa_ == pb_) && (!pa_ || c->hi_value[a] == c->hi_value[b]);
      if (same) {
        r = (int)c->flag[a] - (int)c->flag[b];
      } else {
        if (!pa_ || !pb_) { g_ctx->err = ORC_ERR_HI_MIXED; r = 0; }
        else r = c->hi_value[a] < c->hi_value[b] ? -1 : 1;
      }
    }
  }
  if (r) return r;
  return (a > b) - (a < b); /* stability */
}

int orc_sort_perm(const uint8_t* u, const uint64_t* rec_off, const orc_columns* c, const uint8_t* strand, size_t n,
                  int mode, uint32_t* perm) {
  sort_ctx ctx = {u, rec_off, c, strand, mode, 0};
  for (size_t i = 0; i < n; ++i) perm[i] = (uint32_t)i;
  g_ctx = &ctx;
  qsort(perm, n, sizeof(uint32_t), cmp_perm);
  g_ctx = NULL;
  return ctx.err;
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
typedef struct {
  const uint8_t* file;
  size_t n;
  int n_slots;
  int n_workers;
  /* ring */
  uint8_t** ubuf;      /* n_slots x (1 MiB) */
  uint32_t* ulen;
  const uint8_t** cptr;
  uint32_t* clen;
  int* state;          /* 0 free, 1 queued, 2 inflating, 3 done, 4 eof, 5 error */
  uint64_t head_fill;  /* next block index the reader will fill */
  uint64_t head_work;  /* next block index a worker will take */
  pthread_mutex_t mu;
  pthread_cond_t cv;
  int reader_done;
  int stop;            /* consumer finished early (bounded sample): everybody exits */
  int error;
} cpu_reader;

static void* cpu_reader_thread(void* arg) {
  cpu_reader* r = (cpu_reader*)arg;
  size_t pos = 0;
  for (;;) {
    int slot = (int)(r->head_fill % (uint64_t)r->n_slots);
    pthread_mutex_lock(&r->mu);
    while (r->state[slot] != 0 && !r->stop) pthread_cond_wait(&r->cv, &r->mu);
    int stop = r->stop;
    if (stop) { r->reader_done = 1; pthread_cond_broadcast(&r->cv); }
    pthread_mutex_unlock(&r->mu);
    if (stop) return NULL;
    uint16_t bs = bgzf_block_size(r->file + pos, r->n - pos);
    int st = 1;
    if (bs == 0) st = 4;
    else if (bs < BGZF_HEADER_SIZE + TRAILER_SIZE || r->n - pos < bs) st = 5;
    else {
      r->cptr[slot] = r->file + pos + BGZF_HEADER_SIZE;
      r->clen[slot] = (uint32_t)bs - BGZF_HEADER_SIZE - TRAILER_SIZE;
      pos += bs;
    }
    pthread_mutex_lock(&r->mu);
    r->state[slot] = st;
    r->head_fill++;
    if (st != 1) r->reader_done = 1;
    pthread_cond_broadcast(&r->cv);
    pthread_mutex_unlock(&r->mu);
    if (st != 1) return NULL;
  }
}

static void* cpu_worker_thread(void* arg) {
  cpu_reader* r = (cpu_reader*)arg;
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  inflateInit2(&zs, -15);
  for (;;) {
    pthread_mutex_lock(&r->mu);
    int slot = -1;
    for (;;) {
      if (r->stop) { pthread_mutex_unlock(&r->mu); inflateEnd(&zs); return NULL; }
      if (r->head_work < r->head_fill) {
        int s = (int)(r->head_work % (uint64_t)r->n_slots);
        if (r->state[s] == 1) { slot = s; r->state[s] = 2; r->head_work++; break; }
        if (r->state[s] >= 4) { pthread_mutex_unlock(&r->mu); inflateEnd(&zs); return NULL; }
      } else if (r->reader_done) { pthread_mutex_unlock(&r->mu); inflateEnd(&zs); return NULL; }
      pthread_cond_wait(&r->cv, &r->mu);
    }
    pthread_mutex_unlock(&r->mu);
    long u = inflate_raw(&zs, r->cptr[slot], r->clen[slot], r->ubuf[slot], 1u << 20);
    pthread_mutex_lock(&r->mu);
    if (u < 0) { r->state[slot] = 5; } else { r->ulen[slot] = (uint32_t)u; r->state[slot] = 3; }
    pthread_cond_broadcast(&r->cv);
    pthread_mutex_unlock(&r->mu);
  }
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
  pthread_mutex_lock(&r->mu);
  if (c->slot >= 0) { r->state[c->slot] = 0; pthread_cond_broadcast(&r->cv); }
  int slot = (int)(c->next_block % (uint64_t)r->n_slots);
  while (!(r->state[slot] >= 3 && c->next_block < r->head_fill)) pthread_cond_wait(&r->cv, &r->mu);
  int st = r->state[slot];
  pthread_mutex_unlock(&r->mu);
  if (st == 4) { c->eof = 1; c->slot = -1; return 0; }
  if (st == 5) { c->eof = 1; c->slot = -1; r->error = 1; return -1; }
  c->slot = slot;
  c->cur = 0;
  c->next_block++;
  c->ubytes += r->ulen[slot];
  return 1;
}

/* impl Read for Reader (reader.rs:114-152) driven by read_exact. Returns bytes copied (< want only at EOF). */
static size_t consumer_read_exact(cpu_consumer* c, uint8_t* dst, size_t want) {
  size_t got = 0;
  while (got < want) {
    if (c->eof) break;
    if (c->slot < 0 || c->cur == c->r->ulen[c->slot]) {
      if (consumer_next_block(c) <= 0) break;
      continue; /* ErrorKind::Interrupted -> read_exact retries */
    }
    size_t avail = c->r->ulen[c->slot] - c->cur;
    size_t k = want - got < avail ? want - got : avail;
    memcpy(dst + got, c->r->ubuf[c->slot] + c->cur, k);
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
  r.ubuf = (uint8_t**)calloc((size_t)threads, sizeof(uint8_t*));
  r.ulen = (uint32_t*)calloc((size_t)threads, sizeof(uint32_t));
  r.cptr = (const uint8_t**)calloc((size_t)threads, sizeof(uint8_t*));
  r.clen = (uint32_t*)calloc((size_t)threads, sizeof(uint32_t));
  r.state = (int*)calloc((size_t)threads, sizeof(int));
  for (int i = 0; i < threads; ++i) r.ubuf[i] = (uint8_t*)malloc(1u << 20);
  pthread_mutex_init(&r.mu, NULL);
  pthread_cond_init(&r.cv, NULL);
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
  pthread_mutex_lock(&r.mu);
  r.stop = 1;
  pthread_cond_broadcast(&r.cv);
  pthread_mutex_unlock(&r.mu);
  pthread_join(rt, NULL);
  for (int i = 0; i < r.n_workers; ++i) pthread_join(wt[i], NULL);
  for (int i = 0; i < threads; ++i) free(r.ubuf[i]);
  free(r.ubuf); free(r.ulen); free((void*)r.cptr); free(r.clen); free(r.state); free(wt); free(rec);
  if (records) *records = nrec;
  if (ubytes) *ubytes = c.ubytes;
  if (seconds) *seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  if (checksum) *checksum = sum;
  return rc;
}

/* zlib's CRC-32 (what a BGZF trailer holds), exposed so tests can check the GPU CRC kernel. */
uint32_t orc_crc32(const uint8_t* p, size_t n) { return (uint32_t)crc32(crc32(0, NULL, 0), p, (uInt)n); }
