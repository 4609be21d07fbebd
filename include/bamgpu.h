/*
 * bamgpu.h — C ABI of the B200-native BAM decode path (libbamgpu.so).
 *
 * This is the drop-in boundary for ONE path of NickRoz1/bam_parallel: the
 * multi-threaded BGZF inflater + BAM record parser behind
 *     bam_tools::Reader::new(reader, n, track_progress)      bam_tools/src/reader.rs:28-55
 *     Reader::read_header()                                  reader.rs:87-98
 *     Reader::records().next_rec()                           reader/records.rs:23-29
 *     Reader::read_record / append_record                    reader.rs:57-73
 *     impl Read for Reader                                   reader.rs:114-152
 *     parse_reference_sequences()                            reader.rs:101-112
 * plus the key-extraction slice of the sort that consumes it
 *     create_key_tuple / KeyTuple                            sorting/comparators.rs:15-59
 *     BAMRawRecord fixed/variable field offsets              record/bamrawrecord.rs:49-105
 *     get_hit_count (HI tag)                                 record/tags.rs:113-129
 *
 * The reference has no FFI of its own (pure Rust); these entry points are what a Rust
 * `bam_tools` shim binds (rust/bamgpu-sys).  Plain C types only: no
 * CUDA, torch or C++ types cross this boundary.  No exceptions or panics cross it
 * either: every call returns 0 / a negative bamgpu_status, and bamgpu_last_error()
 * gives the text.  One consumer thread per reader (the reference takes &mut self).
 *
 * There is NO CPU fallback: if no CUDA device is usable, constructors fail with
 * BAMGPU_ERR_CUDA.
 */
#ifndef BAMGPU_H
#define BAMGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum bamgpu_status {
  BAMGPU_OK = 0,
  BAMGPU_EOF = 1,                   /* clean end of stream (reference: Ok(0) / None) */
  BAMGPU_ERR_ARG = -1,
  BAMGPU_ERR_CUDA = -2,             /* CUDA runtime failure or no device */
  BAMGPU_ERR_NOMEM = -3,
  BAMGPU_ERR_IO = -4,               /* the caller's read callback returned < 0 */
  BAMGPU_ERR_BLOCK_TOO_SMALL = -10, /* util.rs:29-38  InvalidData "expected clen >= 26" */
  BAMGPU_ERR_TRUNCATED_BLOCK = -11, /* util.rs:42,44  read_exact fails -> readahead.rs:92 unwrap panic */
  BAMGPU_ERR_INFLATE = -12,         /* readahead.rs:148 expect("Decompression failed.") */
  BAMGPU_ERR_ISIZE = -13,           /* stricter than the reference: inflated size != ISIZE footer */
  BAMGPU_ERR_CRC = -14,             /* stricter than the reference: CRC32 footer mismatch (util.rs:68-71 discards it) */
  BAMGPU_ERR_BAD_MAGIC = -20,       /* reader.rs:90-95 "invalid BAM header" */
  BAMGPU_ERR_SHORT_HEADER = -21,    /* reader.rs:172-190 unwrap()/? on a short header */
  BAMGPU_ERR_SHORT_RECORD = -22,    /* reader.rs:69-70 expect("Failed to read.") */
  BAMGPU_ERR_BUFFER_TOO_SMALL = -23,
  BAMGPU_ERR_STATE = -24,           /* call out of order (e.g. records before read_header) */
  BAMGPU_ERR_SORT_HI_MIXED = -30,   /* sorting/comparators.rs:90 unwrap(): equal names, HI on one record only -> the reference panics */
  BAMGPU_ERR_SORT_BAD_FLAGS = -31,  /* comparators.rs:178 Flags::from_bits().unwrap(): flag bits >= 0x1000 -> the reference panics */
  BAMGPU_ERR_SORT_BAD_TAGS = -32,   /* record/tags.rs:64-129: the aux walk for HI hits a type it panics on */
  BAMGPU_ERR_TOO_LARGE = -33        /* sort: more than 2^31 records, or the sort keys alone do not fit in device memory */
} bamgpu_status;

/* ------------------------------------------------------------------------------------------
 * Host-side BGZF scan (replaces util.rs:18-71 fetch_block/read_block_size/consume_trailer).
 * Unlike the reference we KEEP the trailer: ISIZE places each block's output by exclusive
 * scan, CRC32 is verified on the GPU.
 * ---------------------------------------------------------------------------------------- */
typedef struct bamgpu_block {
  uint64_t in_off;   /* offset of the DEFLATE payload (after the 18-byte header) in the scanned buffer */
  uint64_t out_off;  /* exclusive scan of ISIZE: where this block's bytes land in the inflated stream */
  uint32_t in_len;   /* DEFLATE payload length = BSIZE+1-26 */
  uint32_t isize;    /* ISIZE footer */
  uint32_t crc32;    /* CRC32 footer */
  uint32_t reserved;
} bamgpu_block;

/* Scans `len` bytes of a BGZF stream.  Writes up to `cap` block descriptors; *n_blocks gets the
 * number found; *consumed the bytes covered by whole blocks; *total_isize the sum of ISIZE.
 * `final` != 0 means no more bytes will follow: a trailing partial header (<18 bytes) or a
 * BSIZE of 0xFFFF ends the stream silently (util.rs:58-60,65), a partial block body is
 * BAMGPU_ERR_TRUNCATED_BLOCK.  With final == 0 a partial tail is simply left unconsumed.
 * Returns BAMGPU_OK, BAMGPU_EOF (silent end seen) or an error. */
int bamgpu_scan_blocks(const uint8_t* buf, size_t len, int final, bamgpu_block* blocks, size_t cap,
                       size_t* n_blocks, size_t* consumed, uint64_t* total_isize);

/* ------------------------------------------------------------------------------------------
 * Decoded batch: structure-of-arrays view of consecutive records (K3 output).
 * Every pointer is a DEVICE pointer unless the name starts with h_ (pinned host mirror, only
 * filled when the corresponding BAMGPU_COPY_* bit is set).  Valid until the next
 * bamgpu_next_batch / bamgpu_engine_decode call on the same object.
 * Field meanings: record/bamrawrecord.rs:86-97 (fixed), :61-77 (variable offsets, usize form),
 * sorting/comparators.rs:41-59 (keys), record/tags.rs:113-129 (HI).
 * ---------------------------------------------------------------------------------------- */
typedef struct bamgpu_columns {
  const uint64_t* rec_off;      /* offset of the record BODY (after block_size) in `data` */
  const uint32_t* rec_len;      /* block_size */
  const int32_t* ref_id;
  const int32_t* pos;
  const uint8_t* l_read_name;
  const uint8_t* mapq;
  const uint16_t* bin;
  const uint16_t* n_cigar_op;
  const uint16_t* flag;
  const uint32_t* l_seq;
  const int32_t* next_ref_id;
  const int32_t* next_pos;
  const int32_t* tlen;
  const uint32_t* cigar_off;    /* body-relative: 32 + l_read_name */
  const uint32_t* seq_off;      /* + 4*n_cigar_op */
  const uint32_t* qual_off;     /* + (l_seq+1)/2 */
  const uint32_t* tags_off;     /* + l_seq */
  const uint64_t* coord_key;    /* ((refID==-1 ? i32::MAX : refID) ^ 2^31) << 32 | (pos ^ 2^31); order = (coord_key, strand) */
  const uint8_t* strand;        /* (flag & 0x10) != 0 : comparators.rs:173-180 */
  const uint8_t* hi_present;    /* HI tag found (get_hit_count -> Some) ; 2 = reference would panic walking the tags */
  const int32_t* hi_value;
} bamgpu_columns;

typedef struct bamgpu_batch {
  uint64_t n_records;
  uint64_t data_len;            /* bytes of inflated stream held in `data` */
  const uint8_t* data;          /* device: inflated bytes; record i's body = data[rec_off[i] .. +rec_len[i]] */
  const uint8_t* h_data;        /* pinned host mirror of data (BAMGPU_COPY_DATA) */
  bamgpu_columns dev;           /* device columns */
  bamgpu_columns host;          /* pinned host mirrors (BAMGPU_COPY_COLUMNS), else NULLs */
  uint64_t first_record_index;  /* index of record 0 of this batch in the whole stream */
  uint64_t bad_flag_records;    /* records with flag bits >= 0x1000 (reference coord sort would panic) */
  uint64_t chain_start;         /* offset in `data` of record 0's block_size field (== stream_start unless
                                   BAMGPU_SPECULATIVE_START); ~0 if the batch holds no record start */
  uint64_t chain_repairs;       /* tiles whose guessed entry had to be re-walked serially (0 in practice) */
  uint64_t data_start;          /* first valid byte of `data` / `h_data`: bytes [data_start, data_len) are the inflated stream
                                   (Reader batches keep a gap in front so that the tail of the previous batch - copied in
                                   from whichever GPU decoded it - and this batch's blocks are contiguous); 0 for engine calls */
  uint64_t stream_offset;       /* Reader batches: position in the whole inflated stream of data[data_start] */
} bamgpu_batch;

enum {
  BAMGPU_COPY_DATA = 1,         /* D2H the inflated bytes (needed by next_rec/read/append_record) */
  BAMGPU_COPY_COLUMNS = 2,      /* D2H the SoA columns */
  BAMGPU_WANT_HI = 4,           /* walk aux tags for HI (only the -n -M sort needs it) */
  BAMGPU_NO_CRC = 8,            /* skip the CRC32 check (default: verify) */
  BAMGPU_NO_RECORDS = 16,       /* inflate (+CRC) only: no record kernel (impl Read consumers) */
  BAMGPU_COPY_OFFSETS = 64,     /* D2H only rec_off + rec_len of the columns (with BAMGPU_COPY_DATA this is all that
                                   next_rec / append_record need: the reference lends bodies, nothing else) */
  BAMGPU_SORT_EMIT_BLOCK_SIZE = 128, /* bamgpu_engine_sort: gather (u32 block_size, body)* instead of bodies (sort.rs:340-347) */
  BAMGPU_DEFER_CRC = 256,       /* bamgpu_engine_inflate: do not wait for K1 / K2; the CRC32 pass runs on a side stream next to whatever the
                                   caller queues behind K1 (edge exchange, record kernels) and a failed block (inflate, ISIZE, CRC) is
                                   reported by the bamgpu_engine_records call that follows, before it hands anything out */
  BAMGPU_SPECULATIVE_START = 32 /* shard-edge mode (multi-GPU): stream_start is NOT a known record start; the chain
                                   begins at the first offset >= stream_start that looks like a run of records.
                                   The caller must confirm batch.chain_start with the left-hand shard's tail_off. */
};

#define BAMGPU_MAX_DEVICES 16
typedef struct bamgpu_opts {
  int32_t device;               /* CUDA device ordinal; -1 = current (used when n_devices == 0) */
  uint32_t flags;               /* BAMGPU_* bits */
  uint64_t batch_bytes;         /* compressed bytes per pipeline batch (0 = default: 256 MiB on one GPU, 64 MiB on several) */
  int32_t inflate_impl;         /* K1 variant: 0 = default (one BGZF block per lane pair), 1 = one warp per block
                                   (the north-star's mapping, kept for comparison; DESIGN.md "K1 variants") */
  int32_t jobs_per_device;      /* Reader: batches in flight per GPU (0 = default 8) */
  /* Multi-GPU Reader: batches go to the listed GPUs round robin, i.e. the file is sharded by BGZF block
   * ranges; the partial record at the end of a batch moves to the GPU that decodes the next batch peer-to-peer
   * (cudaMemcpyPeerAsync).  n_devices == 0: one GPU (`device`). */
  int32_t n_devices;
  int32_t devices[BAMGPU_MAX_DEVICES];
  int32_t sort_output;          /* bamgpu_sort_bam: BAMGPU_OUT_* (0 = the reference's output) */
  uint64_t sort_mem_limit;      /* bamgpu_sort_bam: sort.rs:138 `mem_limit` — bytes of DEVICE memory the record bodies may take.
                                   0 = the whole file is decoded and sorted in HBM if it fits there, else the streamed form below;
                                   > 0 = always the streamed form: the file goes through the Reader in batches of about
                                   mem_limit / 8 compressed bytes, only the sort keys (and, for the name orders, the read names)
                                   stay on the device, the bodies wait in host memory and are gathered there in sorted order
                                   (the reference's chunk files + N-way merge, sort.rs:185-308,590-652, without the files) */
  int32_t out_compr_level;      /* bamgpu_sort_bam with BAMGPU_OUT_BAM: sort.rs:143 `_out_compr_level` (which the reference accepts and never
                                   uses).  0 = stored DEFLATE blocks (copy speed); >= 1 = compressed blocks (K6: fixed-Huffman DEFLATE) */
  int32_t reserved;
} bamgpu_opts;

/* What bamgpu_sort_bam writes to the sink. */
enum {
  BAMGPU_OUT_BODIES = 0,   /* record bodies back to back: the reference's output (sort.rs:643, README.md:5) */
  BAMGPU_OUT_RECORDS = 1,  /* (u32 block_size, body)*: the reference's spill wire format (sort.rs:340-347) = a BAM record stream */
  BAMGPU_OUT_BAM = 2       /* a complete BAM file: "BAM\1" + the input's header + the sorted records, in BGZF blocks whose payload is
                              one stored DEFLATE block (0xff00 bytes, CRC32 + ISIZE footers), + the EOF marker - `samtools view -u` style */
};

/* Per-stage device times (CUDA events on the engine's stream) and counters, cumulative. */
typedef struct bamgpu_stats {
  uint64_t blocks, bytes_in, bytes_out, records, batches, kernel_launches;
  double ms_h2d, ms_inflate, ms_crc, ms_records, ms_d2h, ms_total;
  double ms_sort, ms_gather;      /* bamgpu_engine_sort: key sort / body gather */
  uint64_t p2p_bytes;             /* bytes moved between GPUs (batch-edge records, shard halos) */
  double ms_bgzf_write;           /* bamgpu_engine_bgzf_store: CRC + framing */
} bamgpu_stats;

/* ------------------------------------------------------------------------------------------
 * Engine: runs the kernels over a compressed buffer that is ALREADY in device memory.
 * This is what Reader uses per batch, and what bench.py times for the device-resident
 * `value` line.  One engine = one CUDA device + one stream.
 * ---------------------------------------------------------------------------------------- */
typedef struct bamgpu_engine bamgpu_engine;
typedef struct bamgpu_reader bamgpu_reader;

bamgpu_engine* bamgpu_engine_new(const bamgpu_opts* opts /*nullable*/);
void bamgpu_engine_free(bamgpu_engine*);
const char* bamgpu_engine_last_error(const bamgpu_engine*);

/* Device allocation helpers so callers without a CUDA binding (ctypes tests) can stage data. */
void* bamgpu_device_alloc(bamgpu_engine*, size_t bytes);
void bamgpu_device_free(bamgpu_engine*, void* dptr);
void* bamgpu_pinned_alloc(size_t bytes);
void bamgpu_pinned_free(void* p);
int bamgpu_memcpy_h2d(bamgpu_engine*, void* dst, const void* src, size_t bytes);
int bamgpu_memcpy_d2h(bamgpu_engine*, void* dst, const void* src, size_t bytes);

/* Decode `n_blocks` BGZF blocks whose payloads live in device buffer d_comp (in_off relative to it;
 * d_comp must be 16-byte aligned and readable for 64 bytes past the last payload).  `stream_start` is the offset in
 * the inflated stream where the record chain starts (reader.rs:63: right after the BAM header for the
 * first batch, the carry-over position for later ones).  If `cuda_stream` is non-NULL the work is
 * enqueued on that cudaStream_t instead of the engine's own (lets a caller time it with its own
 * events).  Synchronises before returning.  *tail_off gets the offset of the first byte not covered
 * by a complete record (== data_len when the chain ends exactly at the end or on a 0..3-byte tail /
 * block_size 0 with `final`), so a streaming caller can carry the partial record over. */
int bamgpu_engine_decode(bamgpu_engine*, const uint8_t* d_comp, const bamgpu_block* blocks, size_t n_blocks,
                         uint64_t stream_start, int final, uint32_t flags, void* cuda_stream,
                         bamgpu_batch* out, uint64_t* tail_off);

/* ---- The two halves of bamgpu_engine_decode, for callers that shard one file over several GPUs (bench.py, N > 1).
 * Shard g inflates its BGZF block range, copies the first bytes of shard g+1's inflated data behind its own
 * ("halo", so the record straddling the shard edge is whole), then runs the record kernels: records are reported
 * iff they START in the shard's own bytes.  Shards g > 0 pass BAMGPU_SPECULATIVE_START; batch.chain_start of shard
 * g+1 must equal shard g's *tail_off - own length (the caller checks it: one u64 per edge). ---- */
int bamgpu_engine_inflate(bamgpu_engine*, const uint8_t* d_comp, const bamgpu_block* blocks, size_t n_blocks,
                          uint32_t flags, void* cuda_stream);
/* Device pointer / length of the inflated bytes held by the engine (own bytes only). */
int bamgpu_engine_data(const bamgpu_engine*, const uint8_t** d_data, uint64_t* len);
/* Appends `len` bytes (device pointer, <= BAMGPU_MAX_HALO) behind the engine's own inflated bytes. */
#define BAMGPU_MAX_HALO (1u << 20)
int bamgpu_engine_append_halo(bamgpu_engine*, const uint8_t* d_src, size_t len, void* cuda_stream);
int bamgpu_engine_records(bamgpu_engine*, uint64_t stream_start, int final, uint32_t flags, void* cuda_stream,
                          bamgpu_batch* out, uint64_t* tail_off);

/* ---- Shard-edge exchange over NVLink peer memory (one process per GPU or several GPUs in one process) ----
 * bamgpu_engine_edge_open allocates this engine's inbox (two halo buffers + flag words) and exports it; the caller hands
 * the 80-byte handle to the two neighbours (any transport: it is plumbing, 80 bytes once per run) and calls _connect with
 * theirs.  bamgpu_engine_edge_exchange(step) then, on the given stream and without host synchronisation:
 *   - writes the first min(len, BAMGPU_MAX_HALO) inflated bytes of this shard into the LEFT neighbour's inbox
 *     (cudaMemcpyAsync to peer memory) and publishes `step` there;
 *   - waits (on the device) until the RIGHT neighbour has published `step` here and appends that halo behind this
 *     shard's own bytes (bamgpu_engine_append_halo), then acknowledges so the neighbour may reuse the buffer.
 * `step` starts at 1 and increases by one per call on every shard.  No collective, no NCCL. */
typedef struct bamgpu_edge_handle { uint8_t ipc[64]; uint64_t ptr; int32_t device; int32_t pid; } bamgpu_edge_handle;
int bamgpu_engine_edge_open(bamgpu_engine*, bamgpu_edge_handle* mine);
int bamgpu_engine_edge_connect(bamgpu_engine*, const bamgpu_edge_handle* left /*nullable*/, const bamgpu_edge_handle* right /*nullable*/);
int bamgpu_engine_edge_exchange(bamgpu_engine*, uint64_t step, void* cuda_stream);
void bamgpu_engine_edge_close(bamgpu_engine*);

/* ---- Digests of the decoded batch / the sort result, computed on the device from the engine's OUTPUTS (csrc/digest.cu).
 * A size-independent parity check: the CPU oracle folds its own results with the same published function
 * (oracle/bam_oracle.c "DIGEST SPEC").  index_base / stream_base are added to the record index / to rec_off, so shards
 * and batches of one file add up (all sums are mod 2^64).  Synchronises. */
typedef struct bamgpu_digest {
  uint64_t n_records;
  uint64_t body_bytes;   /* sum of block_size */
  uint64_t fields;       /* records: index, stream offset, block_size, 11 fixed fields, 4 offsets, keys, HI;  sort: the permutation */
  uint64_t bodies;       /* every body byte, position- and order-dependent */
} bamgpu_digest;
int bamgpu_engine_digest(bamgpu_engine*, uint64_t index_base, uint64_t stream_base, void* cuda_stream, bamgpu_digest* out);
int bamgpu_engine_sorted_digest(bamgpu_engine*, void* cuda_stream, bamgpu_digest* out);
/* The engine that holds the batch bamgpu_next_batch returned last (for bamgpu_engine_digest / _sort on Reader batches). */
bamgpu_engine* bamgpu_reader_current_engine(bamgpu_reader*);

/* ---- Variable-length fields of the batch the engine decoded last, decoded like the reference's accessors:
 *   BAMGPU_FIELD_CIGAR  decode_cigar(get_bytes(RawCigar))   record/bamrawrecord.rs:223-236 over get_cigar :126-157:
 *                       "<len><op>" per operation ("" for refID < 0, pos < 0 or n_cigar_op == 0); a two-operation CIGAR whose
 *                       first operation is `<x>S` with x == (l_seq + 1) / 2 (sic: the reference compares with the packed
 *                       sequence's BYTE length) is replaced by the value of the CG tag if there is one;
 *   BAMGPU_FIELD_SEQ    decode_seq(get_bytes(RawSequence))  :259-270: "=ACMGRSVTWYHKDBN"[nibble]; every low nibble equal to 0
 *                       is dropped (the reference's stated assumption), not only the padding of an odd-length read;
 *   BAMGPU_FIELD_QUAL   get_bytes(RawQual)                  :99: the l_seq raw phred bytes.
 * Each field comes back as n_records + 1 offsets and the concatenated bytes, in device memory; with BAMGPU_COPY_DATA also in
 * pinned host memory (h_*).  Records for which the reference would panic (slice beyond the record, CIGAR bytes not a multiple
 * of 4, operation code > 8, malformed tag in front of CG) get an empty field and are counted in bad[].  Offsets use the
 * get_bytes closures' u8 arithmetic (`32 + l_read_name` wraps for l_read_name >= 224, :80-84).  Valid until the next decode or
 * decode_fields call on the engine.  Synchronises. */
enum { BAMGPU_FIELD_CIGAR = 1, BAMGPU_FIELD_SEQ = 2, BAMGPU_FIELD_QUAL = 4 };
typedef struct bamgpu_field {
  uint64_t total;               /* bytes in `bytes` */
  uint64_t bad;                 /* records whose field the reference would panic on (empty here) */
  const uint64_t* off;          /* device: n_records + 1 offsets */
  const uint8_t* bytes;         /* device */
  const uint64_t* h_off;        /* pinned host mirrors (BAMGPU_COPY_DATA), else NULL */
  const uint8_t* h_bytes;
} bamgpu_field;
typedef struct bamgpu_fields {
  uint64_t n_records;
  bamgpu_field cigar, seq, qual;
  double ms_measure, ms_decode;  /* device time of the two passes */
} bamgpu_fields;
int bamgpu_engine_decode_fields(bamgpu_engine*, uint32_t which /* BAMGPU_FIELD_* */, uint32_t flags /* BAMGPU_COPY_DATA */,
                                void* cuda_stream, bamgpu_fields* out);
/* bodies-digest (DIGEST SPEC) of one decoded field: n_records, body_bytes = total, bodies = sum_i mix64(h_body(field_i) + i * K). */
int bamgpu_engine_field_digest(bamgpu_engine*, uint32_t field /* one BAMGPU_FIELD_* bit */, void* cuda_stream, bamgpu_digest* out);

/* Number of reference sequences of the BAM header (n_ref), if the caller knows it: tightens the guess of record
 * starts inside K3 (refID must be in [-1, n_ref)).  Results never depend on it.  < 0 = unknown (default). */
void bamgpu_engine_set_n_ref(bamgpu_engine*, int32_t n_ref);

int bamgpu_engine_stats(const bamgpu_engine*, bamgpu_stats* out);
void bamgpu_engine_reset_stats(bamgpu_engine*);

/* ------------------------------------------------------------------------------------------
 * Reader: streaming drop-in for bam_tools::Reader over a caller-supplied byte source.
 * ---------------------------------------------------------------------------------------- */
/* std::io::Read::read: return n > 0 bytes written to dst, 0 at EOF, < 0 on error. May be called from
 * an internal thread (the reference moves `inner` to its reading thread too: readahead.rs:88). */
typedef long (*bamgpu_read_fn)(void* ctx, uint8_t* dst, size_t cap);

/* Reader::new (reader.rs:28-55).  thread_num is accepted for signature parity and clamped like the
 * reference ([3, num_cpus]); the GPU path needs one host feeder thread regardless.  track_progress
 * (nullable) = total compressed size, kept as a byte counter (no progress bar). */
bamgpu_reader* bamgpu_reader_new(bamgpu_read_fn read, void* ctx, size_t thread_num, const uint64_t* track_progress,
                                 const bamgpu_opts* opts /*nullable*/);
/* Convenience: reader over an in-memory BGZF buffer (not copied; must outlive the reader). */
bamgpu_reader* bamgpu_reader_from_memory(const uint8_t* buf, size_t len, size_t thread_num, const bamgpu_opts* opts);
void bamgpu_reader_free(bamgpu_reader*);
const char* bamgpu_last_error(const bamgpu_reader*);

/* Reader::read_header (reader.rs:87-98): *bytes = l_text|text|n_ref|(l_name|name|l_ref)* WITHOUT the
 * 4 magic bytes, owned by the reader; *offset_to_ref_seqs = 4 + l_text. */
int bamgpu_read_header(bamgpu_reader*, const uint8_t** bytes, size_t* len, size_t* offset_to_ref_seqs);

/* Records::next_rec (records.rs:23-29): lends the next record body; valid until the next call.
 * Returns BAMGPU_OK, BAMGPU_EOF (None) or an error. */
int bamgpu_next_rec(bamgpu_reader*, const uint8_t** body, size_t* len);
/* Reader::append_record (reader.rs:63-73): copies the next body to dst[0..cap); *len = block_size,
 * 0 at EOF. BAMGPU_ERR_BUFFER_TOO_SMALL leaves the record unconsumed and sets *len to the size needed. */
int bamgpu_append_record(bamgpu_reader*, uint8_t* dst, size_t cap, size_t* len);
/* impl Read for Reader (reader.rs:114-152): inflated byte stream; returns bytes read, 0 at EOF, <0 error.
 * Unlike the reference it never returns a short "Interrupted" read at block switches.  As in the reference it may be
 * called after read_header() and continues behind the header; mixing it with the record calls is an error here
 * (the reference would hand out torn records). */
long bamgpu_read(bamgpu_reader*, uint8_t* dst, size_t cap);
/* GPU-native batch interface (no reference counterpart): all records of the next pipeline batch. */
int bamgpu_next_batch(bamgpu_reader*, bamgpu_batch* out);
int bamgpu_reader_stats(const bamgpu_reader*, bamgpu_stats* out);

/* parse_reference_sequences (reader.rs:101-112,213-228) over header bytes starting at n_ref.
 * Calls cb(name_without_nul, name_len, l_ref, user) per reference; returns BAMGPU_OK or
 * BAMGPU_ERR_SHORT_HEADER (short input / name not NUL-terminated or with an interior NUL). */
typedef void (*bamgpu_refseq_cb)(const char* name, size_t name_len, uint32_t l_ref, void* user);
int bamgpu_parse_reference_sequences(const uint8_t* bytes, size_t len, bamgpu_refseq_cb cb, void* user);

/* ------------------------------------------------------------------------------------------
 * Sort: what the reference's sort modes do with the records the reader hands them
 *     sort_bam(mem_limit, reader, sorted_sink, tmp_dir, .., reader_thread_num, temp_files_mode, None, sort_by, ..)
 *                                                             bam_tools/src/sorting/sort.rs:138-178
 *     SortBy { Name, NameAndMatchMates, CoordinatesAndStrand }               sort.rs:92-96
 *     comparators                                                            sorting/comparators.rs:61-147
 * The output is the concatenation of the record BODIES in globally stable sorted order (sort.rs:643: no block_size,
 * no header, no BGZF).  On the GPU everything stays resident in HBM: no chunks, temp files or merge
 * (mem_limit / tmp_dir / temp_files_mode have no counterpart).
 * ---------------------------------------------------------------------------------------- */
typedef enum bamgpu_sort_by {
  BAMGPU_SORT_NAME = 0,                    /* bam_binary -n       : read_name bytes incl. NUL, bytewise */
  BAMGPU_SORT_NAME_AND_MATCH_MATES = 1,    /* bam_binary -n -M    : name, then HI, then the full u16 flag; needs BAMGPU_WANT_HI at decode */
  BAMGPU_SORT_COORDINATES_AND_STRAND = 2   /* bam_binary (default): (refID == -1 ? i32::MAX : refID, pos, reverse strand) */
} bamgpu_sort_by;

typedef struct bamgpu_sorted {
  uint64_t n_records;
  uint64_t bodies_len;          /* sum of block_size over all records */
  const uint32_t* perm;         /* device: perm[i] = index in the decoded batch of the i-th record of the sorted order */
  const uint64_t* body_off;     /* device: n_records + 1 offsets of the sorted bodies in `bodies` */
  const uint8_t* bodies;        /* device: the reference's output bytes */
  const uint32_t* h_perm;       /* pinned host mirror (BAMGPU_COPY_COLUMNS), else NULL */
  const uint8_t* h_bodies;      /* pinned host mirror (BAMGPU_COPY_DATA), else NULL */
} bamgpu_sorted;

/* Sorts the records of the batch the engine decoded last (bamgpu_engine_decode / _records / bamgpu_next_batch on the
 * owning reader) and gathers their bodies.  Valid until the next decode or sort on the engine.
 * Errors where the reference would panic: BAMGPU_ERR_SORT_HI_MIXED, _BAD_TAGS (NameAndMatchMates), _BAD_FLAGS
 * (CoordinatesAndStrand). */
int bamgpu_engine_sort(bamgpu_engine*, int sort_by, uint32_t flags /* BAMGPU_COPY_* */, void* cuda_stream, bamgpu_sorted* out);

/* std::io::Write::write_all: return 0, or < 0 on error. */
typedef int (*bamgpu_write_fn)(void* ctx, const uint8_t* src, size_t len);
/* sort_bam (sort.rs:138-178): reads the whole BGZF stream from `read`, decodes and sorts it on the device and writes
 * the sorted bodies to `write`.  The header is read and discarded like the reference does (sort.rs:152-153).
 * A file that does not fit in device memory as one batch (or any file, with bamgpu_opts.sort_mem_limit) is sorted in the
 * streamed form: batches through the Reader, keys on the device, bodies in host memory.  *stats (nullable) gets the stage times. */
int bamgpu_sort_bam(bamgpu_read_fn read, void* read_ctx, bamgpu_write_fn write, void* write_ctx, size_t reader_thread_num,
                    int sort_by, const bamgpu_opts* opts /*nullable*/, bamgpu_stats* stats /*nullable*/,
                    char* err_buf /*nullable*/, size_t err_cap);
/* The same over a BGZF file that is already in host memory (not copied; like bamgpu_reader_from_memory a page-locked buffer is
 * read by the GPUs straight from where it is).
 * Several GPUs (bamgpu_opts.devices[], both entry points; SURVEY.md 8 f1 "multi-GPU => sample sort"): without sort_mem_limit the
 * file is cut into one contiguous BGZF block range per GPU, the ranges are decoded side by side, sampled keys give n - 1 splitters,
 * every GPU partitions its records by bucket and writes bucket d's (u32 block_size, body)* stream into GPU d's memory
 * (cudaMemcpyPeerAsync), and GPU d parses and sorts what it received; the buckets go to the sink in order.  Output bytes are those of
 * the single-GPU sort in every format.  Inputs this does not apply to (fewer blocks than GPUs, a record longer than
 * BAMGPU_MAX_HALO across a range edge, too little device memory) are sorted by the single-GPU / streamed path instead. */
int bamgpu_sort_bam_from_memory(const uint8_t* buf, size_t len, bamgpu_write_fn write, void* write_ctx, size_t reader_thread_num,
                                int sort_by, const bamgpu_opts* opts /*nullable*/, bamgpu_stats* stats /*nullable*/,
                                char* err_buf /*nullable*/, size_t err_cap);

/* BGZF writer with stored DEFLATE blocks: wraps `len` bytes of device memory (a BAM byte stream) into BGZF
 * blocks of 0xff00 payload bytes with CRC32 + ISIZE footers (gz.rs:9-20 framing), optionally followed by the 28-byte EOF
 * marker.  d_out must hold bamgpu_bgzf_bound(len) bytes; *out_len gets the bytes written.  Synchronises. */
size_t bamgpu_bgzf_bound(size_t len);
int bamgpu_engine_bgzf_store(bamgpu_engine*, const uint8_t* d_src, size_t len, uint8_t* d_out, size_t out_cap, size_t* out_len,
                             int eof_marker, void* cuda_stream);

/* The same framing with COMPRESSED payloads (K6, csrc/deflate.cu): level <= 0 = bamgpu_engine_bgzf_store; level >= 1 = every payload becomes
 * one DEFLATE block with the fixed Huffman code (RFC 1951 3.2.6; LZ77 matches from a hash of 4-byte sequences and a distance-1 probe,
 * greedy parse), or a stored block when that is not smaller.  The same bytes on every run.  d_out must hold bamgpu_bgzf_bound(len). */
int bamgpu_engine_bgzf_write(bamgpu_engine*, const uint8_t* d_src, size_t len, uint8_t* d_out, size_t out_cap, size_t* out_len,
                             int eof_marker, int level, void* cuda_stream);

const char* bamgpu_status_str(int status);
/* "sm_100a;<build id>" — lets callers assert the CUDA library (not a fallback) is loaded. */
const char* bamgpu_build_info(void);

#ifdef __cplusplus
}
#endif
#endif /* BAMGPU_H */
