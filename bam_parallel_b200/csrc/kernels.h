// kernels.h — internal launch interface between the host engine (engine.cpp) and the CUDA kernels.
// Not part of the public ABI (that is include/bamgpu.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bamgpu {

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Device-side view of the BGZF block table of one batch (structure of arrays).
struct BlockTableDev {
  const uint64_t* in_off;   // payload offset in the compressed buffer
  const uint32_t* in_len;   // payload bytes
  const uint64_t* out_off;  // offset in the inflated buffer (exclusive scan of ISIZE, + carry)
  const uint32_t* isize;    // ISIZE footer
  const uint32_t* crc;      // CRC32 footer
  uint32_t n_blocks;
};

// ---- K1: DEFLATE inflate, one BGZF block per lane (inflate.cu) -------------------------------
// status[b] = IST_* (inflate_core.cuh).  work_counter: one zeroed u32 in device memory.
void launch_inflate(const uint8_t* comp, const BlockTableDev& t, uint8_t* out, uint32_t* status,
                    uint32_t* work_counter, void* ring_mem, int num_sms, cudaStream_t s);
// Variant 1 (bamgpu_opts.inflate_impl = 1): one warp per BGZF block, the north-star's mapping (inflate_warp.cu).
void launch_inflate_warp_per_block(const uint8_t* comp, const BlockTableDev& t, uint8_t* out, uint32_t* status,
                                   uint32_t* work_counter, int num_sms, cudaStream_t s);
uint32_t inflate_grid(uint32_t n_blocks, int num_sms);   // CTAs launch_inflate uses for that many blocks
size_t inflate_smem_bytes();
size_t inflate_ring_bytes(int num_sms);   // device scratch for the token rings when they live in global memory (else 0)

// ---- K2: CRC32 of every inflated block vs its footer (crc32.cu) ------------------------------
// Sets status[b] = IST_CRC_MISMATCH where status[b]==0 and the CRC differs.
constexpr uint32_t IST_CRC_MISMATCH = 100;
void crc32_init_tables(uint32_t* d_tables /* CRC_TABLE_WORDS words */, cudaStream_t s);
constexpr int CRC_TABLE_WORDS = 2 * 4 * 256;
// small_footprint: the 4-warp / 36 KB shape that fits beside resident K1 CTAs (Reader jobs), else one 32-warp CTA per SM.
void launch_crc32(const uint8_t* out, const BlockTableDev& t, const uint32_t* d_tables, uint32_t* status,
                  int num_sms, bool small_footprint, cudaStream_t s);

// Copies `bytes` (multiple of 4, <= 4096) from device memory to PINNED HOST memory with SM stores.  Result words of a few
// bytes must not go through cudaMemcpyAsync: the D2H copy engine serves transfers in order, so an 8-byte status read queues
// behind a gigabyte of record bodies of the previous batch.
void launch_copy_small_to_host(void* h_pinned, const void* d_src, uint32_t bytes, cudaStream_t s);

// first_bad[0] = min over blocks with status != 0 of (block index << 32 | status); ~0ull if none.
void launch_status_reduce(const uint32_t* status, uint32_t n_blocks, unsigned long long* first_bad, cudaStream_t s);

// ---- K3: record boundaries, fixed fields, sort keys (records.cu) -----------------------------
constexpr uint32_t TILE_BYTES = 1u << 16;   // inflated bytes per K3 tile (one thread walks one tile); 16 KiB tiles measured slower (7.7 vs 6.3 ms)
constexpr uint32_t WILD_JUMP_BYTES = 1u << 16;  // a guessed record may end this far past the data, not farther

// exit_kind values
enum : uint32_t {
  EXIT_NORMAL = 0,      // exit offset is the start of the first record beginning at/after the tile end
  EXIT_TAIL = 1,        // fewer than 4 bytes left at exit offset (reader.rs:78 UnexpectedEof => 0)
  EXIT_ZERO = 2,        // block_size == 0 at exit offset (reader.rs:64-72: iteration ends)
  EXIT_INCOMPLETE = 3,  // record at exit offset runs past the data (needs more bytes / SHORT_RECORD)
  EXIT_NONE = 4,        // tile has no candidate entry
  EXIT_LIMIT = 5,       // exit offset is at/after own_len: the next record starts in the halo, i.e. belongs to the right-hand shard
};

struct TileArrays {
  uint64_t* entry;      // speculated (or known) first record start in the tile; ~0ull if none
  uint64_t* exit_off;   // where the walk stopped
  uint32_t* count;      // records STARTING in this tile when entered at entry
  uint32_t* exit_kind;
  uint32_t* nxt;        // pointer-jumping: successor tile (self if terminal/broken)
  uint32_t* jump;       // doubling scratch
  uint32_t* jump2;
  uint8_t* reach;       // on the true chain
  uint64_t* base;       // exclusive scan of count over reachable tiles
  uint32_t n_tiles;
  uint16_t* ckpt;       // [n_tiles][CK_PER_TILE]: tile-relative start of every CK_STRIDE-th record of the tile
};
constexpr uint32_t CK_STRIDE = 8;
constexpr uint32_t CK_PER_TILE = (TILE_BYTES / 36 + CK_STRIDE) / CK_STRIDE;   // a record takes >= 36 bytes of stream: <= 1821 starts per tile

struct RecordParams {
  const uint8_t* data;  // inflated bytes (carry + new); readable 16 bytes before and after
  uint64_t data_len;    // bytes present (own bytes + halo copied from the right-hand shard)
  uint64_t own_len;     // records are reported iff they START before own_len (== data_len without a halo)
  uint64_t start;       // chain start (absolute offset in data)
  int32_t n_ref;        // number of reference sequences (plausibility filter); <0 = unknown
  int32_t start_known;  // 1: `start` is a record start (reader.rs:63); 0: shard-edge mode, the chain begins at the
                        //    first position >= start that looks like a chain of records (caller verifies it)
};

// Output columns (device pointers, capacity >= n_records). Mirrors bamgpu_columns.
struct ColumnsDev {
  uint64_t* rec_off; uint32_t* rec_len;
  int32_t* ref_id; int32_t* pos; uint8_t* l_read_name; uint8_t* mapq; uint16_t* bin; uint16_t* n_cigar_op;
  uint16_t* flag; uint32_t* l_seq; int32_t* next_ref_id; int32_t* next_pos; int32_t* tlen;
  uint32_t* cigar_off; uint32_t* seq_off; uint32_t* qual_off; uint32_t* tags_off;
  uint64_t* coord_key; uint8_t* strand; uint8_t* hi_present; int32_t* hi_value;
};

// Result block written by the chain kernels (device), copied to the host once per batch.
struct ChainResult {
  unsigned long long n_records;
  unsigned long long tail_off;     // first byte not covered by a complete record
  uint32_t tail_kind;              // EXIT_* at the end of the true chain
  uint32_t broken_tile;            // first reachable tile whose successor's speculation mismatched; ~0u if none
  unsigned long long broken_entry; // the true entry for that successor tile
  unsigned long long bad_flags;    // records with flag bits >= 0x1000
  uint32_t rounds;                 // serial repairs that were needed (0 when every guess was right)
  uint32_t pad;
  unsigned long long chain_start;  // where the chain began (== start when start_known); ~0ull if nowhere
};

void launch_tile_walk(const RecordParams& p, const TileArrays& t, cudaStream_t s);
// Link + pointer-doubling reachability + (rare) serial repair + masked scan of counts + result summary, one launch.
constexpr int RESOLVE_MAX_CTAS = 256;
struct ResolveScratch {  // grid-wide state of k3_resolve (device memory, no initialisation needed)
  uint32_t first, broken_tile, tail_kind, repairs;
  unsigned long long broken_entry, tail_off;
  unsigned long long chunk_sum[RESOLVE_MAX_CTAS];
};
cudaError_t launch_chain_resolve(const RecordParams& p, const TileArrays& t, ChainResult* d_result, ResolveScratch* d_scratch,
                                 int num_sms, cudaStream_t s);
// Second walk: writes rec_off / rec_len for reachable tiles.
void launch_tile_emit(const RecordParams& p, const TileArrays& t, uint64_t* rec_off, uint32_t* rec_len, uint64_t cap, cudaStream_t s);
// Per-record fixed fields, variable offsets, keys (and HI when want_hi).  dev_count: n_records is the columns' capacity and
// the true count is read from d_result->n_records on the device.
void launch_extract(const RecordParams& p, const ColumnsDev& c, uint64_t n_records, bool want_hi,
                    ChainResult* d_result, bool dev_count, cudaStream_t s);

// ---- K5: variable-length field decode (fields.cu) ---------------------------------------------------
// Field k = 0 CIGAR string, 1 sequence string, 2 raw quality bytes (bit 1 << k of `which` = BAMGPU_FIELD_*).
constexpr uint32_t BAMGPU_FIELD_CIGAR_BIT = 1, BAMGPU_FIELD_SEQ_BIT = 2, BAMGPU_FIELD_QUAL_BIT = 4;
struct FieldsDev {
  uint64_t* off[3];       // n + 1 offsets per field (lengths before the scan)
  uint8_t* out[3];        // the decoded bytes
  uint32_t* cig_src;      // body-relative offset of the record's CIGAR bytes (the field, or the CG tag's value)
  uint32_t* cig_nbytes;
  uint8_t* flags;         // per record: bit k set = the reference would panic decoding field k (output empty)
};
size_t fields_scratch_bytes(uint64_t n);
cudaError_t fields_measure(const uint8_t* data, const struct ColumnsDev& c, uint64_t n, uint32_t which, void* scratch, FieldsDev* f,
                           unsigned long long* d_bad, uint64_t* totals, uint64_t* bad, int num_sms, cudaStream_t s);   // synchronises
void launch_fields_decode(const uint8_t* data, const struct ColumnsDev& c, uint64_t n, uint32_t which, const FieldsDev& f, int num_sms,
                          cudaStream_t s);

// ---- digests of the decode / sort outputs (digest.cu; DIGEST SPEC in DESIGN.md) ---------------------
// acc: 4 x u64 in device memory {n_records (unused), body_bytes, fields | perm, bodies}
void launch_digest_records(const uint8_t* data, const ColumnsDev& c, uint64_t n, uint64_t index_base, uint64_t stream_base,
                           unsigned long long* acc, int num_sms, cudaStream_t s);
void launch_digest_sorted(const uint8_t* bodies, const uint64_t* body_off, const uint32_t* perm, uint64_t n,
                          unsigned long long* acc, int num_sms, cudaStream_t s);
// bodies-digest of any ragged byte array (n + 1 offsets): K5's decoded fields
void launch_digest_ragged(const uint8_t* bytes, const uint64_t* off, uint64_t n, unsigned long long* acc, int num_sms, cudaStream_t s);

// ---- shard-edge flags in peer memory (edge.cu) -----------------------------------------------------------------
// Spins (with a time-out) until *flag >= want; on time-out sets *timed_out = 1 and returns.
void launch_wait_flag(const unsigned long long* flag, unsigned long long want, uint32_t* timed_out, cudaStream_t s);
// __threadfence_system(); *flag = value  (flag may live in a peer GPU's memory)
void launch_set_flag(unsigned long long* flag, unsigned long long value, cudaStream_t s);

// ---- sort + gather (sort.cu): the output bytes of the reference's sort_bam (sorting/sort.rs:138-178) -------------
enum : int { SORT_NAME = 0, SORT_NAME_AND_MATCH_MATES = 1, SORT_COORDINATES_AND_STRAND = 2 };   // sort.rs:92-96 SortBy
struct SortResult {
  unsigned long long hi_mixed;     // neighbours with equal names where only one has HI (comparators.rs:90 would panic)
  unsigned long long hi_bad_tags;  // records whose aux walk the reference would panic on (hi_present == 2)
  uint32_t max_name_len;
  uint32_t pad;
};
size_t sort_scratch_bytes(uint64_t n_records);
cudaError_t sort_permutation(const uint8_t* data, const ColumnsDev& c, uint64_t n_records, int sort_by, void* scratch, uint32_t* perm,
                             SortResult* d_res, int* launches, cudaStream_t s);
// prefix = 4: every body is preceded by its u32 block_size (the reference's spill wire format, sort.rs:340-347); 0: bodies only (sort.rs:643)
cudaError_t sorted_offsets(const ColumnsDev& c, const uint32_t* perm, uint64_t n_records, void* scratch, uint64_t* out_off,
                           uint64_t* total_bytes, uint32_t prefix, cudaStream_t s);
cudaError_t gather_sorted_bodies(const uint8_t* data, const ColumnsDev& c, const uint32_t* perm, uint64_t n_records, void* scratch,
                                 uint64_t* out_off, uint8_t* out, uint32_t prefix, cudaStream_t s);

// sample sort over several GPUs (bamgpu_sort_bam with a device list): one sampled key / one splitter
struct SplitKey {
  uint64_t key;          // coord_key          (CoordinatesAndStrand)
  uint32_t strand;
  uint32_t name_len;     // l_read_name        (Name, NameAndMatchMates: buckets are cut by the name alone)
  uint8_t name[256];
};
cudaError_t sample_keys(const uint8_t* data, const ColumnsDev& c, uint64_t n_records, int sort_by, uint32_t n_samples, SplitKey* d_out, cudaStream_t s);
cudaError_t partition_by_splitters(const uint8_t* data, const ColumnsDev& c, uint64_t n_records, int sort_by, const SplitKey* d_split,
                                   uint32_t n_split, void* scratch, uint32_t* perm, uint32_t* d_bounds, int* launches, cudaStream_t s);

// streamed sort (bodies on the host, keys on the device): the names of a batch -> compact store + virtual record offsets
cudaError_t names_total(const ColumnsDev& c, uint64_t n_records, void* scratch, uint64_t* name_bytes, cudaStream_t s);   // synchronises
cudaError_t pack_names(const uint8_t* data, const ColumnsDev& c, uint64_t n_records, void* scratch, uint8_t* dst, uint64_t store_base,
                       uint64_t* voff, cudaStream_t s);

// ---- BGZF writer, stored blocks (bgzf_write.cu) ---------------------------------------------------
// CRC-32 of consecutive `chunk`-byte pieces of src (the last one shorter) -> crc_out[i]   (crc32.cu, K2's warp CRC)
void launch_crc32_chunks(const uint8_t* src, uint64_t len, uint32_t chunk, const uint32_t* d_tables, uint32_t* crc_out, int num_sms, cudaStream_t s);
// Frames: 18-byte BGZF header + 5-byte stored-block header in front of every payload, CRC32 + ISIZE behind it, EOF marker.
constexpr uint32_t BGZF_STORE_PAYLOAD = 0xff00;                       // bytes per block, like htslib
constexpr uint32_t BGZF_STORE_BLOCK = 18 + 5 + BGZF_STORE_PAYLOAD + 8; // a full block on the wire
cudaError_t bgzf_store_blocks(const uint8_t* src, uint64_t len, const uint32_t* crc, uint8_t* out, bool eof_marker, cudaStream_t s);


// ---- K6: BGZF writer, compressed blocks (deflate.cu) -----------------------------------------------
// Every 0xff00-byte payload -> one DEFLATE block: level 1 the fixed Huffman code, level >= 2 a code built for the block (or
// the fixed one if that is smaller); a stored block when neither is smaller than the payload.
size_t bgzf_deflate_scratch_bytes(uint64_t len, int num_sms);
cudaError_t bgzf_deflate_blocks(const uint8_t* src, uint64_t len, const uint32_t* crc, void* scratch, uint8_t* out, bool eof_marker, int level,
                                uint64_t* out_len, int num_sms, cudaStream_t s);   // synchronises

}  // namespace bamgpu
