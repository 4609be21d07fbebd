// bamgen — seeded synthetic BAM writer (C++ + zlib, multi-threaded).
//
// There is no network on the build or GPU boxes, so every workload named in
// BASELINE.json `configs` is produced by this generator at test / bench time.
// It is NOT part of the product path (nothing under bam_parallel_b200/csrc
// links it): it only manufactures inputs.
//
// Output is a spec-conformant BAM: "BAM\1" header + records, cut into BGZF
// blocks of `payload` uncompressed bytes (records freely straddle blocks, as
// htslib/sambamba writers produce for long records), each block raw-DEFLATEd
// with zlib at the requested level, followed by the 28-byte BGZF EOF marker.
//
// Shapes:
//   se150    single-end 150 bp, flag in {0,16}                    (config 1)
//   pe150    paired 2x150 bp, flags {99,147,83,163}, ~2% unmapped (configs 2-4)
//   longname names of 200..222 bytes, HI:i on every record of multi-hit groups (config 5a)
//   long10k  ~10 kbp reads, multi-op CIGAR, ~15 KB bodies        (config 5b)
//
// Every record is a pure function of (seed, shape, record index), so threads
// generate disjoint BGZF block ranges independently and the file is
// byte-identical for any thread count.
//
// Build: see top-level Makefile (libbamgen.so + bamgen CLI).

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Rng {  // splitmix64
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  uint32_t u32() { return (uint32_t)(next() >> 32); }
  uint32_t below(uint32_t n) { return (uint32_t)(((uint64_t)u32() * n) >> 32); }
};

inline uint64_t mix(uint64_t a, uint64_t b, uint64_t c) {
  Rng r(a * 0xD1342543DE82EF95ull + b * 0xA24BAED4963EE407ull + c);
  r.next();
  return r.next();
}

enum Shape { SE150 = 0, PE150 = 1, LONGNAME = 2, LONG10K = 3 };

constexpr int N_REF = 24;
const char* const REF_NAMES[N_REF] = {"chr1",  "chr2",  "chr3",  "chr4",  "chr5",  "chr6",  "chr7",  "chr8",
                                      "chr9",  "chr10", "chr11", "chr12", "chr13", "chr14", "chr15", "chr16",
                                      "chr17", "chr18", "chr19", "chr20", "chr21", "chr22", "chrX",  "chrY"};
const uint32_t REF_LENS[N_REF] = {249250621, 243199373, 198022430, 191154276, 180915260, 171115067,
                                  159138663, 146364022, 141213431, 135534747, 135006516, 133851895,
                                  115169878, 107349540, 102531392, 90354753,  81195210,  78077248,
                                  59128983,  63025520,  48129895,  51304566,  155270560, 59373566};

inline void put_u16(uint8_t* p, uint16_t v) { p[0] = v & 0xff; p[1] = v >> 8; }
inline void put_u32(uint8_t* p, uint32_t v) {
  p[0] = v & 0xff; p[1] = (v >> 8) & 0xff; p[2] = (v >> 16) & 0xff; p[3] = v >> 24;
}

// SAM spec §5.3 reg2bin.
inline uint16_t reg2bin(int64_t beg, int64_t end) {
  --end;
  if (beg >> 14 == end >> 14) return (uint16_t)(((1 << 15) - 1) / 7 + (beg >> 14));
  if (beg >> 17 == end >> 17) return (uint16_t)(((1 << 12) - 1) / 7 + (beg >> 17));
  if (beg >> 20 == end >> 20) return (uint16_t)(((1 << 9) - 1) / 7 + (beg >> 20));
  if (beg >> 23 == end >> 23) return (uint16_t)(((1 << 6) - 1) / 7 + (beg >> 23));
  if (beg >> 26 == end >> 26) return (uint16_t)(((1 << 3) - 1) / 7 + (beg >> 26));
  return 0;
}

struct Fixed {
  int32_t ref_id = -1, pos = -1;
  uint8_t mapq = 0;
  uint16_t flag = 4;
  int32_t next_ref = -1, next_pos = -1, tlen = 0;
  uint32_t ref_span = 0;  // for bin
};

// Writes block_size + body at `out`; returns total bytes (4 + block_size).
size_t write_record(uint8_t* out, const Fixed& f, const char* name, size_t l_name /*incl NUL*/,
                    const uint32_t* cigar, uint32_t n_cigar, uint32_t l_seq, Rng& r, int qual_mode,
                    const uint8_t* tags, size_t l_tags) {
  uint8_t* b = out + 4;
  put_u32(b + 0, (uint32_t)f.ref_id);
  put_u32(b + 4, (uint32_t)f.pos);
  b[8] = (uint8_t)l_name;
  b[9] = f.mapq;
  uint16_t bin = (f.ref_id < 0 || f.pos < 0) ? 4680 : reg2bin(f.pos, (int64_t)f.pos + (f.ref_span ? f.ref_span : 1));
  put_u16(b + 10, bin);
  put_u16(b + 12, (uint16_t)n_cigar);
  put_u16(b + 14, f.flag);
  put_u32(b + 16, l_seq);
  put_u32(b + 20, (uint32_t)f.next_ref);
  put_u32(b + 24, (uint32_t)f.next_pos);
  put_u32(b + 28, (uint32_t)f.tlen);
  uint8_t* p = b + 32;
  memcpy(p, name, l_name);
  p += l_name;
  for (uint32_t i = 0; i < n_cigar; ++i, p += 4) put_u32(p, cigar[i]);
  // seq: 4-bit codes A=1 C=2 G=4 T=8
  static const uint8_t code[4] = {1, 2, 4, 8};
  uint32_t nb = (l_seq + 1) / 2;
  for (uint32_t i = 0; i < nb; i += 16) {
    uint64_t x = r.next();
    for (uint32_t j = 0; j < 16 && i + j < nb; ++j, x >>= 4) {
      uint8_t hi = code[x & 3], lo = code[(x >> 2) & 3];
      p[i + j] = (uint8_t)((hi << 4) | lo);
    }
  }
  if (l_seq & 1) p[nb - 1] &= 0xf0;
  p += nb;
  if (qual_mode == 0) {
    // Illumina-style binned qualities {2,11,25,37}, sticky runs skewed to 37.
    static const uint8_t bins[4] = {37, 25, 11, 2};
    uint32_t cur = 0;
    uint64_t x = 0;
    int have = 0;
    for (uint32_t i = 0; i < l_seq; ++i) {
      if (have == 0) { x = r.next(); have = 16; }
      uint32_t d = x & 15; x >>= 4; --have;
      if (d == 0) cur = (cur + 1) & 3;        // degrade
      else if (d == 1) cur = 0;               // recover
      else if (d == 2 && i > l_seq * 3 / 4) cur = 3;  // tail drop
      p[i] = bins[cur];
    }
  } else {
    // long-read style: noisy phred 1..40
    for (uint32_t i = 0; i < l_seq; i += 8) {
      uint64_t x = r.next();
      for (uint32_t j = 0; j < 8 && i + j < l_seq; ++j, x >>= 8) p[i + j] = (uint8_t)(1 + ((x & 0xff) * 40 >> 8));
    }
  }
  p += l_seq;
  memcpy(p, tags, l_tags);
  p += l_tags;
  uint32_t block_size = (uint32_t)(p - b);
  put_u32(out, block_size);
  return 4 + (size_t)block_size;
}

size_t tag_C(uint8_t* t, char a, char b, uint8_t v) { t[0] = a; t[1] = b; t[2] = 'C'; t[3] = v; return 4; }
size_t tag_i(uint8_t* t, char a, char b, int32_t v) { t[0] = a; t[1] = b; t[2] = 'i'; put_u32(t + 3, (uint32_t)v); return 7; }
size_t tag_Z(uint8_t* t, char a, char b, const char* s) {
  t[0] = a; t[1] = b; t[2] = 'Z';
  size_t n = strlen(s) + 1;
  memcpy(t + 3, s, n);
  return 3 + n;
}

int pick_ref(Rng& r) {
  // weight by length: total ~3.1e9; cheap rejection sampling.
  for (;;) {
    int k = (int)r.below(N_REF);
    if (r.below(249250621u) < REF_LENS[k]) return k;
  }
}

// Upper bound of one record's bytes for each shape (scratch sizing).
size_t max_record_bytes(int shape) { return shape == LONG10K ? 40000 : 1024; }

// Generate record `idx` of `shape` at `out`. Returns bytes written.
size_t gen_record(int shape, uint64_t seed, uint64_t idx, uint8_t* out) {
  uint8_t tags[256];
  char name[256];
  size_t lt = 0;
  if (shape == SE150) {
    Rng r(mix(seed, idx, 1));
    Fixed f;
    f.ref_id = pick_ref(r);
    f.pos = (int32_t)r.below(REF_LENS[f.ref_id] - 150);
    f.flag = (r.u32() & 1) ? 16 : 0;
    f.mapq = r.below(10) ? 60 : (uint8_t)r.below(60);
    f.ref_span = 150;
    int n = snprintf(name, sizeof name, "HWI-ST1234:100:C0ABCACXX:%u:%u:%u:%u", 1 + (uint32_t)((idx >> 18) & 7),
                     1101 + (uint32_t)((idx >> 12) % 1578), 1000 + r.below(20000), 1000 + r.below(200000));
    uint32_t cig = (150u << 4) | 0;
    lt += tag_C(tags + lt, 'N', 'M', (uint8_t)r.below(4));
    lt += tag_C(tags + lt, 'A', 'S', (uint8_t)(130 + r.below(21)));
    char md[32];
    uint32_t k = r.below(8);
    if (k == 0) { uint32_t a = r.below(149); snprintf(md, sizeof md, "%u%c%u", a, "ACGT"[r.below(4)], 149 - a); }
    else snprintf(md, sizeof md, "150");
    lt += tag_Z(tags + lt, 'M', 'D', md);
    lt += tag_Z(tags + lt, 'R', 'G', "grp1");
    return write_record(out, f, name, (size_t)n + 1, &cig, 1, 150, r, 0, tags, lt);
  }
  if (shape == PE150) {
    uint64_t pair = idx >> 1;
    // The two mates of a pair are emitted in a seeded random order, so that the stable name sort (-n, input order
    // inside a name group) and the mate-matching sort (-n -M, flag order inside a group) give DIFFERENT outputs.
    int mate = (int)((idx & 1) ^ (mix(seed, pair, 8) & 1));
    Rng rp(mix(seed, pair, 2));       // shared by both mates
    Rng r(mix(seed, pair * 2 + (uint64_t)mate, 3));   // private to this mate
    Fixed f;
    bool unmapped = rp.below(50) == 0;
    int n = snprintf(name, sizeof name, "A00123:45:HXXXXDSXX:%u:%u:%u:%u", 1 + (uint32_t)((pair >> 20) & 3),
                     1101 + (uint32_t)((pair >> 13) % 1578), 1000 + rp.below(30000), 1000 + rp.below(36000));
    uint32_t cig = (150u << 4) | 0;
    uint32_t n_cig = 1;
    if (unmapped) {
      f.flag = mate ? 141 : 77;
      n_cig = 0;
    } else {
      int ref = pick_ref(rp);
      uint32_t insert = 200 + rp.below(300);
      int32_t p1 = (int32_t)rp.below(REF_LENS[ref] - 600);
      int32_t p2 = p1 + (int32_t)insert - 150;
      bool rev = rp.u32() & 1;
      // mate 0 is read1. !rev: read1 forward at p1 (99), read2 reverse at p2 (147).
      //                   rev: read1 reverse at p2 (83), read2 forward at p1 (163).
      static const uint16_t fl[2][2] = {{99, 147}, {83, 163}};
      f.flag = fl[rev][mate];
      bool me_left = (mate == 0) != rev;
      f.ref_id = ref;
      f.next_ref = ref;
      f.pos = me_left ? p1 : p2;
      f.next_pos = me_left ? p2 : p1;
      f.tlen = me_left ? (int32_t)insert : -(int32_t)insert;
      f.mapq = rp.below(12) ? 60 : (uint8_t)r.below(60);
      f.ref_span = 150;
      lt += tag_C(tags + lt, 'N', 'M', (uint8_t)r.below(4));
      char md[32];
      if (r.below(8) == 0) { uint32_t a = r.below(149); snprintf(md, sizeof md, "%u%c%u", a, "ACGT"[r.below(4)], 149 - a); }
      else snprintf(md, sizeof md, "150");
      lt += tag_Z(tags + lt, 'M', 'D', md);
      lt += tag_Z(tags + lt, 'M', 'C', "150M");
      lt += tag_C(tags + lt, 'A', 'S', (uint8_t)(130 + r.below(21)));
      lt += tag_C(tags + lt, 'M', 'Q', 60);
    }
    lt += tag_Z(tags + lt, 'R', 'G', "grp1");
    return write_record(out, f, name, (size_t)n + 1, &cig, n_cig, 150, r, 0, tags, lt);
  }
  if (shape == LONGNAME) {
    // groups of 4 consecutive records share a name: (hit 1 R1, hit 1 R2, hit 2 R1, hit 2 R2)
    // for multi-hit groups; single-hit groups use only HI:i:1 on all four slots' first two
    // and give the other two a distinct name. Every record carries HI (reference would
    // panic on mixed None/Some within a name group, comparators.rs:90).
    uint64_t grp = idx >> 2;
    // the four slots of a group appear in a seeded random order (one of the 24 permutations), so hits and mates of
    // one name are NOT pre-sorted by (HI, flag): -n keeps the input order inside a name, -n -M must reorder it
    static const uint8_t PERM4[24][4] = {{0,1,2,3},{0,1,3,2},{0,2,1,3},{0,2,3,1},{0,3,1,2},{0,3,2,1},{1,0,2,3},{1,0,3,2},
                                         {1,2,0,3},{1,2,3,0},{1,3,0,2},{1,3,2,0},{2,0,1,3},{2,0,3,1},{2,1,0,3},{2,1,3,0},
                                         {2,3,0,1},{2,3,1,0},{3,0,1,2},{3,0,2,1},{3,1,0,2},{3,1,2,0},{3,2,0,1},{3,2,1,0}};
    uint32_t slot = PERM4[mix(seed, grp, 9) % 24][idx & 3];
    Rng rg(mix(seed, grp, 4));
    Rng r(mix(seed, grp * 4 + slot, 5));
    bool multi = rg.below(3) == 0;
    uint32_t l_name = 200 + rg.below(23);  // 200..222 incl NUL
    uint64_t name_id = (multi || slot < 2) ? grp * 2 : grp * 2 + 1;
    Rng rn(mix(seed, name_id, 6));
    int n = snprintf(name, sizeof name, "m%08llx/", (unsigned long long)(rn.next() & 0xffffffffull));
    for (; n < (int)l_name - 1; ++n) name[n] = "ACGTNacgtn0123456789_:/-"[rn.below(24)];
    name[l_name - 1] = 0;
    Fixed f;
    f.ref_id = pick_ref(r);
    f.pos = (int32_t)r.below(REF_LENS[f.ref_id] - 100);
    bool read2 = slot & 1;
    bool rev = r.u32() & 1;
    f.flag = (uint16_t)(1 | 2 | (rev ? 16 : 32) | (read2 ? 128 : 64) | ((multi && slot >= 2) ? 256 : 0));
    f.next_ref = f.ref_id;
    f.next_pos = f.pos + 120;
    f.tlen = 170;
    f.mapq = (uint8_t)r.below(61);
    f.ref_span = 50;
    uint32_t cig = (50u << 4) | 0;
    lt += tag_C(tags + lt, 'N', 'M', (uint8_t)r.below(3));
    lt += tag_i(tags + lt, 'H', 'I', (multi && slot >= 2) ? 2 : 1);
    lt += tag_C(tags + lt, 'N', 'H', multi ? 2 : 1);
    return write_record(out, f, name, l_name, &cig, 1, 50, r, 0, tags, lt);
  }
  // LONG10K
  Rng r(mix(seed, idx, 7));
  Fixed f;
  f.ref_id = pick_ref(r);
  uint32_t l_seq = 9000 + r.below(2001);
  f.pos = (int32_t)r.below(REF_LENS[f.ref_id] - 20000);
  f.flag = (r.u32() & 1) ? 16 : 0;
  f.mapq = (uint8_t)r.below(61);
  uint32_t cig[512];
  uint32_t n_cig = 0, left = l_seq, span = 0;
  uint32_t clip = r.below(50);
  if (clip) { cig[n_cig++] = (clip << 4) | 4; left -= clip; }
  while (left > 0 && n_cig < 500) {
    uint32_t m = std::min(left, 20 + r.below(200));
    cig[n_cig++] = (m << 4) | 0; left -= m; span += m;
    if (left == 0) break;
    if (r.u32() & 1) { uint32_t d = 1 + r.below(3); cig[n_cig++] = (d << 4) | 2; span += d; }
    else { uint32_t i = std::min(left, 1 + r.below(3)); cig[n_cig++] = (i << 4) | 1; left -= i; }
  }
  if (left) { cig[n_cig++] = (left << 4) | 4; }
  f.ref_span = span;
  int n = snprintf(name, sizeof name, "%08x-%04x-%04x-%04x-%012llx", r.u32(), r.u32() & 0xffff, r.u32() & 0xffff,
                   r.u32() & 0xffff, (unsigned long long)(r.next() & 0xffffffffffffull));
  lt += tag_i(tags + lt, 'N', 'M', (int32_t)r.below(1500));
  lt += tag_Z(tags + lt, 'R', 'G', "ont1");
  return write_record(out, f, name, (size_t)n + 1, cig, n_cig, l_seq, r, 1, tags, lt);
}

std::vector<uint8_t> make_header(const std::string& extra_text) {
  std::string text = "@HD\tVN:1.6\tSO:unsorted\n";
  for (int i = 0; i < N_REF; ++i) text += std::string("@SQ\tSN:") + REF_NAMES[i] + "\tLN:" + std::to_string(REF_LENS[i]) + "\n";
  text += "@RG\tID:grp1\tSM:synthetic\n";
  text += extra_text;
  std::vector<uint8_t> h;
  auto u32 = [&](uint32_t v) { uint8_t b[4]; put_u32(b, v); h.insert(h.end(), b, b + 4); };
  h.insert(h.end(), {'B', 'A', 'M', 1});
  u32((uint32_t)text.size());
  h.insert(h.end(), text.begin(), text.end());
  u32(N_REF);
  for (int i = 0; i < N_REF; ++i) {
    size_t l = strlen(REF_NAMES[i]) + 1;
    u32((uint32_t)l);
    h.insert(h.end(), REF_NAMES[i], REF_NAMES[i] + l);
    u32(REF_LENS[i]);
  }
  return h;
}

const uint8_t BGZF_EOF[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43,
                              0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};

// Compress one BGZF block; appends to `dst`.
void bgzf_block(z_stream& zs, const uint8_t* src, size_t n, std::vector<uint8_t>& dst) {
  size_t base = dst.size();
  dst.resize(base + 18 + 65536 + 8);
  uint8_t* o = dst.data() + base;
  static const uint8_t hdr[16] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0};
  memcpy(o, hdr, 16);
  deflateReset(&zs);
  zs.next_in = const_cast<uint8_t*>(src);
  zs.avail_in = (uInt)n;
  zs.next_out = o + 18;
  zs.avail_out = 65536;
  int rc = deflate(&zs, Z_FINISH);
  if (rc != Z_STREAM_END) { fprintf(stderr, "bamgen: deflate failed rc=%d\n", rc); abort(); }
  size_t clen = 65536 - zs.avail_out;
  size_t total = 18 + clen + 8;
  if (total > 65535) { fprintf(stderr, "bamgen: block too large (%zu)\n", total); abort(); }  // never emit BSIZE=0xFFFF (util.rs:65 wrap)
  put_u16(o + 16, (uint16_t)(total - 1));
  put_u32(o + 18 + clen, (uint32_t)crc32(crc32(0, nullptr, 0), src, (uInt)n));
  put_u32(o + 18 + clen + 4, (uint32_t)n);
  dst.resize(base + total);
}

}  // namespace

extern "C" {

struct bamgen_params {
  int32_t shape;        // 0 se150, 1 pe150, 2 longname, 3 long10k
  int32_t level;        // zlib level 0..9
  uint64_t n_records;
  uint64_t seed;
  uint32_t payload;     // uncompressed bytes per BGZF block, 1..0xff00
  int32_t threads;      // 0 = hardware_concurrency
  int32_t no_eof;       // 1 = omit trailing EOF marker block
  int32_t header_block; // 1 = the BAM header gets a BGZF block of its own (records start at a block boundary)
};

// Generates the whole BAM into a malloc'ed buffer. Returns 0 on success.
// Also reports total uncompressed stream length (header + records) and the
// header length (bytes before the first record, incl. magic).
int bamgen_generate(const bamgen_params* p, uint8_t** out, uint64_t* out_len, uint64_t* ustream_len,
                    uint64_t* header_len) {
  if (!p || !out || !out_len || p->shape < 0 || p->shape > 3 || p->payload == 0 || p->payload > 0xff00 ||
      p->level < 0 || p->level > 9)
    return -1;
  const int shape = p->shape;
  const uint64_t N = p->n_records, seed = p->seed;
  const uint64_t P = p->payload;
  int T = p->threads > 0 ? p->threads : (int)std::thread::hardware_concurrency();
  if (T < 1) T = 1;
  std::vector<uint8_t> header = make_header("");
  const uint64_t H = header.size();

  // Pass 1: bytes per group of G records.
  const uint64_t G = 4096;
  const uint64_t n_groups = (N + G - 1) / G;
  std::vector<uint64_t> goff(n_groups + 1, 0);
  {
    std::atomic<uint64_t> next{0};
    auto work = [&]() {
      std::vector<uint8_t> scratch(max_record_bytes(shape));
      for (;;) {
        uint64_t g = next.fetch_add(1);
        if (g >= n_groups) break;
        uint64_t s = 0, e = std::min(N, (g + 1) * G);
        for (uint64_t i = g * G; i < e; ++i) s += gen_record(shape, seed, i, scratch.data());
        goff[g + 1] = s;
      }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back(work);
    for (auto& x : th) x.join();
  }
  goff[0] = H;
  for (uint64_t g = 0; g < n_groups; ++g) goff[g + 1] += goff[g];
  const uint64_t U = goff[n_groups];  // total uncompressed stream bytes
  // Block b covers stream bytes [bstart(b), bstart(b+1)).  With header_block the header is block 0 on its own, so the
  // record blocks that follow can be repeated verbatim to build a larger valid BAM (bench.py, N > 1 GPUs).
  const bool hb = p->header_block != 0;
  const uint64_t n_blocks = hb ? 1 + (U - H + P - 1) / P : (U + P - 1) / P;
  auto bstart = [&](uint64_t b) -> uint64_t {
    uint64_t s = hb ? (b == 0 ? 0 : H + (b - 1) * P) : b * P;
    return std::min(U, s);
  };

  // Pass 2: threads take contiguous runs of blocks.
  const uint64_t RUN = 64;
  const uint64_t n_runs = (n_blocks + RUN - 1) / RUN;
  std::vector<std::vector<uint8_t>> parts(n_runs);
  {
    std::atomic<uint64_t> next{0};
    auto work = [&]() {
      z_stream zs;
      memset(&zs, 0, sizeof zs);
      if (deflateInit2(&zs, p->level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) abort();
      std::vector<uint8_t> ubuf;
      for (;;) {
        uint64_t run = next.fetch_add(1);
        if (run >= n_runs) break;
        uint64_t b0 = run * RUN, b1 = std::min(n_blocks, b0 + RUN);
        uint64_t s0 = bstart(b0), s1 = bstart(b1);  // stream byte range
        // Materialise stream bytes [s0, s1) into ubuf.
        ubuf.clear();
        ubuf.reserve((size_t)(s1 - s0) + max_record_bytes(shape));
        uint64_t pos = s0;
        if (pos < H) {
          uint64_t e = std::min<uint64_t>(H, s1);
          ubuf.insert(ubuf.end(), header.begin() + pos, header.begin() + e);
          pos = e;
        }
        if (pos < s1) {
          // find group containing pos
          uint64_t g = (uint64_t)(std::upper_bound(goff.begin(), goff.end(), pos) - goff.begin()) - 1;
          uint64_t rec_off = goff[g];
          uint64_t i = g * G;
          std::vector<uint8_t> scratch(max_record_bytes(shape));
          while (pos < s1 && i < N) {
            size_t n = gen_record(shape, seed, i, scratch.data());
            uint64_t r0 = rec_off, r1 = rec_off + n;
            if (r1 > pos) {
              uint64_t a = std::max(r0, pos), bnd = std::min(r1, s1);
              ubuf.insert(ubuf.end(), scratch.begin() + (a - r0), scratch.begin() + (bnd - r0));
              pos = bnd;
            }
            rec_off = r1;
            ++i;
          }
        }
        std::vector<uint8_t>& dst = parts[run];
        dst.reserve((size_t)((s1 - s0) / 2));
        for (uint64_t b = b0; b < b1; ++b) {
          uint64_t a = bstart(b) - s0, e = bstart(b + 1) - s0;
          bgzf_block(zs, ubuf.data() + a, (size_t)(e - a), dst);
        }
      }
      deflateEnd(&zs);
    };
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back(work);
    for (auto& x : th) x.join();
  }
  uint64_t total = p->no_eof ? 0 : sizeof BGZF_EOF;
  for (auto& v : parts) total += v.size();
  uint8_t* buf = (uint8_t*)malloc(total ? total : 1);
  if (!buf) return -2;
  uint64_t o = 0;
  for (auto& v : parts) { memcpy(buf + o, v.data(), v.size()); o += v.size(); std::vector<uint8_t>().swap(v); }
  if (!p->no_eof) { memcpy(buf + o, BGZF_EOF, sizeof BGZF_EOF); o += sizeof BGZF_EOF; }
  *out = buf;
  *out_len = o;
  if (ustream_len) *ustream_len = U;
  if (header_len) *header_len = H;
  return 0;
}

void bamgen_free(uint8_t* p) { free(p); }

}  // extern "C"

#ifdef BAMGEN_MAIN
int main(int argc, char** argv) {
  bamgen_params p{};
  p.shape = PE150; p.level = 6; p.n_records = 100000; p.seed = 0xBA5EBA11ull; p.payload = 0xff00; p.threads = 0;
  const char* path = nullptr;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto val = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
    if (a == "--shape") { std::string s = val(); p.shape = s == "se150" ? 0 : s == "pe150" ? 1 : s == "longname" ? 2 : s == "long10k" ? 3 : -1; }
    else if (a == "--records") p.n_records = strtoull(val(), nullptr, 0);
    else if (a == "--level") p.level = atoi(val());
    else if (a == "--seed") p.seed = strtoull(val(), nullptr, 0);
    else if (a == "--payload") p.payload = (uint32_t)strtoul(val(), nullptr, 0);
    else if (a == "--threads") p.threads = atoi(val());
    else if (a == "--no-eof") p.no_eof = 1;
    else if (a == "--header-block") p.header_block = 1;
    else if (a == "--out") path = val();
    else { fprintf(stderr, "usage: bamgen --shape se150|pe150|longname|long10k --records N --level L --payload B --seed S --threads T --out FILE\n"); return 2; }
  }
  if (!path) { fprintf(stderr, "bamgen: --out required\n"); return 2; }
  uint8_t* buf; uint64_t len, U, H;
  int rc = bamgen_generate(&p, &buf, &len, &U, &H);
  if (rc) { fprintf(stderr, "bamgen: failed rc=%d\n", rc); return 1; }
  FILE* f = fopen(path, "wb");
  if (!f || fwrite(buf, 1, len, f) != len) { perror("bamgen: write"); return 1; }
  fclose(f);
  fprintf(stderr, "bamgen: %llu records, %llu bytes uncompressed, %llu bytes BAM\n", (unsigned long long)p.n_records,
          (unsigned long long)U, (unsigned long long)len);
  bamgen_free(buf);
  return 0;
}
#endif
