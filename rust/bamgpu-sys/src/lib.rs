//! Raw bindings of include/bamgpu.h plus a `bam_tools::Reader`-shaped wrapper.
//! UNBUILT in this repository's image (no cargo/rustc there): source only, see INTEGRATION.md.
#![allow(non_camel_case_types)]
use std::io::{self, Read, Write};
use std::os::raw::{c_char, c_int, c_long, c_void};

#[repr(C)] pub struct bamgpu_reader { _p: [u8; 0] }
pub const BAMGPU_MAX_DEVICES: usize = 16;
/// include/bamgpu.h `bamgpu_opts`, field for field (the layout is part of the ABI).
#[repr(C)] #[derive(Clone, Copy)]
pub struct bamgpu_opts {
    pub device: i32, pub flags: u32, pub batch_bytes: u64, pub inflate_impl: i32, pub jobs_per_device: i32,
    pub n_devices: i32, pub devices: [i32; BAMGPU_MAX_DEVICES], pub sort_output: i32, pub sort_mem_limit: u64,
    pub out_compr_level: i32, pub reserved: i32,
}
impl Default for bamgpu_opts {
    fn default() -> Self {
        bamgpu_opts { device: -1, flags: 0, batch_bytes: 0, inflate_impl: 0, jobs_per_device: 0, n_devices: 0,
                      devices: [0; BAMGPU_MAX_DEVICES], sort_output: 0, sort_mem_limit: 0, out_compr_level: 0, reserved: 0 }
    }
}
impl bamgpu_opts {
    /// One Reader over several GPUs of the box: batches (= BGZF block ranges) go to them round robin.
    pub fn with_devices(devs: &[i32]) -> Self {
        let mut o = Self::default();
        o.n_devices = devs.len().min(BAMGPU_MAX_DEVICES) as i32;
        for (i, d) in devs.iter().take(BAMGPU_MAX_DEVICES).enumerate() { o.devices[i] = *d; }
        o
    }
}
#[repr(C)] pub struct bamgpu_engine { _p: [u8; 0] }
/// include/bamgpu.h `bamgpu_field` / `bamgpu_fields` (bamgpu_engine_decode_fields: decode_cigar / decode_seq / RawQual of a batch).
#[repr(C)] #[derive(Clone, Copy)]
pub struct bamgpu_field { pub total: u64, pub bad: u64, pub off: *const u64, pub bytes: *const u8, pub h_off: *const u64, pub h_bytes: *const u8 }
#[repr(C)] #[derive(Clone, Copy)]
pub struct bamgpu_fields { pub n_records: u64, pub cigar: bamgpu_field, pub seq: bamgpu_field, pub qual: bamgpu_field, pub ms_measure: f64, pub ms_decode: f64 }
pub const BAMGPU_FIELD_CIGAR: u32 = 1;
pub const BAMGPU_FIELD_SEQ: u32 = 2;
pub const BAMGPU_FIELD_QUAL: u32 = 4;
pub const BAMGPU_COPY_DATA: u32 = 1;
pub const BAMGPU_DEFER_CRC: u32 = 256;   // bamgpu_engine_inflate: the CRC32 verdict comes with the following bamgpu_engine_records call
pub type bamgpu_read_fn = unsafe extern "C" fn(ctx: *mut c_void, dst: *mut u8, cap: usize) -> c_long;
pub const BAMGPU_OK: c_int = 0;
pub const BAMGPU_EOF: c_int = 1;

extern "C" {
    pub fn bamgpu_reader_new(read: bamgpu_read_fn, ctx: *mut c_void, thread_num: usize, track_progress: *const u64,
                             opts: *const bamgpu_opts) -> *mut bamgpu_reader;
    pub fn bamgpu_reader_free(r: *mut bamgpu_reader);
    pub fn bamgpu_last_error(r: *const bamgpu_reader) -> *const c_char;
    pub fn bamgpu_read_header(r: *mut bamgpu_reader, bytes: *mut *const u8, len: *mut usize, off_to_refs: *mut usize) -> c_int;
    pub fn bamgpu_next_rec(r: *mut bamgpu_reader, body: *mut *const u8, len: *mut usize) -> c_int;
    pub fn bamgpu_append_record(r: *mut bamgpu_reader, dst: *mut u8, cap: usize, len: *mut usize) -> c_int;
    pub fn bamgpu_read(r: *mut bamgpu_reader, dst: *mut u8, cap: usize) -> c_long;
    pub fn bamgpu_reader_current_engine(r: *mut bamgpu_reader) -> *mut bamgpu_engine;
    pub fn bamgpu_engine_decode_fields(e: *mut bamgpu_engine, which: u32, flags: u32, cuda_stream: *mut c_void, out: *mut bamgpu_fields) -> c_int;
    pub fn bamgpu_sort_bam(read: bamgpu_read_fn, rctx: *mut c_void, write: bamgpu_write_fn, wctx: *mut c_void, reader_thread_num: usize,
                           sort_by: c_int, opts: *const bamgpu_opts, stats: *mut c_void, err: *mut c_char, err_cap: usize) -> c_int;
    pub fn bamgpu_sort_bam_from_memory(buf: *const u8, len: usize, write: bamgpu_write_fn, wctx: *mut c_void, reader_thread_num: usize,
                                       sort_by: c_int, opts: *const bamgpu_opts, stats: *mut c_void, err: *mut c_char, err_cap: usize) -> c_int;
}
pub type bamgpu_write_fn = unsafe extern "C" fn(ctx: *mut c_void, src: *const u8, len: usize) -> c_int;

/// Drop-in for `bam_tools::Reader` (bam_tools/src/reader.rs:17-99).
pub struct Reader { h: *mut bamgpu_reader, _inner: Box<Box<dyn Read + Send>>, rec: Vec<u8> }
unsafe impl Send for Reader {}

unsafe extern "C" fn trampoline(ctx: *mut c_void, dst: *mut u8, cap: usize) -> c_long {
    let inner = &mut *(ctx as *mut Box<dyn Read + Send>);
    match inner.read(std::slice::from_raw_parts_mut(dst, cap)) { Ok(n) => n as c_long, Err(_) => -1 }
}

fn err(h: *const bamgpu_reader) -> io::Error {
    let msg = unsafe { std::ffi::CStr::from_ptr(bamgpu_last_error(h)) }.to_string_lossy().into_owned();
    io::Error::new(io::ErrorKind::InvalidData, msg)
}

impl Reader {
    /// reader.rs:28-32 — `thread_num` is clamped like the reference; `track_progress` is kept as a byte counter.
    pub fn new<RSS: Read + Send + 'static>(inner: RSS, thread_num: usize, track_progress: Option<u64>) -> Self {
        let mut boxed: Box<Box<dyn Read + Send>> = Box::new(Box::new(inner));
        let tp = track_progress.as_ref().map_or(std::ptr::null(), |v| v as *const u64);
        let h = unsafe { bamgpu_reader_new(trampoline, &mut *boxed as *mut _ as *mut c_void, thread_num, tp, std::ptr::null()) };
        assert!(!h.is_null(), "bamgpu: no usable CUDA device (there is no CPU fallback)");
        Reader { h, _inner: boxed, rec: Vec::new() }
    }
    /// reader.rs:87-98
    pub fn read_header(&mut self) -> io::Result<(Vec<u8>, usize)> {
        let (mut p, mut n, mut off) = (std::ptr::null(), 0usize, 0usize);
        if unsafe { bamgpu_read_header(self.h, &mut p, &mut n, &mut off) } < 0 { return Err(err(self.h)); }
        Ok((unsafe { std::slice::from_raw_parts(p, n) }.to_vec(), off))
    }
    /// reader.rs:83
    pub fn records(&mut self) -> Records<'_> { Records { reader: self } }
    /// reader.rs:63-73 — returns Ok(0) at EOF; a short record body panics like the reference's `expect("Failed to read.")`.
    pub fn append_record(&mut self, buf: &mut Vec<u8>) -> io::Result<usize> {
        let (mut p, mut n) = (std::ptr::null(), 0usize);
        match unsafe { bamgpu_next_rec(self.h, &mut p, &mut n) } {
            BAMGPU_EOF => Ok(0),
            BAMGPU_OK => { buf.extend_from_slice(unsafe { std::slice::from_raw_parts(p, n) }); Ok(n) }
            -22 => panic!("Failed to read."),
            _ => Err(err(self.h)),
        }
    }
    /// reader.rs:57-61
    pub fn read_record(&mut self, buf: &mut Vec<u8>) -> io::Result<usize> { buf.clear(); self.append_record(buf) }
}
/// reader.rs:114-152
impl Read for Reader {
    fn read(&mut self, buf: &mut [u8]) -> io::Result<usize> {
        let n = unsafe { bamgpu_read(self.h, buf.as_mut_ptr(), buf.len()) };
        if n < 0 { Err(err(self.h)) } else { Ok(n as usize) }
    }
}
impl Drop for Reader { fn drop(&mut self) { unsafe { bamgpu_reader_free(self.h) } } }

/// reader/records.rs:10-42
pub struct Records<'a> { reader: &'a mut Reader }
impl<'a> Records<'a> {
    /// Lends the next record body (valid until the next call), `None` at EOF — records.rs:23-29.
    pub fn next_rec(&mut self) -> Option<io::Result<&Vec<u8>>> {
        let (mut p, mut n) = (std::ptr::null(), 0usize);
        match unsafe { bamgpu_next_rec(self.reader.h, &mut p, &mut n) } {
            BAMGPU_EOF => None,
            BAMGPU_OK => {
                self.reader.rec.clear();
                self.reader.rec.extend_from_slice(unsafe { std::slice::from_raw_parts(p, n) });
                Some(Ok(&self.reader.rec))
            }
            -22 => panic!("Failed to read."),
            _ => Some(Err(err(self.reader.h))),
        }
    }
}
impl<'a> Iterator for Records<'a> {
    type Item = io::Result<Vec<u8>>;
    fn next(&mut self) -> Option<Self::Item> { self.next_rec().map(|r| r.map(|v| v.clone())) }
}

/// sorting/sort.rs:92-96
#[derive(Clone, Copy, Debug, PartialEq)]
pub enum SortBy { Name = 0, NameAndMatchMates = 1, CoordinatesAndStrand = 2 }

unsafe extern "C" fn write_trampoline(ctx: *mut c_void, src: *const u8, len: usize) -> c_int {
    let sink = &mut *(ctx as *mut &mut dyn Write);
    match sink.write_all(std::slice::from_raw_parts(src, len)) { Ok(()) => 0, Err(_) => -1 }
}

/// sorting/sort.rs:98-103.  Accepted for signature parity; the GPU sort keeps every record in HBM and writes no temp files.
pub enum TempFilesMode { RegularFiles, LZ4CompressedFiles, InMemoryBlocks, InMemoryBlocksLZ4 }

/// Drop-in for `bam_tools::sorting::sort::sort_bam` (sort.rs:138-149): the SAME ten arguments in the same order, so
/// `bam_binary/src/main.rs:66-77` compiles against this crate unchanged.  `tmp_dir` is generic (the reference takes
/// `&tempdir::TempDir`; this crate has no dependencies).  mem_limit / tmp_dir / out_compr_level / temp_files_mode steer the
/// reference's external merge and have no counterpart here; `index_file_to_create` must be `None` (the GBAM index mode,
/// sort.rs:106-134, is unreachable from bam_binary: main.rs:74 passes `Option::<File>::None`).  `bam_file_size` becomes
/// the reader's `track_progress`.  Panics where the reference's comparators panic.  A file that does not fit in HBM is
/// sorted in the streamed form automatically (keys on the device, bodies in host memory); `sort_bam_with` exposes
/// `bamgpu_opts` (device list, output format, `sort_mem_limit` to force the streamed form).
#[allow(clippy::too_many_arguments)]
pub fn sort_bam<R: Read + Send + 'static, W: Write, IndexW: Write, TmpDir>(
    _mem_limit: usize,
    reader: R,
    sorted_sink: &mut W,
    _tmp_dir: &TmpDir,
    _out_compr_level: usize,
    reader_thread_num: usize,
    _temp_files_mode: TempFilesMode,
    index_file_to_create: Option<IndexW>,
    sort_by: SortBy,
    _bam_file_size: Option<u64>,
) -> io::Result<()> {
    if index_file_to_create.is_some() {
        return Err(io::Error::new(io::ErrorKind::Unsupported, "GBAM index sort (sort.rs:106-134) is not built"));
    }
    sort_bam_with(reader, sorted_sink, reader_thread_num, sort_by, &bamgpu_opts::default())
}

/// `sort_bam` with explicit options (bamgpu_opts.sort_output: 0 bodies = the reference's output, 1 (u32 block_size, body)*,
/// 2 a complete BAM in BGZF - stored blocks, or with bamgpu_opts.out_compr_level >= 1 compressed on the GPU;
/// bamgpu_opts.sort_mem_limit: device bytes for record bodies, 0 = whatever fits; bamgpu_opts.devices (`with_devices`): the
/// sample sort over several GPUs, same output bytes).
pub fn sort_bam_with<R: Read + Send + 'static, W: Write>(reader: R, sorted_sink: &mut W, reader_thread_num: usize, sort_by: SortBy,
                                                       opts: &bamgpu_opts) -> io::Result<()> {
    let mut boxed: Box<Box<dyn Read + Send>> = Box::new(Box::new(reader));
    let mut sink: &mut dyn Write = sorted_sink;
    let mut msg = [0 as c_char; 512];
    let rc = unsafe {
        bamgpu_sort_bam(trampoline, &mut *boxed as *mut _ as *mut c_void, write_trampoline, &mut sink as *mut _ as *mut c_void,
                        reader_thread_num, sort_by as c_int, opts as *const bamgpu_opts, std::ptr::null_mut(), msg.as_mut_ptr(), msg.len())
    };
    let text = unsafe { std::ffi::CStr::from_ptr(msg.as_ptr()) }.to_string_lossy().into_owned();
    match rc {
        0 => Ok(()),
        -30 | -31 | -32 => panic!("{}", text),   // comparators.rs:90, :178, tags.rs: the reference panics here
        _ => Err(io::Error::new(io::ErrorKind::InvalidData, text)),
    }
}
