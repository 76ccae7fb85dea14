//! Rust host side of the B200 path: a thin FFI over `include/seq2kminmers.h` plus the batched variant of
//! `KminmersIterator` (reference `src/lib.rs:89-131,179-270`) that yields the same items in the same order.
//!
//! Not compiled in the B200 repository (its image has no Rust toolchain); every signature below restates the C header
//! one to one.  Inside rust-seq2kminmers this file would live as `src/gpu.rs` and use the crate's own `HashMode` and
//! `KminmerHash` (`src/lib.rs:21-27`, `src/kminmer.rs:128-177`) instead of the stand-alone copies kept here.

use std::ffi::{CStr, CString};
use std::os::raw::{c_char, c_int, c_void};

/// `HashMode`, same discriminants as `src/lib.rs:21-27` and `s2k_hash_mode`.
#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum HashMode { Regular = 0, Hpc = 1, Simd = 2, HpcSimd = 3 }

/// Which rolling hash: ntHash1-32 (what the crate compiles) or the 31-bit hybrid of `src/nthash2_avx512_32.rs`.
#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum HashVariant { Nt1_32 = 0, Nt2_31 = 1, Nt1_64 = 2, Nt1_16 = 3 }

/// `KminmerHash` (`src/kminmer.rs:128-135`); equality and order by `hash` only (`src/kminmer.rs:181-203`).
#[derive(Clone, Copy, Debug)]
pub struct KminmerHash { pub hash: u64, pub start: usize, pub end: usize, pub offset: usize, pub rev: bool }
impl PartialEq for KminmerHash { fn eq(&self, o: &Self) -> bool { self.hash == o.hash } }
impl Eq for KminmerHash {}
impl KminmerHash {
    /// `KminmerHash::new_from_hash` (`src/kminmer.rs:169-177`).
    pub fn new_from_hash(hash: u64, start: usize, end: usize, offset: usize, rev: bool) -> Self {
        KminmerHash { hash, start, end, offset, rev }
    }
}

#[repr(C)]
pub struct S2kParams { pub l: u32, pub k: u32, pub density: f64, pub mode: i32, pub variant: i32 }
/// `s2k_count_result`: (hash, count, id of the first item) of the distinct k-min-mer hashes, device pointers.
#[repr(C)]
pub struct S2kCountResult {
    pub n_distinct: u64, pub n_items: u64, pub hash: *const u64, pub count: *const u32, pub first: *const u64,
    pub location: i32, pub reserved: i32,
}

/// Item of `NtHashHPCIterator` / `NtHashSIMDIterator` / `NtHashHPCSIMDIterator` plus the sequence index.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct S2kMinimizer { pub hash: u32, pub start: u32, pub end: u32, pub seq: u32 }

#[repr(C)]
pub struct S2kResult {
    pub n_seqs: u64, pub n_items: u64, pub n_minimizers: u64,
    pub hash: *const u64, pub start: *const u32, pub end: *const u32, pub rev: *const u8,
    pub km_off: *const u64, pub minimizers: *const S2kMinimizer, pub min_off: *const u64, pub min_cnt: *const u32,
    pub location: i32, pub reserved: i32,
}

#[repr(C)]
pub struct S2kRleResult {
    pub n_seqs: u64, pub n_hpc: u64,
    pub hpc: *const u8, pub pos: *const u32, pub hpc_off: *const u64,
    pub location: i32, pub reserved: i32,
}

#[repr(C)]
pub struct S2kCtx { _private: [u8; 0] }

pub const S2K_WANT_MINIMIZERS: u32 = 1;
pub const S2K_NO_TAIL_RULE: u32 = 2;

extern "C" {
    pub fn s2k_abi_version() -> c_int;
    pub fn s2k_ctx_create(device: c_int, out: *mut *mut S2kCtx) -> c_int;
    pub fn s2k_ctx_destroy(ctx: *mut S2kCtx);
    pub fn s2k_ctx_set_flags(ctx: *mut S2kCtx, flags: u32) -> c_int;
    pub fn s2k_ctx_set_slab_bytes(ctx: *mut S2kCtx, bytes: u64) -> c_int;
    pub fn s2k_ctx_set_transport(ctx: *mut S2kCtx, host_threads: c_int, pack_ratio: f64) -> c_int;
    pub fn s2k_run(ctx: *mut S2kCtx, bases: *const u8, seq_off: *const u64, n_seqs: u64,
                   params: *const S2kParams, out: *mut S2kResult) -> c_int;
    pub fn s2k_run_packed2(ctx: *mut S2kCtx, packed: *const u8, seq_off: *const u64, n_seqs: u64,
                           params: *const S2kParams, out: *mut S2kResult) -> c_int;
    pub fn s2k_pack2(bases: *const u8, n_bases: u64, packed_out: *mut u8, host_threads: c_int) -> i64;
    pub fn s2k_run_fastx(ctx: *mut S2kCtx, path: *const c_char, nb_threads: c_int,
                         params: *const S2kParams, out: *mut S2kResult) -> c_int;
    pub fn s2k_last_fastx(ctx: *const S2kCtx, n_seqs: *mut u64, n_bases: *mut u64,
                          bases: *mut *const u8, seq_off: *mut *const u64) -> c_int;
    pub fn s2k_run_device(ctx: *mut S2kCtx, d_bases: *const u8, d_seq_off: *const u64, n_seqs: u64, n_bases: u64,
                          params: *const S2kParams, stream: *mut c_void, out: *mut S2kResult) -> c_int;
    pub fn s2k_encode_rle(ctx: *mut S2kCtx, bases: *const u8, seq_off: *const u64, n_seqs: u64,
                          out: *mut S2kRleResult) -> c_int;
    pub fn s2k_bounds(density: f64, bound_scalar: *mut u32, bound_simd: *mut u32, bound_31: *mut u32);
    pub fn s2k_bound_u64(density: f64) -> u64;
    pub fn s2k_bound_u16(density: f64) -> u32;
    pub fn s2k_last_minimizer_hash_hi(ctx: *const S2kCtx, d_hi: *mut *const u32) -> c_int;
    pub fn s2k_count_device(ctx: *mut S2kCtx, d_hash: *const u64, d_id: *const u64, n_items: u64, id_base: u64,
                            stream: *mut c_void, out: *mut S2kCountResult) -> c_int;
    pub fn s2k_count_partition_device(ctx: *mut S2kCtx, d_hash: *const u64, n_items: u64, id_base: u64, n_parts: u32,
                                      part_counts: *mut u64, d_out_hash: *mut u64, d_out_id: *mut u64, stream: *mut c_void) -> c_int;
    pub fn s2k_count_part(hash: u64, n_parts: u32) -> u32;
    pub fn s2k_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn s2k_host_free(p: *mut c_void);
    pub fn s2k_last_error(ctx: *const S2kCtx) -> *const c_char;
    pub fn s2k_strerror(status: c_int) -> *const c_char;
}

fn io_err(status: c_int, detail: *const c_char) -> std::io::Error {
    let what = unsafe { CStr::from_ptr(s2k_strerror(status)) }.to_string_lossy().into_owned();
    let more = if detail.is_null() { String::new() } else { unsafe { CStr::from_ptr(detail) }.to_string_lossy().into_owned() };
    std::io::Error::new(std::io::ErrorKind::Other, format!("s2k status {} ({}) {}", status, what, more))
}

/// One CUDA device + stream + grow-only buffers.  Single-threaded like one `KminmersIterator`; use one per worker
/// thread (`src/main.rs:65-79` creates one iterator per record inside the worker closure).
pub struct GpuContext(*mut S2kCtx);
unsafe impl Send for GpuContext {}

impl GpuContext {
    pub fn new(device: i32) -> std::io::Result<Self> {
        let mut p = std::ptr::null_mut();
        match unsafe { s2k_ctx_create(device, &mut p) } { 0 => Ok(GpuContext(p)), e => Err(io_err(e, std::ptr::null())) }
    }
    fn check(&self, st: c_int) -> std::io::Result<()> {
        if st == 0 { Ok(()) } else { Err(io_err(st, unsafe { s2k_last_error(self.0) })) }
    }
    /// Share of host slabs packed to 2 bits/base before crossing PCIe, and the threads doing it.
    pub fn set_transport(&mut self, host_threads: i32, pack_ratio: f64) -> std::io::Result<()> {
        self.check(unsafe { s2k_ctx_set_transport(self.0, host_threads, pack_ratio) })
    }
    pub fn set_flags(&mut self, flags: u32) -> std::io::Result<()> { self.check(unsafe { s2k_ctx_set_flags(self.0, flags) }) }

    /// `encode_rle_simd` over a batch (`src/hpc.rs:44-147`): (kept bytes, run starts, per-sequence offsets).
    /// Takes `&mut self`: a run reallocates the context's result buffers, so no `KminmersBatchIterator` (which points into
    /// them) may be alive -- the exclusive borrow lets the compiler enforce that.
    pub fn encode_rle(&mut self, bases: &[u8], seq_off: &[u64]) -> std::io::Result<(Vec<u8>, Vec<u32>, Vec<u64>)> {
        assert!(!seq_off.is_empty() && seq_off[0] == 0 && *seq_off.last().unwrap() as usize <= bases.len());
        let mut r: S2kRleResult = unsafe { std::mem::zeroed() };
        self.check(unsafe { s2k_encode_rle(self.0, bases.as_ptr(), seq_off.as_ptr(), (seq_off.len() - 1) as u64, &mut r) })?;
        let n = r.n_hpc as usize;
        unsafe {
            Ok((std::slice::from_raw_parts(r.hpc, n).to_vec(), std::slice::from_raw_parts(r.pos, n).to_vec(),
                std::slice::from_raw_parts(r.hpc_off, r.n_seqs as usize + 1).to_vec()))
        }
    }
}
impl Drop for GpuContext { fn drop(&mut self) { unsafe { s2k_ctx_destroy(self.0) } } }

/// `hash_bound` of `src/lib.rs:91` and its re-derivations (`src/nthash_avx512_32.rs:47-48`, `src/nthash2_avx512_32.rs:52-54`).
pub fn bounds(density: f64) -> (u32, u32, u32) {
    let (mut a, mut b, mut c) = (0u32, 0u32, 0u32);
    unsafe { s2k_bounds(density, &mut a, &mut b, &mut c) };
    (a, b, c)
}

/// Batched variant of `KminmersIterator`: one call for many reads; yields `(read index, KminmerHash)` in the order
/// `for read in reads { for kminmer in KminmersIterator::new(read, l, k, density, mode) { .. } }` would.
/// The result buffers belong to the context and are reallocated by its next run: the iterator holds the context's
/// EXCLUSIVE borrow, so safe code cannot start another run (or `encode_rle`) while it is alive.
pub struct KminmersBatchIterator<'c> { _ctx: &'c mut GpuContext, res: S2kResult, read: usize, i: u64 }

impl<'c> KminmersBatchIterator<'c> {
    /// `bases`: concatenated reads (ASCII); `seq_off`: n+1 offsets starting at 0.  Mirrors `KminmersIterator::new`
    /// (`src/lib.rs:89`); where the reference panics (`l > 31` in the SIMD modes, `l >= 256`) this returns `Err`.
    pub fn new(ctx: &'c mut GpuContext, bases: &[u8], seq_off: &[u64], l: usize, k: usize, density: f64, mode: HashMode)
        -> std::io::Result<Self> {
        Self::with_variant(ctx, bases, seq_off, l, k, density, mode, HashVariant::Nt1_32)
    }
    pub fn with_variant(ctx: &'c mut GpuContext, bases: &[u8], seq_off: &[u64], l: usize, k: usize, density: f64,
                        mode: HashMode, variant: HashVariant) -> std::io::Result<Self> {
        assert!(!seq_off.is_empty() && *seq_off.last().unwrap() as usize <= bases.len());
        let p = S2kParams { l: l as u32, k: k as u32, density, mode: mode as i32, variant: variant as i32 };
        let mut res: S2kResult = unsafe { std::mem::zeroed() };
        ctx.check(unsafe { s2k_run(ctx.0, bases.as_ptr(), seq_off.as_ptr(), (seq_off.len() - 1) as u64, &p, &mut res) })?;
        Ok(Self { _ctx: ctx, res, read: 0, i: 0 })
    }
    /// Reads already packed four bases per byte (A=0 C=1 T=2 G=3, `s2k_pack2`): a quarter of the PCIe bytes.
    pub fn from_packed2(ctx: &'c mut GpuContext, packed: &[u8], seq_off: &[u64], l: usize, k: usize, density: f64,
                        mode: HashMode) -> std::io::Result<Self> {
        assert!(!seq_off.is_empty() && seq_off[0] == 0 && (*seq_off.last().unwrap() as usize + 3) / 4 <= packed.len());
        let p = S2kParams { l: l as u32, k: k as u32, density, mode: mode as i32, variant: 0 };
        let mut res: S2kResult = unsafe { std::mem::zeroed() };
        ctx.check(unsafe { s2k_run_packed2(ctx.0, packed.as_ptr(), seq_off.as_ptr(), (seq_off.len() - 1) as u64, &p, &mut res) })?;
        Ok(Self { _ctx: ctx, res, read: 0, i: 0 })
    }
    /// The file mode of `src/main.rs:50-81`: `parallel_fastx(&filename, nb_threads, task)` and the iterator per record.
    pub fn from_fastx(ctx: &'c mut GpuContext, path: &str, nb_threads: usize, l: usize, k: usize, density: f64,
                      mode: HashMode) -> std::io::Result<Self> {
        let c = CString::new(path).map_err(|e| std::io::Error::new(std::io::ErrorKind::InvalidInput, e))?;
        let p = S2kParams { l: l as u32, k: k as u32, density, mode: mode as i32, variant: 0 };
        let mut res: S2kResult = unsafe { std::mem::zeroed() };
        ctx.check(unsafe { s2k_run_fastx(ctx.0, c.as_ptr(), nb_threads as c_int, &p, &mut res) })?;
        Ok(Self { _ctx: ctx, res, read: 0, i: 0 })
    }
    pub fn n_items(&self) -> u64 { self.res.n_items }
    pub fn n_seqs(&self) -> u64 { self.res.n_seqs }
    /// Items of read `r` as SoA slices (hash, start, end, rev) without going through the iterator.
    pub fn read_items(&self, r: usize) -> (&[u64], &[u32], &[u32], &[u8]) {
        let km_off = unsafe { std::slice::from_raw_parts(self.res.km_off, self.res.n_seqs as usize + 1) };
        let (a, n) = (km_off[r] as usize, (km_off[r + 1] - km_off[r]) as usize);
        unsafe {
            (std::slice::from_raw_parts(self.res.hash.add(a), n), std::slice::from_raw_parts(self.res.start.add(a), n),
             std::slice::from_raw_parts(self.res.end.add(a), n), std::slice::from_raw_parts(self.res.rev.add(a), n))
        }
    }
}

impl<'c> Iterator for KminmersBatchIterator<'c> {
    type Item = (usize, KminmerHash);
    fn next(&mut self) -> Option<Self::Item> {
        if self.i >= self.res.n_items { return None; }
        let km_off = unsafe { std::slice::from_raw_parts(self.res.km_off, self.res.n_seqs as usize + 1) };
        while km_off[self.read + 1] <= self.i { self.read += 1; }      // reads without items are skipped
        let i = self.i as usize;
        self.i += 1;
        unsafe {
            Some((self.read, KminmerHash::new_from_hash(*self.res.hash.add(i), *self.res.start.add(i) as usize,
                 *self.res.end.add(i) as usize, i - km_off[self.read] as usize, *self.res.rev.add(i) != 0)))
        }
    }
}

/// Pinned staging buffer (`s2k_host_alloc`): lets `s2k_run` copy at the full PCIe rate.
pub struct PinnedBuf { p: *mut u8, len: usize }
unsafe impl Send for PinnedBuf {}
impl PinnedBuf {
    pub fn new(len: usize) -> std::io::Result<Self> {
        let mut p: *mut c_void = std::ptr::null_mut();
        match unsafe { s2k_host_alloc(len, &mut p) } { 0 => Ok(PinnedBuf { p: p as *mut u8, len }), e => Err(io_err(e, std::ptr::null())) }
    }
    pub fn as_mut_slice(&mut self) -> &mut [u8] { unsafe { std::slice::from_raw_parts_mut(self.p, self.len) } }
    pub fn as_slice(&self) -> &[u8] { unsafe { std::slice::from_raw_parts(self.p, self.len) } }
}
impl Drop for PinnedBuf { fn drop(&mut self) { unsafe { s2k_host_free(self.p as *mut c_void) } } }

#[cfg(test)]
mod tests {
    //! Twin of the reference's `tests/main.rs:41-57` (KAT-1): needs a B200 and libs2k_b200.so at run time.
    use super::*;
    #[test]
    fn bounds_follow_the_reference_recipe() {
        assert_eq!(bounds(0.01), (42949672, 42949672, 21474836));
        assert_eq!(bounds(0.007), (30064771, 30064772, 15032386));
    }
    #[test]
    fn short_sequences_yield_nothing() {
        let ctx = GpuContext::new(0).unwrap();
        let it = KminmersBatchIterator::new(&ctx, b"ACGTACGT", &[0, 8], 31, 5, 0.01, HashMode::HpcSimd).unwrap();
        assert_eq!(it.count(), 0);                              // src/lib.rs:97
    }
}
