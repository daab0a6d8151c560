//! rapidfuzz-b200 -- Rust host shim over the C ABI of `include/rfgpu.h` (librfgpu.so).
//!
//! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no cargo/rustc.  It is the
//! binding a maintainer of rapidfuzz-rs would add to route the one-vs-many `BatchComparator` hot path
//! to the B200 engine; it is kept mechanical on purpose.  Reference signatures it mirrors:
//! `distance::levenshtein::{Args, BatchComparator}` (src/distance/levenshtein.rs:86-126, :1636-1818) and the
//! same shape for indel / lcs_seq / osa / jaro / jaro_winkler and `fuzz::RatioBatchComparator`
//! (src/fuzz.rs:98-150).  `None` (score worse than `score_cutoff`, src/common.rs:43-45, :83-85) crosses the
//! boundary as `u32::MAX` / NaN.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rf_args {
    pub has_cutoff: u8,
    pub cutoff_u: u64,
    pub cutoff_f: f64,
    pub has_hint: u8,
    pub hint_u: u64,
    pub hint_f: f64,
    pub insertion_cost: u64,
    pub deletion_cost: u64,
    pub substitution_cost: u64,
    pub prefix_weight: f64,
    pub reference_quirks: u8,
    pub pad: u8, // hamming::Args::pad
}
#[repr(C)] pub struct rf_corpus { _p: [u8; 0] }
#[repr(C)] pub struct rf_batch { _p: [u8; 0] }
#[repr(C)] pub struct rf_sharded_corpus { _p: [u8; 0] }
#[repr(C)] pub struct rf_sharded_batch { _p: [u8; 0] }
#[repr(C)] pub struct rf_comm { _p: [u8; 0] }

/// `rf_elem_type`: the reference's `HashableChar` element types (src/details/common.rs:29-37)
pub const RF_ELEM_U8: c_int = 0;
pub const RF_ELEM_U16: c_int = 1;
pub const RF_ELEM_U32: c_int = 2;
pub const RF_ELEM_U64: c_int = 3;
pub const RF_ELEM_I8: c_int = 4;
pub const RF_ELEM_I16: c_int = 5;
pub const RF_ELEM_I32: c_int = 6;
pub const RF_ELEM_I64: c_int = 7;

pub const RF_LEVENSHTEIN: c_int = 0;
pub const RF_INDEL: c_int = 1;
pub const RF_LCS_SEQ: c_int = 2;
pub const RF_OSA: c_int = 3;
pub const RF_JARO: c_int = 4;
pub const RF_JARO_WINKLER: c_int = 5;
pub const RF_RATIO: c_int = 6;
pub const RF_HAMMING: c_int = 7;
pub const RF_PREFIX: c_int = 8;
pub const RF_POSTFIX: c_int = 9;
pub const RF_DAMERAU_LEVENSHTEIN: c_int = 10;
pub const RF_DISTANCE: c_int = 0;
pub const RF_SIMILARITY: c_int = 1;
pub const RF_NORMALIZED_DISTANCE: c_int = 2;
pub const RF_NORMALIZED_SIMILARITY: c_int = 3;

#[link(name = "rfgpu")]
extern "C" {
    pub fn rf_args_default(a: *mut rf_args);
    pub fn rf_last_error() -> *const c_char;
    pub fn rf_corpus_create_u8(chars: *const u8, offsets: *const u64, n: u64, device: c_int, out: *mut *mut rf_corpus) -> c_int;
    // `char` / u32 elements: per-query alphabet renaming on the device, exact (see include/rfgpu.h)
    pub fn rf_corpus_create_u32(elems: *const u32, offsets: *const u64, n: u64, device: c_int, out: *mut *mut rf_corpus) -> c_int;
    pub fn rf_batch_create_u32(metric: c_int, query: *const u32, len: u32, device: c_int, out: *mut *mut rf_batch) -> c_int;
    pub fn rf_corpus_destroy(c: *mut rf_corpus) -> c_int;
    pub fn rf_corpus_release_csr(c: *mut rf_corpus) -> c_int;
    pub fn rf_corpus_has_csr(c: *const rf_corpus) -> c_int;
    pub fn rf_corpus_size(c: *const rf_corpus) -> u64;
    pub fn rf_batch_create_u8(metric: c_int, query: *const u8, len: u32, device: c_int, out: *mut *mut rf_batch) -> c_int;
    pub fn rf_batch_destroy(b: *mut rf_batch) -> c_int;
    pub fn rf_batch_score_u32(b: *const rf_batch, c: *const rf_corpus, kind: c_int, args: *const rf_args, out: *mut u32) -> c_int;
    pub fn rf_batch_score_u8(b: *const rf_batch, c: *const rf_corpus, kind: c_int, args: *const rf_args, out: *mut u8) -> c_int;
    pub fn rf_batch_score_f64(b: *const rf_batch, c: *const rf_corpus, kind: c_int, args: *const rf_args, out: *mut f64) -> c_int;
    pub fn rf_cdist_topk_u8(q_chars: *const u8, q_offsets: *const u64, nq: u32, c: *const rf_corpus, args: *const rf_args,
                            k: u32, idx: *mut u32, dist: *mut u32) -> c_int;
    // sharded corpora (one process per GPU): per-shard lists stay on the device, are exchanged with one all-gather
    // (NCCL) and merged by (distance, global index)
    pub fn rf_cdist_topk_u8_device(q_chars: *const u8, q_offsets: *const u64, nq: u32, c: *const rf_corpus, args: *const rf_args,
                                   k: u32, idx_device: *mut u32, dist_device: *mut u32, stream: *mut c_void) -> c_int;
    pub fn rf_topk_merge_device(idx_parts: *const u32, dist_parts: *const u32, part_stride: u64, index_base_device: *const u64,
                                parts: u32, nq: u32, k: u32, idx_out_device: *mut u64, dist_out_device: *mut u32,
                                device: c_int, stream: *mut c_void) -> c_int;
    pub fn rf_set_option(name: *const c_char, value: c_int) -> c_int;
    // one-shot scoring of host-resident candidates (chunked H2D / scan / D2H pipeline)
    pub fn rf_batch_stream_u32(b: *const rf_batch, chars: *const u8, offsets: *const u64, n: u64, kind: c_int, args: *const rf_args, out: *mut u32) -> c_int;
    pub fn rf_batch_stream_f64(b: *const rf_batch, chars: *const u8, offsets: *const u64, n: u64, kind: c_int, args: *const rf_args, out: *mut f64) -> c_int;
    // post-processing on the GPU: k best by (score, index) / every candidate within the cutoff in index order
    pub fn rf_batch_extract_u32(b: *const rf_batch, c: *const rf_corpus, kind: c_int, args: *const rf_args, k: u32, idx: *mut u32, score: *mut u32, n_out: *mut u32) -> c_int;
    pub fn rf_batch_extract_f64(b: *const rf_batch, c: *const rf_corpus, kind: c_int, args: *const rf_args, k: u32, idx: *mut u32, score: *mut f64, n_out: *mut u32) -> c_int;
    pub fn rf_batch_filter_u32(b: *const rf_batch, c: *const rf_corpus, kind: c_int, args: *const rf_args, capacity: u64, idx: *mut u32, score: *mut u32, n_hits: *mut u64) -> c_int;
    pub fn rf_batch_filter_f64(b: *const rf_batch, c: *const rf_corpus, kind: c_int, args: *const rf_args, capacity: u64, idx: *mut u32, score: *mut f64, n_hits: *mut u64) -> c_int;
    // any HashableChar element type, widened BY VALUE inside the library
    pub fn rf_corpus_create_elems(elems: *const c_void, elem_type: c_int, offsets: *const u64, n: u64, device: c_int, out: *mut *mut rf_corpus) -> c_int;
    pub fn rf_batch_create_elems(metric: c_int, query: *const c_void, elem_type: c_int, len: u32, device: c_int, out: *mut *mut rf_batch) -> c_int;
    pub fn rf_batch_set_option(b: *mut rf_batch, name: *const c_char, value: c_int) -> c_int;
    // fewer bytes on the link: one length byte per candidate in, one score byte per candidate out
    pub fn rf_batch_stream_u32_len8(b: *const rf_batch, chars: *const u8, lens: *const u8, n: u64, kind: c_int, args: *const rf_args, out: *mut u32) -> c_int;
    pub fn rf_batch_stream_u8_len8(b: *const rf_batch, chars: *const u8, lens: *const u8, n: u64, kind: c_int, args: *const rf_args, out: *mut u8) -> c_int;
    pub fn rf_batch_stream_u32_elems32(b: *const rf_batch, elems: *const u32, offsets: *const u64, n: u64, kind: c_int, args: *const rf_args, out: *mut u32) -> c_int;
    pub fn rf_cdist_topk_u32(q_elems: *const u32, q_offsets: *const u64, nq: u32, c: *const rf_corpus, args: *const rf_args, k: u32, idx: *mut u32, dist: *mut u32) -> c_int;
    // ONE process, several GPUs: the corpus sharded by bytes over `devices`, results identical to the single-GPU calls
    pub fn rf_corpus_create_sharded_u8(chars: *const u8, offsets: *const u64, n: u64, devices: *const c_int, ndev: c_int, out: *mut *mut rf_sharded_corpus) -> c_int;
    pub fn rf_sharded_corpus_destroy(c: *mut rf_sharded_corpus) -> c_int;
    pub fn rf_sharded_corpus_size(c: *const rf_sharded_corpus) -> u64;
    pub fn rf_sharded_batch_create_u8(metric: c_int, query: *const u8, len: u32, devices: *const c_int, ndev: c_int, out: *mut *mut rf_sharded_batch) -> c_int;
    pub fn rf_sharded_batch_destroy(b: *mut rf_sharded_batch) -> c_int;
    pub fn rf_sharded_score_u32(b: *const rf_sharded_batch, c: *const rf_sharded_corpus, kind: c_int, args: *const rf_args, out: *mut u32) -> c_int;
    pub fn rf_sharded_score_f64(b: *const rf_sharded_batch, c: *const rf_sharded_corpus, kind: c_int, args: *const rf_args, out: *mut f64) -> c_int;
    pub fn rf_sharded_score_u32_allgather_device(b: *const rf_sharded_batch, c: *const rf_sharded_corpus, kind: c_int, args: *const rf_args, out_device: *const *mut u32) -> c_int;
    pub fn rf_sharded_extract_u32(b: *const rf_sharded_batch, c: *const rf_sharded_corpus, kind: c_int, args: *const rf_args, k: u32, idx: *mut u64, score: *mut u32, n_out: *mut u32) -> c_int;
    pub fn rf_sharded_cdist_topk_u8(q_chars: *const u8, q_offsets: *const u64, nq: u32, c: *const rf_sharded_corpus, args: *const rf_args, k: u32, idx: *mut u64, dist: *mut u32) -> c_int;
    pub fn rf_sharded_stream_u32(b: *const rf_sharded_batch, chars: *const u8, offsets: *const u64, n: u64, kind: c_int, args: *const rf_args, out: *mut u32) -> c_int;
    // one process per GPU: communicator + scoring with the final all-gather of the score vectors
    pub fn rf_comm_unique_id(out128: *mut c_void) -> c_int;
    pub fn rf_comm_create_rank(id128: *const c_void, nranks: c_int, rank: c_int, device: c_int, out: *mut *mut rf_comm) -> c_int;
    pub fn rf_comm_destroy(comm: *mut rf_comm) -> c_int;
    pub fn rf_batch_score_u32_allgather_device(b: *const rf_batch, c: *const rf_corpus, comm: *mut rf_comm, kind: c_int, args: *const rf_args,
                                               out_device: *mut u32, out_capacity: u64, counts_out: *mut u64, stream: *mut c_void) -> c_int;
    // packing + corpus files
    pub fn rf_pack_u8(strings: *const *const u8, lengths: *const u64, n: u64, offsets_out: *mut u64, chars_out: *mut u8, nthreads: c_int) -> c_int;
    pub fn rf_corpus_file_write(path: *const c_char, chars: *const u8, offsets: *const u64, n: u64) -> c_int;
    pub fn rf_corpus_create_from_file(path: *const c_char, device: c_int, out: *mut *mut rf_corpus) -> c_int;
    #[allow(dead_code)]
    fn rf_kernel_launch_count() -> u64;
    #[allow(dead_code)]
    fn rf_device_count() -> c_int;
    #[allow(dead_code)]
    fn rf_status_string(s: c_int) -> *const c_char;
    #[allow(dead_code)]
    fn rf_result_is_float(metric: c_int, kind: c_int) -> c_int;
    #[allow(dead_code)]
    fn rf_corpus_total_chars(c: *const rf_corpus) -> u64;
    #[allow(dead_code)]
    fn rf_corpus_device(c: *const rf_corpus) -> c_int;
    #[allow(dead_code)]
    fn rf_batch_score_u32_device(b: *const rf_batch, c: *const rf_corpus, kind: c_int, args: *const rf_args, out: *mut u32, stream: *mut c_void) -> c_int;
}

fn check(status: c_int) {
    if status != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(rf_last_error()) }.to_string_lossy().into_owned();
        panic!("rfgpu status {status}: {msg}"); // the reference has no Result on this path either
    }
}

/// Packed candidates resident in GPU memory (new: the reference takes one iterator per call).
pub struct Corpus { h: *mut rf_corpus }
unsafe impl Send for Corpus {}
unsafe impl Sync for Corpus {}
impl Corpus {
    pub fn new<I, S>(candidates: I, device: i32) -> Self where I: IntoIterator<Item = S>, S: AsRef<[u8]> {
        let mut chars = Vec::new();
        let mut offsets = vec![0u64];
        for s in candidates { chars.extend_from_slice(s.as_ref()); offsets.push(chars.len() as u64); }
        let mut h = std::ptr::null_mut();
        check(unsafe { rf_corpus_create_u8(chars.as_ptr(), offsets.as_ptr(), (offsets.len() - 1) as u64, device, &mut h) });
        Corpus { h }
    }
    /// Candidates as `char` sequences (`str::chars()`), the reference's default element type for `&str` inputs.
    pub fn from_chars<I, S>(candidates: I, device: i32) -> Self where I: IntoIterator<Item = S>, S: AsRef<str> {
        let mut elems: Vec<u32> = Vec::new();
        let mut offsets = vec![0u64];
        for s in candidates { elems.extend(s.as_ref().chars().map(|c| c as u32)); offsets.push(elems.len() as u64); }
        let mut h = std::ptr::null_mut();
        check(unsafe { rf_corpus_create_u32(elems.as_ptr(), offsets.as_ptr(), (offsets.len() - 1) as u64, device, &mut h) });
        Corpus { h }
    }
    /// Corpus file written by `rf_corpus_file_write` (mmap + upload).
    pub fn from_file(path: &std::path::Path, device: i32) -> Self {
        let p = std::ffi::CString::new(path.to_string_lossy().as_bytes()).expect("path contains NUL");
        let mut h = std::ptr::null_mut();
        check(unsafe { rf_corpus_create_from_file(p.as_ptr(), device, &mut h) });
        Corpus { h }
    }
    pub fn len(&self) -> usize { unsafe { rf_corpus_size(self.h) as usize } }
    /// frees the CSR copy (45 % of the footprint); afterwards only what the interleaved layout serves works (see rfgpu.h)
    pub fn release_csr(&mut self) { check(unsafe { rf_corpus_release_csr(self.h) }); }
    pub fn has_csr(&self) -> bool { unsafe { rf_corpus_has_csr(self.h) != 0 } }
}
impl Drop for Corpus { fn drop(&mut self) { unsafe { rf_corpus_destroy(self.h); } } }

/// The candidates sharded over several GPUs of one box, driven by THIS process (no MPI, no PyTorch): the counterpart of
/// `Corpus` for `rf_corpus_create_sharded_u8`.  `BatchComparator` is `Clone + Send + Sync` in the reference
/// (src/distance/levenshtein.rs:1635-1639); so are these handles.
pub struct ShardedCorpus { h: *mut rf_sharded_corpus, devices: Vec<c_int> }
unsafe impl Send for ShardedCorpus {}
unsafe impl Sync for ShardedCorpus {}
impl ShardedCorpus {
    pub fn new<I, S>(candidates: I, devices: &[i32]) -> Self where I: IntoIterator<Item = S>, S: AsRef<[u8]> {
        let mut chars = Vec::new();
        let mut offsets = vec![0u64];
        for s in candidates { chars.extend_from_slice(s.as_ref()); offsets.push(chars.len() as u64); }
        let devs: Vec<c_int> = devices.iter().map(|&d| d as c_int).collect();
        let mut h = std::ptr::null_mut();
        check(unsafe { rf_corpus_create_sharded_u8(chars.as_ptr(), offsets.as_ptr(), (offsets.len() - 1) as u64, devs.as_ptr(), devs.len() as c_int, &mut h) });
        ShardedCorpus { h, devices: devs }
    }
    pub fn len(&self) -> usize { unsafe { rf_sharded_corpus_size(self.h) as usize } }
    /// `levenshtein::BatchComparator::new(query).distance(c)` for every candidate of the sharded corpus, in candidate order.
    pub fn levenshtein_distance<Q: AsRef<[u8]>>(&self, query: Q) -> Vec<usize> {
        let q = query.as_ref();
        let mut b = std::ptr::null_mut();
        check(unsafe { rf_sharded_batch_create_u8(RF_LEVENSHTEIN, q.as_ptr(), q.len() as u32, self.devices.as_ptr(), self.devices.len() as c_int, &mut b) });
        let mut out = vec![0u32; self.len()];
        let st = unsafe { rf_sharded_score_u32(b, self.h, RF_DISTANCE, std::ptr::null(), out.as_mut_ptr()) };
        unsafe { rf_sharded_batch_destroy(b) };
        check(st);
        out.into_iter().map(|v| v as usize).collect()
    }
    /// Many-vs-many over all shards: per query the `k` best (global index, distance); one NCCL all-gather of the per-shard lists.
    pub fn cdist_topk<I, S>(&self, queries: I, k: u32) -> Vec<Vec<(u64, u32)>> where I: IntoIterator<Item = S>, S: AsRef<[u8]> {
        let mut chars = Vec::new();
        let mut offsets = vec![0u64];
        for q in queries { chars.extend_from_slice(q.as_ref()); offsets.push(chars.len() as u64); }
        let nq = offsets.len() - 1;
        let mut idx = vec![u64::MAX; nq * k as usize];
        let mut dist = vec![u32::MAX; nq * k as usize];
        check(unsafe { rf_sharded_cdist_topk_u8(chars.as_ptr(), offsets.as_ptr(), nq as u32, self.h, std::ptr::null(), k, idx.as_mut_ptr(), dist.as_mut_ptr()) });
        (0..nq).map(|q| (0..k as usize).map(|i| (idx[q * k as usize + i], dist[q * k as usize + i]))
                                       .take_while(|&(i, _)| i != u64::MAX).collect()).collect()
    }
}
impl Drop for ShardedCorpus { fn drop(&mut self) { unsafe { rf_sharded_corpus_destroy(self.h); } } }

/// Element types other than `u8` / `char` go to the library as they are (`rf_*_create_elems` widens BY VALUE).
pub trait Elem: Copy { const TYPE: c_int; }
impl Elem for u8 { const TYPE: c_int = RF_ELEM_U8; }
impl Elem for u16 { const TYPE: c_int = RF_ELEM_U16; }
impl Elem for u32 { const TYPE: c_int = RF_ELEM_U32; }
impl Elem for u64 { const TYPE: c_int = RF_ELEM_U64; }
impl Elem for i8 { const TYPE: c_int = RF_ELEM_I8; }
impl Elem for i16 { const TYPE: c_int = RF_ELEM_I16; }
impl Elem for i32 { const TYPE: c_int = RF_ELEM_I32; }
impl Elem for i64 { const TYPE: c_int = RF_ELEM_I64; }
impl Corpus {
    pub fn from_elems<T: Elem>(candidates: &[&[T]], device: i32) -> Self {
        let mut elems: Vec<T> = Vec::new();
        let mut offsets = vec![0u64];
        for s in candidates { elems.extend_from_slice(s); offsets.push(elems.len() as u64); }
        let mut h = std::ptr::null_mut();
        check(unsafe { rf_corpus_create_elems(elems.as_ptr() as *const c_void, T::TYPE, offsets.as_ptr(), (offsets.len() - 1) as u64, device, &mut h) });
        Corpus { h }
    }
}

/// Many-vs-many (new on this side): for every query the `k` best candidates of `corpus` by (Levenshtein distance, index);
/// rows with fewer than `k` hits are shorter.  Queries of at most 64 bytes.
pub fn cdist_topk<I, S>(queries: I, corpus: &Corpus, k: u32, score_cutoff: Option<u64>) -> Vec<Vec<(u32, u32)>>
where I: IntoIterator<Item = S>, S: AsRef<[u8]> {
    let mut chars = Vec::new();
    let mut offsets = vec![0u64];
    for q in queries { chars.extend_from_slice(q.as_ref()); offsets.push(chars.len() as u64); }
    let nq = offsets.len() - 1;
    let mut a: rf_args = unsafe { std::mem::zeroed() };
    unsafe { rf_args_default(&mut a) };
    if let Some(c) = score_cutoff { a.has_cutoff = 1; a.cutoff_u = c; }
    let mut idx = vec![u32::MAX; nq * k as usize];
    let mut dist = vec![u32::MAX; nq * k as usize];
    check(unsafe { rf_cdist_topk_u8(chars.as_ptr(), offsets.as_ptr(), nq as u32, corpus.h, &a, k, idx.as_mut_ptr(), dist.as_mut_ptr()) });
    (0..nq).map(|q| (0..k as usize).map(|i| (idx[q * k as usize + i], dist[q * k as usize + i]))
                                   .take_while(|&(i, _)| i != u32::MAX).collect()).collect()
}

/// Elements of the other `HashableChar` integer types (src/details/common.rs:29-37) widened to the ABI's u32 BY VALUE
/// (the reference compares elements numerically): unsigned types zero-extend, negative values keep their
/// two's-complement 32-bit pattern.  `None` when a value does not fit or the sequence would be ambiguous.
pub fn widen_elements<T: Copy + Into<i128>>(elems: &[T]) -> Option<Vec<u32>> {
    let (mut neg, mut big) = (false, false);
    let mut out = Vec::with_capacity(elems.len());
    for &e in elems {
        let v: i128 = e.into();
        if v < -(1i128 << 31) || v >= (1i128 << 32) { return None; }
        neg |= v < 0;
        big |= v >= (1i128 << 31);
        out.push((v as i64) as u32);
    }
    if neg && big { None } else { Some(out) }
}

/// `NoScoreCutoff` / `WithScoreCutoff<T>` exactly as in src/common.rs:4-86: the cutoff type picks the output type.
#[derive(Default, Copy, Clone)] pub struct NoScoreCutoff;
#[derive(Default, Copy, Clone)] pub struct WithScoreCutoff<T>(pub T);
pub trait Cutoff<T: Copy> { type Output; fn cutoff(&self) -> Option<T>; fn wrap(raw: T, none: bool) -> Self::Output; }
impl<T: Copy> Cutoff<T> for NoScoreCutoff { type Output = T; fn cutoff(&self) -> Option<T> { None } fn wrap(raw: T, _: bool) -> T { raw } }
impl<T: Copy> Cutoff<T> for WithScoreCutoff<T> { type Output = Option<T>; fn cutoff(&self) -> Option<T> { Some(self.0) } fn wrap(raw: T, none: bool) -> Option<T> { (!none).then_some(raw) } }

macro_rules! metric_module {
    ($modname:ident, $metric:expr, $int_t:ty, $int_is_f64:expr) => {
        pub mod $modname {
            use super::*;
            #[derive(Copy, Clone)]
            pub struct Args<ResultType, CutoffType> {
                pub(crate) score_cutoff: CutoffType,
                pub(crate) score_hint: Option<ResultType>,
                pub(crate) weights: (u64, u64, u64),
                pub(crate) prefix_weight: f64,
            }
            impl<R> Default for Args<R, NoScoreCutoff> {
                fn default() -> Self { Args { score_cutoff: NoScoreCutoff, score_hint: None, weights: (1, 1, 1), prefix_weight: 0.1 } }
            }
            impl<R: Copy, C> Args<R, C> {
                pub fn score_hint(mut self, h: R) -> Self { self.score_hint = Some(h); self }
                pub fn score_cutoff(self, c: R) -> Args<R, WithScoreCutoff<R>> {
                    Args { score_cutoff: WithScoreCutoff(c), score_hint: self.score_hint, weights: self.weights, prefix_weight: self.prefix_weight }
                }
                pub fn weights(mut self, ins: u64, del: u64, sub: u64) -> Self { self.weights = (ins, del, sub); self }
                pub fn prefix_weight(mut self, w: f64) -> Self { self.prefix_weight = w; self }
            }
            /// `BatchComparator::new(query)`: caches s1 and its pattern-match bit table on the GPU.
            pub struct BatchComparator { h: *mut rf_batch }
            unsafe impl Send for BatchComparator {}
            unsafe impl Sync for BatchComparator {}
            impl Drop for BatchComparator { fn drop(&mut self) { unsafe { rf_batch_destroy(self.h); } } }
            impl BatchComparator {
                pub fn new<Q: AsRef<[u8]>>(query: Q, device: i32) -> Self {
                    let q = query.as_ref();
                    let mut h = std::ptr::null_mut();
                    check(unsafe { rf_batch_create_u8($metric, q.as_ptr(), q.len() as u32, device, &mut h) });
                    BatchComparator { h }
                }
                fn c_args_f(cut: Option<f64>, w: (u64, u64, u64), pw: f64) -> rf_args {
                    let mut a: rf_args = unsafe { std::mem::zeroed() };
                    unsafe { rf_args_default(&mut a) };
                    a.insertion_cost = w.0; a.deletion_cost = w.1; a.substitution_cost = w.2; a.prefix_weight = pw;
                    if let Some(c) = cut { a.has_cutoff = 1; a.cutoff_f = c; }
                    a
                }
                fn run_u32(&self, c: &Corpus, kind: c_int, a: &rf_args) -> Vec<u32> {
                    let mut out = vec![0u32; c.len()];
                    check(unsafe { rf_batch_score_u32(self.h, c.h, kind, a, out.as_mut_ptr()) });
                    out
                }
                /// integer scores as bytes (`None` = 0xFF): a quarter of the result download of `run_u32`
                pub fn run_u8(&self, c: &Corpus, kind: c_int, a: &rf_args) -> Vec<u8> {
                    let mut out = vec![0u8; c.len()];
                    check(unsafe { rf_batch_score_u8(self.h, c.h, kind, a, out.as_mut_ptr()) });
                    out
                }
                fn run_f64(&self, c: &Corpus, kind: c_int, a: &rf_args) -> Vec<f64> {
                    let mut out = vec![0f64; c.len()];
                    check(unsafe { rf_batch_score_f64(self.h, c.h, kind, a, out.as_mut_ptr()) });
                    out
                }
                /// one score per candidate == the user's `for c in candidates { scorer.normalized_similarity_with_args(c, args) }`
                pub fn normalized_similarity_with_args<C: Cutoff<f64>>(&self, c: &Corpus, args: &Args<f64, C>) -> Vec<C::Output> {
                    let a = Self::c_args_f(args.score_cutoff.cutoff(), args.weights, args.prefix_weight);
                    self.run_f64(c, RF_NORMALIZED_SIMILARITY, &a).into_iter().map(|v| C::wrap(v, v.is_nan())).collect()
                }
                pub fn normalized_distance_with_args<C: Cutoff<f64>>(&self, c: &Corpus, args: &Args<f64, C>) -> Vec<C::Output> {
                    let a = Self::c_args_f(args.score_cutoff.cutoff(), args.weights, args.prefix_weight);
                    self.run_f64(c, RF_NORMALIZED_DISTANCE, &a).into_iter().map(|v| C::wrap(v, v.is_nan())).collect()
                }
                pub fn normalized_similarity(&self, c: &Corpus) -> Vec<f64> { self.normalized_similarity_with_args(c, &Args::default()) }
                pub fn normalized_distance(&self, c: &Corpus) -> Vec<f64> { self.normalized_distance_with_args(c, &Args::default()) }
            }
            metric_module!(@intmethods $int_t, $int_is_f64);
        }
    };
    // distance / similarity: usize-valued for the edit-distance family ...
    (@intmethods $int_t:ty, false) => {
        impl BatchComparator {
            pub fn distance_with_args<C: Cutoff<usize>>(&self, c: &Corpus, args: &Args<usize, C>) -> Vec<C::Output> {
                let mut a = Self::c_args_f(None, args.weights, args.prefix_weight);
                if let Some(k) = args.score_cutoff.cutoff() { a.has_cutoff = 1; a.cutoff_u = k as u64; }
                self.run_u32(c, RF_DISTANCE, &a).into_iter().map(|v| C::wrap(v as usize, v == u32::MAX)).collect()
            }
            pub fn similarity_with_args<C: Cutoff<usize>>(&self, c: &Corpus, args: &Args<usize, C>) -> Vec<C::Output> {
                let mut a = Self::c_args_f(None, args.weights, args.prefix_weight);
                if let Some(k) = args.score_cutoff.cutoff() { a.has_cutoff = 1; a.cutoff_u = k as u64; }
                self.run_u32(c, RF_SIMILARITY, &a).into_iter().map(|v| C::wrap(v as usize, v == u32::MAX)).collect()
            }
            pub fn distance(&self, c: &Corpus) -> Vec<usize> { self.distance_with_args(c, &Args::default()) }
            pub fn similarity(&self, c: &Corpus) -> Vec<usize> { self.similarity_with_args(c, &Args::default()) }
        }
    };
    // ... and f64-valued for Jaro / Jaro-Winkler
    (@intmethods $int_t:ty, true) => {
        impl BatchComparator {
            pub fn distance_with_args<C: Cutoff<f64>>(&self, c: &Corpus, args: &Args<f64, C>) -> Vec<C::Output> {
                let a = Self::c_args_f(args.score_cutoff.cutoff(), args.weights, args.prefix_weight);
                self.run_f64(c, RF_DISTANCE, &a).into_iter().map(|v| C::wrap(v, v.is_nan())).collect()
            }
            pub fn similarity_with_args<C: Cutoff<f64>>(&self, c: &Corpus, args: &Args<f64, C>) -> Vec<C::Output> {
                let a = Self::c_args_f(args.score_cutoff.cutoff(), args.weights, args.prefix_weight);
                self.run_f64(c, RF_SIMILARITY, &a).into_iter().map(|v| C::wrap(v, v.is_nan())).collect()
            }
            pub fn distance(&self, c: &Corpus) -> Vec<f64> { self.distance_with_args(c, &Args::default()) }
            pub fn similarity(&self, c: &Corpus) -> Vec<f64> { self.similarity_with_args(c, &Args::default()) }
        }
    };
}

pub mod distance {
    use super::*;
    metric_module!(levenshtein, RF_LEVENSHTEIN, usize, false);
    metric_module!(indel, RF_INDEL, usize, false);
    metric_module!(lcs_seq, RF_LCS_SEQ, usize, false);
    metric_module!(osa, RF_OSA, usize, false);
    metric_module!(jaro, RF_JARO, f64, true);
    metric_module!(jaro_winkler, RF_JARO_WINKLER, f64, true);
    metric_module!(prefix, RF_PREFIX, usize, false);
    metric_module!(postfix, RF_POSTFIX, usize, false);
    metric_module!(damerau_levenshtein, RF_DAMERAU_LEVENSHTEIN, usize, false);
    // hamming: same shape with `Args::pad` -> `rf_args.pad`; without it a candidate of another length makes the call
    // return status 1 ("Differing length arguments provided"), which maps to Err(hamming::Error::DifferentLengthArgs).
    metric_module!(hamming, RF_HAMMING, usize, false);
}

/// `fuzz::RatioBatchComparator` (src/fuzz.rs:98-150); documented semantics (== `fuzz::ratio`), see DESIGN.md Q1.
pub mod fuzz {
    use super::*;
    pub struct RatioBatchComparator { h: *mut rf_batch }
    impl Drop for RatioBatchComparator { fn drop(&mut self) { unsafe { rf_batch_destroy(self.h); } } }
    impl RatioBatchComparator {
        pub fn new<Q: AsRef<[u8]>>(query: Q, device: i32) -> Self {
            let q = query.as_ref();
            let mut h = std::ptr::null_mut();
            check(unsafe { rf_batch_create_u8(RF_RATIO, q.as_ptr(), q.len() as u32, device, &mut h) });
            RatioBatchComparator { h }
        }
        pub fn similarity(&self, c: &Corpus) -> Vec<f64> {
            let mut out = vec![0f64; c.len()];
            check(unsafe { rf_batch_score_f64(self.h, c.h, RF_SIMILARITY, std::ptr::null(), out.as_mut_ptr()) });
            out
        }
    }
}
