//! Raw bindings of include/umigpu.h (one declaration per C entry point, same order) plus a small
//! safe wrapper.  NOT COMPILED HERE (no Rust toolchain in the build image); kept in sync with the
//! header by hand — tests/test_abi.py checks the header against the built library.
#![allow(non_camel_case_types)]
use libc::{c_char, c_float, c_int, c_void};

pub const UMIGPU_ALGO_DIR: i32 = 0;
pub const UMIGPU_ALGO_ADJ: i32 = 1;
pub const UMIGPU_ALGO_ADJ_UPSTREAM: i32 = 2;
pub const UMIGPU_ALGO_CC: i32 = 3;
pub const UMIGPU_MERGE_ANY: i32 = 0;
pub const UMIGPU_MERGE_AVGQUAL: i32 = 1;
pub const UMIGPU_MERGE_MAPQUAL: i32 = 2;
pub const UMIGPU_FLAG_LABELS: u32 = 1;

#[repr(C)]
pub struct umigpu_ctx { _private: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy)]
pub struct umigpu_config {
    pub k: i32,
    pub percentage: c_float,
    pub algo: i32,
    pub merge: i32,
    pub umi_len: u32,
    pub device: i32,
    pub flags: u32,
    pub reserved: u32,
    pub stream: *mut c_void,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct umigpu_counters {
    pub total_reads: u64, pub n_buckets: u64, pub total_umis: u64, pub max_umis: u64, pub n_kept: u64,
    pub unordered_pairs: u64, pub pairs_evaluated: u64, pub n_edges: u64, pub n_tile_items: u64,
    pub n_tile_candidates: u64, pub n_sweeps: u64, pub n_block_pairs: u64, pub n_unmapped: u64,
    pub n_unpaired: u64, pub n_chimeric: u64, pub n_mates_skipped: u64, pub key_bits: u64,
}

#[repr(C)]
pub struct umigpu_group { _private: [u8; 0] }

/// The bucket whose neighbour search is split over the devices of a shard group.
#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct umigpu_hot { pub present: i32, pub owner: i32, pub read_index: u64, pub reads_est: u64 }

#[repr(C)]
pub struct umigpu_result {
    pub n_kept: u64,
    pub kept_read_index: *const u64,
    pub n_reads: u64,
    pub read_cluster_root: *const u64,
    pub counters: umigpu_counters,
    pub read_umi_rep: *const u64,
}

extern "C" {
    pub fn umigpu_version() -> *const c_char;
    pub fn umigpu_device_init(device: i32) -> c_int;
    pub fn umigpu_last_error(ctx: *const umigpu_ctx) -> *const c_char;
    pub fn umigpu_create(cfg: *const umigpu_config, out: *mut *mut umigpu_ctx) -> c_int;
    pub fn umigpu_destroy(ctx: *mut umigpu_ctx);
    pub fn umigpu_reset(ctx: *mut umigpu_ctx) -> c_int;
    pub fn umigpu_push_reads(ctx: *mut umigpu_ctx, n: u64, tid: *const i32, unclipped_pos: *const i64, is_reverse: *const u8,
                             umi_ascii: *const u8, score: *const i32, weight: *const i32, first_read_index: u64) -> c_int;
    pub fn umigpu_push_reads_device(ctx: *mut umigpu_ctx, n: u64, tid: *const i32, unclipped_pos: *const i64, is_reverse: *const u8,
                                    umi_ascii: *const u8, score: *const i32, weight: *const i32, first_read_index: u64) -> c_int;
    pub fn umigpu_push_reads_paired(ctx: *mut umigpu_ctx, n: u64, tid: *const i32, unclipped_pos: *const i64, is_reverse: *const u8,
                                    tlen: *const i64, umi_ascii: *const u8, score: *const i32, weight: *const i32,
                                    first_read_index: u64) -> c_int;
    pub fn umigpu_push_reads_packed(ctx: *mut umigpu_ctx, n: u64, tid: *const i32, pos32: *const i32, is_reverse: *const u8,
                                    umi_2bit: *const c_void, n_mask: *const u32, score8: *const u8, first_read_index: u64) -> c_int;
    pub fn umigpu_push_bam_records(ctx: *mut umigpu_ctx, n: u64, records: *const u8, offsets: *const u64, umi_sep: u8,
                                   first_read_index: u64, n_unmapped: *mut u64) -> c_int;
    pub fn umigpu_bam_record_offsets(buf: *const u8, len: u64, offsets: *mut u64, max_records: u64, n_records: *mut u64,
                                     consumed: *mut u64) -> c_int;
    pub fn umigpu_run(ctx: *mut umigpu_ctx) -> c_int;
    pub fn umigpu_fetch(ctx: *mut umigpu_ctx, out: *mut umigpu_result) -> c_int;
    pub fn umigpu_finish(ctx: *mut umigpu_ctx, out: *mut umigpu_result) -> c_int;
    pub fn umigpu_get_counters(ctx: *mut umigpu_ctx, out: *mut umigpu_counters) -> c_int;
    pub fn umigpu_cluster_bucket(ctx: *mut umigpu_ctx, n: u64, umi_ascii: *const u8, freq: *const i32, keep: *mut u8, label: *mut i32) -> c_int;
    pub fn umigpu_remove_near(ctx: *mut umigpu_ctx, n: u64, umi_ascii: *const u8, freq: *const i32, query: *const u8,
                              k: i32, max_freq: i32, out: *mut u8) -> c_int;
    pub fn umigpu_neighbours(ctx: *mut umigpu_ctx, n: u64, umi_ascii: *const u8, freq: *const i32, apply_rule: i32,
                             row_ptr: *mut u64, col: *mut u32, col_capacity: u64, n_edges: *mut u64) -> c_int;
    pub fn umigpu_avg_qual(ctx: *mut umigpu_ctx, n: u64, qual: *const u8, offsets: *const u64, out: *mut i32) -> c_int;
    pub fn umigpu_stage_ms(ctx: *mut umigpu_ctx, stage: c_int, ms: *mut c_float) -> c_int;
    pub fn umigpu_launch_count(ctx: *mut umigpu_ctx, reset: c_int) -> u64;
    pub fn umigpu_result_free(ctx: *mut umigpu_ctx);
    pub fn umigpu_shard_plan(n: u64, tid: *const i32, unclipped_pos: *const i64, is_reverse: *const u8, n_shards: i32,
                             shard_of_read: *mut i32, shard_cost: *mut u64) -> c_int;
    pub fn umigpu_dedup_sharded(cfg: *const umigpu_config, n_devices: i32, device_ids: *const i32, n: u64, tid: *const i32,
                                unclipped_pos: *const i64, is_reverse: *const u8, umi_ascii: *const u8, score: *const i32,
                                kept: *mut *mut u64, n_kept: *mut u64, counters: *mut umigpu_counters) -> c_int;
    pub fn umigpu_free(p: *mut c_void);
    // several devices, ONE dataset
    pub fn umigpu_pos_key(tid: i32, unclipped_pos: i64) -> i64;
    pub fn umigpu_shard_plan_sorted(n: u64, tid: *const i32, unclipped_pos: *const i64, is_reverse: *const u8, n_shards: i32,
                                    hot_min_reads: u64, cuts: *mut u64, cut_keys: *mut i64, hot: *mut umigpu_hot, shard_cost: *mut f64) -> c_int;
    pub fn umigpu_xchg_create(ctx: *mut umigpu_ctx, rank: i32, n_ranks: i32, max_hot_uniques: u64, max_hot_edges: u64,
                              ipc_handle_out: *mut u8) -> c_int;
    pub fn umigpu_xchg_attach_ipc(ctx: *mut umigpu_ctx, handles: *const u8) -> c_int;
    pub fn umigpu_xchg_attach_local(ctx: *mut umigpu_ctx, group: *const *mut umigpu_ctx) -> c_int;
    pub fn umigpu_run_sharded(ctx: *mut umigpu_ctx, hot: *const umigpu_hot, key_lo: i64, key_hi: i64) -> c_int;
    pub fn umigpu_group_create(cfg: *const umigpu_config, n_devices: i32, device_ids: *const i32, out: *mut *mut umigpu_group) -> c_int;
    pub fn umigpu_group_destroy(g: *mut umigpu_group);
    pub fn umigpu_group_context(g: *mut umigpu_group, rank: i32) -> *mut umigpu_ctx;
    pub fn umigpu_group_dedup(g: *mut umigpu_group, n: u64, tid: *const i32, unclipped_pos: *const i64, is_reverse: *const u8,
                              umi_ascii: *const u8, score: *const i32, kept: *mut *mut u64, n_kept: *mut u64,
                              counters: *mut umigpu_counters, rank_ms: *mut c_float) -> c_int;
    pub fn umigpu_int_peak(ctx: *mut umigpu_ctx, lop3_ops_per_s: *mut f64, popc_ops_per_s: *mut f64) -> c_int;
}

/// Safe owner of a context.  Errors become panics, like every error in the reference
/// (`expect`/`panic!` with panic = "abort", Cargo.toml:19).
pub struct Context { raw: *mut umigpu_ctx }

impl Context {
    pub fn new(cfg: umigpu_config) -> Self {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { umigpu_create(&cfg, &mut raw) };
        if rc != 0 { panic!("umigpu_create failed ({rc}): {}", last_error(std::ptr::null())); }
        Self { raw }
    }
    fn check(&self, rc: c_int, what: &str) {
        if rc != 0 { panic!("{what} failed ({rc}): {}", last_error(self.raw)); }
    }
    pub fn push_reads(&mut self, tid: &[i32], pos: &[i64], rev: &[u8], umi: &[u8], score: Option<&[i32]>, first_read_index: u64) {
        let n = tid.len();
        assert!(pos.len() == n && rev.len() == n && umi.len() % n.max(1) == 0);
        let rc = unsafe {
            umigpu_push_reads(self.raw, n as u64, tid.as_ptr(), pos.as_ptr(), rev.as_ptr(), umi.as_ptr(),
                              score.map_or(std::ptr::null(), |s| s.as_ptr()), std::ptr::null(), first_read_index)
        };
        self.check(rc, "umigpu_push_reads");
    }
    /// `--paired`: the template length joins the bucket key (PairedAlignment, deduplicate_sam.rs:545-565).
    pub fn push_reads_paired(&mut self, tid: &[i32], pos: &[i64], rev: &[u8], tlen: &[i64], umi: &[u8], score: Option<&[i32]>, first_read_index: u64) {
        let n = tid.len();
        assert!(pos.len() == n && rev.len() == n && tlen.len() == n && umi.len() % n.max(1) == 0);
        let rc = unsafe {
            umigpu_push_reads_paired(self.raw, n as u64, tid.as_ptr(), pos.as_ptr(), rev.as_ptr(), tlen.as_ptr(), umi.as_ptr(),
                                     score.map_or(std::ptr::null(), |s| s.as_ptr()), std::ptr::null(), first_read_index)
        };
        self.check(rc, "umigpu_push_reads_paired");
    }
    /// Kept read indices (ascending input order) and the end-of-run counters.
    pub fn finish(&mut self) -> (&[u64], umigpu_counters) {
        let mut res = std::mem::MaybeUninit::<umigpu_result>::zeroed();
        let rc = unsafe { umigpu_finish(self.raw, res.as_mut_ptr()) };
        self.check(rc, "umigpu_finish");
        let res = unsafe { res.assume_init() };
        let kept = if res.n_kept == 0 { &[][..] } else { unsafe { std::slice::from_raw_parts(res.kept_read_index, res.n_kept as usize) } };
        (kept, res.counters)
    }
    pub fn reset(&mut self) { let rc = unsafe { umigpu_reset(self.raw) }; self.check(rc, "umigpu_reset"); }
}

impl Drop for Context { fn drop(&mut self) { unsafe { umigpu_destroy(self.raw) } } }

fn last_error(ctx: *const umigpu_ctx) -> String {
    unsafe { std::ffi::CStr::from_ptr(umigpu_last_error(ctx)).to_string_lossy().into_owned() }
}

/// Several GPUs of one box in one process: contiguous coordinate slices, hot bucket split over NVLink (umigpu_group_*).
pub struct Group { raw: *mut umigpu_group }

impl Group {
    pub fn new(cfg: umigpu_config, devices: &[i32]) -> Self {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { umigpu_group_create(&cfg, devices.len() as i32, devices.as_ptr(), &mut raw) };
        if rc != 0 { panic!("umigpu_group_create failed ({rc}): {}", last_error(std::ptr::null())); }
        Self { raw }
    }
    /// Kept read indices of the whole dataset (ascending) and the summed counters.
    pub fn dedup(&mut self, tid: &[i32], pos: &[i64], rev: &[u8], umi: &[u8], score: Option<&[i32]>) -> (Vec<u64>, umigpu_counters) {
        let n = tid.len();
        assert!(pos.len() == n && rev.len() == n && umi.len() % n.max(1) == 0);
        let (mut kept, mut n_kept, mut ctr) = (std::ptr::null_mut(), 0u64, umigpu_counters::default());
        let rc = unsafe {
            umigpu_group_dedup(self.raw, n as u64, tid.as_ptr(), pos.as_ptr(), rev.as_ptr(), umi.as_ptr(),
                               score.map_or(std::ptr::null(), |s| s.as_ptr()), &mut kept, &mut n_kept, &mut ctr, std::ptr::null_mut())
        };
        if rc != 0 { panic!("umigpu_group_dedup failed ({rc}): {}", last_error(std::ptr::null())); }
        let out = if n_kept == 0 { Vec::new() } else { unsafe { std::slice::from_raw_parts(kept, n_kept as usize) }.to_vec() };
        unsafe { umigpu_free(kept as *mut c_void) };
        (out, ctr)
    }
}

impl Drop for Group { fn drop(&mut self) { unsafe { umigpu_group_destroy(self.raw) } } }
