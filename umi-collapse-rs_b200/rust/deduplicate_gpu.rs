//! The one file a maintainer adds to the reference (`src/deduplicate_gpu.rs`): a new
//! `impl DeduplicateInterface` that replaces HOT LOOP A's map updates and all of HOT LOOP B
//! (src/deduplicate_sam.rs:148-233) with libumigpu, keeping htslib I/O, the CLI and every flag.
//! NOT COMPILED HERE (no Rust toolchain in the build image).  See INTEGRATION.md, which also lists the two one-word
//! visibility changes this file needs in the reference (`pub(crate) struct UcWriter`, `pub(crate) fn umi_pattern`).
use std::time::SystemTime;

use memchr::arch::x86_64::avx2::memchr::One;
use rust_htslib::bam::{Read, Reader, Record};
use tracing::{debug, info};
use umigpu_sys as gpu;

use crate::cli::Cli;
use crate::deduplicate_sam::{DeduplicateInterface, UcWriter};
use crate::utils::get_unclipped_pos;
use crate::utils::read::UcSAMRead;

const CHUNK: usize = 1 << 22; // reads per umigpu_push_reads call (H2D overlaps BAM decoding)

pub struct DeduplicateGPU { algo: i32, merge: i32 }

impl DeduplicateGPU {
    /// main.rs:52-92: ("dir"|"adj"|"cc", "any"|"avgqual"|"mapqual"); `--data` is ignored as in the reference.
    pub fn new(args: &Cli) -> Self {
        let algo = match args.algo_str.as_str() { "dir" => gpu::UMIGPU_ALGO_DIR, "adj" => gpu::UMIGPU_ALGO_ADJ, "cc" => gpu::UMIGPU_ALGO_CC,
            _ => panic!("Invalid algorithm combination: {} , {} and {}", args.algo_str, args.merge_str.as_ref().unwrap(), args.data_str) };
        let merge = match args.merge_str.as_ref().unwrap().as_str() { "any" => gpu::UMIGPU_MERGE_ANY, "avgqual" => gpu::UMIGPU_MERGE_AVGQUAL,
            "mapqual" => gpu::UMIGPU_MERGE_MAPQUAL, m => panic!("Invalid merge {m}") };
        Self { algo, merge }
    }
}

impl DeduplicateInterface for DeduplicateGPU {
    fn deduplicate_and_merge(&mut self, args: &Cli, start_time: &SystemTime) {
        // the reference collects ClusterTrackers under --tag and then writes nothing (deduplicate_sam.rs:236-239); the GPU arm
        // refuses instead of silently ignoring the flag (umicollapse_gpu, the C++ twin, implements it from FLAG_LABELS)
        if args.track_clusters { panic!("--tag is not wired into the GPU arm of the Rust host yet"); }
        let pattern = UcSAMRead::umi_pattern(args.umi_separator);           // utils/read.rs:65-75, used once for the autodetection
        let one = One::new(args.umi_separator).expect("failed to create a new searcher");
        let mut reader = Reader::from_path(&args.input).expect("Invalid input path");
        reader.set_threads(args.num_threads).expect("Failed to set the number of threads for reader.");
        let mut writer = UcWriter::new(&args.input, &args.output, &reader, args.paired, args);

        let (mut tid, mut pos, mut rev, mut umi, mut score) = (Vec::new(), Vec::new(), Vec::new(), Vec::new(), Vec::new());
        let mut file_index: Vec<u64> = Vec::new();                           // pushed read number -> record number in the input
        let mut tlen: Vec<i64> = Vec::new();                                 // --paired only: PairedAlignment's fourth field
        let (mut unpaired, mut chimeric) = (0u64, 0u64);
        let mut ctx: Option<gpu::Context> = None;
        let mut umi_len = args.umi_length;
        let (mut total, mut unmapped, mut index) = (0u64, 0u64, 0u64);
        let mut first_of_chunk = 0u64;
        let mut record = Record::new();
        while let Some(r) = reader.read(&mut record) {
            r.expect("Failed to parse record");
            let this = index; index += 1;                                   // index in the input file
            // --paired filters, in the order of deduplicate_sam.rs:96-129 (mates are written back by UcWriter::close)
            if args.paired && record.is_paired() && record.is_last_in_template() { continue; }
            total += 1;
            if record.is_unmapped() { unmapped += 1; if args.keep_unmapped { writer.write(&record).unwrap(); } continue; }
            if args.paired {
                if !record.is_paired() { unpaired += 1; if args.remove_unpaired { continue; } }
                if record.is_paired() && record.is_mate_unmapped() { unmapped += 1; continue; }
                if record.is_paired() && record.tid() != record.mtid() { chimeric += 1; if args.remove_chimeric { continue; } }
            }
            let qname = record.qname();
            let p = one.find(qname).expect("failed to get the umi");        // utils/read.rs:100-110
            // utils/read.rs:87-94: the length is the regex's first group — the FIRST separator that is followed by [ATCGN]+
            // (caseless), which may be a later separator than the one get_umi cuts at
            if umi_len == 0 {
                let caps = pattern.captures(qname).unwrap().unwrap();
                umi_len = caps.get(1).expect("No UMI group found in pattern match").as_bytes().len();
            }
            // reads are numbered densely in push order; file_index maps a kept number back to its record in the input
            if tid.is_empty() { first_of_chunk = file_index.len() as u64; }
            file_index.push(this);
            tid.push(record.tid());
            pos.push(get_unclipped_pos(&record));                           // utils/mod.rs:96-104
            rev.push(record.is_reverse() as u8);
            if args.paired { tlen.push(record.insert_size()); }
            umi.extend_from_slice(&qname[p + 1..p + 1 + umi_len]);
            score.push(if self.merge == gpu::UMIGPU_MERGE_MAPQUAL { record.mapq() as i32 }
                       else { let q = record.qual(); (q.iter().map(|&b| b as f32).sum::<f32>() / record.seq_len() as f32) as i32 }); // read.rs:56-63
            if tid.len() == CHUNK { flush(&mut ctx, args, self, umi_len, &mut tid, &mut pos, &mut rev, &mut tlen, &mut umi, &mut score, first_of_chunk); }
        }
        flush(&mut ctx, args, self, umi_len, &mut tid, &mut pos, &mut rev, &mut tlen, &mut umi, &mut score, first_of_chunk);
        info!("UMI collapsing reading finished in {:?} seconds", SystemTime::now().duration_since(*start_time).unwrap().as_secs_f32());
        drop(reader);

        // all buckets at once on the GPU, then a second pass over the input writes the survivors in input order
        let Some(mut ctx) = ctx else {
            // no read survived the filters: the reference writes a header-only BAM (its bucket map is simply empty)
            writer.close();
            debug!("Number of input reads: {}", total);
            debug!("Number of removed unmapped reads: {}", unmapped);
            debug!("Number of reads after deduplicating: 0");
            return;
        };
        let (kept, ctr) = ctx.finish();
        let mut reader = Reader::from_path(&args.input).expect("Invalid input path");
        reader.set_threads(args.num_threads).unwrap();
        let (mut i, mut k) = (0u64, 0usize);
        while k < kept.len() { reader.read(&mut record).unwrap().unwrap(); if i == file_index[kept[k] as usize] { writer.write(&record).unwrap(); k += 1; } i += 1; }
        writer.close();

        debug!("Number of input reads: {}", total);                         // deduplicate_sam.rs:243-267
        debug!("Number of removed unmapped reads: {}", unmapped);
        if args.paired { debug!("Number of unpaired reads: {}", unpaired); debug!("Number of chimeric reads: {}", chimeric); }
        debug!("Number of unique alignment positions: {}", ctr.n_buckets);
        debug!("Number of UMIs: {}", ctr.total_umis);
        debug!("Average number of UMIs per alignment position: {}", ctr.total_umis as f64 / ctr.n_buckets as f64);
        debug!("Max number of UMIs over all alignment positions: {}", ctr.max_umis);
        debug!("Number of reads after deduplicating: {}", ctr.n_kept);
    }
}

#[allow(clippy::too_many_arguments)]
fn flush(ctx: &mut Option<gpu::Context>, args: &Cli, me: &DeduplicateGPU, umi_len: usize, tid: &mut Vec<i32>, pos: &mut Vec<i64>,
         rev: &mut Vec<u8>, tlen: &mut Vec<i64>, umi: &mut Vec<u8>, score: &mut Vec<i32>, first: u64) {
    if tid.is_empty() { return; }
    let c = ctx.get_or_insert_with(|| gpu::Context::new(gpu::umigpu_config {
        k: args.k, percentage: args.percentage, algo: me.algo, merge: me.merge, umi_len: umi_len as u32, device: 0,
        flags: 0, reserved: 0, stream: std::ptr::null_mut() }));
    if args.paired { c.push_reads_paired(tid, pos, rev, tlen, umi, Some(score), first); } else { c.push_reads(tid, pos, rev, umi, Some(score), first); }
    tid.clear(); pos.clear(); rev.clear(); tlen.clear(); umi.clear(); score.clear();
}

// ---- several GPUs of one box (UMICOLLAPSE_GPUS=8): the same reader loop keeps ALL reads' arrays (25 B per read) instead of
// flushing chunks, then ONE call shards the coordinate-sorted stream into contiguous slices, one per device, splits the
// hot locus' neighbour search over NVLink, and returns the merged kept list (include/umigpu.h, "Several devices, ONE
// dataset"):
//
//     let group = gpu::Group::new(cfg, &(0..n_gpus).collect::<Vec<i32>>());      // umigpu_group_create
//     let (kept, ctr) = group.dedup(&tid, &pos, &rev, &umi, Some(&score));       // umigpu_group_dedup -> ascending indices
//
// followed by the same second pass over the input.  Input that is not coordinate-sorted is detected by the devices and
// re-routed through the hash plan inside the call.
