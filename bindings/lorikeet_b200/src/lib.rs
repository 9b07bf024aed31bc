//! Safe wrapper over `lorikeet_b200-sys`, shaped after the two `gkl` seams Lorikeet uses
//! (`gkl::pairhmm::forward`, `src/pair_hmm/pair_hmm.rs:348-366`; `gkl::smithwaterman::align`,
//! `src/smith_waterman/smith_waterman_aligner.rs:56-61`) and after the callers either side of them.
//! Errors of the C ABI become `Err(Error)`; the reference `unwrap()`s / panics at the same places
//! (`pair_hmm.rs:348,372`, `smith_waterman_aligner.rs:56-60`), which an `AVXMode::Gpu` arm reproduces with `.unwrap()`.
//!
//! NOT compiled in this repository's build container (no Rust toolchain there); see `tests/test_bindings.py`.
use lorikeet_b200_sys as sys;
use std::ffi::CStr;
use std::ptr;

#[derive(Debug, Clone)]
pub struct Error {
    pub status: i32,
    pub message: String,
}
pub type Result<T> = std::result::Result<T, Error>;

/// `gkl::smithwaterman::OverhangStrategy`
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum OverhangStrategy {
    SoftClip = 0,
    InDel = 1,
    LeadingInDel = 2,
    Ignore = 3,
}

/// `gkl::smithwaterman::Parameters`
pub type Parameters = sys::lk_sw_params;
/// the named parameter sets of `smith_waterman_aligner.rs:12-25`
pub const ORIGINAL_DEFAULT: Parameters = Parameters { match_value: 3, mismatch_penalty: -1, gap_open_penalty: -4, gap_extend_penalty: -3 };
pub const STANDARD_NGS: Parameters = Parameters { match_value: 25, mismatch_penalty: -50, gap_open_penalty: -110, gap_extend_penalty: -6 };
pub const NEW_SW_PARAMETERS: Parameters = Parameters { match_value: 200, mismatch_penalty: -150, gap_open_penalty: -260, gap_extend_penalty: -11 };
pub const ALIGNMENT_TO_BEST_HAPLOTYPE_SW_PARAMETERS: Parameters = Parameters { match_value: 10, mismatch_penalty: -15, gap_open_penalty: -30, gap_extend_penalty: -5 };

/// `PCRErrorModel` (`pair_hmm_likelihood_calculation_engine.rs:60-70`)
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum PCRErrorModel {
    None = 0,
    Hostile = 1,
    Aggressive = 2,
    Conservative = 3,
}

/// Arguments of `PairHMMLikelihoodCalculationEngine::new` (`:129-141`), same order.
pub type EngineArgs = sys::lk_engine_args;

/// A batch of active regions in the packed layout that crosses the ABI (CSR offsets + byte arenas).
#[derive(Default, Clone)]
pub struct RegionBatch {
    pub region_read_start: Vec<i64>,
    pub region_hap_start: Vec<i64>,
    pub read_off: Vec<i64>,
    pub read_bases: Vec<u8>,
    pub read_qual: Vec<u8>,
    pub read_ins: Vec<u8>,
    pub read_del: Vec<u8>,
    pub read_gcp: Option<Vec<u8>>,
    pub gcp_constant: u8,
    pub hap_off: Vec<i64>,
    pub hap_bases: Vec<u8>,
    /// several samples per region (`None`: one each); see `lk_pairhmm_batch`
    pub samples: Option<(Vec<i64>, Vec<i64>)>,
}

impl RegionBatch {
    pub fn new(gcp_constant: u8) -> Self {
        RegionBatch { region_read_start: vec![0], region_hap_start: vec![0], read_off: vec![0], hap_off: vec![0], gcp_constant, ..Default::default() }
    }
    /// Appends one region: its haplotypes (insertion order of `AssemblyResultSet.haplotypes`) and, per sample, its reads as
    /// `(bases, quals, insertion quals, deletion quals)`.
    pub fn push_region(&mut self, haplotypes: &[&[u8]], samples: &[Vec<(&[u8], &[u8], &[u8], &[u8])>]) {
        for h in haplotypes {
            self.hap_bases.extend_from_slice(h);
            self.hap_off.push(self.hap_bases.len() as i64);
        }
        self.region_hap_start.push(self.hap_off.len() as i64 - 1);
        if samples.len() != 1 || self.samples.is_some() {
            let n_regions_before = self.region_read_start.len() - 1;
            let (rss, srs) = self.samples.get_or_insert_with(|| (vec![0], vec![0]));
            while rss.len() <= n_regions_before {
                // regions pushed before the first multi-sample one were one sample each
                let g = rss.len() - 1;
                srs.push(self.region_read_start[g + 1]);
                rss.push(srs.len() as i64 - 1);
            }
        }
        for reads in samples {
            for (b, q, i, d) in reads {
                assert!(b.len() == q.len() && b.len() == i.len() && b.len() == d.len() && !b.is_empty()); // pair_hmm.rs:417-440
                self.read_bases.extend_from_slice(b);
                self.read_qual.extend_from_slice(q);
                self.read_ins.extend_from_slice(i);
                self.read_del.extend_from_slice(d);
                self.read_off.push(self.read_bases.len() as i64);
            }
            if let Some((_, srs)) = self.samples.as_mut() {
                srs.push(self.read_off.len() as i64 - 1);
            }
        }
        if let Some((rss, srs)) = self.samples.as_mut() {
            rss.push(srs.len() as i64 - 1);
        }
        self.region_read_start.push(self.read_off.len() as i64 - 1);
    }
    pub fn n_regions(&self) -> usize {
        self.region_read_start.len() - 1
    }
    pub fn n_reads(&self) -> usize {
        *self.region_read_start.last().unwrap() as usize
    }
    pub fn n_pairs(&self) -> usize {
        (0..self.n_regions())
            .map(|g| ((self.region_read_start[g + 1] - self.region_read_start[g]) * (self.region_hap_start[g + 1] - self.region_hap_start[g])) as usize)
            .sum()
    }
    /// The borrowed view the C ABI takes; valid while `self` is not modified.
    pub fn as_raw(&self) -> sys::lk_pairhmm_batch {
        sys::lk_pairhmm_batch {
            n_regions: self.n_regions() as i64,
            region_read_start: self.region_read_start.as_ptr(),
            region_hap_start: self.region_hap_start.as_ptr(),
            read_off: self.read_off.as_ptr(),
            read_bases: self.read_bases.as_ptr(),
            read_qual: self.read_qual.as_ptr(),
            read_ins: self.read_ins.as_ptr(),
            read_del: self.read_del.as_ptr(),
            read_gcp: self.read_gcp.as_ref().map_or(ptr::null(), |g| g.as_ptr()),
            gcp_constant: self.gcp_constant,
            hap_off: self.hap_off.as_ptr(),
            hap_bases: self.hap_bases.as_ptr(),
            region_sample_start: self.samples.as_ref().map_or(ptr::null(), |s| s.0.as_ptr()),
            sample_read_start: self.samples.as_ref().map_or(ptr::null(), |s| s.1.as_ptr()),
        }
    }
}

/// One GPU.  Thread-safe (calls serialise on an internal lock); clone-free sharing through `&Context` / `Arc<Context>`.
pub struct Context {
    raw: *mut sys::lk_ctx,
    owned: bool,
}
unsafe impl Send for Context {}
unsafe impl Sync for Context {}

impl Context {
    /// `device < 0`: the current CUDA device.  Fails with `LK_ERR_CUDA` when there is no GPU: there is no CPU fallback.
    pub fn new(device: i32) -> Result<Context> {
        let mut raw = ptr::null_mut();
        let st = unsafe { sys::lk_create(device, &mut raw) };
        if st != sys::LK_OK {
            return Err(Error { status: st, message: "lk_create failed (no CUDA device?)".into() });
        }
        Ok(Context { raw, owned: true })
    }
    fn check(&self, st: sys::lk_status) -> Result<()> {
        if st == sys::LK_OK {
            return Ok(());
        }
        let message = unsafe { CStr::from_ptr(sys::lk_last_error(self.raw)) }.to_string_lossy().into_owned();
        Err(Error { status: st, message })
    }
    /// `PairHMM::do_not_use_tristate_correction` (`pair_hmm.rs:189-191`)
    pub fn do_not_use_tristate_correction(&self) -> Result<()> {
        self.check(unsafe { sys::lk_set_tristate_correction(self.raw, 0) })
    }
    /// The function `gkl::pairhmm::forward()` returns: `(hap, read, qual, ins, del, gcp) -> log10 likelihood`.
    pub fn forward(&self, hap: &[u8], read: &[u8], qual: &[u8], ins: &[u8], del: &[u8], gcp: &[u8]) -> Result<f64> {
        assert!(read.len() == qual.len() && read.len() == ins.len() && read.len() == del.len() && read.len() == gcp.len());
        let mut out = 0f64;
        self.check(unsafe {
            sys::lk_pairhmm_forward(self.raw, hap.as_ptr(), hap.len() as i64, read.as_ptr(), qual.as_ptr(), ins.as_ptr(), del.as_ptr(), gcp.as_ptr(), read.len() as i64, &mut out)
        })?;
        Ok(out)
    }
    /// `PairHMM::compute_likelihoods` for a whole batch: `m_log_likelihood_array[r * A + h]` per region (`pair_hmm.rs:349-369`).
    pub fn pairhmm_batch(&self, batch: &RegionBatch) -> Result<(Vec<f64>, Vec<u8>)> {
        let n = batch.n_pairs();
        let (mut out, mut flags) = (vec![0f64; n], vec![0u8; n]);
        let raw = batch.as_raw();
        self.check(unsafe { sys::lk_pairhmm_batch_host(self.raw, &raw, out.as_mut_ptr(), flags.as_mut_ptr()) })?;
        Ok((out, flags))
    }
    /// `PairHMMLikelihoodCalculationEngine::compute_read_likelihoods` (`…engine.rs:195-242`) for a batch of regions:
    /// per region (and sample) an `[A x R]` row-major block, normalised and filtered, plus the keep mask per read.
    pub fn compute_read_likelihoods(&self, args: &EngineArgs, batch: &RegionBatch, read_mapq: &[u8], ref_allele_index: Option<&[i32]>) -> Result<(Vec<f64>, Vec<u8>)> {
        assert_eq!(read_mapq.len(), batch.n_reads());
        let (mut values, mut keep) = (vec![0f64; batch.n_pairs()], vec![1u8; batch.n_reads()]);
        let raw = batch.as_raw();
        self.check(unsafe {
            sys::lk_engine_compute_read_likelihoods(self.raw, args, &raw, read_mapq.as_ptr(), ref_allele_index.map_or(ptr::null(), |r| r.as_ptr()), values.as_mut_ptr(), keep.as_mut_ptr())
        })?;
        Ok((values, keep))
    }
    /// The function `gkl::smithwaterman::align()` returns: ASCII CIGAR bytes (parsed with `CigarString::try_from`) + offset.
    pub fn align(&self, reference: &[u8], alternate: &[u8], parameters: Parameters, strategy: OverhangStrategy) -> Result<(Vec<u8>, i32)> {
        let mut text = vec![0u8; 16 * (reference.len() + alternate.len()) + 32];
        let mut offset = 0i32;
        self.check(unsafe {
            sys::lk_sw_align(self.raw, reference.as_ptr(), reference.len() as i64, alternate.as_ptr(), alternate.len() as i64, &parameters, strategy as i32,
                             text.as_mut_ptr() as *mut _, text.len(), &mut offset)
        })?;
        let n = text.iter().position(|&c| c == 0).unwrap_or(text.len());
        text.truncate(n);
        Ok((text, offset))
    }
    /// Batch of independent alignments; BAM-encoded CIGAR elements (`len << 4 | op`) per pair, `None` where the reference
    /// would assert (empty sequence) or the CIGAR exceeded `max_ops`.
    pub fn align_batch(&self, pairs: &[(&[u8], &[u8])], parameters: Parameters, strategy: OverhangStrategy, max_ops: usize) -> Result<Vec<Option<(Vec<u32>, i32)>>> {
        let (ro, refs, ao, alts) = pack_pairs(pairs);
        let n = pairs.len();
        let (mut ops, mut n_ops, mut offs) = (vec![0u32; n * max_ops], vec![0i32; n], vec![0i32; n]);
        self.check(unsafe {
            sys::lk_sw_batch_host(self.raw, n as i64, ro.as_ptr(), refs.as_ptr(), ao.as_ptr(), alts.as_ptr(), &parameters, strategy as i32, max_ops as i32,
                                  ops.as_mut_ptr(), n_ops.as_mut_ptr(), offs.as_mut_ptr())
        })?;
        Ok((0..n).map(|i| if n_ops[i] < 0 { None } else { Some((ops[i * max_ops..i * max_ops + n_ops[i] as usize].to_vec(), offs[i])) }).collect())
    }
    /// `CigarUtils::calculate_cigar` (`cigar_utils.rs:358-457`) for many (reference, haplotype) pairs in one launch.
    /// `None` where the reference returns `None` (`is_s_w_failure`).
    pub fn calculate_cigar_batch(&self, pairs: &[(&[u8], &[u8])], parameters: Parameters, strategy: OverhangStrategy, max_ops: usize) -> Result<Vec<Option<Vec<u32>>>> {
        let (ro, refs, ao, alts) = pack_pairs(pairs);
        let n = pairs.len();
        let (mut ops, mut n_ops) = (vec![0u32; n * max_ops], vec![0i32; n]);
        self.check(unsafe {
            sys::lk_calculate_cigar_batch(self.raw, n as i64, ro.as_ptr(), refs.as_ptr(), ao.as_ptr(), alts.as_ptr(), &parameters, strategy as i32, max_ops as i32,
                                          ops.as_mut_ptr(), n_ops.as_mut_ptr())
        })?;
        Ok((0..n).map(|i| if n_ops[i] < 0 { None } else { Some(ops[i * max_ops..i * max_ops + n_ops[i] as usize].to_vec()) }).collect())
    }
    /// The alignment step of the dangling-end merges (`read_threading_graph.rs:1408,1471`): LeadingInDel, no substring
    /// shortcut needed, only the CIGAR is used.
    pub fn dangling_end_sw_batch(&self, pairs: &[(&[u8], &[u8])], parameters: Parameters, max_ops: usize) -> Result<Vec<Option<Vec<u32>>>> {
        let (ro, refs, ao, alts) = pack_pairs(pairs);
        let n = pairs.len();
        let (mut ops, mut n_ops) = (vec![0u32; n * max_ops], vec![0i32; n]);
        self.check(unsafe {
            sys::lk_dangling_end_sw_batch(self.raw, n as i64, ro.as_ptr(), refs.as_ptr(), ao.as_ptr(), alts.as_ptr(), &parameters, max_ops as i32, ops.as_mut_ptr(), n_ops.as_mut_ptr())
        })?;
        Ok((0..n).map(|i| if n_ops[i] < 0 { None } else { Some(ops[i * max_ops..i * max_ops + n_ops[i] as usize].to_vec()) }).collect())
    }
    /// `realign_reads_to_their_best_haplotype` (`assembly_based_caller_utils.rs:208-246`): per read its best allele, whether it
    /// is informative, the new CIGAR (`None`: not realigned) and the new alignment start.
    ///
    /// # Safety
    /// `batch` holds raw pointers (see `lk_realign_batch`); they must be valid for the call.
    pub unsafe fn realign_reads_batch(&self, batch: &sys::lk_realign_batch, n_reads: usize, max_ops: usize) -> Result<(Vec<i32>, Vec<u8>, Vec<Option<Vec<u32>>>, Vec<i64>)> {
        let (mut best, mut info) = (vec![-1i32; n_reads], vec![0u8; n_reads]);
        let (mut ops, mut n_ops, mut pos) = (vec![0u32; n_reads * max_ops], vec![0i32; n_reads], vec![0i64; n_reads]);
        self.check(sys::lk_realign_reads_batch(self.raw, batch, max_ops as i32, best.as_mut_ptr(), info.as_mut_ptr(), ops.as_mut_ptr(), n_ops.as_mut_ptr(), pos.as_mut_ptr()))?;
        let cigars = (0..n_reads).map(|i| if n_ops[i] < 0 { None } else { Some(ops[i * max_ops..i * max_ops + n_ops[i] as usize].to_vec()) }).collect();
        Ok((best, info, cigars, pos))
    }
    pub fn raw(&self) -> *mut sys::lk_ctx {
        self.raw
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        if self.owned {
            unsafe { sys::lk_destroy(self.raw) }
        }
    }
}

/// `PairHMM` of `src/pair_hmm/pair_hmm.rs` with the GPU behind it: the same method names, argument meaning and asserts, so the
/// `AVXMode::Gpu` arm can stand in for the struct the engine owns today.  There is nothing to allocate on this path
/// (`initialize` / `quick_initialize` / `reinitialize` only book-keep the sizes, as the gkl arm does, `:63-207`); the
/// haplotype-prefix caching arguments are accepted and ignored exactly as the gkl arm ignores them (`:442-496`).
pub struct PairHMM<'a> {
    ctx: &'a Context,
    initialized: bool,
    max_read_length: usize,
    max_haplotype_length: usize,
    haplotypes: Vec<Vec<u8>>,
    log_likelihood_array: Vec<f64>,
}

impl<'a> PairHMM<'a> {
    /// `PairHMM::initialize` (`:63-127`): keeps the haplotypes in list order and the maximum lengths.
    pub fn initialize(ctx: &'a Context, haplotypes: &[&[u8]], max_read_length: usize) -> PairHMM<'a> {
        let max_haplotype_length = haplotypes.iter().map(|h| h.len()).max().unwrap_or(0);
        PairHMM { ctx, initialized: true, max_read_length, max_haplotype_length, haplotypes: haplotypes.iter().map(|h| h.to_vec()).collect(), log_likelihood_array: Vec::new() }
    }
    /// `PairHMM::quick_initialize` (`:130-168`): not initialised when both sizes are zero (`:138-141`).
    pub fn quick_initialize(ctx: &'a Context, max_read_length: usize, haplotype_max_length: usize) -> PairHMM<'a> {
        PairHMM { ctx, initialized: !(max_read_length == 0 && haplotype_max_length == 0), max_read_length, max_haplotype_length: haplotype_max_length, haplotypes: Vec::new(), log_likelihood_array: Vec::new() }
    }
    /// `PairHMM::reinitialize` (`:193-207`): sizes only, as the reference (its re-allocations are commented out).
    pub fn reinitialize(&mut self, max_read_length: usize, haplotype_max_length: usize) {
        self.max_read_length = max_read_length;
        self.max_haplotype_length = haplotype_max_length;
        self.initialized = true;
    }
    pub fn do_not_use_tristate_correction(&self) {
        self.ctx.do_not_use_tristate_correction().unwrap()
    }
    /// `PairHMM::compute_log10_likelihoods` (`:217-341`) for one region: `m_log_likelihood_array[r * A + h]`.
    pub fn compute_log10_likelihoods(&mut self, batch: &RegionBatch) -> &[f64] {
        assert!(self.initialized, "Must call initialize before calling compute_log10_likelihoods");
        self.log_likelihood_array = self.ctx.pairhmm_batch(batch).unwrap().0; // the reference unwraps at pair_hmm.rs:348,372
        &self.log_likelihood_array
    }
    /// `PairHMM::compute_read_likelihood_given_haplotype_log10` (`:404-500`), same asserts and messages.
    #[allow(clippy::too_many_arguments)]
    pub fn compute_read_likelihood_given_haplotype_log10(&self, haplotype_bases: &[u8], read_bases: &[u8], read_quals: &[u8], insertion_gop: &[u8],
                                                         deletion_gop: &[u8], overall_gcp: &[u8], _recache_read_values: bool, _next_haplotype_bases: Option<&[u8]>) -> f64 {
        assert!(self.initialized, "Must call initialize before calling compute_read_likelihood_given_haplotype_log10");
        assert!(haplotype_bases.len() <= self.max_haplotype_length, "Haplotype bases is too long");
        assert!(read_quals.len() == read_bases.len(), "Read bases and read quals aren't the same size");
        assert!(insertion_gop.len() == read_bases.len(), "Read bases and insertion gcp aren't the same size");
        assert!(deletion_gop.len() == read_bases.len(), "Read bases and deletion gcp aren't the same size");
        assert!(overall_gcp.len() == read_bases.len(), "Read bases and overal GCP aren't the same size");
        let result = self.ctx.forward(haplotype_bases, read_bases, read_quals, insertion_gop, deletion_gop, overall_gcp).unwrap();
        assert!(result <= 0.0, "PairHMM Log Probability cannot be greater than 0.0"); // :478-481
        result
    }
    /// `PairHMM::get_log_likelihood_array` (`:377-379`)
    pub fn get_log_likelihood_array(&self) -> &[f64] {
        &self.log_likelihood_array
    }
    pub fn haplotypes(&self) -> &[Vec<u8>] {
        &self.haplotypes
    }
    /// `PairHMM::find_first_position_where_haplotypes_differ` (`:706-717`)
    pub fn find_first_position_where_haplotypes_differ(h1: &[u8], h2: &[u8]) -> usize {
        h1.iter().zip(h2.iter()).position(|(a, b)| a != b).unwrap_or_else(|| h1.len().min(h2.len()))
    }
}

fn pack_pairs(pairs: &[(&[u8], &[u8])]) -> (Vec<i64>, Vec<u8>, Vec<i64>, Vec<u8>) {
    let (mut ro, mut refs, mut ao, mut alts) = (vec![0i64], Vec::new(), vec![0i64], Vec::new());
    for (r, a) in pairs {
        refs.extend_from_slice(r);
        ro.push(refs.len() as i64);
        alts.extend_from_slice(a);
        ao.push(alts.len() as i64);
    }
    (ro, refs, ao, alts)
}

/// Several GPUs behind one handle in ONE process: what a Lorikeet build links to reach every device of the box.
pub struct Pool {
    raw: *mut sys::lk_pool,
}
unsafe impl Send for Pool {}
unsafe impl Sync for Pool {}

impl Pool {
    /// `devices = None`: every visible device.
    pub fn new(devices: Option<&[i32]>) -> Result<Pool> {
        let mut raw = ptr::null_mut();
        let st = unsafe { sys::lk_pool_create(devices.map_or(ptr::null(), |d| d.as_ptr()), devices.map_or(0, |d| d.len() as i32), &mut raw) };
        if st != sys::LK_OK {
            return Err(Error { status: st, message: "lk_pool_create failed".into() });
        }
        Ok(Pool { raw })
    }
    pub fn device_count(&self) -> usize {
        unsafe { sys::lk_pool_device_count(self.raw) as usize }
    }
    /// The i-th device's context, borrowed from the pool.
    pub fn context(&self, i: usize) -> Context {
        Context { raw: unsafe { sys::lk_pool_ctx(self.raw, i as i32) }, owned: false }
    }
    fn check(&self, st: sys::lk_status) -> Result<()> {
        if st == sys::LK_OK {
            return Ok(());
        }
        Err(Error { status: st, message: unsafe { CStr::from_ptr(sys::lk_pool_last_error(self.raw)) }.to_string_lossy().into_owned() })
    }
    pub fn do_not_use_tristate_correction(&self) -> Result<()> {
        self.check(unsafe { sys::lk_pool_set_tristate_correction(self.raw, 0) })
    }
    pub fn pairhmm_batch(&self, batch: &RegionBatch) -> Result<(Vec<f64>, Vec<u8>)> {
        let n = batch.n_pairs();
        let (mut out, mut flags) = (vec![0f64; n], vec![0u8; n]);
        let raw = batch.as_raw();
        self.check(unsafe { sys::lk_pool_pairhmm_batch_host(self.raw, &raw, out.as_mut_ptr(), flags.as_mut_ptr()) })?;
        Ok((out, flags))
    }
    pub fn compute_read_likelihoods(&self, args: &EngineArgs, batch: &RegionBatch, read_mapq: &[u8], ref_allele_index: Option<&[i32]>) -> Result<(Vec<f64>, Vec<u8>)> {
        let (mut values, mut keep) = (vec![0f64; batch.n_pairs()], vec![1u8; batch.n_reads()]);
        let raw = batch.as_raw();
        self.check(unsafe {
            sys::lk_pool_compute_read_likelihoods(self.raw, args, &raw, read_mapq.as_ptr(), ref_allele_index.map_or(ptr::null(), |r| r.as_ptr()), values.as_mut_ptr(), keep.as_mut_ptr())
        })?;
        Ok((values, keep))
    }
    /// The cross-thread batcher over every device of the pool (`args = None`: bare PairHMM calls).
    pub fn batcher(&self, args: Option<&EngineArgs>, max_regions: i32, max_wait_us: i32) -> Result<Batcher> {
        let mut raw = ptr::null_mut();
        self.check(unsafe { sys::lk_pool_batcher_create(self.raw, args.map_or(ptr::null(), |a| a as *const _), max_regions, max_wait_us, &mut raw) })?;
        Ok(Batcher { raw, engine: args.is_some() })
    }
}

impl Drop for Pool {
    fn drop(&mut self) {
        unsafe { sys::lk_pool_destroy(self.raw) }
    }
}

/// Front-end for Lorikeet's call pattern: Rayon workers call the engine with ONE region each
/// (`assembly_region_walker.rs:140-277`); the batcher merges what is pending from all threads into few launches.
/// Must be dropped before the `Context` / `Pool` it was created from.
pub struct Batcher {
    raw: *mut sys::lk_batcher,
    engine: bool,
}
unsafe impl Send for Batcher {}
unsafe impl Sync for Batcher {}

/// A region queued with `submit_async`: owns everything the library reads and writes until `wait`.
pub struct Pending {
    ticket: *mut sys::lk_ticket,
    _region: Box<RegionBatch>,
    _desc: Box<sys::lk_pairhmm_batch>, // stable address: the dispatcher reads it when it gathers the launch
    mapq: Vec<u8>,
    values: Vec<f64>,
    aux: Vec<u8>,
}
unsafe impl Send for Pending {}

impl Batcher {
    pub fn new(ctx: &Context, args: Option<&EngineArgs>, max_regions: i32, max_wait_us: i32) -> Result<Batcher> {
        let mut raw = ptr::null_mut();
        ctx.check(unsafe { sys::lk_batcher_create(ctx.raw, args.map_or(ptr::null(), |a| a as *const _), max_regions, max_wait_us, &mut raw) })?;
        Ok(Batcher { raw, engine: args.is_some() })
    }
    /// Blocks the calling worker until its region is done: `(values, keep)` in engine mode, `(log10 likelihoods, flags)` otherwise.
    pub fn submit(&self, region: RegionBatch, read_mapq: Vec<u8>, ref_allele_index: i32) -> Result<(Vec<f64>, Vec<u8>)> {
        let p = self.submit_async(region, read_mapq, ref_allele_index)?;
        self.wait(p)
    }
    /// Queues a region and returns at once, so one worker can keep several regions in flight.
    pub fn submit_async(&self, region: RegionBatch, read_mapq: Vec<u8>, ref_allele_index: i32) -> Result<Pending> {
        assert_eq!(region.n_regions(), 1);
        let n_aux = if self.engine { region.n_reads() } else { region.n_pairs() };
        let n_pairs = region.n_pairs();
        let region = Box::new(region);
        let desc = Box::new(region.as_raw());
        let mut p = Pending { ticket: ptr::null_mut(), _region: region, _desc: desc, mapq: read_mapq, values: vec![0f64; n_pairs], aux: vec![0u8; n_aux] };
        let st = unsafe {
            sys::lk_batcher_submit_async(self.raw, &*p._desc, if p.mapq.is_empty() { ptr::null() } else { p.mapq.as_ptr() }, ref_allele_index, p.values.as_mut_ptr(), p.aux.as_mut_ptr(), &mut p.ticket)
        };
        if st != sys::LK_OK {
            return Err(Error { status: st, message: "lk_batcher_submit_async rejected the region".into() });
        }
        Ok(p)
    }
    /// Every `Pending` must be waited on exactly once, before the batcher is dropped.
    pub fn wait(&self, p: Pending) -> Result<(Vec<f64>, Vec<u8>)> {
        let st = unsafe { sys::lk_batcher_wait(self.raw, p.ticket) };
        if st != sys::LK_OK {
            return Err(Error { status: st, message: "batched launch failed".into() });
        }
        Ok((p.values, p.aux))
    }
    pub fn stats(&self) -> (i64, i64) {
        let (mut b, mut r) = (0i64, 0i64);
        unsafe { sys::lk_batcher_stats(self.raw, &mut b, &mut r) };
        (b, r)
    }
}

impl Drop for Batcher {
    fn drop(&mut self) {
        unsafe { sys::lk_batcher_destroy(self.raw) }
    }
}
