// host_cigar.cpp — CigarUtils::calculate_cigar around the batched Smith-Waterman kernel.
//
// Reference:
//   CigarUtils::calculate_cigar / is_s_w_failure      src/reads/cigar_utils.rs:358-487 (SW_PAD :10)
//   AlignmentUtils::trim_cigar_by_bases / trim_cigar   src/reads/alignment_utils.rs:321-386
//   AlignmentUtils::left_align_indels                  src/reads/alignment_utils.rs:425-555
//   AlignmentUtils::normalize_alleles                  src/reads/alignment_utils.rs:583-680
//   CigarBuilder                                       src/reads/cigar_builder.rs
//
// The O(n*m) part (one SW per haplotype, <= 128 per region: read_threading_assembler.rs:787,827) runs on the GPU as
// ONE lk_sw batch; what remains here is O(n+m) CIGAR surgery per pair, kept on the host like the reference.
// CIGAR elements are BAM-encoded u32 (len << 4 | op; M=0 I=1 D=2 N=3 S=4 H=5 P=6 '='=7 X=8).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/lorikeet_b200.h"
#include "best_allele.h"
#include "host_parallel.h"
#include "lk_api_internal.h"

namespace {

enum Op : uint32_t { M = 0, I = 1, D = 2, N = 3, S = 4, H = 5, P = 6, EQ = 7, X = 8 };
constexpr int kSwPad = 10;  // cigar_utils.rs:10

inline uint32_t el(uint32_t op, uint32_t len) { return (len << 4) | op; }
inline uint32_t op_of(uint32_t e) { return e & 0xF; }
inline uint32_t len_of(uint32_t e) { return e >> 4; }
inline bool is_clipping(uint32_t e) { return op_of(e) == S || op_of(e) == H; }
inline bool is_indel(uint32_t e) { return op_of(e) == I || op_of(e) == D; }
inline bool is_alignment(uint32_t e) { return op_of(e) == M || op_of(e) == EQ || op_of(e) == X; }
inline bool consumes_read(uint32_t e) { const uint32_t o = op_of(e); return o == M || o == I || o == S || o == EQ || o == X; }
inline bool consumes_ref(uint32_t e) { const uint32_t o = op_of(e); return o == M || o == D || o == N || o == EQ || o == X; }
inline uint32_t len_on_read(uint32_t e) { return consumes_read(e) ? len_of(e) : 0; }
inline uint32_t len_on_ref(uint32_t e) { return consumes_ref(e) ? len_of(e) : 0; }

struct BuilderResult {
    std::vector<uint32_t> cigar;
    uint32_t leading_deletion_bases_removed = 0, trailing_deletion_bases_removed = 0;
    bool ok = true;
};

// CigarBuilder (cigar_builder.rs), remove_deletions_at_ends as a constructor flag
class CigarBuilder {
  public:
    explicit CigarBuilder(bool remove_deletions_at_ends) : remove_(remove_deletions_at_ends) {}

    bool add(uint32_t e) {
        if (len_of(e) == 0) return true;
        if (remove_ && op_of(e) == D && leading_position()) {
            leading_removed_ += len_of(e);
            return true;
        }
        if (!advance_section(e)) return false;
        if (has_last_ && op_of(e) == op_of(last_)) {  // same type as the last operator: merge into the last element
            // (combine_cigar_operators(..).unwrap_or(old): if the vector's tail is of another type the element is dropped)
            if (op_of(els_.back()) == op_of(e)) els_.back() = el(op_of(e), len_of(e) + len_of(els_.back()));
            return true;
        }
        if (!has_last_) { push(e); return true; }
        if (is_clipping(e)) {
            const size_t n = els_.size();
            if (remove_ && !consumes_read(last_) && !is_clipping(last_)) {  // trailing deletion before the clip
                trailing_removed_ += len_of(els_[n - 1]);
                els_[n - 1] = e;
                last_ = e;
            } else if (remove_ && last_two_del_ins()) {
                trailing_removed_ += len_of(els_[n - 2]);
                els_[n - 2] = els_[n - 1];
                els_[n - 1] = e;  // last_operator deliberately left as is (as in the reference)
            } else {
                push(e);
            }
            return true;
        }
        if (op_of(e) == D && op_of(last_) == I) {  // keep D before I
            const size_t n = els_.size();
            if (n > 1 && op_of(els_[n - 2]) == D) els_[n - 2] = el(D, len_of(els_[n - 2]) + len_of(e));
            else els_.insert(els_.begin() + (long)(n - 1), e);
            return true;
        }
        push(e);
        return true;
    }

    BuilderResult make_and_record_deletions_removed_result() {
        BuilderResult r;
        r.leading_deletion_bases_removed = leading_removed_;
        if (error_ || !has_last_) { r.ok = false; return r; }
        if (section_ == LeftSoftClip && op_of(els_[0]) == S) { r.ok = false; return r; }  // completely soft clipped
        uint32_t in_make = 0;
        if (remove_ && op_of(last_) == D) {
            in_make = len_of(els_.back());
            els_.pop_back();
        } else if (remove_ && last_two_del_ins()) {
            in_make = len_of(els_[els_.size() - 2]);
            els_.erase(els_.end() - 2);
        }
        if (els_.empty()) { r.ok = false; return r; }  // make(allow_empty = false)
        r.cigar = els_;
        r.trailing_deletion_bases_removed = trailing_removed_ + in_make;
        return r;
    }

  private:
    enum Section { LeftHardClip, LeftSoftClip, Middle, RightSoftClip, RightHardClip };
    bool remove_;
    std::vector<uint32_t> els_;
    uint32_t last_ = 0;
    bool has_last_ = false, error_ = false;
    Section section_ = LeftHardClip;
    uint32_t leading_removed_ = 0, trailing_removed_ = 0;

    void push(uint32_t e) { els_.push_back(e); last_ = e; has_last_ = true; }
    bool leading_position() const {  // a deletion here would lead the alignment
        if (!has_last_) return true;
        if (is_clipping(last_)) return true;
        if (op_of(last_) == I) return els_.size() == 1 || is_clipping(els_[els_.size() - 2]);
        return false;
    }
    bool last_two_del_ins() const {
        return has_last_ && els_.size() > 1 && op_of(last_) == I && op_of(els_[els_.size() - 2]) == D;
    }
    bool advance_section(uint32_t e) {
        if (op_of(e) == H) {
            if (section_ == LeftSoftClip || section_ == Middle || section_ == RightSoftClip) section_ = RightHardClip;
        } else if (op_of(e) == S) {
            if (section_ == RightHardClip) { error_ = true; return false; }
            if (section_ == LeftHardClip) section_ = LeftSoftClip;
            else if (section_ == Middle) section_ = RightSoftClip;
        } else {
            if (section_ == RightSoftClip || section_ == RightHardClip) { error_ = true; return false; }
            if (section_ == LeftHardClip || section_ == LeftSoftClip) section_ = Middle;
        }
        return true;
    }
};

// AlignmentUtils::trim_cigar(cigar, start, end, by_reference = false)  (alignment_utils.rs:334-386)
BuilderResult trim_cigar_by_bases(const std::vector<uint32_t> &cigar, uint32_t start, uint32_t end) {
    BuilderResult bad;
    bad.ok = false;
    if (end < start) return bad;
    CigarBuilder nb(true);
    uint32_t element_start, element_end = 0;
    for (uint32_t e : cigar) {
        element_start = element_end;
        element_end = element_start + len_on_read(e);
        if (element_end < start || (element_end == start && element_start < start)) continue;
        if (element_start > end && element_end > end + 1) break;
        const uint32_t overlap = (element_end == element_start) ? len_of(e)
                                                                 : std::min(end + 1, element_end) - std::max(start, element_start);
        if (!nb.add(el(op_of(e), overlap))) return bad;
    }
    if (element_end < end) return bad;  // "Cigar elements don't reach end position"
    return nb.make_and_record_deletions_removed_result();
}

struct Range { int start, end; int len() const { return end > start ? end - start : 0; } };

// AlignmentUtils::normalize_alleles for two sequences, trim = true (alignment_utils.rs:583-640)
bool normalize_alleles(const uint8_t *s0, int n0, const uint8_t *s1, int n1, Range &b0, Range &b1, int max_shift, int &start_shift,
                       int &end_shift) {
    if (max_shift > b0.start || max_shift > b1.start) return false;
    auto at = [&](const uint8_t *s, int n, int i, bool &ok) -> uint8_t { if (i < 0 || i >= n) { ok = false; return 0; } return s[i]; };
    bool ok = true;
    start_shift = 0; end_shift = 0;
    int min_size = std::min(b0.len(), b1.len());
    while (min_size > 0 && at(s0, n0, b0.end - 1, ok) == at(s1, n1, b1.end - 1, ok) && ok) {
        b0.end--; b1.end--; min_size--; end_shift++;
    }
    while (ok && min_size > 0 && at(s0, n0, b0.start, ok) == at(s1, n1, b1.start, ok) && ok) {
        b0.start++; b1.start++; min_size--; start_shift--;
    }
    while (ok && start_shift < max_shift && at(s0, n0, b0.start - 1, ok) == at(s1, n1, b1.start - 1, ok) && ok &&
           at(s0, n0, b0.end - 1, ok) == at(s1, n1, b1.end - 1, ok) && ok) {
        b0.start--; b0.end--; b1.start--; b1.end--; start_shift++; end_shift++;
    }
    return ok;  // an out-of-range access is a panic in the reference
}

// AlignmentUtils::left_align_indels (alignment_utils.rs:425-555)
BuilderResult left_align_indels(const std::vector<uint32_t> &cigar, const uint8_t *ref, int ref_len, const uint8_t *read, int read_len,
                                uint32_t read_start) {
    BuilderResult r;
    bool any_indel = false;
    for (uint32_t e : cigar) any_indel |= is_indel(e);
    if (!any_indel) { r.cigar = cigar; return r; }
    BuilderResult bad;
    bad.ok = false;
    int last_indel = 0;
    for (int n = 0; n < (int)cigar.size(); n++) if (is_indel(cigar[(size_t)n])) last_indel = n;
    uint32_t necessary_ref = read_start;
    for (int n = 0; n <= last_indel; n++) necessary_ref += len_on_ref(cigar[(size_t)n]);
    if (necessary_ref > (uint32_t)ref_len) return bad;  // "Read goes past end of reference"
    std::vector<uint32_t> rtl;
    uint32_t ref_length = 0;
    for (uint32_t e : cigar) ref_length += len_on_ref(e);
    Range ref_rng{ (int)(read_start + ref_length), (int)(read_start + ref_length) };
    Range read_rng{ read_len, read_len };
    for (int n = (int)cigar.size() - 1; n >= 0; n--) {
        const uint32_t e = cigar[(size_t)n];
        if (is_indel(e)) {
            ref_rng.start -= (int)len_on_ref(e);
            read_rng.start -= (int)len_on_read(e);
        } else if (ref_rng.len() == 0 && read_rng.len() == 0) {
            ref_rng.start -= (int)len_on_ref(e); read_rng.start -= (int)len_on_read(e);
            ref_rng.end -= (int)len_on_ref(e); read_rng.end -= (int)len_on_read(e);
            rtl.push_back(e);
        } else {
            const int max_shift = is_alignment(e) ? (int)len_of(e) : 0;
            int s0, s1;
            if (!normalize_alleles(ref, ref_len, read, read_len, ref_rng, read_rng, max_shift, s0, s1)) return bad;
            rtl.push_back(el(M, (uint32_t)s1));
            const bool emit_indel = n == 0 || s0 < max_shift || !is_alignment(e);
            const int new_match_left = s0 < 0 ? -s0 : 0;
            const int remaining_left = s0 < 0 ? (int)len_of(e) : (int)len_of(e) - s0;
            if (emit_indel) {
                rtl.push_back(el(D, (uint32_t)ref_rng.len()));
                rtl.push_back(el(I, (uint32_t)read_rng.len()));
                ref_rng.end -= ref_rng.len();
                read_rng.end -= read_rng.len();
                const int dref = new_match_left + (consumes_ref(e) ? remaining_left : 0);
                const int dread = new_match_left + (consumes_read(e) ? remaining_left : 0);
                ref_rng.start -= dref; ref_rng.end -= dref;
                read_rng.start -= dread; read_rng.end -= dread;
            }
            rtl.push_back(el(M, (uint32_t)new_match_left));
            rtl.push_back(el(op_of(e), (uint32_t)remaining_left));
        }
    }
    rtl.push_back(el(D, (uint32_t)ref_rng.len()));
    rtl.push_back(el(I, (uint32_t)read_rng.len()));
    if (read_rng.start != 0) return bad;  // "Given cigar does not account for all bases of the read"
    CigarBuilder cb(true);
    for (auto it = rtl.rbegin(); it != rtl.rend(); ++it)
        if (!cb.add(*it)) return bad;
    return cb.make_and_record_deletions_removed_result();
}

int write_ops(const std::vector<uint32_t> &c, int32_t max_ops, uint32_t *out) {
    if ((int)c.size() > max_ops) return -2;
    for (size_t i = 0; i < c.size(); i++) out[i] = c[i];
    return (int)c.size();
}

}  // namespace

// First half of calculate_cigar (cigar_utils.rs:365-385): cases that need no alignment.  Returns 1 and fills the
// result if resolved, 0 if the padded pair has to be aligned.
extern "C" int lk_calculate_cigar_shortcut(const uint8_t *ref, int64_t ref_len, const uint8_t *alt, int64_t alt_len, int32_t max_ops,
                                           uint32_t *out_ops, int32_t *out_n_ops) {
    if (alt_len == 0) {  // "horrible edge case from the unit tests, where this path has no bases"
        if (max_ops < 1) { *out_n_ops = -2; return 1; }
        out_ops[0] = el(D, (uint32_t)ref_len);
        *out_n_ops = 1;
        return 1;
    }
    if (alt_len == ref_len) {
        int64_t mism = 0;
        for (int64_t i = 0; i < ref_len && mism <= 2; i++) mism += alt[i] != ref[i];
        if (mism <= 2) {
            if (max_ops < 1) { *out_n_ops = -2; return 1; }
            out_ops[0] = el(M, (uint32_t)ref_len);
            *out_n_ops = 1;
            return 1;
        }
    }
    return 0;
}

// Second half (cigar_utils.rs:399-457): from the SW result on the 10xN-padded pair to the final CIGAR.
// out_n_ops: >= 0 number of elements; -2 capacity; -3 the reference returns None (is_s_w_failure); -4 the reference would panic.
extern "C" lk_status lk_calculate_cigar_finish(const uint8_t *ref, int64_t ref_len, const uint8_t *alt, int64_t alt_len,
                                               const uint32_t *sw_ops, int32_t sw_n_ops, int32_t sw_offset, int32_t max_ops,
                                               uint32_t *out_ops, int32_t *out_n_ops) {
    if (!out_n_ops || (sw_n_ops > 0 && !sw_ops) || !out_ops) return LK_ERR_INVALID;
    if (sw_n_ops < 0) { *out_n_ops = -4; return LK_OK; }
    // is_s_w_failure (:469-487)
    if (sw_offset > 0) { *out_n_ops = -3; return LK_OK; }
    for (int i = 0; i < sw_n_ops; i++)
        if (op_of(sw_ops[i]) == S) { *out_n_ops = -3; return LK_OK; }
    const std::vector<uint32_t> aligned(sw_ops, sw_ops + sw_n_ops);
    const int64_t padded_path_len = alt_len + 2 * kSwPad;
    BuilderResult trimmed = trim_cigar_by_bases(aligned, (uint32_t)kSwPad, (uint32_t)(padded_path_len - kSwPad - 1));
    if (!trimmed.ok) { *out_n_ops = -4; return LK_OK; }
    std::vector<uint32_t> non_standard = trimmed.cigar;
    if (trimmed.trailing_deletion_bases_removed > 0) non_standard.push_back(el(D, trimmed.trailing_deletion_bases_removed));
    BuilderResult la = left_align_indels(non_standard, ref, (int)ref_len, alt, (int)alt_len, trimmed.leading_deletion_bases_removed);
    if (!la.ok) { *out_n_ops = -4; return LK_OK; }
    const uint32_t lead = trimmed.leading_deletion_bases_removed + la.leading_deletion_bases_removed;
    const uint32_t trail = la.trailing_deletion_bases_removed;
    std::vector<uint32_t> result;
    if (lead > 0) result.push_back(el(D, lead));
    result.insert(result.end(), la.cigar.begin(), la.cigar.end());
    if (trail > 0) result.push_back(el(D, trail));
    *out_n_ops = write_ops(result, max_ops, out_ops);
    return LK_OK;
}

// CigarUtils::calculate_cigar for a batch of (reference, alternate) pairs: shortcuts on the host, every remaining pair
// padded with 10 x 'N' and aligned in ONE lk_sw_batch_host call, CIGAR surgery on the host.
namespace {
// The context's page-locked staging for one assembled SW batch (see lk_ctx::stage_*): held for the duration of the batch.
struct SwStage {
    std::unique_lock<std::mutex> lock;
    uint8_t *refs = nullptr, *alts = nullptr;
    uint32_t *ops = nullptr;
    int32_t *n = nullptr, *off = nullptr;
    bool reserve(lk_ctx *ctx, size_t ref_bytes, size_t alt_bytes, size_t pairs, size_t max_ops) {
        lock = std::unique_lock<std::mutex>(ctx->stage_mu);
        cudaSetDevice(ctx->device);
        const bool ok = ctx->stage_refs.reserve(std::max<size_t>(ref_bytes, 16)) == cudaSuccess &&
                        ctx->stage_alts.reserve(std::max<size_t>(alt_bytes, 16)) == cudaSuccess &&
                        ctx->stage_ops.reserve(std::max<size_t>(pairs * max_ops * 4, 16)) == cudaSuccess &&
                        ctx->stage_n.reserve(std::max<size_t>(pairs * 4, 16)) == cudaSuccess &&
                        ctx->stage_off.reserve(std::max<size_t>(pairs * 4, 16)) == cudaSuccess;
        if (!ok) { (void)cudaGetLastError(); return false; }
        refs = (uint8_t *)ctx->stage_refs.p; alts = (uint8_t *)ctx->stage_alts.p; ops = (uint32_t *)ctx->stage_ops.p;
        n = (int32_t *)ctx->stage_n.p; off = (int32_t *)ctx->stage_off.p;
        return true;
    }
};
}  // namespace

extern "C" lk_status lk_calculate_cigar_batch(lk_ctx *ctx, int64_t n_pairs, const int64_t *ref_off, const uint8_t *refs,
                                              const int64_t *alt_off, const uint8_t *alts, const lk_sw_params *params,
                                              lk_sw_overhang strategy, int32_t max_ops, uint32_t *cigar_ops, int32_t *n_ops) {
    LK_ABI_GUARD_BEGIN
    if (!ctx) return LK_ERR_INVALID;
    if (n_pairs < 0 || !params || max_ops < 1) return lk_fail(ctx, LK_ERR_INVALID, "bad calculate_cigar arguments");
    if (n_pairs == 0) return LK_OK;
    if (!ref_off || !alt_off || !cigar_ops || !n_ops) return lk_fail(ctx, LK_ERR_INVALID, "null calculate_cigar arrays");
    for (int64_t p = 0; p < n_pairs; p++)
        if (ref_off[p + 1] < ref_off[p] || alt_off[p + 1] < alt_off[p]) return lk_fail(ctx, LK_ERR_INVALID, "decreasing offsets");
    // the per-pair host work (shortcuts, padding, CIGAR surgery) is spread over the host cores for large batches
    constexpr int64_t kGrain = 2048;
    std::vector<uint8_t> need((size_t)n_pairs, 0);
    lk::parallel_for(n_pairs, kGrain, [&](int64_t p_begin, int64_t p_end) {
        for (int64_t p = p_begin; p < p_end; p++)
            need[(size_t)p] = !lk_calculate_cigar_shortcut(refs + ref_off[p], ref_off[p + 1] - ref_off[p], alts + alt_off[p],
                                                           alt_off[p + 1] - alt_off[p], max_ops, cigar_ops + p * max_ops, &n_ops[p]);
    });
    std::vector<int64_t> todo;
    for (int64_t p = 0; p < n_pairs; p++)
        if (need[(size_t)p]) todo.push_back(p);
    if (todo.empty()) return LK_OK;
    // padded copies (cigar_utils.rs:387-398)
    const int64_t T = (int64_t)todo.size();
    std::vector<int64_t> pro((size_t)T + 1, 0), pao((size_t)T + 1, 0);
    for (int64_t i = 0; i < T; i++) {
        const int64_t p = todo[(size_t)i];
        pro[(size_t)i + 1] = pro[(size_t)i] + (ref_off[p + 1] - ref_off[p]) + 2 * kSwPad;
        pao[(size_t)i + 1] = pao[(size_t)i] + (alt_off[p + 1] - alt_off[p]) + 2 * kSwPad;
    }
    const int64_t sw_max_ops = std::max<int64_t>(8, (int64_t)max_ops + 8);
    SwStage stage;
    if (!stage.reserve(ctx, (size_t)pro.back(), (size_t)pao.back(), (size_t)T, (size_t)sw_max_ops)) return lk_fail(ctx, LK_ERR_NOMEM, "out of page-locked host memory");
    uint8_t *const prefs = stage.refs, *const palts = stage.alts;
    uint32_t *const sw_ops = stage.ops;
    int32_t *const sw_n = stage.n, *const sw_off = stage.off;
    lk::parallel_for(T, kGrain, [&](int64_t i_begin, int64_t i_end) {
        for (int64_t i = i_begin; i < i_end; i++) {
            const int64_t p = todo[(size_t)i];
            const int64_t rl = ref_off[p + 1] - ref_off[p], al = alt_off[p + 1] - alt_off[p];
            uint8_t *r = prefs + pro[(size_t)i], *a = palts + pao[(size_t)i];
            memset(r, 'N', (size_t)kSwPad); memcpy(r + kSwPad, refs + ref_off[p], (size_t)rl); memset(r + kSwPad + rl, 'N', (size_t)kSwPad);
            memset(a, 'N', (size_t)kSwPad); memcpy(a + kSwPad, alts + alt_off[p], (size_t)al); memset(a + kSwPad + al, 'N', (size_t)kSwPad);
        }
    });
    lk_status st = lk_sw_batch_host(ctx, T, pro.data(), prefs, pao.data(), palts, params, strategy, (int32_t)sw_max_ops, sw_ops, sw_n, sw_off);
    if (st != LK_OK) return st;
    std::atomic<int> failed{ (int)LK_OK };
    lk::parallel_for(T, kGrain, [&](int64_t i_begin, int64_t i_end) {
        for (int64_t i = i_begin; i < i_end; i++) {
            const int64_t p = todo[(size_t)i];
            if (sw_n[(size_t)i] == -2) { n_ops[p] = -2; continue; }
            const lk_status fs = lk_calculate_cigar_finish(refs + ref_off[p], ref_off[p + 1] - ref_off[p], alts + alt_off[p],
                                                           alt_off[p + 1] - alt_off[p], sw_ops + (size_t)i * (size_t)sw_max_ops,
                                                           sw_n[(size_t)i], sw_off[(size_t)i], max_ops, cigar_ops + p * max_ops, &n_ops[p]);
            if (fs != LK_OK) failed.store((int)fs);
        }
    });
    if (failed.load() != (int)LK_OK) return lk_fail(ctx, (lk_status)failed.load(), "calculate_cigar_finish failed");
    return LK_OK;
    LK_ABI_GUARD_END(return lk_fail(ctx, LK_ERR_NOMEM, "out of host memory"), return lk_fail(ctx, LK_ERR_INVALID, "internal error"))
}

// =================================================================================================
// N2: AssemblyBasedCallerUtils::realign_reads_to_their_best_haplotype
//   best allele per read with tie-breaking   src/model/allele_likelihoods.rs:457-554, :1043-1096
//   tie-breaking priority                    src/assembly/assembly_based_caller_utils.rs:187-195
//   create_read_aligned_to_ref               src/reads/alignment_utils.rs:40-164
//   append_clipped_elements_from_cigar...    :173-205      apply_cigar_to_cigar  :240-281 (+ CigarPairTransform :969-1063)
//   read_start_on_reference_haplotype        :283-312      Haplotype::get_consolidated_padded_cigar  haplotype.rs:248-256
// The O(R*H) part -- one Smith-Waterman per read against its best haplotype, SoftClip,
// ALIGNMENT_TO_BEST_HAPLOTYPE_SW_PARAMETERS (10, -15, -30, -5) -- is ONE lk_sw batch on the GPU; best-allele
// selection and the O(R+H) CIGAR composition stay on the host like the reference.
// =================================================================================================
namespace {

constexpr uint32_t kHapPad = 1000;  // get_consolidated_padded_cigar(1000), alignment_utils.rs:89

// CigarBuilder::new(true).add_all(c).make(false)
BuilderResult consolidate(const std::vector<uint32_t> &c) {
    CigarBuilder b(true);
    for (uint32_t e : c)
        if (!b.add(e)) { BuilderResult bad; bad.ok = false; return bad; }
    return b.make_and_record_deletions_removed_result();
}

uint32_t read_length_of(const std::vector<uint32_t> &c) {
    uint32_t n = 0;
    for (uint32_t e : c) n += len_on_read(e);
    return n;
}

// alignment_utils.rs:283-312; ok = false where the reference panics ("Cigar doesn't reach the read start")
uint32_t read_start_on_reference_haplotype(const std::vector<uint32_t> &hap_vs_ref, uint32_t read_start_on_hap, bool &ok) {
    ok = true;
    if (read_start_on_hap == 0) return 0;
    uint32_t ref_consumed = 0, hap_consumed = 0;
    for (uint32_t e : hap_vs_ref) {
        ref_consumed += len_on_ref(e);
        hap_consumed += len_on_read(e);
        if (hap_consumed >= read_start_on_hap) {
            const uint32_t excess = consumes_ref(e) ? hap_consumed - read_start_on_hap : 0;
            return ref_consumed > excess ? ref_consumed - excess : 0;  // saturating_sub
        }
    }
    ok = false;
    return 0;
}

// CigarPairTransform::new (alignment_utils.rs:969-1063): op13 (or none), advance12, advance23; false = panic
bool pair_transform(uint32_t op12, uint32_t op23, int &op13, uint32_t &adv12, uint32_t &adv23) {
    const bool m12 = op12 == M || op12 == EQ || op12 == X, i12 = op12 == I || op12 == S, d12 = op12 == D;
    const bool m23 = op23 == M || op23 == EQ || op23 == X, i23 = op23 == I || op23 == S, d23 = op23 == D;
    if (!(m23 || i23 || d23)) return false;
    if (m12) {
        if (m23) { op13 = M; adv12 = 1; adv23 = 1; }
        else if (i23) { op13 = I; adv12 = 1; adv23 = 1; }
        else { op13 = D; adv12 = 0; adv23 = 1; }
        return true;
    }
    if (i12) { op13 = I; adv12 = 1; adv23 = 0; return true; }
    if (d12) {
        if (m23) { op13 = D; adv12 = 1; adv23 = 1; }
        else if (i23) { op13 = -1; adv12 = 1; adv23 = 1; }
        else { op13 = D; adv12 = 0; adv23 = 1; }
        return true;
    }
    return false;
}

// alignment_utils.rs:240-281
BuilderResult apply_cigar_to_cigar(const std::vector<uint32_t> &c12, const std::vector<uint32_t> &c23) {
    BuilderResult bad;
    bad.ok = false;
    CigarBuilder nb(true);
    size_t i12 = 0, i23 = 0;
    uint32_t e12 = 0, e23 = 0;
    while (i12 < c12.size() && i23 < c23.size()) {
        int op13;
        uint32_t a12, a23;
        if (!pair_transform(op_of(c12[i12]), op_of(c23[i23]), op13, a12, a23)) return bad;
        e12 += a12;
        e23 += a23;
        if (op13 >= 0 && !nb.add(el((uint32_t)op13, 1))) return bad;
        if (e12 == len_of(c12[i12])) { i12++; e12 = 0; }
        if (e23 == len_of(c23[i23])) { i23++; e23 = 0; }
    }
    return nb.make_and_record_deletions_removed_result();
}

// alignment_utils.rs:173-205
std::vector<uint32_t> append_clipped(const std::vector<uint32_t> &cigar, const std::vector<uint32_t> &orig) {
    if (orig.empty()) return cigar;
    size_t first = 0, last = orig.size() - 1;
    std::vector<uint32_t> out, tail;
    while (is_clipping(orig[first]) && first != last) out.push_back(orig[first++]);
    out.insert(out.end(), cigar.begin(), cigar.end());
    while (is_clipping(orig[last]) && first != last) tail.push_back(orig[last--]);
    out.insert(out.end(), tail.rbegin(), tail.rend());
    return out;
}

}  // namespace

extern "C" lk_status lk_cigar_apply_to_cigar(const uint32_t *c12, int32_t n12, const uint32_t *c23, int32_t n23, int32_t max_ops,
                                             uint32_t *out_ops, int32_t *out_n_ops) {
    if (!out_n_ops || !out_ops || n12 < 0 || n23 < 0 || (n12 > 0 && !c12) || (n23 > 0 && !c23)) return LK_ERR_INVALID;
    BuilderResult r = apply_cigar_to_cigar(std::vector<uint32_t>(c12, c12 + n12), std::vector<uint32_t>(c23, c23 + n23));
    *out_n_ops = r.ok ? write_ops(r.cigar, max_ops, out_ops) : -4;
    return LK_OK;
}

extern "C" lk_status lk_cigar_append_clipped_elements(const uint32_t *cigar, int32_t n, const uint32_t *original, int32_t n_orig,
                                                      int32_t max_ops, uint32_t *out_ops, int32_t *out_n_ops) {
    if (!out_n_ops || !out_ops || n < 0 || n_orig < 1 || !original || (n > 0 && !cigar)) return LK_ERR_INVALID;
    *out_n_ops = write_ops(append_clipped(std::vector<uint32_t>(cigar, cigar + n), std::vector<uint32_t>(original, original + n_orig)),
                           max_ops, out_ops);
    return LK_OK;
}

extern "C" int64_t lk_read_start_on_reference_haplotype(const uint32_t *hap_vs_ref, int32_t n, uint32_t read_start_on_haplotype) {
    bool ok = true;
    const uint32_t r = read_start_on_reference_haplotype(std::vector<uint32_t>(hap_vs_ref, hap_vs_ref + (n > 0 ? n : 0)),
                                                         read_start_on_haplotype, ok);
    return ok ? (int64_t)r : -1;
}

// AlleleLikelihoods::search_best_allele(sample, r, can_be_reference = true, Some(priorities)) for every column of one
// [A x R] row-major block (allele_likelihoods.rs:457-554) and BestAllele::new / is_informative (:1141-1166).
extern "C" lk_status lk_best_alleles_breaking_ties(const double *values, int64_t n_alleles, int64_t n_reads, int64_t n_evidence,
                                                   const int32_t *priorities, double informative_threshold, int32_t *best_allele,
                                                   double *confidence, uint8_t *informative) {
    if (n_alleles < 0 || n_reads < 0 || n_evidence < 0 || n_evidence > n_reads || (n_alleles > 0 && n_reads > 0 && !values))
        return LK_ERR_INVALID;
    if (!best_allele) return LK_ERR_INVALID;
    for (int64_t r = 0; r < n_evidence; r++)
        lk::best_allele_of_read(values, n_alleles, n_reads, r, priorities, informative_threshold, best_allele + r,
                                confidence ? confidence + r : nullptr, informative ? informative + r : nullptr);
    return LK_OK;
}

// create_read_aligned_to_ref after the alignment (alignment_utils.rs:62-163).  Inputs: the read WITHOUT its soft clips
// (ReadClipper::hard_clip_soft_clipped_bases is the caller's), its original CIGAR, the read->haplotype SW result, the
// haplotype's CIGAR against the reference and alignment_start_hap_wrt_ref, the reference haplotype's bases and
// reference_start.  out_n_ops: >= 0 elements of the new CIGAR (clips re-attached); -1 the read is not realigned
// (alignment_offset == -1, :63-66); -2 capacity; -4 where the reference would panic.
extern "C" lk_status lk_realign_finish(const uint8_t *read, int64_t read_len, const uint32_t *orig_cigar, int32_t orig_n,
                                       int64_t orig_read_len, const uint32_t *sw_ops, int32_t sw_n_ops, int32_t sw_offset,
                                       const uint32_t *hap_cigar, int32_t hap_cigar_n, int64_t hap_start_wrt_ref,
                                       const uint8_t *ref_hap, int64_t ref_hap_len, int64_t reference_start, int32_t max_ops,
                                       uint32_t *out_ops, int32_t *out_n_ops, int64_t *out_pos) {
    if (!out_n_ops || !out_ops || !out_pos || orig_n < 1 || !orig_cigar || hap_cigar_n < 0 || (hap_cigar_n > 0 && !hap_cigar))
        return LK_ERR_INVALID;
    if (sw_offset == -1) { *out_n_ops = -1; return LK_OK; }
    if (sw_n_ops < 0) { *out_n_ops = -4; return LK_OK; }
    BuilderResult sw_cigar = consolidate(std::vector<uint32_t>(sw_ops, sw_ops + sw_n_ops));
    if (!sw_cigar.ok) { *out_n_ops = -4; return LK_OK; }
    std::vector<uint32_t> padded_in(hap_cigar, hap_cigar + hap_cigar_n);
    padded_in.push_back(el(M, kHapPad));
    BuilderResult padded = consolidate(padded_in);
    if (!padded.ok) { *out_n_ops = -4; return LK_OK; }
    bool ok = true;
    const uint32_t start_on_ref_hap = read_start_on_reference_haplotype(padded.cigar, (uint32_t)sw_offset, ok);
    if (!ok) { *out_n_ops = -4; return LK_OK; }
    const int64_t read_start_on_reference = reference_start + hap_start_wrt_ref + (int64_t)start_on_ref_hap;
    const uint32_t padded_read_len = read_length_of(padded.cigar);
    if (padded_read_len == 0) { *out_n_ops = -4; return LK_OK; }
    BuilderResult hap_to_ref = trim_cigar_by_bases(padded.cigar, (uint32_t)sw_offset, padded_read_len - 1);
    if (!hap_to_ref.ok) { *out_n_ops = -4; return LK_OK; }
    BuilderResult read_to_ref = apply_cigar_to_cigar(sw_cigar.cigar, hap_to_ref.cigar);
    if (!read_to_ref.ok) { *out_n_ops = -4; return LK_OK; }
    BuilderResult la = left_align_indels(read_to_ref.cigar, ref_hap, (int)ref_hap_len, read, (int)read_len, start_on_ref_hap);
    if (!la.ok) { *out_n_ops = -4; return LK_OK; }
    *out_pos = read_start_on_reference + (int64_t)la.leading_deletion_bases_removed;
    const std::vector<uint32_t> with_clips = append_clipped(la.cigar, std::vector<uint32_t>(orig_cigar, orig_cigar + orig_n));
    // :148-160 "Cigar ... with read length ... != read length"
    const int64_t soft_clipped = orig_read_len - read_len;
    if ((int64_t)read_length_of(la.cigar) + soft_clipped != orig_read_len) { *out_n_ops = -4; return LK_OK; }
    *out_n_ops = write_ops(with_clips, max_ops, out_ops);
    return LK_OK;
}

// realign_reads_to_their_best_haplotype for a batch of regions (assembly_based_caller_utils.rs:208-246).
extern "C" lk_status lk_realign_reads_batch(lk_ctx *ctx, const lk_realign_batch *b, int32_t max_ops, int32_t *best_allele,
                                            uint8_t *informative, uint32_t *cigar_ops, int32_t *n_ops, int64_t *pos) {
    LK_ABI_GUARD_BEGIN
    if (!ctx) return LK_ERR_INVALID;
    if (!b || max_ops < 1 || b->n_regions < 0) return lk_fail(ctx, LK_ERR_INVALID, "bad realign arguments");
    if (b->n_regions == 0) return LK_OK;
    if (!b->region_read_start || !b->region_hap_start || !b->values || !b->hap_off || !b->hap_bases || !b->hap_cigar_off ||
        !b->hap_cigars || !b->hap_start_wrt_ref || !b->read_off || !b->read_bases || !b->read_cigar_off || !b->read_cigars ||
        !b->read_orig_len || !b->reference_start || !best_allele || !cigar_ops || !n_ops || !pos)
        return lk_fail(ctx, LK_ERR_INVALID, "null realign arrays");
    const int64_t n_reads = b->region_read_start[b->n_regions];
    const double thr = b->informative_threshold > 0.0 ? b->informative_threshold : 0.2;  // LOG_10_INFORMATIVE_THRESHOLD
    // ---- best haplotype per read: one search per (region, sample) block; on the host, or as a kernel when the matrices are
    // in device memory (the caller kept lk_engine_compute_read_likelihoods' output in HBM) ----
    if ((b->region_sample_start != nullptr) != (b->sample_read_start != nullptr)) return lk_fail(ctx, LK_ERR_INVALID, "region_sample_start and sample_read_start go together");
    std::vector<uint8_t> info((size_t)n_reads, 0);
    std::vector<int64_t> pair_of((size_t)n_reads, -1);
    std::vector<int64_t> ro(1, 0), ao(1, 0);
    std::vector<int64_t> todo;
    std::vector<lk::BestAlleleBlock> blocks;
    std::vector<int32_t> prio_all;
    {
        int64_t vbase = 0;
        for (int64_t g = 0; g < b->n_regions; g++) {
            const int64_t r0 = b->region_read_start[g], r1 = b->region_read_start[g + 1];
            const int64_t h0 = b->region_hap_start[g], h1 = b->region_hap_start[g + 1];
            const int64_t A = h1 - h0, R = r1 - r0;
            if (A < 0 || R < 0) return lk_fail(ctx, LK_ERR_INVALID, "decreasing region offsets");
            // haplotype_alignment_tiebreaking_priority: (is_ref ? 1 : 0) + 1 - cigar elements
            const int64_t prio_off = (int64_t)prio_all.size();
            for (int64_t a = 0; a < A; a++) {
                const int32_t n_el = (int32_t)(b->hap_cigar_off[h0 + a + 1] - b->hap_cigar_off[h0 + a]);
                const bool is_ref = b->ref_allele_index ? (b->ref_allele_index[g] == a) : (a == 0);
                prio_all.push_back((is_ref ? 1 : 0) + 1 - n_el);
            }
            const int64_t s0 = b->region_sample_start ? b->region_sample_start[g] : g, s1 = b->region_sample_start ? b->region_sample_start[g + 1] : g + 1;
            int64_t covered = r0;
            for (int64_t sm = s0; sm < s1; sm++) {
                const int64_t q0 = b->region_sample_start ? b->sample_read_start[sm] : r0, q1 = b->region_sample_start ? b->sample_read_start[sm + 1] : r1;
                if (q0 != covered || q1 < q0 || q1 > r1) return lk_fail(ctx, LK_ERR_INVALID, "samples must tile their region's reads");
                covered = q1;
                const int64_t Rs = q1 - q0;
                const int64_t n_ev = b->region_sample_start ? (b->sample_n_evidence ? b->sample_n_evidence[sm] : Rs)
                                                            : (b->region_n_evidence ? b->region_n_evidence[g] : Rs);
                if (n_ev < 0 || n_ev > Rs) return lk_fail(ctx, LK_ERR_INVALID, "bad evidence count");
                lk::BestAlleleBlock B;
                B.values_off = vbase + A * (q0 - r0); B.first_read = q0; B.prio_off = prio_off;
                B.n_alleles = (int32_t)A; B.n_reads = (int32_t)Rs; B.n_evidence = (int32_t)n_ev; B.pad = 0;
                blocks.push_back(B);
            }
            if (covered != r1) return lk_fail(ctx, LK_ERR_INVALID, "samples must tile their region's reads");
            vbase += A * R;
        }
    }
    cudaPointerAttributes pattr;
    const bool values_on_device = cudaPointerGetAttributes(&pattr, b->values) == cudaSuccess && pattr.type == cudaMemoryTypeDevice;
    (void)cudaGetLastError();
    if (values_on_device) {
        const lk_status bs = lk_best_alleles_device(ctx, b->values, blocks.data(), (int64_t)blocks.size(), prio_all.data(), (int64_t)prio_all.size(),
                                                    thr, n_reads, best_allele, info.data());
        if (bs != LK_OK) return bs;
    } else {
        for (const lk::BestAlleleBlock &B : blocks) {
            lk_status st = lk_best_alleles_breaking_ties(b->values + B.values_off, B.n_alleles, B.n_reads, B.n_evidence,
                                                         prio_all.data() + B.prio_off, thr, best_allele + B.first_read, nullptr,
                                                         info.data() + B.first_read);
            if (st != LK_OK) return lk_fail(ctx, st, "best_alleles failed");
            for (int64_t r = B.n_evidence; r < B.n_reads; r++) best_allele[B.first_read + r] = -1;
        }
    }
    for (int64_t g = 0; g < b->n_regions; g++) {
        const int64_t h0 = b->region_hap_start[g];
        for (int64_t r = b->region_read_start[g]; r < b->region_read_start[g + 1]; r++) {
            if (best_allele[r] < 0) { best_allele[r] = -1; n_ops[r] = -1; pos[r] = 0; continue; }
            const int64_t h = h0 + best_allele[r];
            pair_of[(size_t)r] = (int64_t)todo.size();
            todo.push_back(r);
            ro.push_back(ro.back() + (b->hap_off[h + 1] - b->hap_off[h]));
            ao.push_back(ao.back() + (b->read_off[r + 1] - b->read_off[r]));
        }
    }
    if (informative) memcpy(informative, info.data(), (size_t)n_reads);
    if (todo.empty()) return LK_OK;
    // ---- one SW batch: reference = best haplotype, alternate = read without soft clips ----
    const int64_t T = (int64_t)todo.size();
    std::vector<int64_t> hap_of((size_t)T), href_of((size_t)T), region_of((size_t)T);
    {
        size_t i = 0;
        for (int64_t g = 0; g < b->n_regions; g++) {
            const int64_t h0 = b->region_hap_start[g];
            for (int64_t r = b->region_read_start[g]; r < b->region_read_start[g + 1]; r++) {
                if (pair_of[(size_t)r] < 0) continue;
                hap_of[i] = h0 + best_allele[r];
                href_of[i] = h0 + (b->ref_allele_index ? b->ref_allele_index[g] : 0);
                region_of[i] = g;
                i++;
            }
        }
    }
    constexpr int64_t kGrain = 2048;
    const lk_sw_params prm = { 10, -15, -30, -5 };  // ALIGNMENT_TO_BEST_HAPLOTYPE_SW_PARAMETERS, smith_waterman_aligner.rs:22-25
    const int32_t sw_max_ops = std::max<int32_t>(max_ops, 64);
    SwStage stage;
    if (!stage.reserve(ctx, (size_t)std::max<int64_t>(ro.back(), 1), (size_t)std::max<int64_t>(ao.back(), 1), (size_t)T, (size_t)sw_max_ops))
        return lk_fail(ctx, LK_ERR_NOMEM, "out of page-locked host memory");
    uint8_t *const refs = stage.refs, *const alts = stage.alts;
    uint32_t *const sw_ops = stage.ops;
    int32_t *const sw_n = stage.n, *const sw_off = stage.off;
    lk::parallel_for(T, kGrain, [&](int64_t i_begin, int64_t i_end) {
        for (int64_t i = i_begin; i < i_end; i++) {
            const int64_t r = todo[(size_t)i], h = hap_of[(size_t)i];
            memcpy(alts + ao[(size_t)i], b->read_bases + b->read_off[r], (size_t)(b->read_off[r + 1] - b->read_off[r]));
            memcpy(refs + ro[(size_t)i], b->hap_bases + b->hap_off[h], (size_t)(b->hap_off[h + 1] - b->hap_off[h]));
        }
    });
    lk_status st = lk_sw_batch_host(ctx, T, ro.data(), refs, ao.data(), alts, &prm, LK_SW_SOFTCLIP, sw_max_ops, sw_ops, sw_n, sw_off);
    if (st != LK_OK) return st;
    // ---- CIGAR composition per read (host, spread over the cores for large batches) ----
    std::atomic<int> failed{ (int)LK_OK };
    lk::parallel_for(T, kGrain, [&](int64_t i_begin, int64_t i_end) {
        for (int64_t i = i_begin; i < i_end; i++) {
            const int64_t r = todo[(size_t)i], h = hap_of[(size_t)i], href = href_of[(size_t)i], g = region_of[(size_t)i];
            if (sw_n[(size_t)i] == -2) { n_ops[r] = -2; continue; }
            const lk_status fs = lk_realign_finish(
                b->read_bases + b->read_off[r], b->read_off[r + 1] - b->read_off[r], b->read_cigars + b->read_cigar_off[r],
                (int32_t)(b->read_cigar_off[r + 1] - b->read_cigar_off[r]), b->read_orig_len[r],
                sw_ops + (size_t)i * (size_t)sw_max_ops, sw_n[(size_t)i], sw_off[(size_t)i], b->hap_cigars + b->hap_cigar_off[h],
                (int32_t)(b->hap_cigar_off[h + 1] - b->hap_cigar_off[h]), b->hap_start_wrt_ref[h], b->hap_bases + b->hap_off[href],
                b->hap_off[href + 1] - b->hap_off[href], b->reference_start[g], max_ops, cigar_ops + r * (int64_t)max_ops, &n_ops[r], &pos[r]);
            if (fs != LK_OK) failed.store((int)fs);
        }
    });
    if (failed.load() != (int)LK_OK) return lk_fail(ctx, (lk_status)failed.load(), "realign_finish failed");
    return LK_OK;
    LK_ABI_GUARD_END(return lk_fail(ctx, LK_ERR_NOMEM, "out of host memory"), return lk_fail(ctx, LK_ERR_INVALID, "internal error"))
}

// The alignment step of the read-threading graph's dangling-end merges (src/read_threading/read_threading_graph.rs:1370-1486:
// generate_cigar_against_downwards_reference_path :1408-1414 and ..._upwards_... :1471-1477): every dangling tail / head of a
// region's graph is aligned to its reference path with OverhangStrategy::LeadingInDel (the callers pass the
// dangling-end parameters, STANDARD_NGS by default) and AlignmentUtils::remove_trailing_deletions (alignment_utils.rs:702-707)
// is applied to the result; only the CIGAR is used.  One Smith-Waterman launch for all pairs.
extern "C" lk_status lk_dangling_end_sw_batch(lk_ctx *ctx, int64_t n_pairs, const int64_t *ref_off, const uint8_t *refs,
                                              const int64_t *alt_off, const uint8_t *alts, const lk_sw_params *params,
                                              int32_t max_ops, uint32_t *cigar_ops, int32_t *n_ops) {
    if (!ctx) return LK_ERR_INVALID;
    if (n_pairs < 0 || max_ops < 1 || !params) return lk_fail(ctx, LK_ERR_INVALID, "bad dangling-end arguments");
    if (n_pairs == 0) return LK_OK;
    if (!ref_off || !refs || !alt_off || !alts || !cigar_ops || !n_ops) return lk_fail(ctx, LK_ERR_INVALID, "null dangling-end arrays");
    try {
        // one element of slack: a CIGAR that fits max_ops only after its trailing deletion is dropped still comes back
        const int32_t sw_max = max_ops + 1;
        std::vector<uint32_t> ops((size_t)n_pairs * (size_t)sw_max);
        std::vector<int32_t> offs((size_t)n_pairs);
        const lk_status st = lk_sw_batch_host(ctx, n_pairs, ref_off, refs, alt_off, alts, params, LK_SW_LEADING_INDEL, sw_max, ops.data(),
                                              n_ops, offs.data());
        if (st != LK_OK) return st;
        for (int64_t p = 0; p < n_pairs; p++) {
            int32_t n = n_ops[p];
            if (n < 0) continue;  // -1 empty sequence (the reference asserts), -2 capacity
            const uint32_t *src = ops.data() + (size_t)p * (size_t)sw_max;
            if (n > 0 && (src[n - 1] & 0xFu) == D) n--;  // remove_trailing_deletions
            if (n > max_ops) { n_ops[p] = -2; continue; }
            memcpy(cigar_ops + (size_t)p * (size_t)max_ops, src, (size_t)n * 4);
            n_ops[p] = n;
        }
        return LK_OK;
    } catch (...) {
        return lk_fail(ctx, LK_ERR_NOMEM, "out of host memory");
    }
}
