/*
 * trace_lists.cpp -- the complete flank indel / flank SNP lists of a trace (AlignmentTrace::flank_indel_data /
 * flank_snp_data, SeqAlignment/AlignmentTraceback.h:29-33), rebuilt on the host from K5's operation string.
 *
 * K5 returns these lists in fixed-size slots (HIPSTR_MAX_TRACE_INDELS / _SNPS per trace) together with the TRUE counts.
 * A read with more entries than slots -- a chimeric or mismapped read re-aligned by the left aligner can carry dozens of
 * flank mismatches -- is rare, so instead of sizing every slot for the worst case the caller asks for the full lists of
 * exactly those traces here.  The walk is the accounting half of HapAligner::retrace (HapAligner.cpp:363-571) driven by
 * the operations K5 already chose: from the seed outwards on both sides, block by block; a flank SNP is a matched base
 * that differs from the haplotype with log P(base correct) above MIN_SNP_LOG_PROB_CORRECT (:24); an indel is recorded
 * when its run ENDS inside a flank block (a run cut by a block end or by the end of the read is not recorded -- the
 * reference's loop has no flush there, and neither has this).
 */
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hipstr_b200.h"
#include "../csrc/flatten.h"

namespace {

const double kMinSnpLogProbCorrect = -0.0043648054;   /* HapAligner.cpp:24 */

struct Block { int len; bool rep; int32_t start_fw, start_rv; int fw_index; };

struct Lists {
  int32_t cap_indels, cap_snps, n_indels = 0, n_snps = 0;
  int32_t *indels, *snps;
  void indel(int pos, int size) {
    if (n_indels < cap_indels) { indels[2 * n_indels] = pos; indels[2 * n_indels + 1] = size; }
    n_indels++;
  }
  void snp(int pos, char base) {
    if (n_snps < cap_snps) { snps[2 * n_snps] = pos; snps[2 * n_snps + 1] = base; }
    n_snps++;
  }
};

/* one side of the seed; `ops` in WALK order (from the seed outwards), `blocks` / `hap` oriented like the walk */
void replay_side(bool rev, const std::string& ops, int n_side, int n_read, const char* bases, const char* quals, const std::string& hap,
                 const std::vector<Block>& blocks, const int32_t* stutter_size, int block_index, int base_index, Lists& acc) {
  const hipstr::HostTables& T = hipstr::host_tables();
  auto ridx = [&](int k) { return rev ? n_read - 1 - k : k; };
  int seq_index = n_side - 1;
  size_t c = 0;
  std::vector<int> row_start(blocks.size());
  for (size_t b = 0, row = 0; b < blocks.size(); row += blocks[b].len, b++) row_start[b] = (int)row;
  while (block_index >= 0) {
    const Block& blk = blocks[(size_t)block_index];
    if (blk.rep) {
      const int size = stutter_size[blk.fw_index], len = blk.len;
      c += (size_t)std::min(len + size, seq_index + 1) + (size < 0 ? (size_t)-size : 0);
      if (len + size >= seq_index + 1) return;
      seq_index -= len + size;
    } else {
      int prev_type = -1;
      int pos = (rev ? blk.start_rv : blk.start_fw) + (rev ? -base_index : base_index);
      const int step = rev ? 1 : -1;
      int indel_seq_index = -1, indel_pos = -1;
      while (base_index >= 0 && seq_index >= 0) {
        if (c >= ops.size()) return;
        const char op = ops[c++];
        const int type = op == 'M' ? 0 : (op == 'D' ? 1 : 2);
        if (type != prev_type) {
          if (prev_type == 1) { if (rev) acc.indel(indel_pos, indel_pos - pos); else acc.indel(pos + 1, pos - indel_pos); }
          else if (prev_type == 2) acc.indel(indel_pos + (rev ? 0 : 1), indel_seq_index - seq_index);
          if (type == 1 || type == 2) { indel_seq_index = seq_index; indel_pos = pos; }
          prev_type = type;
        }
        if (type == 0) {
          const int r = ridx(seq_index);
          if (hap[(size_t)(row_start[(size_t)block_index] + base_index)] != bases[r] && T.qual_lut[(unsigned char)quals[r]][0] > kMinSnpLogProbCorrect)
            acc.snp(pos, bases[r]);
          seq_index--; base_index--; pos += step;
        } else if (type == 1) {
          base_index--; pos += step;
        } else {
          seq_index--;
        }
        if (seq_index == -1 || (base_index == -1 && block_index == 0)) return;
      }
    }
    --block_index;
    if (block_index >= 0) base_index = blocks[(size_t)block_index].len - 1;
  }
}

}  // namespace

extern "C" hipstr_status_t hipstr_trace_flank_lists(const hipstr_align_batch_t* b, const int32_t* block_start, int32_t pool, int32_t hap_index,
                                                    const char* hap_aln, int32_t seed_hap_pos, const int32_t* stutter_size,
                                                    const char* own_quals, int32_t cap_indels, int32_t* n_indels, int32_t* indels,
                                                    int32_t cap_snps, int32_t* n_snps, int32_t* snps) {
  if (!b || !block_start || !hap_aln || !stutter_size || !n_indels || !n_snps || pool < 0 || pool >= b->n_pools ||
      (cap_indels > 0 && !indels) || (cap_snps > 0 && !snps))
    return HIPSTR_ERR_BAD_ARG;
  int l = (int)(std::upper_bound(b->locus_pool_off, b->locus_pool_off + b->n_loci + 1, pool) - b->locus_pool_off) - 1;
  if (l < 0 || l >= b->n_loci) return HIPSTR_ERR_BAD_ARG;
  const int b0 = b->locus_block_off[l], nb = b->locus_block_off[l + 1] - b0;
  if (nb < 1 || nb > HIPSTR_MAX_BLOCKS_PER_LOCUS) return HIPSTR_ERR_BAD_ARG;
  std::vector<int32_t> n_opts((size_t)nb), choice((size_t)nb);
  for (int k = 0; k < nb; k++) n_opts[(size_t)k] = b->block_opt_off[b0 + k + 1] - b->block_opt_off[b0 + k];
  if (hap_index < 0 || hap_index >= b->locus_hap_off[l + 1] - b->locus_hap_off[l]) return HIPSTR_ERR_BAD_ARG;
  hipstr::haplotype_options(nb, n_opts.data(), hap_index, choice.data());
  std::vector<Block> fw((size_t)nb);
  std::string hap_fw;
  for (int k = 0; k < nb; k++) {
    const int o = b->block_opt_off[b0 + k] + choice[(size_t)k], o0 = b->block_opt_off[b0 + k];
    const int len = b->opt_seq_off[o + 1] - b->opt_seq_off[o];
    hap_fw.append(b->opt_seq + b->opt_seq_off[o], (size_t)len);
    Block& blk = fw[(size_t)k];
    blk.len = len;
    blk.rep = b->block_period[b0 + k] > 0;
    blk.start_fw = block_start[b0 + k];
    blk.start_rv = block_start[b0 + k] + (b->opt_seq_off[o0 + 1] - b->opt_seq_off[o0]) - 1;   // end of the reference allele - 1
    blk.fw_index = k;
  }
  std::vector<Block> rv(fw.rbegin(), fw.rend());
  const std::string hap_rv(hap_fw.rbegin(), hap_fw.rend());
  const int s0 = b->pool_seq_off[pool], n = b->pool_seq_off[pool + 1] - s0, seed = b->pool_seed[pool];
  if (seed <= 0 || seed >= n - 1) return HIPSTR_ERR_INVALID_SEED;
  const char* bases = b->pool_bases + s0;
  const char* quals = own_quals ? own_quals : b->pool_quals + s0;
  const int hs_len = (int)hap_fw.size();
  if (seed_hap_pos < 0 || seed_hap_pos >= hs_len) return HIPSTR_ERR_BAD_ARG;
  // split the operation string at the seed: the left part ends with the operation of read base seed-1
  const std::string ops(hap_aln);
  size_t at = 0;
  for (int consumed = 0; at < ops.size() && consumed < seed; at++) consumed += ops[at] != 'D';
  if (at >= ops.size() || ops[at] != 'M') return HIPSTR_ERR_BAD_ARG;
  std::string left(ops.begin(), ops.begin() + (long)at), right(ops.begin() + (long)at + 1, ops.end());
  std::reverse(left.begin(), left.end());   // walk order: from the seed towards the read start
  Lists acc;
  acc.cap_indels = cap_indels; acc.cap_snps = cap_snps; acc.indels = indels; acc.snps = snps;
  auto locate = [](const std::vector<Block>& blocks, int position, int& blk, int& off) {
    blk = 0; off = position;
    while (off >= blocks[(size_t)blk].len) { off -= blocks[(size_t)blk].len; blk++; }
  };
  if (seed_hap_pos != 0) {   // (a seed on the first haplotype base soft-clips everything to its left)
    int fb, fc;
    locate(fw, seed_hap_pos, fb, fc);
    if (fc == 0) replay_side(false, left, seed, n, bases, quals, hap_fw, fw, stutter_size, fb - 1, fw[(size_t)fb - 1].len - 1, acc);
    else replay_side(false, left, seed, n, bases, quals, hap_fw, fw, stutter_size, fb, fc - 1, acc);
  }
  const int rmax = hs_len - 1 - seed_hap_pos;
  if (rmax != 0) {
    int rb, rc;
    locate(rv, rmax, rb, rc);
    if (rc == 0) replay_side(true, right, n - 1 - seed, n, bases, quals, hap_rv, rv, stutter_size, rb - 1, rv[(size_t)rb - 1].len - 1, acc);
    else replay_side(true, right, n - 1 - seed, n, bases, quals, hap_rv, rv, stutter_size, rb, rc - 1, acc);
  }
  *n_indels = acc.n_indels;
  *n_snps = acc.n_snps;
  return HIPSTR_OK;
}
