/*
 * hipstr_synth.cpp -- synthetic locus generator (see hipstr_synth.h).
 *
 * Per locus: a 2200-bp random chromosome with a period-p motif x ref_copies at
 * [1000, 1000+p*ref_copies), non-motif bases forced either side; candidate
 * alleles = ref +/- 1, 2, ... copies; a repeat block padded by 5 bp
 * (HaplotypeGenerator pads the region, HaplotypeGenerator.h:54-62) between two
 * 35-bp flank blocks; per sample a diploid genotype; per read one of the two
 * alleles (+/- 1 copy of stutter with probability stutter_rate), the indel
 * placed at the STR start as a left-aligner would, Phred 20-39 qualities,
 * flank substitutions with quality '+'.  Reads are pooled and seeded with the
 * product's own host ops (hipstr_pool_reads, hipstr_calc_seeds).
 */
#include "hipstr_synth.h"

#include <climits>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <random>
#include <string>
#include <vector>

struct hipstr_synth {
  hipstr_synth_view_t view;
  std::vector<int32_t> locus_block_off, locus_pool_off, block_period, block_opt_off, opt_seq_off, pool_seq_off, pool_seed;
  std::vector<int64_t> locus_hap_off, locus_out_off;
  std::vector<double> block_stutter;
  std::vector<char> opt_seq, pool_bases, pool_quals;
  std::vector<int32_t> locus_read_off, locus_sample_off, pool_index, sample_label, read_weight, n_haps, true_gt, read_bp_diff;
  std::vector<uint8_t> second_mate, haploid, read_rev_strand;
  std::vector<double> log_p1, log_p2;
  std::vector<int32_t> read_stop;
  std::vector<int32_t> read_seq_off, read_start, read_cigar_off, read_cigar_len, read_name_id, block_start, block_end;
  std::vector<char> read_bases, read_quals, read_cigar_type, chrom_seqs;
};

namespace {

const int kStrStart = 1000, kPad = 5, kFlank = 35, kTrim = 40, kChromLen = 2200;

struct SimRead {
  std::string seq, qual;
  int32_t start;
  std::vector<char> ctype;
  std::vector<int32_t> clen;
  int32_t bp_diff;
};

void push_cigar(SimRead& r, char t, int n) {
  if (n <= 0) return;
  if (!r.ctype.empty() && r.ctype.back() == t) r.clen.back() += n;
  else { r.ctype.push_back(t); r.clen.push_back(n); }
}

}  // namespace

extern "C" hipstr_synth_t* hipstr_synth_create(const hipstr_synth_cfg_t* cfg_in) {
  hipstr_synth_cfg_t cfg = *cfg_in;
  if (cfg.period <= 0) cfg.period = 4;
  if (cfg.ref_copies <= 0) cfg.ref_copies = 12;
  if (cfg.stutter_rate < 0) cfg.stutter_rate = 0.05;
  if (cfg.sub_rate < 0) cfg.sub_rate = 1.0 / 200;
  const int p = cfg.period, ref_bp = p * cfg.ref_copies, str_end = kStrStart + ref_bp;
  hipstr_synth* S = new hipstr_synth();
  const char* ACGT = "ACGT";
  S->locus_block_off.push_back(0); S->locus_pool_off.push_back(0); S->locus_hap_off.push_back(0);
  S->locus_out_off.push_back(0); S->block_opt_off.push_back(0); S->opt_seq_off.push_back(0);
  S->pool_seq_off.push_back(0); S->locus_read_off.push_back(0); S->locus_sample_off.push_back(0);
  S->read_seq_off.push_back(0); S->read_cigar_off.push_back(0);
  int32_t next_name = 0;
  int64_t read_ll_size = 0, post_size = 0;

  for (int l = 0; l < cfg.n_loci; l++) {
    std::mt19937 rng((uint32_t)(cfg.seed * 1000003ull + (uint64_t)l));
    auto uni = [&](int lo, int hi) { return lo + (int)(rng() % (uint32_t)(hi - lo + 1)); };
    auto unif = [&]() { return (rng() >> 8) * (1.0 / 16777216.0); };
    std::string chrom(kChromLen, 'A');
    for (auto& c : chrom) c = ACGT[rng() & 3];
    std::string motif(p, 'A');
    do { for (auto& c : motif) c = ACGT[rng() & 3]; }
    while (p > 1 && motif == std::string(p, motif[0]));
    for (int k = 0; k < ref_bp; k++) chrom[kStrStart + k] = motif[k % p];
    while (chrom[kStrStart - 1] == motif[p - 1]) chrom[kStrStart - 1] = ACGT[rng() & 3];
    while (chrom[str_end] == motif[0]) chrom[str_end] = ACGT[rng() & 3];

    // candidate alleles: ref, ref-1, ref+1, ref-2, ... copies (>= 2 copies)
    std::vector<int> copies;
    copies.push_back(cfg.ref_copies);
    for (int d = 1; (int)copies.size() < cfg.n_alleles && d < 200; d++) {
      if (cfg.ref_copies - d >= 2) copies.push_back(cfg.ref_copies - d);
      if ((int)copies.size() < cfg.n_alleles) copies.push_back(cfg.ref_copies + d);
    }
    const int A = (int)copies.size();
    const int blk_start = kStrStart - kPad, blk_end = str_end + kPad;
    const int first_start = blk_start - kFlank, last_end = blk_end + kFlank;
    auto rep_seq = [&](int k) {
      std::string s = chrom.substr(blk_start, kPad);
      for (int i = 0; i < k * p; i++) s += motif[i % p];
      return s + chrom.substr(str_end, kPad);
    };
    // blocks: flank, repeat, flank
    auto add_opt = [&](const std::string& s) {
      S->opt_seq.insert(S->opt_seq.end(), s.begin(), s.end());
      S->opt_seq_off.push_back((int32_t)S->opt_seq.size());
    };
    const double def_model[6] = {0.95, 0.05, 0.05, 0.95, 0.01, 0.01};  // hipstr_main.cpp:343
    for (int b = 0; b < 3; b++) {
      S->block_period.push_back(b == 1 ? p : 0);
      for (int k = 0; k < 6; k++) S->block_stutter.push_back(def_model[k]);
      if (b == 0) add_opt(chrom.substr(first_start, kFlank));
      else if (b == 2) add_opt(chrom.substr(blk_end, kFlank));
      else for (int a = 0; a < A; a++) add_opt(rep_seq(copies[a]));
      S->block_opt_off.push_back((int32_t)S->opt_seq_off.size() - 1);
      S->block_start.push_back(b == 0 ? first_start : b == 1 ? blk_start : blk_end);
      S->block_end.push_back(b == 0 ? blk_start : b == 1 ? blk_end : last_end);
    }
    S->chrom_seqs.insert(S->chrom_seqs.end(), chrom.begin(), chrom.end());
    S->locus_block_off.push_back((int32_t)S->block_period.size());

    // reads
    std::vector<SimRead> reads;
    std::vector<int32_t> labels;
    std::vector<uint8_t> mates;
    const int snp_left = kStrStart - 15, snp_right = str_end + 12;   // reference coordinates of the planted SNPs
    char alt_left = chrom[snp_left], alt_right = chrom[snp_right];
    if (cfg.flank_snp_freq > 0) {
      while (alt_left == chrom[snp_left]) alt_left = ACGT[rng() & 3];
      while (alt_right == chrom[snp_right]) alt_right = ACGT[rng() & 3];
    }
    for (int s = 0; s < cfg.n_samples; s++) {
      int gt[2] = {uni(0, A - 1), uni(0, A - 1)};
      if (cfg.haploid) gt[1] = gt[0];
      S->true_gt.push_back(gt[0]); S->true_gt.push_back(gt[1]);
      bool carries[2][2] = {{false, false}, {false, false}};   // [chromosome copy][left / right SNP]
      if (cfg.flank_snp_freq > 0)
        for (int c = 0; c < 2; c++)
          for (int side = 0; side < 2; side++) carries[c][side] = (cfg.haploid && c == 1) ? carries[0][side] : unif() < cfg.flank_snp_freq;
      for (int r = 0; r < cfg.reads_per_sample; r++) {
        const int copy = rng() & 1;
        int k = copies[gt[copy]];
        if (unif() < cfg.stutter_rate) k += (rng() & 1) ? 1 : -1;
        if (k < 1) k = 1;
        const int delta = (k - cfg.ref_copies) * p;           // bp difference vs reference
        const int hap_str_end = str_end + delta;               // in sample-haplotype coordinates
        std::string hap = chrom.substr(0, kStrStart);
        for (int i = 0; i < k * p; i++) hap += motif[i % p];
        hap += chrom.substr(str_end);
        if (carries[copy][0]) hap[snp_left] = alt_left;
        if (carries[copy][1]) hap[snp_right + delta] = alt_right;
        const bool mate = unif() < cfg.mate_rate;
        for (int m = 0; m < (mate ? 2 : 1); m++) {
          SimRead rd;
          int start = uni(kStrStart - 40, kStrStart - 10);
          int end = start + cfg.read_len;                      // haplotype coordinates, exclusive
          if (cfg.trim) end = std::min(end, hap_str_end + kTrim);
          end = std::min(end, (int)hap.size());
          rd.start = start;
          rd.bp_diff = delta;
          rd.seq = hap.substr(start, end - start);
          rd.qual.resize(rd.seq.size());
          for (auto& q : rd.qual) q = (char)('!' + uni(20, 39));
          // CIGAR vs the reference with the indel at the STR start; substitutions only in flanks
          int i = 0;
          const int n = (int)rd.seq.size();
          auto flank_run = [&](int count, int ref_shift) {
            for (int e = i + count; i < e; i++) {
              if (unif() < cfg.sub_rate) {
                char c;
                do c = ACGT[rng() & 3]; while (c == rd.seq[i]);
                rd.seq[i] = c; rd.qual[i] = '+';
                push_cigar(rd, 'X', 1);
              } else
                push_cigar(rd, rd.seq[i] == chrom[start + i - ref_shift] ? '=' : 'X', 1);   // planted SNPs read as mismatches
            }
          };
          flank_run(std::min(n, kStrStart - start), 0);
          if (i < n) {
            if (delta > 0) { int ins = std::min(delta, n - i); push_cigar(rd, 'I', ins); i += ins; }
            else if (delta < 0) push_cigar(rd, 'D', -delta);
            int in_str = std::min(n - i, std::max(0, hap_str_end - (start + i)));
            push_cigar(rd, '=', in_str); i += in_str;
            flank_run(n - i, delta);
          }
          reads.push_back(rd);
          labels.push_back(s);
          mates.push_back(m == 1);
        }
      }
    }
    // pool + seed with the product's host ops
    const int R = (int)reads.size();
    std::vector<int32_t> seq_off(R + 1, 0), cig_off(R + 1, 0), starts(R), lens(R);
    std::vector<char> bases, quals, ctype;
    std::vector<int32_t> clen;
    for (int r = 0; r < R; r++) {
      bases.insert(bases.end(), reads[r].seq.begin(), reads[r].seq.end());
      quals.insert(quals.end(), reads[r].qual.begin(), reads[r].qual.end());
      seq_off[r + 1] = (int32_t)bases.size();
    }
    std::vector<int32_t> pidx(R), pfirst(R), pseq_off(R + 1);
    std::vector<char> pbases(bases.size()), pquals(bases.size());
    int32_t P = 0;
    hipstr_pool_reads(R, seq_off.data(), bases.data(), quals.data(), pidx.data(), &P, pfirst.data(), pseq_off.data(),
                      pbases.data(), pquals.data());
    for (int q = 0; q < P; q++) {
      const SimRead& f = reads[pfirst[q]];
      starts[q] = f.start; lens[q] = (int32_t)f.seq.size();
      ctype.insert(ctype.end(), f.ctype.begin(), f.ctype.end());
      clen.insert(clen.end(), f.clen.begin(), f.clen.end());
      cig_off[q + 1] = (int32_t)ctype.size();
    }
    std::vector<int32_t> seeds(P);
    const int32_t rs = blk_start, re = blk_end;
    hipstr_status_t st = hipstr_calc_seeds(P, starts.data(), lens.data(), cig_off.data(), ctype.data(), clen.data(),
                                           first_start, last_end, 1, &rs, &re, seeds.data());
    if (st != HIPSTR_OK) std::fill(seeds.begin(), seeds.end(), -1);
    const int32_t base = (int32_t)S->pool_bases.size();
    S->pool_bases.insert(S->pool_bases.end(), pbases.begin(), pbases.begin() + pseq_off[P]);
    S->pool_quals.insert(S->pool_quals.end(), pquals.begin(), pquals.begin() + pseq_off[P]);
    for (int q = 0; q < P; q++) {
      S->pool_seq_off.push_back(base + pseq_off[q + 1]);
      S->pool_seed.push_back(seeds[q]);
    }
    S->locus_pool_off.push_back((int32_t)S->pool_seed.size());
    S->locus_hap_off.push_back(S->locus_hap_off.back() + A);
    S->locus_out_off.push_back(S->locus_out_off.back() + (int64_t)P * A);
    for (int r = 0; r < R; r++) {
      S->pool_index.push_back(pidx[r]);
      S->sample_label.push_back(labels[r]);
      S->second_mate.push_back(mates[r]);
      S->read_weight.push_back(mates[r] ? 0 : 1);
      S->log_p1.push_back(0.0); S->log_p2.push_back(0.0);
      S->read_bp_diff.push_back(reads[r].bp_diff);
      S->read_bases.insert(S->read_bases.end(), reads[r].seq.begin(), reads[r].seq.end());
      S->read_quals.insert(S->read_quals.end(), reads[r].qual.begin(), reads[r].qual.end());
      // offsets of the C-ABI are 32-bit: a request that does not fit is refused (callers split their locus lists)
      if (S->read_bases.size() > (size_t)INT32_MAX) { delete S; return nullptr; }
      S->read_seq_off.push_back((int32_t)S->read_bases.size());
      S->read_start.push_back(reads[r].start);
      {
        int32_t stop = reads[r].start - 1;
        for (size_t c = 0; c < reads[r].ctype.size(); c++)
          if (reads[r].ctype[c] != 'I') stop += reads[r].clen[c];
        S->read_stop.push_back(stop);
      }
      S->read_cigar_type.insert(S->read_cigar_type.end(), reads[r].ctype.begin(), reads[r].ctype.end());
      S->read_cigar_len.insert(S->read_cigar_len.end(), reads[r].clen.begin(), reads[r].clen.end());
      S->read_cigar_off.push_back((int32_t)S->read_cigar_type.size());
      if (!mates[r]) next_name++;
      S->read_name_id.push_back(next_name);
      S->read_rev_strand.push_back((uint8_t)((((uint32_t)S->read_name_id.size() * 2654435761u) >> 13) & 1));
    }
    S->locus_read_off.push_back((int32_t)S->pool_index.size());
    S->locus_sample_off.push_back(S->locus_sample_off.back() + cfg.n_samples);
    S->n_haps.push_back(A);
    S->haploid.push_back(cfg.haploid ? 1 : 0);
    read_ll_size += (int64_t)R * A;
    post_size += (int64_t)cfg.n_samples * A * A;
  }

  hipstr_align_batch_t& b = S->view.batch;
  std::memset(&S->view, 0, sizeof(S->view));
  b.n_loci = cfg.n_loci;
  b.n_blocks = (int32_t)S->block_period.size();
  b.n_options = (int32_t)S->opt_seq_off.size() - 1;
  b.n_pools = (int32_t)S->pool_seed.size();
  b.n_haps = S->locus_hap_off.back();
  b.locus_block_off = S->locus_block_off.data();
  b.locus_pool_off = S->locus_pool_off.data();
  b.locus_hap_off = S->locus_hap_off.data();
  b.locus_out_off = S->locus_out_off.data();
  b.block_period = S->block_period.data();
  b.block_opt_off = S->block_opt_off.data();
  b.block_stutter = S->block_stutter.data();
  b.opt_seq_off = S->opt_seq_off.data();
  b.opt_seq = S->opt_seq.data();
  b.pool_seq_off = S->pool_seq_off.data();
  b.pool_bases = S->pool_bases.data();
  b.pool_quals = S->pool_quals.data();
  b.pool_seed = S->pool_seed.data();
  b.realign_pool = NULL;
  b.realign_hap = NULL;
  S->view.n_reads = (int64_t)S->pool_index.size();
  S->view.locus_read_off = S->locus_read_off.data();
  S->view.locus_sample_off = S->locus_sample_off.data();
  S->view.pool_index = S->pool_index.data();
  S->view.sample_label = S->sample_label.data();
  S->view.second_mate = S->second_mate.data();
  S->view.read_weight = S->read_weight.data();
  S->view.log_p1 = S->log_p1.data();
  S->view.log_p2 = S->log_p2.data();
  S->view.n_haps = S->n_haps.data();
  S->view.haploid = S->haploid.data();
  S->view.true_gt = S->true_gt.data();
  S->view.read_bp_diff = S->read_bp_diff.data();
  S->view.read_ll_size = read_ll_size;
  S->view.post_size = post_size;
  S->view.read_seq_off = S->read_seq_off.data();
  S->view.read_bases = S->read_bases.data();
  S->view.read_quals = S->read_quals.data();
  S->view.read_start = S->read_start.data();
  S->view.read_cigar_off = S->read_cigar_off.data();
  S->view.read_cigar_type = S->read_cigar_type.data();
  S->view.read_cigar_len = S->read_cigar_len.data();
  S->view.read_name_id = S->read_name_id.data();
  S->view.block_start = S->block_start.data();
  S->view.block_end = S->block_end.data();
  S->view.chrom_len = kChromLen;
  S->view.chrom_seqs = S->chrom_seqs.data();
  S->view.region_start = kStrStart;
  S->view.region_stop = str_end;
  S->view.read_stop = S->read_stop.data();
  S->view.read_rev_strand = S->read_rev_strand.data();
  return S;
}

extern "C" const hipstr_synth_view_t* hipstr_synth_view(const hipstr_synth_t* s) { return &s->view; }
extern "C" void hipstr_synth_destroy(hipstr_synth_t* s) { delete s; }
