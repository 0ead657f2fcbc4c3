/*
 * region_driver.cpp -- hipstr_process_regions: BAM files -> VCF records for a WINDOW of regions.
 *
 * The decisions of the reference's per-region driver
 *   BamProcessor::process_regions                      src/bam_processor.cpp:521-617
 *   SNPBamProcessor::process_reads                     src/snp_bam_processor.cpp:36-118
 *   GenotyperBamProcessor::analyze_reads_and_phasing   src/genotyper_bam_processor.cpp:160-289
 *   GenotyperBamProcessor::learn_stutter_model         src/genotyper_bam_processor.cpp:104-158
 * re-arranged around window-sized device calls: the regions of the window are read and filtered on all host threads
 * (every thread owns its BAM handles), then ONE K7 launch gives the phasing log-likelihoods of every read of the
 * window, ONE K4 call trains all stutter models, the K6 launches left-align all reads, and the lockstep genotyper
 * (K1 K2 K3 K5) advances all loci together before K3b / K5 write the records.  The reference does the same steps one
 * region at a time on one core.
 */
#include <algorithm>
#include <atomic>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hipstr_b200.h"
#include "ingest_handles.h"
#include "seq_stutter_genotyper.h"

namespace {

enum RegionStatus { GENOTYPED = 0, TOO_LONG = 1, NEAR_CONTIG_END = 2, TOO_FEW_READS = 3, TOO_MANY_READS = 4, NO_STUTTER_MODEL = 5,
                    GENOTYPING_FAILED = 6, UNKNOWN_CHROMOSOME = 7 };

struct Locus {
  int region = -1;
  int chrom = -1;
  hipstr::FilteredReads kept;
  // STR reads in the order of SNPBamProcessor::process_reads: per sample the paired ones, then the unpaired ones
  std::vector<const BamRecord*> reads, mates;   // mates[i] == nullptr for unpaired reads
  std::vector<int32_t> label;
  std::vector<double> log_p1, log_p2;
  double stutter[6];
  bool has_model = false;
};

/* alignments of one region packed into flat arrays (built per region on the host threads, then concatenated) */
struct AlnPiece {
  std::vector<int32_t> pos, end, seq_len, cigar_n, cigar_len, group_n;
  std::string bases, quals, cigar_type;
  void add(const BamRecord& a) {
    pos.push_back(a.pos); end.push_back(a.end_pos);
    bases += a.bases; quals += a.quals;
    seq_len.push_back((int32_t)a.bases.size());
    for (const auto& op : a.cigar) { cigar_type += op.first; cigar_len.push_back(op.second); }
    cigar_n.push_back((int32_t)a.cigar.size());
  }
  void append_to(std::vector<int32_t>& all_pos, std::vector<int32_t>& all_end, std::vector<int32_t>& seq_off, std::string& all_bases,
                 std::string& all_quals, std::vector<int32_t>& cigar_off, std::string& all_types, std::vector<int32_t>& all_lens) const {
    all_pos.insert(all_pos.end(), pos.begin(), pos.end());
    all_end.insert(all_end.end(), end.begin(), end.end());
    for (int32_t n : seq_len) seq_off.push_back(seq_off.back() + n);
    for (int32_t n : cigar_n) cigar_off.push_back(cigar_off.back() + n);
    all_bases += bases; all_quals += quals; all_types += cigar_type;
    all_lens.insert(all_lens.end(), cigar_len.begin(), cigar_len.end());
  }
};

thread_local std::string g_driver_error;

}  // namespace

struct hipstr_region_results {
  struct Region { int32_t status = 0, pos = 0, n_reads = 0; std::string text; };
  std::vector<Region> regions;
  std::vector<std::string> samples;
  std::string sample_text;
  double seconds[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double genotyper_seconds[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // hipstr_genotyper_timing of the window
  int64_t genotyper_stats[3] = {0, 0, 0};                       // alignments, traces, rounds
  int64_t counters[4] = {0, 0, 0, 0};   // alignments read, reads kept, reads with phase information, left-alignment failures
};

extern "C" {

const char* hipstr_process_regions_last_error(void) { return g_driver_error.c_str(); }

void hipstr_pipeline_default_options(hipstr_pipeline_options_t* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  hipstr_filter_default_options(&o->filter);
  o->max_str_length = 100;
  o->min_total_reads = 100;
  o->max_total_haplotypes = 1000;
  o->max_flank_haplotypes = 4;
  o->min_flank_freq = 0.01;
  o->max_em_iter = 100;
  o->abs_ll_converge = 0.01;
  o->frac_ll_converge = 0.001;
  o->use_def_stutter_model = 0;
  const double def[6] = {0.95, 0.05, 0.05, 0.95, 0.01, 0.01};
  std::memcpy(o->def_stutter_model, def, sizeof(def));
  o->recalc_stutter_model = 0;
  o->skip_padding = 15;
  o->host_threads = 0;
}

hipstr_status_t hipstr_process_regions(hipstr_ctx_t* ctx, int32_t n_files, const char* const* bam_paths, hipstr_snp_vcf_t* snp_vcf,
                                       int32_t n_chroms, const char* const* chrom_names, const char* const* chrom_seqs,
                                       int32_t n_regions, const char* const* region_chrom, const int32_t* region_start,
                                       const int32_t* region_stop, const int32_t* region_period, const char* const* region_name,
                                       const hipstr_pipeline_options_t* opt, const hipstr_vcf_options_t* vcf_opt,
                                       hipstr_region_results_t** out) {
  using hipstr::now_s;
  if (!ctx) return HIPSTR_ERR_NO_DEVICE;
  if (n_files < 1 || !bam_paths || n_chroms < 1 || !chrom_names || !chrom_seqs || n_regions < 0 || !opt || !vcf_opt || !out) return HIPSTR_ERR_BAD_ARG;
  if (n_regions > 0 && (!region_chrom || !region_start || !region_stop || !region_period)) return HIPSTR_ERR_BAD_ARG;
  std::unique_ptr<hipstr_region_results> R(new hipstr_region_results());
  R->regions.resize(n_regions);
  double t = now_s();

  // read groups -> sample / library (hipstr_main.cpp:442-468) and the sorted sample list of the VCF
  std::vector<std::string> paths(bam_paths, bam_paths + n_files);
  std::map<std::string, std::string> rg_to_sample, rg_to_library;
  std::vector<std::string> ref_names;
  {
    std::set<std::string> samples;
    for (int f = 0; f < n_files; f++) {
      hipstr::BamFile file;
      if (!file.open(paths[f])) { g_driver_error = file.error(); return HIPSTR_ERR_BAD_ARG; }
      if (f == 0) ref_names = file.ref_names();
      else if (file.ref_names() != ref_names) { g_driver_error = "BAM header mismatch issue: the files must have the same reference sequences"; return HIPSTR_ERR_BAD_ARG; }
      if (file.read_groups().empty()) { g_driver_error = "Provided BAM files don't contain read groups in the header"; return HIPSTR_ERR_BAD_ARG; }
      for (const hipstr::BamReadGroup& g : file.read_groups()) {
        if (!g.has_sample || !g.has_library) { g_driver_error = "RG in BAM header is lacking the SM or LB tag"; return HIPSTR_ERR_BAD_ARG; }
        rg_to_sample[paths[f] + g.id] = g.sample;
        rg_to_library[paths[f] + g.id] = g.library;
        samples.insert(g.sample);
      }
    }
    R->samples.assign(samples.begin(), samples.end());
    for (const std::string& s : R->samples) { R->sample_text += s; R->sample_text += '\n'; }
  }
  auto is_haploid = [&](const char* chrom) {
    for (int c = 0; c < opt->n_haploid_chroms; c++)
      if (std::strcmp(opt->haploid_chroms[c], chrom) == 0) return true;
    return false;
  };
  std::map<std::string, int> chrom_index;
  std::vector<std::string> seqs(n_chroms);
  for (int c = 0; c < n_chroms; c++) chrom_index[chrom_names[c]] = c;
  // only the chromosomes this window's regions name are materialised (a caller passes the whole genome on every window)
  for (int i = 0; i < n_regions; i++) {
    auto ci = chrom_index.find(region_chrom[i]);
    if (ci != chrom_index.end() && seqs[ci->second].empty() && chrom_seqs[ci->second]) seqs[ci->second] = chrom_seqs[ci->second];
  }

  // ---- per region, on all host threads: region query, read filters, mate pairing, PCR duplicates -----------------
  std::vector<std::unique_ptr<Locus> > slots(n_regions);
  std::vector<std::string> errors(n_regions);
  std::atomic<int> next(0);
  std::atomic<int64_t> n_records(0);
  const int n_threads = std::max(1, std::min(opt->host_threads > 0 ? opt->host_threads : hipstr::host_threads(), std::max(n_regions, 1)));
  auto worker = [&]() {
    std::vector<std::unique_ptr<hipstr::BamFile> > files;   // this thread's handles, opened on first use
    for (int i; (i = next.fetch_add(1)) < n_regions;) {
      hipstr_region_results::Region& res = R->regions[i];
      const int32_t start = region_start[i], stop = region_stop[i];
      if (stop - start > opt->max_str_length) { res.status = TOO_LONG; continue; }
      auto ci = chrom_index.find(region_chrom[i]);
      if (ci == chrom_index.end()) { res.status = UNKNOWN_CHROMOSOME; continue; }
      const std::string& seq = seqs[ci->second];
      if (start < 50 || stop + 50 >= (int64_t)seq.size()) { res.status = NEAR_CONTIG_END; continue; }
      try {
        if (files.empty())
          for (const std::string& p : paths) {
            files.emplace_back(new hipstr::BamFile());
            if (!files.back()->open(p)) throw hipstr::FilterError(files.back()->error());
          }
        std::vector<BamRecord> records;
        const int32_t dist = opt->filter.max_mate_dist;
        const int32_t str_region[2] = {start, stop};
        for (size_t f = 0; f < files.size(); f++)
          if (!files[f]->fetch(region_chrom[i], start < dist ? 0 : start - dist, stop + dist, (int32_t)f, records, str_region))
            throw hipstr::FilterError(files[f]->error());
        n_records += (int64_t)records.size();
        std::unique_ptr<Locus> L(new Locus());
        L->region = i;
        L->chrom = ci->second;
        hipstr::ReadFilter filter;
        hipstr::FilterOptions& fo = filter.options;
        fo.max_mate_dist = opt->filter.max_mate_dist; fo.min_bp_before_indel = opt->filter.min_bp_before_indel; fo.min_flank = opt->filter.min_flank;
        fo.min_read_end_match = opt->filter.min_read_end_match; fo.maximal_end_match_window = opt->filter.maximal_end_match_window;
        fo.require_paired_reads = opt->filter.require_paired_reads; fo.min_sum_qual_log_prob = opt->filter.min_sum_qual_log_prob;
        fo.max_total_reads = opt->filter.max_total_reads; fo.base_qual_trim = (char)opt->filter.base_qual_trim;
        fo.remove_pcr_dups = opt->filter.remove_pcr_dups != 0; fo.trim_adapters = opt->filter.trim_adapters != 0;
        filter.run(records, ref_names, paths, seq, {std::make_pair(start, stop)}, rg_to_sample, L->kept);
        if (fo.remove_pcr_dups) hipstr::ReadFilter::remove_pcr_duplicates(rg_to_library, paths, L->kept);
        const hipstr::FilteredReads& k = L->kept;
        for (size_t g = 0; g < k.rg_names.size(); g++) {
          for (size_t j = 0; j < k.paired[g].size(); j++) { L->reads.push_back(&k.paired[g][j]); L->mates.push_back(&k.mates[g][j]); L->label.push_back((int32_t)g); }
          for (size_t j = 0; j < k.unpaired[g].size(); j++) { L->reads.push_back(&k.unpaired[g][j]); L->mates.push_back(nullptr); L->label.push_back((int32_t)g); }
        }
        res.n_reads = (int32_t)L->reads.size();
        if ((int32_t)L->reads.size() < opt->min_total_reads) { res.status = TOO_FEW_READS; continue; }
        if (k.counts.too_many_reads) { res.status = TOO_MANY_READS; continue; }
        L->log_p1.assign(L->reads.size(), 0.0);
        L->log_p2.assign(L->reads.size(), 0.0);
        slots[i] = std::move(L);
      } catch (const std::exception& e) {
        errors[i] = e.what();
      }
    }
  };
  {
    std::vector<std::thread> pool;
    for (int w = 1; w < n_threads; w++) pool.emplace_back(worker);
    worker();
    for (std::thread& th : pool) th.join();
  }
  for (int i = 0; i < n_regions; i++)
    if (!errors[i].empty()) { g_driver_error = std::string(region_chrom[i]) + ":" + std::to_string(region_start[i]) + ": " + errors[i]; return HIPSTR_ERR_BAD_ARG; }
  std::vector<Locus*> loci;
  for (auto& s : slots)
    if (s) loci.push_back(s.get());
  R->counters[0] = n_records;
  for (Locus* L : loci) R->counters[1] += (int64_t)L->reads.size();
  R->seconds[0] = now_s() - t; t = now_s();

  // ---- phasing log-likelihoods: from the 10X haplotype tags, or every read (+ mate) of the window in ONE K7 launch --------
  if (opt->bams_from_10x) {   // process_10x_reads: the pair's HP tag decides; mates that disagree (or lack it) carry no information
    for (Locus* L : loci)
      for (size_t r = 0; r < L->reads.size(); r++) {
        int64_t hap = L->reads[r]->has_hp ? L->reads[r]->hp : -1;
        if (L->mates[r]) {
          const int64_t mate_hap = L->mates[r]->has_hp ? L->mates[r]->hp : -1;
          if (mate_hap != hap) hap = -1;
        }
        if (hap == -1) continue;
        if (hap != 1 && hap != 2) { g_driver_error = "HP tag of " + L->reads[r]->name + " is neither 1 nor 2"; return HIPSTR_ERR_BAD_ARG; }
        L->log_p1[r] = hap == 1 ? -0.01 : -1000.0;     // FROM_HAP_LL / OTHER_HAP_LL (snp_bam_processor.h:17-18)
        L->log_p2[r] = hap == 2 ? -0.01 : -1000.0;
        R->counters[2]++;
      }
  } else if (snp_vcf && !loci.empty()) {
    const int32_t n_vcf = hipstr_snp_vcf_num_samples(snp_vcf);
    std::map<std::string, int> vcf_index;
    {
      const std::string names = hipstr_snp_vcf_samples(snp_vcf);
      size_t at = 0;
      for (int s = 0; s < n_vcf; s++) { const size_t eol = names.find('\n', at); vcf_index[names.substr(at, eol - at)] = s; at = eol + 1; }
    }
    std::vector<int32_t> entry_aln_off(1, 0), entry_set, aln_pos, aln_end, aln_seq_off(1, 0), aln_cigar_off(1, 0), cigar_len, set_off(1, 0);
    std::vector<uint32_t> snp_pos;
    std::string bases, quals, cigar_type, base1, base2;
    std::vector<std::pair<Locus*, size_t> > entry_owner;
    // the SNP sets of every region (binary searches on the loaded VCF; the handle's result buffers are not shared)
    struct Sets { bool found = false; std::vector<int32_t> off; std::vector<uint32_t> pos; std::string b1, b2; };
    std::vector<Sets> sets(loci.size());
    for (size_t l = 0; l < loci.size(); l++) {
      const int i = loci[l]->region;
      const int32_t start = region_start[i], stop = region_stop[i], dist = opt->filter.max_mate_dist;
      int32_t found = 0;
      const int32_t* off; const uint32_t* pos; const char* b1; const char* b2;
      hipstr_status_t st = hipstr_snp_vcf_region_sets(snp_vcf, region_chrom[i], start > dist ? start - dist : 1, stop + dist, 1, &start, &stop,
                                                      opt->skip_padding, &found, &off, &pos, &b1, &b2);
      if (st != HIPSTR_OK) return st;
      if (!found) continue;    // chromosome not in the VCF: no SNP information for this region
      sets[l].found = true;
      sets[l].off.assign(off, off + n_vcf + 1);
      sets[l].pos.assign(pos, pos + off[n_vcf]);
      sets[l].b1.assign(b1, (size_t)off[n_vcf]);
      sets[l].b2.assign(b2, (size_t)off[n_vcf]);
    }
    // the reads (+ mates) of every region, packed per region on the host threads and then laid end to end
    std::vector<AlnPiece> pieces(loci.size());
    std::vector<std::vector<int32_t> > piece_sets(loci.size());
    hipstr::parallel_for(loci.size(), [&](size_t l) {
      if (!sets[l].found) return;
      const Locus* L = loci[l];
      std::vector<int32_t> sample_set(L->kept.rg_names.size());
      for (size_t g = 0; g < sample_set.size(); g++) {
        auto vi = vcf_index.find(L->kept.rg_names[g]);
        sample_set[g] = vi == vcf_index.end() ? -1 : vi->second;
      }
      for (size_t r = 0; r < L->reads.size(); r++) {
        pieces[l].add(*L->reads[r]);
        if (L->mates[r]) pieces[l].add(*L->mates[r]);
        pieces[l].group_n.push_back(L->mates[r] ? 2 : 1);
        piece_sets[l].push_back(sample_set[L->label[r]]);
      }
    });
    for (size_t l = 0; l < loci.size(); l++) {
      if (!sets[l].found) continue;
      const int32_t set_base = (int32_t)set_off.size() - 1, snp_base = (int32_t)snp_pos.size();
      for (int sm = 0; sm < n_vcf; sm++) set_off.push_back(snp_base + sets[l].off[sm + 1]);
      snp_pos.insert(snp_pos.end(), sets[l].pos.begin(), sets[l].pos.end());
      base1 += sets[l].b1;
      base2 += sets[l].b2;
      const AlnPiece& P = pieces[l];
      for (int32_t n : P.group_n) entry_aln_off.push_back(entry_aln_off.back() + n);
      for (int32_t v : piece_sets[l]) entry_set.push_back(v < 0 ? -1 : set_base + v);
      P.append_to(aln_pos, aln_end, aln_seq_off, bases, quals, aln_cigar_off, cigar_type, cigar_len);
      for (size_t r = 0; r < loci[l]->reads.size(); r++) entry_owner.emplace_back(loci[l], r);
    }
    if (!entry_set.empty()) {
      cigar_len.push_back(0);
      if (snp_pos.empty()) snp_pos.push_back(0);
      hipstr_snp_phasing_t b;
      std::memset(&b, 0, sizeof(b));
      b.n_entries = (int32_t)entry_set.size(); b.entry_aln_off = entry_aln_off.data(); b.entry_snp_set = entry_set.data();
      b.n_alns = (int32_t)aln_pos.size(); b.aln_pos = aln_pos.data(); b.aln_end = aln_end.data(); b.aln_seq_off = aln_seq_off.data();
      b.bases = bases.c_str(); b.quals = quals.c_str(); b.aln_cigar_off = aln_cigar_off.data(); b.cigar_type = cigar_type.c_str();
      b.cigar_len = cigar_len.data(); b.n_sets = (int32_t)set_off.size() - 1; b.set_off = set_off.data(); b.snp_pos = snp_pos.data();
      b.snp_base1 = base1.c_str(); b.snp_base2 = base2.c_str();
      std::vector<double> p1(entry_set.size()), p2(entry_set.size());
      std::vector<int32_t> counts(entry_set.size() * 4);
      R->seconds[6] = now_s() - t;           // SNP sets + packing the reads of the window
      hipstr_status_t st = hipstr_snp_phasing_batch_host(ctx, &b, p1.data(), p2.data(), counts.data());
      R->seconds[7] = now_s() - t - R->seconds[6];   // the K7 call (uploads, launch, downloads)
      if (st != HIPSTR_OK) { g_driver_error = hipstr_last_error(ctx); return st; }
      for (size_t e = 0; e < entry_owner.size(); e++) {
        entry_owner[e].first->log_p1[entry_owner[e].second] = p1[e];
        entry_owner[e].first->log_p2[entry_owner[e].second] = p2[e];
        R->counters[2] += p1[e] != p2[e];
      }
    }
  }
  R->seconds[1] = now_s() - t; t = now_s();

  // ---- stutter models: the default, or learn_stutter_model for all loci in ONE K4 call --------------------------------
  if (opt->use_def_stutter_model) {
    for (Locus* L : loci) { std::memcpy(L->stutter, opt->def_stutter_model, sizeof(L->stutter)); L->has_model = true; }
  } else if (!loci.empty()) {
    std::vector<int32_t> lro(1, 0), lso(1, 0), num_bps, label, motif, ref_allele, informative;
    std::vector<double> p1, p2;
    std::vector<uint8_t> haploid;
    for (Locus* L : loci) {
      const int i = L->region;
      const size_t S = L->kept.rg_names.size();
      std::vector<std::vector<size_t> > by_sample(S);   // read indices with a usable length, per sample
      std::vector<int32_t> bp_of(L->reads.size());
      int inf_reads = 0;
      size_t r = 0;
      for (size_t s = 0; s < S && inf_reads <= 10000; s++)     // MAX_INF_READS is tested between samples
        for (; r < L->reads.size() && L->label[r] == (int32_t)s; r++) {
          const BamRecord& a = *L->reads[r];
          std::string types;
          std::vector<int32_t> lens;
          for (const auto& op : a.cigar) { types += op.first; lens.push_back(op.second); }
          int32_t bp = 0;
          if (!hipstr_extract_cigar(types.c_str(), lens.data(), (int32_t)lens.size(), a.pos, region_start[i] - region_period[i],
                                    region_stop[i] + region_period[i], &bp))
            continue;
          if (bp < -(region_stop[i] - region_start[i] + 1)) continue;
          inf_reads++;
          bp_of[r] = bp;
          by_sample[s].push_back(r);
        }
      informative.push_back(inf_reads);
      for (size_t s = 0; s < S; s++)
        for (size_t idx : by_sample[s]) { num_bps.push_back(bp_of[idx]); label.push_back((int32_t)s); p1.push_back(L->log_p1[idx]); p2.push_back(L->log_p2[idx]); }
      lro.push_back((int32_t)num_bps.size());
      lso.push_back(lso.back() + (int32_t)S);
      motif.push_back(region_period[i]);
      ref_allele.push_back(0);
      haploid.push_back(is_haploid(region_chrom[i]) ? 1 : 0);
    }
    hipstr_em_batch_t em;
    std::memset(&em, 0, sizeof(em));
    const int32_t zero = 0; const double zd = 0;
    em.n_loci = (int32_t)loci.size(); em.locus_read_off = lro.data(); em.locus_sample_off = lso.data();
    em.num_bps = num_bps.empty() ? &zero : num_bps.data(); em.sample_label = label.empty() ? &zero : label.data();
    em.log_p1 = p1.empty() ? &zd : p1.data(); em.log_p2 = p2.empty() ? &zd : p2.data();
    em.motif_len = motif.data(); em.ref_allele = ref_allele.data(); em.haploid = haploid.data();
    std::vector<double> params(6 * loci.size()), ll(loci.size());
    std::vector<uint8_t> converged(loci.size());
    std::vector<int32_t> iters(loci.size());
    hipstr_status_t st = hipstr_em_train_host(ctx, &em, opt->max_em_iter, opt->abs_ll_converge, opt->frac_ll_converge, params.data(), converged.data(),
                                              iters.data(), ll.data());
    if (st != HIPSTR_OK) { g_driver_error = hipstr_last_error(ctx); return st; }
    for (size_t l = 0; l < loci.size(); l++) {
      if (informative[l] < opt->min_total_reads) R->regions[loci[l]->region].status = TOO_FEW_READS;
      else if (!converged[l]) R->regions[loci[l]->region].status = NO_STUTTER_MODEL;
      else { std::memcpy(loci[l]->stutter, &params[6 * l], sizeof(loci[l]->stutter)); loci[l]->has_model = true; }
    }
    loci.erase(std::remove_if(loci.begin(), loci.end(), [](Locus* L) { return !L->has_model; }), loci.end());
  }
  R->seconds[2] = now_s() - t; t = now_s();
  if (loci.empty()) { *out = R.release(); return HIPSTR_OK; }

  // ---- left alignment of every read of the window (K6), then the lockstep genotyper and the records -----------------
  const int32_t n = (int32_t)loci.size();
  std::vector<int32_t> lro(1, 0), lso(1, 0), seq_off(1, 0), read_start, read_stop, cigar_off(1, 0), cigar_len, label, name_id, starts, stops,
      periods, trim_start, trim_stop;
  std::string bases, quals, cigar_type;
  std::vector<double> p1, p2, stutter;
  std::vector<uint8_t> haploid, rev, use;
  std::vector<const char*> seq_ptr, chrom_ptr, name_ptr, sample_ptr, out_sample_ptr;
  struct RawPiece { AlnPiece alns; std::vector<int32_t> name_id; std::vector<uint8_t> rev, use; };
  std::vector<RawPiece> raw_pieces(loci.size());
  hipstr::parallel_for(loci.size(), [&](size_t l) {
    const Locus* L = loci[l];
    RawPiece& P = raw_pieces[l];
    std::map<std::string, int32_t> ids;      // equal names on adjacent reads = mates that both span the STR
    for (size_t r = 0; r < L->reads.size(); r++) {
      const BamRecord& a = *L->reads[r];
      P.alns.add(a);
      P.name_id.push_back(ids.insert(std::make_pair(a.name, (int32_t)ids.size())).first->second);
      P.rev.push_back(a.reverse() ? 1 : 0);
      P.use.push_back(!a.passes.empty() && a.passes[0] == '1' ? 1 : 0);
    }
  });
  for (size_t l = 0; l < loci.size(); l++) {
    const Locus* L = loci[l];
    const int i = L->region;
    const RawPiece& P = raw_pieces[l];
    P.alns.append_to(read_start, read_stop, seq_off, bases, quals, cigar_off, cigar_type, cigar_len);
    label.insert(label.end(), L->label.begin(), L->label.end());
    name_id.insert(name_id.end(), P.name_id.begin(), P.name_id.end());
    p1.insert(p1.end(), L->log_p1.begin(), L->log_p1.end());
    p2.insert(p2.end(), L->log_p2.begin(), L->log_p2.end());
    rev.insert(rev.end(), P.rev.begin(), P.rev.end());
    use.insert(use.end(), P.use.begin(), P.use.end());
    lro.push_back((int32_t)read_start.size());
    lso.push_back(lso.back() + (int32_t)L->kept.rg_names.size());
    haploid.push_back(is_haploid(region_chrom[i]) ? 1 : 0);
    starts.push_back(region_start[i]); stops.push_back(region_stop[i]); periods.push_back(region_period[i]);
    trim_start.push_back(region_start[i] > 40 ? region_start[i] - 40 : 1);
    trim_stop.push_back(region_stop[i] + 40);
    stutter.insert(stutter.end(), L->stutter, L->stutter + 6);
    seq_ptr.push_back(seqs[L->chrom].c_str());
    chrom_ptr.push_back(region_chrom[i]);
    name_ptr.push_back(region_name && region_name[i] ? region_name[i] : "");
    for (const std::string& s : L->kept.rg_names) sample_ptr.push_back(s.c_str());
  }
  cigar_len.push_back(0);
  hipstr_locus_reads_t raw;
  std::memset(&raw, 0, sizeof(raw));
  raw.locus_read_off = lro.data(); raw.locus_sample_off = lso.data(); raw.read_seq_off = seq_off.data(); raw.bases = bases.c_str();
  raw.quals = quals.c_str(); raw.read_start = read_start.data(); raw.cigar_off = cigar_off.data(); raw.cigar_type = cigar_type.c_str();
  raw.cigar_len = cigar_len.data(); raw.sample_label = label.data(); raw.name_id = name_id.data(); raw.log_p1 = p1.data(); raw.log_p2 = p2.data();
  raw.haploid = haploid.data(); raw.rev_strand = rev.data(); raw.read_stop = read_stop.data(); raw.use_for_haps = use.data();
  hipstr_left_aligned_t* aligned = nullptr;
  hipstr_status_t st = hipstr_left_align_reads_host(ctx, n, &raw, seq_ptr.data(), trim_start.data(), trim_stop.data(), &aligned);
  if (st != HIPSTR_OK) { g_driver_error = std::string("left alignment: ") + hipstr_last_error(ctx); return st; }
  int64_t failed = 0, nw = 0;
  hipstr_left_aligned_counts(aligned, &failed, &nw);
  R->counters[3] = failed;
  R->seconds[3] = now_s() - t; t = now_s();

  hipstr_genotyper_t* g = nullptr;
  if (opt->ref_vcf) {   // the alleles of the reference panel's record for every region (read_vcf_alleles)
    std::vector<int32_t> allele_pos(n, -1), allele_off(1, 0);
    std::vector<std::string> allele_text;
    for (int32_t l = 0; l < n; l++) {
      int32_t pos = -1, count = 0;
      const char* text = nullptr;
      if (hipstr_str_vcf_alleles(opt->ref_vcf, chrom_ptr[l], starts[l], stops[l], &pos, &count, &text) == 1) {
        allele_pos[l] = pos;
        const std::string all(text);
        for (size_t at = 0; at < all.size();) { const size_t eol = all.find('\n', at); allele_text.push_back(all.substr(at, eol - at)); at = eol + 1; }
      }
      allele_off.push_back((int32_t)allele_text.size());
    }
    std::vector<const char*> allele_ptr;
    for (const std::string& a : allele_text) allele_ptr.push_back(a.c_str());
    allele_ptr.push_back("");
    st = hipstr_genotyper_create_with_ref_alleles(ctx, n, starts.data(), stops.data(), periods.data(), seq_ptr.data(), stutter.data(),
                                                  hipstr_left_aligned_reads(aligned), allele_pos.data(), allele_off.data(), allele_ptr.data(), &g);
  } else
    st = hipstr_genotyper_create_from_reads(ctx, n, starts.data(), stops.data(), periods.data(), seq_ptr.data(), stutter.data(),
                                            hipstr_left_aligned_reads(aligned), &g);
  if (st != HIPSTR_OK) { hipstr_left_aligned_free(aligned); g_driver_error = "constructing the genotypers failed"; return st; }
  std::vector<uint8_t> ok(n, 0);
  st = hipstr_genotyper_genotype(g, opt->max_total_haplotypes, opt->max_flank_haplotypes, opt->min_flank_freq, 1, ok.data());
  if (st == HIPSTR_OK && opt->recalc_stutter_model)
    st = hipstr_genotyper_recompute_stutter_models(g, opt->max_total_haplotypes, opt->max_flank_haplotypes, opt->min_flank_freq, opt->max_em_iter,
                                                   opt->abs_ll_converge, opt->frac_ll_converge, ok.data());
  R->seconds[4] = now_s() - t; t = now_s();
  if (st == HIPSTR_OK) {
    for (const std::string& s : R->samples) out_sample_ptr.push_back(s.c_str());
    hipstr_vcf_loci_t vl;
    std::memset(&vl, 0, sizeof(vl));
    vl.chrom = chrom_ptr.data(); vl.name = name_ptr.data(); vl.region_start = starts.data(); vl.region_stop = stops.data(); vl.period = periods.data();
    vl.chrom_seq = seq_ptr.data(); vl.locus_sample_names = sample_ptr.data(); vl.n_out_samples = (int32_t)out_sample_ptr.size();
    vl.out_sample_names = out_sample_ptr.data();
    st = hipstr_genotyper_write_vcf(g, &vl, vcf_opt);
  }
  if (st != HIPSTR_OK) {
    g_driver_error = hipstr_genotyper_last_error(g);
    hipstr_genotyper_destroy(g);
    hipstr_left_aligned_free(aligned);
    return st;
  }
  std::vector<char> buf(1 << 20);
  for (int32_t l = 0; l < n; l++) {
    hipstr_region_results::Region& res = R->regions[loci[l]->region];
    int32_t pos = 0;
    int32_t len = hipstr_genotyper_locus_record(g, l, &pos, buf.data(), (int32_t)buf.size());
    if (len < 0) { buf.resize((size_t)-len + 16); len = hipstr_genotyper_locus_record(g, l, &pos, buf.data(), (int32_t)buf.size()); }
    if (ok[l] && len > 0) { res.status = GENOTYPED; res.pos = pos; res.text.assign(buf.data(), (size_t)len); }
    else res.status = GENOTYPING_FAILED;
  }
  R->seconds[5] = now_s() - t;
  hipstr_genotyper_timing(g, R->genotyper_seconds);
  {
    int32_t rounds = 0;
    hipstr_genotyper_stats(g, &R->genotyper_stats[0], &R->genotyper_stats[1], &rounds);
    R->genotyper_stats[2] = rounds;
  }
  hipstr_genotyper_destroy(g);
  hipstr_left_aligned_free(aligned);
  *out = R.release();
  return HIPSTR_OK;
}

int32_t hipstr_region_results_count(const hipstr_region_results_t* r) { return r ? (int32_t)r->regions.size() : -1; }
int32_t hipstr_region_results_status(const hipstr_region_results_t* r, int32_t region, int32_t* pos, int32_t* n_reads) {
  if (!r || region < 0 || region >= (int32_t)r->regions.size()) return -1;
  if (pos) *pos = r->regions[region].pos;
  if (n_reads) *n_reads = r->regions[region].n_reads;
  return r->regions[region].status;
}
const char* hipstr_region_results_record(const hipstr_region_results_t* r, int32_t region) {
  return r && region >= 0 && region < (int32_t)r->regions.size() ? r->regions[region].text.c_str() : nullptr;
}
const char* hipstr_region_results_samples(const hipstr_region_results_t* r) { return r ? r->sample_text.c_str() : nullptr; }
void hipstr_region_results_timing(const hipstr_region_results_t* r, double* seconds6 /* [8] */, int64_t* counters4) {
  if (!r) return;
  if (seconds6) std::memcpy(seconds6, r->seconds, 8 * sizeof(double));
  if (counters4) std::memcpy(counters4, r->counters, 4 * sizeof(int64_t));
}
void hipstr_region_results_genotyper_timing(const hipstr_region_results_t* r, double* seconds9, int64_t* stats3) {
  if (!r) return;
  if (seconds9) std::memcpy(seconds9, r->genotyper_seconds, sizeof(r->genotyper_seconds));
  if (stats3) std::memcpy(stats3, r->genotyper_stats, sizeof(r->genotyper_stats));
}
void hipstr_region_results_free(hipstr_region_results_t* r) { delete r; }

}  // extern "C"
