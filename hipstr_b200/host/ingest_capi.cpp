/*
 * ingest_capi.cpp -- C-ABI of the stages upstream of the genotyper (include/hipstr_b200.h, "section 8(f) row 4"):
 * indexed BAM access (bam_reader.h) and read filtering / pairing / PCR-duplicate removal (read_filter.h).
 */
#include <cstring>
#include <memory>
#include <sstream>

#include "../../include/hipstr_b200.h"
#include "bam_reader.h"
#include "read_filter.h"
#include "ingest_handles.h"


namespace {

thread_local std::string g_error;

int64_t give_text(const std::string& s, int64_t cap, char* out) {
  if (!out || (int64_t)s.size() + 1 > cap) return -(int64_t)s.size() - 1;
  std::memcpy(out, s.c_str(), s.size() + 1);
  return (int64_t)s.size();
}

std::string name_of(int32_t id, const std::vector<std::string>& names) { return id >= 0 && id < (int32_t)names.size() ? names[id] : std::string("*"); }

void fill_record(BamRecord& a, int32_t flag, int32_t pos, int32_t end_pos, const char* bases, const char* quals, int32_t n_cigar,
                 const char* cigar_type, const int32_t* cigar_len) {
  a.flag = (uint16_t)flag;
  a.pos = pos;
  a.end_pos = end_pos;
  a.bases = bases;
  a.quals = quals;
  for (int i = 0; i < n_cigar; i++) a.cigar.emplace_back(cigar_type[i], cigar_len[i]);
}

}  // namespace

extern "C" {

const char* hipstr_ingest_last_error(void) { return g_error.c_str(); }

hipstr_status_t hipstr_bam_reader_open(int32_t n_files, const char* const* paths, hipstr_bam_reader_t** out) {
  if (n_files < 1 || !paths || !out) return HIPSTR_ERR_BAD_ARG;
  std::unique_ptr<hipstr_bam_reader> r(new hipstr_bam_reader());
  for (int i = 0; i < n_files; i++) {
    std::unique_ptr<hipstr::BamFile> f(new hipstr::BamFile());
    if (!f->open(paths[i])) { g_error = f->error(); return HIPSTR_ERR_BAD_ARG; }
    if (i > 0) {   // compare_bam_headers (bam_io.cpp:364-381)
      if (f->ref_names() != r->files[0]->ref_names() || f->ref_lengths() != r->files[0]->ref_lengths()) {
        g_error = std::string("BAM header mismatch issue. BAM headers for files ") + paths[0] + " and " + paths[i] + " must have the same reference sequences";
        return HIPSTR_ERR_BAD_ARG;
      }
    }
    r->files.push_back(std::move(f));
    r->paths.push_back(paths[i]);
  }
  *out = r.release();
  return HIPSTR_OK;
}

void hipstr_bam_reader_close(hipstr_bam_reader_t* r) { delete r; }

int64_t hipstr_bam_reader_read_groups(const hipstr_bam_reader_t* r, int64_t cap, char* out_text) {
  if (!r) return -1;
  std::ostringstream out;
  for (size_t f = 0; f < r->files.size(); f++)
    for (const hipstr::BamReadGroup& g : r->files[f]->read_groups())
      out << r->paths[f] << '\t' << g.id << '\t' << (g.has_sample ? g.sample : std::string("-")) << '\t' << (g.has_library ? g.library : std::string("-")) << '\n';
  return give_text(out.str(), cap, out_text);
}

hipstr_status_t hipstr_bam_reader_fetch(hipstr_bam_reader_t* r, const char* chrom, int32_t start, int32_t end, hipstr_bam_records_t** out) {
  if (!r || !chrom || !out) return HIPSTR_ERR_BAD_ARG;
  std::unique_ptr<hipstr_bam_records> recs(new hipstr_bam_records());
  recs->ref_names = r->files[0]->ref_names();
  recs->file_names = r->paths;
  for (size_t f = 0; f < r->files.size(); f++)
    if (!r->files[f]->fetch(chrom, start, end, (int32_t)f, recs->records)) { g_error = r->files[f]->error(); return HIPSTR_ERR_BAD_ARG; }
  *out = recs.release();
  return HIPSTR_OK;
}

int32_t hipstr_bam_records_count(const hipstr_bam_records_t* recs) { return recs ? (int32_t)recs->records.size() : -1; }
void hipstr_bam_records_free(hipstr_bam_records_t* recs) { delete recs; }

int64_t hipstr_bam_records_text(const hipstr_bam_records_t* recs, int64_t cap, char* out_text) {
  if (!recs) return -1;
  std::ostringstream out;
  for (const BamRecord& a : recs->records) {
    out << a.name << '\t' << a.flag << '\t' << name_of(a.ref_id, recs->ref_names) << '\t' << a.pos << '\t' << a.end_pos << '\t' << (int)a.mapq << '\t'
        << (a.cigar.empty() ? std::string("*") : hipstr::cigar_string(a.cigar)) << '\t' << name_of(a.mate_ref_id, recs->ref_names) << '\t' << a.mate_pos
        << '\t' << a.bases << '\t' << a.quals << '\t' << (a.has_rg ? a.rg : std::string("-")) << '\t' << (a.has_xa ? a.xa : std::string("-")) << '\t'
        << (a.has_sa ? a.sa : std::string("-")) << '\t';
    if (a.has_as) out << a.as; else out << '-';
    out << '\t';
    if (a.has_xs) out << a.xs; else out << '-';
    out << '\t' << recs->file_names[a.file] << '\n';
  }
  return give_text(out.str(), cap, out_text);
}

void hipstr_filter_default_options(hipstr_filter_options_t* o) {
  if (!o) return;
  const hipstr::FilterOptions d;
  o->max_mate_dist = d.max_mate_dist;
  o->min_bp_before_indel = d.min_bp_before_indel;
  o->min_flank = d.min_flank;
  o->min_read_end_match = d.min_read_end_match;
  o->maximal_end_match_window = d.maximal_end_match_window;
  o->require_paired_reads = d.require_paired_reads;
  o->min_sum_qual_log_prob = d.min_sum_qual_log_prob;
  o->max_total_reads = d.max_total_reads;
  o->base_qual_trim = d.base_qual_trim;
  o->remove_pcr_dups = d.remove_pcr_dups;
  o->trim_adapters = d.trim_adapters;
}

hipstr_status_t hipstr_filter_reads(const hipstr_bam_records_t* recs, const char* chrom_seq, int32_t n_regions, const int32_t* region_start,
                                    const int32_t* region_stop, const hipstr_filter_options_t* opt, int32_t n_rg, const char* const* rg_keys,
                                    const char* const* rg_samples, const char* const* rg_libraries, hipstr_filtered_reads_t** out) {
  if (!recs || !chrom_seq || n_regions < 1 || !region_start || !region_stop || !opt || n_rg < 0 || !out) return HIPSTR_ERR_BAD_ARG;
  if (n_rg > 0 && (!rg_keys || !rg_samples || !rg_libraries)) return HIPSTR_ERR_BAD_ARG;
  hipstr::ReadFilter filter;
  filter.options.max_mate_dist = opt->max_mate_dist;
  filter.options.min_bp_before_indel = opt->min_bp_before_indel;
  filter.options.min_flank = opt->min_flank;
  filter.options.min_read_end_match = opt->min_read_end_match;
  filter.options.maximal_end_match_window = opt->maximal_end_match_window;
  filter.options.require_paired_reads = opt->require_paired_reads;
  filter.options.min_sum_qual_log_prob = opt->min_sum_qual_log_prob;
  filter.options.max_total_reads = opt->max_total_reads;
  filter.options.base_qual_trim = (char)opt->base_qual_trim;
  filter.options.remove_pcr_dups = opt->remove_pcr_dups != 0;
  filter.options.trim_adapters = opt->trim_adapters != 0;
  std::map<std::string, std::string> rg_to_sample, rg_to_library;
  for (int i = 0; i < n_rg; i++) { rg_to_sample[rg_keys[i]] = rg_samples[i]; rg_to_library[rg_keys[i]] = rg_libraries[i]; }
  std::vector<std::pair<int32_t, int32_t> > regions;
  for (int i = 0; i < n_regions; i++) {
    if (region_stop[i] <= region_start[i]) return HIPSTR_ERR_BAD_ARG;
    regions.emplace_back(region_start[i], region_stop[i]);
  }
  std::unique_ptr<hipstr_filtered_reads> h(new hipstr_filtered_reads());
  try {
    // a view, not a std::string: the chromosome is not copied per call
    filter.run(recs->records, recs->ref_names, recs->file_names, std::string_view(chrom_seq, std::strlen(chrom_seq)), regions, rg_to_sample,
               h->reads);
    if (filter.options.remove_pcr_dups) hipstr::ReadFilter::remove_pcr_duplicates(rg_to_library, recs->file_names, h->reads);
  } catch (const std::exception& e) {
    g_error = e.what();
    return HIPSTR_ERR_BAD_ARG;
  }
  h->adapter_stats = filter.adapter_trimmer.stats_message();
  *out = h.release();
  return HIPSTR_OK;
}

void hipstr_filtered_reads_free(hipstr_filtered_reads_t* h) { delete h; }

void hipstr_filtered_reads_counts(const hipstr_filtered_reads_t* h, int32_t* c) {
  if (!h || !c) return;
  const hipstr::FilterCounts& n = h->reads.counts;
  int32_t passed = 0;
  for (size_t g = 0; g < h->reads.paired.size(); g++) passed += (int32_t)(h->reads.paired[g].size() + h->reads.unpaired[g].size());
  c[0] = n.read_count; c[1] = n.hard_clip; c[2] = n.read_has_n; c[3] = n.low_qual_score; c[4] = n.unique_mapping;
  c[5] = n.unpaired_filtered; c[6] = n.pcr_duplicates; c[7] = n.too_many_reads ? 1 : 0; c[8] = passed;
}

int64_t hipstr_filtered_reads_text(const hipstr_filtered_reads_t* h, int64_t cap, char* out_text) {
  if (!h) return -1;
  std::ostringstream out;
  auto dump = [&out](const char* kind, const BamRecord& a) {
    out << kind << '\t' << a.name << '\t' << a.flag << '\t' << a.pos << '\t' << a.end_pos << '\t'
        << (a.cigar.empty() ? std::string("*") : hipstr::cigar_string(a.cigar)) << '\t' << a.bases << '\t' << a.quals << '\t'
        << (a.passes.empty() ? std::string("-") : a.passes) << '\n';
  };
  const hipstr::FilteredReads& r = h->reads;
  for (size_t g = 0; g < r.rg_names.size(); g++) {
    out << "G\t" << r.rg_names[g] << '\n';
    for (size_t i = 0; i < r.paired[g].size(); i++) { dump("P", r.paired[g][i]); dump("M", r.mates[g][i]); }
    for (size_t i = 0; i < r.unpaired[g].size(); i++) dump("U", r.unpaired[g][i]);
  }
  out << "T\t" << (r.counts.too_many_reads ? 1 : 0) << '\n';
  return give_text(out.str(), cap, out_text);
}

hipstr_status_t hipstr_filtered_reads_view(hipstr_filtered_reads_t* h, hipstr_filtered_view_t* v) {
  if (!h || !v) return HIPSTR_ERR_BAD_ARG;
  const hipstr::FilteredReads& r = h->reads;
  h->sample_entry_off.assign(1, 0); h->entry_aln_off.assign(1, 0); h->aln_seq_off.assign(1, 0); h->aln_cigar_off.assign(1, 0);
  h->entry_snp_set.clear(); h->aln_pos.clear(); h->aln_end.clear(); h->cigar_len.clear(); h->aln_flag.clear();
  h->bases.clear(); h->quals.clear(); h->cigar_type.clear(); h->passes.clear(); h->sample_names.clear();
  h->entry_names.clear(); h->entry_name_off.assign(1, 0);
  auto add = [&](const BamRecord& a) {
    h->aln_pos.push_back(a.pos);
    h->aln_end.push_back(a.end_pos);
    h->aln_flag.push_back(a.flag);
    h->bases += a.bases;
    h->quals += a.quals;
    h->aln_seq_off.push_back((int32_t)h->bases.size());
    for (const auto& op : a.cigar) { h->cigar_type += op.first; h->cigar_len.push_back(op.second); }
    h->aln_cigar_off.push_back((int32_t)h->cigar_type.size());
  };
  for (size_t g = 0; g < r.rg_names.size(); g++) {
    h->sample_names.push_back(r.rg_names[g].c_str());
    // the order of SNPBamProcessor::process_reads: paired STR reads (each with its mate), then the unpaired ones
    for (size_t i = 0; i < r.paired[g].size(); i++) {
      add(r.paired[g][i]);
      add(r.mates[g][i]);
      h->entry_aln_off.push_back((int32_t)h->aln_pos.size());
      h->entry_snp_set.push_back((int32_t)g);
      h->passes += r.paired[g][i].passes.empty() ? '0' : r.paired[g][i].passes[0];
      h->entry_names += r.paired[g][i].name;
      h->entry_name_off.push_back((int32_t)h->entry_names.size());
    }
    for (size_t i = 0; i < r.unpaired[g].size(); i++) {
      add(r.unpaired[g][i]);
      h->entry_aln_off.push_back((int32_t)h->aln_pos.size());
      h->entry_snp_set.push_back((int32_t)g);
      h->passes += r.unpaired[g][i].passes.empty() ? '0' : r.unpaired[g][i].passes[0];
      h->entry_names += r.unpaired[g][i].name;
      h->entry_name_off.push_back((int32_t)h->entry_names.size());
    }
    h->sample_entry_off.push_back((int32_t)h->entry_snp_set.size());
  }
  h->cigar_len.push_back(0);
  v->n_samples = (int32_t)r.rg_names.size();
  v->sample_names = h->sample_names.data();
  v->sample_entry_off = h->sample_entry_off.data();
  v->entry_passes = h->passes.c_str();
  v->aln_flag = h->aln_flag.data();
  v->entry_name_off = h->entry_name_off.data();
  v->entry_names = h->entry_names.c_str();
  hipstr_snp_phasing_t& b = v->reads;
  std::memset(&b, 0, sizeof(b));
  b.n_entries = (int32_t)h->entry_snp_set.size();
  b.entry_aln_off = h->entry_aln_off.data();
  b.entry_snp_set = h->entry_snp_set.data();
  b.n_alns = (int32_t)h->aln_pos.size();
  b.aln_pos = h->aln_pos.data();
  b.aln_end = h->aln_end.data();
  b.aln_seq_off = h->aln_seq_off.data();
  b.bases = h->bases.c_str();
  b.quals = h->quals.c_str();
  b.aln_cigar_off = h->aln_cigar_off.data();
  b.cigar_type = h->cigar_type.c_str();
  b.cigar_len = h->cigar_len.data();
  return HIPSTR_OK;
}

/* single-read steps, exported so that each can be checked on its own */
int32_t hipstr_trim_one(int32_t what, int32_t arg, int32_t arg2, int32_t flag, int32_t pos, int32_t end_pos, const char* bases, const char* quals,
                        int32_t n_cigar, const char* cigar_type, const int32_t* cigar_len, int32_t* out_pos, char* out_seq, char* out_qual,
                        int32_t* n_out_cigar, char* out_ctype, int32_t* out_clen) {
  if (!bases || !quals || (n_cigar > 0 && (!cigar_type || !cigar_len)) || !out_pos || !out_seq || !out_qual || !n_out_cigar || !out_ctype || !out_clen) return -2;
  BamRecord a;
  fill_record(a, flag, pos, end_pos, bases, quals, n_cigar, cigar_type, cigar_len);
  try {
    if (what == 0) hipstr::trim_low_quality_ends(a, (char)arg);
    else if (what == 1) { hipstr::AdapterTrimmer trimmer; trimmer.trim_adapters(a); }
    else hipstr::trim_num_bases(a, arg, arg2);
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
  out_pos[0] = a.pos; out_pos[1] = a.end_pos; out_pos[2] = a.length();
  std::strcpy(out_seq, a.bases.c_str());
  std::strcpy(out_qual, a.quals.c_str());
  *n_out_cigar = (int32_t)a.cigar.size();
  for (size_t i = 0; i < a.cigar.size(); i++) { out_ctype[i] = a.cigar[i].first; out_clen[i] = a.cigar[i].second; }
  return 0;
}

int32_t hipstr_alignment_filters(int32_t pos, int32_t end_pos, const char* bases, const char* quals, int32_t n_cigar, const char* cigar_type,
                                 const int32_t* cigar_len, const char* chrom_seq, int32_t window, int32_t* out, double* sum_qual) {
  if (!bases || !quals || (n_cigar > 0 && (!cigar_type || !cigar_len)) || !chrom_seq || !out || !sum_qual) return -2;
  BamRecord a;
  fill_record(a, 0, pos, end_pos, bases, quals, n_cigar, cigar_type, cigar_len);
  const std::string_view ref(chrom_seq, std::strlen(chrom_seq));
  try {
    out[0] = hipstr::filters::has_largest_end_matches(a, ref, 0, window, window) ? 1 : 0;
    const std::pair<int, int> m = hipstr::filters::num_end_matches(a, ref, 0);
    out[1] = m.first; out[2] = m.second;
    const std::pair<int, int> d = hipstr::filters::end_dist_to_indel(a);
    out[3] = d.first; out[4] = d.second;
    *sum_qual = hipstr::filters::sum_log_prob_correct(a.quals);
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1;
  }
  return 0;
}

}  // extern "C"
